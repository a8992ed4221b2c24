// Peer mailboxes: the NVLink exchange layer of the Z-sharded ops (one process per GPU).
//
// Every rank owns one MAILBOX in its own HBM (cudaMalloc'ed, exported with cudaIpc, mapped by
// every other rank). Ranks never pull: a rank WRITES into its peers' mailboxes over NVLink
// (plain stores + __threadfence_system, then a release store of a signal word) and reads only
// its own. All waits are polls of LOCAL words with a clock64() timeout, so a mismatched call
// sequence ends in B2V_ERR_NOCONV instead of a hung GPU.
//
// Mailbox layout (byte offsets from its base; `pc` = plane capacity in bytes = dy * ceil(dx/32)
// * 4, `rc` = record capacity = 4 * pc):
//   0     uint32 sig[64]            signal words, see PB_SIG_*
//   256   uint32 flags[2][16]       flood: "my shard gained bits" of every rank, by epoch parity
//   512   int64  counts[2][16][2]   marching cubes: (V, T) of every rank, by epoch parity
//   1024  uint32 cnt_tag[2][16]     epoch tags of the counts
//   4096  flood inbox from the LOWER neighbour: 2 planes of reached bits (its [last own, halo])
//   +2pc  flood inbox from the UPPER neighbour: 2 planes                 (its [halo, first own])
//   +4pc  marching-cubes inbox [2][rc]: the UPPER neighbour's plane-0 records, by epoch parity
#pragma once
#include "b2v_common.cuh"

constexpr int kPeerMaxWorld = 16;
enum { PB_SIG_FF_FROM_LO = 0, PB_SIG_FF_FROM_HI = 1, PB_SIG_MC_FROM_HI = 2 };
constexpr int64_t kPbFlags = 256, kPbCounts = 512, kPbCntTag = 1024, kPbData = 4096;

struct PeerBox {
  char* base;
  int64_t pc;   // plane capacity (bytes)
  __host__ __device__ uint32_t* sig(int i) const { return (uint32_t*)base + i; }
  __host__ __device__ uint32_t* flags(int parity) const { return (uint32_t*)(base + kPbFlags) + parity * kPeerMaxWorld; }
  __host__ __device__ long long* counts(int parity) const {
    return (long long*)(base + kPbCounts) + parity * kPeerMaxWorld * 2;
  }
  __host__ __device__ uint32_t* cnt_tag(int parity) const {
    return (uint32_t*)(base + kPbCntTag) + parity * kPeerMaxWorld;
  }
  __host__ __device__ uint32_t* ff_from_lo() const { return (uint32_t*)(base + kPbData); }
  __host__ __device__ uint32_t* ff_from_hi() const { return (uint32_t*)(base + kPbData + 2 * pc); }
  __host__ __device__ uint4* mc_from_hi(int parity) const { return (uint4*)(base + kPbData + 4 * pc + parity * 4 * pc); }
};

static inline int64_t peer_mailbox_bytes(int64_t pc) { return kPbData + 4 * pc + 8 * pc; }

struct PeerSet {           // passed by value to the exchange kernels
  int rank, world;
  uint32_t epoch;          // first epoch of this call (epochs are consumed in lockstep by all ranks)
  long long timeout;       // clock64() cycles a poll may take
  int64_t pc;
  char* box[kPeerMaxWorld];
  __host__ __device__ PeerBox of(int r) const { return PeerBox{box[r], pc}; }
  __host__ __device__ PeerBox mine() const { return of(rank); }
};

// ---- system-scope accessors ------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_s64(long long* p, long long v) {
  asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_relaxed_sys_s64(const long long* p) {
  long long v;
  asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Poll a LOCAL word until it equals `want` (acquire). Returns false on timeout.
__device__ __forceinline__ bool peer_wait_eq(const uint32_t* p, uint32_t want, long long timeout) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) != want) {
    if (clock64() - t0 > timeout) return false;
    __nanosleep(64);
  }
  return true;
}

// host side (peer.cu): fill a PeerSet from the host array of mailbox base pointers
int peer_make_set(int rank, int world, const void* const* mailboxes_host, int64_t plane_bytes, uint32_t epoch,
                  PeerSet* out);
