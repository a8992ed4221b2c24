// Marching cubes: classify -> scan -> emit, on a bit-packed inside volume.
// Replaces the per-piece vtkContourFilter step of invesalius/data/surface_process.py:71-201
// (geometry: converters.py:34-101). Canonical output order is defined in DESIGN.md and
// restated by the CPU checker; VTK itself is not available, so parity to VTK is unpinned.
//
//   1. bits   (HBM-bound, the only pass over the volume: 1 B/voxel uint8, 2 B/voxel int16):
//             inside(p) = S[p] >= iso packed 32 voxels per word along x.
//   2. count  (on bits, L2 resident): per word the crossing masks towards +x/+y/+z, the
//             number of owned vertices (popcounts) and of triangles (table lookups over
//             the active cells only).
//   3. scan   exclusive prefix sums of both counts (block sums -> one block -> apply).
//   4. emit   one thread per vertex and one thread per active cell, each finding its word by
//             searching the offsets; a vertex id is  voff[owner word] + popcounts below the
//             owner bit.
#include "b2v_common.cuh"
#define B2V_MC_QUAL __device__
#include "mc_tables.h"

namespace {

struct McGeom {
  int64_t nz, ny, nx;
  int wx;           // words per row
  int64_t nwords;
};

McGeom make_geom(int64_t nz, int64_t ny, int64_t nx) {
  McGeom g;
  g.nz = nz; g.ny = ny; g.nx = nx;
  g.wx = (int)ceil_div64(nx, 32);
  g.nwords = nz * ny * g.wx;
  return g;
}

constexpr int kScanBlock = 256;

struct McWs {
  uint32_t* bits;    // [nwords] inside bits
  uint4* info;       // [nwords] (cx, cy, cz, vertex offset)
  unsigned long long* toff;   // [nwords] (active-cell offset << 32) | triangle offset
  uint32_t* amask;   // [nwords] cells of the word that produce triangles
  uint32_t* bsum_v;  // [nblocks]
  unsigned long long* bsum_t;  // [nblocks] packed like toff
  unsigned long long* totals;  // [0] V, [1] (C << 32) | T
  int64_t nblocks;
  int64_t bytes;
};

McWs carve(void* base, const McGeom& g) {
  McWs w;
  auto align = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
  char* p = (char*)base;
  int64_t off = 0;
  w.nblocks = ceil_div64(g.nwords, kScanBlock);
  w.bits = (uint32_t*)(p + off); off += align(g.nwords * 4 + 4);
  w.info = (uint4*)(p + off); off += align(g.nwords * 16);
  w.toff = (unsigned long long*)(p + off); off += align(g.nwords * 8);
  w.amask = (uint32_t*)(p + off); off += align(g.nwords * 4);
  w.bsum_v = (uint32_t*)(p + off); off += align(w.nblocks * 4);
  w.bsum_t = (unsigned long long*)(p + off); off += align(w.nblocks * 8);
  w.totals = (unsigned long long*)(p + off); off += 256;
  w.bytes = off;
  return w;
}

// ---- 1. inside bits ---------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_mc_bits(const T* __restrict__ vol, McGeom g, int ithr,
                                                 uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < g.nwords; wi += nwarps) {
    int64_t row = wi / g.wx;
    int64_t x = (wi - row * g.wx) * 32 + lane;
    bool in = x < g.nx && (int)vol[row * g.nx + x] >= ithr;
    uint32_t b = __ballot_sync(0xffffffffu, in);
    if (lane == 0) bits[wi] = b;
  }
}

// uint8, nx % 16 == 0, aligned: one 128-bit load = 16 voxels per lane, 2 lanes per word.
// LINEAR: nx % 32 == 0, rows hold no padding groups: group gi is voxels [16 gi, 16 gi + 16)
// and half-word gi of the bit volume (no 64-bit division per group).
template <bool LINEAR>
__global__ void __launch_bounds__(256) k_mc_bits_u8_vec(const uint8_t* __restrict__ vol, McGeom g, int ithr,
                                                        uint32_t* __restrict__ bits) {
  const int gx = g.wx * 2;  // 16-voxel groups per row (padded)
  const int64_t ngroups = g.nz * g.ny * gx;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  const int lane = threadIdx.x & 31;
  const uint32_t thr = (uint32_t)(ithr > 255 ? 255 : (ithr < 0 ? 0 : ithr));
  const uint32_t t7 = (thr & 0x7fu) * 0x01010101u;
  const bool t_high = thr >= 128u;
  const bool none = ithr > 255;
  // four 128-bit loads in flight per thread; bytes are compared four at a time
  for (int64_t g0 = (int64_t)blockIdx.x * blockDim.x * 4; g0 < ngroups; g0 += stride) {
    uint4 v[4];
    int64_t row[4];
    int q[4];
    bool ok[4], in[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t gi = g0 + k * blockDim.x + threadIdx.x;
      ok[k] = gi < ngroups;
      if (LINEAR) {
        row[k] = 0; q[k] = 0;
        in[k] = ok[k];
      } else {
        row[k] = ok[k] ? gi / gx : 0;
        q[k] = ok[k] ? (int)(gi - row[k] * gx) : 0;
        in[k] = ok[k] && (int64_t)q[k] * 16 < g.nx;   // a padded group contributes zero bits
      }
      v[k] = make_uint4(0u, 0u, 0u, 0u);
      if (in[k]) v[k] = ld_stream((const uint4*)(vol + (LINEAR ? gi * 16 : row[k] * g.nx + (int64_t)q[k] * 16)));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t b = 0;
      if (in[k] && !none) {
        b = flags_to_nibble(ge_flags_u8x4(v[k].x, t7, t_high)) | (flags_to_nibble(ge_flags_u8x4(v[k].y, t7, t_high)) << 4) |
            (flags_to_nibble(ge_flags_u8x4(v[k].z, t7, t_high)) << 8) |
            (flags_to_nibble(ge_flags_u8x4(v[k].w, t7, t_high)) << 12);
      }
      uint32_t word = b << (16 * (lane & 1));
      word |= __shfl_xor_sync(0xffffffffu, word, 1);
      if ((lane & 1) == 0 && ok[k]) {
        const int64_t gi = g0 + k * blockDim.x + threadIdx.x;
        bits[LINEAR ? gi >> 1 : row[k] * g.wx + (q[k] >> 1)] = word;
      }
    }
  }
}

// int16, nx % 8 == 0, aligned: 8 voxels per lane, 4 lanes per word
__global__ void __launch_bounds__(256) k_mc_bits_i16_vec(const int16_t* __restrict__ vol, McGeom g, int ithr,
                                                         uint32_t* __restrict__ bits) {
  const int gx = g.wx * 4;
  const int64_t ngroups = g.nz * g.ny * gx;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (int64_t g0 = (int64_t)blockIdx.x * blockDim.x; g0 < ngroups; g0 += stride) {
    int64_t gi = g0 + threadIdx.x;
    uint32_t b = 0;
    int64_t row = 0;
    int q = 0;
    if (gi < ngroups) {
      row = gi / gx;
      q = (int)(gi - row * gx);
      int64_t x = (int64_t)q * 8;
      if (x < g.nx) {
        int4 v = ld_stream((const int4*)(vol + row * g.nx + x));
        int vv[8] = {(int16_t)(v.x & 0xffff), v.x >> 16, (int16_t)(v.y & 0xffff), v.y >> 16,
                     (int16_t)(v.z & 0xffff), v.z >> 16, (int16_t)(v.w & 0xffff), v.w >> 16};
#pragma unroll
        for (int k = 0; k < 8; ++k) b |= (uint32_t)(vv[k] >= ithr) << k;
      }
    }
    uint32_t word = b << (8 * (lane & 3));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    if ((lane & 3) == 0 && gi < ngroups) bits[row * g.wx + (q >> 2)] = word;
  }
}

// ---- 2. count -------------------------------------------------------------------------------
struct Rows {       // the four bit rows a cell row touches, word w and bit 0 of word w+1
  uint32_t i00, i01, i10, i11;  // [cz][cy]
  uint32_t n00, n01, n10, n11;  // next word in x (0 past the row end)
};

__device__ __forceinline__ Rows load_rows(const uint32_t* __restrict__ bits, const McGeom& g, int64_t z, int64_t y,
                                          int w) {
  Rows r;
  const bool hy = y + 1 < g.ny, hz = z + 1 < g.nz, hn = w + 1 < g.wx;
  const int64_t b00 = (z * g.ny + y) * g.wx + w;
  const int64_t b01 = b00 + g.wx, b10 = b00 + (int64_t)g.ny * g.wx, b11 = b10 + g.wx;
  r.i00 = __ldg(bits + b00);
  r.i01 = hy ? __ldg(bits + b01) : 0u;
  r.i10 = hz ? __ldg(bits + b10) : 0u;
  r.i11 = (hy && hz) ? __ldg(bits + b11) : 0u;
  r.n00 = hn ? __ldg(bits + b00 + 1) : 0u;
  r.n01 = (hn && hy) ? __ldg(bits + b01 + 1) : 0u;
  r.n10 = (hn && hz) ? __ldg(bits + b10 + 1) : 0u;
  r.n11 = (hn && hy && hz) ? __ldg(bits + b11 + 1) : 0u;
  return r;
}

// bit i of the result = bit i+1 of the 64-bit (hi:lo)
__device__ __forceinline__ uint32_t shift_in(uint32_t lo, uint32_t hi) { return (lo >> 1) | (hi << 31); }

// bits i with x = 32 w + i and x + 1 < nx
__device__ __forceinline__ uint32_t valid_x1(const McGeom& g, int w) {
  int64_t rem = g.nx - 1 - (int64_t)w * 32;  // number of valid bits
  return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

__device__ __forceinline__ int cell_case(const Rows& r, int i) {
  uint32_t a = __funnelshift_r(r.i00, r.n00, i) & 3u;
  uint32_t b = __funnelshift_r(r.i01, r.n01, i) & 3u;
  uint32_t c = __funnelshift_r(r.i10, r.n10, i) & 3u;
  uint32_t d = __funnelshift_r(r.i11, r.n11, i) & 3u;
  return (int)(a | (b << 2) | (c << 4) | (d << 6));
}

__global__ void __launch_bounds__(kScanBlock) k_mc_count(const uint32_t* __restrict__ bits, McGeom g, int skip_last,
                                                         uint4* __restrict__ info,
                                                         unsigned long long* __restrict__ toff,
                                                         uint32_t* __restrict__ amask, uint32_t* __restrict__ bsum_v,
                                                         unsigned long long* __restrict__ bsum_t) {
  __shared__ unsigned char s_ntri[256];
  __shared__ unsigned long long s_red[2][kScanBlock / 32];
  s_ntri[threadIdx.x] = B2V_MC_NTRI[threadIdx.x];
  __syncthreads();
  const int64_t wi = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  uint32_t nv = 0, nt = 0, nact = 0;
  if (wi < g.nwords) {
    int64_t row = wi / g.wx;
    int w = (int)(wi - row * g.wx);
    int64_t z = row / g.ny, y = row - z * g.ny;
    Rows r = load_rows(bits, g, z, y, w);
    const uint32_t vx = valid_x1(g, w);
    const bool hy = y + 1 < g.ny, hz = z + 1 < g.nz;
    uint32_t cx = (r.i00 ^ shift_in(r.i00, r.n00)) & vx;
    uint32_t cy = hy ? (r.i00 ^ r.i01) : 0u;
    uint32_t cz = hz ? (r.i00 ^ r.i10) : 0u;
    // a Z shard does not own the vertices of its last (shared) plane: the next shard does
    nv = (skip_last && z == g.nz - 1) ? 0 : __popc(cx) + __popc(cy) + __popc(cz);
    uint32_t act = 0;
    if (hy && hz) {
      uint32_t s00 = shift_in(r.i00, r.n00), s01 = shift_in(r.i01, r.n01), s10 = shift_in(r.i10, r.n10),
               s11 = shift_in(r.i11, r.n11);
      uint32_t any = r.i00 | r.i01 | r.i10 | r.i11 | s00 | s01 | s10 | s11;
      uint32_t all = r.i00 & r.i01 & r.i10 & r.i11 & s00 & s01 & s10 & s11;
      act = any & ~all & vx;
      nact = __popc(act);
      uint32_t m = act;
      while (m) {
        int i = __ffs(m) - 1;
        m &= m - 1;
        nt += s_ntri[cell_case(r, i)];
      }
    }
    info[wi] = make_uint4(cx, cy, cz, nv);
    toff[wi] = ((unsigned long long)nact << 32) | nt;
    amask[wi] = act;
  }
  // block sums (cells and triangles travel packed: no carry can cross, totals < 2^31)
  unsigned long long a = nv, b = ((unsigned long long)nact << 32) | nt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = a;
    s_red[1][threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long sa = 0, sb = 0;
#pragma unroll
    for (int k = 0; k < kScanBlock / 32; ++k) {
      sa += s_red[0][k];
      sb += s_red[1][k];
    }
    bsum_v[blockIdx.x] = (uint32_t)sa;
    bsum_t[blockIdx.x] = sb;
  }
}

// ---- 3. scan -------------------------------------------------------------------------------
// one block: exclusive scan of the block sums (in place, as 64-bit running totals truncated
// to 32 bits on store; the host rejects totals >= 2^32), totals[0..1] = V, T
__global__ void __launch_bounds__(1024) k_mc_scan_bsums(uint32_t* bsum_v, unsigned long long* bsum_t, int64_t nblocks,
                                                        unsigned long long* totals) {
  __shared__ unsigned long long s_w[2][32];
  __shared__ unsigned long long s_carry[2];
  if (threadIdx.x == 0) s_carry[0] = s_carry[1] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < nblocks; base += 1024) {
    int64_t i = base + threadIdx.x;
    unsigned long long v = i < nblocks ? bsum_v[i] : 0ull, t = i < nblocks ? bsum_t[i] : 0ull;
    unsigned long long iv = v, it = t;  // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long pv = __shfl_up_sync(0xffffffffu, iv, o), pt = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) { iv += pv; it += pt; }
    }
    if (lane == 31) { s_w[0][wid] = iv; s_w[1][wid] = it; }
    __syncthreads();
    if (wid == 0) {
      unsigned long long a = s_w[0][lane], b = s_w[1][lane];
      unsigned long long ia = a, ib = b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned long long pa = __shfl_up_sync(0xffffffffu, ia, o), pb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += pa; ib += pb; }
      }
      s_w[0][lane] = ia - a;  // exclusive over warps
      s_w[1][lane] = ib - b;
    }
    __syncthreads();
    unsigned long long ev = s_carry[0] + s_w[0][wid] + iv - v;
    unsigned long long et = s_carry[1] + s_w[1][wid] + it - t;
    if (i < nblocks) {
      bsum_v[i] = (uint32_t)ev;
      bsum_t[i] = et;
    }
    __syncthreads();
    if (threadIdx.x == 1023) {
      s_carry[0] = ev + v;
      s_carry[1] = et + t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[0] = s_carry[0];
    totals[1] = s_carry[1];
  }
}

__global__ void __launch_bounds__(kScanBlock) k_mc_scan_apply(uint4* __restrict__ info,
                                                              unsigned long long* __restrict__ toff, int64_t nwords,
                                                              const uint32_t* __restrict__ bsum_v,
                                                              const unsigned long long* __restrict__ bsum_t) {
  __shared__ unsigned long long s_w[2][kScanBlock / 32];
  const int64_t wi = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long v = wi < nwords ? info[wi].w : 0u, t = wi < nwords ? toff[wi] : 0ull;
  unsigned long long iv = v, it = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long pv = __shfl_up_sync(0xffffffffu, iv, o), pt = __shfl_up_sync(0xffffffffu, it, o);
    if (lane >= o) { iv += pv; it += pt; }
  }
  if (lane == 31) { s_w[0][wid] = iv; s_w[1][wid] = it; }
  __syncthreads();
  unsigned long long ov = bsum_v[blockIdx.x], ot = bsum_t[blockIdx.x];
  for (int k = 0; k < wid; ++k) { ov += s_w[0][k]; ot += s_w[1][k]; }
  if (wi < nwords) {
    info[wi].w = (uint32_t)(ov + iv - v);
    toff[wi] = ot + it - t;
  }
}

// ---- 4. emit -------------------------------------------------------------------------------
// The surface touches a few cells per word, so a warp-per-word emitter leaves most lanes idle
// (measured: 510 us, instruction-issue bound). Instead: one thread per VERTEX and one thread
// per ACTIVE CELL. A thread finds its word by a two-level search over the exclusive offsets
// (block sums: bisected once per warp, then a gallop from there; the 256 words of the block:
// bisected), then its voxel / cell inside the word by popcounts. Outputs of consecutive
// threads are consecutive in memory. (One thread per TRIANGLE was measured too: uniform
// work, but every triangle repeats the search and the eight row loads: 144 us against
// 128 us at 512^3.)
struct McXform {
  float sx, sy, sz;
  int ox, oy, oz;
  int flip_y;
  float iso;
};

// per edge id: axis and owner offset, packed (a | ox<<2 | oy<<3 | oz<<4)
__device__ __forceinline__ int edge_code(int e) {
  int a = e >> 2, cu = e & 1, cv = (e >> 1) & 1;
  int ox = 0, oy = 0, oz = 0;
  if (a == 0) { oy = cu; oz = cv; }
  else if (a == 1) { ox = cu; oz = cv; }
  else { ox = cu; oy = cv; }
  return a | (ox << 2) | (oy << 3) | (oz << 4);
}

// largest i in [0, n) with key(i) <= k; key is non-decreasing and key(0) <= k
template <typename F>
__device__ __forceinline__ int64_t last_le(int64_t n, uint32_t k, F key) {
  int64_t lo = 0, hi = n;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (key(mid) <= k) lo = mid; else hi = mid;
  }
  return lo;
}

// the same, knowing key(start) <= k and expecting the answer near start: gallop, then bisect
template <typename F>
__device__ __forceinline__ int64_t last_le_from(int64_t start, int64_t n, uint32_t k, F key) {
  int64_t lo = start, hi = start + 1, step = 1;
  while (hi < n && key(hi) <= k) { lo = hi; step <<= 1; hi = lo + step; }
  if (hi > n) hi = n;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (key(mid) <= k) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ uint32_t below(int i) { return i >= 32 ? 0xffffffffu : ((1u << i) - 1u); }

template <typename T>
__global__ void __launch_bounds__(256) k_mc_emit_verts(const T* __restrict__ vol, McGeom g,
                                                       const uint4* __restrict__ info,
                                                       const uint32_t* __restrict__ bsum_v, int64_t nblocks,
                                                       const unsigned long long* __restrict__ totals, McXform xf,
                                                       float* __restrict__ verts) {
  const uint32_t V = (uint32_t)totals[0];
  const uint32_t stride = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // a warp's 32 consecutive outputs start in the same block of words or close to it: lane 0
  // bisects the block sums once, the others gallop on from its answer
  for (uint32_t kb = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); kb < V; kb += stride) {
    const uint32_t k = kb + lane;
    auto bkey = [&](int64_t i) { return __ldg(bsum_v + i); };
    int64_t blk = lane == 0 ? last_le(nblocks, kb, bkey) : 0;
    blk = __shfl_sync(0xffffffffu, blk, 0);
    if (k >= V) continue;
    blk = last_le_from(blk, nblocks, k, bkey);
    const int64_t w0 = blk * kScanBlock;
    const int64_t nw = g.nwords - w0 < kScanBlock ? g.nwords - w0 : kScanBlock;
    const int64_t wi = w0 + last_le(nw, k, [&](int64_t i) { return __ldg(&info[w0 + i].w); });
    const uint4 inf = __ldg(info + wi);
    const uint32_t r = k - inf.w;
    // voxel: largest i with (#vertices of voxels below i) <= r
    int lo = 0, hi = 32;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      uint32_t m = below(mid);
      if ((uint32_t)(__popc(inf.x & m) + __popc(inf.y & m) + __popc(inf.z & m)) <= r) lo = mid; else hi = mid;
    }
    const int i = lo;
    const uint32_t m = below(i);
    int rr = (int)(r - (uint32_t)(__popc(inf.x & m) + __popc(inf.y & m) + __popc(inf.z & m)));
    const int bx = (inf.x >> i) & 1, by = (inf.y >> i) & 1;
    // rr-th crossing edge of the voxel in x, y, z order
    int axis;
    if (bx && rr == 0) axis = 0;
    else { rr -= bx; if (by && rr == 0) axis = 1; else axis = 2; }
    const int64_t row = wi / g.wx;
    const int w = (int)(wi - row * g.wx);
    const int64_t z = row / g.ny, y = row - z * g.ny;
    const int64_t x = (int64_t)w * 32 + i;
    const int64_t p = row * g.nx + x;
    const int64_t step = axis == 0 ? 1 : (axis == 1 ? g.nx : g.nx * g.ny);
    const float s0 = (float)vol[p], s1 = (float)vol[p + step];
    const float t = __fdiv_rn(__fsub_rn(xf.iso, s0), __fsub_rn(s1, s0));
    float fx = (float)((int)x + xf.ox), fy = (float)((int)y + xf.oy), fz = (float)((int)z + xf.oz);
    if (axis == 0) fx = __fadd_rn(fx, t); else if (axis == 1) fy = __fadd_rn(fy, t); else fz = __fadd_rn(fz, t);
    const float py = __fmul_rn(fy, xf.sy);
    float* o = verts + 3ll * k;
    o[0] = __fmul_rn(fx, xf.sx);
    o[1] = xf.flip_y ? -py : py;
    o[2] = __fmul_rn(fz, xf.sz);
  }
}

__global__ void __launch_bounds__(256) k_mc_emit_tris(McGeom g, const uint32_t* __restrict__ bits,
                                                      const uint4* __restrict__ info,
                                                      const unsigned long long* __restrict__ toff,
                                                      const uint32_t* __restrict__ amask,
                                                      const unsigned long long* __restrict__ bsum_t, int64_t nblocks,
                                                      const unsigned long long* __restrict__ totals, int flip_y,
                                                      int skip_last, int vbase, const uint4* __restrict__ foreign,
                                                      int foreign_base, int* __restrict__ tris) {
  __shared__ signed char s_tri[256][16];  // 15 edge ids + triangle count
  for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
    int c = i >> 4, k = i & 15;
    s_tri[c][k] = k < 15 ? B2V_MC_TRI[c][k] : (signed char)B2V_MC_NTRI[c];
  }
  __syncthreads();
  const uint32_t C = (uint32_t)(totals[1] >> 32);
  const uint32_t stride = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (uint32_t kb = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); kb < C; kb += stride) {
    const uint32_t k = kb + lane;
    auto bkey = [&](int64_t i) { return (uint32_t)(__ldg(bsum_t + i) >> 32); };
    int64_t blk = lane == 0 ? last_le(nblocks, kb, bkey) : 0;
    blk = __shfl_sync(0xffffffffu, blk, 0);
    if (k >= C) continue;
    blk = last_le_from(blk, nblocks, k, bkey);
    const int64_t w0 = blk * kScanBlock;
    const int64_t nw = g.nwords - w0 < kScanBlock ? g.nwords - w0 : kScanBlock;
    const int64_t wi = w0 + last_le(nw, k, [&](int64_t i) { return (uint32_t)(__ldg(toff + w0 + i) >> 32); });
    const unsigned long long tq = __ldg(toff + wi);
    const uint32_t act = __ldg(amask + wi);
    const int i = (int)__fns(act, 0, (int)(k - (uint32_t)(tq >> 32)) + 1);   // this thread's cell bit
    const int64_t row = wi / g.wx;
    const int w = (int)(wi - row * g.wx);
    const int64_t z = row / g.ny, y = row - z * g.ny;
    const Rows r = load_rows(bits, g, z, y, w);
    // triangles of the active cells before mine in this word
    int64_t tbase = (uint32_t)tq;
    for (uint32_t e = act & below(i); e; e &= e - 1) tbase += s_tri[cell_case(r, __ffs(e) - 1)][15];
    const int c = cell_case(r, i);
    const int ntri = s_tri[c][15];
    for (int t = 0; t < ntri; ++t) {
      int id[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int code = edge_code(s_tri[c][3 * t + m]);
        const int a = code & 3;
        const int qx = i + ((code >> 2) & 1);                      // 0..32 within the word pair
        const int64_t qy = y + ((code >> 3) & 1), qz = z + ((code >> 4) & 1);
        const int ob = qx & 31;
        const uint32_t ol = (1u << ob) - 1u;
        uint4 oi;
        int v;
        if (skip_last && qz == g.nz - 1) {
          // owned by the next shard: its records of that plane, its numbering
          oi = __ldg(foreign + qy * g.wx + (w + (qx >> 5)));
          v = foreign_base;
        } else {
          oi = __ldg(info + (qz * g.ny + qy) * g.wx + (w + (qx >> 5)));
          v = vbase;
        }
        v += (int)(oi.w + __popc(oi.x & ol) + __popc(oi.y & ol) + __popc(oi.z & ol));
        if (a > 0) v += (oi.x >> ob) & 1;
        if (a > 1) v += (oi.y >> ob) & 1;
        id[m] = v;
      }
      int* o = tris + 3 * (tbase + t);
      o[0] = id[0];
      o[1] = flip_y ? id[2] : id[1];
      o[2] = flip_y ? id[1] : id[2];
    }
  }
}

int grid_for(int64_t items, int per_block) {
  int64_t blocks = ceil_div64(items, per_block);
  int64_t cap = (int64_t)b2v_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// inside(p) <=> (double)S >= iso  <=>  S >= ceil(iso) for integer S
int int_threshold(double iso, int lo, int hi) {
  double c = ceil(iso);
  if (c <= (double)lo) return lo;          // everything inside
  if (c > (double)hi) return hi + 1;       // nothing inside
  return (int)c;
}

}  // namespace

extern "C" int64_t b2v_mc_workspace_bytes(int64_t nz, int64_t ny, int64_t nx) {
  if (nz <= 0 || ny <= 0 || nx <= 0) return 0;
  return carve(nullptr, make_geom(nz, ny, nx)).bytes;
}

static int mc_count_impl(const void* vol, int dtype, int64_t nz, int64_t ny, int64_t nx, double iso, int skip_last,
                         void* workspace, void* stream, int64_t* nverts_host, int64_t* ntris_host) {
  B2V_REQUIRE(vol && workspace && nverts_host && ntris_host, B2V_ERR_ARG, "mc_count: null pointer");
  B2V_REQUIRE(nz > 0 && ny > 0 && nx > 0, B2V_ERR_ARG, "mc_count: empty volume");
  B2V_REQUIRE(dtype == B2V_U8 || dtype == B2V_I16, B2V_ERR_ARG, "mc_count: dtype must be uint8 or int16");
  B2V_REQUIRE(iso == iso, B2V_ERR_ARG, "mc_count: iso is NaN");
  McGeom g = make_geom(nz, ny, nx);
  B2V_REQUIRE(g.nwords < (1ll << 31), B2V_ERR_ARG, "mc_count: volume too large for one call; shard along z");
  McWs w = carve(workspace, g);
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if (dtype == B2V_U8) {
    int thr = int_threshold(iso, 0, 255);
    if (nx % 16 == 0 && b2v_aligned16(vol))
      if (nx % 32 == 0)
        k_mc_bits_u8_vec<true><<<grid_for(g.nwords * 2, 1024), 256, 0, s>>>((const uint8_t*)vol, g, thr, w.bits);
      else
        k_mc_bits_u8_vec<false><<<grid_for(g.nwords * 2, 1024), 256, 0, s>>>((const uint8_t*)vol, g, thr, w.bits);
    else
      k_mc_bits<uint8_t><<<grid_for(g.nwords, 8), 256, 0, s>>>((const uint8_t*)vol, g, thr, w.bits);
  } else {
    int thr = int_threshold(iso, -32768, 32767);
    if (nx % 8 == 0 && b2v_aligned16(vol))
      k_mc_bits_i16_vec<<<grid_for(g.nwords * 4, 256), 256, 0, s>>>((const int16_t*)vol, g, thr, w.bits);
    else
      k_mc_bits<int16_t><<<grid_for(g.nwords, 8), 256, 0, s>>>((const int16_t*)vol, g, thr, w.bits);
  }
  if ((rc = b2v_check_launch("k_mc_bits"))) return rc;
  k_mc_count<<<(unsigned)w.nblocks, kScanBlock, 0, s>>>(w.bits, g, skip_last, w.info, w.toff, w.amask, w.bsum_v,
                                                        w.bsum_t);
  if ((rc = b2v_check_launch("k_mc_count"))) return rc;
  k_mc_scan_bsums<<<1, 1024, 0, s>>>(w.bsum_v, w.bsum_t, w.nblocks, w.totals);
  if ((rc = b2v_check_launch("k_mc_scan_bsums"))) return rc;
  k_mc_scan_apply<<<(unsigned)w.nblocks, kScanBlock, 0, s>>>(w.info, w.toff, g.nwords, w.bsum_v, w.bsum_t);
  if ((rc = b2v_check_launch("k_mc_scan_apply"))) return rc;
  unsigned long long tot[2] = {0, 0};
  B2V_CUDA(cudaMemcpyAsync(tot, w.totals, sizeof(tot), cudaMemcpyDeviceToHost, s));
  B2V_CUDA(cudaStreamSynchronize(s));
  const unsigned long long ntri = tot[1] & 0xffffffffull;
  B2V_REQUIRE(tot[0] < (1ull << 31) && ntri < (1ull << 31) && (tot[1] >> 32) < (1ull << 31), B2V_ERR_RANGE,
              "mc_count: %llu vertices / %llu triangles exceed int32 indices; shard along z", tot[0], ntri);
  *nverts_host = (int64_t)tot[0];
  *ntris_host = (int64_t)ntri;
  return B2V_OK;
}

static int mc_emit_impl(const void* vol, int dtype, int64_t nz, int64_t ny, int64_t nx, double iso,
                        const void* workspace, float sx, float sy, float sz, int32_t ox, int32_t oy, int32_t oz,
                        int flip_y, int skip_last, int32_t vbase, const void* foreign, int32_t foreign_base,
                        float* verts, int32_t* tris, void* stream) {
  B2V_REQUIRE(vol && workspace, B2V_ERR_ARG, "mc_emit: null pointer");
  B2V_REQUIRE(dtype == B2V_U8 || dtype == B2V_I16, B2V_ERR_ARG, "mc_emit: dtype must be uint8 or int16");
  B2V_REQUIRE(!skip_last || foreign, B2V_ERR_ARG, "mc_emit: a shard that skips its last plane needs the next shard's records");
  McGeom g = make_geom(nz, ny, nx);
  McWs w = carve(const_cast<void*>(workspace), g);
  cudaStream_t s = (cudaStream_t)stream;
  McXform xf = {sx, sy, sz, ox, oy, oz, flip_y ? 1 : 0, (float)iso};
  const int grid = b2v_sm_count() * 8;   // grid-stride over the counts that live on the device
  int rc;
  if (verts) {
    if (dtype == B2V_U8)
      k_mc_emit_verts<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)vol, g, w.info, w.bsum_v, w.nblocks, w.totals, xf,
                                                    verts);
    else
      k_mc_emit_verts<int16_t><<<grid, 256, 0, s>>>((const int16_t*)vol, g, w.info, w.bsum_v, w.nblocks, w.totals, xf,
                                                    verts);
    if ((rc = b2v_check_launch("k_mc_emit_verts"))) return rc;
  }
  if (tris) {
    k_mc_emit_tris<<<grid, 256, 0, s>>>(g, w.bits, w.info, w.toff, w.amask, w.bsum_t, w.nblocks, w.totals,
                                        flip_y ? 1 : 0, skip_last, vbase, (const uint4*)foreign, foreign_base, tris);
    if ((rc = b2v_check_launch("k_mc_emit_tris"))) return rc;
  }
  return B2V_OK;
}

extern "C" int b2v_mc_count(const void* vol, int dtype, int64_t nz, int64_t ny, int64_t nx, double iso,
                            void* workspace, void* stream, int64_t* nverts_host, int64_t* ntris_host) {
  return mc_count_impl(vol, dtype, nz, ny, nx, iso, 0, workspace, stream, nverts_host, ntris_host);
}

extern "C" int b2v_mc_emit(const void* vol, int dtype, int64_t nz, int64_t ny, int64_t nx, double iso,
                           const void* workspace, float sx, float sy, float sz, int32_t ox, int32_t oy, int32_t oz,
                           int flip_y, float* verts, int32_t* tris, void* stream) {
  return mc_emit_impl(vol, dtype, nz, ny, nx, iso, workspace, sx, sy, sz, ox, oy, oz, flip_y, 0, 0, nullptr, 0, verts,
                      tris, stream);
}

// ---- Z-sharded variants (dist.py): the slab passed in ends with the plane it shares with
// the next shard; that plane's vertices belong to the next shard (skip_last_plane).
extern "C" int b2v_mc_count_shard(const void* vol, int dtype, int64_t nz, int64_t ny, int64_t nx, double iso,
                                  int skip_last_plane, void* workspace, void* stream, int64_t* nverts_host,
                                  int64_t* ntris_host) {
  return mc_count_impl(vol, dtype, nz, ny, nx, iso, skip_last_plane ? 1 : 0, workspace, stream, nverts_host,
                       ntris_host);
}

extern "C" int b2v_mc_emit_shard(const void* vol, int dtype, int64_t nz, int64_t ny, int64_t nx, double iso,
                                 const void* workspace, float sx, float sy, float sz, int32_t ox, int32_t oy,
                                 int32_t oz, int flip_y, int skip_last_plane, int32_t vertex_base,
                                 const void* next_shard_plane0_records, int32_t next_shard_vertex_base, float* verts,
                                 int32_t* tris, void* stream) {
  return mc_emit_impl(vol, dtype, nz, ny, nx, iso, workspace, sx, sy, sz, ox, oy, oz, flip_y, skip_last_plane ? 1 : 0,
                      vertex_base, next_shard_plane0_records, next_shard_vertex_base, verts, tris, stream);
}

extern "C" int b2v_mc_layout(int64_t nz, int64_t ny, int64_t nx, int64_t* layout_out) {
  B2V_REQUIRE(nz > 0 && ny > 0 && nx > 0 && layout_out, B2V_ERR_ARG, "mc_layout: bad arguments");
  McGeom g = make_geom(nz, ny, nx);
  McWs w = carve(nullptr, g);
  layout_out[0] = (int64_t)((char*)w.info - (char*)nullptr);   // byte offset of the per-word records
  layout_out[1] = (int64_t)ny * g.wx * 16;                      // bytes of records per z-plane
  return B2V_OK;
}
