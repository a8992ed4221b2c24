/* b2v.h — C ABI of libb2v.so, the Blackwell-native (sm_100a) replacement for the
 * per-voxel hot path of InVesalius 3.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - volumes are dense C-order [dz][dy][dx] (x fastest); strided / file-backed host
 *     views are packed by b2v_copy3d_* before the call;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value is a B2V_* status; b2v_last_error() gives the message of the
 *     last failure on the calling thread;
 *   - nothing is retained after return; scratch memory is supplied by the caller
 *     (query the size with the matching *_workspace_bytes function);
 *   - no entry point synchronises the stream unless documented.
 *
 * Each function cites the reference interface it replaces (paths relative to the
 * invesalius3 checkout).
 */
#ifndef B2V_H
#define B2V_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2V_OK 0
#define B2V_ERR_ARG 1      /* bad argument (shape, dtype code, axis, alignment) */
#define B2V_ERR_CUDA 2     /* CUDA runtime failure */
#define B2V_ERR_RANGE 3    /* value not representable (mirrors a Rust NumCast panic) */
#define B2V_ERR_NOCONV 4   /* iterative kernel hit its round cap */

/* element type codes (reference: invesalius_rs/src/types.rs:5-70) */
#define B2V_I16 0
#define B2V_U8 1
#define B2V_F64 2

/* projection kinds for b2v_mip */
#define B2V_MIP_MAX 0
#define B2V_MIP_MIN 1
#define B2V_MIP_MEAN 2

const char* b2v_last_error(void);
int b2v_version(void);
/* number of kernels launched by this library on the calling thread since the last reset */
int64_t b2v_launch_count(void);
void b2v_launch_count_reset(void);

/* ---- host<->device packing of strided views --------------------------------
 * Replaces the implicit stride handling of rust-numpy `as_array()` for callers that
 * pass `mask.matrix[1:,1:,1:]` (invesalius/data/styles.py:2493,2917,3157).
 * Copies a [dz][dy][dx] box of `elem` byte elements between a host view with byte
 * pitches (row_pitch, plane_pitch; x contiguous) and a dense device buffer. */
int b2v_copy3d_h2d(void* dst_dev, const void* src_host, int64_t dz, int64_t dy, int64_t dx, int64_t elem,
                   int64_t src_row_pitch, int64_t src_plane_pitch, void* stream);
int b2v_copy3d_d2h(void* dst_host, const void* src_dev, int64_t dz, int64_t dy, int64_t dx, int64_t elem,
                   int64_t dst_row_pitch, int64_t dst_plane_pitch, void* stream);

/* ---- threshold ---------------------------------------------------------------
 * Slice.SetMaskThreshold (whole-volume branch) invesalius/data/slice_.py:1238-1246:
 *     mask[v] = 255 if lo <= img[v] <= hi else 0               (preserve_markers = 0)
 * Slice.do_threshold_to_a_slice / do_threshold_to_all_slices slice_.py:1722-1769:
 *     same, except voxels whose OLD mask value is 1, 2, 253 or 254 keep it
 *                                                              (preserve_markers = 1)
 * img: int16 [n]; mask: uint8 [n] (read only when preserve_markers). Elementwise over
 * n voxels of a dense buffer. Algorithmic bytes: 3 B/voxel (4 with preservation). */
int b2v_threshold_i16(const int16_t* img, int64_t n, int32_t lo, int32_t hi, uint8_t* mask,
                      int preserve_markers, void* stream);
/* same on the padded Mask layout of invesalius/data/mask.py:422-431: mask has shape
 * [dz+1][dy+1][dx+1], voxel (z,y,x) at [z+1][y+1][x+1]; also sets the axial flag
 * mask[z+1][0][0] = 1 (slice_.py:1246,1767). When only_dirty != 0, slices whose flag is
 * already non-zero are skipped (slice_.py:1762-1763). */
int b2v_threshold_i16_masklayout(const int16_t* img, int64_t dz, int64_t dy, int64_t dx, int32_t lo, int32_t hi,
                                 uint8_t* mask_padded, int preserve_markers, int only_dirty, void* stream);

/* ---- intensity projections ---------------------------------------------------
 * NumPy reductions in Slice.get_image_slice, invesalius/data/slice_.py:881-886 /
 * 970-975 / 1057-1062: tmp_array.max(axis) / .min(axis) / .mean(axis).
 * img dense [dz][dy][dx] of dtype code `dtype` (int16 or uint8); axis 0/1/2.
 * out: same dtype for MAX/MIN, float64 for MEAN; shape [dy][dx] / [dz][dx] / [dz][dy].
 * workspace: b2v_mip_workspace_bytes() bytes (may be 0). 2 B/voxel for int16. */
int64_t b2v_mip_workspace_bytes(int dtype, int64_t dz, int64_t dy, int64_t dx, int axis, int kind);
int b2v_mip(const void* img, int dtype, int64_t dz, int64_t dy, int64_t dx, int axis, int kind, void* out,
            void* workspace, void* stream);

/* global min/max of a dense buffer as float32 (first step of MIDA,
 * invesalius_rs/src/mips.rs:113-122). minmax_out: float[2] on the device. */
int64_t b2v_minmax_workspace_bytes(int64_t n);
int b2v_minmax_f32(const void* img, int dtype, int64_t n, float* minmax_out, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2V_H */
