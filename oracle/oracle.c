/* oracle.c — CPU restatement of the InVesalius per-voxel hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under invesalius3_b200/ may import, link or call
 * this file; it exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs can check and time the reference semantics on the CPU.
 *
 * The reference's native code is a Rust crate (invesalius_rs) that cannot be built in
 * this image (no cargo/rustc), so each function below restates the Rust/NumPy source
 * line by line, citing file:line under the invesalius3 checkout. Compiled with
 * -ffp-contract=off: Rust never fuses a*b+c, and neither may we.
 *
 * Parity pinning (see tests/test_oracle_golden.py):
 *   threshold        pinned  (reference tests + samples/Cranium.inv3 masks)
 *   flood fill       pinned  (reference tests' known answers)
 *   fill holes       pinned  (reference test's known answer)
 *   MIP/MIDA/LMIP/contour-MIP   parity unpinned (the reference has no test/golden)
 *
 * All arrays are addressed through explicit ELEMENT strides so that the strided
 * memmap views the reference's callers pass work unchanged.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX3(z, y, x, s) ((z) * (s)[0] + (y) * (s)[1] + (x) * (s)[2])

/* ------------------------------------------------------------------------- */
/* threshold — invesalius/data/slice_.py:1238-1246 (no preservation) and       */
/* :1731-1737 (markers 1, 2, 253, 254 of the old mask survive)                 */
/* ------------------------------------------------------------------------- */
void orc_threshold_i16(const int16_t* img, const int64_t* is, int64_t dz, int64_t dy, int64_t dx, double lo,
                       double hi, uint8_t* mask, const int64_t* ms, int preserve) {
  for (int64_t z = 0; z < dz; ++z)
    for (int64_t y = 0; y < dy; ++y)
      for (int64_t x = 0; x < dx; ++x) {
        double v = (double)img[IDX3(z, y, x, is)];
        uint8_t m = (v >= lo && v <= hi) ? 255 : 0;
        uint8_t* o = &mask[IDX3(z, y, x, ms)];
        if (preserve && (*o == 1 || *o == 2 || *o == 253 || *o == 254)) m = *o;
        *o = m;
      }
}

/* ------------------------------------------------------------------------- */
/* flood fill family — invesalius_rs/src/floodfill.rs                          */
/* ------------------------------------------------------------------------- */
typedef struct {
  int64_t* buf;
  int64_t cap, head, tail; /* deque over a growable ring is overkill: a vector with a
                              moving head reproduces VecDeque push_back/pop_back/pop_front */
} vec_t;
static void vec_init(vec_t* v) {
  v->cap = 1 << 16;
  v->buf = (int64_t*)malloc(sizeof(int64_t) * 3 * v->cap);
  v->head = v->tail = 0;
}
static void vec_push(vec_t* v, int64_t x, int64_t y, int64_t z) {
  if (v->tail == v->cap) {
    if (v->head > v->cap / 2) { /* compact */
      memmove(v->buf, v->buf + 3 * v->head, sizeof(int64_t) * 3 * (v->tail - v->head));
      v->tail -= v->head;
      v->head = 0;
    } else {
      v->cap *= 2;
      v->buf = (int64_t*)realloc(v->buf, sizeof(int64_t) * 3 * v->cap);
    }
  }
  int64_t* p = v->buf + 3 * v->tail++;
  p[0] = x; p[1] = y; p[2] = z;
}
static int vec_pop_back(vec_t* v, int64_t* x, int64_t* y, int64_t* z) {
  if (v->tail == v->head) return 0;
  int64_t* p = v->buf + 3 * --v->tail;
  *x = p[0]; *y = p[1]; *z = p[2];
  return 1;
}
static int vec_pop_front(vec_t* v, int64_t* x, int64_t* y, int64_t* z) {
  if (v->tail == v->head) return 0;
  int64_t* p = v->buf + 3 * v->head++;
  *x = p[0]; *y = p[1]; *z = p[2];
  return 1;
}

/* generic_floodfill_threshold, floodfill.rs:96-166. Returns -1 for a seed outside the
 * volume (the Rust code panics on the out-of-bounds index, floodfill.rs:122). */
#define DEF_FF_THRESHOLD(NAME, T)                                                                             \
  int NAME(const T* data, const int64_t* ds, int64_t dz, int64_t dy, int64_t dx, const int64_t* seeds,        \
           int64_t nseeds, T t0, T t1, uint8_t fill, const uint8_t* strct, int64_t odz, int64_t ody,          \
           int64_t odx, uint8_t* out, const int64_t* os) {                                                    \
    int64_t offz = odz / 2, offy = ody / 2, offx = odx / 2;                                                   \
    for (int64_t s = 0; s < nseeds; ++s) {                                                                    \
      int64_t i = seeds[3 * s], j = seeds[3 * s + 1], k = seeds[3 * s + 2];                                   \
      if (i < 0 || j < 0 || k < 0 || i >= dx || j >= dy || k >= dz) return -1;                                \
    }                                                                                                         \
    vec_t st;                                                                                                 \
    vec_init(&st);                                                                                            \
    for (int64_t s = 0; s < nseeds; ++s) {                                                                    \
      int64_t i = seeds[3 * s], j = seeds[3 * s + 1], k = seeds[3 * s + 2];                                   \
      T val = data[IDX3(k, j, i, ds)];                                                                        \
      if (val >= t0 && val <= t1) {                                                                           \
        vec_push(&st, i, j, k);                                                                               \
        out[IDX3(k, j, i, os)] = fill;                                                                        \
      }                                                                                                       \
    }                                                                                                         \
    int64_t x, y, z;                                                                                          \
    while (vec_pop_back(&st, &x, &y, &z)) {                                                                   \
      out[IDX3(z, y, x, os)] = fill;                                                                          \
      for (int64_t kk = 0; kk < odz; ++kk) {                                                                  \
        int64_t zo = z + kk - offz;                                                                           \
        if (zo < 0 || zo >= dz) continue;                                                                     \
        for (int64_t jj = 0; jj < ody; ++jj) {                                                                \
          int64_t yo = y + jj - offy;                                                                         \
          if (yo < 0 || yo >= dy) continue;                                                                   \
          for (int64_t ii = 0; ii < odx; ++ii) {                                                              \
            if (strct[(kk * ody + jj) * odx + ii] != 0) {                                                     \
              int64_t xo = x + ii - offx;                                                                     \
              if (xo < 0 || xo >= dx) continue;                                                               \
              if (out[IDX3(zo, yo, xo, os)] != fill) {                                                        \
                T val = data[IDX3(zo, yo, xo, ds)];                                                           \
                if (val >= t0 && val <= t1) {                                                                 \
                  out[IDX3(zo, yo, xo, os)] = fill;                                                           \
                  vec_push(&st, xo, yo, zo);                                                                  \
                }                                                                                             \
              }                                                                                               \
            }                                                                                                 \
          }                                                                                                   \
        }                                                                                                     \
      }                                                                                                       \
    }                                                                                                         \
    free(st.buf);                                                                                             \
    return 0;                                                                                                 \
  }
DEF_FF_THRESHOLD(orc_floodfill_threshold_i16, int16_t)
DEF_FF_THRESHOLD(orc_floodfill_threshold_u8, uint8_t)
DEF_FF_THRESHOLD(orc_floodfill_threshold_f64, double)

/* generic_floodfill_threshold_inplace, floodfill.rs:168-237 */
#define DEF_FF_INPLACE(NAME, T)                                                                               \
  int NAME(T* data, const int64_t* ds, int64_t dz, int64_t dy, int64_t dx, const int64_t* seeds,              \
           int64_t nseeds, T t0, T t1, T fill, const uint8_t* strct, int64_t odz, int64_t ody, int64_t odx) { \
    int64_t offz = odz / 2, offy = ody / 2, offx = odx / 2;                                                   \
    for (int64_t s = 0; s < nseeds; ++s) {                                                                    \
      int64_t i = seeds[3 * s], j = seeds[3 * s + 1], k = seeds[3 * s + 2];                                   \
      if (i < 0 || j < 0 || k < 0 || i >= dx || j >= dy || k >= dz) return -1;                                \
    }                                                                                                         \
    vec_t st;                                                                                                 \
    vec_init(&st);                                                                                            \
    for (int64_t s = 0; s < nseeds; ++s) {                                                                    \
      int64_t i = seeds[3 * s], j = seeds[3 * s + 1], k = seeds[3 * s + 2];                                   \
      T val = data[IDX3(k, j, i, ds)];                                                                        \
      if (val >= t0 && val <= t1) {                                                                           \
        vec_push(&st, i, j, k);                                                                               \
        data[IDX3(k, j, i, ds)] = fill;                                                                       \
      }                                                                                                       \
    }                                                                                                         \
    int64_t x, y, z;                                                                                          \
    while (vec_pop_back(&st, &x, &y, &z)) {                                                                   \
      data[IDX3(z, y, x, ds)] = fill;                                                                         \
      for (int64_t kk = 0; kk < odz; ++kk) {                                                                  \
        int64_t zo = z + kk - offz;                                                                           \
        if (zo < 0 || zo >= dz) continue;                                                                     \
        for (int64_t jj = 0; jj < ody; ++jj) {                                                                \
          int64_t yo = y + jj - offy;                                                                         \
          if (yo < 0 || yo >= dy) continue;                                                                   \
          for (int64_t ii = 0; ii < odx; ++ii) {                                                              \
            if (strct[(kk * ody + jj) * odx + ii] != 0) {                                                     \
              int64_t xo = x + ii - offx;                                                                     \
              if (xo < 0 || xo >= dx) continue;                                                               \
              if (data[IDX3(zo, yo, xo, ds)] != fill) {                                                       \
                T val = data[IDX3(zo, yo, xo, ds)];                                                           \
                if (val >= t0 && val <= t1) {                                                                 \
                  data[IDX3(zo, yo, xo, ds)] = fill;                                                          \
                  vec_push(&st, xo, yo, zo);                                                                  \
                }                                                                                             \
              }                                                                                               \
            }                                                                                                 \
          }                                                                                                   \
        }                                                                                                     \
      }                                                                                                       \
    }                                                                                                         \
    free(st.buf);                                                                                             \
    return 0;                                                                                                 \
  }
DEF_FF_INPLACE(orc_floodfill_threshold_inplace_i16, int16_t)
DEF_FF_INPLACE(orc_floodfill_threshold_inplace_u8, uint8_t)
DEF_FF_INPLACE(orc_floodfill_threshold_inplace_f64, double)

/* floodfill_internal, floodfill.rs:5-49: 6-connected BFS over data == v; the seed is
 * marked unconditionally (floodfill.rs:21). (i, j, k) = (x, y, z). */
#define DEF_FF_EQ(NAME, T)                                                                                  \
  int NAME(const T* data, const int64_t* ds, int64_t d, int64_t h, int64_t w, int64_t i, int64_t j,         \
           int64_t k, T v, uint8_t fill, uint8_t* out, const int64_t* os) {                                 \
    if (i < 0 || j < 0 || k < 0 || i >= w || j >= h || k >= d) return -1;                                   \
    vec_t st;                                                                                               \
    vec_init(&st);                                                                                          \
    vec_push(&st, i, j, k);                                                                                 \
    out[IDX3(k, j, i, os)] = fill;                                                                          \
    int64_t x, y, z;                                                                                        \
    while (vec_pop_front(&st, &x, &y, &z)) {                                                                \
      if (z + 1 < d && data[IDX3(z + 1, y, x, ds)] == v && out[IDX3(z + 1, y, x, os)] != fill) {            \
        out[IDX3(z + 1, y, x, os)] = fill; vec_push(&st, x, y, z + 1);                                      \
      }                                                                                                     \
      if (z > 0 && data[IDX3(z - 1, y, x, ds)] == v && out[IDX3(z - 1, y, x, os)] != fill) {                \
        out[IDX3(z - 1, y, x, os)] = fill; vec_push(&st, x, y, z - 1);                                      \
      }                                                                                                     \
      if (y + 1 < h && data[IDX3(z, y + 1, x, ds)] == v && out[IDX3(z, y + 1, x, os)] != fill) {            \
        out[IDX3(z, y + 1, x, os)] = fill; vec_push(&st, x, y + 1, z);                                      \
      }                                                                                                     \
      if (y > 0 && data[IDX3(z, y - 1, x, ds)] == v && out[IDX3(z, y - 1, x, os)] != fill) {                \
        out[IDX3(z, y - 1, x, os)] = fill; vec_push(&st, x, y - 1, z);                                      \
      }                                                                                                     \
      if (x + 1 < w && data[IDX3(z, y, x + 1, ds)] == v && out[IDX3(z, y, x + 1, os)] != fill) {            \
        out[IDX3(z, y, x + 1, os)] = fill; vec_push(&st, x + 1, y, z);                                      \
      }                                                                                                     \
      if (x > 0 && data[IDX3(z, y, x - 1, ds)] == v && out[IDX3(z, y, x - 1, os)] != fill) {                \
        out[IDX3(z, y, x - 1, os)] = fill; vec_push(&st, x - 1, y, z);                                      \
      }                                                                                                     \
    }                                                                                                       \
    free(st.buf);                                                                                           \
    return 0;                                                                                               \
  }
DEF_FF_EQ(orc_floodfill_i16, int16_t)
DEF_FF_EQ(orc_floodfill_u8, uint8_t)
DEF_FF_EQ(orc_floodfill_f64, double)

/* fill_holes_automatically_internal, floodfill.rs:51-94 */
int orc_fill_holes_automatically(uint8_t* mask, const int64_t* ms, const uint32_t* labels, const int64_t* ls,
                                 int64_t dz, int64_t dy, int64_t dx, uint32_t nlabels, uint32_t max_size) {
  uint32_t* sizes = (uint32_t*)calloc((size_t)nlabels + 1, sizeof(uint32_t));
  for (int64_t z = 0; z < dz; ++z)
    for (int64_t y = 0; y < dy; ++y)
      for (int64_t x = 0; x < dx; ++x) {
        uint32_t l = labels[IDX3(z, y, x, ls)];
        if (l > nlabels) { free(sizes); return -1; } /* Rust: index out of bounds panic */
        sizes[l] += 1;
      }
  int modified = 0;
  for (uint32_t l = 0; l <= nlabels; ++l)
    if (sizes[l] > 0 && sizes[l] <= max_size) { modified = 1; break; }
  if (!modified) { free(sizes); return 0; }
  for (int64_t z = 0; z < dz; ++z)
    for (int64_t y = 0; y < dy; ++y)
      for (int64_t x = 0; x < dx; ++x)
        if (sizes[labels[IDX3(z, y, x, ls)]] <= max_size) mask[IDX3(z, y, x, ms)] = 254;
  free(sizes);
  return 1;
}

/* ------------------------------------------------------------------------- */
/* projections — invesalius_rs/src/mips.rs                                     */
/* ------------------------------------------------------------------------- */
/* Views: image [sz][sy][sx] with element strides; out 2-D with element strides:
 * axis 0 -> out[y][x], axis 1 -> out[z][x], axis 2 -> out[z][y]  (mips.rs:125-133). */

/* get_opacity, mips.rs:88-100 */
static float get_opacity(float vl, float wl, float ww) {
  float min_value = wl - (ww / 2.0f);
  float max_value = wl + (ww / 2.0f);
  if (vl < min_value) return 0.0f;
  else if (vl > max_value) return 1.0f;
  else return (vl - min_value) / (max_value - min_value);
}

/* NumCast::from(f32) -> integer: truncation toward zero, fails outside the type's
 * range or for NaN (num-traits float_to_int). Returns 0 on failure. */
static int cast_f32_i16(float f, int16_t* o) {
  if (!(f > -32769.0f && f < 32768.0f)) return 0;
  *o = (int16_t)f;
  return 1;
}
static int cast_f32_u8(float f, uint8_t* o) {
  if (!(f > -1.0f && f < 256.0f)) return 0;
  *o = (uint8_t)f;
  return 1;
}
static int cast_f32_f64(float f, double* o) { *o = (double)f; return 1; }

#define RAY_GEOM(axis, sz, sy, sx, is, n_r, n_c, n_l, s_r, s_c, s_l)               \
  int64_t n_r, n_c, n_l, s_r, s_c, s_l;                                            \
  if (axis == 0) { n_r = sy; n_c = sx; n_l = sz; s_r = is[1]; s_c = is[2]; s_l = is[0]; }       \
  else if (axis == 1) { n_r = sz; n_c = sx; n_l = sy; s_r = is[0]; s_c = is[2]; s_l = is[1]; }  \
  else { n_r = sz; n_c = sy; n_l = sx; s_r = is[0]; s_c = is[1]; s_l = is[2]; }

/* mida_internal, mips.rs:102-168. Returns 0 ok, 3 if a cast panicked. */
#define DEF_MIDA(NAME, T, U, CAST)                                                                           \
  int NAME(const T* image, const int64_t* is, int64_t sz, int64_t sy, int64_t sx, int axis, float wl,       \
           float ww, U* out, const int64_t* os) {                                                            \
    float img_min = 0, img_max = 0;                                                                          \
    int first = 1;                                                                                           \
    for (int64_t z = 0; z < sz; ++z)                                                                         \
      for (int64_t y = 0; y < sy; ++y)                                                                       \
        for (int64_t x = 0; x < sx; ++x) {                                                                   \
          float v = (float)image[IDX3(z, y, x, is)];                                                         \
          if (first) { img_min = img_max = v; first = 0; }                                                   \
          else { img_min = fminf(img_min, v); img_max = fmaxf(img_max, v); }                                 \
        }                                                                                                    \
    float range = img_max - img_min;                                                                         \
    RAY_GEOM(axis, sz, sy, sx, is, n_r, n_c, n_l, s_r, s_c, s_l)                                             \
    int bad = 0;                                                                                             \
    for (int64_t r = 0; r < n_r; ++r)                                                                        \
      for (int64_t c = 0; c < n_c; ++c) {                                                                    \
        const T* lane = image + r * s_r + c * s_c;                                                           \
        float fmax = 0.0f, alpha_p = 0.0f, colour_p = 0.0f, final_colour = 0.0f;                             \
        for (int64_t l = 0; l < n_l; ++l) {                                                                  \
          float vl = (float)lane[l * s_l];                                                                   \
          float fpi = (1.0f / range) * (vl - img_min);                                                       \
          float dl;                                                                                          \
          if (fpi > fmax) { dl = fpi - fmax; fmax = fpi; } else dl = 0.0f;                                   \
          float bt = 1.0f - dl;                                                                              \
          float alpha = get_opacity(vl, wl, ww);                                                             \
          float colour = (bt * colour_p) + (1.0f - bt * alpha_p) * fpi * alpha;                              \
          float current_alpha = (bt * alpha_p) + (1.0f - bt * alpha_p) * alpha;                              \
          colour_p = colour; alpha_p = current_alpha; final_colour = colour;                                 \
          if (current_alpha >= 1.0f) break;                                                                  \
        }                                                                                                    \
        U o;                                                                                                 \
        if (!CAST(range * final_colour + img_min, &o)) { bad = 1; continue; }                                \
        out[r * os[0] + c * os[1]] = o;                                                                      \
      }                                                                                                      \
    return bad ? 3 : 0;                                                                                      \
  }
DEF_MIDA(orc_mida_i16, int16_t, int16_t, cast_f32_i16)
DEF_MIDA(orc_mida_u8, uint8_t, uint8_t, cast_f32_u8)
DEF_MIDA(orc_mida_f64, double, uint8_t, cast_f32_u8)

/* lmip, mips.rs:7-86 (T -> U cast is the identity for the pairs the module offers) */
#define DEF_LMIP(NAME, T)                                                                                   \
  void NAME(const T* image, const int64_t* is, int64_t sz, int64_t sy, int64_t sx, int axis, T tmin, T tmax, \
            T* out, const int64_t* os) {                                                                     \
    RAY_GEOM(axis, sz, sy, sx, is, n_r, n_c, n_l, s_r, s_c, s_l)                                             \
    for (int64_t r = 0; r < n_r; ++r)                                                                        \
      for (int64_t c = 0; c < n_c; ++c) {                                                                    \
        const T* lane = image + r * s_r + c * s_c;                                                           \
        T max_val = lane[0];                                                                                 \
        int start = max_val >= tmin && max_val <= tmax;                                                      \
        for (int64_t l = 0; l < n_l; ++l) {                                                                  \
          T val = lane[l * s_l];                                                                             \
          if (val > max_val) max_val = val;                                                                  \
          else if (val < max_val && start) break;                                                            \
          if (val >= tmin && val <= tmax) start = 1;                                                         \
        }                                                                                                    \
        out[r * os[0] + c * os[1]] = max_val;                                                                \
      }                                                                                                      \
  }
DEF_LMIP(orc_lmip_i16, int16_t)
DEF_LMIP(orc_lmip_u8, uint8_t)
DEF_LMIP(orc_lmip_f64, double)

/* finite_difference + calc_fcm_intensity, mips.rs:170-213. The difference is taken in
 * T (wrapping for the integer types in a release build) before the f32 conversion.
 * Writes the T-typed temp volume `tmp` (dense [sz][sy][sx]); returns 3 if a cast to T
 * would have panicked (mips.rs:241). */
#define DEF_FCM(NAME, T, WRAPDIFF, CAST)                                                                    \
  int NAME(const T* image, const int64_t* is, int64_t sz, int64_t sy, int64_t sx, float n, int axis,        \
           T* tmp) {                                                                                         \
    float dir[3] = {0.0f, 0.0f, 0.0f};                                                                       \
    if (axis == 0) dir[2] = 1.0f; else if (axis == 1) dir[1] = 1.0f; else if (axis == 2) dir[0] = 1.0f;      \
    int bad = 0;                                                                                             \
    for (int64_t z = 0; z < sz; ++z)                                                                         \
      for (int64_t y = 0; y < sy; ++y)                                                                       \
        for (int64_t x = 0; x < sx; ++x) {                                                                   \
          int64_t px = x == 0 ? 0 : x - 1, fx = x == sx - 1 ? sx - 1 : x + 1;                                \
          int64_t py = y == 0 ? 0 : y - 1, fy = y == sy - 1 ? sy - 1 : y + 1;                                \
          int64_t pz = z == 0 ? 0 : z - 1, fz = z == sz - 1 ? sz - 1 : z + 1;                                \
          float h = 1.0f;                                                                                    \
          float gx = (float)WRAPDIFF(image[IDX3(z, y, fx, is)], image[IDX3(z, y, px, is)]) / (2.0f * h);     \
          float gy = (float)WRAPDIFF(image[IDX3(z, fy, x, is)], image[IDX3(z, py, x, is)]) / (2.0f * h);     \
          float gz = (float)WRAPDIFF(image[IDX3(fz, y, x, is)], image[IDX3(pz, y, x, is)]) / (2.0f * h);     \
          float gm = sqrtf(gx * gx + gy * gy + gz * gz);                                                     \
          float val;                                                                                         \
          if (gm == 0.0f) val = 0.0f;                                                                        \
          else {                                                                                             \
            float d = gx * dir[0] + gy * dir[1] + gz * dir[2];                                               \
            float sf = powf(1.0f - fabsf(d / gm), n);                                                        \
            val = gm * sf;                                                                                   \
          }                                                                                                  \
          T o;                                                                                               \
          if (!CAST(val, &o)) { bad = 1; o = 0; }                                                            \
          tmp[(z * sy + y) * sx + x] = o;                                                                    \
        }                                                                                                    \
    return bad ? 3 : 0;                                                                                      \
  }
#define WRAP_I16(a, b) ((int16_t)((int)(a) - (int)(b)))
#define WRAP_U8(a, b) ((uint8_t)((int)(a) - (int)(b)))
#define DIFF_F64(a, b) ((a) - (b))
DEF_FCM(orc_fcm_i16, int16_t, WRAP_I16, cast_f32_i16)
DEF_FCM(orc_fcm_u8, uint8_t, WRAP_U8, cast_f32_u8)
DEF_FCM(orc_fcm_f64, double, DIFF_F64, cast_f32_f64)

/* ------------------------------------------------------------------------- */
/* marching cubes — canonical restatement of the per-piece contour step of      */
/* invesalius/data/surface_process.py:71-201 (vtkContourFilter on a padded,     */
/* y-flipped vtkImageData; geometry from converters.py:34-101).                 */
/*                                                                              */
/* PARITY UNPINNED: VTK is neither installed nor vendored and the reference has */
/* no golden mesh. The canonical order defined here (and in DESIGN.md) is:      */
/*   inside(p)   <=> (double)S[p] >= iso                                        */
/*   vertices    voxels in raveled order; per voxel the crossing edges towards   */
/*               +x, +y, +z in that order; t = (iso - s0)/(s1 - s0) in float32   */
/*   position    x = ((i + ox) [+ t]) * sx ; y = -(((j + oy) [+ t]) * sy) when   */
/*               flip_y ; z = ((k + oz) [+ t]) * sz   (all float32, no FMA)       */
/*   triangles   cells in raveled order, table order within a cell (the table    */
/*               is generated by tools/gen_mc_tables.py); winding reversed when  */
/*               flip_y so normals keep pointing from inside to outside          */
/* ------------------------------------------------------------------------- */
#include "../invesalius3_b200/csrc/mc_tables.h"

static inline double mc_val(const void* vol, int dtype, int64_t i) {
  return dtype == 0 ? (double)((const int16_t*)vol)[i] : (double)((const uint8_t*)vol)[i];
}

/* Pass `verts`/`tris` NULL to only count. Returns 0; counts in *nv, *nt. */
int orc_marching_cubes(const void* vol, int dtype, int64_t nz, int64_t ny, int64_t nx, double iso, float sx, float sy,
                       float sz, int64_t ox, int64_t oy, int64_t oz, int flip_y, float* verts, int64_t* tris,
                       int64_t* nv_out, int64_t* nt_out) {
  int64_t n = nz * ny * nx;
  int64_t* voff = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
  uint8_t* in = (uint8_t*)malloc((size_t)n);
  for (int64_t i = 0; i < n; ++i) in[i] = mc_val(vol, dtype, i) >= iso;
  int64_t nv = 0;
  const int64_t step[3] = {1, nx, nx * ny};
  for (int64_t z = 0; z < nz; ++z)
    for (int64_t y = 0; y < ny; ++y)
      for (int64_t x = 0; x < nx; ++x) {
        int64_t p = (z * ny + y) * nx + x;
        voff[p] = nv;
        int64_t idx[3] = {x, y, z};
        const int64_t lim[3] = {nx, ny, nz};
        for (int a = 0; a < 3; ++a) {
          if (idx[a] + 1 >= lim[a]) continue;
          int64_t q = p + step[a];
          if (in[p] == in[q]) continue;
          if (verts) {
            float s0 = (float)mc_val(vol, dtype, p), s1 = (float)mc_val(vol, dtype, q);
            float t = ((float)iso - s0) / (s1 - s0);
            float fx = (float)(x + ox), fy = (float)(y + oy), fz = (float)(z + oz);
            if (a == 0) fx = fx + t; else if (a == 1) fy = fy + t; else fz = fz + t;
            float px = fx * sx, py = fy * sy, pz = fz * sz;
            if (flip_y) py = -py;
            verts[3 * nv] = px; verts[3 * nv + 1] = py; verts[3 * nv + 2] = pz;
          }
          ++nv;
        }
      }
  voff[n] = nv;
  int64_t nt = 0;
  for (int64_t z = 0; z + 1 < nz; ++z)
    for (int64_t y = 0; y + 1 < ny; ++y)
      for (int64_t x = 0; x + 1 < nx; ++x) {
        int64_t p = (z * ny + y) * nx + x;
        int c = 0;
        for (int k = 0; k < 8; ++k) c |= in[p + (k & 1) + ((k >> 1) & 1) * nx + ((k >> 2) & 1) * nx * ny] << k;
        int ntri = B2V_MC_NTRI[c];
        for (int t = 0; t < ntri; ++t) {
          if (tris) {
            int64_t id[3];
            for (int m = 0; m < 3; ++m) {
              int e = B2V_MC_TRI[c][3 * t + m];
              int a = e >> 2, j = e & 3, cu = j & 1, cv = j >> 1;
              int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
              int64_t off[3] = {0, 0, 0};
              off[u] = cu; off[v] = cv;
              int64_t q = p + off[0] + off[1] * nx + off[2] * nx * ny;
              /* rank of axis a among the owner's crossing edges (x, y, z order) */
              int64_t qi[3] = {x + off[0], y + off[1], z + off[2]};
              const int64_t lim[3] = {nx, ny, nz};
              int rank = 0;
              for (int b = 0; b < a; ++b)
                if (qi[b] + 1 < lim[b] && in[q] != in[q + step[b]]) ++rank;
              id[m] = voff[q] + rank;
            }
            tris[3 * nt] = id[0];
            tris[3 * nt + 1] = flip_y ? id[2] : id[1];
            tris[3 * nt + 2] = flip_y ? id[1] : id[2];
          }
          ++nt;
        }
      }
  free(voff);
  free(in);
  *nv_out = nv;
  *nt_out = nt;
  return 0;
}
