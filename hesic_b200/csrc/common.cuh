// Shared device/host helpers for the hesic_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <string>

#include "../../include/hesic_b200.h"

namespace hesic {

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int cuda_fail(cudaError_t e, const char *what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return HESIC_E_CUDA;
}

#define HESIC_CUDA(expr)                                   \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return hesic::cuda_fail(_e, #expr); \
  } while (0)

#define HESIC_REQUIRE(cond, ...)   \
  do {                             \
    if (!(cond)) {                 \
      hesic::set_error(__VA_ARGS__); \
      return HESIC_E_INVALID;      \
    }                              \
  } while (0)

// call after every kernel launch
inline int launched(const char *name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, name);
  return HESIC_OK;
}
#define HESIC_LAUNCHED(name)           \
  do {                                 \
    int _r = hesic::launched(name);    \
    if (_r != HESIC_OK) return _r;     \
  } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline int current_device() {
  int d = 0;
  return cudaGetDevice(&d) == cudaSuccess ? d : -1;
}
// cudaFuncSetAttribute and friends are per device: first() is true once per device of this process (one process per
// GPU is the normal deployment, but a model moved to another device of the same process must still launch)
struct PerDeviceOnce {
  unsigned long long done = 0;
  bool first() {
    const int d = current_device();
    if (d < 0 || d > 63) return true;
    if ((done >> d) & 1ull) return false;
    done |= 1ull << d;
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// Device view of a hesic_tensor
struct TView {
  const void *p0;
  const void *p1;
  int fmt, B, C, H, W, Cs;
};

inline TView view(const hesic_tensor *t) {
  TView v;
  v.p0 = t->p0; v.p1 = t->p1; v.fmt = t->fmt; v.B = t->B; v.C = t->C; v.H = t->H; v.W = t->W;
  v.Cs = t->Cs > 0 ? t->Cs : t->C;
  return v;
}

inline bool same_shape(const hesic_tensor *a, const hesic_tensor *b) {
  return a->B == b->B && a->C == b->C && a->H == b->H && a->W == b->W;
}

inline int check_tensor(const hesic_tensor *t, const char *name, bool allow_null_data = false) {
  HESIC_REQUIRE(t != nullptr, "%s: null tensor descriptor", name);
  HESIC_REQUIRE(t->fmt >= 0 && t->fmt <= 4, "%s: bad format %d", name, t->fmt);
  HESIC_REQUIRE(t->B >= 0 && t->C >= 0 && t->H >= 0 && t->W >= 0, "%s: negative size", name);
  HESIC_REQUIRE(t->Cs == 0 || t->Cs >= t->C, "%s: Cs < C", name);
  if (!allow_null_data && (int64_t)t->B * t->C * t->H * t->W > 0) {
    HESIC_REQUIRE(t->p0 != nullptr, "%s: null data pointer", name);
    if (t->fmt == HESIC_FMT_NHWC_SPLIT || t->fmt == HESIC_FMT_ROWPAD8_SPLIT)
      HESIC_REQUIRE(t->p1 != nullptr, "%s: null lo plane", name);
    if (t->fmt == HESIC_FMT_ROWPAD8_SPLIT)
      HESIC_REQUIRE((t->Cs == 8 || (t->Cs == 4 && (t->H & 1) == 0)) && t->C <= t->Cs,
                    "%s: ROWPAD needs Cs == 8 or 4 channel slots (4: even height)", name);
  }
  return HESIC_OK;
}

// ROWPAD element offset of channel slot 0 of pixel (y, x).  Cs = 8: [B][H+4][W+8][8].  Cs = 4: rows are
// interleaved in pairs, [B][(H+4)/2][W+8][2][4], so that 8 pixels x 2 rows x 4 slots are 64 contiguous elements.
__device__ __forceinline__ size_t rowpad_off(const TView &t, int b, int y, int x) {
  const int Wp = t.W + HESIC_ROWPAD_X, Hp = t.H + HESIC_ROWPAD_Y;
  if (t.Cs == 8) return (((size_t)b * Hp + y + 2) * Wp + x + 2) * 8;
  return (((size_t)b * (Hp / 2) + ((y + 2) >> 1)) * Wp + x + 2) * 8 + ((y + 2) & 1) * 4;
}

__device__ __forceinline__ size_t toff(const TView &t, int b, int c, int y, int x) {
  if (t.fmt == HESIC_FMT_NCHW_F32) return (((size_t)b * t.Cs + c) * t.H + y) * t.W + x;
  if (t.fmt == HESIC_FMT_ROWPAD8_SPLIT) return rowpad_off(t, b, y, x) + c;
  if (t.fmt == HESIC_FMT_NHWC_HILO) return (((size_t)b * t.H + y) * t.W + x) * 2 * t.Cs + c;   // hi; lo is Cs further
  return (((size_t)b * t.H + y) * t.W + x) * t.Cs + c;
}

__device__ __forceinline__ float tload(const TView &t, int b, int c, int y, int x) {
  size_t o = toff(t, b, c, y, x);
  if (t.fmt == HESIC_FMT_NHWC_HILO)
    return __bfloat162float(((const __nv_bfloat16 *)t.p0)[o]) + __bfloat162float(((const __nv_bfloat16 *)t.p0)[o + t.Cs]);
  if (t.fmt >= HESIC_FMT_NHWC_SPLIT) {
    return __bfloat162float(((const __nv_bfloat16 *)t.p0)[o]) + __bfloat162float(((const __nv_bfloat16 *)t.p1)[o]);
  }
  return ((const float *)t.p0)[o];
}

// value = hi + lo with hi = rn_bf16(v), lo = rn_bf16(v - hi): |v - hi - lo| <= 2^-18 |v|
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ void tstore(const TView &t, int b, int c, int y, int x, float v) {
  size_t o = toff(t, b, c, y, x);
  if (t.fmt >= HESIC_FMT_NHWC_SPLIT) {
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    ((__nv_bfloat16 *)t.p0)[o] = hi;
    if (t.fmt == HESIC_FMT_NHWC_HILO) ((__nv_bfloat16 *)t.p0)[o + t.Cs] = lo;
    else ((__nv_bfloat16 *)t.p1)[o] = lo;
  } else {
    ((float *)t.p0)[o] = v;
  }
}

// One whole pixel of a ROWPAD tensor (all Cs = 8 or 4 channel slots, zeros above n): one 16- or 8-byte store
// per plane.  v holds n <= Cs values.
template <int N>
__device__ __forceinline__ void store_rowpad_pixel(const TView &t, int b, int y, int x, const float (&v)[N]) {
  const size_t o = rowpad_off(t, b, y, x);
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (c < N) split_bf16(v[c < N ? c : 0], hi[c], lo[c]);
    else { hi[c] = __float2bfloat16_rn(0.f); lo[c] = hi[c]; }
  }
  if (t.Cs == 8) {
    *reinterpret_cast<uint4 *>((__nv_bfloat16 *)t.p0 + o) = *reinterpret_cast<const uint4 *>(hi);
    *reinterpret_cast<uint4 *>((__nv_bfloat16 *)t.p1 + o) = *reinterpret_cast<const uint4 *>(lo);
  } else {
    *reinterpret_cast<uint2 *>((__nv_bfloat16 *)t.p0 + o) = *reinterpret_cast<const uint2 *>(hi);
    *reinterpret_cast<uint2 *>((__nv_bfloat16 *)t.p1 + o) = *reinterpret_cast<const uint2 *>(lo);
  }
}

// Asynchronous global -> shared copies (LDGSTS): all of a block's staging loads are in flight at once instead
// of one load->store round trip per element.  pred == false writes zeros (src-size 0) without touching src.
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = pred ? BYTES : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(d), "l"(gsrc), "n"(BYTES), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == HESIC_ACT_RELU) return fmaxf(v, 0.f);
  if (act == HESIC_ACT_LEAKY_RELU) return v > 0.f ? v : 0.01f * v;
  return v;
}

// NonNegativeParametrizer.forward (compressai/ops/parametrizers.py:41-44)
__device__ __forceinline__ float nonneg_reparam(float p, float minimum) {
  const float pedestal = 1.4551915228366852e-11f;  // (2^-18)^2
  float bound = sqrtf(minimum + pedestal);
  float o = fmaxf(p, bound);
  return o * o - pedestal;
}

// Diagnostic switches (A/B measurements of kernel variants and fallbacks) exist only in a library
// built with -DHESIC_DIAG (`python -m hesic_b200.build --diag`): the production build reads no environment variable.
#ifdef HESIC_DIAG
inline const char *diag_env(const char *name) { return getenv(name); }
#else
inline const char *diag_env(const char *) { return nullptr; }
#endif

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hesic
