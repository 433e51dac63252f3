// HBM-bound kernels of the HESIC forward path: homography warp, entropy-model quantise/likelihood,
// global max + mixture softmax, bilinear upsample, layout conversion, symbol/index preparation and
// the rate-distortion partial sums.  All sm_100a, all asynchronous on the caller's stream.
#include "common.cuh"

namespace hesic {

__device__ __forceinline__ void unflatten(const TView &t, size_t i, int &b, int &c, int &y, int &x) {
  // walk in the tensor's own memory order so that consecutive threads touch consecutive addresses
  if (t.fmt == HESIC_FMT_NCHW_F32) {
    x = i % t.W; i /= t.W; y = i % t.H; i /= t.H; c = i % t.C; b = i / t.C;
  } else {
    c = i % t.C; i /= t.C; x = i % t.W; i /= t.W; y = i % t.H; b = i / t.H;
  }
}

__device__ __forceinline__ void block_add_double(double v, double *acc) {
  v = warp_sum(v);
  __shared__ double part[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) part[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    double s = lane < nw ? part[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(acc, s);
  }
}

// ---------------------------------------------------------------------------------------------
// kornia.warp_perspective.  Thread = one destination pixel, all channels.  The 3x3 chain
// N_dst * M * N_src^-1 and its inverse are evaluated once per block, and the per-pixel sampling
// coordinate (normalised grid -> S*g -> divide -> grid_sample unnormalise) per thread, both in fp64;
// the bilinear blend itself is fp32 like grid_sample's.
__global__ void __launch_bounds__(256) warp_kernel(const TView src, const float *__restrict__ Mx, const TView dst,
                                                  int align_corners) {
  __shared__ double S[9];
  const int b = blockIdx.z;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const float *m = Mx + b * 9;
    double h = src.H, w = src.W, ho = dst.H, wo = dst.W;
    // A = M * N_src^-1
    double nsi[9] = {(w - 1) / 2, 0, (w - 1) / 2, 0, (h - 1) / 2, (h - 1) / 2, 0, 0, 1};
    double nd[9] = {2 / (wo - 1), 0, -1, 0, 2 / (ho - 1), -1, 0, 0, 1};
    double A[9], T[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (double)m[i * 3 + k] * nsi[k * 3 + j];
        A[i * 3 + j] = s;
      }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += nd[i * 3 + k] * A[k * 3 + j];
        T[i * 3 + j] = s;
      }
    double c00 = T[4] * T[8] - T[5] * T[7], c01 = T[5] * T[6] - T[3] * T[8], c02 = T[3] * T[7] - T[4] * T[6];
    double det = T[0] * c00 + T[1] * c01 + T[2] * c02;
    double id = 1.0 / det;
    S[0] = (c00 * id); S[1] = ((T[2] * T[7] - T[1] * T[8]) * id); S[2] = ((T[1] * T[5] - T[2] * T[4]) * id);
    S[3] = (c01 * id); S[4] = ((T[0] * T[8] - T[2] * T[6]) * id); S[5] = ((T[2] * T[3] - T[0] * T[5]) * id);
    S[6] = (c02 * id); S[7] = ((T[1] * T[6] - T[0] * T[7]) * id); S[8] = ((T[0] * T[4] - T[1] * T[3]) * id);
  }
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.W || y >= dst.H) return;
  // Sampling coordinates in fp64: at 512x512 the fp32 chain ((u+1)/2)*(W-1) alone loses ~2e-4 px,
  // which is visible at the 1e-4 level on textured images; the fp64 evaluation is exact to ~1e-12 px.
  double gx = dst.W > 1 ? -1.0 + 2.0 * x / (double)(dst.W - 1) : -1.0;
  double gy = dst.H > 1 ? -1.0 + 2.0 * y / (double)(dst.H - 1) : -1.0;
  double u = gx * S[0] + gy * S[1] + S[2];
  double v = gx * S[3] + gy * S[4] + S[5];
  double z = gx * S[6] + gy * S[7] + S[8];
  double sc = fabs(z) > 1e-8 ? 1.0 / z : 1.0;
  u *= sc; v *= sc;
  double ix, iy;
  if (align_corners) {
    ix = ((u + 1.0) / 2.0) * (double)(src.W - 1);
    iy = ((v + 1.0) / 2.0) * (double)(src.H - 1);
  } else {
    ix = ((u + 1.0) * (double)src.W - 1.0) / 2.0;
    iy = ((v + 1.0) * (double)src.H - 1.0) / 2.0;
  }
  bool finite = isfinite(ix) && isfinite(iy) && fabs(ix) < 1e9 && fabs(iy) < 1e9;
  double fx = finite ? floor(ix) : 0.0, fy = finite ? floor(iy) : 0.0;
  int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  float ax = (float)(ix - fx), ay = (float)(iy - fy);
  float wnw = (1.f - ax) * (1.f - ay), wne = ax * (1.f - ay);
  float wsw = (1.f - ax) * ay, wse = ax * ay;
  bool vx0 = x0 >= 0 && x0 < src.W, vx1 = x1 >= 0 && x1 < src.W;
  bool vy0 = y0 >= 0 && y0 < src.H, vy1 = y1 >= 0 && y1 < src.H;
  for (int c = 0; c < src.C; ++c) {
    float o = 0.f;
    if (finite) {
      if (vy0 && vx0) o += tload(src, b, c, y0, x0) * wnw;
      if (vy0 && vx1) o += tload(src, b, c, y0, x1) * wne;
      if (vy1 && vx0) o += tload(src, b, c, y1, x0) * wsw;
      if (vy1 && vx1) o += tload(src, b, c, y1, x1) * wse;
    }
    tstore(dst, b, c, y, x, o);
  }
}

// ---------------------------------------------------------------------------------------------
// EntropyBottleneck forward (eval).  params: 60 floats per channel, see hesic_b200.h.
__device__ __forceinline__ float eb_logits(const float *__restrict__ p, float v) {
  float h[3], g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t = p[j] * v + p[3 + j];
    h[j] = t + p[6 + j] * tanhf(t);
  }
  const float *q = p + 9;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float t = q[j * 3 + 0] * h[0];
      t += q[j * 3 + 1] * h[1];
      t += q[j * 3 + 2] * h[2];
      t += q[9 + j];
      g[j] = t + q[12 + j] * tanhf(t);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) h[j] = g[j];
    q += 15;
  }
  float t = q[0] * h[0];
  t += q[1] * h[1];
  t += q[2] * h[2];
  return t + q[3];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) eb_kernel(const TView z, const float *__restrict__ params, float bound,
                                                const TView z_hat, const TView lik, double *log2_sum, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  double lg = 0.0;
  if (i < n) {
    int b, c, y, x;
    unflatten(z, i, b, c, y, x);
    const float *p = params + (size_t)c * HESIC_EB_PARAMS_PER_CHANNEL;
    float med = p[58];
    float q = rintf(tload(z, b, c, y, x) - med) + med;
    float lower = eb_logits(p, q - 0.5f), upper = eb_logits(p, q + 0.5f);
    float t = lower + upper;
    float sg = t > 0.f ? -1.f : (t < 0.f ? 1.f : 0.f);
    float l = fabsf(sigmoidf_(sg * upper) - sigmoidf_(sg * lower));
    if (bound > 0.f) l = fmaxf(l, bound);
    if (z_hat.p0) tstore(z_hat, b, c, y, x, q);
    if (lik.p0) tstore(lik, b, c, y, x, l);
    lg = (double)log2f(l);
  }
  if (log2_sum) block_add_double(lg, log2_sum);
}

__global__ void eb_pack_kernel(const float *m0, const float *m1, const float *m2, const float *m3, const float *m4,
                               const float *b0, const float *b1, const float *b2, const float *b3, const float *b4,
                               const float *f0, const float *f1, const float *f2, const float *f3,
                               const float *quantiles, int C, float *out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float *p = out + (size_t)c * HESIC_EB_PARAMS_PER_CHANNEL;
  auto sp = [](float v) { return v > 20.f ? v : log1pf(expf(v)); };  // F.softplus (beta=1, threshold=20)
  for (int j = 0; j < 3; ++j) { p[j] = sp(m0[c * 3 + j]); p[3 + j] = b0[c * 3 + j]; p[6 + j] = tanhf(f0[c * 3 + j]); }
  const float *ms[3] = {m1, m2, m3}, *bs[3] = {b1, b2, b3}, *fs[3] = {f1, f2, f3};
  float *q = p + 9;
  for (int l = 0; l < 3; ++l) {
    for (int j = 0; j < 9; ++j) q[j] = sp(ms[l][c * 9 + j]);
    for (int j = 0; j < 3; ++j) { q[9 + j] = bs[l][c * 3 + j]; q[12 + j] = tanhf(fs[l][c * 3 + j]); }
    q += 15;
  }
  for (int j = 0; j < 3; ++j) q[j] = sp(m4[c * 3 + j]);
  q[3] = b4[c];
  p[58] = quantiles[c * 3 + 1];
  p[59] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// GaussianMixtureConditional / GaussianConditional forward (eval)
__device__ __forceinline__ float std_cumulative(float v) { return 0.5f * erfcf(-0.70710678118654752440f * v); }

__global__ void __launch_bounds__(256) gaussian_kernel(const TView y, const TView scales, const TView means,
                                                      const float *__restrict__ weights, int K, int mixture,
                                                      float scale_bound, float lik_bound, const TView y_hat,
                                                      const TView lik, double *log2_sum, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  double lg = 0.0;
  if (i < n) {
    int b, c, yy, xx;
    unflatten(y, i, b, c, yy, xx);
    const int M = y.C;
    float v = tload(y, b, c, yy, xx);
    float q, l;
    if (mixture) {
      q = rintf(v);
      l = 0.f;
      for (int k = 0; k < K; ++k) {
        int ck = k * M + c;
        float d = fabsf(q - tload(means, b, ck, yy, xx));
        float s = fmaxf(tload(scales, b, ck, yy, xx), scale_bound);
        float term = (std_cumulative((0.5f - d) / s) - std_cumulative((-0.5f - d) / s)) * weights[(size_t)b * K * M + ck];
        l = k == 0 ? term : l + term;
      }
    } else {
      float d;
      if (means.p0) {
        float mu = tload(means, b, c, yy, xx);
        float r = rintf(v - mu);
        q = r + mu;
        d = fabsf(q - mu);
      } else {
        q = rintf(v);
        d = fabsf(q);
      }
      float s = fmaxf(tload(scales, b, c, yy, xx), scale_bound);
      l = std_cumulative((0.5f - d) / s) - std_cumulative((-0.5f - d) / s);
    }
    if (lik_bound > 0.f) l = fmaxf(l, lik_bound);
    if (y_hat.p0) tstore(y_hat, b, c, yy, xx, q);
    if (lik.p0) tstore(lik, b, c, yy, xx, l);
    lg = (double)log2f(l);
  }
  if (log2_sum) block_add_double(lg, log2_sum);
}

// ---------------------------------------------------------------------------------------------
// spatial_pool2d: global max per (b, c).
// NHWC: block = 32 channels x 8 pixel lanes (a warp reads 32 consecutive channels of one pixel).
__global__ void __launch_bounds__(256) spatial_max_nhwc_kernel(const TView x, float *__restrict__ out) {
  __shared__ float red[8][33];
  const int b = blockIdx.y;
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float m = -INFINITY;
  const int P = x.H * x.W;
  if (c < x.C)
    for (int p = pl; p < P; p += 8) m = fmaxf(m, tload(x, b, c, p / x.W, p % x.W));
  red[pl][cl] = m;
  __syncthreads();
  if (pl == 0 && c < x.C) {
#pragma unroll
    for (int j = 1; j < 8; ++j) m = fmaxf(m, red[j][cl]);
    out[(size_t)b * x.C + c] = m;
  }
}

// NCHW: one warp per (b, c) plane, lanes walk pixels.
__global__ void __launch_bounds__(256) spatial_max_nchw_kernel(const TView x, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int plane = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (plane >= x.B * x.C) return;
  const int b = plane / x.C, c = plane % x.C;
  const int P = x.H * x.W;
  float m = -INFINITY;
  for (int p = lane; p < P; p += 32) m = fmaxf(m, tload(x, b, c, p / x.W, p % x.W));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) out[plane] = m;
}

// LeakyReLU -> conv1x1 -> softmax over the K mixture components (newnet1.py:498-512)
__global__ void mixture_weights_kernel(const float *__restrict__ pooled, const float *__restrict__ w,
                                       const float *__restrict__ bias, int K, int M, float *__restrict__ out) {
  extern __shared__ float logit[];  // [K]
  const int b = blockIdx.y, m = blockIdx.x;
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KM = K * M;
  const int o = k * M + m;
  const float *wr = w + (size_t)o * KM;
  const float *pin = pooled + (size_t)b * KM;
  float s = 0.f;
  for (int i = lane; i < KM; i += 32) {
    float v = pin[i];
    v = v > 0.f ? v : 0.01f * v;
    s = fmaf(wr[i], v, s);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) logit[k] = s + (bias ? bias[o] : 0.f);
  __syncthreads();
  if (threadIdx.x < K) {
    float mx = -INFINITY;
    for (int j = 0; j < K; ++j) mx = fmaxf(mx, logit[j]);
    float den = 0.f;
    for (int j = 0; j < K; ++j) den += expf(logit[j] - mx);
    out[(size_t)b * KM + threadIdx.x * M + m] = expf(logit[threadIdx.x] - mx) / den;
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_kernel(const TView x, const TView y, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b, c, oy, ox;
  unflatten(y, i, b, c, oy, ox);
  // at::native area_pixel_compute_scale / upsample_bilinear2d, align_corners=True
  float sh = y.H > 1 ? (float)(x.H - 1) / (float)(y.H - 1) : 0.f;
  float sw = y.W > 1 ? (float)(x.W - 1) / (float)(y.W - 1) : 0.f;
  float fy = sh * oy, fx = sw * ox;
  int y0 = (int)fy, x0 = (int)fx;
  int yp = y0 < x.H - 1 ? 1 : 0, xp = x0 < x.W - 1 ? 1 : 0;
  float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  float v = hy * (hx * tload(x, b, c, y0, x0) + lx * tload(x, b, c, y0, x0 + xp)) +
            ly * (hx * tload(x, b, c, y0 + yp, x0) + lx * tload(x, b, c, y0 + yp, x0 + xp));
  tstore(y, b, c, oy, ox, v);
}

__global__ void __launch_bounds__(256) convert_kernel(const TView x, const TView y, int op, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b, c, yy, xx;
  unflatten(y, i, b, c, yy, xx);
  float v = tload(x, b, c, yy, xx);
  if (op == HESIC_OP_ABS) v = fabsf(v);
  else if (op == HESIC_OP_ROUND) v = rintf(v);
  tstore(y, b, c, yy, xx, v);
}

// NCHW fp32 (C <= 8) -> ROWPAD8 split planes: thread = pixel, one 16-byte store per plane (all 8
// channel slots, zeros above C), reads coalesced per source plane.
__global__ void __launch_bounds__(256) rowpad_kernel(const TView x, const TView y, int op, size_t npix) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  int xx = i % x.W; size_t r = i / x.W;
  int yy = r % x.H; int b = r / x.H;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    v[c] = c < x.C ? ((const float *)x.p0)[(((size_t)b * x.Cs + c) * x.H + yy) * x.W + xx] : 0.f;
    if (op == HESIC_OP_ABS) v[c] = fabsf(v[c]);
    else if (op == HESIC_OP_ROUND) v[c] = rintf(v[c]);
  }
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) split_bf16(v[c], hi[c], lo[c]);
  size_t o = toff(y, b, 0, yy, xx);
  *reinterpret_cast<uint4 *>((__nv_bfloat16 *)y.p0 + o) = *reinterpret_cast<const uint4 *>(hi);
  *reinterpret_cast<uint4 *>((__nv_bfloat16 *)y.p1 + o) = *reinterpret_cast<const uint4 *>(lo);
}

// ---------------------------------------------------------------------------------------------
// integer preparation for the host rANS coder (bit-exact with the reference)
__global__ void __launch_bounds__(256) symbols_kernel(const TView x, const float *__restrict__ cmeans, const TView means,
                                                     int32_t *__restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int xx = i % x.W; size_t r = i / x.W;
  int yy = r % x.H; r /= x.H;
  int c = r % x.C; int b = r / x.C;
  float v = tload(x, b, c, yy, xx);
  if (cmeans) v = v - cmeans[c];
  else if (means.p0) v = v - tload(means, b, c, yy, xx);
  out[i] = (int32_t)rintf(v);
}

__global__ void __launch_bounds__(256) indexes_channel_kernel(int C, int HW, int32_t *__restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)((i / HW) % C);
}

__global__ void __launch_bounds__(256) indexes_scale_kernel(const TView sc, const float *__restrict__ table, int nt,
                                                           float bound, int32_t *__restrict__ out, size_t n) {
  extern __shared__ float tab[];
  for (int j = threadIdx.x; j < nt; j += blockDim.x) tab[j] = table[j];
  __syncthreads();
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int xx = i % sc.W; size_t r = i / sc.W;
  int yy = r % sc.H; r /= sc.H;
  int c = r % sc.C; int b = r / sc.C;
  float s = fmaxf(tload(sc, b, c, yy, xx), bound);
  int idx = nt - 1;
  for (int j = 0; j < nt - 1; ++j) idx -= (s <= tab[j]) ? 1 : 0;
  out[i] = idx;
}

__global__ void __launch_bounds__(256) sse_kernel(const TView a, const TView b_, double *acc, size_t n) {
  double s = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int b, c, y, x;
    unflatten(a, i, b, c, y, x);
    float d = tload(a, b, c, y, x) - tload(b_, b, c, y, x);
    s += (double)d * (double)d;
  }
  block_add_double(s, acc);
}

static inline size_t numel(const hesic_tensor *t) { return (size_t)t->B * t->C * t->H * t->W; }
static inline unsigned nblk(size_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace hesic

using namespace hesic;

extern "C" int hesic_warp_perspective(const hesic_tensor *src, const float *M, const hesic_tensor *dst,
                                      int align_corners, void *stream) {
  int r;
  if ((r = check_tensor(src, "warp src")) != HESIC_OK) return r;
  if ((r = check_tensor(dst, "warp dst")) != HESIC_OK) return r;
  HESIC_REQUIRE(M != nullptr, "warp: null homography");
  HESIC_REQUIRE(src->B == dst->B && src->C == dst->C, "warp: batch/channel mismatch");
  HESIC_REQUIRE(src->p0 != dst->p0, "warp: in-place is not supported");
  if (numel(dst) == 0) return HESIC_OK;
  dim3 blk(32, 8), grid((dst->W + 31) / 32, (dst->H + 7) / 8, dst->B);
  warp_kernel<<<grid, blk, 0, as_stream(stream)>>>(view(src), M, view(dst), align_corners);
  HESIC_LAUNCHED("warp_kernel");
  return HESIC_OK;
}

extern "C" int hesic_eb_pack(const float *const *m, const float *const *b, const float *const *f, const float *quantiles,
                             int C, float *params_out, void *stream) {
  HESIC_REQUIRE(m && b && f && quantiles && params_out && C > 0, "eb_pack: null argument");
  eb_pack_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(m[0], m[1], m[2], m[3], m[4], b[0], b[1], b[2], b[3], b[4],
                                                                f[0], f[1], f[2], f[3], quantiles, C, params_out);
  HESIC_LAUNCHED("eb_pack_kernel");
  return HESIC_OK;
}

static hesic_tensor null_tensor() {
  hesic_tensor t;
  t.p0 = nullptr; t.p1 = nullptr; t.fmt = 0; t.B = t.C = t.H = t.W = t.Cs = 0;
  return t;
}

extern "C" int hesic_entropy_bottleneck(const hesic_tensor *z, const float *params, float likelihood_bound,
                                        const hesic_tensor *z_hat, const hesic_tensor *lik, double *log2_sum,
                                        void *stream) {
  int r;
  if ((r = check_tensor(z, "eb input")) != HESIC_OK) return r;
  HESIC_REQUIRE(params != nullptr, "eb: null params");
  hesic_tensor nt = null_tensor();
  const hesic_tensor *zh = z_hat ? z_hat : &nt, *lk = lik ? lik : &nt;
  if (zh->p0) HESIC_REQUIRE(same_shape(z, zh), "eb: z_hat shape mismatch");
  if (lk->p0) HESIC_REQUIRE(same_shape(z, lk), "eb: likelihood shape mismatch");
  size_t n = numel(z);
  if (n == 0) return HESIC_OK;
  eb_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(z), params, likelihood_bound, view(zh), view(lk), log2_sum, n);
  HESIC_LAUNCHED("eb_kernel");
  return HESIC_OK;
}

extern "C" int hesic_gaussian_conditional(const hesic_tensor *y, const hesic_tensor *scales, const hesic_tensor *means,
                                          const float *weights, int K, int mixture, float scale_bound,
                                          float likelihood_bound, const hesic_tensor *y_hat, const hesic_tensor *lik,
                                          double *log2_sum, void *stream) {
  int r;
  if ((r = check_tensor(y, "gaussian input")) != HESIC_OK) return r;
  if ((r = check_tensor(scales, "gaussian scales")) != HESIC_OK) return r;
  hesic_tensor nt = null_tensor();
  const hesic_tensor *mu = means ? means : &nt, *yh = y_hat ? y_hat : &nt, *lk = lik ? lik : &nt;
  HESIC_REQUIRE(K >= 1, "gaussian: K must be >= 1");
  if (mixture) {
    HESIC_REQUIRE(weights != nullptr && mu->p0 != nullptr, "gaussian mixture: means and weights are required");
    HESIC_REQUIRE(scales->C == K * y->C && mu->C == K * y->C, "gaussian mixture: scales/means need K*M channels");
  } else {
    HESIC_REQUIRE(K == 1 && scales->C == y->C, "gaussian: scales channel mismatch");
    if (mu->p0) HESIC_REQUIRE(mu->C == y->C, "gaussian: means channel mismatch");
  }
  HESIC_REQUIRE(scales->B == y->B && scales->H == y->H && scales->W == y->W, "gaussian: scales shape mismatch");
  if (mu->p0) HESIC_REQUIRE(mu->B == y->B && mu->H == y->H && mu->W == y->W, "gaussian: means shape mismatch");
  if (yh->p0) HESIC_REQUIRE(same_shape(y, yh), "gaussian: y_hat shape mismatch");
  if (lk->p0) HESIC_REQUIRE(same_shape(y, lk), "gaussian: likelihood shape mismatch");
  size_t n = numel(y);
  if (n == 0) return HESIC_OK;
  gaussian_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(y), view(scales), view(mu), weights, K, mixture,
                                                         scale_bound, likelihood_bound, view(yh), view(lk), log2_sum, n);
  HESIC_LAUNCHED("gaussian_kernel");
  return HESIC_OK;
}

extern "C" int hesic_spatial_max(const hesic_tensor *x, float *out_max, void *stream) {
  int r;
  if ((r = check_tensor(x, "spatial_max input")) != HESIC_OK) return r;
  HESIC_REQUIRE(out_max != nullptr, "spatial_max: null output");
  HESIC_REQUIRE(x->H * x->W > 0, "spatial_max: empty spatial extent");
  if (x->B * x->C == 0) return HESIC_OK;
  if (x->fmt == HESIC_FMT_NCHW_F32) {
    spatial_max_nchw_kernel<<<(x->B * x->C + 7) / 8, 256, 0, as_stream(stream)>>>(view(x), out_max);
  } else {
    dim3 grid((x->C + 31) / 32, x->B);
    spatial_max_nhwc_kernel<<<grid, 256, 0, as_stream(stream)>>>(view(x), out_max);
  }
  HESIC_LAUNCHED("spatial_max_kernel");
  return HESIC_OK;
}

extern "C" int hesic_mixture_weights(const float *pooled, const float *w1x1, const float *bias, int B, int K, int M,
                                     float *out, void *stream) {
  HESIC_REQUIRE(pooled && w1x1 && out, "mixture_weights: null argument");
  HESIC_REQUIRE(K >= 1 && K <= 32 && M >= 1 && B >= 0, "mixture_weights: bad sizes");
  if (B == 0) return HESIC_OK;
  mixture_weights_kernel<<<dim3(M, B), K * 32, K * sizeof(float), as_stream(stream)>>>(pooled, w1x1, bias, K, M, out);
  HESIC_LAUNCHED("mixture_weights_kernel");
  return HESIC_OK;
}

extern "C" int hesic_upsample_bilinear(const hesic_tensor *x, const hesic_tensor *y, int scale, void *stream) {
  int r;
  if ((r = check_tensor(x, "upsample input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "upsample output")) != HESIC_OK) return r;
  HESIC_REQUIRE(scale >= 1 && y->H == x->H * scale && y->W == x->W * scale && y->B == x->B && y->C == x->C,
                "upsample: output shape mismatch");
  size_t n = numel(y);
  if (n == 0) return HESIC_OK;
  upsample_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(x), view(y), n);
  HESIC_LAUNCHED("upsample_kernel");
  return HESIC_OK;
}

extern "C" int hesic_convert(const hesic_tensor *x, const hesic_tensor *y, int op, void *stream) {
  int r;
  if ((r = check_tensor(x, "convert input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "convert output")) != HESIC_OK) return r;
  HESIC_REQUIRE(same_shape(x, y), "convert: shape mismatch");
  HESIC_REQUIRE(op >= 0 && op <= 2, "convert: bad op");
  size_t n = numel(y);
  if (n == 0) return HESIC_OK;
  if (y->fmt == HESIC_FMT_ROWPAD8_SPLIT && x->fmt == HESIC_FMT_NCHW_F32 && ((uintptr_t)y->p0 & 15) == 0 &&
      ((uintptr_t)y->p1 & 15) == 0) {
    // whole-pixel fast path (the view starts at channel slot 0 and owns all 8 slots)
    size_t npix = (size_t)y->B * y->H * y->W;
    rowpad_kernel<<<nblk(npix), 256, 0, as_stream(stream)>>>(view(x), view(y), op, npix);
    HESIC_LAUNCHED("rowpad_kernel");
    return HESIC_OK;
  }
  convert_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(x), view(y), op, n);
  HESIC_LAUNCHED("convert_kernel");
  return HESIC_OK;
}

extern "C" int hesic_prepare_symbols(const hesic_tensor *x, const float *channel_means, const hesic_tensor *means,
                                     int32_t *out_symbols, void *stream) {
  int r;
  if ((r = check_tensor(x, "symbols input")) != HESIC_OK) return r;
  HESIC_REQUIRE(out_symbols != nullptr || numel(x) == 0, "symbols: null output");
  hesic_tensor nt = null_tensor();
  const hesic_tensor *mu = means ? means : &nt;
  if (mu->p0) HESIC_REQUIRE(same_shape(x, mu), "symbols: means shape mismatch");
  size_t n = numel(x);
  if (n == 0) return HESIC_OK;
  symbols_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(x), channel_means, view(mu), out_symbols, n);
  HESIC_LAUNCHED("symbols_kernel");
  return HESIC_OK;
}

extern "C" int hesic_build_indexes_channel(int B, int C, int H, int W, int32_t *out, void *stream) {
  HESIC_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0, "indexes: negative size");
  size_t n = (size_t)B * C * H * W;
  if (n == 0) return HESIC_OK;
  HESIC_REQUIRE(out != nullptr, "indexes: null output");
  indexes_channel_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(C, H * W, out, n);
  HESIC_LAUNCHED("indexes_channel_kernel");
  return HESIC_OK;
}

extern "C" int hesic_build_indexes_scale(const hesic_tensor *scales, const float *table, int n_table, float scale_bound,
                                         int32_t *out, void *stream) {
  int r;
  if ((r = check_tensor(scales, "indexes scales")) != HESIC_OK) return r;
  HESIC_REQUIRE(table && n_table >= 1 && n_table <= 4096, "indexes: bad scale table");
  size_t n = numel(scales);
  if (n == 0) return HESIC_OK;
  HESIC_REQUIRE(out != nullptr, "indexes: null output");
  indexes_scale_kernel<<<nblk(n), 256, n_table * sizeof(float), as_stream(stream)>>>(view(scales), table, n_table,
                                                                                    scale_bound, out, n);
  HESIC_LAUNCHED("indexes_scale_kernel");
  return HESIC_OK;
}

extern "C" int hesic_sum_squared_error(const hesic_tensor *a, const hesic_tensor *b, double *acc, void *stream) {
  int r;
  if ((r = check_tensor(a, "sse a")) != HESIC_OK) return r;
  if ((r = check_tensor(b, "sse b")) != HESIC_OK) return r;
  HESIC_REQUIRE(same_shape(a, b) && acc, "sse: shape mismatch or null accumulator");
  size_t n = numel(a);
  if (n == 0) return HESIC_OK;
  unsigned blocks = nblk(n) < 148 * 8 ? nblk(n) : 148 * 8;
  sse_kernel<<<blocks, 256, 0, as_stream(stream)>>>(view(a), view(b), acc, n);
  HESIC_LAUNCHED("sse_kernel");
  return HESIC_OK;
}
