// Operators that only the DSIC variant (ywz/DSIC/mynet6_plus.py) adds to the forward path: GroupNorm (+ReLU),
// softmax over the disparity channels of a cost volume, and dense_warp (disparity-weighted horizontal
// shift-sum).  All NCHW fp32 like the reference's tensors at those points; HBM-bound except dense_warp,
// which is a 32-tap 1-D correlation staged through shared memory.
#include <algorithm>

#include "common.cuh"

namespace hesic {

// ---------------------------------------------------------------------------------------------
// nn.GroupNorm: statistics over (C/G, H, W) per (b, g) -- contiguous in NCHW -- accumulated in fp64.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float *__restrict__ x, size_t group_elems, size_t group_stride,
                                                      double *__restrict__ stats) {
  const int bg = blockIdx.y;
  const float *p = x + (size_t)bg * group_stride;
  double s = 0.0, ss = 0.0;
  const size_t n4 = ((((uintptr_t)p) & 15u) == 0) ? group_elems / 4 : 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p) + i);
    s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
    ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < group_elems; i += (size_t)gridDim.x * blockDim.x) {
    const double v = p[i];
    s += v; ss += v * v;
  }
  s = warp_sum(s); ss = warp_sum(ss);
  __shared__ double ps[8], pss[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { ps[w] = s; pss[w] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; ++i) { a += ps[i]; b += pss[i]; }
    atomicAdd(&stats[2 * bg], a);
    atomicAdd(&stats[2 * bg + 1], b);
  }
}

// y = (x - mean) * rstd * weight[c] + bias[c]  (+ ReLU).  x: channels [0, C) of a buffer with Cs channels.
__global__ void __launch_bounds__(256) gn_apply_kernel(const TView x, const TView y, int G, const double *__restrict__ stats,
                                                      const float *__restrict__ weight, const float *__restrict__ bias,
                                                      float eps, int relu, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t HW = (size_t)x.H * x.W;
  const size_t pos = i % HW;
  const int c = (int)((i / HW) % x.C), b = (int)(i / (HW * x.C));
  const int cpg = x.C / G, g = c / cpg;
  const double cnt = (double)cpg * (double)HW;
  const double mean = stats[2 * (b * G + g)] / cnt;
  const double var = fmax(stats[2 * (b * G + g) + 1] / cnt - mean * mean, 0.0);
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float v = ((const float *)x.p0)[((size_t)b * x.Cs + c) * HW + pos];
  float o = (v - (float)mean) * rstd;
  if (weight) o = o * weight[c] + (bias ? bias[c] : 0.f);
  if (relu) o = fmaxf(o, 0.f);
  ((float *)y.p0)[((size_t)b * y.Cs + c) * HW + pos] = o;
}

// ---------------------------------------------------------------------------------------------
// softmax over the channel dimension, thread = pixel (channel stride H*W: coalesced across threads)
__global__ void __launch_bounds__(256) softmax_channels_kernel(const TView x, const TView y, size_t npix) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const size_t HW = (size_t)x.H * x.W;
  const size_t pos = i % HW;
  const int b = (int)(i / HW);
  const float *xp = (const float *)x.p0 + (size_t)b * x.Cs * HW + pos;
  float *yp = (float *)y.p0 + (size_t)b * y.Cs * HW + pos;
  float m = -INFINITY;
  for (int c = 0; c < x.C; ++c) m = fmaxf(m, xp[c * HW]);
  float s = 0.f;
  for (int c = 0; c < x.C; ++c) s += expf(xp[c * HW] - m);
  const float inv = 1.f / s;
  for (int c = 0; c < x.C; ++c) yp[c * HW] = expf(xp[c * HW] - m) * inv;
}

// ---------------------------------------------------------------------------------------------
// dense_warp: out[b,c,y,x] = sum_{d < D, x + d < W} cost[b,d,y,x] * h1[b,c,y,x+d]
// Block = one image row segment of 64 pixels; 256 threads = 64 pixels x 4 channel lanes.  The pixel's D cost
// values live in registers, the feature row segment (64 + D - 1 pixels) of 32 channels at a time in shared memory.
constexpr int DW_TX = 64, DW_CB = 32, DW_MAXD = 64;

__global__ void __launch_bounds__(256) dense_warp_kernel(const TView h1, const TView cost, const TView out) {
  __shared__ float hs[DW_CB][DW_TX + DW_MAXD];
  const int D = cost.C, W = h1.W, H = h1.H;
  const int x0 = blockIdx.x * DW_TX, yy = blockIdx.y, b = blockIdx.z;
  const int px = threadIdx.x & 63, cl = threadIdx.x >> 6;
  const size_t HW = (size_t)H * W;
  const int x = x0 + px;
  float cst[DW_MAXD];
#pragma unroll
  for (int d = 0; d < DW_MAXD; ++d)
    cst[d] = (d < D && x < W && x + d < W) ? __ldg((const float *)cost.p0 + ((size_t)b * cost.Cs + d) * HW + (size_t)yy * W + x) : 0.f;
  const int span = DW_TX + D - 1;
  for (int c0 = 0; c0 < h1.C; c0 += DW_CB) {
    __syncthreads();
    for (int i = threadIdx.x; i < DW_CB * span; i += 256) {
      const int c = i / span, col = i - c * span;
      const int gx = x0 + col;
      float v = 0.f;
      if (c0 + c < h1.C && gx < W) v = __ldg((const float *)h1.p0 + ((size_t)b * h1.Cs + c0 + c) * HW + (size_t)yy * W + gx);
      hs[c][col] = v;
    }
    __syncthreads();
    if (x < W) {
#pragma unroll
      for (int j = 0; j < DW_CB / 4; ++j) {
        const int c = cl + 4 * j;
        if (c0 + c >= h1.C) break;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < DW_MAXD; ++d)
          if (d < D) acc = fmaf(cst[d], hs[c][px + d], acc);
        ((float *)out.p0)[((size_t)b * out.Cs + c0 + c) * HW + (size_t)yy * W + x] = acc;
      }
    }
  }
}

static inline size_t numel(const hesic_tensor *t) { return (size_t)t->B * t->C * t->H * t->W; }

}  // namespace hesic

using namespace hesic;

extern "C" int hesic_group_norm(const hesic_tensor *x, const hesic_tensor *y, int groups, const float *weight,
                                const float *bias, float eps, int relu, void *stream) {
  int r;
  if ((r = check_tensor(x, "group_norm input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "group_norm output")) != HESIC_OK) return r;
  HESIC_REQUIRE(x->fmt == HESIC_FMT_NCHW_F32 && y->fmt == HESIC_FMT_NCHW_F32, "group_norm: NCHW fp32 tensors required");
  HESIC_REQUIRE(same_shape(x, y), "group_norm: shape mismatch");
  HESIC_REQUIRE(groups >= 1 && x->C % groups == 0, "group_norm: %d channels are not divisible into %d groups", x->C, groups);
  const int xCs = x->Cs > 0 ? x->Cs : x->C;
  HESIC_REQUIRE(xCs == x->C, "group_norm: input must not be a channel slice");
  size_t n = numel(x);
  if (n == 0) return HESIC_OK;
  cudaStream_t st = as_stream(stream);
  const int BG = x->B * groups;
  HESIC_REQUIRE(BG <= 65535, "group_norm: too many (batch, group) pairs");
  double *stats = nullptr;
  HESIC_CUDA(cudaMallocAsync(&stats, (size_t)BG * 2 * sizeof(double), st));
  cudaMemsetAsync(stats, 0, (size_t)BG * 2 * sizeof(double), st);
  const size_t ge = (size_t)(x->C / groups) * x->H * x->W;
  unsigned bx = (unsigned)std::min<size_t>((ge / 4 + 255) / 256 + 1, 64);
  gn_stats_kernel<<<dim3(bx, BG), 256, 0, st>>>((const float *)x->p0, ge, ge, stats);
  int rc = launched("gn_stats_kernel");
  if (rc == HESIC_OK) {
    gn_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(view(x), view(y), groups, stats, weight, bias, eps, relu, n);
    rc = launched("gn_apply_kernel");
  }
  cudaFreeAsync(stats, st);
  return rc;
}

extern "C" int hesic_softmax_channels(const hesic_tensor *x, const hesic_tensor *y, void *stream) {
  int r;
  if ((r = check_tensor(x, "softmax input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "softmax output")) != HESIC_OK) return r;
  HESIC_REQUIRE(x->fmt == HESIC_FMT_NCHW_F32 && y->fmt == HESIC_FMT_NCHW_F32, "softmax: NCHW fp32 tensors required");
  HESIC_REQUIRE(same_shape(x, y), "softmax: shape mismatch");
  if (numel(x) == 0) return HESIC_OK;
  size_t npix = (size_t)x->B * x->H * x->W;
  softmax_channels_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, as_stream(stream)>>>(view(x), view(y), npix);
  HESIC_LAUNCHED("softmax_channels_kernel");
  return HESIC_OK;
}

extern "C" int hesic_dense_warp(const hesic_tensor *h1, const hesic_tensor *cost, const hesic_tensor *out, void *stream) {
  int r;
  if ((r = check_tensor(h1, "dense_warp features")) != HESIC_OK) return r;
  if ((r = check_tensor(cost, "dense_warp cost")) != HESIC_OK) return r;
  if ((r = check_tensor(out, "dense_warp output")) != HESIC_OK) return r;
  HESIC_REQUIRE(h1->fmt == HESIC_FMT_NCHW_F32 && cost->fmt == HESIC_FMT_NCHW_F32 && out->fmt == HESIC_FMT_NCHW_F32,
                "dense_warp: NCHW fp32 tensors required");
  HESIC_REQUIRE(same_shape(h1, out), "dense_warp: output shape mismatch");
  HESIC_REQUIRE(cost->B == h1->B && cost->H == h1->H && cost->W == h1->W, "dense_warp: cost volume shape mismatch");
  HESIC_REQUIRE(cost->C >= 1 && cost->C <= DW_MAXD, "dense_warp: 1..%d disparities supported", DW_MAXD);
  HESIC_REQUIRE(h1->p0 != out->p0, "dense_warp: in-place is not supported");
  if (numel(out) == 0) return HESIC_OK;
  HESIC_REQUIRE(h1->H <= 65535 && h1->B <= 65535, "dense_warp: image too tall / batch too large");
  dim3 grid((h1->W + DW_TX - 1) / DW_TX, h1->H, h1->B);
  dense_warp_kernel<<<grid, 256, 0, as_stream(stream)>>>(view(h1), view(cost), view(out));
  HESIC_LAUNCHED("dense_warp_kernel");
  return HESIC_OK;
}
