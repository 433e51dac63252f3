// Operators that only the DSIC variant (ywz/DSIC/mynet6_plus.py) adds to the forward path: GroupNorm (+ReLU),
// softmax over the disparity channels of a cost volume, and dense_warp (disparity-weighted horizontal
// shift-sum).  All NCHW fp32 like the reference's tensors at those points; HBM-bound except dense_warp,
// which is a 32-tap 1-D correlation staged through shared memory.
#include <algorithm>

#include "common.cuh"
#include "conv.h"

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


// ---------------------------------------------------------------------------------------------
// Channels-last variants (the fused DSIC engine keeps activations channels-last between tensor-core convs):
// GroupNorm reads the conv output as NHWC fp32 and writes the next conv's input as bf16 (hi, lo) SPLIT planes,
// possibly a channel slice of a concatenation buffer.
//
// Statistics: block = R pixel rows x C/4 channel quads (a thread keeps ONE channel quad, hence one group, for the
// whole kernel), fp64 accumulation, shared-memory atomics per group, one global atomic per (block, group).
__global__ void __launch_bounds__(256) gn_stats_nhwc_kernel(const TView x, int G, int rows_per_block, double *__restrict__ stats,
                                                           int slot_stride) {
  extern __shared__ double gsum[];   // [2 * G]
  const int C4 = x.C >> 2, cpg = x.C / G;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) gsum[i] = 0.0;
  __syncthreads();
  const int cq = threadIdx.x % C4, r = threadIdx.x / C4;
  if (r < rows_per_block) {
    const size_t HW = (size_t)x.H * x.W;
    const float *base = (const float *)x.p0 + (size_t)b * HW * x.Cs + 4 * cq;
    double s = 0.0, ss = 0.0;
    for (size_t pix = (size_t)blockIdx.x * rows_per_block + r; pix < HW; pix += (size_t)gridDim.x * rows_per_block) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(base + pix * x.Cs));
      s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
      ss += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
    }
    const int g = (4 * cq) / cpg;
    atomicAdd(&gsum[2 * g], s);
    atomicAdd(&gsum[2 * g + 1], ss);
  }
  __syncthreads();
  // stats[b][g][slot][2]: `slot_stride` doubles between groups (2 for the plain layout, 2 * HESIC_GN_SLOTS for the slotted one)
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x)
    atomicAdd(&stats[((size_t)b * G + (i >> 1)) * slot_stride + (slot_stride > 2 ? 2 * (blockIdx.x & (HESIC_GN_SLOTS - 1)) : 0) + (i & 1)], gsum[i]);
}

// per (b, c): y = x * scale + shift with scale = rstd * weight[c], shift = bias[c] - mean * scale
__global__ void gn_finalize_kernel(const double *__restrict__ stats, int slots, int B, int C, int G, double count,
                                   const float *__restrict__ weight, const float *__restrict__ bias, float eps,
                                   float *__restrict__ scale, float *__restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C, g = c / (C / G);
  double s = 0.0, ss = 0.0;
  for (int k = 0; k < slots; ++k) {       // partial sums in a fixed order
    s += stats[((size_t)(b * G + g) * slots + k) * 2];
    ss += stats[((size_t)(b * G + g) * slots + k) * 2 + 1];
  }
  const double mean = s / count;
  const double var = fmax(ss / count - mean * mean, 0.0);
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double w = weight ? (double)weight[c] : 1.0, bb = (weight && bias) ? (double)bias[c] : 0.0;
  scale[i] = (float)(rstd * w);
  shift[i] = (float)(bb - mean * rstd * w);
}

__global__ void __launch_bounds__(256) gn_apply_nhwc_kernel(const TView x, const TView y, const float *__restrict__ scale,
                                                           const float *__restrict__ shift, int relu, size_t n4) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int C4 = x.C >> 2;
  const int cq = (int)(i % C4);
  const size_t pixb = i / C4;                       // b * HW + pix
  const int b = (int)(pixb / ((size_t)x.H * x.W));
  const float4 v = __ldg(reinterpret_cast<const float4 *>((const float *)x.p0 + pixb * x.Cs + 4 * cq));
  const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + (size_t)b * x.C + 4 * cq));
  const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift + (size_t)b * x.C + 4 * cq));
  float o[4] = {fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w)};
  if (relu) {
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = fmaxf(o[k], 0.f);
  }
  const size_t off = pixb * y.Cs + 4 * cq;
  if (y.fmt == HESIC_FMT_NHWC_SPLIT) {
    __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) split_bf16(o[k], hi[k], lo[k]);
    *reinterpret_cast<uint2 *>((__nv_bfloat16 *)y.p0 + off) = *reinterpret_cast<const uint2 *>(hi);
    *reinterpret_cast<uint2 *>((__nv_bfloat16 *)y.p1 + off) = *reinterpret_cast<const uint2 *>(lo);
  } else {
    *reinterpret_cast<float4 *>((float *)y.p0 + off) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// softmax over <= 64 contiguous channels, thread = pixel (NHWC fp32 in and out)
__global__ void __launch_bounds__(128) softmax_nhwc_kernel(const TView x, const TView y, size_t npix) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int C4 = x.C >> 2;
  const float4 *xp = reinterpret_cast<const float4 *>((const float *)x.p0 + i * x.Cs);
  float4 *yp = reinterpret_cast<float4 *>((float *)y.p0 + i * y.Cs);
  float4 v[16];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k < C4) { v[k] = __ldg(xp + k); m = fmaxf(fmaxf(m, fmaxf(v[k].x, v[k].y)), fmaxf(v[k].z, v[k].w)); }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k < C4) {
      v[k].x = expf(v[k].x - m); v[k].y = expf(v[k].y - m); v[k].z = expf(v[k].z - m); v[k].w = expf(v[k].w - m);
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  const float inv = 1.f / s;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k < C4) yp[k] = make_float4(v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv);
}

// dense_warp, channels-last: features and output as (hi, lo) SPLIT planes (channel slices allowed), cost NHWC fp32.
// Block = 32 pixels of one image row; the 32 + D - 1 feature pixels (all channels, fp32) and the 32 x D cost
// values are staged in shared memory; warp w owns 4 pixels, a lane 4 consecutive channels (stride 128 channels).
constexpr int DWN_TX = 32;
__global__ void __launch_bounds__(256) dense_warp_nhwc_kernel(const TView h1, const TView cost, const TView out) {
  extern __shared__ float dsm[];
  const int D = cost.C, W = h1.W, H = h1.H, Cn = h1.C;
  const int span = DWN_TX + D - 1;
  float *feat = dsm;                      // [span][Cn]
  float *cs = dsm + (size_t)span * Cn;    // [DWN_TX][D]
  const int x0 = blockIdx.x * DWN_TX, yy = blockIdx.y, b = blockIdx.z;
  const size_t rowpix = ((size_t)b * H + yy) * W;
  const int C8 = Cn >> 3;
  for (int i = threadIdx.x; i < span * C8; i += 256) {
    const int col = i / C8, c8 = i - col * C8;
    const int gx = x0 + col;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (gx < W) {
      const size_t o = (rowpix + gx) * h1.Cs + 8 * c8;
      const uint4 hv = __ldg(reinterpret_cast<const uint4 *>((const __nv_bfloat16 *)h1.p0 + o));
      const uint4 lv = __ldg(reinterpret_cast<const uint4 *>((const __nv_bfloat16 *)h1.p1 + o));
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        f[2 * k] = __uint_as_float(hw[k] << 16) + __uint_as_float(lw[k] << 16);
        f[2 * k + 1] = __uint_as_float(hw[k] & 0xffff0000u) + __uint_as_float(lw[k] & 0xffff0000u);
      }
    }
    float4 *d = reinterpret_cast<float4 *>(feat + (size_t)col * Cn + 8 * c8);
    d[0] = make_float4(f[0], f[1], f[2], f[3]);
    d[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  for (int i = threadIdx.x; i < DWN_TX * D; i += 256) {
    const int px = i / D, d = i - px * D;
    const int gx = x0 + px;
    cs[i] = (gx < W && gx + d < W) ? __ldg((const float *)cost.p0 + (rowpix + gx) * cost.Cs + d) : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pp = 0; pp < DWN_TX / 8; ++pp) {
    const int px = warp * (DWN_TX / 8) + pp;
    const int gx = x0 + px;
    if (gx >= W) continue;
    for (int c4 = lane; c4 < (Cn >> 2); c4 += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int d = 0; d < D; ++d) {
        const float w = cs[px * D + d];
        const float4 f = *reinterpret_cast<const float4 *>(feat + (size_t)(px + d) * Cn + 4 * c4);
        acc.x = fmaf(w, f.x, acc.x); acc.y = fmaf(w, f.y, acc.y); acc.z = fmaf(w, f.z, acc.z); acc.w = fmaf(w, f.w, acc.w);
      }
      const float o[4] = {acc.x, acc.y, acc.z, acc.w};
      __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split_bf16(o[k], hi[k], lo[k]);
      const size_t off = (rowpix + gx) * out.Cs + 4 * c4;
      *reinterpret_cast<uint2 *>((__nv_bfloat16 *)out.p0 + off) = *reinterpret_cast<const uint2 *>(hi);
      *reinterpret_cast<uint2 *>((__nv_bfloat16 *)out.p1 + off) = *reinterpret_cast<const uint2 *>(lo);
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
  HESIC_REQUIRE(same_shape(x, y), "group_norm: shape mismatch");
  HESIC_REQUIRE(groups >= 1 && x->C % groups == 0, "group_norm: %d channels are not divisible into %d groups", x->C, groups);
  const int xCs = x->Cs > 0 ? x->Cs : x->C, yCs = y->Cs > 0 ? y->Cs : y->C;
  size_t n = numel(x);
  if (n == 0) return HESIC_OK;
  cudaStream_t st = as_stream(stream);
  const int BG = x->B * groups;
  HESIC_REQUIRE(BG <= 65535, "group_norm: too many (batch, group) pairs");
  if (x->fmt == HESIC_FMT_NHWC_F32) {
    // channels-last: conv output (fp32) -> next conv's input (SPLIT planes or fp32), channel slices allowed
    HESIC_REQUIRE(y->fmt == HESIC_FMT_NHWC_SPLIT || y->fmt == HESIC_FMT_NHWC_F32, "group_norm: an NHWC input needs an NHWC output");
    const int cpg = x->C / groups, C4 = x->C / 4;
    HESIC_REQUIRE(x->C % 4 == 0 && cpg % 4 == 0 && xCs % 4 == 0 && yCs % 4 == 0 && C4 <= 256,
                  "group_norm (NHWC): channels per group and strides must be multiples of 4, C <= 1024");
    HESIC_REQUIRE(((uintptr_t)x->p0 & 15) == 0 && ((uintptr_t)y->p0 & 7) == 0 && (y->fmt != HESIC_FMT_NHWC_SPLIT || ((uintptr_t)y->p1 & 7) == 0),
                  "group_norm (NHWC): misaligned tensor");
    const size_t HW = (size_t)x->H * x->W;
    double *stats = nullptr;
    const size_t stat_bytes = (size_t)BG * 2 * sizeof(double), aff_bytes = (size_t)x->B * x->C * sizeof(float);
    HESIC_CUDA(cudaMallocAsync(&stats, stat_bytes + 2 * aff_bytes, st));
    float *scale = reinterpret_cast<float *>(reinterpret_cast<char *>(stats) + stat_bytes), *shift = scale + (size_t)x->B * x->C;
    cudaMemsetAsync(stats, 0, stat_bytes, st);
    const int rows = 256 / C4;
    const unsigned bx = (unsigned)std::max<size_t>(1, std::min<size_t>((HW + rows - 1) / rows, 2048 / std::max(1, x->B)));
    gn_stats_nhwc_kernel<<<dim3(bx, x->B), 256, 2 * groups * sizeof(double), st>>>(view(x), groups, rows, stats, 2);
    int rc = launched("gn_stats_nhwc_kernel");
    if (rc == HESIC_OK) {
      gn_finalize_kernel<<<(x->B * x->C + 255) / 256, 256, 0, st>>>(stats, 1, x->B, x->C, groups, (double)cpg * (double)HW, weight, bias, eps,
                                                                  scale, shift);
      rc = launched("gn_finalize_kernel");
    }
    if (rc == HESIC_OK) {
      const size_t n4 = n / 4;
      gn_apply_nhwc_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(view(x), view(y), scale, shift, relu, n4);
      rc = launched("gn_apply_nhwc_kernel");
    }
    cudaFreeAsync(stats, st);
    return rc;
  }
  HESIC_REQUIRE(x->fmt == HESIC_FMT_NCHW_F32 && y->fmt == HESIC_FMT_NCHW_F32, "group_norm: NCHW fp32 (or NHWC) tensors required");
  HESIC_REQUIRE(xCs == x->C, "group_norm: input must not be a channel slice");
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

// Statistics of an NHWC fp32 tensor into the slotted layout [B][groups][HESIC_GN_SLOTS][2] (the route for outputs whose
// producing convolution could not accumulate them in its epilogue); `stats` zeroed by the caller.
namespace hesic {
int gn_stats_slotted(const hesic_tensor *x, int groups, double *stats, cudaStream_t st) {
  const int C4 = x->C / 4;
  HESIC_REQUIRE(x->fmt == HESIC_FMT_NHWC_F32 && x->C % 4 == 0 && (x->C / groups) % 4 == 0 && C4 <= 256 && x->B <= 65535,
                "group-norm statistics: NHWC fp32 input with channels per group a multiple of 4 required");
  const size_t HW = (size_t)x->H * x->W;
  const int rows = 256 / C4;
  const unsigned bx = (unsigned)std::max<size_t>(1, std::min<size_t>((HW + rows - 1) / rows, 2048 / std::max(1, x->B)));
  gn_stats_nhwc_kernel<<<dim3(bx, x->B), 256, 2 * groups * sizeof(double), st>>>(view(x), groups, rows, stats, 2 * HESIC_GN_SLOTS);
  return launched("gn_stats_nhwc_kernel");
}
}  // namespace hesic

extern "C" int hesic_group_norm_apply(const hesic_tensor *x, const hesic_tensor *y, int groups, const float *weight,
                                      const float *bias, float eps, int relu, const double *stats, void *stream) {
  int r;
  if ((r = check_tensor(x, "group_norm input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "group_norm output")) != HESIC_OK) return r;
  HESIC_REQUIRE(same_shape(x, y) && stats, "group_norm_apply: shape mismatch or null statistics");
  HESIC_REQUIRE(groups >= 1 && x->C % groups == 0, "group_norm: %d channels are not divisible into %d groups", x->C, groups);
  HESIC_REQUIRE(x->fmt == HESIC_FMT_NHWC_F32 && (y->fmt == HESIC_FMT_NHWC_SPLIT || y->fmt == HESIC_FMT_NHWC_F32),
                "group_norm_apply: NHWC fp32 input, NHWC fp32 or SPLIT output");
  const int xCs = x->Cs > 0 ? x->Cs : x->C, yCs = y->Cs > 0 ? y->Cs : y->C;
  HESIC_REQUIRE(x->C % 4 == 0 && xCs % 4 == 0 && yCs % 4 == 0 && ((uintptr_t)x->p0 & 15) == 0 && ((uintptr_t)y->p0 & 7) == 0 &&
                    (y->fmt != HESIC_FMT_NHWC_SPLIT || ((uintptr_t)y->p1 & 7) == 0),
                "group_norm_apply: channel counts / strides must be multiples of 4 and the tensors aligned");
  const size_t n = numel(x);
  if (n == 0) return HESIC_OK;
  cudaStream_t st = as_stream(stream);
  float *scale = nullptr;
  const size_t aff = (size_t)x->B * x->C;
  HESIC_CUDA(cudaMallocAsync(&scale, 2 * aff * sizeof(float), st));
  float *shift = scale + aff;
  const int cpg = x->C / groups;
  gn_finalize_kernel<<<(unsigned)((aff + 255) / 256), 256, 0, st>>>(stats, HESIC_GN_SLOTS, x->B, x->C, groups,
                                                                  (double)cpg * (double)x->H * (double)x->W, weight, bias, eps, scale, shift);
  int rc = launched("gn_finalize_kernel");
  if (rc == HESIC_OK) {
    gn_apply_nhwc_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(view(x), view(y), scale, shift, relu, n / 4);
    rc = launched("gn_apply_nhwc_kernel");
  }
  cudaFreeAsync(scale, st);
  return rc;
}

extern "C" int hesic_softmax_channels(const hesic_tensor *x, const hesic_tensor *y, void *stream) {
  int r;
  if ((r = check_tensor(x, "softmax input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "softmax output")) != HESIC_OK) return r;
  HESIC_REQUIRE(same_shape(x, y), "softmax: shape mismatch");
  if (numel(x) == 0) return HESIC_OK;
  size_t npix = (size_t)x->B * x->H * x->W;
  if (x->fmt == HESIC_FMT_NHWC_F32) {
    const int xCs = x->Cs > 0 ? x->Cs : x->C, yCs = y->Cs > 0 ? y->Cs : y->C;
    HESIC_REQUIRE(y->fmt == HESIC_FMT_NHWC_F32, "softmax: an NHWC input needs an NHWC fp32 output");
    HESIC_REQUIRE(x->C % 4 == 0 && x->C <= 64 && xCs % 4 == 0 && yCs % 4 == 0 && ((uintptr_t)x->p0 & 15) == 0 && ((uintptr_t)y->p0 & 15) == 0,
                  "softmax (NHWC): at most 64 channels, multiples of 4, 16-byte aligned");
    softmax_nhwc_kernel<<<(unsigned)((npix + 127) / 128), 128, 0, as_stream(stream)>>>(view(x), view(y), npix);
    HESIC_LAUNCHED("softmax_nhwc_kernel");
    return HESIC_OK;
  }
  HESIC_REQUIRE(x->fmt == HESIC_FMT_NCHW_F32 && y->fmt == HESIC_FMT_NCHW_F32, "softmax: NCHW fp32 (or NHWC fp32) tensors required");
  softmax_channels_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, as_stream(stream)>>>(view(x), view(y), npix);
  HESIC_LAUNCHED("softmax_channels_kernel");
  return HESIC_OK;
}

extern "C" int hesic_dense_warp(const hesic_tensor *h1, const hesic_tensor *cost, const hesic_tensor *out, void *stream) {
  int r;
  if ((r = check_tensor(h1, "dense_warp features")) != HESIC_OK) return r;
  if ((r = check_tensor(cost, "dense_warp cost")) != HESIC_OK) return r;
  if ((r = check_tensor(out, "dense_warp output")) != HESIC_OK) return r;
  HESIC_REQUIRE(same_shape(h1, out), "dense_warp: output shape mismatch");
  if (h1->fmt == HESIC_FMT_NHWC_SPLIT) {
    HESIC_REQUIRE(cost->fmt == HESIC_FMT_NHWC_F32 && out->fmt == HESIC_FMT_NHWC_SPLIT, "dense_warp: SPLIT features need an NHWC fp32 cost and a SPLIT output");
    HESIC_REQUIRE(cost->B == h1->B && cost->H == h1->H && cost->W == h1->W, "dense_warp: cost volume shape mismatch");
    HESIC_REQUIRE(cost->C >= 1 && cost->C <= DW_MAXD, "dense_warp: 1..%d disparities supported", DW_MAXD);
    const int hCs = h1->Cs > 0 ? h1->Cs : h1->C, oCs = out->Cs > 0 ? out->Cs : out->C;
    HESIC_REQUIRE(h1->C % 8 == 0 && hCs % 8 == 0 && oCs % 4 == 0 && ((uintptr_t)h1->p0 & 15) == 0 && ((uintptr_t)h1->p1 & 15) == 0 &&
                      ((uintptr_t)out->p0 & 7) == 0 && ((uintptr_t)out->p1 & 7) == 0,
                  "dense_warp (channels-last): channel counts must be multiples of 8, planes 16-byte aligned");
    if (numel(out) == 0) return HESIC_OK;
    HESIC_REQUIRE(h1->H <= 65535 && h1->B <= 65535, "dense_warp: image too tall / batch too large");
    const size_t smem = ((size_t)(DWN_TX + cost->C - 1) * h1->C + (size_t)DWN_TX * cost->C) * sizeof(float);
    HESIC_REQUIRE(smem <= 200 * 1024, "dense_warp (channels-last): %d channels x %d disparities do not fit shared memory", h1->C, cost->C);
    static size_t attr[64] = {0};
    const int dv = current_device() & 63;
    if (smem > attr[dv]) {
      HESIC_CUDA(cudaFuncSetAttribute(dense_warp_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr[dv] = smem;
    }
    dim3 grid((h1->W + DWN_TX - 1) / DWN_TX, h1->H, h1->B);
    dense_warp_nhwc_kernel<<<grid, 256, smem, as_stream(stream)>>>(view(h1), view(cost), view(out));
    HESIC_LAUNCHED("dense_warp_nhwc_kernel");
    return HESIC_OK;
  }
  HESIC_REQUIRE(h1->fmt == HESIC_FMT_NCHW_F32 && cost->fmt == HESIC_FMT_NCHW_F32 && out->fmt == HESIC_FMT_NCHW_F32,
                "dense_warp: NCHW fp32 tensors (or SPLIT features) required");
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
