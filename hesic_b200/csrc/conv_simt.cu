// Generic fp32 implicit-GEMM convolution on CUDA cores (sm_100a).
//
// Role in the design: (1) the layers whose GEMM shape cannot feed a tensor-core tile -- the
// full-resolution edge layers with Cin or Cout in {3, 6} (newnet1.py:583,612,629,670) and the
// 3-channel GDNs -- and (2) the correctness anchor the tcgen05 path (conv_tc.cu) is validated
// against on the device.  One kernel covers nn.Conv2d (any stride) and nn.ConvTranspose2d in gather
// form, with bias + activation (+ GDN as a 1x1 contraction over x^2) fused.
#include <algorithm>

#include <stdlib.h>

#include <vector>

#include "conv.h"

namespace hesic {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct ConvArgs {
  TView x, y;
  const float *w;     // [K][Cout]
  const float *bias;  // [Cout]
  int Cin, Cout, kh, kw, stride, pad, transposed;
  int Hout, Wout, K, M;
  int act;
  int gdn;  // 0 conv, 1 GDN, 2 inverse GDN (A = x^2, epilogue x * (r)sqrt)
};

template <bool NHWC_IN>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const ConvArgs a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // A-load mapping: NHWC -> consecutive threads walk k (channels); NCHW -> walk pixels.
  int am[4], ak[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int e = tid + i * NT;
    if (NHWC_IN) { ak[i] = e % BK; am[i] = e / BK; }
    else         { am[i] = e % BM; ak[i] = e / BM; }
  }
  int ab[4], aoy[4], aox[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + am[i];
    if (m < a.M) {
      ab[i] = m / (a.Hout * a.Wout);
      int r = m - ab[i] * a.Hout * a.Wout;
      aoy[i] = r / a.Wout;
      aox[i] = r - aoy[i] * a.Wout;
    } else {
      ab[i] = -1; aoy[i] = 0; aox[i] = 0;
    }
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = 0.f;
      int kg = k0 + ak[i];
      if (kg < a.K && ab[i] >= 0) {
        int tap = kg / a.Cin, ci = kg - tap * a.Cin;
        int ky = tap / a.kw, kx = tap - ky * a.kw;
        int iy, ix;
        bool ok = true;
        if (!a.transposed) {
          iy = aoy[i] * a.stride + ky - a.pad;
          ix = aox[i] * a.stride + kx - a.pad;
        } else {
          int ty_ = aoy[i] + a.pad - ky, tx_ = aox[i] + a.pad - kx;
          ok = ty_ >= 0 && tx_ >= 0 && (ty_ % a.stride) == 0 && (tx_ % a.stride) == 0;
          iy = ty_ / a.stride;
          ix = tx_ / a.stride;
        }
        if (ok && iy >= 0 && iy < a.x.H && ix >= 0 && ix < a.x.W) {
          v = tload(a.x, ab[i], ci, iy, ix);
          if (a.gdn) v = v * v;
        }
      }
      As[ak[i]][am[i]] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * NT;
      int n = e % BN, k = e / BN;
      int kg = k0 + k;
      Bs[k][n] = (kg < a.K && n0 + n < a.Cout) ? a.w[(size_t)kg * a.Cout + n0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    int b = m / (a.Hout * a.Wout);
    int r = m - b * a.Hout * a.Wout;
    int oy = r / a.Wout, ox = r - oy * a.Wout;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= a.Cout) continue;
      float v = acc[i][j] + a.bias[n];
      if (a.gdn) {
        float xin = tload(a.x, b, n, oy, ox);
        v = xin * (a.gdn == 2 ? sqrtf(v) : rsqrtf(v));
      }
      v = apply_act(v, a.act);
      tstore(a.y, b, n, oy, ox, v);
    }
  }
}

int conv_out_size(const hesic_conv *c, int in, int k) {
  if (!c->transposed) return (in + 2 * c->pad - k) / c->stride + 1;
  return (in - 1) * c->stride - 2 * c->pad + k + c->out_pad;
}

static int launch_simt(const ConvArgs &a, cudaStream_t s) {
  if (a.M == 0 || a.Cout == 0) return HESIC_OK;
  dim3 grid((a.M + BM - 1) / BM, (a.Cout + BN - 1) / BN);
  if (a.x.fmt == HESIC_FMT_NCHW_F32) conv_simt_kernel<false><<<grid, NT, 0, s>>>(a);
  else conv_simt_kernel<true><<<grid, NT, 0, s>>>(a);
  HESIC_LAUNCHED("conv_simt_kernel");
  return HESIC_OK;
}

int conv_forward_simt(hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y, int act, cudaStream_t s) {
  ConvArgs a;
  a.x = view(x); a.y = view(y);
  a.w = c->w_simt; a.bias = c->bias;
  a.Cin = c->Cin; a.Cout = c->Cout; a.kh = c->kh; a.kw = c->kw; a.stride = c->stride; a.pad = c->pad;
  a.transposed = c->transposed;
  a.Hout = y->H; a.Wout = y->W; a.K = c->kh * c->kw * c->Cin; a.M = y->B * y->H * y->W;
  a.act = c->has_gdn ? HESIC_ACT_NONE : act; a.gdn = 0;
  if (!c->has_gdn) return launch_simt(a, s);
  // unfused: conv -> scratch in y's own storage is not possible (GDN mixes channels), so run the conv
  // into a temporary fp32 NHWC buffer and the GDN contraction from there.
  float *tmp = nullptr;
  size_t n = (size_t)y->B * y->H * y->W * c->Cout;
  HESIC_CUDA(cudaMallocAsync(&tmp, n * sizeof(float), s));
  hesic_tensor t = *y;
  t.p0 = tmp; t.p1 = nullptr; t.fmt = HESIC_FMT_NHWC_F32; t.Cs = c->Cout;
  a.y = view(&t);
  int r = launch_simt(a, s);
  if (r == HESIC_OK) {
    r = gdn_simt(&t, y, c->gdn_beta, c->gdn_w_simt, c->gdn_inverse, s);
    if (r == HESIC_OK && act != HESIC_ACT_NONE) { set_error("activation after fused GDN is not supported"); r = HESIC_E_UNSUPPORTED; }
  }
  cudaFreeAsync(tmp, s);
  return r;
}

int gdn_simt(const hesic_tensor *x, const hesic_tensor *y, const float *beta_rp, const float *w_simt, int inverse,
             cudaStream_t s) {
  ConvArgs a;
  a.x = view(x); a.y = view(y);
  a.w = w_simt; a.bias = beta_rp;
  a.Cin = x->C; a.Cout = x->C; a.kh = 1; a.kw = 1; a.stride = 1; a.pad = 0; a.transposed = 0;
  a.Hout = x->H; a.Wout = x->W; a.K = x->C; a.M = x->B * x->H * x->W;
  a.act = HESIC_ACT_NONE; a.gdn = inverse ? 2 : 1;
  return launch_simt(a, s);
}

// ---------------------------------------------------------------------------------------------
// operand packing
__global__ void pack_conv_kernel(const float *__restrict__ w, const float *__restrict__ mask, int Cin, int Cout,
                                 int kh, int kw, int transposed, int CoutPad, float *__restrict__ w_simt,
                                 __nv_bfloat16 *__restrict__ w_hi, __nv_bfloat16 *__restrict__ w_lo) {
  size_t total = (size_t)kh * kw * Cin * CoutPad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ci = i % Cin;
    size_t r = i / Cin;
    int co = r % CoutPad;
    int tap = r / CoutPad;
    int ky = tap / kw, kx = tap % kw;
    float v = 0.f;
    if (co < Cout) {
      size_t src = transposed ? ((((size_t)ci * Cout + co) * kh + ky) * kw + kx)
                              : ((((size_t)co * Cin + ci) * kh + ky) * kw + kx);
      v = w[src];
      if (mask) v *= mask[src];
      w_simt[((size_t)tap * Cin + ci) * Cout + co] = v;
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    w_hi[i] = hi;  // [tap][CoutPad][Cin]
    w_lo[i] = lo;
  }
}

// ROW formulation: [ky][CoutPad][kx*8 + c]; a stride-1 transposed conv is the correlation with the
// kernel flipped in both axes.
__global__ void pack_conv_row_kernel(const float *__restrict__ w, const float *__restrict__ mask, int Cin, int Cout,
                                     int transposed, int CoutPad, __nv_bfloat16 *__restrict__ w_hi,
                                     __nv_bfloat16 *__restrict__ w_lo) {
  int total = 5 * CoutPad * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int e = i % 64, r = i / 64;
    int co = r % CoutPad, ky = r / CoutPad;
    int kx = e / 8, ci = e % 8;
    float v = 0.f;
    if (kx < 5 && ci < Cin && co < Cout) {
      size_t src = transposed ? ((((size_t)ci * Cout + co) * 5 + (4 - ky)) * 5 + (4 - kx))
                              : ((((size_t)co * Cin + ci) * 5 + ky) * 5 + kx);
      v = w[src];
      if (mask) v *= mask[src];
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    w_hi[i] = hi;
    w_lo[i] = lo;
  }
}

// ROW2 formulation: [s = kernel-row pair][CoutPad][kx*8 + r*4 + c], ky = 2*s + r (rows >= 5 and kx >= 5 zero)
__global__ void pack_conv_row2_kernel(const float *__restrict__ w, const float *__restrict__ mask, int Cin, int Cout,
                                      int CoutPad, __nv_bfloat16 *__restrict__ w_hi, __nv_bfloat16 *__restrict__ w_lo) {
  int total = 3 * CoutPad * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int e = i % 64, r = i / 64;
    int co = r % CoutPad, s = r / CoutPad;
    int ky = 2 * s + (e % 8) / 4, kx = e / 8, ci = e % 4;
    float v = 0.f;
    if (ky < 5 && kx < 5 && ci < Cin && co < Cout) {
      size_t src = (((size_t)co * Cin + ci) * 5 + ky) * 5 + kx;
      v = w[src];
      if (mask) v *= mask[src];
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    w_hi[i] = hi;
    w_lo[i] = lo;
  }
}

// SCATTER formulation (RGB synthesis head): [n = (ky*5 + kx)*Cout + co][ci], rows >= 25*Cout zero
__global__ void pack_conv_scatter_kernel(const float *__restrict__ w, int Cin, int Cout, int NPAD,
                                         __nv_bfloat16 *__restrict__ w_hi, __nv_bfloat16 *__restrict__ w_lo) {
  int total = NPAD * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int ci = i % Cin, n = i / Cin;
    float v = 0.f;
    if (n < 25 * Cout) {
      int tap = n / Cout, co = n % Cout;
      v = w[((size_t)ci * Cout + co) * 25 + tap];   // ConvTranspose2d weight [Cin][Cout][5][5]
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    w_hi[i] = hi;
    w_lo[i] = lo;
  }
}

__global__ void pack_gdn_kernel(const float *__restrict__ beta, const float *__restrict__ gamma, int C, float beta_bound,
                                float gamma_bound, float pedestal, float *__restrict__ beta_rp,
                                float *__restrict__ w_simt, __nv_bfloat16 *__restrict__ g_hi,
                                __nv_bfloat16 *__restrict__ g_lo) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < C) {
    float o = fmaxf(beta[idx], beta_bound);
    beta_rp[idx] = o * o - pedestal;
  }
  if (idx < C * C) {
    int i = idx / C, j = idx % C;  // gamma[i][j]: output channel i, input channel j
    float o = fmaxf(gamma[idx], gamma_bound);
    float g = o * o - pedestal;
    w_simt[(size_t)j * C + i] = g;
    if (g_hi) {
      __nv_bfloat16 hi, lo;
      split_bf16(g, hi, lo);
      g_hi[idx] = hi;
      g_lo[idx] = lo;
    }
  }
}

__global__ void fill_zero_kernel(float *p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

static void reparam_bounds(float minimum, float *bound, float *pedestal) {
  // compressai/ops/parametrizers.py:27-36 -- computed in Python doubles, stored as fp32 buffers
  double off = ldexp(1.0, -18);
  *pedestal = (float)(off * off);
  *bound = (float)sqrt((double)minimum + off * off);
}

int pack_gdn(const float *beta, const float *gamma, int C, float beta_min, float *beta_rp, float *w_simt,
             __nv_bfloat16 *g_hi, __nv_bfloat16 *g_lo, cudaStream_t s) {
  float bb, gb, ped;
  reparam_bounds(beta_min, &bb, &ped);
  reparam_bounds(0.f, &gb, &ped);
  int n = C * C;
  pack_gdn_kernel<<<(n + 255) / 256, 256, 0, s>>>(beta, gamma, C, bb, gb, ped, beta_rp, w_simt, g_hi, g_lo);
  HESIC_LAUNCHED("pack_gdn_kernel");
  return HESIC_OK;
}

}  // namespace hesic

using namespace hesic;

extern "C" hesic_conv *hesic_conv_create(int Cin, int Cout, int kh, int kw, int stride, int pad, int transposed,
                                         int output_padding) {
  if (Cin <= 0 || Cout <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0 || output_padding < 0) {
    set_error("hesic_conv_create: invalid geometry");
    return nullptr;
  }
  hesic_conv *c = new hesic_conv();
  c->Cin = Cin; c->Cout = Cout; c->kh = kh; c->kw = kw; c->stride = stride; c->pad = pad;
  c->transposed = transposed ? 1 : 0; c->out_pad = output_padding;
  c->CoutPad = (Cout + 15) / 16 * 16;
  c->tc_kind = HESIC_TC_GENERIC; c->tc_taps = kh * kw; c->tc_k = Cin;
  const bool k5 = kh == 5 && kw == 5 && pad == 2;
  if (Cin <= 4 && k5 && !c->transposed && stride == 2) {
    c->tc_kind = HESIC_TC_ROW2; c->tc_taps = 3; c->tc_k = 64;
  } else if (Cin <= 8 && k5 && ((!c->transposed && (stride == 1 || stride == 2)) ||
                         (c->transposed && stride == 1 && output_padding == 0))) {
    c->tc_kind = HESIC_TC_ROW; c->tc_taps = 5; c->tc_k = 64;
  } else if (c->transposed && stride == 2 && k5 && output_padding == 1 && Cout <= 4 && Cin % 64 == 0 && Cin <= 256) {
    c->tc_kind = HESIC_TC_SCATTER; c->tc_taps = 1; c->tc_k = Cin;
  }
  return c;
}

static void conv_release(hesic_conv *c) {
  cudaFree(c->w_simt); cudaFree(c->bias); cudaFree(c->w_hi); cudaFree(c->w_lo);
  cudaFree(c->gdn_beta); cudaFree(c->gdn_w_simt); cudaFree(c->gdn_g_hi); cudaFree(c->gdn_g_lo);
  c->w_simt = c->bias = nullptr; c->w_hi = c->w_lo = nullptr;
  c->gdn_beta = c->gdn_w_simt = nullptr; c->gdn_g_hi = c->gdn_g_lo = nullptr;
  c->tc_maps_bn = 0; c->tc_maps_gdn = -1;     // cached tensor maps point into the freed operands
  c->loaded = false; c->has_gdn = false;
}

extern "C" void hesic_conv_destroy(hesic_conv *c) {
  if (!c) return;
  conv_release(c);
  free(c->tc_maps);
  free(c->w_host);
  delete c;
}

// The packed operands live on the device that was current when they were packed.  A model moved to another device
// (`model.to('cuda:1')`) re-packs (the host side keys on the device index): free the old device's copies first.
static void conv_claim_device(hesic_conv *c) {
  const int dev = current_device();
  if (c->device >= 0 && c->device != dev) conv_release(c);
  c->device = dev;
}

extern "C" int hesic_conv_load(hesic_conv *c, const float *weight, const float *bias, const float *mask, void *stream) {
  HESIC_REQUIRE(c && weight, "hesic_conv_load: null argument");
  cudaStream_t s = as_stream(stream);
  size_t taps = (size_t)c->kh * c->kw;
  conv_claim_device(c);
  if (!c->w_simt) {
    HESIC_CUDA(cudaMalloc(&c->w_simt, taps * c->Cin * c->Cout * sizeof(float)));
    HESIC_CUDA(cudaMalloc(&c->bias, c->Cout * sizeof(float)));
    size_t tc_elems = std::max(taps * c->Cin, (size_t)c->tc_taps * c->tc_k) * c->CoutPad;
    HESIC_CUDA(cudaMalloc(&c->w_hi, tc_elems * sizeof(__nv_bfloat16)));
    HESIC_CUDA(cudaMalloc(&c->w_lo, tc_elems * sizeof(__nv_bfloat16)));
  }
  c->live_taps = ~0ull;
  c->kband_bn = c->kband_chunks = 0;
  if (mask && taps <= 64) {
    // once per load: which taps does the mask keep?  (host copy + sync; weights are loaded once per model)
    const size_t n = (size_t)c->Cout * c->Cin * taps;
    std::vector<float> hm(n);
    HESIC_CUDA(cudaMemcpyAsync(hm.data(), mask, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    HESIC_CUDA(cudaStreamSynchronize(s));
    uint64_t live = 0;
    for (size_t i = 0; i < n; ++i)
      if (hm[i] != 0.f) live |= 1ull << (i % taps);
    c->live_taps = live ? live : ~0ull;   // an all-zero mask keeps every (zero-weight) tap
  }
  size_t total = taps * c->Cin * c->CoutPad;
  int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_conv_kernel<<<blocks, 256, 0, s>>>(weight, mask, c->Cin, c->Cout, c->kh, c->kw, c->transposed, c->CoutPad,
                                          c->w_simt, c->w_hi, c->w_lo);
  HESIC_LAUNCHED("pack_conv_kernel");
  // the tensor-core operand planes of the ROW / SCATTER formulations replace the generic ones
  if (c->tc_kind == HESIC_TC_ROW) {
    pack_conv_row_kernel<<<(5 * c->CoutPad * 64 + 255) / 256, 256, 0, s>>>(weight, mask, c->Cin, c->Cout, c->transposed,
                                                                       c->CoutPad, c->w_hi, c->w_lo);
    HESIC_LAUNCHED("pack_conv_row_kernel");
  } else if (c->tc_kind == HESIC_TC_ROW2) {
    pack_conv_row2_kernel<<<(3 * c->CoutPad * 64 + 255) / 256, 256, 0, s>>>(weight, mask, c->Cin, c->Cout, c->CoutPad,
                                                                        c->w_hi, c->w_lo);
    HESIC_LAUNCHED("pack_conv_row2_kernel");
  } else if (c->tc_kind == HESIC_TC_SCATTER) {
    HESIC_REQUIRE(mask == nullptr, "hesic_conv_load: masked transposed RGB head is not supported");
    const int NPAD = (25 * c->Cout + 15) / 16 * 16;
    pack_conv_scatter_kernel<<<(NPAD * c->Cin + 255) / 256, 256, 0, s>>>(weight, c->Cin, c->Cout, NPAD, c->w_hi, c->w_lo);
    HESIC_LAUNCHED("pack_conv_scatter_kernel");
  }
  if (bias) {
    HESIC_CUDA(cudaMemcpyAsync(c->bias, bias, c->Cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
    fill_zero_kernel<<<(c->Cout + 255) / 256, 256, 0, s>>>(c->bias, c->Cout);
    HESIC_LAUNCHED("fill_zero_kernel");
  }
  if (c->Cin <= 8 && c->Cout <= 4 && c->kh == 5 && c->kw == 5) {
    // once per load (host copy + sync, like the mask above): the stencil kernel's parameter-space weights
    const size_t n = taps * c->Cin * c->Cout;
    if (!c->w_host) c->w_host = (float *)malloc(n * sizeof(float));
    HESIC_CUDA(cudaMemcpyAsync(c->w_host, c->w_simt, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    HESIC_CUDA(cudaStreamSynchronize(s));
  }
  c->loaded = true;
  return HESIC_OK;
}

extern "C" int hesic_conv_set_gdn(hesic_conv *c, const float *beta, const float *gamma, int inverse, float beta_min,
                                  void *stream) {
  HESIC_REQUIRE(c, "hesic_conv_set_gdn: null conv");
  if (!beta || !gamma) { c->has_gdn = false; return HESIC_OK; }
  HESIC_REQUIRE(c->device < 0 || c->device == current_device(),
                "hesic_conv_set_gdn: the layer's operands live on device %d, the current device is %d", c->device, current_device());
  int C = c->Cout;
  if (!c->gdn_beta) {
    HESIC_CUDA(cudaMalloc(&c->gdn_beta, C * sizeof(float)));
    HESIC_CUDA(cudaMalloc(&c->gdn_w_simt, (size_t)C * C * sizeof(float)));
    HESIC_CUDA(cudaMalloc(&c->gdn_g_hi, (size_t)C * C * sizeof(__nv_bfloat16)));
    HESIC_CUDA(cudaMalloc(&c->gdn_g_lo, (size_t)C * C * sizeof(__nv_bfloat16)));
  }
  int r = pack_gdn(beta, gamma, C, beta_min, c->gdn_beta, c->gdn_w_simt, c->gdn_g_hi, c->gdn_g_lo, as_stream(stream));
  if (r != HESIC_OK) return r;
  c->has_gdn = true;
  c->gdn_inverse = inverse ? 1 : 0;
  return HESIC_OK;
}

// one block per (N tile, K chunk): does the block of weights hold any non-zero entry?
__global__ void kband_kernel(const float *__restrict__ w, int Cin, int Cout, int taps, int BN, int kchunks, int *__restrict__ flags) {
  const int nt = blockIdx.x / kchunks, kc = blockIdx.x - nt * kchunks;
  const int co0 = nt * BN, co1 = min(co0 + BN, Cout), ci0 = kc * 64, ci1 = min(ci0 + 64, Cin);
  const int ncin = ci1 - ci0;
  const size_t n = (size_t)(co1 - co0) * ncin * taps;
  int any = 0;
  for (size_t i = threadIdx.x; i < n && !any; i += blockDim.x) {
    const int t = (int)(i % taps);
    const size_t r = i / taps;
    const int ci = ci0 + (int)(r % ncin), co = co0 + (int)(r / ncin);
    if (w[((size_t)co * Cin + ci) * taps + t] != 0.f) any = 1;
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) flags[blockIdx.x] = 1;
}

extern "C" int hesic_conv_detect_kband(hesic_conv *c, const float *weight, void *stream) {
  HESIC_REQUIRE(c && weight, "hesic_conv_detect_kband: null argument");
  c->kband_bn = c->kband_chunks = 0;
  if (c->transposed || c->tc_kind != HESIC_TC_GENERIC) return HESIC_OK;     // plain convolutions on the generic tensor-core path only
  const int BN = std::min(128, (c->Cout + 31) / 32 * 32), kchunks = (c->Cin + 63) / 64, n_tiles = (c->Cout + BN - 1) / BN;
  if (kchunks < 2 || n_tiles > 8 || kchunks > 127) return HESIC_OK;
  cudaStream_t s = as_stream(stream);
  int *flags = nullptr;
  const int nb = n_tiles * kchunks;
  HESIC_CUDA(cudaMallocAsync(&flags, nb * sizeof(int), s));
  HESIC_CUDA(cudaMemsetAsync(flags, 0, nb * sizeof(int), s));
  kband_kernel<<<nb, 256, 0, s>>>(weight, c->Cin, c->Cout, c->kh * c->kw, BN, kchunks, flags);
  int rc = launched("kband_kernel");
  std::vector<int> h(nb, 1);
  if (rc == HESIC_OK && cudaMemcpyAsync(h.data(), flags, nb * sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = HESIC_E_CUDA;
  if (rc == HESIC_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = HESIC_E_CUDA;
  cudaFreeAsync(flags, s);
  if (rc != HESIC_OK) return rc;
  for (int nt = 0; nt < n_tiles; ++nt) {
    int lo = kchunks, hi = 0;
    for (int kc = 0; kc < kchunks; ++kc)
      if (h[nt * kchunks + kc]) { lo = std::min(lo, kc); hi = std::max(hi, kc + 1); }
    if (hi <= lo) { lo = 0; hi = 1; }     // an all-zero tile still needs one step to define the accumulator
    c->kc_lo[nt] = (int8_t)lo; c->kc_hi[nt] = (int8_t)hi;
  }
  c->kband_bn = BN; c->kband_chunks = kchunks;
  return HESIC_OK;
}

extern "C" int hesic_conv_enable_gdn(hesic_conv *c, int enable) {
  HESIC_REQUIRE(c, "hesic_conv_enable_gdn: null conv");
  HESIC_REQUIRE(!enable || c->gdn_beta, "hesic_conv_enable_gdn: no GDN has been packed for this layer");
  c->has_gdn = enable != 0;
  return HESIC_OK;
}

static int conv_forward_any(hesic_conv *c, const hesic_tensor *x, const hesic_tensor *xb, const hesic_tensor *y, int act,
                            int path, void *stream) {
  HESIC_REQUIRE(c && c->loaded, "hesic_conv_forward: weights not loaded");
  HESIC_REQUIRE(c->device == current_device(), "hesic_conv_forward: the layer's operands live on device %d, the current device is %d "
                "(load the weights and launch with the tensors' device current)", c->device, current_device());
  int r;
  if ((r = check_tensor(x, "conv input")) != HESIC_OK) return r;
  if (xb && (r = check_tensor(xb, "conv input (second part)")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "conv output")) != HESIC_OK) return r;
  const int Cin = x->C + (xb ? xb->C : 0);
  HESIC_REQUIRE(Cin == c->Cin, "conv: input has %d channels, layer expects %d", Cin, c->Cin);
  if (xb) HESIC_REQUIRE(xb->B == x->B && xb->H == x->H && xb->W == x->W, "conv: the two input parts differ in shape");
  HESIC_REQUIRE(y->C == c->Cout, "conv: output has %d channels, layer produces %d", y->C, c->Cout);
  HESIC_REQUIRE(y->B == x->B, "conv: batch mismatch");
  int Ho = conv_out_size(c, x->H, c->kh), Wo = conv_out_size(c, x->W, c->kw);
  HESIC_REQUIRE(y->H == Ho && y->W == Wo, "conv: output is %dx%d, expected %dx%d", y->H, y->W, Ho, Wo);
  HESIC_REQUIRE(act >= 0 && act <= 2, "conv: bad activation %d", act);
  cudaStream_t s = as_stream(stream);
  // full-resolution few-channel layers: exact-fp32 stencil (AUTO and SIMT both mean "CUDA cores" here)
  if (path != HESIC_PATH_TCGEN05 && conv_small_supported(c, x, xb, y) && (path == HESIC_PATH_AUTO || xb))
    return conv_forward_small(c, x, xb, y, act, s);
  if (xb) {
    set_error("conv: a two-part (concatenated) input is only supported by the few-channel stencil path");
    return HESIC_E_UNSUPPORTED;
  }
  bool tc_ok = conv_tc_supported(c, x, y);
  if (path == HESIC_PATH_TCGEN05 && !tc_ok) {
    set_error("conv: shape/format not supported by the tcgen05 path");
    return HESIC_E_UNSUPPORTED;
  }
  if (path == HESIC_PATH_TCGEN05 || (path == HESIC_PATH_AUTO && tc_ok)) return conv_forward_tc(c, x, y, act, s);
  static const bool trace = diag_env("HESIC_TRACE_SIMT") != nullptr;   // which layers miss the tensor-core path
  if (trace)
    fprintf(stderr, "[hesic] CUDA-core conv: %d->%d k%dx%d s%d tr%d  in fmt%d %dx%dx%d  out fmt%d %dx%d path=%d\n", c->Cin, c->Cout,
            c->kh, c->kw, c->stride, c->transposed, x->fmt, x->B, x->H, x->W, y->fmt, y->H, y->W, path);
  return conv_forward_simt(c, x, y, act, s);
}

extern "C" int hesic_conv_forward(hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y, int act, int path,
                                  void *stream) {
  return conv_forward_any(c, x, nullptr, y, act, path, stream);
}

namespace hesic { int gn_stats_slotted(const hesic_tensor *x, int groups, double *stats, cudaStream_t st); }

// Convolution whose NHWC fp32 output is followed by nn.GroupNorm: also produces the group statistics
// stats[B][groups][HESIC_GN_SLOTS][2] = partial (sum, sum of squares) -- in the tensor-core epilogue where the tile
// geometry allows (no second pass over the output), else with the statistics kernel.
extern "C" int hesic_conv_forward_gn(hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y, int path, double *stats, int groups,
                                     void *stream) {
  HESIC_REQUIRE(c && y && stats && groups >= 1, "hesic_conv_forward_gn: null argument");
  HESIC_REQUIRE(y->fmt == HESIC_FMT_NHWC_F32 && c->Cout % groups == 0, "hesic_conv_forward_gn: NHWC fp32 output, channels divisible by groups");
  cudaStream_t s = as_stream(stream);
  HESIC_CUDA(cudaMemsetAsync(stats, 0, (size_t)y->B * groups * HESIC_GN_SLOTS * 2 * sizeof(double), s));
  static const bool unfused = diag_env("HESIC_GN_UNFUSED") != nullptr;     // diagnostic: always use the statistics kernel
  c->gn_stats = unfused ? nullptr : stats; c->gn_groups = groups; c->gn_fused = false;
  const int r = conv_forward_any(c, x, nullptr, y, HESIC_ACT_NONE, path, stream);
  const bool fused = c->gn_fused;
  c->gn_stats = nullptr; c->gn_groups = 0; c->gn_fused = false;
  if (r != HESIC_OK || fused) return r;
  return gn_stats_slotted(y, groups, stats, s);
}

extern "C" int hesic_conv_forward_cat(hesic_conv *c, const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y,
                                      int act, int path, void *stream) {
  HESIC_REQUIRE(xb != nullptr, "hesic_conv_forward_cat: null second input");
  return conv_forward_any(c, xa, xb, y, act, path, stream);
}

// Convolution whose NCHW fp32 output is a reconstruction compared with a target image (the MSE term of RateDistortionLoss,
// ywz/mywork/test3real.py:99-111): *sse += sum((y - target)^2) -- in the epilogue of the full-resolution stencil (the
// output is still in registers, no second pass over it), else with the squared-error kernel on the written image.
extern "C" int hesic_conv_forward_sse(hesic_conv *c, const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y, int act,
                                      int path, const hesic_tensor *target, double *sse, void *stream) {
  HESIC_REQUIRE(c && y && target && sse, "hesic_conv_forward_sse: null argument");
  int r;
  if ((r = check_tensor(target, "sse target")) != HESIC_OK) return r;
  HESIC_REQUIRE(y->fmt == HESIC_FMT_NCHW_F32 && target->fmt == HESIC_FMT_NCHW_F32 && same_shape(y, target),
                "hesic_conv_forward_sse: output and target must be NCHW fp32 of the same shape");
  c->sse_target = (const float *)target->p0; c->sse_Cs = target->Cs > 0 ? target->Cs : target->C;
  c->sse_acc = sse; c->sse_fused = false;
  r = conv_forward_any(c, xa, xb, y, act, path, stream);
  const bool fused = c->sse_fused;
  c->sse_target = nullptr; c->sse_Cs = 0; c->sse_acc = nullptr; c->sse_fused = false;
  if (r != HESIC_OK || fused) return r;
  return hesic_sum_squared_error(y, target, sse, stream);
}

extern "C" int hesic_gdn(const hesic_tensor *x, const hesic_tensor *y, const float *beta, const float *gamma,
                         int inverse, float beta_min, void *stream) {
  int r;
  if ((r = check_tensor(x, "gdn input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "gdn output")) != HESIC_OK) return r;
  HESIC_REQUIRE(same_shape(x, y), "gdn: shape mismatch");
  HESIC_REQUIRE(beta && gamma, "gdn: null parameters");
  cudaStream_t s = as_stream(stream);
  int C = x->C;
  float *buf = nullptr;
  HESIC_CUDA(cudaMallocAsync(&buf, ((size_t)C * C + C) * sizeof(float), s));
  r = pack_gdn(beta, gamma, C, beta_min, buf + (size_t)C * C, buf, nullptr, nullptr, s);
  if (r == HESIC_OK) r = gdn_simt(x, y, buf + (size_t)C * C, buf, inverse, s);
  cudaFreeAsync(buf, s);
  return r;
}
