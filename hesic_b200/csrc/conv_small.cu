// Full-resolution few-channel 5x5 layers on CUDA cores: Encoder2.pre_conv (+pre_gdn) and
// Decoder2.after_conv (newnet1.py:629-630,643-644,669-670,685-686), i.e. conv / stride-1 transposed
// conv with Cin <= 8, Cout <= 4, k5, pad 2, evaluated in exact fp32 FMA.
//
// Why not the tensor-core path: with 3 output channels the GEMM has N = 3 (padded to 16) and every
// 128-pixel A tile is streamed through shared memory five times (one 16 KB box per kernel row, hi and
// lo planes) for 0.24 GFLOP per pair -- the layer is bound by L2->SM fills at ~5 TFLOP/s.  As a direct
// stencil the whole layer is 450 FMA per output pixel and one read of the input (24 B) + one write of
// the output (12 B): HBM-bound on paper, FMA-issue-bound in practice.
//
// Block = 64 x 16 output pixels, 128 threads, each thread 2 rows x 4 consecutive pixels x Cout
// accumulators.  The zero-padded input tile (Cin x 20 x 72 fp32, one TMA box per source tensor) and the weights ([ci][ky][kx][4], the
// kernel flipped for the transposed form) sit in shared memory / the kernel parameters; inner loop per (ci, ky): four 16-byte
// input loads feed 120 FMAs whose weight operand is a uniform register.  torch.cat on the reference side is a second
// input pointer (channels [Ca, Cin) come from xb), so the concatenation is never materialised.  The
// output is NCHW fp32 or, when the consumer is the 3->128 tensor-core layer, directly its ROWPAD8
// bf16 (hi, lo) input format.
#include <stdlib.h>
#include <string.h>

#include "conv.h"
#include "tc_ptx.cuh"

namespace hesic {
namespace small {

// the input tile starts FOUR columns left of the output tile (two of halo, two of alignment: a TMA box has to start on a
// 16-byte boundary of the innermost dimension) and is 72 floats wide
constexpr int TW = 64, TH = 16, NT = 128, PITCH = TW + 8, ROWS = TH + 4, XOFF = 4;
constexpr int PLANE = ROWS * PITCH;      // floats of one channel of the input tile

struct Args {
  const float *xa, *xb;   // NCHW fp32; channels [0, Ca) from xa, [Ca, Cin) from xb
  int Ca, CsA, CsB;
  int B, H, W;
  const float *w;         // hesic_conv::w_simt  [(ky*5+kx)*Cin + ci][Cout]
  const float *bias;
  int transposed;
  int gdn;                // 0 none, 1 GDN, 2 inverse GDN over the Cout channels
  const float *beta, *gamma;   // reparametrised beta [Cout]; gamma as [j][i] (hesic_conv::gdn_w_simt)
  int act;
  int out_fmt, out_Cs;    // NCHW fp32 (channel slice of out_Cs) or ROWPAD split (out_Cs = 8 or 4 slots)
  void *y0, *y1;
  // Weights as kernel parameters, [ci][ky][kx][4] with the kernel flipped for the transposed form: the FMA's weight operand
  // is a uniform register filled from the constant bank (LDCU), no shared-memory weight loads.  r03 measurements of the
  // alternatives (all bit-identical): weights in shared memory 142-155 us; this form 135-140 us; packed fma.rn.f32x2 with
  // pixel pairs 155-159 us (one MOV per packed FMA to build the unaligned pairs); packed FMAs over row pairs interleaved in
  // shared memory, no MOVs, 182-185 us.  (B200 issues three-source FFMAs at full rate, tools/micro/ffma_rate.cu: what held
  // this kernel back was the tile fill, see `tma` below.)
  float wk[8 * 25 * 4];
  // tile fill by TMA (r03): one box [channels][ROWS][PITCH] per source tensor, out-of-image elements zero-filled by the
  // copy engine.  The per-element cp.async loop it replaces (bounds tests, 64-bit addresses, predicates) was 40 % of the
  // kernel's executed instructions (ncu source page: FFMA 46 %).  gap_b: floats between the end of xa's box and the
  // (128-byte aligned) start of xb's.
  int tma, gap_b;
  // squared error of the NCHW output against a target image (MSE partial of RateDistortionLoss, test3real.py:99-111),
  // accumulated from the registers the output is stored from: one atomic per block
  const float *target;
  int tgt_Cs;
  double *sse;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(NT) conv_small_kernel(const __grid_constant__ Args a, const __grid_constant__ CUtensorMap map_a,
                                                       const __grid_constant__ CUtensorMap map_b) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long tile_bar;
  float *in = sm;                        // [CIN][ROWS][PITCH] (+ gap_b floats before xb's channels)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;

  if (a.tma) {
    const uint32_t bar = tc::smem_u32(&tile_bar), dst = tc::smem_u32(in);
    if (tid == 0) {
      tc::mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      tc::mbar_expect_tx(bar, (uint32_t)(CIN * PLANE * sizeof(float)));
      tc::tma_load_4d(&map_a, dst, bar, x0 - XOFF, y0 - 2, 0, b);
      if (a.Ca < CIN) tc::tma_load_4d(&map_b, dst + (uint32_t)((a.Ca * PLANE + a.gap_b) * sizeof(float)), bar, x0 - XOFF, y0 - 2, 0, b);
    }
    __syncthreads();
    tc::mbar_wait(bar, 0, 20);
  } else
  // input tile: one warp per (channel, row), lanes along x.  The tile starts at x0 - 4 (even), so with an even
  // image width every float2 is 8-byte aligned and entirely inside or outside the image.
  {
    const int warp = tid >> 5, lane = tid & 31;
    const bool vec2 = (a.W & 1) == 0 && ((((uintptr_t)a.xa) | ((uintptr_t)a.xb)) & 7u) == 0;
    for (int cr = warp; cr < CIN * ROWS; cr += NT / 32) {
      const int c = cr / ROWS, r = cr - c * ROWS;
      const int gy = y0 - 2 + r;
      float *dst = in + (size_t)cr * PITCH;
      const float *src = nullptr;
      if (gy >= 0 && gy < a.H)
        src = (c < a.Ca ? a.xa + ((size_t)b * a.CsA + c) * a.H * a.W : a.xb + ((size_t)b * a.CsB + (c - a.Ca)) * a.H * a.W) +
              (size_t)gy * a.W;
      const float *safe = a.xa;   // any valid address for the zero-fill form
      if (vec2) {
        for (int col = 2 * lane; col < PITCH; col += 64) {
          const int gx = x0 - XOFF + col;
          const bool ok = src && gx >= 0 && gx < a.W;
          cp_async<8>(dst + col, ok ? src + gx : safe, ok);
        }
      } else {
        for (int col = lane; col < PITCH; col += 32) {
          const int gx = x0 - XOFF + col;
          const bool ok = src && gx >= 0 && gx < a.W;
          cp_async<4>(dst + col, ok ? src + gx : safe, ok);
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();
  }

  float acc[2][4][COUT];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int co = 0; co < COUT; ++co) acc[r][p][co] = 0.f;

#pragma unroll 1
  for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      float v[2][8];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float *row = in + ((size_t)ci * ROWS + ty + 8 * r + ky) * PITCH + 4 * tx + (ci >= a.Ca ? a.gap_b : 0);
        // columns [4 tx + 2, 4 tx + 10) of the tile (XOFF - 2 = 2 floats into the row): 8 + 16 + 8 bytes
        const float2 q0 = *reinterpret_cast<const float2 *>(row + 2), q2 = *reinterpret_cast<const float2 *>(row + 8);
        const float4 q1 = *reinterpret_cast<const float4 *>(row + 4);
        v[r][0] = q0.x; v[r][1] = q0.y; v[r][2] = q1.x; v[r][3] = q1.y;
        v[r][4] = q1.z; v[r][5] = q1.w; v[r][6] = q2.x; v[r][7] = q2.y;
      }
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const float *wq = a.wk + ((ci * 5 + ky) * 5 + kx) * 4;
        const float wv[4] = {wq[0], wq[1], wq[2], wq[3]};
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[r][p][co] = fmaf(v[r][p + kx], wv[co], acc[r][p][co]);
      }
    }
  }

  float bia[COUT], bet[COUT], gam[COUT][COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) {
    bia[c] = __ldg(a.bias + c);
    bet[c] = a.gdn ? __ldg(a.beta + c) : 1.f;
#pragma unroll
    for (int j = 0; j < COUT; ++j) gam[j][c] = a.gdn ? __ldg(a.gamma + j * COUT + c) : 0.f;
  }
  double se = 0.0;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int oy = y0 + ty + 8 * r;
    if (oy >= a.H) continue;
    float o[COUT][4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float x[COUT];
#pragma unroll
      for (int c = 0; c < COUT; ++c) x[c] = acc[r][p][c] + bia[c];
      if (a.gdn) {
        float sq[COUT], t[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) sq[c] = x[c] * x[c];
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          float nrm = bet[c];
#pragma unroll
          for (int j = 0; j < COUT; ++j) nrm = fmaf(gam[j][c], sq[j], nrm);
          t[c] = x[c] * (a.gdn == 2 ? sqrtf(nrm) : rsqrtf(nrm));
        }
#pragma unroll
        for (int c = 0; c < COUT; ++c) x[c] = t[c];
      }
#pragma unroll
      for (int c = 0; c < COUT; ++c) o[c][p] = apply_act(x[c], a.act);
    }
    const int ox = x0 + 4 * tx;
    if (ox >= a.W) continue;
    if (a.out_fmt == HESIC_FMT_NCHW_F32) {
      float se_row = 0.f;
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        float *dst = (float *)a.y0 + (((size_t)b * a.out_Cs + c) * a.H + oy) * a.W + ox;
        if (ox + 3 < a.W && (((uintptr_t)dst) & 15u) == 0) {
          *reinterpret_cast<float4 *>(dst) = make_float4(o[c][0], o[c][1], o[c][2], o[c][3]);
        } else {
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (ox + p < a.W) dst[p] = o[c][p];
        }
        if (a.sse) {
          const float *tg = a.target + (((size_t)b * a.tgt_Cs + c) * a.H + oy) * a.W + ox;
          float t[4];
          if (ox + 3 < a.W && (((uintptr_t)tg) & 15u) == 0) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(tg));
            t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
          } else {
#pragma unroll
            for (int p = 0; p < 4; ++p) t[p] = ox + p < a.W ? __ldg(tg + p) : o[c][p];
          }
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float d = o[c][p] - t[p];
            se_row = fmaf(d, d, se_row);
          }
        }
      }
      se += (double)se_row;   // fp32 over the row's 4 * COUT squares, fp64 across rows and threads
    } else {   // ROWPAD split: one store of all channel slots per pixel and plane
      TView yv;
      yv.p0 = a.y0; yv.p1 = a.y1; yv.fmt = HESIC_FMT_ROWPAD8_SPLIT; yv.B = a.B; yv.C = COUT; yv.H = a.H; yv.W = a.W; yv.Cs = a.out_Cs;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (ox + p >= a.W) continue;
        float v[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) v[c] = o[c][p];
        store_rowpad_pixel(yv, b, oy, ox + p, v);
      }
    }
  }
  if (a.sse) {   // uniform
    __shared__ double part[NT / 32];
    se = warp_sum(se);
    if ((tid & 31) == 0) part[tid >> 5] = se;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) s += part[w];
      atomicAdd(a.sse, s);
    }
  }
}

template <int CIN, int COUT>
static int launch(Args &a, const hesic_tensor *xa, const hesic_tensor *xb, cudaStream_t s) {
  CUtensorMap ma, mb;
  memset(&ma, 0, sizeof(ma));
  memset(&mb, 0, sizeof(mb));
  a.tma = 0; a.gap_b = 0;
  static const bool tma_on = diag_env("HESIC_SMALL_NO_TMA") == nullptr;
  const int Cb = CIN - a.Ca;
  if (tma_on && a.W % 4 == 0 && ((uintptr_t)a.xa & 15u) == 0 && (!xb || ((uintptr_t)a.xb & 15u) == 0) && (Cb == 0 || xb)) {
    auto mk = [&](CUtensorMap *m, const float *base, int C, int Cs) {
      uint64_t dims[4] = {(uint64_t)a.W, (uint64_t)a.H, (uint64_t)C, (uint64_t)a.B};
      uint64_t strides[3] = {(uint64_t)a.W * 4, (uint64_t)a.H * a.W * 4, (uint64_t)Cs * a.H * a.W * 4};
      uint32_t box[4] = {(uint32_t)PITCH, (uint32_t)ROWS, (uint32_t)C, 1u};
      return tc::make_tensor_map(m, base, 4, dims, strides, box, true, false);
    };
    int r = mk(&ma, a.xa, a.Ca, a.CsA);
    if (r == HESIC_OK && Cb > 0) r = mk(&mb, a.xb, Cb, a.CsB);
    if (r == HESIC_OK) {
      a.tma = 1;
      const int bytes_a = a.Ca * PLANE * (int)sizeof(float);
      a.gap_b = Cb > 0 ? ((bytes_a + 127) / 128 * 128 - bytes_a) / (int)sizeof(float) : 0;
    }
  }
  const int smem = (CIN * PLANE + a.gap_b) * (int)sizeof(float);
  static PerDeviceOnce once;
  if (once.first())
    HESIC_CUDA(cudaFuncSetAttribute(conv_small_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (CIN * PLANE + 32) * (int)sizeof(float)));
  dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, a.B);
  conv_small_kernel<CIN, COUT><<<grid, NT, smem, s>>>(a, ma, mb);
  HESIC_LAUNCHED("conv_small_kernel");
  return HESIC_OK;
}

}  // namespace small

bool conv_small_supported(const hesic_conv *c, const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y) {
  if (c->kh != 5 || c->kw != 5 || c->pad != 2 || c->stride != 1 || c->out_pad != 0) return false;
  if (!((c->Cin == 6 && c->Cout == 3) || (c->Cin == 3 && c->Cout == 3))) return false;
  if (xa->fmt != HESIC_FMT_NCHW_F32 || (xb && xb->fmt != HESIC_FMT_NCHW_F32)) return false;
  if (y->fmt == HESIC_FMT_ROWPAD8_SPLIT) {
    if ((((uintptr_t)y->p0 | (uintptr_t)y->p1) & 15u) != 0 || c->Cout > y->Cs) return false;
  } else if (y->fmt != HESIC_FMT_NCHW_F32) {
    return false;
  }
  return (int64_t)y->B * y->H * y->W > 0 && y->B <= 65535 && (y->H + small::TH - 1) / small::TH <= 65535;
}

int conv_forward_small(hesic_conv *c, const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y, int act,
                       cudaStream_t s) {
  small::Args a;
  a.xa = (const float *)xa->p0; a.Ca = xa->C; a.CsA = xa->Cs > 0 ? xa->Cs : xa->C;
  a.xb = xb ? (const float *)xb->p0 : nullptr; a.CsB = xb ? (xb->Cs > 0 ? xb->Cs : xb->C) : 0;
  a.B = y->B; a.H = y->H; a.W = y->W;
  a.w = c->w_simt; a.bias = c->bias; a.transposed = c->transposed;
  if (!c->w_host) { set_error("conv stencil: weights not loaded"); return HESIC_E_INVALID; }
  memset(a.wk, 0, sizeof(a.wk));
  for (int ci = 0; ci < c->Cin; ++ci)
    for (int ky = 0; ky < 5; ++ky)
      for (int kx = 0; kx < 5; ++kx) {
        const int tap = c->transposed ? (4 - ky) * 5 + (4 - kx) : ky * 5 + kx;
        for (int co = 0; co < c->Cout; ++co) a.wk[((ci * 5 + ky) * 5 + kx) * 4 + co] = c->w_host[(size_t)(tap * c->Cin + ci) * c->Cout + co];
      }
  a.gdn = c->has_gdn ? (c->gdn_inverse ? 2 : 1) : 0;
  a.beta = c->gdn_beta; a.gamma = c->gdn_w_simt;
  if (c->has_gdn && act != HESIC_ACT_NONE) { set_error("activation after fused GDN is not supported"); return HESIC_E_UNSUPPORTED; }
  a.act = act;
  a.out_fmt = y->fmt; a.out_Cs = y->Cs > 0 ? y->Cs : y->C; a.y0 = y->p0; a.y1 = y->p1;
  a.target = nullptr; a.tgt_Cs = 0; a.sse = nullptr;
  if (c->sse_acc && y->fmt == HESIC_FMT_NCHW_F32) {
    a.target = c->sse_target; a.tgt_Cs = c->sse_Cs; a.sse = c->sse_acc;
    c->sse_fused = true;
  }
  if (c->Cin == 6) return small::launch<6, 3>(a, xa, xb, s);
  return small::launch<3, 3>(a, xa, xb, s);
}

}  // namespace hesic
