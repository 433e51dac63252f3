// Convolution plan shared by the SIMT and tcgen05 paths.
#pragma once
#include "common.cuh"

// Operand formulation of the tcgen05 path (decided by the layer geometry at creation):
//   GENERIC  one GEMM k-step per (tap, 64-channel chunk); weights [tap][CoutPad][Cin]
//   ROW      Cin <= 8, k5 (the full-resolution RGB / 6-channel layers): input in ROWPAD8 format, one k-step
//            per kernel ROW with K = 8 pixels x 8 channel slots; weights [ky][CoutPad][64]
//   SCATTER  stride-2 k5 transposed conv with Cout <= 4 (the RGB synthesis head): one GEMM per input pixel onto
//            its 5x5xCout output patch, overlap-add in the epilogue (conv_head.cuh); weights [NPAD = 25*Cout
//            rounded up to 16][Cin], row n = (ky*5 + kx)*Cout + co
//   ROW2     Cin <= 4, k5, stride 2 (the first analysis layer, RGB -> N): input in ROWPAD format with 4 channel
//            slots; one k-step per PAIR of kernel rows, K = 2 rows x 8 pixels x 4 slots, 3 k-steps instead of 5;
//            weights [3][CoutPad][64] stay resident in shared memory for the whole kernel
constexpr int HESIC_GN_SLOTS = 8;   // partial-sum slots per (image, group): spreads the epilogues' atomics
enum { HESIC_TC_GENERIC = 0, HESIC_TC_ROW = 1, HESIC_TC_SCATTER = 2, HESIC_TC_ROW2 = 3 };

struct hesic_conv {
  int Cin = 0, Cout = 0, kh = 0, kw = 0, stride = 1, pad = 0, transposed = 0, out_pad = 0;
  // SIMT operand: fp32 [kh*kw*Cin][Cout] (tap-major rows, Cout contiguous)
  float *w_simt = nullptr;
  float *bias = nullptr;   // [Cout] (zeros when the layer has no bias)
  // host copy of w_simt for the few-channel stencil layers (Cin <= 8, Cout <= 4, k5): conv_small_kernel takes its weights
  // as kernel parameters
  float *w_host = nullptr;
  // tcgen05 operands: bf16 hi / lo planes [kh*kw][CoutPad][Cin] (K-major per tap)
  __nv_bfloat16 *w_hi = nullptr, *w_lo = nullptr;
  int CoutPad = 0;
  int tc_kind = HESIC_TC_GENERIC, tc_taps = 0, tc_k = 0;   // w_hi/w_lo are [tc_taps][CoutPad][tc_k]
  bool loaded = false;
  int device = -1;   // device that owns the packed operands below (the device current at the last hesic_conv_load)
  // MaskedConv2d: bit t set = tap t (ky*kw + kx) has a non-zero mask entry; dead taps are skipped by the tcgen05 path
  // (mask type 'A' of a 5x5 kernel keeps 12 of 25 taps, compressai/layers/layers.py:36-40)
  uint64_t live_taps = ~0ull;
  // block-banded weights (hesic_conv_detect_kband): per 128-wide N tile the range of 64-channel K chunks with non-zero
  // weights; valid for the tile geometry (kband_bn, kband_chunks) it was computed for, 0 = not computed
  int kband_bn = 0, kband_chunks = 0;
  int8_t kc_lo[8] = {0}, kc_hi[8] = {0};
  // fused GDN
  bool has_gdn = false;
  int gdn_inverse = 0;
  float *gdn_beta = nullptr;    // reparametrised beta [Cout]
  float *gdn_w_simt = nullptr;  // fp32 [Cout(j)][Cout(i)] = gamma[i][j]
  __nv_bfloat16 *gdn_g_hi = nullptr, *gdn_g_lo = nullptr;  // bf16 planes [Cout(i)][Cout(j)] (K-major)
  // GroupNorm statistics fused into the epilogue (hesic_conv_forward_gn): set for the duration of one launch
  double *gn_stats = nullptr;   // [B][groups][HESIC_GN_SLOTS][2] partial (sum, sum of squares), zeroed by the caller
  int gn_groups = 0;
  bool gn_fused = false;        // out: the launched kernel accumulated the statistics
  // squared error against a target image fused into the epilogue (hesic_conv_forward_sse): set for the duration of one launch
  const float *sse_target = nullptr;   // NCHW fp32, the output's shape; channel stride of an image = sse_Cs channels
  int sse_Cs = 0;
  double *sse_acc = nullptr;    // += sum((y - target)^2), fp64
  bool sse_fused = false;       // out: the launched kernel accumulated the sum
  // tcgen05 path: cached TMA tensor maps of the static operands (w_hi, w_lo, gamma_hi, gamma_lo)
  unsigned char *tc_maps = nullptr;
  int tc_maps_bn = 0, tc_maps_gdn = -1;
};

namespace hesic {
int conv_out_size(const hesic_conv *c, int in, int k);
int conv_forward_simt(hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y, int act, cudaStream_t s);
int conv_forward_tc(hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y, int act, cudaStream_t s);
bool conv_tc_supported(const hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y);
// conv_small.cu: exact-fp32 CUDA-core stencil for the full-resolution 6->3 / 3->3 k5 s1 layers; xb (may be
// null) supplies the channels after xa's (torch.cat on the reference side)
bool conv_small_supported(const hesic_conv *c, const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y);
int conv_forward_small(hesic_conv *c, const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y, int act,
                       cudaStream_t s);
int gdn_simt(const hesic_tensor *x, const hesic_tensor *y, const float *beta_rp, const float *w_simt, int inverse,
             cudaStream_t s);
}  // namespace hesic
