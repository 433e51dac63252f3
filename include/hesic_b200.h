/* hesic_b200 -- C ABI of the B200-native HESIC stereo-compression forward path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The
 * Python host side (hesic_b200/compat: `compressai.*`, `newnet1`, `newnet1_joint`,
 * `newnet9`, `kornia` stand-in) binds it with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.  Each entry point names the reference
 * interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *  - every pointer marked "dev" is a CUDA device pointer on the current device;
 *    "host" pointers are ordinary host memory;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    all device entry points are asynchronous on that stream;
 *  - return value 0 = success, <0 = error (HESIC_E_*); hesic_last_error() gives the
 *    message for the calling thread.  There is NO CPU fallback: calling a device entry
 *    point without a usable sm_100 device returns HESIC_E_CUDA.
 */
#ifndef HESIC_B200_H_
#define HESIC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HESIC_ABI_VERSION 1

enum {
  HESIC_OK = 0,
  HESIC_E_INVALID = -1,   /* bad argument (maps to ValueError on the Python side) */
  HESIC_E_CUDA = -2,      /* CUDA runtime / driver error, or no sm_100 device      */
  HESIC_E_UNSUPPORTED = -3,
  HESIC_E_OVERFLOW = -4   /* output buffer too small                               */
};

/* Activation layouts in HBM.  SPLIT is the native inter-layer format of the conv stack:
 * two bf16 planes (hi, lo) with value = float(hi) + float(lo), i.e. a 16-bit-mantissa
 * decomposition of the fp32 value that feeds the bf16x3 tcgen05 path without conversion. */
enum { HESIC_FMT_NCHW_F32 = 0, HESIC_FMT_NHWC_F32 = 1, HESIC_FMT_NHWC_SPLIT = 2, HESIC_FMT_ROWPAD8_SPLIT = 3,
       HESIC_FMT_NHWC_HILO = 4 };
/* NHWC_HILO is the activation format of the enhancement network (32 channels at full resolution): ONE
 * bf16 buffer [B][H][W][2*Cs] holding, per pixel, the Cs 'hi' values followed by the Cs 'lo' values
 * (Cs = 32: 128 bytes per pixel = one 128B-swizzled shared-memory row, so a TMA box of pixels is directly
 * the K-major A operand [hi | lo] of the bf16x3 contraction).  p1 is unused. */
/* ROWPAD8_SPLIT is the input format of the full-resolution edge layers (Cin <= 8: the RGB images and
 * the 6-channel concatenations of newnet1.py:643,686): bf16 (hi, lo) planes of
 * [B][H + HESIC_ROWPAD_Y][W + HESIC_ROWPAD_X][8] (Cs = 8 channel slots), image at row/column offset 2, border
 * and unused channels zero.  Cs = 4 (<= 4 channels, even H; input of the 3 -> N stride-2 first layer) interleaves
 * rows in pairs: [B][(H + 4) / 2][W + 8][2][4], so 8 pixels x 2 rows x 4 slots are 64 contiguous elements.  One TMA box row of 64 elements = the 8 pixels x 8 channels a 5-tap kernel row
 * needs, so the tensor-core path reads it as an implicit im2col with K = 64 per kernel row. */
#define HESIC_ROWPAD_Y 4
#define HESIC_ROWPAD_X 8

typedef struct {
  void *p0;        /* dev: float data, or bf16 'hi' plane, already offset to channel 0 of the view */
  void *p1;        /* dev: bf16 'lo' plane (SPLIT), else NULL                                        */
  int32_t fmt;     /* HESIC_FMT_*                                                                    */
  int32_t B, C, H, W;
  int32_t Cs;      /* channels of the underlying buffer (>= C): lets a producer write a channel
                      slice of a concatenation buffer (torch.cat on the reference side)            */
} hesic_tensor;

enum { HESIC_ACT_NONE = 0, HESIC_ACT_RELU = 1, HESIC_ACT_LEAKY_RELU = 2 /* slope 0.01 */ };
enum { HESIC_PATH_AUTO = 0, HESIC_PATH_SIMT = 1, HESIC_PATH_TCGEN05 = 2 };

int hesic_abi_version(void);
const char *hesic_last_error(void);
/* 0 when a sm_100 device is current and the kernels can launch; fills name (may be NULL). */
int hesic_device_check(char *name, int name_len);
/* number of kernels this library launched since the last reset (bench.py "gpu_launches") */
int64_t hesic_launch_count(int reset);

/* ---------------------------------------------------------------------------------------------
 * Convolution / transposed convolution (+ bias, activation, optional fused GDN).
 * Replaces nn.Conv2d / nn.ConvTranspose2d as built by compressai/models/utils.py:104-118
 * (`conv`, `deconv`), MaskedConv2d (compressai/layers/layers.py:21-45) and the GDN that follows
 * them in newnet1.py:580-692 (compressai/layers/gdn.py:55-70).
 */
typedef struct hesic_conv hesic_conv;

hesic_conv *hesic_conv_create(int Cin, int Cout, int kh, int kw, int stride, int pad, int transposed,
                              int output_padding);
void hesic_conv_destroy(hesic_conv *c);
/* weight: dev fp32 in the reference layout (conv [Cout,Cin,kh,kw]; transposed [Cin,Cout,kh,kw]);
 * bias: dev fp32 [Cout] or NULL; mask: dev fp32 like weight or NULL (MaskedConv2d). Packs the
 * operand planes the kernels read (fp32 tap-major for the SIMT path, bf16 hi/lo K-major per tap for
 * tcgen05). */
int hesic_conv_load(hesic_conv *c, const float *weight, const float *bias, const float *mask, void *stream);
/* Attach a GDN/IGDN to the conv epilogue. beta [Cout], gamma [Cout,Cout] are the RAW parameters of
 * compressai.layers.GDN; the non-negative reparametrisation (ops/parametrizers.py:41-44) is applied
 * on the device.  Pass NULLs to detach. */
int hesic_conv_set_gdn(hesic_conv *c, const float *beta, const float *gamma, int inverse, float beta_min,
                       void *stream);
/* Block-banded layers (the nn.Conv3d of the DSIC cost volumes evaluated as a 2-D convolution over the stacked
 * (depth, feature) channels, ywz/DSIC/mynet6_plus.py:273-283): finds, per 128-wide tile of output channels, the range of
 * 64-channel input chunks that hold any non-zero weight, and the tensor-core path then skips the rest.  weight: the dev
 * fp32 tensor last given to hesic_conv_load.  Synchronises the stream (a few flags are read back); optional. */
int hesic_conv_detect_kband(hesic_conv *c, const float *weight, void *stream);
/* Detach (0) / re-attach (1) the GDN packed by the last hesic_conv_set_gdn without re-packing it: the same layer
 * object serves the fused engine (GDN in the epilogue) and stand-alone operator calls (plain convolution). */
int hesic_conv_enable_gdn(hesic_conv *c, int enable);
int hesic_conv_forward(hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y, int act, int path,
                       void *stream);
/* Same layer applied to torch.cat((xa, xb), dim=1) without materialising the concatenation
 * (newnet1.py:643-644 pre_conv(cat(x1_warp, x2)), :686 after_conv(cat(.., x1_hat_warp))): the
 * layer's input channels [0, xa->C) are read from xa, the rest from xb.  NCHW fp32 inputs,
 * full-resolution few-channel k5 s1 layers only (else HESIC_E_UNSUPPORTED). */
int hesic_conv_forward_cat(hesic_conv *c, const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y,
                           int act, int path, void *stream);
/* The layers that emit a reconstruction (decoder1.g_s_conv4 -> x1_hat, newnet1.py:612; decoder2.after_conv -> x2_hat,
 * :686), with the MSE partial of RateDistortionLoss (ywz/mywork/test3real.py:99-111) taken from the same epilogue:
 * *sse += sum((y - target)^2) in fp64, y and target NCHW fp32 of one shape; xb may be NULL (no concatenation).  The
 * full-resolution stencil (after_conv) accumulates it from the registers it stores from; the RGB synthesis head (measured
 * slower with the target read in its epilogue) and any other layer run hesic_sum_squared_error on the written output
 * (same result to 1e-7 relative: the stencil sums the 12 squares of a store group in fp32 before the fp64 accumulator). */
int hesic_conv_forward_sse(hesic_conv *c, const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y,
                           int act, int path, const hesic_tensor *target, double *sse, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Enhancement network layers (Independent_EN, ywz/mywork/newnet1.py:272-311,1278-1300; ResidualBlock,
 * compressai/layers/layers.py:125-147): conv3x3 (stride 1, padding 1) with Cin <= 32 at full resolution.
 *   Cout == 32: y is NHWC_HILO;  y = act(conv(x) + bias) [+ res1] [+ res2]   (res*: NHWC_HILO, may be NULL) --
 *               the LeakyReLU and the identity additions of ResidualBlock / Enhancement_Block in the epilogue;
 *   Cout <= 4:  y is NCHW fp32;  y = conv(x) + bias [+ res1]   (res1: NCHW fp32 -- Enhancement.forward's
 *               `self.conv2(out) + identity`, newnet1.py:309-310).
 * x is NHWC_HILO with Cs == 32 (channel slots >= Cin must hold zeros).  weight: dev fp32 [Cout,Cin,3,3]. */
typedef struct hesic_en_conv hesic_en_conv;
hesic_en_conv *hesic_en_conv_create(int Cin, int Cout);
void hesic_en_conv_destroy(hesic_en_conv *c);
int hesic_en_conv_load(hesic_en_conv *c, const float *weight, const float *bias, void *stream);
int hesic_en_conv_forward(hesic_en_conv *c, const hesic_tensor *x, const hesic_tensor *y, int act,
                          const hesic_tensor *res1, const hesic_tensor *res2, void *stream);
/* torch.cat((xa, xb), dim=-3) (newnet1.py:302; xb may be NULL) written as NHWC_HILO with 32 channel slots,
 * zeros above xa->C + xb->C: the input of Enhancement.conv1. */
int hesic_en_pack_input(const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y, void *stream);

/* Watchdog of the tcgen05 path: every in-kernel barrier wait is time-bounded, so a protocol error
 * cannot hang the GPU.  Returns 0 when no wait has timed out since the last call, else HESIC_E_CUDA
 * (details in hesic_last_error()).  Synchronises the device; meant for tests and smoke checks. */
int hesic_tc_status(void);

/* Stand-alone GDN (compressai/layers/gdn.py:55-70) with raw parameters, any C. */
int hesic_gdn(const hesic_tensor *x, const hesic_tensor *y, const float *beta, const float *gamma, int inverse,
              float beta_min, void *stream);

/* ---------------------------------------------------------------------------------------------
 * kornia.warp_perspective(src, M, dsize) as called at newnet1.py:746,753,767,1287,1291:
 * dst(x,y) = bilinear(src, M^-1 (x,y,1)), zero padding, kornia's normalise/invert/denormalise
 * arithmetic, align_corners selectable (1 = the convention used throughout this repository).
 * M: dev fp32 [B,3,3].  src/dst: NCHW fp32 (dst may be a channel slice).  dst_rowpad (may be NULL): the
 * same result written a second time in ROWPAD format, for the warped image that feeds the next 3->128
 * layer (newnet1.py:753-754). */
int hesic_warp_perspective(const hesic_tensor *src, const float *M, const hesic_tensor *dst,
                           const hesic_tensor *dst_rowpad, int align_corners, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Homography front-end glue (the step that produces h_matrix, ywz/mywork/test3real.py:171-181):
 * kornia.get_perspective_transform(src, dst) -- the 4-point direct linear transform, test3real.py:179,
 * udh/udh/model.py:27 -- optionally followed by torch.inverse (test3real.py:180).  src, dst: dev fp32
 * [B,4,2] corner coordinates; H: dev fp32 [B,3,3]; invert != 0 writes H^-1 (H rounded to fp32 first, as on
 * the reference path).  Solved per pair in fp64 with partial pivoting; degenerate corners give inf/nan
 * entries (torch.linalg.solve raises instead).
 * hesic_max_pool2x2: nn.MaxPool2d(2, 2) of udh/udh/model.py:66 on dense NCHW fp32, y = [B,C,H/2,W/2]. */
int hesic_perspective_transform(const float *src, const float *dst, int B, int invert, float *H, void *stream);
int hesic_max_pool2x2(const hesic_tensor *x, const hesic_tensor *y, void *stream);

/* ---------------------------------------------------------------------------------------------
 * EntropyBottleneck.forward, eval mode (compressai/entropy_models/entropy_models.py:384-411,
 * 350-382): z_hat = round(z - median) + median; likelihood = |sigmoid(s*u) - sigmoid(s*l)| clamped.
 * params: dev fp32, per channel 58 floats packed as
 *   [softplus(M0) 3 | b0 3 | tanh(f0) 3 | softplus(M1) 9 | b1 3 | tanh(f1) 3 | M2 9 | b2 3 | f2 3 |
 *    M3 9 | b3 3 | f3 3 | softplus(M4) 3 | b4 1 | median 1 | pad 1]  (hesic_eb_pack builds it from the
 * raw _matrices/_biases/_factors/quantiles tensors).
 * z_hat / lik may be NULL-p0 tensors to skip an output.  log2_sum: dev double[1] accumulator for
 * sum(log2 lik) (the bpp partial of ywz/mywork/test3real.py:115-122) or NULL. */
#define HESIC_EB_PARAMS_PER_CHANNEL 60
int hesic_eb_pack(const float *const *matrices /*5*/, const float *const *biases /*5*/,
                  const float *const *factors /*4*/, const float *quantiles, int C, float *params_out,
                  void *stream);
int hesic_entropy_bottleneck(const hesic_tensor *z, const float *params, float likelihood_bound,
                             const hesic_tensor *z_hat, const hesic_tensor *lik, double *log2_sum, void *stream);

/* GaussianMixtureConditional.forward, eval mode (entropy_models.py:661-702): y_hat = round(y);
 * lik = sum_k w[k*M+m] * (Phi((.5-|y_hat-mu|)/s) - Phi((-.5-|y_hat-mu|)/s)), s = max(sigma, bound).
 * scales/means: [B, K*M, H, W]; weights: dev fp32 [B, K*M].  K = 1 with weights == NULL is
 * GaussianConditional.forward (entropy_models.py:528-554): y_hat = round(y - mu) + mu (mu may be
 * absent: means->p0 == NULL).  y_hat_split (may be NULL): a second copy of y_hat as SPLIT planes, the
 * input format of the synthesis stack that consumes it (newnet1.py:744,762). */
int hesic_gaussian_conditional(const hesic_tensor *y, const hesic_tensor *scales, const hesic_tensor *means,
                               const float *weights, int K, int mixture, float scale_bound,
                               float likelihood_bound, const hesic_tensor *y_hat, const hesic_tensor *lik,
                               const hesic_tensor *y_hat_split, double *log2_sum, void *stream);

/* spatial_pool2d (newnet1.py:441-453) + LeakyReLU + conv1x1(K*M -> K*M) + softmax over the K
 * components (newnet1.py:498-512,572-574).  x: [B, K*M, H, W]; w1x1: dev fp32 [K*M, K*M] (the
 * reference's [Cout,Cin,1,1]); out: dev fp32 [B, K*M].  pooled (dev [B,K*M], may be NULL) receives
 * the raw spatial maximum. */
int hesic_spatial_max(const hesic_tensor *x, float *out_max, void *stream);
int hesic_mixture_weights(const float *pooled, const float *w1x1, const float *bias, int B, int K, int M,
                          float *out, void *stream);

/* nn.UpsamplingBilinear2d(scale_factor=s) (align_corners=True; newnet1.py:524,564). */
int hesic_upsample_bilinear(const hesic_tensor *x, const hesic_tensor *y, int scale, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Operators of the DSIC variant (ywz/DSIC/mynet6_plus.py).  Each takes the reference's NCHW fp32 tensors, or -- for
 * the fused DSIC engine, which keeps activations channels-last between tensor-core convolutions -- the
 * channels-last forms named below (channel slices of wider buffers allowed, Cs > C).
 * nn.GroupNorm(groups, C, eps, affine) (+ the nn.ReLU that always follows it, mynet6_plus.py:224-238,262-290):
 * statistics over (C/groups, H, W) per sample; weight/bias: dev fp32 [C] or NULL.
 * Formats: NCHW fp32 -> NCHW fp32, or NHWC fp32 -> NHWC_SPLIT / NHWC fp32 (channels per group a multiple of 4). */
int hesic_group_norm(const hesic_tensor *x, const hesic_tensor *y, int groups, const float *weight, const float *bias,
                     float eps, int relu, void *stream);
/* conv -> nn.GroupNorm(+ReLU) (mynet6_plus.py:224-290) without a separate statistics pass: hesic_conv_forward_gn writes the
 * convolution's NHWC fp32 output AND the per-(image, group) partial sums stats[B][groups][8][2] (fp64: sum, sum of squares;
 * accumulated by the tensor-core epilogue where the tile geometry allows, else by the statistics kernel);
 * hesic_group_norm_apply normalises with them (x NHWC fp32 -> y NHWC fp32 or SPLIT, channel slices allowed). */
int hesic_conv_forward_gn(hesic_conv *conv, const hesic_tensor *x, const hesic_tensor *y, int path, double *stats, int groups,
                          void *stream);
int hesic_group_norm_apply(const hesic_tensor *x, const hesic_tensor *y, int groups, const float *weight, const float *bias,
                           float eps, int relu, const double *stats, void *stream);
/* nn.functional.softmax(x, dim=-3): over the disparity channels of a cost volume (mynet6_plus.py:311).
 * NCHW fp32, or NHWC fp32 with C <= 64. */
int hesic_softmax_channels(const hesic_tensor *x, const hesic_tensor *y, void *stream);
/* dense_warp.forward (mynet6_plus.py:316-345): out[b,c,y,x] = sum_{d, x+d<W} cost[b,d,y,x] * h1[b,c,y,x+d].
 * All NCHW fp32, or h1 / out NHWC_SPLIT with an NHWC fp32 cost. */
int hesic_dense_warp(const hesic_tensor *h1, const hesic_tensor *cost, const hesic_tensor *out, void *stream);

/* Layout / format conversion with an optional pointwise op: 0 copy, 1 abs (newnet1.py:435),
 * 2 round-half-even (EntropyModel._quantize 'dequantize', entropy_models.py:98-125). */
enum { HESIC_OP_COPY = 0, HESIC_OP_ABS = 1, HESIC_OP_ROUND = 2 };
int hesic_convert(const hesic_tensor *x, const hesic_tensor *y, int op, void *stream);
/* 8-bit interleaved images -> the path's input tensors: src dev uint8 [B][H][W][C] (as cv2.imread / PIL deliver them and
 * the reference's loader holds them until transforms.ToTensor(), compressai/datasets/utils.py:101-102,
 * ywz/mywork/test3real.py:323), dst NCHW fp32 = float(u8) / 255 -- ToTensor's arithmetic, one IEEE division.  Lets a
 * caller ship 1 byte per sample over PCIe instead of 4. */
int hesic_images_from_u8(const uint8_t *src, int B, int H, int W, int C, const hesic_tensor *dst, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Symbol / CDF-index preparation for the host rANS coder -- integer outputs, bit-exact with
 * EntropyModel.compress (entropy_models.py:165-196): symbols = int32(round_half_even(x - mean)),
 * flattened per image in C,H,W order.  means: NULL, per-channel (dev fp32 [C]) or a full tensor.
 * indexes: EntropyBottleneck._build_indexes (entropy_models.py:413-418): channel id; or
 * GaussianConditional.build_indexes (entropy_models.py:556-562):
 *   (n_table-1) - #{ s in table[:-1] : max(scale, bound) <= s }.
 * out_*: dev int32 [B, C*H*W]. */
int hesic_prepare_symbols(const hesic_tensor *x, const float *channel_means, const hesic_tensor *means,
                          int32_t *out_symbols, void *stream);
int hesic_build_indexes_channel(int B, int C, int H, int W, int32_t *out_indexes, void *stream);
int hesic_build_indexes_scale(const hesic_tensor *scales, const float *table, int n_table, float scale_bound,
                              int32_t *out_indexes, void *stream);

/* Rate-distortion partial sums (RateDistortionLoss, ywz/mywork/test3real.py:90-124):
 * acc[0] += sum((a-b)^2) over all elements (double).  The likelihood kernels above accumulate
 * sum(log2 p) themselves. */
int hesic_sum_squared_error(const hesic_tensor *a, const hesic_tensor *b, double *acc, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Host-side entropy coder (stays on the host, as in the reference).
 * compressai._CXX.pmf_to_quantized_cdf (compressai/cpp_exts/ops/ops.cpp:24-81); cdf_out holds n+1. */
int hesic_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf_out);
/* compressai.ans.RansEncoder.encode_with_indexes / BufferedRansEncoder (rans_interface.cpp:99-204).
 * cdfs: host int32 [n_cdfs, cdf_pitch] dense table.  An encoder object buffers symbols across
 * encode calls until flush (BufferedRansEncoder semantics). */
typedef struct hesic_rans_encoder hesic_rans_encoder;
hesic_rans_encoder *hesic_rans_encoder_create(void);
void hesic_rans_encoder_destroy(hesic_rans_encoder *e);
int hesic_rans_encoder_push(hesic_rans_encoder *e, const int32_t *symbols, const int32_t *indexes, int64_t n,
                            const int32_t *cdfs, int n_cdfs, int cdf_pitch, const int32_t *cdf_sizes,
                            const int32_t *offsets);
/* returns the number of bytes the stream needs; writes it if out_cap is large enough, else
 * HESIC_E_OVERFLOW is NOT raised and nothing is consumed: call again with a larger buffer. */
int64_t hesic_rans_encoder_flush(hesic_rans_encoder *e, uint8_t *out, int64_t out_cap);
/* compressai.ans.RansDecoder (rans_interface.cpp:206-350): set_stream + decode_stream. */
typedef struct hesic_rans_decoder hesic_rans_decoder;
hesic_rans_decoder *hesic_rans_decoder_create(void);
void hesic_rans_decoder_destroy(hesic_rans_decoder *d);
int hesic_rans_decoder_set_stream(hesic_rans_decoder *d, const uint8_t *stream, int64_t nbytes);
int hesic_rans_decoder_decode(hesic_rans_decoder *d, const int32_t *indexes, int64_t n, const int32_t *cdfs,
                              int n_cdfs, int cdf_pitch, const int32_t *cdf_sizes, const int32_t *offsets,
                              int32_t *out_symbols);

/* ---------------------------------------------------------------------------------------------
 * File codec of the stereo models (HSIC.compress / decompress, ywz/mywork/newnet1.py:823-1273).
 * Device side: per latent element, the 16-bit cumulative-frequency row the reference builds on the host in a Python
 * double loop (newnet1.py:934-978): pmf[s] = sum_k w_k (Phi((.5-|s-(mu_k+minmax)|)/sigma_k) - Phi((-.5-|..|)/sigma_k)),
 * s = 0..2*minmax, sigma_k = max(sigma_k, bound); clip to [1/65536, 1]; pmf / sum(pmf) * 65536 (fp32, numpy's pairwise
 * summation order); round half even; cumulative sum.  scales / means: [1, K*M, H, W] (any layout), weights: dev fp32
 * [K*M]; channels: dev int32 [n_channels] (the non-zero channels, newnet1.py:882-886).  out_cdf: dev int32
 * [n_channels*H*W, 2*minmax+2], rows in the coding order (channel, h, w). */
int hesic_gmm_cdf_tables(const hesic_tensor *scales, const hesic_tensor *means, const float *weights, int K, int M,
                         const int32_t *channels, int n_channels, int minmax, float scale_bound, int32_t *out_cdf,
                         void *stream);
/* Host range coder with the calling pattern of the un-vendored PyPI `range_coder` the reference uses there (one
 * symbol per cumulative row, totals need not be powers of two).  Its byte stream is this library's own (carry-less
 * range coder, 64-bit low, byte renormalisation): the reference's package is neither vendored nor pinned.
 * cdfs: host int32 [n, cdf_pitch], row i = cumulative frequencies (cdf_len entries, first 0) of element i. */
typedef struct hesic_range_encoder hesic_range_encoder;
hesic_range_encoder *hesic_range_encoder_create(void);
void hesic_range_encoder_destroy(hesic_range_encoder *e);
int hesic_range_encoder_push(hesic_range_encoder *e, const int32_t *symbols, int64_t n, const int32_t *cdfs, int cdf_pitch,
                             int cdf_len);
/* returns the stream length; writes it (and resets the encoder) if out_cap is large enough */
int64_t hesic_range_encoder_finish(hesic_range_encoder *e, uint8_t *out, int64_t out_cap);
typedef struct hesic_range_decoder hesic_range_decoder;
hesic_range_decoder *hesic_range_decoder_create(const uint8_t *stream, int64_t nbytes);
void hesic_range_decoder_destroy(hesic_range_decoder *d);
int hesic_range_decoder_decode(hesic_range_decoder *d, int64_t n, const int32_t *cdfs, int cdf_pitch, int cdf_len,
                               int32_t *out_symbols);

#ifdef __cplusplus
}
#endif
#endif /* HESIC_B200_H_ */
