// tcgen05 implicit-GEMM convolution / transposed convolution with fused bias, activation and GDN.
//
// GEMM view (one "task" = one 128-pixel x BN-channel output tile of one sub-pixel phase):
//     D[pixel, co] = sum_{tap} sum_{ci} A_tap[pixel, ci] * W_tap[co, ci]
// * A is never materialised (no im2col buffer).  Activations live in HBM as channels-last bf16
//   (hi, lo) planes; for every (tap, 64-channel chunk) one TMA box of bw x bh x bb pixels x 64
//   channels lands in shared memory already in the 128B-swizzled K-major layout tcgen05.mma reads.
//   TMA's out-of-bounds zero fill IS the convolution's zero padding.
//     - stride-1 conv:       tensor map (C, W, 1, H, B), box origin shifted by the tap offset
//     - stride-2 conv:       space-to-depth VIEW of the same memory, (2C, W/2, 2, H/2, B): tap
//                            ky-p = 2*dy+py selects the parity plane py and a stride-1 shift dy
//     - transposed conv s2:  four sub-pixel phases; phase (ry,rx) is a stride-1 conv over the input
//                            with the taps ky = ry+p (mod 2) (9/6/6/4 taps for k5), written to the
//                            interleaved output positions -- no zero-insertion work
// * fp32 parity on bf16 tensor cores: value = hi + lo (16 mantissa bits), and every k-step issues
//   Ah*Wh into a MAIN fp32 TMEM accumulator and Ah*Wl + Al*Wh into a second, SMALL one ("bf16x3").
//   The tensor core's fp32 accumulate truncates (measured: error grows linearly with the number of
//   accumulate steps); keeping the 2^-8-sized correction terms out of the main accumulator cuts its
//   step count 3x, and their own truncation is relative to their small magnitude.  The epilogue adds
//   main + small in fp32.
// * GDN / IGDN is a second contraction over the squared conv output: the epilogue warps square the
//   accumulator tile, write it back to TMEM as a bf16 (hi, lo) A operand, and the MMA warp runs
//   x^2 * gamma^T (A from TMEM, gamma from SMEM) into a norm accumulator; a second epilogue pass
//   forms x * rsqrt(beta + norm) (or * sqrt for IGDN).  The GDN MMAs of tile t are issued in the
//   middle of tile t+1's main loop so the tensor pipe never waits for an epilogue.
// * persistent, warp-specialised CTA (1 per SM): warp 0 = TMA producer, warp 1 = MMA issuer and
//   TMEM owner, warps 2-9 = epilogue (two groups of four, splitting the tile's columns).  TMEM (512 columns): two buffers of {main 128, small 128}.  For
//   GDN the epilogue keeps x in registers, overwrites main with the bf16 x^2 operand (hi 64 + lo 64
//   columns) and the norm accumulates into small.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "conv.h"
#include "tc_ptx.cuh"

namespace hesic {
namespace tc {

constexpr int BK = 64;                  // channels per k-step: 128 B of bf16 = one SW128 row
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 320;     // TMA warp + MMA warp + 2 x 4 epilogue warps
constexpr int EPI_THREADS = 256;
constexpr int MAX_TAPS = 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t ACC_STRIDE = 256, COL_SMALL = 128;   // per-buffer TMEM layout
constexpr int GDN_AT = 6;               // k-step of tile t+1 before which GDN(t) is issued
constexpr int SMEM_LIMIT = 232448;      // 227 KB
constexpr int BAR_BYTES = 256;
constexpr int CHAN_BYTES = 1024;      // bias + beta tables of the current n-tile
constexpr int STAGING_BYTES = 2 * A_TILE_BYTES;   // two [128 px][128 B] swizzled store tiles

struct Tap {
  int8_t dy, dx, py, px;
  int16_t w;
  int16_t pad_;
};

struct Params {
  int tiles_x, tiles_y, tiles_b;
  int lbw, lbh;            // log2 of the box width / height (bw*bh*bb == 128)
  int bw, bh, bb;
  int n_tiles, BN;
  int n_phases, os;        // transposed conv: os = stride, phases = os*os; conv: 1, 1
  int tap_begin[5];
  Tap taps[MAX_TAPS];
  int kchunks;
  int in_Cs;
  int Cout, CoutPad16;
  int out_fmt, out_Cs;
  void *y0, *y1;
  int Hout, Wout, B;
  const float *bias;
  int act;
  int gdn;                 // 0 none, 1 GDN, 2 inverse GDN (128-channel, TMEM contraction)
  const float *beta;
  // planar epilogue (Cout <= 4): NCHW fp32 output, optional sub-pixel phases in N, optional 3-channel GDN
  int planar;              // 0 = channels-last epilogue, 1 = planar
  int pl_phases, pl_os;    // N index = phase*4 + channel; output pixel = q*pl_os + (ry, rx)
  int pl_gdn;              // 0 none, 1 GDN, 2 inverse GDN over the Cout channels (registers)
  const float *pl_gamma;   // fp32 [Cout(j)][Cout(i)] reparametrised (hesic_conv::gdn_w_simt)
  int tma_store;           // 1: epilogue stages 128-byte-wide tiles in smem and writes them with TMA
  int w_resident;          // all weight tiles ([tap][kchunk][hi, lo]) are loaded once and stay in smem
  int n_wtiles;            // taps * kchunks (w_resident)
  int wide_n;              // BN == 128: issue Ah.[Wh; Wl] as one N = 256 MMA
  int16_t kc_lo[8], kc_hi[8];  // per N tile: the 64-channel K chunks [kc_lo, kc_hi) that hold non-zero weights (block-banded layers)
  int gdn_at;              // k-step of tile t+1 before which the GDN MMAs of tile t are issued
  int stg_sets;            // 1 or 2 staging tile pairs for the TMA-store epilogue (2: a pair does not wait for the previous pair's store)
  int stages;
  uint32_t stage_bytes, b_bytes;
  int n_tasks;
  // GroupNorm statistics of the output (nn.GroupNorm follows the conv, mynet6_plus.py:224-290): every epilogue warp adds
  // the (sum, sum of squares) of its 32 pixels x 32 channels to stats[b][group][slot] -- the separate statistics pass
  // over the conv output (one full read of the tensor) disappears.  Requires one image per tile and 32-channel chunks
  // that do not straddle groups.
  double *stats;
  int stat_cpg, stat_groups;
};

// (sum, sum of squares) of v[32] over the warp's valid pixels -> one pair of fp64 atomics per warp and chunk
__device__ __forceinline__ void gn_accumulate(const Params &p, const float (&v)[32], bool valid, int b, int chan0) {
  float s = 0.f, ss = 0.f;
  if (valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) { s += v[j]; ss = fmaf(v[j], v[j], ss); }
  }
  double ds = (double)s, dss = (double)ss;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dss += __shfl_xor_sync(0xffffffffu, dss, o);
  }
  if ((threadIdx.x & 31) == 0) {
    double *d = p.stats + ((((size_t)b * p.stat_groups + chan0 / p.stat_cpg) * HESIC_GN_SLOTS) + (blockIdx.x & (HESIC_GN_SLOTS - 1))) * 2;
    atomicAdd(d, ds);
    atomicAdd(d + 1, dss);
  }
}

struct TaskCoord {
  int mt, ph, nt;
};
// Tasks are dealt round-robin to the persistent CTAs.  The sub-pixel phases of a transposed conv have
// different lengths (9/6/6/4 taps for k5 s2) and the grid size is a multiple of 4, so the phase is rotated
// by the pixel-tile index: every CTA then sees all four phases instead of always the same one.
__device__ __forceinline__ TaskCoord decode_task(const Params &p, int task) {
  TaskCoord t;
  t.nt = task % p.n_tiles;
  int r = task / p.n_tiles;
  t.mt = r / p.n_phases;
  t.ph = (r + t.mt) % p.n_phases;
  return t;
}

// Same step sequence for the producer and the MMA warp.
template <typename ConvStep, typename GdnStep>
__device__ __forceinline__ void walk_schedule(const Params &p, ConvStep &&conv_step, GdnStep &&gdn_step) {
  int lt = 0;
  for (int task = blockIdx.x; task < p.n_tasks; task += gridDim.x, ++lt) {
    TaskCoord tk = decode_task(p, task);
    const int t0 = p.tap_begin[tk.ph], t1 = p.tap_begin[tk.ph + 1];
    const int kc0 = p.kc_lo[tk.nt & 7], kc1 = p.kc_hi[tk.nt & 7];
    const int ksteps = (t1 - t0) * (kc1 - kc0);
    const int g = (p.gdn && lt > 0) ? min(p.gdn_at, ksteps - 1) : -1;
    int ks = 0;
    for (int t = t0; t < t1; ++t) {
      for (int kc = kc0; kc < kc1; ++kc, ++ks) {
        if (ks == g) {
          gdn_step(lt - 1, 0);
          gdn_step(lt - 1, 1);
        }
        conv_step(lt, tk, t, kc, ks == 0, ks == ksteps - 1);
      }
    }
  }
  if (p.gdn && lt > 0) {
    gdn_step(lt - 1, 0);
    gdn_step(lt - 1, 1);
  }
}

__device__ __forceinline__ void store_chunk(const Params &p, const float (&v)[32], size_t pix, int n_base) {
  if (p.out_fmt == HESIC_FMT_NHWC_SPLIT) {
    __nv_bfloat16 *h = (__nv_bfloat16 *)p.y0 + pix * p.out_Cs + n_base;
    __nv_bfloat16 *l = (__nv_bfloat16 *)p.y1 + pix * p.out_Cs + n_base;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (n_base + g * 8 < p.Cout) {
        uint4 hv, lv;
        split_pair(v[g * 8 + 0], v[g * 8 + 1], hv.x, lv.x);
        split_pair(v[g * 8 + 2], v[g * 8 + 3], hv.y, lv.y);
        split_pair(v[g * 8 + 4], v[g * 8 + 5], hv.z, lv.z);
        split_pair(v[g * 8 + 6], v[g * 8 + 7], hv.w, lv.w);
        *reinterpret_cast<uint4 *>(h + g * 8) = hv;
        *reinterpret_cast<uint4 *>(l + g * 8) = lv;
      }
    }
  } else {
    float *o = (float *)p.y0 + pix * p.out_Cs + n_base;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (n_base + g * 4 < p.Cout)
        *reinterpret_cast<float4 *>(o + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
    }
  }
}

// EPI_WARPS = 8: two epilogue groups of four warps, 64 channels per thread (all layers).
// EPI_WARPS = 16 (fused-GDN layers whose K loop is too short to hide the epilogue -- the first analysis layer): four
// groups, 32 channels per thread.  r02 profile of that layer (profiles/r02_ncu_conv1_stalls.md): the epilogue warps issue
// one instruction every ~7 cycles (fixed-latency dependencies, TMEM / shared-memory round trips, instruction fetch) and
// with two warps per scheduler nothing covers those gaps; twice the warps at half the registers do.
template <int EPI_WARPS>
__device__ __forceinline__ void epi_bar_all() {
  if (EPI_WARPS == 16) asm volatile("bar.sync 1, 512;" ::: "memory");
  else asm volatile("bar.sync 1, 256;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_half(int half) {
  if (half == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
  else asm volatile("bar.sync 3, 256;" ::: "memory");
}

template <int EPI_WARPS>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
               const __grid_constant__ CUtensorMap map_y0, const __grid_constant__ CUtensorMap map_y1,
               const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t w_base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // resident weights (w_resident)
  const uint32_t smem_base = w_base + (p.w_resident ? (uint32_t)p.n_wtiles * 2u * p.b_bytes : 0u);
  const uint32_t b_off = p.w_resident ? 0u : 2u * A_TILE_BYTES;             // B operand inside a stage
  const uint32_t stg_base = smem_base + (uint32_t)p.stages * p.stage_bytes;   // epilogue store staging (tma_store)
  const uint32_t bar_base = stg_base + (p.tma_store ? (uint32_t)(STAGING_BYTES * p.stg_sets) : 0u);
  // barrier map (8 B each)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto acc_full = [&](int b) { return bar_base + 128u + 8u * b; };
  auto acc_empty = [&](int b) { return bar_base + 144u + 8u * b; };
  const uint32_t x2_full = bar_base + 160u, norm_full = bar_base + 168u, w_full = bar_base + 176u;
  const uint32_t tmem_slot = bar_base + 192u;
  const uint32_t bias_s = bar_base + BAR_BYTES, beta_s = bias_s + 512u;   // per-tile channel constants (fp32 x 128 each)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_map(&map_a_hi); prefetch_map(&map_a_lo); prefetch_map(&map_w_hi); prefetch_map(&map_w_lo);
    if (p.gdn) { prefetch_map(&map_g_hi); prefetch_map(&map_g_lo); }
    if (p.tma_store) { prefetch_map(&map_y0); prefetch_map(&map_y1); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), p.planar ? 128 : 32 * EPI_WARPS); }
    mbar_init(x2_full, 32 * EPI_WARPS); mbar_init(norm_full, 1); mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() { if (++stage == p.stages) { stage = 0; phase ^= 1u; } };
      const int txy = p.tiles_x * p.tiles_y;
      if (p.w_resident) {
        mbar_expect_tx(w_full, (uint32_t)p.n_wtiles * 2u * p.b_bytes);
        for (int i = 0; i < p.n_wtiles; ++i) {
          const int t = i / p.kchunks, kc = i - t * p.kchunks;
          tma_load_3d(&map_w_hi, w_base + (uint32_t)(2 * i) * p.b_bytes, w_full, kc * BK, 0, p.taps[t].w);
          tma_load_3d(&map_w_lo, w_base + (uint32_t)(2 * i + 1) * p.b_bytes, w_full, kc * BK, 0, p.taps[t].w);
        }
      }
      walk_schedule(
          p,
          [&](int, const TaskCoord &tk, int t, int kc, bool, bool) {
            mbar_wait(empty_bar(stage), phase ^ 1u, 1);
            const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
            const uint32_t fb = full_bar(stage);
            mbar_expect_tx(fb, 2u * A_TILE_BYTES + (p.w_resident ? 0u : 2u * p.b_bytes));
            const int tb = tk.mt / txy, rr = tk.mt - tb * txy;
            const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
            const Tap tap = p.taps[t];
            const int c = kc * BK + tap.px * p.in_Cs;
            const int x = tx * p.bw + tap.dx, y = ty * p.bh + tap.dy, b = tb * p.bb;
            tma_load_5d(&map_a_hi, sa, fb, c, x, tap.py, y, b);
            tma_load_5d(&map_a_lo, sa + A_TILE_BYTES, fb, c, x, tap.py, y, b);
            if (!p.w_resident) {
              tma_load_3d(&map_w_hi, sa + 2 * A_TILE_BYTES, fb, kc * BK, tk.nt * p.BN, tap.w);
              tma_load_3d(&map_w_lo, sa + 2 * A_TILE_BYTES + p.b_bytes, fb, kc * BK, tk.nt * p.BN, tap.w);
            }
            advance();
          },
          [&](int, int c) {
            mbar_wait(empty_bar(stage), phase ^ 1u, 9);
            const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
            const uint32_t fb = full_bar(stage);
            mbar_expect_tx(fb, 2u * 128u * 128u);
            tma_load_2d(&map_g_hi, sa + b_off, fb, c * BK, 0);
            tma_load_2d(&map_g_lo, sa + b_off + p.b_bytes, fb, c * BK, 0);
            advance();
          });
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() { if (++stage == p.stages) { stage = 0; phase ^= 1u; } };
      const uint32_t idesc_g = instr_desc(128);
      if (p.w_resident) {
        mbar_wait(w_full, 0, 5);
        tc_fence_after();
      }
      walk_schedule(
          p,
          [&](int lt, const TaskCoord &tk, int t, int kc, bool first, bool last) {
            const int buf = lt & 1;
            if (first) {
              mbar_wait(acc_empty(buf), (((uint32_t)lt >> 1) & 1u) ^ 1u, 2);
              tc_fence_after();
            }
            mbar_wait(full_bar(stage), phase, 3);
            tc_fence_after();
            const int nvalid = min(p.BN, p.CoutPad16 - tk.nt * p.BN);
            const uint32_t idesc = instr_desc(nvalid);
            const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
            const uint64_t a_hi = smem_desc(sa), a_lo = smem_desc(sa + A_TILE_BYTES);
            const uint32_t sb = p.w_resident ? w_base + (uint32_t)(2 * (t * p.kchunks + kc)) * p.b_bytes : sa + 2 * A_TILE_BYTES;
            const uint64_t b_hi = smem_desc(sb), b_lo = smem_desc(sb + p.b_bytes);
            const uint32_t d_main = tmem_base + (uint32_t)buf * ACC_STRIDE, d_small = d_main + COL_SMALL;
            if (p.wide_n && nvalid == 128) {
              // Full 128-column tile: the (hi, lo) weight tiles are adjacent in shared memory and {main, small} are
              // adjacent in TMEM, so Ah.[Wh; Wl] is ONE N = 256 MMA -- Ah is read from shared memory once instead of
              // twice and the issue count drops from 3 to 2 per K = 16 step (the kernel is bound by the tensor
              // core's shared-memory operand reads: 24 -> 20 KB per step).
              const uint32_t idesc256 = instr_desc(256);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t o = (uint64_t)(k * 2);   // 32 B per K=16 step, encoded >> 4
                mma_ss(d_main, a_hi + o, b_hi + o, idesc256, (first && k == 0) ? 0u : 1u);
                mma_ss(d_small, a_lo + o, b_hi + o, idesc, 1u);
              }
            } else {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t o = (uint64_t)(k * 2);   // 32 B per K=16 step, encoded >> 4
                const uint32_t acc = (first && k == 0) ? 0u : 1u;
                mma_ss(d_main, a_hi + o, b_hi + o, idesc, acc);
                mma_ss(d_small, a_hi + o, b_lo + o, idesc, acc);
                mma_ss(d_small, a_lo + o, b_hi + o, idesc, 1u);
              }
            }
            tc_commit(empty_bar(stage));
            if (last) tc_commit(acc_full(buf));
            advance();
          },
          [&](int lt, int c) {
            const int buf = lt & 1;
            if (c == 0) {
              mbar_wait(x2_full, (uint32_t)lt & 1u, 4);
              tc_fence_after();
            }
            mbar_wait(full_bar(stage), phase, 6);
            tc_fence_after();
            const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
            const uint64_t b_hi = smem_desc(sa + b_off), b_lo = smem_desc(sa + b_off + p.b_bytes);
            // A operand (x^2 hi | lo, bf16) sits in this buffer's main columns; the norm overwrites small.
            const uint32_t x2 = tmem_base + (uint32_t)buf * ACC_STRIDE, d = x2 + COL_SMALL;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t o = (uint64_t)(k * 2);
              const uint32_t ah = x2 + (uint32_t)(c * 32 + k * 8);
              const uint32_t al = x2 + 64u + (uint32_t)(c * 32 + k * 8);
              mma_ts(d, ah, b_hi + o, idesc_g, (c == 0 && k == 0) ? 0u : 1u);
              mma_ts(d, ah, b_lo + o, idesc_g, 1u);
              mma_ts(d, al, b_hi + o, idesc_g, 1u);
            }
            tc_commit(empty_bar(stage));
            if (c == 1) tc_commit(norm_full);
            advance();
          });
    }
  } else if (EPI_WARPS == 16) {
    // ===================== epilogue, four groups (fused GDN, SPLIT output through TMA stores) =====================
    // group g = (warp - 2) / 4 owns the 32-channel chunk g; chunks (0, 1) and (2, 3) each fill one staging tile pair
    // (64 channels: hi tile + lo tile) and have their own store barrier and issuing thread.
    const int quad = warp & 3, grp = (warp - 2) >> 2, half = grp >> 1;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int txy = p.tiles_x * p.tiles_y;
    const bool is_issuer = warp == 2 + 8 * half && elect_one();   // one lane of the half's first warp
    const uint32_t stg_set = stg_base + (uint32_t)half * (uint32_t)STAGING_BYTES;
    const uint32_t row_off = stg_set + (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    const bool igdn = p.gdn == 2;
    int lt = 0;
    for (int task = blockIdx.x; task < p.n_tasks; task += gridDim.x, ++lt) {
      const TaskCoord tk = decode_task(p, task);
      const int buf = lt & 1;
      const int tb = tk.mt / txy, rr = tk.mt - tb * txy;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      const int ry = tk.ph / p.os, rx = tk.ph - ry * p.os;
      const int n0 = tk.nt * p.BN;
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)buf * ACC_STRIDE;
      const int c_fold = rx * p.out_Cs, mx = tx * p.bw, my = ty * p.bh, mb = tb * p.bb;

      epi_bar_all<16>();
      const int ci = (int)threadIdx.x - 64;
      if (ci < 128) {
        st_shared_f32(bias_s + 4u * ci, (n0 + ci < p.Cout) ? __ldg(p.bias + n0 + ci) : 0.f);
        st_shared_f32(beta_s + 4u * ci, __ldg(p.beta + ci));
      }
      epi_bar_all<16>();

      mbar_wait(acc_full(buf), ((uint32_t)lt >> 1) & 1u, 7);
      tc_fence_after();
      // pass 1: x = main + small + bias, 16 columns at a time
      float xs[32];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[16], q[16];
        tmem_ld16(acc + grp * 32 + hh * 16, r);
        tmem_ld16(acc + COL_SMALL + grp * 32 + hh * 16, q);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = ld_shared_f4(bias_s + 128u * grp + 64u * hh + 16u * j);
          xs[hh * 16 + 4 * j + 0] = (__uint_as_float(r[4 * j + 0]) + __uint_as_float(q[4 * j + 0])) + t.x;
          xs[hh * 16 + 4 * j + 1] = (__uint_as_float(r[4 * j + 1]) + __uint_as_float(q[4 * j + 1])) + t.y;
          xs[hh * 16 + 4 * j + 2] = (__uint_as_float(r[4 * j + 2]) + __uint_as_float(q[4 * j + 2])) + t.z;
          xs[hh * 16 + 4 * j + 3] = (__uint_as_float(r[4 * j + 3]) + __uint_as_float(q[4 * j + 3])) + t.w;
        }
      }
      // every group has its chunk in registers before the main columns are overwritten by the x^2 operand
      tc_fence_before();
      epi_bar_all<16>();
      tc_fence_after();
      {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = xs[2 * j], c = xs[2 * j + 1];
          split_pair(a * a, c * c, hi[j], lo[j]);
        }
        tmem_st16(acc + grp * 16, hi);
        tmem_st16(acc + 64u + grp * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(x2_full);
      // pass 2: y = x * rsqrt(beta + norm)   (IGDN: * sqrt)
      mbar_wait(norm_full, (uint32_t)lt & 1u, 8);
      tc_fence_after();
      {
        uint32_t q[32];
        tmem_ld32(acc + COL_SMALL + grp * 32, q);
        tmem_ld_wait();
#define HESIC_GDN_SCALE16(FN)                                                         \
  _Pragma("unroll") for (int jj = 0; jj < 8; ++jj) {                                  \
    const float4 bt = ld_shared_f4(beta_s + 128u * grp + 16u * jj);                   \
    xs[4 * jj + 0] *= FN(__uint_as_float(q[4 * jj + 0]) + bt.x);                      \
    xs[4 * jj + 1] *= FN(__uint_as_float(q[4 * jj + 1]) + bt.y);                      \
    xs[4 * jj + 2] *= FN(__uint_as_float(q[4 * jj + 2]) + bt.z);                      \
    xs[4 * jj + 3] *= FN(__uint_as_float(q[4 * jj + 3]) + bt.w);                      \
  }
        if (igdn) {
          HESIC_GDN_SCALE16(sqrt_approx)
        } else {
          HESIC_GDN_SCALE16(rsqrt_approx)
        }
#undef HESIC_GDN_SCALE16
      }
      tc_fence_before();
      mbar_arrive(acc_empty(buf));
      // store: this half's 64 channels as one (hi, lo) staging tile pair; the group fills bytes [64 (grp & 1), +64) of a row
      if (is_issuer) bulk_wait_read<0>();      // the half's previous tile pair has left shared memory
      epi_bar_half(half);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        split_pair(xs[g * 8 + 0], xs[g * 8 + 1], h0, l0);
        split_pair(xs[g * 8 + 2], xs[g * 8 + 3], h1, l1);
        split_pair(xs[g * 8 + 4], xs[g * 8 + 5], h2, l2);
        split_pair(xs[g * 8 + 6], xs[g * 8 + 7], h3, l3);
        const uint32_t o = row_off + (((uint32_t)((grp & 1) * 4 + g) ^ sw) << 4);
        st_shared_v4(o, h0, h1, h2, h3);
        st_shared_v4(o + A_TILE_BYTES, l0, l1, l2, l3);
      }
      fence_async_smem();
      epi_bar_half(half);
      if (is_issuer) {
        const int c0 = c_fold + n0 + half * 64;
        tma_store_5d(&map_y0, stg_set, c0, mx, ry, my, mb);
        tma_store_5d(&map_y1, stg_set + A_TILE_BYTES, c0, mx, ry, my, mb);
        bulk_commit();
      }
    }
    if (is_issuer) bulk_wait_all();
  } else if (!(p.planar && warp >= 6)) {
    // ===================== epilogue =====================
    // Two groups of four warps; warp w reads TMEM lane quadrant w % 4 (= 32 pixels of the tile) and group
    // g = (w - 2) / 4 takes the 32-channel chunks g, g + 2 of the tile's columns, so a thread holds at most
    // 64 channels of one pixel.  (One group doing all 128 channels ran 2900 straight-line instructions per
    // tile alone on its scheduler and stalled on instruction fetch; two groups halve the code per warp and
    // give every scheduler two warps.)  The planar path (Cout <= 4) needs one group only.
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int xi = row & (p.bw - 1), yi = (row >> p.lbw) & (p.bh - 1), bi = row >> (p.lbw + p.lbh);
    const int txy = p.tiles_x * p.tiles_y;
    const int nchunks = p.BN / 32, npairs = (nchunks + 1) / 2;
    const bool is_issuer = warp == 2 && elect_one();
    const uint32_t row_off0 = stg_base + (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    uint32_t pair_ctr = 0;   // store pairs issued so far (selects the staging set)
    // activation as max(v, slope * v): slope 1 = none, 0 = ReLU, 0.01 = LeakyReLU
    const float act_slope = p.act == HESIC_ACT_RELU ? 0.f : (p.act == HESIC_ACT_LEAKY_RELU ? 0.01f : 1.f);
    int lt = 0;
    for (int task = blockIdx.x; task < p.n_tasks; task += gridDim.x, ++lt) {
      const TaskCoord tk = decode_task(p, task);
      const int buf = lt & 1;
      const int tb = tk.mt / txy, rr = tk.mt - tb * txy;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      const int ry = tk.ph / p.os, rx = tk.ph - ry * p.os;
      const int ox = (tx * p.bw + xi) * p.os + rx, oy = (ty * p.bh + yi) * p.os + ry, b = tb * p.bb + bi;
      const bool valid = ox < p.Wout && oy < p.Hout && b < p.B;
      const size_t pix = ((size_t)b * p.Hout + oy) * p.Wout + ox;
      const int n0 = tk.nt * p.BN;
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)buf * ACC_STRIDE;
      const int c_fold = rx * p.out_Cs, mx = tx * p.bw, my = ty * p.bh, mb = tb * p.bb;

      // Output of chunk pair i (chunks 2i and 2i+1, one per group; `active` = this thread's chunk exists).
      // tma_store: both groups write their halves into 128B-swizzled [128 px][128 B] staging tiles
      // (conflict-free 16-byte stores), one thread hands the tiles to TMA, which writes whole pixel rows and
      // clips everything outside the tensor; else each thread stores its pixel directly.
      auto emit = [&](const float (&v)[32], int i, bool active) {
        const int nb = n0 + (2 * i + grp) * 32;
        if (!p.tma_store) {
          if (valid && active) store_chunk(p, v, pix, nb);
          return;
        }
        // staging set of this pair; with two sets only the pair before the previous one must have left shared memory
        const uint32_t stg_set = stg_base + (p.stg_sets == 2 ? (pair_ctr & 1u) * (uint32_t)STAGING_BYTES : 0u);
        const uint32_t row_off = row_off0 + (stg_set - stg_base);
        ++pair_ctr;
        if (is_issuer) {
          if (p.stg_sets == 2) bulk_wait_read<1>();
          else bulk_wait_read<0>();
        }
        epi_bar();
        if (p.out_fmt == HESIC_FMT_NHWC_F32) {
          // one [128 px][32 ch fp32] tile per group
          if (active) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_shared_v4(row_off + (uint32_t)grp * A_TILE_BYTES + (((uint32_t)j ^ sw) << 4), __float_as_uint(v[4 * j]),
                           __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          }
          fence_async_smem();
          epi_bar();
          if (is_issuer) {
            const int nb0 = n0 + 2 * i * 32;
            tma_store_5d(&map_y0, stg_set, c_fold + nb0, mx, ry, my, mb);
            if (2 * i + 1 < nchunks && nb0 + 32 < p.Cout)
              tma_store_5d(&map_y0, stg_set + A_TILE_BYTES, c_fold + nb0 + 32, mx, ry, my, mb);
            bulk_commit();
          }
        } else {
          // the pair fills one 64-channel (hi, lo) tile pair: group g owns bytes [64g, 64g + 64) of every row
          if (active) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
              split_pair(v[g * 8 + 0], v[g * 8 + 1], h0, l0);
              split_pair(v[g * 8 + 2], v[g * 8 + 3], h1, l1);
              split_pair(v[g * 8 + 4], v[g * 8 + 5], h2, l2);
              split_pair(v[g * 8 + 6], v[g * 8 + 7], h3, l3);
              const uint32_t o = row_off + (((uint32_t)(grp * 4 + g) ^ sw) << 4);
              st_shared_v4(o, h0, h1, h2, h3);
              st_shared_v4(o + A_TILE_BYTES, l0, l1, l2, l3);
            }
          }
          fence_async_smem();
          epi_bar();
          if (is_issuer) {
            const int c0 = c_fold + n0 + 2 * i * 32;
            tma_store_5d(&map_y0, stg_set, c0, mx, ry, my, mb);
            tma_store_5d(&map_y1, stg_set + A_TILE_BYTES, c0, mx, ry, my, mb);
            bulk_commit();
          }
        }
      };

      if (!p.planar) {
        // channel constants of this n-tile -> smem (overlaps the tile's MMAs)
        epi_bar();
        const int ci = (int)threadIdx.x - 64;
        if (ci < 128) {
          st_shared_f32(bias_s + 4u * ci, (n0 + ci < p.Cout) ? __ldg(p.bias + n0 + ci) : 0.f);
          if (p.gdn) st_shared_f32(beta_s + 4u * ci, __ldg(p.beta + ci));
        }
        epi_bar();
      }

      mbar_wait(acc_full(buf), ((uint32_t)lt >> 1) & 1u, 7);
      tc_fence_after();
      if (p.planar) {
        uint32_t r[16], q[16];
        tmem_ld16(acc, r);
        tmem_ld16(acc + COL_SMALL, q);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(acc_empty(buf));   // the tile is in registers: release the accumulator early
        const int qx = tx * p.bw + xi, qy = ty * p.bh + yi;
        if (b < p.B) {
          float bia[4], bet[4], gam[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            bia[c] = c < p.Cout ? __ldg(p.bias + c) : 0.f;
            bet[c] = (p.pl_gdn && c < p.Cout) ? __ldg(p.beta + c) : 1.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) gam[j * 4 + c] = (p.pl_gdn && c < p.Cout && j < p.Cout) ? __ldg(p.pl_gamma + j * p.Cout + c) : 0.f;
          }
#pragma unroll
          for (int ph = 0; ph < 4; ++ph) {
            if (ph < p.pl_phases) {
              const int oy2 = qy * p.pl_os + (ph >> 1), ox2 = qx * p.pl_os + (ph & 1);
              float x[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) x[c] = (__uint_as_float(r[ph * 4 + c]) + __uint_as_float(q[ph * 4 + c])) + bia[c];
              if (p.pl_gdn) {
                float sq[4], o[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) sq[c] = x[c] * x[c];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  float nrm = bet[c];
#pragma unroll
                  for (int j = 0; j < 4; ++j) nrm = fmaf(gam[j * 4 + c], sq[j], nrm);
                  o[c] = x[c] * (p.pl_gdn == 2 ? sqrtf(nrm) : rsqrtf(nrm));
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) x[c] = o[c];
              }
              if (oy2 < p.Hout && ox2 < p.Wout) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  if (c < p.Cout)
                    ((float *)p.y0)[(((size_t)b * p.out_Cs + c) * p.Hout + oy2) * p.Wout + ox2] = apply_act(x[c], p.act);
              }
            }
          }
        }
      } else if (!p.gdn) {
#pragma unroll 1
        for (int i = 0; i < npairs; ++i) {
          if (n0 + 2 * i * 32 >= p.Cout) break;
          const int ch = 2 * i + grp;
          const bool active = ch < nchunks && n0 + ch * 32 < p.Cout;
          float v[32];
          if (active) {
            uint32_t r[32], q[32];
            tmem_ld32(acc + ch * 32, r);
            tmem_ld32(acc + COL_SMALL + ch * 32, q);
            tmem_ld_wait();
            ld_chan32(bias_s + 128u * ch, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float t = (__uint_as_float(r[j]) + __uint_as_float(q[j])) + v[j];
              v[j] = fmaxf(t, t * act_slope);
            }
            if (p.stats) gn_accumulate(p, v, valid, b, n0 + ch * 32);
          }
          emit(v, i, active);
        }
        tc_fence_before();
        mbar_arrive(acc_empty(buf));
      } else {
        // pass 1: x = conv + bias (kept in registers); x^2 -> bf16 (hi | lo) A operand over the main columns
        float xs[64];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ch = 2 * i + grp;
          uint32_t r[32], q[32];
          tmem_ld32(acc + ch * 32, r);
          tmem_ld32(acc + COL_SMALL + ch * 32, q);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = ld_shared_f4(bias_s + 128u * ch + 16u * j);
            xs[i * 32 + 4 * j + 0] = (__uint_as_float(r[4 * j + 0]) + __uint_as_float(q[4 * j + 0])) + t.x;
            xs[i * 32 + 4 * j + 1] = (__uint_as_float(r[4 * j + 1]) + __uint_as_float(q[4 * j + 1])) + t.y;
            xs[i * 32 + 4 * j + 2] = (__uint_as_float(r[4 * j + 2]) + __uint_as_float(q[4 * j + 2])) + t.z;
            xs[i * 32 + 4 * j + 3] = (__uint_as_float(r[4 * j + 3]) + __uint_as_float(q[4 * j + 3])) + t.w;
          }
        }
        // every thread of both groups has its chunks in registers before main is overwritten
        tc_fence_before();
        epi_bar();
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ch = 2 * i + grp;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = xs[i * 32 + 2 * j], c = xs[i * 32 + 2 * j + 1];
            split_pair(a * a, c * c, hi[j], lo[j]);
          }
          tmem_st16(acc + ch * 16, hi);
          tmem_st16(acc + 64u + ch * 16, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(x2_full);
        // pass 2: y = x * rsqrt(beta + norm)   (IGDN: * sqrt), formed in place so that the accumulator buffer
        // can be handed back to the MMA warp before the (longer) store phase
        mbar_wait(norm_full, (uint32_t)lt & 1u, 8);
        tc_fence_after();
        const bool igdn = p.gdn == 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ch = 2 * i + grp;
          uint32_t q[32];
          tmem_ld32(acc + COL_SMALL + ch * 32, q);
          tmem_ld_wait();
          // beta four at a time (a 32-float table would sit on top of xs[64] + q[32] and spill)
          // one uniform branch per 32-column chunk (not per element: `igdn ? sqrt : rsqrt` inside the unrolled loop
          // became a branch around every group of four, 16 basic blocks whose LDS -> FADD -> MUFU -> FMUL latencies
          // could not overlap); the chunk boundary also keeps the next chunk's TMEM load from being hoisted over
          // this one's arithmetic, which would spill xs[]
#define HESIC_GDN_SCALE_CHUNK(FN)                                                     \
  _Pragma("unroll") for (int jj = 0; jj < 8; ++jj) {                                  \
    const float4 bt = ld_shared_f4(beta_s + 128u * ch + 16u * jj);                    \
    xs[i * 32 + 4 * jj + 0] *= FN(__uint_as_float(q[4 * jj + 0]) + bt.x);             \
    xs[i * 32 + 4 * jj + 1] *= FN(__uint_as_float(q[4 * jj + 1]) + bt.y);             \
    xs[i * 32 + 4 * jj + 2] *= FN(__uint_as_float(q[4 * jj + 2]) + bt.z);             \
    xs[i * 32 + 4 * jj + 3] *= FN(__uint_as_float(q[4 * jj + 3]) + bt.w);             \
  }
          if (igdn) {
            HESIC_GDN_SCALE_CHUNK(sqrt_approx)
          } else {
            HESIC_GDN_SCALE_CHUNK(rsqrt_approx)
          }
#undef HESIC_GDN_SCALE_CHUNK
        }
        tc_fence_before();
        mbar_arrive(acc_empty(buf));
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = xs[i * 32 + j];
          emit(v, i, true);
        }
      }
    }
  }

  if (EPI_WARPS == 8 && p.tma_store && warp == 2) bulk_wait_all();   // staged tiles fully written before the CTA retires
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace hesic
#include "conv_head.cuh"
#include "conv_tc_pair.cuh"
#include "conv_tc_first.cuh"
namespace hesic {
namespace tc {

// ---------------------------------------------------------------------------------------------
// host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

// rank-n tensor map (bf16 unless stated), 128B swizzle, zero OOB fill.  dims/box innermost first; strides (bytes) for dims 1..n-1.
static int make_map(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides,
                    const uint32_t *box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return HESIC_E_CUDA; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  CUresult r = fn(m, dtype, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,..] box=[%u,%u,%u,..]", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0);
    return HESIC_E_CUDA;
  }
  return HESIC_OK;
}

// exported to the other tcgen05 translation units (enhance.cu)
int make_tensor_map(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides,
                    const uint32_t *box, bool f32, bool swizzle128) {
  return make_map(m, base, rank, dims, strides, box, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
}

static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
static int posmod(int a, int m) { return ((a % m) + m) % m; }
static bool aligned16(const void *q) { return ((uintptr_t)q & 15u) == 0; }

static void add_tap(Params &p, int &n, int dy, int dx, int py, int px, int w) {
  Tap t;
  t.dy = (int8_t)dy; t.dx = (int8_t)dx; t.py = (int8_t)py; t.px = (int8_t)px; t.w = (int16_t)w; t.pad_ = 0;
  p.taps[n++] = t;
}

// tap tables of the three operand formulations (see conv.h: hesic_conv::tc_kind)
static int build_taps(const hesic_conv *c, Params &p) {
  int n = 0;
  p.n_phases = 1; p.os = 1;
  p.tap_begin[0] = 0;
  if (c->tc_kind == HESIC_TC_ROW2) {
    // one tap per pair of kernel rows (ky = 2s, 2s + 1): box row origin = output row + s in units of two rows
    for (int s2 = 0; s2 < 3; ++s2) add_tap(p, n, s2, 0, 0, 0, s2);
    p.tap_begin[1] = n;
  } else if (c->tc_kind == HESIC_TC_ROW) {
    // one tap per kernel row; the 5 x-taps x 8 channel slots are the K dimension of the box row
    for (int ky = 0; ky < 5; ++ky) {
      if (!c->transposed && c->stride == 2) add_tap(p, n, ky >> 1, 0, ky & 1, 0, ky);
      else add_tap(p, n, ky, 0, 0, 0, ky);
    }
    p.tap_begin[1] = n;
  } else if (!c->transposed) {
    for (int ky = 0; ky < c->kh; ++ky)
      for (int kx = 0; kx < c->kw; ++kx) {
        if (c->kh * c->kw <= 64 && !((c->live_taps >> (ky * c->kw + kx)) & 1ull)) continue;   // masked-out tap (MaskedConv2d)
        const int oy = ky - c->pad, ox = kx - c->pad;
        if (c->stride == 1) add_tap(p, n, oy, ox, 0, 0, ky * c->kw + kx);
        else {
          const int py = posmod(oy, 2), px = posmod(ox, 2);
          add_tap(p, n, (oy - py) / 2, (ox - px) / 2, py, px, ky * c->kw + kx);
        }
      }
    p.tap_begin[1] = n;
  } else {
    const int s = c->stride;
    p.n_phases = s * s; p.os = s;
    for (int ph = 0; ph < s * s; ++ph) {
      const int ry = ph / s, rx = ph % s;
      p.tap_begin[ph] = n;
      for (int ky = 0; ky < c->kh; ++ky) {
        if (posmod(ry + c->pad - ky, s) != 0) continue;
        for (int kx = 0; kx < c->kw; ++kx) {
          if (posmod(rx + c->pad - kx, s) != 0) continue;
          add_tap(p, n, (ry + c->pad - ky) / s, (rx + c->pad - kx) / s, 0, 0, ky * c->kw + kx);
        }
      }
      if (n == p.tap_begin[ph]) return -1;   // a phase without taps
    }
    p.tap_begin[s * s] = n;
  }
  return n;
}

}  // namespace tc

bool conv_tc_supported(const hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y) {
  using namespace tc;
  const int xCs = x->Cs > 0 ? x->Cs : x->C, yCs = y->Cs > 0 ? y->Cs : y->C;
  if ((int64_t)x->B * x->H * x->W == 0 || (int64_t)y->B * y->H * y->W == 0) return false;
  if (!aligned16(x->p0) || !aligned16(x->p1)) return false;
  const bool planar_out = c->Cout <= 4;
  if (planar_out) {
    if (y->fmt != HESIC_FMT_NCHW_F32) return false;
    if (c->has_gdn && c->Cout < 2) return false;
  } else {
    if (y->fmt != HESIC_FMT_NHWC_SPLIT && y->fmt != HESIC_FMT_NHWC_F32) return false;
    if (c->Cout % 8 || c->Cout < 16 || yCs % 8 || !aligned16(y->p0)) return false;
    if (y->fmt == HESIC_FMT_NHWC_SPLIT && !aligned16(y->p1)) return false;
    if (c->has_gdn && c->Cout != 128) return false;
  }
  if (c->tc_kind == HESIC_TC_ROW2) {
    if (x->fmt != HESIC_FMT_ROWPAD8_SPLIT || x->Cs != 4 || ((x->H | x->W) & 1)) return false;
  } else if (c->tc_kind == HESIC_TC_ROW) {
    if (x->fmt != HESIC_FMT_ROWPAD8_SPLIT || x->Cs != 8) return false;
    if (!c->transposed && c->stride == 2 && ((x->H | x->W) & 1)) return false;
  } else if (c->tc_kind == HESIC_TC_SCATTER) {
    if (x->fmt != HESIC_FMT_NHWC_SPLIT || !planar_out) return false;
    if (c->Cin % 64 || c->Cin > 256 || xCs % 8) return false;
  } else {
    if (x->fmt != HESIC_FMT_NHWC_SPLIT || planar_out) return false;
    if (c->Cin % 8 || c->Cin < 16 || xCs % 8) return false;
    if (c->kh * c->kw > MAX_TAPS || c->kh != c->kw) return false;
    if (c->stride != 1 && c->stride != 2) return false;
    if (!c->transposed && c->stride == 2 && ((x->H | x->W) & 1)) return false;
    if (c->transposed && (c->kh < c->stride)) return false;
  }
  return encode_fn() != nullptr;
}

// RGB synthesis head (conv_head.cuh)
template <int COUT>
static int launch_head(hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y, int act, cudaStream_t s) {
  using namespace tc;
  static int num_sms = 0;
  static PerDeviceOnce once;
  if (once.first()) {
    int dev = 0;
    HESIC_CUDA(cudaGetDevice(&dev));
    HESIC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    HESIC_CUDA(cudaFuncSetAttribute(head::conv_head_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  }
  const int xCs = x->Cs > 0 ? x->Cs : x->C;
  head::HParams p;
  memset(&p, 0, sizeof(p));
  p.H = x->H; p.W = x->W; p.B = x->B;
  p.tiles_x = (x->W + head::IN_W - 1) / head::IN_W; p.tiles_y = (x->H + head::IN_H - 1) / head::IN_H;
  p.n_tasks = p.tiles_x * p.tiles_y * x->B;
  p.NPAD = (25 * COUT + 15) / 16 * 16;
  p.kchunks = c->Cin / BK;
  p.y = (float *)y->p0; p.out_Cs = y->Cs > 0 ? y->Cs : y->C;
  p.bias = c->bias; p.act = act;
  p.gdn = c->has_gdn ? (c->gdn_inverse ? 2 : 1) : 0;
  p.beta = c->gdn_beta; p.gamma = c->gdn_w_simt;
  // hesic_conv_forward_sse: NOT fused here (c->sse_fused stays false, the caller runs the squared-error kernel on the
  // written image).  Measured r04, 16 x 3 x 512 x 512: this kernel 124.5 us + sse_dense_kernel 27 us, against 172-183 us
  // with the target read and the squares in this epilogue (whether the target values were fetched where they are used,
  // at the top of the tile, or one tile ahead): the epilogue is what bounds this kernel, and its 14-pixel-wide tile rows
  // make the extra loads as expensive as its stores.
  const int fixed = 1024 + 256 + p.kchunks * 2 * p.NPAD * 128 + BM * (p.NPAD + 1) * 4;
  p.stages = std::min(8, (SMEM_LIMIT - fixed) / head::STAGE_BYTES);
  if (p.stages < 2) { set_error("conv head: operands do not fit shared memory"); return HESIC_E_UNSUPPORTED; }
  const int smem_bytes = fixed + p.stages * head::STAGE_BYTES;
  CUtensorMap ma_hi, ma_lo;
  {
    const uint64_t e = 2;
    uint64_t dims[5] = {(uint64_t)x->C, (uint64_t)x->W, 1, (uint64_t)x->H, (uint64_t)x->B};
    uint64_t strides[4] = {(uint64_t)xCs * e, (uint64_t)x->W * xCs * e, (uint64_t)x->W * xCs * e,
                           (uint64_t)x->H * x->W * xCs * e};
    uint32_t box[5] = {(uint32_t)BK, (uint32_t)head::TILE_W, 1u, (uint32_t)head::TILE_H, 1u};
    int r = make_map(&ma_hi, x->p0, 5, dims, strides, box);
    if (r == HESIC_OK) r = make_map(&ma_lo, x->p1, 5, dims, strides, box);
    if (r != HESIC_OK) return r;
  }
  if (!c->tc_maps || c->tc_maps_bn != p.NPAD) {
    if (!c->tc_maps) c->tc_maps = (unsigned char *)aligned_alloc(128, 4 * sizeof(CUtensorMap));
    CUtensorMap *m = (CUtensorMap *)c->tc_maps;
    uint64_t dims[2] = {(uint64_t)c->Cin, (uint64_t)p.NPAD}, strides[1] = {(uint64_t)c->Cin * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)p.NPAD};
    int r = make_map(&m[0], c->w_hi, 2, dims, strides, box);
    if (r == HESIC_OK) r = make_map(&m[1], c->w_lo, 2, dims, strides, box);
    if (r != HESIC_OK) return r;
    c->tc_maps_bn = p.NPAD;
  }
  const CUtensorMap *m = (const CUtensorMap *)c->tc_maps;
  const int grid = std::min(p.n_tasks, num_sms);
  head::conv_head_kernel<COUT><<<grid, head::NT, smem_bytes, s>>>(ma_hi, ma_lo, m[0], m[1], p);
  HESIC_LAUNCHED("conv_head_kernel");
  return HESIC_OK;
}

int conv_forward_tc(hesic_conv *c, const hesic_tensor *x, const hesic_tensor *y, int act, cudaStream_t s) {
  using namespace tc;
  if (c->has_gdn && act != HESIC_ACT_NONE) {
    set_error("activation after fused GDN is not supported");
    return HESIC_E_UNSUPPORTED;
  }
  if (c->tc_kind == HESIC_TC_SCATTER) {
    switch (c->Cout) {
      case 1: return launch_head<1>(c, x, y, act, s);
      case 2: return launch_head<2>(c, x, y, act, s);
      case 3: return launch_head<3>(c, x, y, act, s);
      default: return launch_head<4>(c, x, y, act, s);
    }
  }
  static int num_sms = 0;
  static PerDeviceOnce once;
  if (once.first()) {
    int dev = 0;
    HESIC_CUDA(cudaGetDevice(&dev));
    HESIC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    HESIC_CUDA(cudaFuncSetAttribute(conv_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    HESIC_CUDA(cudaFuncSetAttribute(conv_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  }
  const int xCs = x->Cs > 0 ? x->Cs : x->C, yCs = y->Cs > 0 ? y->Cs : y->C;
  const bool planar = c->Cout <= 4;
  Params p;
  memset(&p, 0, sizeof(p));
  const int ntaps = build_taps(c, p);
  if (ntaps <= 0) { set_error("conv tcgen05: unsupported tap structure"); return HESIC_E_UNSUPPORTED; }
  p.planar = planar ? 1 : 0;
  p.pl_phases = 1;
  p.pl_os = 1;
  const int tile_os = planar ? p.pl_os : p.os;     // output pixels per tile-space pixel, per axis
  const int Hq = (y->H + tile_os - 1) / tile_os, Wq = (y->W + tile_os - 1) / tile_os;
  p.bw = Wq >= 16 ? 16 : pow2ceil(Wq);
  p.bh = std::min(BM / p.bw, pow2ceil(Hq));
  p.bb = BM / (p.bw * p.bh);
  p.lbw = ilog2(p.bw); p.lbh = ilog2(p.bh);
  p.tiles_x = (Wq + p.bw - 1) / p.bw; p.tiles_y = (Hq + p.bh - 1) / p.bh; p.tiles_b = (y->B + p.bb - 1) / p.bb;
  // N tile: at most 128 columns (two double-buffered {main, small} accumulators fill the 512 TMEM columns);
  // the last tile of a layer may be narrower (its MMAs are issued with the remaining N).
  p.BN = planar ? 16 : std::min(128, (c->Cout + 31) / 32 * 32);
  p.Cout = c->Cout;
  p.CoutPad16 = planar ? 16 : (c->Cout + 15) / 16 * 16;
  p.n_tiles = planar ? 1 : (c->Cout + p.BN - 1) / p.BN;
  const bool rowk = c->tc_kind == HESIC_TC_ROW || c->tc_kind == HESIC_TC_ROW2;
  p.kchunks = rowk ? 1 : (c->Cin + BK - 1) / BK;
  for (int i = 0; i < 8; ++i) { p.kc_lo[i] = 0; p.kc_hi[i] = (int16_t)p.kchunks; }
  if (p.kchunks > 32767) { set_error("conv tcgen05: too many input channels"); return HESIC_E_UNSUPPORTED; }
  p.in_Cs = xCs;
  p.out_fmt = y->fmt; p.out_Cs = yCs; p.y0 = y->p0; p.y1 = y->p1;
  p.Hout = y->H; p.Wout = y->W; p.B = y->B;
  p.bias = c->bias; p.act = act;
  p.gdn = (!planar && c->has_gdn) ? (c->gdn_inverse ? 2 : 1) : 0;
  p.pl_gdn = (planar && c->has_gdn) ? (c->gdn_inverse ? 2 : 1) : 0;
  p.beta = c->gdn_beta;
  p.pl_gamma = c->gdn_w_simt;
  p.b_bytes = (uint32_t)p.BN * 128u;
  c->gn_fused = false;
  if (c->gn_stats && c->gn_groups >= 1 && !planar && !p.gdn && y->fmt == HESIC_FMT_NHWC_F32 && p.bb == 1 && c->Cout % 32 == 0 &&
      c->Cout % c->gn_groups == 0 && (c->gn_groups == 1 || (c->Cout / c->gn_groups) % 32 == 0) && c->tc_kind == HESIC_TC_GENERIC) {
    p.stats = c->gn_stats; p.stat_groups = c->gn_groups; p.stat_cpg = c->Cout / c->gn_groups;
    c->gn_fused = true;
  }
  // ROW2: the whole weight set (3 tiles, hi + lo) fits next to the pipeline -> loaded once per CTA; a stage then
  // carries only the A tiles (or, for a GDN step, the gamma tiles in the same space)
  p.n_wtiles = ntaps * p.kchunks;
  if (c->kband_bn == p.BN && c->kband_chunks == p.kchunks && p.n_tiles <= 8 && !rowk) {
    // block-banded weights (hesic_conv_detect_kband): skip the K chunks of an N tile that hold only zeros
    for (int i = 0; i < p.n_tiles; ++i) { p.kc_lo[i] = c->kc_lo[i]; p.kc_hi[i] = c->kc_hi[i]; }
  }
  p.w_resident = (c->tc_kind == HESIC_TC_ROW2 && p.n_tiles == 1 && p.b_bytes <= (uint32_t)A_TILE_BYTES) ? 1 : 0;
  // diagnostic switches, read once per process
  static const bool env_no_resident = diag_env("HESIC_TC_NO_RESIDENT") != nullptr, env_narrow = diag_env("HESIC_TC_NARROW") != nullptr,
                    env_direct_store = diag_env("HESIC_TC_DIRECT_STORE") != nullptr,
                    env_one_staging = diag_env("HESIC_TC_ONE_STAGING") != nullptr;
  if (env_no_resident) p.w_resident = 0;
  p.wide_n = (!planar && p.BN == 128 && !env_narrow) ? 1 : 0;
  static const int gdn_at_env = diag_env("HESIC_TC_GDN_AT") ? atoi(diag_env("HESIC_TC_GDN_AT")) : GDN_AT;
  p.gdn_at = std::max(0, gdn_at_env);
  p.stage_bytes = 2u * A_TILE_BYTES + (p.w_resident ? 0u : 2u * p.b_bytes);
  // TMA-store epilogue: channels-last outputs; for the sub-pixel phases of a transposed conv the phase
  // column is folded into the channel dimension of the output map, which needs whole store tiles.
  const int store_ch = y->fmt == HESIC_FMT_NHWC_SPLIT ? 64 : 32;
  const int esz = y->fmt == HESIC_FMT_NHWC_SPLIT ? 2 : 4;
  p.tma_store = (!planar && (yCs * esz) % 16 == 0 &&
                 (p.os == 1 || (p.os == 2 && c->Cout % store_ch == 0 && !((y->H | y->W) & 1)))) ? 1 : 0;
  if (env_direct_store) p.tma_store = 0;
  int fixed = 1024 + BAR_BYTES + CHAN_BYTES + (p.tma_store ? STAGING_BYTES : 0) +
              (p.w_resident ? p.n_wtiles * 2 * (int)p.b_bytes : 0);
  p.stg_sets = 1;
  if (p.tma_store && !env_one_staging) {
    // a second staging set where the operand pipeline keeps its depth: short-K layers are epilogue-bound and every store
    // pair otherwise waits for the previous pair to leave shared memory (r01 profile of the first analysis layer: 11 %)
    const int st2 = (SMEM_LIMIT - fixed - STAGING_BYTES) / (int)p.stage_bytes;
    // (measured: giving up the third pipeline stage for it slows the K-heavy layers by 5-20 %)
    if (st2 >= 3 || (p.w_resident && st2 >= 2)) { p.stg_sets = 2; fixed += STAGING_BYTES; }
  }
  p.stages = std::min(8, (SMEM_LIMIT - fixed) / (int)p.stage_bytes);
  if (p.stages < 2) { set_error("conv tcgen05: tile does not fit shared memory"); return HESIC_E_UNSUPPORTED; }
  p.n_tasks = p.tiles_x * p.tiles_y * p.tiles_b * p.n_phases * p.n_tiles;
  const int smem_bytes = p.stages * (int)p.stage_bytes + fixed;

  CUtensorMap my0, my1;
  memset(&my0, 0, sizeof(my0));
  memset(&my1, 0, sizeof(my1));
  if (p.tma_store) {
    uint64_t dims[5], strides[4];
    uint32_t box[5] = {(uint32_t)store_ch, (uint32_t)p.bw, 1u, (uint32_t)p.bh, (uint32_t)p.bb};
    const uint64_t e = (uint64_t)esz;
    if (p.os == 2) {
      dims[0] = (uint64_t)yCs + y->C; dims[1] = y->W / 2; dims[2] = 2; dims[3] = y->H / 2; dims[4] = y->B;
      strides[0] = 2ull * yCs * e; strides[1] = (uint64_t)y->W * yCs * e; strides[2] = 2ull * y->W * yCs * e;
      strides[3] = (uint64_t)y->H * y->W * yCs * e;
    } else {
      dims[0] = y->C; dims[1] = y->W; dims[2] = 1; dims[3] = y->H; dims[4] = y->B;
      strides[0] = (uint64_t)yCs * e; strides[1] = (uint64_t)y->W * yCs * e; strides[2] = (uint64_t)y->W * yCs * e;
      strides[3] = (uint64_t)y->H * y->W * yCs * e;
    }
    const CUtensorMapDataType dt = esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    int r = make_map(&my0, y->p0, 5, dims, strides, box, dt);
    if (r == HESIC_OK && esz == 2) r = make_map(&my1, y->p1, 5, dims, strides, box, dt);
    if (r != HESIC_OK) return r;
  }

  // activation maps (depend on the input pointer -> encoded per call, host-only work)
  CUtensorMap ma_hi, ma_lo;
  {
    uint64_t dims[5], strides[4];
    uint32_t box[5] = {(uint32_t)BK, (uint32_t)p.bw, 1u, (uint32_t)p.bh, (uint32_t)p.bb};
    const uint64_t e = 2;
    if (c->tc_kind == HESIC_TC_ROW2) {
      // row-pair interleaved ROWPAD (4 slots): element (k, ox, pair) = padded pixel 2*ox + k/8, padded row
      // 2*pair + (k%8)/4, channel slot k%4 -- 64 contiguous elements = 8 pixels x 2 rows x 4 slots
      const uint64_t pairB = (uint64_t)(x->W + HESIC_ROWPAD_X) * 8 * e, Hp2 = ((uint64_t)x->H + HESIC_ROWPAD_Y) / 2;
      dims[0] = BK; dims[1] = x->W / 2; dims[2] = 1; dims[3] = Hp2; dims[4] = x->B;
      strides[0] = 16 * e; strides[1] = pairB; strides[2] = pairB; strides[3] = Hp2 * pairB;
    } else if (c->tc_kind == HESIC_TC_ROW) {
      // overlapping rows: element (k, ox) = padded pixel (stride*ox + k/8), channel slot k%8
      const uint64_t rowB = (uint64_t)(x->W + HESIC_ROWPAD_X) * 8 * e, Hp = (uint64_t)x->H + HESIC_ROWPAD_Y;
      if (!c->transposed && c->stride == 2) {
        dims[0] = BK; dims[1] = x->W / 2; dims[2] = 2; dims[3] = Hp / 2; dims[4] = x->B;
        strides[0] = 16 * e; strides[1] = rowB; strides[2] = 2 * rowB; strides[3] = Hp * rowB;
      } else {
        dims[0] = BK; dims[1] = x->W; dims[2] = 1; dims[3] = Hp; dims[4] = x->B;
        strides[0] = 8 * e; strides[1] = rowB; strides[2] = rowB; strides[3] = Hp * rowB;
      }
    } else if (c->tc_kind == HESIC_TC_GENERIC && !c->transposed && c->stride == 2) {
      dims[0] = (uint64_t)xCs + x->C; dims[1] = x->W / 2; dims[2] = 2; dims[3] = x->H / 2; dims[4] = x->B;
      strides[0] = 2ull * xCs * e; strides[1] = (uint64_t)x->W * xCs * e; strides[2] = 2ull * x->W * xCs * e;
      strides[3] = (uint64_t)x->H * x->W * xCs * e;
    } else {
      dims[0] = x->C; dims[1] = x->W; dims[2] = 1; dims[3] = x->H; dims[4] = x->B;
      strides[0] = (uint64_t)xCs * e; strides[1] = (uint64_t)x->W * xCs * e; strides[2] = (uint64_t)x->W * xCs * e;
      strides[3] = (uint64_t)x->H * x->W * xCs * e;
    }
    int r = make_map(&ma_hi, x->p0, 5, dims, strides, box);
    if (r == HESIC_OK) r = make_map(&ma_lo, x->p1, 5, dims, strides, box);
    if (r != HESIC_OK) return r;
  }
  // weight / gamma maps (static per layer; re-encoded when the tile shape changes)
  const int want_gdn = p.gdn ? 1 : 0;
  if (!c->tc_maps || c->tc_maps_bn != p.BN || c->tc_maps_gdn != want_gdn) {
    if (!c->tc_maps) c->tc_maps = (unsigned char *)aligned_alloc(128, 4 * sizeof(CUtensorMap));
    CUtensorMap *m = (CUtensorMap *)c->tc_maps;
    uint64_t dims[3] = {(uint64_t)c->tc_k, (uint64_t)c->CoutPad, (uint64_t)c->tc_taps};
    uint64_t strides[2] = {(uint64_t)c->tc_k * 2, (uint64_t)c->tc_k * c->CoutPad * 2};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)p.BN, 1u};
    int r = make_map(&m[0], c->w_hi, 3, dims, strides, box);
    if (r == HESIC_OK) r = make_map(&m[1], c->w_lo, 3, dims, strides, box);
    if (r == HESIC_OK && want_gdn) {
      uint64_t gd[2] = {128, 128}, gs[1] = {256};
      uint32_t gb[2] = {(uint32_t)BK, 128u};
      r = make_map(&m[2], c->gdn_g_hi, 2, gd, gs, gb);
      if (r == HESIC_OK) r = make_map(&m[3], c->gdn_g_lo, 2, gd, gs, gb);
    } else if (r == HESIC_OK) {
      m[2] = m[0]; m[3] = m[1];
    }
    if (r != HESIC_OK) return r;
    c->tc_maps_bn = p.BN; c->tc_maps_gdn = want_gdn;
  }
  const CUtensorMap *m = (const CUtensorMap *)c->tc_maps;
  // CTA-pair kernel (conv_tc_pair.cuh): plain K-heavy layers with full 128-column N tiles
  static const bool pair_on = diag_env("HESIC_TC_SINGLE_CTA") == nullptr;
  if (pair_on && c->tc_kind == HESIC_TC_GENERIC && !planar && p.tma_store && p.BN == 128 && !p.w_resident &&
      num_sms >= 2 && p.kchunks * ntaps >= 8) {
    static PerDeviceOnce pair_once;
    if (pair_once.first())
      HESIC_CUDA(cudaFuncSetAttribute(conv_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    CUtensorMap mw64;
    {
      uint64_t dims[3] = {(uint64_t)c->tc_k, (uint64_t)c->CoutPad, (uint64_t)c->tc_taps};
      uint64_t strides[2] = {(uint64_t)c->tc_k * 2, (uint64_t)c->tc_k * c->CoutPad * 2};
      uint32_t box[3] = {(uint32_t)BK, 64u, 1u};
      int r = make_map(&mw64, c->w_hi, 3, dims, strides, box);
      if (r != HESIC_OK) return r;
    }
    CUtensorMap mg_hi = mw64, mg_lo = mw64;
    if (p.gdn) {
      uint64_t gd[2] = {128, 128}, gs[1] = {256};
      uint32_t gb[2] = {(uint32_t)BK, 64u};
      int r = make_map(&mg_hi, c->gdn_g_hi, 2, gd, gs, gb);
      if (r == HESIC_OK) r = make_map(&mg_lo, c->gdn_g_lo, 2, gd, gs, gb);
      if (r != HESIC_OK) return r;
    }
    Params q = p;
    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
    q.n_tiles = (c->Cout + 127) / 128;
    q.n_tasks = ((m_tiles + 1) / 2) * p.n_phases * q.n_tiles;     // PAIR tasks
    const int pfixed = 1024 + BAR_BYTES + CHAN_BYTES + STAGING_BYTES;
    q.stages = std::min(8, (SMEM_LIMIT - pfixed) / PAIR_STAGE_BYTES);
    q.stg_sets = 1;
    const int pgrid = std::min(2 * q.n_tasks, num_sms & ~1);
    conv_tc_pair_kernel<<<pgrid, NUM_THREADS, q.stages * PAIR_STAGE_BYTES + pfixed, s>>>(ma_hi, ma_lo, m[0], m[1], mw64, mg_hi, mg_lo, my0, my1, q);
    HESIC_LAUNCHED("conv_tc_pair_kernel");
    return HESIC_OK;
  }
  // first analysis layer (conv_tc_first.cuh): two tiles interleaved in the epilogue, one activation box per tile
  static const bool first_on = diag_env("HESIC_TC_NO_FIRST") == nullptr;
  if (first_on && c->tc_kind == HESIC_TC_ROW2 && p.gdn == 1 && p.w_resident && y->fmt == HESIC_FMT_NHWC_SPLIT && c->Cout == 128 &&
      p.tma_store && y->W >= FIRST_BW && y->H >= FIRST_BH) {
    static PerDeviceOnce first_once;
    if (first_once.first())
      HESIC_CUDA(cudaFuncSetAttribute(conv_tc_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    Params q = p;
    q.bw = FIRST_BW; q.bh = FIRST_BH; q.bb = 1;
    q.tiles_x = (y->W + FIRST_BW - 1) / FIRST_BW; q.tiles_y = (y->H + FIRST_BH - 1) / FIRST_BH; q.tiles_b = y->B;
    q.n_tasks = q.tiles_x * q.tiles_y * q.tiles_b;
    CUtensorMap fa_hi, fa_lo, fw_hi, fw_lo, fy0, fy1;
    const uint64_t e = 2;
    {
      // K = 16 slices (32 B) of the row-pair interleaved ROWPAD input, 18 row pairs x 8 pixels
      const uint64_t pairB = (uint64_t)(x->W + HESIC_ROWPAD_X) * 8 * e, Hp2 = ((uint64_t)x->H + HESIC_ROWPAD_Y) / 2;
      uint64_t dims[5] = {(uint64_t)BK, (uint64_t)x->W / 2, 1, Hp2, (uint64_t)x->B};
      uint64_t strides[4] = {16 * e, pairB, pairB, Hp2 * pairB};
      uint32_t box[5] = {16u, (uint32_t)FIRST_BW, 1u, (uint32_t)FIRST_BH + 2u, 1u};
      int r = make_map(&fa_hi, x->p0, 5, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_32B);
      if (r == HESIC_OK) r = make_map(&fa_lo, x->p1, 5, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_32B);
      if (r != HESIC_OK) return r;
    }
    {
      uint64_t dims[3] = {(uint64_t)c->tc_k, (uint64_t)c->CoutPad, (uint64_t)c->tc_taps};
      uint64_t strides[2] = {(uint64_t)c->tc_k * 2, (uint64_t)c->tc_k * c->CoutPad * 2};
      uint32_t box[3] = {16u, 128u, 1u};
      int r = make_map(&fw_hi, c->w_hi, 3, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_32B);
      if (r == HESIC_OK) r = make_map(&fw_lo, c->w_lo, 3, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_32B);
      if (r != HESIC_OK) return r;
    }
    {
      // hi and lo planes of one allocation (lo above hi, 16-byte aligned distance): the unused third dimension of the map
      // becomes the plane and one store writes both
      const ptrdiff_t plane_d = (const char *)y->p1 - (const char *)y->p0;
      const bool planes = plane_d > 0 && plane_d % 16 == 0 && plane_d < ((ptrdiff_t)1 << 40);
      q.pl_phases = planes ? 2 : 1;
      uint64_t dims[5] = {(uint64_t)y->C, (uint64_t)y->W, planes ? 2u : 1u, (uint64_t)y->H, (uint64_t)y->B};
      uint64_t strides[4] = {(uint64_t)yCs * e, planes ? (uint64_t)plane_d : (uint64_t)y->W * yCs * e, (uint64_t)y->W * yCs * e,
                             (uint64_t)y->H * y->W * yCs * e};
      uint32_t box[5] = {16u, (uint32_t)FIRST_BW, planes ? 2u : 1u, 4u, 1u};
      int r = make_map(&fy0, y->p0, 5, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_32B);
      if (r == HESIC_OK) r = make_map(&fy1, y->p1, 5, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_32B);
      if (r != HESIC_OK) return r;
    }
    const int fgrid = std::min(q.n_tasks, num_sms);
    conv_tc_first_kernel<<<fgrid, FIRST_THREADS, FIRST_SMEM_BYTES, s>>>(fa_hi, fa_lo, fw_hi, fw_lo, m[2], m[3], fy0, fy1, q);
    HESIC_LAUNCHED("conv_tc_first_kernel");
    return HESIC_OK;
  }
  const int grid = std::min(p.n_tasks, num_sms);
  // fused GDN on a short K loop (the first analysis layer): 16 epilogue warps, 32 channels per thread
  static const bool epi16_on = diag_env("HESIC_TC_EPI8") == nullptr;
  if (epi16_on && p.gdn && !planar && p.tma_store && y->fmt == HESIC_FMT_NHWC_SPLIT && p.BN == 128 && p.n_tiles == 1 &&
      p.stg_sets == 2 && c->Cout == 128) {
    conv_tc_kernel<16><<<grid, 64 + 32 * 16, smem_bytes, s>>>(ma_hi, ma_lo, m[0], m[1], m[2], m[3], my0, my1, p);
    HESIC_LAUNCHED("conv_tc_kernel<16>");
    return HESIC_OK;
  }
  conv_tc_kernel<8><<<grid, NUM_THREADS, smem_bytes, s>>>(ma_hi, ma_lo, m[0], m[1], m[2], m[3], my0, my1, p);
  HESIC_LAUNCHED("conv_tc_kernel");
  return HESIC_OK;
}

}  // namespace hesic

namespace hesic { namespace en { int watchdog_status(unsigned int *dbg); } }
// 0 when no tcgen05 kernel has hit its watchdog since the last call; otherwise HESIC_E_CUDA with the
// first timed-out wait in the error text (tag: 1/9 producer empty, 2 acc_empty, 3/6 full, 4 x2_full,
// 7 acc_full, 8 norm_full).  Synchronises the device.
extern "C" int hesic_tc_status(void) {
  using namespace hesic;
  unsigned int flag = 0, dbg[8] = {0};
  HESIC_CUDA(cudaDeviceSynchronize());
  HESIC_CUDA(cudaMemcpyFromSymbol(&flag, tc::g_tc_abort, sizeof(flag)));
  if (!flag) {
    const int e = en::watchdog_status(dbg);   // the enhancement kernels keep their own flag (enhance.cu)
    if (e < 0) return cuda_fail(cudaGetLastError(), "hesic_tc_status");
    if (e == 0) return HESIC_OK;
    set_error("tcgen05 enhancement-conv watchdog: wait tag %u timed out (block %u thread %u parity %u)", dbg[0], dbg[1],
              dbg[2], dbg[3]);
    return HESIC_E_CUDA;
  }
  HESIC_CUDA(cudaMemcpyFromSymbol(dbg, tc::g_tc_dbg, sizeof(dbg)));
  unsigned int zero = 0;
  HESIC_CUDA(cudaMemcpyToSymbol(tc::g_tc_abort, &zero, sizeof(zero)));
  set_error("tcgen05 conv watchdog: wait tag %u timed out (block %u thread %u parity %u)", dbg[0], dbg[1], dbg[2], dbg[3]);
  return HESIC_E_CUDA;
}
