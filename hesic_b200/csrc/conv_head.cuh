// RGB synthesis head: nn.ConvTranspose2d(Cin -> Cout <= 4, k5, s2, p2, op1) (compressai/models/utils.py:112-118,
// newnet1.py:612,670) on tcgen05 in SCATTER form.  Included by conv_tc.cu (shares its PTX wrappers).
//
// A transposed convolution writes, for every INPUT pixel q, a 5x5xCout patch  P[q][ky][kx][co] =
// sum_ci x[q][ci] * w[ci][co][ky][kx]  onto the output at (2q - 2 + k).  The patch is a plain GEMM:
//     P[pixel, n] = sum_ci A[pixel, ci] * Wn[n, ci],   n = (ky*5 + kx)*Cout + co,  N = 25*Cout (75 -> 80)
// so every activation is fetched ONCE (the gather form needed nine shifted 128-pixel boxes per tile and was
// bound by L2->SM fills at 8 % tensor-pipe utilisation), the weights (N x Cin, 40 KB as bf16 hi/lo) stay
// resident in shared memory for the whole kernel, and the MMAs have N = 80 instead of N = 16.
// The overlap-add (col2im) happens in the epilogue: the tile's patches go TMEM -> shared memory, then every
// output pixel sums its 9/6/6/4 contributions in a fixed order.  Tiles are 8 x 16 input pixels with a
// one-pixel halo that is recomputed by the neighbouring tile (interior 6 x 14), so no atomics and no
// cross-CTA traffic; pixels outside the image are TMA zero fill and contribute zeros.
// bf16x3 split accumulation, warp roles, mbarrier rings and the watchdog are those of conv_tc_kernel.
#pragma once

namespace hesic {
namespace tc {
namespace head {

constexpr int NT = 320;                  // TMA warp, MMA warp, 8 epilogue warps (r03: 4 left the overlap-add of a tile -- 168
                                         // items -- to 128 threads in two unequal rounds: 153-162 us per launch)
constexpr int EPI_T = NT - 64;
constexpr int TILE_W = 16, TILE_H = 8;   // input pixels per tile (UMMA M = 128), halo included
constexpr int IN_W = TILE_W - 2, IN_H = TILE_H - 2;
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES;   // one 64-channel chunk, hi + lo

struct HParams {
  int tiles_x, tiles_y, n_tasks;
  int H, W, B;             // input size
  int NPAD, kchunks, stages;
  float *y;                // NCHW fp32 [B][out_Cs][2H][2W], offset to channel 0 of the view
  int out_Cs;
  const float *bias;
  int gdn;                 // 0 none, 1 GDN, 2 inverse GDN over the Cout channels (registers)
  const float *beta, *gamma;
  int act;
};

__device__ __forceinline__ void epi_bar_head() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

template <int COUT>
__global__ void __launch_bounds__(NT, 1)
conv_head_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                 const __grid_constant__ HParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_bytes = (uint32_t)p.NPAD * 128u;                        // one (chunk, plane) weight tile
  const uint32_t w_base = smem_base;                                        // [kchunks][hi, lo][NPAD][128 B]
  const uint32_t stage_base = w_base + (uint32_t)p.kchunks * 2u * w_bytes;  // [stages][hi, lo][128 px][128 B]
  const uint32_t d_base = stage_base + (uint32_t)p.stages * STAGE_BYTES;    // fp32 [128 px][NPAD + 1]
  const int pitch = p.NPAD + 1;
  const uint32_t bar_base = d_base + (uint32_t)(BM * pitch * 4);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto acc_full = [&](int b) { return bar_base + 128u + 8u * b; };
  auto acc_empty = [&](int b) { return bar_base + 144u + 8u * b; };
  const uint32_t w_full = bar_base + 160u, tmem_slot = bar_base + 192u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_map(&map_a_hi); prefetch_map(&map_a_lo); prefetch_map(&map_w_hi); prefetch_map(&map_w_lo);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), EPI_T); }
    mbar_init(w_full, 1);
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
  const int txy = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (elect_one()) {
      // weights: resident for the whole kernel
      mbar_expect_tx(w_full, (uint32_t)p.kchunks * 2u * w_bytes);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        tma_load_2d(&map_w_hi, w_base + (uint32_t)(2 * kc) * w_bytes, w_full, kc * BK, 0);
        tma_load_2d(&map_w_lo, w_base + (uint32_t)(2 * kc + 1) * w_bytes, w_full, kc * BK, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int task = blockIdx.x; task < p.n_tasks; task += gridDim.x) {
        const int b = task / txy, rr = task - b * txy;
        const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
        const int x0 = tx * IN_W - 1, y0 = ty * IN_H - 1;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 1);
          const uint32_t sa = stage_base + (uint32_t)stage * STAGE_BYTES, fb = full_bar(stage);
          mbar_expect_tx(fb, STAGE_BYTES);
          tma_load_5d(&map_a_hi, sa, fb, kc * BK, x0, 0, y0, b);
          tma_load_5d(&map_a_lo, sa + A_TILE_BYTES, fb, kc * BK, x0, 0, y0, b);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(w_full, 0, 5);
      tc_fence_after();
      const uint32_t idesc = instr_desc(p.NPAD);
      int stage = 0, lt = 0;
      uint32_t phase = 0;
      for (int task = blockIdx.x; task < p.n_tasks; task += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(acc_empty(buf), (((uint32_t)lt >> 1) & 1u) ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)buf * ACC_STRIDE, d_small = d_main + COL_SMALL;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(full_bar(stage), phase, 3);
          tc_fence_after();
          const uint32_t sa = stage_base + (uint32_t)stage * STAGE_BYTES;
          const uint64_t a_hi = smem_desc(sa), a_lo = smem_desc(sa + A_TILE_BYTES);
          const uint64_t b_hi = smem_desc(w_base + (uint32_t)(2 * kc) * w_bytes);
          const uint64_t b_lo = smem_desc(w_base + (uint32_t)(2 * kc + 1) * w_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t o = (uint64_t)(k * 2);
            const uint32_t acc = (kc == 0 && k == 0) ? 0u : 1u;
            mma_ss(d_main, a_hi + o, b_hi + o, idesc, acc);
            mma_ss(d_small, a_hi + o, b_lo + o, idesc, acc);
            mma_ss(d_small, a_lo + o, b_hi + o, idesc, 1u);
          }
          tc_commit(empty_bar(stage));
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(acc_full(buf));
      }
    }
  } else {
    // ===================== epilogue: TMEM -> smem patches -> overlap-add -> NCHW =====================
    // warp w reads TMEM lane quadrant w % 4; the two warps of a quadrant split the tile's N columns in 16-column chunks
    const int quad = warp & 3, row = quad * 32 + lane, tid = (int)threadIdx.x - 64, half = (warp - 2) >> 2;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int Ho = 2 * p.H, Wo = 2 * p.W;
    float bia[COUT], bet[COUT], gam[COUT][COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      bia[c] = __ldg(p.bias + c);
      bet[c] = p.gdn ? __ldg(p.beta + c) : 1.f;
#pragma unroll
      for (int j = 0; j < COUT; ++j) gam[j][c] = p.gdn ? __ldg(p.gamma + j * COUT + c) : 0.f;
    }
    int lt = 0;
    for (int task = blockIdx.x; task < p.n_tasks; task += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int b = task / txy, rr = task - b * txy;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      const int x0 = tx * IN_W - 1, y0 = ty * IN_H - 1;
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)buf * ACC_STRIDE;
      mbar_wait(acc_full(buf), ((uint32_t)lt >> 1) & 1u, 7);
      tc_fence_after();
      epi_bar_head();   // the previous tile's overlap-add has finished reading the patch buffer
      const uint32_t drow = d_base + (uint32_t)(row * pitch) * 4u;
      for (int c0 = 16 * half; c0 < p.NPAD; c0 += 32) {
        uint32_t r[16], q[16];
        tmem_ld16(acc + c0, r);
        tmem_ld16(acc + COL_SMALL + c0, q);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) st_shared_f32(drow + 4u * (c0 + j), __uint_as_float(r[j]) + __uint_as_float(q[j]));
      }
      tc_fence_before();
      mbar_arrive(acc_empty(buf));   // patches are in shared memory: release the accumulator
      epi_bar_head();
      // one item = interior input pixel (iy, ix) x output row parity ry -> two horizontally adjacent outputs
      for (int item = tid; item < IN_H * IN_W * 2; item += EPI_T) {
        const int ry = item / (IN_H * IN_W), ip = item - ry * (IN_H * IN_W);
        const int iy = ip / IN_W + 1, ix = ip - (ip / IN_W) * IN_W + 1;
        const int gy = y0 + iy, gx = x0 + ix;
        if (gy >= p.H || gx >= p.W) continue;
        float o[2][COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) { o[0][c] = 0.f; o[1][c] = 0.f; }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int ky = ry + 2 - 2 * dy;          // out row 2*gy + ry = 2*(gy + dy) - 2 + ky
          if (ky > 4) continue;                     // (ry = 1, dy = -1)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const uint32_t src = d_base + (uint32_t)(((iy + dy) * TILE_W + ix + dx) * pitch) * 4u;
            const int n0 = (ky * 5 + 2 - 2 * dx) * COUT;       // rx = 0: kx = 2 - 2 dx
#pragma unroll
            for (int c = 0; c < COUT; ++c) o[0][c] += ld_shared_f32(src + 4u * (n0 + c));
            if (dx >= 0) {                                        // rx = 1: kx = 3 - 2 dx
#pragma unroll
              for (int c = 0; c < COUT; ++c) o[1][c] += ld_shared_f32(src + 4u * (n0 + COUT + c));
            }
          }
        }
#pragma unroll
        for (int rx = 0; rx < 2; ++rx) {
          float x[COUT];
#pragma unroll
          for (int c = 0; c < COUT; ++c) x[c] = o[rx][c] + bia[c];
          if (p.gdn) {
            float sq[COUT], t[COUT];
#pragma unroll
            for (int c = 0; c < COUT; ++c) sq[c] = x[c] * x[c];
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
              float nrm = bet[c];
#pragma unroll
              for (int j = 0; j < COUT; ++j) nrm = fmaf(gam[j][c], sq[j], nrm);
              t[c] = x[c] * (p.gdn == 2 ? sqrt_approx(nrm) : rsqrt_approx(nrm));
            }
#pragma unroll
            for (int c = 0; c < COUT; ++c) x[c] = t[c];
          }
#pragma unroll
          for (int c = 0; c < COUT; ++c) o[rx][c] = apply_act(x[c], p.act);
        }
        const int oy = 2 * gy + ry, ox = 2 * gx;
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          float *dst = p.y + (((size_t)b * p.out_Cs + c) * Ho + oy) * Wo + ox;
          if ((((uintptr_t)dst) & 7u) == 0) *reinterpret_cast<float2 *>(dst) = make_float2(o[0][c], o[1][c]);
          else { dst[0] = o[0][c]; dst[1] = o[1][c]; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace head
}  // namespace tc
}  // namespace hesic
