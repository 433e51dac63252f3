// First analysis layer (conv 3 -> 128, k5, s2, + fused GDN; newnet1.py:583-601, compressai/layers/gdn.py:55-70) as its own
// kernel: conv_tc_first_kernel.
//
// Why a layer-specific kernel.  The layer is HBM-sized work (587 MB per call at 16 x 512^2 -> 90 us at the measured peak) with a
// K loop of only 3 steps, so nothing hides the fused-GDN epilogue.  In conv_tc_kernel<16> a tile's chain
//     conv MMAs -> x (registers) -> x^2 -> TMEM -> GDN MMAs -> TMEM -> x * rsqrt(beta + norm) -> staging -> TMA store
// is ~5 k cycles of SERIAL latency and only one tile can be in the epilogue at a time, because x waits in registers between the
// two passes (r02 profile: 33 % of the epilogue's samples sit on norm_full, one instruction issued every ~7 cycles; 277 us).
// Here nothing waits in registers:
//   * pass 2 rebuilds |x| from the x^2 operand that is still in TMEM:  y = sign(x) * x^2 * rsqrt(x^2 * (beta + norm))
//     (IGDN: sign(x) * sqrt(x^2 * (beta + norm))) -- one MUFU per element as before; x^2 = hi + lo carries 16-17 mantissa
//     bits, i.e. a relative error <= 2^-18 in y, below the 2^-17 the SPLIT output format itself keeps.  The signs of a thread's
//     32 values travel as eight registers of packed sign bytes.
//   * so the epilogue warps run the passes of TWO tiles interleaved -- P1(a) P1(b) P2(a) P2(b) -- and every wait for the
//     tensor pipe (accumulator ready, norm ready) has a whole pass of the other tile in front of it.
//   * every epilogue warp owns a 32-lane x 32-column block of the accumulator from the first read to the TMA store: its x^2
//     operand is written into the columns it has just read (K = 16 slice m of the GDN contraction sits at columns
//     32 (m / 2) + 8 (m % 2), the lo part 16 columns further), its output goes through its own staging rows and its own
//     TMA store.  No block-level barrier is left in the loop (conv_tc_kernel<16>: three 512-thread and two 256-thread
//     barriers per tile).
//   * everything but the activations is RESIDENT in shared memory.  Only K = 48 of each 64-element row pair is contracted
//     (kx < 5 fills k < 40 of the [kx][row][slot] order), so the operands are kept as 32B-swizzled K = 16 slices (rows of
//     32 B, SWIZZLE_32B tensor maps and UMMA descriptors) instead of 128-byte rows: the weights take 72 KB instead of 96 KB,
//     an activation tile 27 KB, and next to them fit gamma (64 KB, 128B-swizzled as before), TWO activation buffers and the
//     staging rows -- the per-tile TMA traffic is the 27 KB of activations (conv_tc_kernel<16>: 96 KB + 64 KB of gamma
//     through a 2-stage ring; a first version of this kernel that still streamed gamma through a 5-slot ring ran 207 us and
//     went to 272 / 300 us with 3 / 2 slots).
//   * ONE activation box per tile, plane and K slice: the three kernel-row pairs read the same 18 x 8 window of row pairs
//     through UMMA descriptors shifted by one 256-byte swizzle atom (tile = 8 x 16 output pixels, so a row of the tile is
//     one atom).
// Warp roles: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..17 = epilogue (4 lane quadrants x 4 column groups).
#pragma once

namespace hesic {
namespace tc {

constexpr int FIRST_THREADS = 64 + 32 * 16;
constexpr int FIRST_BW = 8, FIRST_BH = 16;                       // tile: 8 x 16 output pixels
constexpr int FIRST_A_ROWS = (FIRST_BH + 2) * FIRST_BW;           // 144 box rows per plane and K = 16 step
constexpr int FIRST_AK_BYTES = FIRST_A_ROWS * 32;                 // one K = 16 slice of one plane: 144 rows x 32 B
constexpr int FIRST_A_BYTES = 2 * 3 * FIRST_AK_BYTES;             // (hi, lo) x 3 slices = 27 KB per tile
constexpr int FIRST_WK_BYTES = 256 * 32;                          // [Wh ; Wl] rows of one (row pair, K = 16 slice): 8 KB
constexpr int FIRST_W_BYTES = 9 * FIRST_WK_BYTES;                 // 72 KB
constexpr int FIRST_G_BYTES = 4 * 128 * 128;                      // gamma: 2 K chunks x (hi, lo) x [128][64], 64 KB
constexpr int FIRST_RC = 16;                                      // channels per store round of an epilogue warp
constexpr int FIRST_STG_BYTES = 16 * 2 * 32 * FIRST_RC * 2;       // 2 KB per epilogue warp
constexpr int FIRST_SMEM_BYTES = 1024 + FIRST_W_BYTES + FIRST_G_BYTES + 2 * FIRST_A_BYTES + FIRST_STG_BYTES + 512 + 1024;
static_assert(FIRST_SMEM_BYTES <= SMEM_LIMIT, "first-layer kernel: shared memory plan does not fit");

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
// y | (s & 0x80000000): the sign of s onto a non-negative y
__device__ __forceinline__ float with_sign(float y, uint32_t s) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xf8;" : "=r"(r) : "r"(__float_as_uint(y)), "r"(s));
  return __uint_as_float(r);
}
// K-major, 32B-swizzled shared-memory matrix descriptor: rows of 32 B (one K = 16 slice of bf16), 8-row groups 256 B apart
__device__ __forceinline__ uint64_t smem_desc_sw32(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}

__global__ void __launch_bounds__(FIRST_THREADS, 1)
conv_tc_first_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                     const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                     const __grid_constant__ CUtensorMap map_y0, const __grid_constant__ CUtensorMap map_y1,
                     const __grid_constant__ Params p) {
  constexpr int RC = FIRST_RC;
  constexpr uint32_t STG_WARP = 2u * 32u * (uint32_t)RC * 2u;     // hi rows + lo rows of one round
  constexpr uint32_t ROW_B = (uint32_t)RC * 2u;                   // bytes of a staging row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t w_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t g_base = w_base + (uint32_t)FIRST_W_BYTES;
  const uint32_t a_base = g_base + (uint32_t)FIRST_G_BYTES;
  const uint32_t stg_base = a_base + 2u * (uint32_t)FIRST_A_BYTES;
  const uint32_t bar_base = stg_base + (uint32_t)FIRST_STG_BYTES;
  auto a_full = [&](int b) { return bar_base + 8u * b; };
  auto a_empty = [&](int b) { return bar_base + 16u + 8u * b; };
  auto acc_full = [&](int b) { return bar_base + 128u + 8u * b; };
  auto acc_empty = [&](int b) { return bar_base + 144u + 8u * b; };
  auto x2_full = [&](int b) { return bar_base + 160u + 8u * b; };
  auto norm_full = [&](int b) { return bar_base + 176u + 8u * b; };
  const uint32_t w_full = bar_base + 192u, tmem_slot = bar_base + 200u;
  const uint32_t bias_s = bar_base + 512u, beta_s = bias_s + 512u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_local = ((int)blockIdx.x < p.n_tasks) ? (p.n_tasks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0 && lane == 0) {
    prefetch_map(&map_a_hi); prefetch_map(&map_a_lo); prefetch_map(&map_w_hi); prefetch_map(&map_w_lo);
    prefetch_map(&map_g_hi); prefetch_map(&map_g_lo); prefetch_map(&map_y0); prefetch_map(&map_y1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(a_full(b), 1); mbar_init(a_empty(b), 1);
      mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 512);
      mbar_init(x2_full(b), 512); mbar_init(norm_full(b), 1);
    }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // channel constants (one N tile: the same for every tile of the launch)
  if (threadIdx.x >= 64 && threadIdx.x < 192) {
    const int ci = (int)threadIdx.x - 64;
    st_shared_f32(bias_s + 4u * ci, __ldg(p.bias + ci));
    st_shared_f32(beta_s + 4u * ci, __ldg(p.beta + ci));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int txy = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // weights and gamma: once per CTA, resident
      mbar_expect_tx(w_full, (uint32_t)(FIRST_W_BYTES + FIRST_G_BYTES));
      for (int s = 0; s < 3; ++s)
        for (int k = 0; k < 3; ++k) {
          const uint32_t d = w_base + (uint32_t)(s * 3 + k) * FIRST_WK_BYTES;
          tma_load_3d(&map_w_hi, d, w_full, 16 * k, 0, s);
          tma_load_3d(&map_w_lo, d + 128u * 32u, w_full, 16 * k, 0, s);
        }
      for (int c = 0; c < 2; ++c) {
        tma_load_2d(&map_g_hi, g_base + (uint32_t)(2 * c) * 16384u, w_full, c * BK, 0);
        tma_load_2d(&map_g_lo, g_base + (uint32_t)(2 * c + 1) * 16384u, w_full, c * BK, 0);
      }
      for (int lt = 0; lt < n_local; ++lt) {
        const int buf = lt & 1;
        const int mt = (int)blockIdx.x + lt * (int)gridDim.x;
        const int tb = mt / txy, rr = mt - tb * txy;
        const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
        mbar_wait(a_empty(buf), (((uint32_t)lt >> 1) & 1u) ^ 1u, 1);
        mbar_expect_tx(a_full(buf), (uint32_t)FIRST_A_BYTES);
        const uint32_t d = a_base + (uint32_t)buf * FIRST_A_BYTES;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          tma_load_5d(&map_a_hi, d + (uint32_t)k * FIRST_AK_BYTES, a_full(buf), 16 * k, tx * FIRST_BW, 0, ty * FIRST_BH, tb);
          tma_load_5d(&map_a_lo, d + (uint32_t)(3 + k) * FIRST_AK_BYTES, a_full(buf), 16 * k, tx * FIRST_BW, 0, ty * FIRST_BH, tb);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc128 = instr_desc(128), idesc256 = instr_desc(256);
      mbar_wait(w_full, 0, 5);
      tc_fence_after();
      auto conv = [&](int lt) {
        const int buf = lt & 1;
        const uint32_t par = ((uint32_t)lt >> 1) & 1u;
        mbar_wait(acc_empty(buf), par ^ 1u, 2);
        mbar_wait(a_full(buf), par, 3);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)buf * ACC_STRIDE, d_small = d_main + COL_SMALL;
        const uint32_t a0 = a_base + (uint32_t)buf * FIRST_A_BYTES;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            // kernel-row pair s: the tile's A rows start s row pairs (8 pixels x 32 B = one swizzle atom) into the box
            const uint64_t a_hi = smem_desc_sw32(a0 + (uint32_t)k * FIRST_AK_BYTES + (uint32_t)s * 256u);
            const uint64_t a_lo = smem_desc_sw32(a0 + (uint32_t)(3 + k) * FIRST_AK_BYTES + (uint32_t)s * 256u);
            const uint64_t b = smem_desc_sw32(w_base + (uint32_t)(s * 3 + k) * FIRST_WK_BYTES);   // [Wh ; Wl]: 256 rows
            mma_ss(d_main, a_hi, b, idesc256, (s == 0 && k == 0) ? 0u : 1u);
            mma_ss(d_small, a_lo, b, idesc128, 1u);
          }
        }
        tc_commit(a_empty(buf));
        tc_commit(acc_full(buf));
      };
      auto gdn = [&](int lt) {
        const int buf = lt & 1;
        mbar_wait(x2_full(buf), ((uint32_t)lt >> 1) & 1u, 4);
        tc_fence_after();
        const uint32_t x2 = tmem_base + (uint32_t)buf * ACC_STRIDE, d = x2 + COL_SMALL;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint64_t b_hi = smem_desc(g_base + (uint32_t)(2 * c) * 16384u), b_lo = smem_desc(g_base + (uint32_t)(2 * c + 1) * 16384u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t o = (uint64_t)(k * 2);
            const int m = 4 * c + k;                         // K = 16 slice: channels [16 m, 16 m + 16)
            const uint32_t ah = x2 + (uint32_t)(32 * (m >> 1) + 8 * (m & 1)), al = ah + 16u;
            mma_ts(d, ah, b_hi + o, idesc128, (m == 0) ? 0u : 1u);
            mma_ts(d, ah, b_lo + o, idesc128, 1u);
            mma_ts(d, al, b_hi + o, idesc128, 1u);
          }
        }
        tc_commit(norm_full(buf));
      };
      for (int lt = 0; lt < n_local; lt += 2) {
        const bool two = lt + 1 < n_local;
        conv(lt);
        if (two) conv(lt + 1);
        gdn(lt);
        if (two) gdn(lt + 1);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t col0 = (uint32_t)(grp * 32);
    const uint32_t stg = stg_base + (uint32_t)(warp - 2) * STG_WARP;
    // staging rows: one pixel (lane) per row of ROW_B bytes, 16-byte chunks XOR-swizzled as the tensor map's swizzle mode
    // (32B: chunk ^ bit 7 of the address)
    const uint32_t sw = ((uint32_t)lane >> 2) & 1u;
    const uint32_t row_hi = stg + (uint32_t)lane * ROW_B, row_lo = row_hi + 32u * ROW_B;

    auto pass1 = [&](int lt, uint32_t (&sg)[8]) {
      const int buf = lt & 1;
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)buf * ACC_STRIDE + col0;
      mbar_wait(acc_full(buf), ((uint32_t)lt >> 1) & 1u, 7);
      tc_fence_after();
      float xs[32];
      {
        uint32_t r[32], q[32];
        tmem_ld32(acc, r);
        tmem_ld32(acc + COL_SMALL, q);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = ld_shared_f4(bias_s + 4u * col0 + 16u * j);
          xs[4 * j + 0] = (__uint_as_float(r[4 * j + 0]) + __uint_as_float(q[4 * j + 0])) + t.x;
          xs[4 * j + 1] = (__uint_as_float(r[4 * j + 1]) + __uint_as_float(q[4 * j + 1])) + t.y;
          xs[4 * j + 2] = (__uint_as_float(r[4 * j + 2]) + __uint_as_float(q[4 * j + 2])) + t.z;
          xs[4 * j + 3] = (__uint_as_float(r[4 * j + 3]) + __uint_as_float(q[4 * j + 3])) + t.w;
        }
      }
      // sign bytes: sg[j] = top bytes of xs[4j .. 4j+3]
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t t01 = prmt(__float_as_uint(xs[4 * j]), __float_as_uint(xs[4 * j + 1]), 0x0073u);
        const uint32_t t23 = prmt(__float_as_uint(xs[4 * j + 2]), __float_as_uint(xs[4 * j + 3]), 0x0073u);
        sg[j] = prmt(t01, t23, 0x5410u);
      }
      {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = xs[2 * j], c = xs[2 * j + 1];
          split_pair(a * a, c * c, hi[j], lo[j]);
        }
        // the x^2 operand goes into the columns this warp has just read: hi pairs at +0 (two K = 16 slices), lo pairs at +16
        tmem_st16(acc, hi);
        tmem_st16(acc + 16u, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(x2_full(buf));
    };

    auto pass2 = [&](int lt, const uint32_t (&sg)[8]) {
      const int buf = lt & 1;
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)buf * ACC_STRIDE + col0;
      const int mt = (int)blockIdx.x + lt * (int)gridDim.x;
      const int tb = mt / txy, rr = mt - tb * txy;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      mbar_wait(norm_full(buf), ((uint32_t)lt >> 1) & 1u, 8);
      tc_fence_after();
      uint32_t q[32], w[32];
      tmem_ld32(acc + COL_SMALL, q);
      tmem_ld32(acc, w);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(acc_empty(buf));          // norm and x^2 are in registers: the buffer goes back to the MMA warp
#pragma unroll
      for (int rd = 0; rd < 32 / RC; ++rd) {
        uint32_t oh[RC / 2], ol[RC / 2];
#pragma unroll
        for (int jj = 0; jj < RC / 4; ++jj) {
          const int c4 = rd * (RC / 4) + jj;                    // channels 4 c4 .. 4 c4 + 3 of the warp's 32
          const float4 bt = ld_shared_f4(beta_s + 4u * col0 + 16u * c4);
          const float be[4] = {bt.x, bt.y, bt.z, bt.w};
          const uint32_t s = sg[c4];
          float y[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t hw = w[2 * c4 + (e >> 1)], lw = w[16 + 2 * c4 + (e >> 1)];
            const float x2 = (e & 1) ? __uint_as_float(hw & 0xffff0000u) + __uint_as_float(lw & 0xffff0000u)
                                     : __uint_as_float(hw << 16) + __uint_as_float(lw << 16);
            const float n = __uint_as_float(q[4 * c4 + e]) + be[e];
            // |x| / sqrt(n) = x^2 * rsqrt(x^2 * n); the 1e-30 keeps x = 0 away from 0 * inf
            const float v = x2 * rsqrt_approx(fmaf(x2, n, 1e-30f));
            y[e] = with_sign(v, s << (24 - 8 * e));
          }
          split_pair(y[0], y[1], oh[2 * jj], ol[2 * jj]);
          split_pair(y[2], y[3], oh[2 * jj + 1], ol[2 * jj + 1]);
        }
        // the warp's previous store out of these staging rows has been read (it was issued a whole round of arithmetic ago)
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < RC / 8; ++ch) {
          const uint32_t o = ((uint32_t)ch ^ sw) << 4;
          st_shared_v4(row_hi + o, oh[4 * ch], oh[4 * ch + 1], oh[4 * ch + 2], oh[4 * ch + 3]);
          st_shared_v4(row_lo + o, ol[4 * ch], ol[4 * ch + 1], ol[4 * ch + 2], ol[4 * ch + 3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && !(p.pl_os & 1)) {
          const int c0 = (int)col0 + rd * RC;
          tma_store_5d(&map_y0, stg, c0, tx * FIRST_BW, 0, ty * FIRST_BH + 4 * quad, tb);
          tma_store_5d(&map_y1, stg + 32u * ROW_B, c0, tx * FIRST_BW, 0, ty * FIRST_BH + 4 * quad, tb);
          bulk_commit();
        }
      }
    };

    uint32_t sg[2][8];
    for (int lt = 0; lt < n_local; lt += 2) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (lt + h < n_local) pass1(lt + h, sg[h]);
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (lt + h < n_local) pass2(lt + h, sg[h]);
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace hesic
