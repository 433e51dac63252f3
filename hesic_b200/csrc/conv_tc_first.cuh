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
//     -- one MUFU per element as before; x^2 = hi + lo carries 16-17 mantissa bits, i.e. a relative error <= 2^-18 in y,
//     below the 2^-17 the SPLIT output format itself keeps.  The signs of a thread's 32 values travel as 16 registers
//     holding the sign bytes of a channel pair at the sign positions of a packed bf16 pair (one LOP3 applies them to the
//     hi and to the lo word of the output).  (The layer is GDN; an IGDN would be sign(x) * sqrt(x^2 * (beta + norm)).)
//   * so the epilogue warps run the passes of TWO tiles interleaved -- P1(a) P1(b) P2(a) P2(b) -- and every wait for the
//     tensor pipe (accumulator ready, norm ready) has a whole pass of the other tile in front of it.
//   * every epilogue warp owns a 32-lane x 32-column block of the accumulator from the first read to the TMA store: its x^2
//     operand is written into the columns it has just read (K = 16 slice m of the GDN contraction sits at columns
//     32 (m / 2) + 8 (m % 2), the lo part 16 columns further), beta is written into its norm columns (the GDN MMAs accumulate
//     onto it), its output goes through its own 2 KB of staging rows and its own TMA store (both planes in one store when
//     they are planes of one allocation).  No block-level barrier is left in the loop (conv_tc_kernel<16>: three 512-thread
//     and two 256-thread barriers per tile).
//   * K is 144: the three bf16x3 terms share ONE fp32 accumulator, so pass 1 reads 32 columns instead of 64.
//   * producer, MMA and store-issuing threads are chosen with elect.sync (tc_ptx.cuh: elect_one): with `lane == 0` every
//     UTCHMMA sat in a ~10-instruction serialisation loop and the MMA thread (51 small MMAs per tile) bound the kernel.
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
// Measured (16 x 512^2, L2 flushed, profiles/r03_first_layer.md): 277 us (conv_tc_kernel<16>) -> 183 us; 150 us with the TMA
// stores switched off -- what is left is the store engine's row rate (a warp's round is 32-byte rows: 2048 rows per tile).
// Two variants with wider rows were built and measured slower: the four warps of a lane quadrant sharing 128-byte rows
// (mbarrier hand-shake inside the quad: 206 us) and plane-wise rounds of 64-byte rows (the second round waits for the first
// store to drain: 210 us).  Four-kilobyte rounds per warp (64-byte rows, no waiting) miss the shared-memory budget by 2.5 KB.
// No staging at all -- every thread writing its 16 channels of a plane with one 256-bit st.global.v8 (one full sector) -- was
// measured at 217 us: 32 distinct lines per store instruction keep the LSU busier than the staged path keeps the TMA engine.
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
constexpr int FIRST_STG_BYTES = 16 * 2048;                        // per epilogue warp: 32 pixels x 16 channels x (hi, lo)

constexpr int FIRST_SMEM_BYTES = 1024 + FIRST_W_BYTES + FIRST_G_BYTES + 2 * FIRST_A_BYTES + FIRST_STG_BYTES + 512 + 1024;
static_assert(FIRST_SMEM_BYTES <= SMEM_LIMIT, "first-layer kernel: shared memory plan does not fit");

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
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
    if (elect_one()) {
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
    if (elect_one()) {
      const uint32_t idesc128 = instr_desc(128);
      mbar_wait(w_full, 0, 5);
      tc_fence_after();
      auto conv = [&](int lt) {
        const int buf = lt & 1;
        const uint32_t par = ((uint32_t)lt >> 1) & 1u;
        mbar_wait(acc_empty(buf), par ^ 1u, 2);
        mbar_wait(a_full(buf), par, 3);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)buf * ACC_STRIDE;
        const uint32_t a0 = a_base + (uint32_t)buf * FIRST_A_BYTES;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            // kernel-row pair s: the tile's A rows start s row pairs (8 pixels x 32 B = one swizzle atom) into the box
            const uint64_t a_hi = smem_desc_sw32(a0 + (uint32_t)k * FIRST_AK_BYTES + (uint32_t)s * 256u);
            const uint64_t a_lo = smem_desc_sw32(a0 + (uint32_t)(3 + k) * FIRST_AK_BYTES + (uint32_t)s * 256u);
            const uint64_t b_hi = smem_desc_sw32(w_base + (uint32_t)(s * 3 + k) * FIRST_WK_BYTES);
            const uint64_t b_lo = smem_desc_sw32(w_base + (uint32_t)(s * 3 + k) * FIRST_WK_BYTES + 128u * 32u);
            // K is 144 here: the three bf16x3 terms can share ONE fp32 accumulator (the main / small separation of the
            // K-heavy layers guards against the tensor core's truncating accumulate over thousands of steps)
            mma_ss(d_main, a_hi, b_hi, idesc128, (s == 0 && k == 0) ? 0u : 1u);
            mma_ss(d_main, a_hi, b_lo, idesc128, 1u);
            mma_ss(d_main, a_lo, b_hi, idesc128, 1u);
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
            mma_ts(d, ah, b_hi + o, idesc128, 1u);          // onto beta (written by pass 1)
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
    // Warp (quad, grp) owns TMEM lanes [32 quad, +32) and columns / channels [32 grp, +32) from the first read to the store.
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t col0 = (uint32_t)(grp * 32);
    constexpr uint32_t ROW_B = 32u;                               // a staging row: 16 channels of one pixel and plane
    const uint32_t stg = stg_base + (uint32_t)(warp - 2) * 2048u;
    // staging rows: 16-byte chunks XOR-swizzled as the tensor map's SWIZZLE_32B mode (chunk ^ bit 7 of the address)
    const uint32_t sw = ((uint32_t)lane >> 2) & 1u;
    // one store per round covers both planes when they are planes of one allocation (the '1' dimension of the output map
    // becomes the plane): staging rows ordered [y][plane][x]; otherwise hi rows, then lo rows, and one store per plane
    const bool planes = p.pl_phases == 2;
    const uint32_t row_hi = stg + (planes ? (uint32_t)((lane >> 3) * 16 + (lane & 7)) : (uint32_t)lane) * ROW_B;
    const uint32_t row_lo = row_hi + (planes ? 8u : 32u) * ROW_B;
    const bool issuer = elect_one();

    // pass 1: x = conv + bias; x^2 -> bf16 (hi, lo) operand over the columns just read; beta -> the norm columns, so that the
    // GDN contraction accumulates onto it
    auto pass1 = [&](int lt, uint32_t (&sg)[16]) {
      const int buf = lt & 1;
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)buf * ACC_STRIDE + col0;
      mbar_wait(acc_full(buf), ((uint32_t)lt >> 1) & 1u, 7);
      tc_fence_after();
      float xs[32];
      {
        uint32_t r[32];
        tmem_ld32(acc, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = ld_shared_f4(bias_s + 4u * col0 + 16u * j);
          xs[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + t.x;
          xs[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + t.y;
          xs[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + t.z;
          xs[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + t.w;
        }
      }
      {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = xs[2 * j], c = xs[2 * j + 1];
          // sign bytes of the pair at the sign positions of a packed bf16 pair (bytes 1 and 3)
          sg[j] = prmt(__float_as_uint(a), __float_as_uint(c), 0x7030u);
          split_pair(a * a, c * c, hi[j], lo[j]);
        }
        // the x^2 operand goes into the columns this warp has just read: hi pairs at +0 (two K = 16 slices), lo pairs at +16
        tmem_st16(acc, hi);
        tmem_st16(acc + 16u, lo);
      }
      {
        uint32_t bt[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = ld_shared_f4(beta_s + 4u * col0 + 16u * j);
          bt[4 * j] = __float_as_uint(t.x); bt[4 * j + 1] = __float_as_uint(t.y);
          bt[4 * j + 2] = __float_as_uint(t.z); bt[4 * j + 3] = __float_as_uint(t.w);
        }
        tmem_st32(acc + COL_SMALL, bt);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(x2_full(buf));
    };

    // pass 2: y = sign(x) * x^2 * rsqrt(x^2 * (beta + norm)) from the x^2 operand and the norm, both still in TMEM
    auto pass2 = [&](int lt, const uint32_t (&sg)[16]) {
      const int buf = lt & 1;
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)buf * ACC_STRIDE + col0;
      const int mt = (int)blockIdx.x + lt * (int)gridDim.x;
      const int tb = mt / txy, rr = mt - tb * txy;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      mbar_wait(norm_full(buf), ((uint32_t)lt >> 1) & 1u, 8);
      tc_fence_after();
#pragma unroll
      for (int rd = 0; rd < 2; ++rd) {
        // this round's 16 channels: norm (16 columns), x^2 hi pairs (8) and lo pairs (8)
        uint32_t q[16], wh[8], wl[8];
        tmem_ld16(acc + COL_SMALL + 16u * rd, q);
        tmem_ld8(acc + 8u * rd, wh);
        tmem_ld8(acc + 16u + 8u * rd, wl);
        tmem_ld_wait();
        if (rd == 1) {
          tc_fence_before();
          mbar_arrive(acc_empty(buf));      // norm and x^2 are in registers: the buffer goes back to the MMA warp
        }
        uint32_t oh[8], ol[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = rd * 8 + jj;                            // channel pair (2j, 2j + 1) of the warp's 32
          const uint32_t hw = wh[jj], lw = wl[jj];
          // (leaving the first value's bits below the second lo part's mantissa would save a mask per pair but costs 2^-17 of
          // y: measured 4.5e-5 -> 8.0e-5 of the rms on the largest outputs)
          const float xa = __uint_as_float(hw << 16) + __uint_as_float(lw << 16);
          const float xb = __uint_as_float(hw & 0xffff0000u) + __uint_as_float(lw & 0xffff0000u);
          // |x| / sqrt(n) = x^2 * rsqrt(x^2 * n); the 1e-30 keeps x = 0 away from 0 * inf
          const float ya = xa * rsqrt_approx(fmaf(xa, __uint_as_float(q[2 * jj]), 1e-30f));
          const float yb = xb * rsqrt_approx(fmaf(xb, __uint_as_float(q[2 * jj + 1]), 1e-30f));
          uint32_t h, l;
          split_pair(ya, yb, h, l);
          // -(hi + lo) = (-hi) + (-lo): the signs go onto both packed pairs
          asm("lop3.b32 %0, %1, %2, 0x80008000, 0x78;" : "=r"(oh[jj]) : "r"(h), "r"(sg[j]));
          asm("lop3.b32 %0, %1, %2, 0x80008000, 0x78;" : "=r"(ol[jj]) : "r"(l), "r"(sg[j]));
        }
        // the warp's previous store out of these staging rows has been read (it was issued a whole round of arithmetic ago)
        if (issuer) bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const uint32_t o = ((uint32_t)ch ^ sw) << 4;
          st_shared_v4(row_hi + o, oh[4 * ch], oh[4 * ch + 1], oh[4 * ch + 2], oh[4 * ch + 3]);
          st_shared_v4(row_lo + o, ol[4 * ch], ol[4 * ch + 1], ol[4 * ch + 2], ol[4 * ch + 3]);
        }
        fence_async_smem();
        __syncwarp();
        if (issuer) {
          const int c0 = (int)col0 + rd * 16;
          tma_store_5d(&map_y0, stg, c0, tx * FIRST_BW, 0, ty * FIRST_BH + 4 * quad, tb);
          if (!planes) tma_store_5d(&map_y1, stg + 32u * ROW_B, c0, tx * FIRST_BW, 0, ty * FIRST_BH + 4 * quad, tb);
          bulk_commit();
        }
      }
    };

    uint32_t sg[2][16];
    for (int lt = 0; lt < n_local; lt += 2) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (lt + h < n_local) pass1(lt + h, sg[h]);
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (lt + h < n_local) pass2(lt + h, sg[h]);
    }
    bulk_wait_all();     // bulk groups belong to the thread that committed them (the elected lane); a no-op for the others
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
