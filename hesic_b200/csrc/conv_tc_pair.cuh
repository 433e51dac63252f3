// CTA-pair (cta_group::2) variant of conv_tc_kernel for the K-heavy layers (plain or with the fused GDN): two SMs of one TPC
// work on two pixel tiles of the same (phase, N tile) and share the weight operand.
//
// Why: conv_tc_kernel is bound by shared-memory bandwidth, not by the tensor pipe (r01 ncu of g_a_conv2: tensor pipe
// active 68 %, tensor-core operand reads 56 % of the smem pipe + 64 KB of TMA fills per k-step).  Per K = 16 step one
// CTA reads A (hi, lo) 8 KB + B 12 KB; in a pair each CTA supplies only HALF of every B operand:
//     MMA1  D[main | small] (M 256 x N 256)  = Ah . [Wh ; Wl]     B rows 0..127 (Wh) from CTA 0, 128..255 (Wl) from CTA 1
//     MMA2  D[small]        (M 256 x N 128) += Al . Wh            B rows 0..63 from CTA 0, 64..127 from CTA 1
// so a CTA reads 8 + 6 = 14 KB per step instead of 20 and receives 56 KB of TMA fills per k-step instead of 64.
// Each CTA keeps its own 128 accumulator rows in its own TMEM; the epilogue is the single-CTA one.
//
// Protocol (rank 0 = leader): both producers load their A tiles and B parts with cp.async.bulk.tensor .cta_group::2
// completing on the LEADER's full barrier; the leader's MMA thread issues tcgen05.mma.cta_group::2 and multicasts its
// commits to both CTAs' empty / acc_full barriers; the peer's epilogue threads release the accumulator buffer by a
// remote arrive on the leader's acc_empty barrier.
#pragma once

namespace hesic {
namespace tc {

constexpr int PAIR_B1_BYTES = 128 * 128, PAIR_B2_BYTES = 64 * 128;
constexpr int PAIR_STAGE_BYTES = 2 * A_TILE_BYTES + PAIR_B1_BYTES + PAIR_B2_BYTES;   // 56 KB

// pair task -> (pixel-tile pair, phase, N tile); CTA `rank` of the pair takes pixel tile 2 * mtp + rank
__device__ __forceinline__ TaskCoord decode_pair_task(const Params &p, int task, int rank) {
  // N-tile major: the last N tile of a layer may be narrower (Cout = 192: 128 + 64 columns), and with the N tile as the
  // fastest index an even grid stride gave every CTA pair the same N tile in every round -- half of the pairs two wide tasks,
  // the other half two narrow ones (conv 128 -> 192 at 32 x 32: 128 tasks on 74 pairs).  Wide tasks first, narrow ones on top.
  TaskCoord t;
  const int per_nt = p.n_tasks / p.n_tiles;
  t.nt = task / per_nt;
  const int r = task - t.nt * per_nt;
  const int mtp = r / p.n_phases;
  t.ph = (r + mtp) % p.n_phases;
  t.mt = 2 * mtp + rank;
  return t;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                    const __grid_constant__ CUtensorMap map_w_hi64, const __grid_constant__ CUtensorMap map_g_hi64,
                    const __grid_constant__ CUtensorMap map_g_lo64, const __grid_constant__ CUtensorMap map_y0,
                    const __grid_constant__ CUtensorMap map_y1, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + (uint32_t)p.stages * PAIR_STAGE_BYTES;
  const uint32_t bar_base = stg_base + (uint32_t)STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto acc_full = [&](int b) { return bar_base + 128u + 8u * b; };
  auto acc_empty = [&](int b) { return bar_base + 144u + 8u * b; };
  const uint32_t x2_full = bar_base + 160u, norm_full = bar_base + 168u;
  const uint32_t tmem_slot = bar_base + 192u;
  const uint32_t bias_s = bar_base + BAR_BYTES, beta_s = bias_s + 512u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = p.n_tasks;                 // host passes the number of PAIR tasks
  const int first = blockIdx.x >> 1, step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_map(&map_a_hi); prefetch_map(&map_a_lo); prefetch_map(&map_w_hi); prefetch_map(&map_w_lo);
    prefetch_map(&map_w_hi64); prefetch_map(&map_y0); prefetch_map(&map_y1);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 2 * EPI_THREADS); }
    mbar_init(x2_full, 2 * EPI_THREADS); mbar_init(norm_full, 1);
    if (p.gdn) { prefetch_map(&map_g_hi64); prefetch_map(&map_g_lo64); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  const int txy = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // the gamma tiles of a GDN step: each CTA loads its 64-row half of gamma_hi and gamma_lo (16 KB) for K chunk c
      auto gdn_step = [&](int c) {
        mbar_wait(empty_bar(stage), phase ^ 1u, 9);
        const uint32_t sa = smem_base + (uint32_t)stage * PAIR_STAGE_BYTES;
        const uint32_t fb = mapa_cta(full_bar(stage), 0);
        if (leader) mbar_expect_tx(full_bar(stage), 2u * 2u * (uint32_t)PAIR_B2_BYTES);
        tma_load_2d_pair(&map_g_hi64, sa + 2 * A_TILE_BYTES, fb, c * BK, 64 * (int)rank);
        tma_load_2d_pair(&map_g_lo64, sa + 2 * A_TILE_BYTES + PAIR_B2_BYTES, fb, c * BK, 64 * (int)rank);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      };
      int lt = 0;
      for (int task = first; task < n_pairs; task += step, ++lt) {
        const TaskCoord tk = decode_pair_task(p, task, (int)rank);
        const int tb = tk.mt / txy, rr = tk.mt - tb * txy;
        const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
        const int t0 = p.tap_begin[tk.ph], t1 = p.tap_begin[tk.ph + 1];
        const int kc0 = p.kc_lo[tk.nt & 7], kc1 = p.kc_hi[tk.nt & 7];   // block-banded layers skip empty K chunks
        const int ksteps = (t1 - t0) * (kc1 - kc0);
        const int g = (p.gdn && lt > 0) ? min(p.gdn_at, ksteps - 1) : -1;
        int ks = 0;
        for (int t = t0; t < t1; ++t) {
          const Tap tap = p.taps[t];
          for (int kc = kc0; kc < kc1; ++kc, ++ks) {
            if (ks == g) { gdn_step(0); gdn_step(1); }
            mbar_wait(empty_bar(stage), phase ^ 1u, 1);
            const uint32_t sa = smem_base + (uint32_t)stage * PAIR_STAGE_BYTES;
            const uint32_t fb = mapa_cta(full_bar(stage), 0);       // the leader's barrier collects both CTAs' bytes
            if (leader) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)PAIR_STAGE_BYTES);
            const int c = kc * BK + tap.px * p.in_Cs;
            const int x = tx * p.bw + tap.dx, y = ty * p.bh + tap.dy, b = tb * p.bb;
            tma_load_5d_pair(&map_a_hi, sa, fb, c, x, tap.py, y, b);
            tma_load_5d_pair(&map_a_lo, sa + A_TILE_BYTES, fb, c, x, tap.py, y, b);
            // B1: the leader holds Wh (-> main columns), the peer Wl (-> small columns); B2: each a 64-row half of Wh
            tma_load_3d_pair(leader ? &map_w_hi : &map_w_lo, sa + 2 * A_TILE_BYTES, fb, kc * BK, tk.nt * 128, tap.w);
            tma_load_3d_pair(&map_w_hi64, sa + 2 * A_TILE_BYTES + PAIR_B1_BYTES, fb, kc * BK, tk.nt * 128 + 64 * (int)rank, tap.w);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (p.gdn && lt > 0) { gdn_step(0); gdn_step(1); }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      int stage = 0, lt = 0;
      uint32_t phase = 0;
      const uint32_t idesc256 = instr_desc_pair(256), idesc128 = instr_desc_pair(128);
      // GDN(t): norm = x^2 . gamma^T with the x^2 operand (hi | lo) in each CTA's TMEM main columns, accumulated into small
      auto gdn_step = [&](int glt, int c) {
        const int gbuf = glt & 1;
        if (c == 0) {
          mbar_wait(x2_full, (uint32_t)glt & 1u, 4);
          tc_fence_after();
        }
        mbar_wait(full_bar(stage), phase, 6);
        tc_fence_after();
        const uint32_t sa = smem_base + (uint32_t)stage * PAIR_STAGE_BYTES;
        const uint64_t g_hi = smem_desc(sa + 2 * A_TILE_BYTES), g_lo = smem_desc(sa + 2 * A_TILE_BYTES + PAIR_B2_BYTES);
        const uint32_t x2 = tmem_base + (uint32_t)gbuf * ACC_STRIDE, d = x2 + COL_SMALL;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t o = (uint64_t)(k * 2);
          const uint32_t ah = x2 + (uint32_t)(c * 32 + k * 8);
          const uint32_t al = x2 + 64u + (uint32_t)(c * 32 + k * 8);
          mma_ts_pair(d, ah, g_hi + o, idesc128, (c == 0 && k == 0) ? 0u : 1u);
          mma_ts_pair(d, ah, g_lo + o, idesc128, 1u);
          mma_ts_pair(d, al, g_hi + o, idesc128, 1u);
        }
        tc_commit_pair(empty_bar(stage));
        if (c == 1) tc_commit_pair(norm_full);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      };
      for (int task = first; task < n_pairs; task += step, ++lt) {
        const TaskCoord tk = decode_pair_task(p, task, 0);
        const int buf = lt & 1;
        const int ksteps = (p.tap_begin[tk.ph + 1] - p.tap_begin[tk.ph]) * (p.kc_hi[tk.nt & 7] - p.kc_lo[tk.nt & 7]);
        const int g = (p.gdn && lt > 0) ? min(p.gdn_at, ksteps - 1) : -1;
        const uint32_t d_main = tmem_base + (uint32_t)buf * ACC_STRIDE, d_small = d_main + COL_SMALL;
        for (int ks = 0; ks < ksteps; ++ks) {
          if (ks == g) { gdn_step(lt - 1, 0); gdn_step(lt - 1, 1); }
          if (ks == 0) {
            mbar_wait(acc_empty(buf), (((uint32_t)lt >> 1) & 1u) ^ 1u, 2);
            tc_fence_after();
          }
          mbar_wait(full_bar(stage), phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)stage * PAIR_STAGE_BYTES;
          const uint64_t a_hi = smem_desc(sa), a_lo = smem_desc(sa + A_TILE_BYTES);
          const uint64_t b1 = smem_desc(sa + 2 * A_TILE_BYTES), b2 = smem_desc(sa + 2 * A_TILE_BYTES + PAIR_B1_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t o = (uint64_t)(k * 2);
            mma_ss_pair(d_main, a_hi + o, b1 + o, idesc256, (ks == 0 && k == 0) ? 0u : 1u);
            mma_ss_pair(d_small, a_lo + o, b2 + o, idesc128, 1u);
          }
          tc_commit_pair(empty_bar(stage));
          if (ks == ksteps - 1) tc_commit_pair(acc_full(buf));
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
      if (p.gdn && lt > 0) { gdn_step(lt - 1, 0); gdn_step(lt - 1, 1); }
    }
  } else {
    // ===================== epilogue (both CTAs; the single-CTA non-GDN path) =====================
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int xi = row & (p.bw - 1), yi = (row >> p.lbw) & (p.bh - 1), bi = row >> (p.lbw + p.lbh);
    const bool is_issuer = warp == 2 && elect_one();
    const uint32_t row_off = stg_base + (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    const float act_slope = p.act == HESIC_ACT_RELU ? 0.f : (p.act == HESIC_ACT_LEAKY_RELU ? 0.01f : 1.f);
    const uint32_t acc_empty_leader0 = mapa_cta(acc_empty(0), 0), acc_empty_leader1 = mapa_cta(acc_empty(1), 0);
    const uint32_t x2_full_leader = mapa_cta(x2_full, 0);
    int lt = 0;
    for (int task = first; task < n_pairs; task += step, ++lt) {
      const TaskCoord tk = decode_pair_task(p, task, (int)rank);
      const int buf = lt & 1;
      const int tb = tk.mt / txy, rr = tk.mt - tb * txy;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      const int ry = tk.ph / p.os, rx = tk.ph - ry * p.os;
      const int n0 = tk.nt * 128;
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)buf * ACC_STRIDE;
      const int c_fold = rx * p.out_Cs, mx = tx * p.bw, my = ty * p.bh, mb = tb * p.bb;
      // only the GroupNorm statistics need to know which of the tile's pixels exist (the TMA stores clip by themselves)
      const bool valid = (tx * p.bw + xi) * p.os + rx < p.Wout && (ty * p.bh + yi) * p.os + ry < p.Hout && tb * p.bb + bi < p.B;

      epi_bar();
      const int ci = (int)threadIdx.x - 64;
      if (ci < 128) {
        st_shared_f32(bias_s + 4u * ci, (n0 + ci < p.Cout) ? __ldg(p.bias + n0 + ci) : 0.f);
        if (p.gdn) st_shared_f32(beta_s + 4u * ci, __ldg(p.beta + ci));
      }
      epi_bar();

      // chunk pair i (chunks 2i, 2i+1: one per group) -> swizzled staging tiles -> TMA store (as in conv_tc_kernel)
      auto emit = [&](const float (&v)[32], int i, bool active) {
        if (is_issuer) bulk_wait_read<0>();
        epi_bar();
        if (p.out_fmt == HESIC_FMT_NHWC_F32) {
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
            tma_store_5d(&map_y0, stg_base, c_fold + nb0, mx, ry, my, mb);
            if (nb0 + 32 < p.Cout) tma_store_5d(&map_y0, stg_base + A_TILE_BYTES, c_fold + nb0 + 32, mx, ry, my, mb);
            bulk_commit();
          }
        } else {
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
            tma_store_5d(&map_y0, stg_base, c0, mx, ry, my, mb);
            tma_store_5d(&map_y1, stg_base + A_TILE_BYTES, c0, mx, ry, my, mb);
            bulk_commit();
          }
        }
      };

      mbar_wait(acc_full(buf), ((uint32_t)lt >> 1) & 1u, 7);
      tc_fence_after();
      const uint32_t acc_empty_leader = buf ? acc_empty_leader1 : acc_empty_leader0;
      if (!p.gdn) {
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          if (n0 + 2 * i * 32 >= p.Cout) break;
          const int ch = 2 * i + grp;
          const bool active = n0 + ch * 32 < p.Cout;
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
            if (p.stats) gn_accumulate(p, v, valid, tb * p.bb + bi, n0 + ch * 32);
          }
          emit(v, i, active);
        }
        tc_fence_before();
        // both CTAs' epilogues release the accumulator buffer on the LEADER's barrier (the only MMA issuer)
        mbar_arrive_remote(acc_empty_leader);
      } else {
        // fused GDN / IGDN, as in conv_tc_kernel; the x^2-ready signal of both CTAs goes to the leader's barrier
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
        mbar_arrive_remote(x2_full_leader);
        mbar_wait(norm_full, (uint32_t)lt & 1u, 8);
        tc_fence_after();
        const bool igdn = p.gdn == 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ch = 2 * i + grp;
          uint32_t q[32];
          tmem_ld32(acc + COL_SMALL + ch * 32, q);
          tmem_ld_wait();
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
        mbar_arrive_remote(acc_empty_leader);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = xs[i * 32 + j];
          emit(v, i, true);
        }
      }
    }
    if (is_issuer) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace hesic
