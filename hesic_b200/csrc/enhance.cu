// Enhancement network (Independent_EN, ywz/mywork/newnet1.py:272-311,1278-1300): conv3x3 s1 p1 over 32-channel
// full-resolution activations, with the LeakyReLU / identity additions of ResidualBlock
// (compressai/layers/layers.py:125-147) and Enhancement_Block fused into the epilogue.
//
// 36 of the 40 convolutions of Independent_EN are 32 -> 32 at 512x512: 4.8 GF and 2 x 128 B/pixel of HBM traffic
// each, i.e. 36 FLOP/B -- the layer is bound by HBM and by shared-memory fill, not by the tensor pipe, so the
// kernel is built around moving every activation byte ONCE per layer:
//   * activations live in HBM as NHWC_HILO: per pixel 32 bf16 'hi' then 32 bf16 'lo' = 128 B = one 128B-swizzled
//     shared-memory row.  A TMA box of pixels is therefore directly the K-major A operand  A = [Ah | Al]  (K = 64).
//   * one output tile = 16 x 8 pixels (UMMA M = 128).  The producer loads the tile with a one-row halo THREE times,
//     shifted by dx = -1, 0, +1 pixels (TMA zero fill = the zero padding): 3 x 20 KB instead of 9 tap loads.  Inside
//     a copy the three dy taps are plain address offsets of one 16-pixel row (2 KB, swizzle-atom aligned).
//   * the weights of a tap are ONE resident tile of 2N rows x 128 B:  rows [0, N) = [Wh | 0],  rows [N, 2N) =
//     [Wl | Wh].  One K = 64 MMA chain per tap then yields, with fp32 parity as in conv_tc.cu (bf16x3, main + small
//     accumulators in adjacent TMEM columns):
//         D[:, 0:N]  (main)  += [Ah | Al] . [Wh | 0]  = Ah.Wh
//         D[:, N:2N] (small) += [Ah | Al] . [Wl | Wh] = Ah.Wl + Al.Wh
//     (K chunks 0,1 = the Ah half: two M128 x N(2N) x K16 MMAs; chunks 2,3 = the Al half, where the main rows are zero:
//     two N-wide MMAs on the small rows only.)  36 MMAs per tile, every A row read from shared memory once per tap --
//     the kernel is bound by shared-memory bandwidth (r01 profile: tensor-core operand reads alone are 53 % of it;
//     a first version with separate main / small chains re-read A 1.5x and was 30 % slower); all nine tap tiles
//     (72 KB) stay in shared memory.
//   * warp-specialised persistent CTA: warp 0 TMA producer, warp 1 MMA issuer (4 TMEM accumulator buffers),
//     warps 2-5 and 6-9 two epilogue groups taking alternate tiles: bias, LeakyReLU, up to two residuals read
//     straight from HBM (L2-prefetched one turn ahead, loaded before the accumulator is waited for), hi/lo split,
//     one swizzled staging tile per group, one TMA store per tile.
//   * Cout <= 4 (the RGB output layer): same main loop with N = 16 (2N = 32) and a planar NCHW fp32 epilogue that adds the
//     NCHW identity image.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tc_ptx.cuh"

struct hesic_en_conv {
  int Cin = 0, Cout = 0, N = 0;
  __nv_bfloat16 *w = nullptr;   // [9][2N][64]: rows [0,N) = (hi[ci 0..31] | 0), rows [N,2N) = (lo[ci] | hi[ci]) of tap t
  float *bias = nullptr;        // [32]
  bool loaded = false;
  int device = -1;              // device that owns w / bias
  CUtensorMap map_w;
};

namespace hesic {
namespace en {
using namespace tc;

constexpr int TW = 16, TH = 8, HR = TH + 2;
constexpr int PIX_BYTES = 128;
constexpr int ROW_BYTES = TW * PIX_BYTES;           // 2 KB: one tile row = two swizzle atoms
constexpr int SLOT_BYTES = HR * ROW_BYTES;          // 20 KB: one dx-shifted copy of the tile + halo rows
constexpr int NSLOTS = 6;
constexpr int W_BYTES = 9 * 64 * 128;               // resident weights (2N = 64 rows per tap)
constexpr int STG_BYTES = BM * PIX_BYTES;
constexpr int NT = 320;                             // TMA warp + MMA warp + 2 x 4 epilogue warps
constexpr int NACC = 4;
constexpr uint32_t TMEM_COLS_EN = 256;              // 4 x {main 32, small 32}
constexpr int SMEM_BYTES = 1024 + W_BYTES + NSLOTS * SLOT_BYTES + 2 * STG_BYTES + 512;

struct EParams {
  int H, W, B, tiles_x, tiles_y, n_tasks;
  int N, Cout, act, planar;
  const float *bias;
  const __nv_bfloat16 *res1, *res2;   // NHWC_HILO residuals (32 slots) or null
  const float *ident;                 // planar: NCHW fp32 image added to the output, or null
  int ident_Cs;
  float *y;                           // planar output
  int y_Cs;
};

__device__ __forceinline__ void epi_bar128(int grp) {
  if (grp == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
  else asm volatile("bar.sync 4, 128;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// v[8g .. 8g+7] += hi + lo of one 8-channel chunk pair
__device__ __forceinline__ void add_chunk(float *v, const uint4 &h, const uint4 &l) {
  v[0] += bf_lo(h.x) + bf_lo(l.x); v[1] += bf_hi(h.x) + bf_hi(l.x);
  v[2] += bf_lo(h.y) + bf_lo(l.y); v[3] += bf_hi(h.y) + bf_hi(l.y);
  v[4] += bf_lo(h.z) + bf_lo(l.z); v[5] += bf_hi(h.z) + bf_hi(l.z);
  v[6] += bf_lo(h.w) + bf_lo(l.w); v[7] += bf_hi(h.w) + bf_hi(l.w);
}

// (A CTA-pair variant that shared the weight tiles between two SMs was built and measured in r01 -- 381-388 us per layer
// against 340 us: only the weight reads shrink, 9 % of the shared-memory traffic, while two SMs become lock-stepped --
// and removed in r02; profiles/r01_en_pair_layer_times.txt keeps the measurement.)
__global__ void __launch_bounds__(NT, 1)
en_conv_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_r1,
               const __grid_constant__ CUtensorMap map_r2, const __grid_constant__ EParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t w_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slot_base = w_base + (uint32_t)W_BYTES;
  const uint32_t stg = slot_base + (uint32_t)(NSLOTS * SLOT_BYTES);
  const uint32_t bar_base = stg + 2u * (uint32_t)STG_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto acc_full = [&](int b) { return bar_base + 128u + 8u * b; };
  auto acc_empty = [&](int b) { return bar_base + 160u + 8u * b; };
  const uint32_t w_full = bar_base + 192u, tmem_slot = bar_base + 200u;
  auto res_full_bar = [&](int g) { return bar_base + 208u + 8u * g; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wt_rows = 2u * (uint32_t)p.N;   // rows of one resident tap tile
  const uint32_t wt_bytes = wt_rows * 128u;
  // work units = tiles, dealt round-robin to the persistent CTAs
  const int n_units = p.n_tasks;
  const int ufirst = (int)blockIdx.x, ustep = (int)gridDim.x;
  auto tile_of = [&](int u) { return u; };

  if (warp == 0 && lane == 0) {
    prefetch_map(&map_x); prefetch_map(&map_w);
    if (!p.planar) prefetch_map(&map_y);
    if (p.res1) prefetch_map(&map_r1);
    if (p.res2) prefetch_map(&map_r2);
    for (int s = 0; s < NSLOTS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < NACC; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 128); }
    mbar_init(w_full, 1);
    mbar_init(res_full_bar(0), 1); mbar_init(res_full_bar(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS_EN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
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
      mbar_expect_tx(w_full, 9u * wt_bytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(&map_w, w_base + (uint32_t)t * wt_bytes, w_full, 0, t * 2 * p.N);
      int slot = 0;
      uint32_t phase = 0;
      for (int u = ufirst; u < n_units; u += ustep) {
        const int task = tile_of(u);
        const int tb = task / txy, rr = task - tb * txy;
        const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
        for (int c = 0; c < 3; ++c) {
          mbar_wait(empty_bar(slot), phase ^ 1u, 1);
          const uint32_t dst = slot_base + (uint32_t)slot * SLOT_BYTES;
          mbar_expect_tx(full_bar(slot), (uint32_t)SLOT_BYTES);
          tma_load_4d(&map_x, dst, full_bar(slot), 0, tx * TW + c - 1, ty * TH - 1, tb);
          if (++slot == NSLOTS) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = instr_desc(2 * p.N), idesc_n = instr_desc(p.N);
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) { mma_ss(d, a, b, id, acc); };
      auto commit = [&](uint32_t bar) { tc_commit(bar); };
      mbar_wait(w_full, 0, 5);
      tc_fence_after();
      int slot = 0, lt = 0;
      uint32_t phase = 0;
      for (int u = ufirst; u < n_units; u += ustep, ++lt) {
        const int buf = lt & (NACC - 1);
        mbar_wait(acc_empty(buf), (((uint32_t)lt >> 2) & 1u) ^ 1u, 2);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)buf * 64u;   // columns [0, N) main, [N, 2N) small
        for (int c = 0; c < 3; ++c) {
          mbar_wait(full_bar(slot), phase, 3);
          tc_fence_after();
          const uint32_t sa = slot_base + (uint32_t)slot * SLOT_BYTES;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t a = smem_desc(sa + (uint32_t)dy * ROW_BYTES);
            const uint64_t b = smem_desc(w_base + (uint32_t)(dy * 3 + c) * wt_bytes);
            const uint32_t first = (c == 0 && dy == 0) ? 0u : 1u;
            // K chunks 0,1 (A = Ah): all 2N rows; chunks 2,3 (A = Al): the main rows are zero there -> small rows only
            // (single CTA: rows N..2N of the tile; pair: the N/2 extra small rows each CTA keeps behind its first N rows)
            mma(d, a, b, idesc, first);
            mma(d, a + 2, b + 2, idesc, 1u);
            const uint64_t bs = smem_desc(w_base + (uint32_t)(dy * 3 + c) * wt_bytes + (uint32_t)p.N * 128u);
            mma(d + (uint32_t)p.N, a + 4, bs + 4, idesc_n, 1u);
            mma(d + (uint32_t)p.N, a + 6, bs + 6, idesc_n, 1u);
          }
          commit(empty_bar(slot));
          if (++slot == NSLOTS) { slot = 0; phase ^= 1u; }
        }
        commit(acc_full(buf));
      }
    }
  } else {
    // ===================== epilogue (TMEM lane quadrant = warp % 4; group g = tiles lt % 2 == g) =====================
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int xi = row & (TW - 1), yi = row >> 4;
    const bool is_issuer = (warp & 3) == 2 && elect_one();
    const uint32_t my_stg = stg + (uint32_t)grp * STG_BYTES;
    const uint32_t row_off = my_stg + (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    const uint32_t res_full = res_full_bar(grp);
    uint32_t res_phase = 0;
    float bias[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias[j] = j < p.Cout ? __ldg(p.bias + j) : 0.f;
    auto release_acc = [&](int b) { mbar_arrive(acc_empty(b)); };
    for (int lt = grp, u = ufirst + grp * ustep; u < n_units; u += 2 * ustep, lt += 2) {
      const int task = tile_of(u);
      const int buf = lt & (NACC - 1);
      const int tb = task / txy, rr = task - tb * txy;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      const int ox = tx * TW + xi, oy = ty * TH + yi;
      const bool valid = ox < p.W && oy < p.H && tb < p.B;
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)buf * 64u;
      if (p.planar) {
        float idv[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.ident && valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < p.Cout) idv[c] = __ldg(p.ident + (((size_t)tb * p.ident_Cs + c) * p.H + oy) * p.W + ox);
        }
        mbar_wait(acc_full(buf), ((uint32_t)lt >> 2) & 1u, 7);
        tc_fence_after();
        uint32_t r[16], q[16];
        tmem_ld16(acc, r);
        tmem_ld16(acc + 16u, q);
        tmem_ld_wait();
        tc_fence_before();
        release_acc(buf);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < p.Cout) {
              const float v = apply_act((__uint_as_float(r[c]) + __uint_as_float(q[c])) + bias[c], p.act) + idv[c];
              p.y[(((size_t)tb * p.y_Cs + c) * p.H + oy) * p.W + ox] = v;
            }
        }
        continue;
      }
      // Residual tiles land in the group's staging buffer by TMA ([128 px][128 B], swizzled like the output tile):
      // a thread reads its own row, adds, and later writes its output row to the same place.  (Per-thread 128-byte
      // global loads cost 32 L1 wavefronts per instruction -- r01 profile: +55 / +230 us per layer for one / two
      // residuals.)  The next tile's residual tiles are prefetched into L2 meanwhile.
      const bool has1 = p.res1 != nullptr, has2 = p.res2 != nullptr;
      if (is_issuer) {
        bulk_wait_read<0>();   // the group's previous output tile has left the staging buffer
        if (has1) {
          mbar_expect_tx(res_full, (uint32_t)STG_BYTES);
          tma_load_4d(&map_r1, my_stg, res_full, 0, tx * TW, ty * TH, tb);
          const int nt = tile_of(u + 2 * ustep);
          if (u + 2 * ustep < n_units && nt < p.n_tasks) {
            const int nb = nt / txy, nr = nt - nb * txy;
            tma_prefetch_4d(&map_r1, 0, (nr % p.tiles_x) * TW, (nr / p.tiles_x) * TH, nb);
            if (has2) tma_prefetch_4d(&map_r2, 0, (nr % p.tiles_x) * TW, (nr / p.tiles_x) * TH, nb);
          }
        }
      }
      mbar_wait(acc_full(buf), ((uint32_t)lt >> 2) & 1u, 7);
      tc_fence_after();
      float v[32];
      {
        uint32_t r[32], q[32];
        tmem_ld32(acc, r);
        tmem_ld32(acc + 32u, q);
        tmem_ld_wait();
        tc_fence_before();
        release_acc(buf);
        const float slope = p.act == HESIC_ACT_RELU ? 0.f : (p.act == HESIC_ACT_LEAKY_RELU ? 0.01f : 1.f);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float t = (__uint_as_float(r[j]) + __uint_as_float(q[j])) + bias[j];
          v[j] = fmaxf(t, t * slope);
        }
      }
      if (has1) {
        mbar_wait(res_full, res_phase, 10);
        res_phase ^= 1u;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 h = ld_shared_u4(row_off + (((uint32_t)g ^ sw) << 4)), l = ld_shared_u4(row_off + (((uint32_t)(g + 4) ^ sw) << 4));
          add_chunk(v + 8 * g, h, l);
        }
        if (has2) {
          fence_async_smem();
          epi_bar128(grp);   // every row of the first residual tile has been read
          if (is_issuer) {
            mbar_expect_tx(res_full, (uint32_t)STG_BYTES);
            tma_load_4d(&map_r2, my_stg, res_full, 0, tx * TW, ty * TH, tb);
          }
          mbar_wait(res_full, res_phase, 11);
          res_phase ^= 1u;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 h = ld_shared_u4(row_off + (((uint32_t)g ^ sw) << 4)), l = ld_shared_u4(row_off + (((uint32_t)(g + 4) ^ sw) << 4));
            add_chunk(v + 8 * g, h, l);
          }
        }
      } else {
        epi_bar128(grp);   // staging buffer free (issuer's bulk_wait_read above)
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        split_pair(v[g * 8 + 0], v[g * 8 + 1], h0, l0);
        split_pair(v[g * 8 + 2], v[g * 8 + 3], h1, l1);
        split_pair(v[g * 8 + 4], v[g * 8 + 5], h2, l2);
        split_pair(v[g * 8 + 6], v[g * 8 + 7], h3, l3);
        st_shared_v4(row_off + (((uint32_t)g ^ sw) << 4), h0, h1, h2, h3);
        st_shared_v4(row_off + (((uint32_t)(g + 4) ^ sw) << 4), l0, l1, l2, l3);
      }
      fence_async_smem();
      epi_bar128(grp);
      if (is_issuer) {
        tma_store_4d(&map_y, my_stg, 0, tx * TW, ty * TH, tb);
        bulk_commit();
      }
    }
    if (is_issuer) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS_EN) : "memory");
  }
}

// weight [Cout][Cin][3][3] fp32 -> [9][2N][64] bf16: rows [0,N) = (hi | 0), rows [N,2N) = (lo | hi), zero padded
__global__ void en_pack_w_kernel(const float *__restrict__ w, int Cin, int Cout, int N, __nv_bfloat16 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * N * 32) return;
  const int ci = i & 31, co = (i >> 5) % N, t = i / (32 * N);
  float v = 0.f;
  if (ci < Cin && co < Cout) v = w[((size_t)co * Cin + ci) * 9 + t];
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  __nv_bfloat16 *m = out + ((size_t)t * 2 * N + co) * 64, *sm = m + (size_t)N * 64;
  m[ci] = hi;
  m[32 + ci] = __float2bfloat16_rn(0.f);
  sm[ci] = lo;
  sm[32 + ci] = hi;
}

// cat(xa, xb) NCHW fp32 -> NHWC_HILO (32 slots): thread = pixel, 128 B written per pixel
__global__ void __launch_bounds__(256) en_pack_input_kernel(const TView xa, const TView xb, __nv_bfloat16 *__restrict__ y,
                                                           size_t npix) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int x = i % xa.W;
  const size_t r = i / xa.W;
  const int yy = r % xa.H, b = r / xa.H;
  const int Ca = xa.C, Cb = xb.p0 ? xb.C : 0;
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 2 * j + e;
      v[e] = 0.f;
      if (c < Ca) v[e] = ((const float *)xa.p0)[(((size_t)b * xa.Cs + c) * xa.H + yy) * xa.W + x];
      else if (c < Ca + Cb) v[e] = ((const float *)xb.p0)[(((size_t)b * xb.Cs + (c - Ca)) * xb.H + yy) * xb.W + x];
    }
    split_pair(v[0], v[1], hi[j], lo[j]);
  }
  uint4 *o = reinterpret_cast<uint4 *>(y) + i * 8;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    o[g] = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
    o[g + 4] = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
  }
}

// watchdog of this translation unit (hesic_tc_status, conv_tc.cu)
int watchdog_status(unsigned int *dbg) {
  unsigned int flag = 0;
  if (cudaMemcpyFromSymbol(&flag, g_tc_abort, sizeof(flag)) != cudaSuccess) return -1;
  if (!flag) return 0;
  cudaMemcpyFromSymbol(dbg, g_tc_dbg, 8 * sizeof(unsigned int));
  unsigned int zero = 0;
  cudaMemcpyToSymbol(g_tc_abort, &zero, sizeof(zero));
  return 1;
}

}  // namespace en
}  // namespace hesic

using namespace hesic;

extern "C" hesic_en_conv *hesic_en_conv_create(int Cin, int Cout) {
  if (Cin <= 0 || Cin > 32 || !(Cout == 32 || (Cout >= 1 && Cout <= 4))) {
    set_error("hesic_en_conv_create: Cin must be <= 32 and Cout 32 or <= 4 (got %d -> %d)", Cin, Cout);
    return nullptr;
  }
  hesic_en_conv *c = new hesic_en_conv();
  c->Cin = Cin; c->Cout = Cout; c->N = Cout == 32 ? 32 : 16;
  return c;
}

extern "C" void hesic_en_conv_destroy(hesic_en_conv *c) {
  if (!c) return;
  cudaFree(c->w); cudaFree(c->bias);
  delete c;
}

extern "C" int hesic_en_conv_load(hesic_en_conv *c, const float *weight, const float *bias, void *stream) {
  HESIC_REQUIRE(c && weight, "hesic_en_conv_load: null argument");
  cudaStream_t s = as_stream(stream);
  if (c->device >= 0 && c->device != current_device()) {     // the model moved to another device: re-pack there
    cudaFree(c->w); cudaFree(c->bias);
    c->w = nullptr; c->bias = nullptr; c->loaded = false;
  }
  c->device = current_device();
  if (!c->w) {
    HESIC_CUDA(cudaMalloc(&c->w, (size_t)9 * 2 * c->N * 64 * sizeof(__nv_bfloat16)));
    HESIC_CUDA(cudaMalloc(&c->bias, 32 * sizeof(float)));
    const uint64_t dims[2] = {64, (uint64_t)9 * 2 * c->N}, strides[1] = {128};
    const uint32_t box[2] = {64, (uint32_t)(2 * c->N)};
    int r = tc::make_tensor_map(&c->map_w, c->w, 2, dims, strides, box);
    if (r != HESIC_OK) return r;
  }
  en::en_pack_w_kernel<<<(9 * c->N * 32 + 255) / 256, 256, 0, s>>>(weight, c->Cin, c->Cout, c->N, c->w);
  HESIC_LAUNCHED("en_pack_w_kernel");
  HESIC_CUDA(cudaMemsetAsync(c->bias, 0, 32 * sizeof(float), s));
  if (bias) HESIC_CUDA(cudaMemcpyAsync(c->bias, bias, c->Cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
  c->loaded = true;
  return HESIC_OK;
}

static int check_hilo32(const hesic_tensor *t, const char *name) {
  int r = check_tensor(t, name);
  if (r != HESIC_OK) return r;
  HESIC_REQUIRE(t->fmt == HESIC_FMT_NHWC_HILO && (t->Cs == 32 || (t->Cs == 0 && t->C == 32)),
                "%s: expected an NHWC_HILO tensor with 32 channel slots", name);
  HESIC_REQUIRE(((uintptr_t)t->p0 & 127) == 0, "%s: data must be 128-byte aligned", name);
  return HESIC_OK;
}

extern "C" int hesic_en_conv_forward(hesic_en_conv *c, const hesic_tensor *x, const hesic_tensor *y, int act,
                                     const hesic_tensor *res1, const hesic_tensor *res2, void *stream) {
  using namespace en;
  HESIC_REQUIRE(c && c->loaded, "hesic_en_conv_forward: weights not loaded");
  HESIC_REQUIRE(c->device == current_device(), "hesic_en_conv_forward: the layer's operands live on device %d, the current device is %d",
                c->device, current_device());
  int r;
  if ((r = check_hilo32(x, "en conv input")) != HESIC_OK) return r;
  HESIC_REQUIRE(x->C >= c->Cin, "en conv: input has %d channels, layer expects %d", x->C, c->Cin);
  HESIC_REQUIRE(act >= 0 && act <= 2, "en conv: bad activation %d", act);
  const bool planar = c->Cout <= 4;
  if (planar) {
    if ((r = check_tensor(y, "en conv output")) != HESIC_OK) return r;
    HESIC_REQUIRE(y->fmt == HESIC_FMT_NCHW_F32 && y->C == c->Cout, "en conv: a %d-channel layer writes NCHW fp32", c->Cout);
    if (res1) {
      if ((r = check_tensor(res1, "en conv identity")) != HESIC_OK) return r;
      HESIC_REQUIRE(res1->fmt == HESIC_FMT_NCHW_F32 && same_shape(res1, y), "en conv: identity must be NCHW fp32 like the output");
    }
    HESIC_REQUIRE(res2 == nullptr, "en conv: the planar output takes one identity tensor");
  } else {
    if ((r = check_hilo32(y, "en conv output")) != HESIC_OK) return r;
    HESIC_REQUIRE(y->C == 32, "en conv: output must have 32 channels");
    for (const hesic_tensor *t : {res1, res2}) {
      if (!t) continue;
      if ((r = check_hilo32(t, "en conv residual")) != HESIC_OK) return r;
      HESIC_REQUIRE(t->B == y->B && t->H == y->H && t->W == y->W, "en conv: residual shape mismatch");
    }
  }
  HESIC_REQUIRE(y->B == x->B && y->H == x->H && y->W == x->W, "en conv: output is %dx%dx%d, expected %dx%dx%d", y->B, y->H,
                y->W, x->B, x->H, x->W);
  if ((int64_t)x->B * x->H * x->W == 0) return HESIC_OK;

  static int num_sms = 0;
  static PerDeviceOnce once;
  if (once.first()) {
    int dev = 0;
    HESIC_CUDA(cudaGetDevice(&dev));
    HESIC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    HESIC_CUDA(cudaFuncSetAttribute(en_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  EParams p;
  memset(&p, 0, sizeof(p));
  p.H = x->H; p.W = x->W; p.B = x->B;
  p.tiles_x = (x->W + TW - 1) / TW; p.tiles_y = (x->H + TH - 1) / TH;
  p.n_tasks = p.tiles_x * p.tiles_y * x->B;
  p.N = c->N; p.Cout = c->Cout; p.act = act; p.planar = planar ? 1 : 0;
  p.bias = c->bias;
  if (planar) {
    p.ident = res1 ? (const float *)res1->p0 : nullptr;
    p.ident_Cs = res1 ? (res1->Cs > 0 ? res1->Cs : res1->C) : 0;
    p.y = (float *)y->p0; p.y_Cs = y->Cs > 0 ? y->Cs : y->C;
  } else {
    p.res1 = res1 ? (const __nv_bfloat16 *)res1->p0 : nullptr;
    p.res2 = res2 ? (const __nv_bfloat16 *)res2->p0 : nullptr;
  }
  CUtensorMap mx, my;
  const uint64_t dims[4] = {64, (uint64_t)x->W, (uint64_t)x->H, (uint64_t)x->B};
  const uint64_t strides[3] = {128, (uint64_t)x->W * 128, (uint64_t)x->H * x->W * 128};
  const uint32_t box_in[4] = {64, TW, HR, 1}, box_out[4] = {64, TW, TH, 1};
  if ((r = tc::make_tensor_map(&mx, x->p0, 4, dims, strides, box_in)) != HESIC_OK) return r;
  if (planar) my = mx;
  else if ((r = tc::make_tensor_map(&my, y->p0, 4, dims, strides, box_out)) != HESIC_OK) return r;
  CUtensorMap mr1 = mx, mr2 = mx;
  if (!planar && res1 && (r = tc::make_tensor_map(&mr1, res1->p0, 4, dims, strides, box_out)) != HESIC_OK) return r;
  if (!planar && res2 && (r = tc::make_tensor_map(&mr2, res2->p0, 4, dims, strides, box_out)) != HESIC_OK) return r;
  if (!planar && !res1 && res2) { mr1 = mr2; p.res1 = p.res2; p.res2 = nullptr; }
  const int grid = std::min(p.n_tasks, num_sms);
  en_conv_kernel<<<grid, NT, SMEM_BYTES, as_stream(stream)>>>(mx, c->map_w, my, mr1, mr2, p);
  HESIC_LAUNCHED("en_conv_kernel");
  return HESIC_OK;
}

extern "C" int hesic_en_pack_input(const hesic_tensor *xa, const hesic_tensor *xb, const hesic_tensor *y, void *stream) {
  int r;
  if ((r = check_tensor(xa, "en pack input")) != HESIC_OK) return r;
  if (xb && (r = check_tensor(xb, "en pack input (second part)")) != HESIC_OK) return r;
  if ((r = check_hilo32(y, "en pack output")) != HESIC_OK) return r;
  HESIC_REQUIRE(xa->fmt == HESIC_FMT_NCHW_F32 && (!xb || xb->fmt == HESIC_FMT_NCHW_F32), "en pack: inputs must be NCHW fp32");
  HESIC_REQUIRE(xa->B == y->B && xa->H == y->H && xa->W == y->W, "en pack: shape mismatch");
  HESIC_REQUIRE(!xb || (xb->B == xa->B && xb->H == xa->H && xb->W == xa->W), "en pack: the two inputs differ in shape");
  HESIC_REQUIRE(xa->C + (xb ? xb->C : 0) <= 32, "en pack: more than 32 channels");
  const size_t npix = (size_t)y->B * y->H * y->W;
  if (npix == 0) return HESIC_OK;
  TView vb;
  memset(&vb, 0, sizeof(vb));
  if (xb) vb = view(xb);
  en::en_pack_input_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, as_stream(stream)>>>(view(xa), vb, (__nv_bfloat16 *)y->p0, npix);
  HESIC_LAUNCHED("en_pack_input_kernel");
  return HESIC_OK;
}
