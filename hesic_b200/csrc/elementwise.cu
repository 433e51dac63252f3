// HBM-bound kernels of the HESIC forward path: homography warp, entropy-model quantise/likelihood,
// global max + mixture softmax, bilinear upsample, layout conversion, symbol/index preparation and
// the rate-distortion partial sums.  All sm_100a, all asynchronous on the caller's stream.
#include <limits.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace hesic {

__device__ __forceinline__ void unflatten(const TView &t, size_t i, int &b, int &c, int &y, int &x) {
  // walk in the tensor's own memory order so that consecutive threads touch consecutive addresses
  if (t.fmt == HESIC_FMT_NCHW_F32) {
    x = i % t.W; i /= t.W; y = i % t.H; i /= t.H; c = i % t.C; b = i / t.C;
  } else {
    c = i % t.C; i /= t.C; x = i % t.W; i /= t.W; y = i % t.H; b = i / t.H;
  }
}

__device__ __forceinline__ void block_add_double(double v, double *acc) {
  v = warp_sum(v);
  __shared__ double part[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) part[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    double s = lane < nw ? part[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(acc, s);
  }
}

// ---------------------------------------------------------------------------------------------
// kornia.warp_perspective.  Block = 32 x 8 destination pixels of one image, thread = one pixel, all channels.
//  * The 3x3 chain N_dst * M * N_src^-1 and its inverse are evaluated once per block in fp64; the per-pixel
//    sampling coordinate (normalised grid -> S*g -> divide -> grid_sample unnormalise) is fp64 as well (at
//    512x512 the fp32 chain loses ~2e-4 px, visible at the 1e-4 level on textured images) but costs ~35 DP
//    instructions: FMA form, and 1/z by two Newton steps on a MUFU seed instead of an IEEE division.
//  * Gather staged through shared memory: every thread packs its top-left source tap (x0, y0) into one word,
//    warp shuffles (redux) reduce the packed words to the warp's source bounding box, the 8 warps merge theirs
//    with shared-memory atomics, and the block then copies that source window (all channels) row by row with
//    coalesced loads; the 4 x C bilinear taps read shared memory.  Windows larger than the staging buffer
//    (strong perspective / minification) fall back to gathering straight from global memory.
//  * The blend itself is fp32 in grid_sample's order.  Optional second output in ROWPAD format: the warped
//    image is the input of the next 3->128 tensor-core layer (newnet1.py:753-754), which saves a repack pass.
constexpr int WARP_STAGE_FLOATS = 6144;   // 24 KB: e.g. 3 channels x 40 x 51 source pixels
constexpr int WARP_TILE = 32;   // block = 32 x (8 * WARP_ROWS) destination pixels, WARP_ROWS rows per thread

template <int WARP_ROWS>
__global__ void __launch_bounds__(256, 3) warp_kernel(const TView src, const float *__restrict__ Mx, const TView dst,
                                                  const TView dst2, int align_corners) {
  __shared__ double T[9], S[12];
  __shared__ int bbox[4];   // min x0, min y0, max x0, max y0 over the block's finite pixels
  __shared__ float stage[WARP_STAGE_FLOATS];
  const int b = blockIdx.z;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  // S = (N_dst * M * N_src^-1)^-1 in fp64, nine threads in parallel (one matrix entry each)
  if (tid < 9) {
    const int i = tid / 3, j = tid - 3 * i;
    const float *m = Mx + b * 9;
    const double h = src.H, w = src.W, ho = dst.H, wo = dst.W;
    const double nsi[9] = {(w - 1) / 2, 0, (w - 1) / 2, 0, (h - 1) / 2, (h - 1) / 2, 0, 0, 1};
    const double nd[9] = {2 / (wo - 1), 0, -1, 0, 2 / (ho - 1), -1, 0, 0, 1};
    double t = 0;
    for (int k = 0; k < 3; ++k) {
      double a = 0;   // (M * N_src^-1)[k][j]
      for (int l = 0; l < 3; ++l) a += (double)m[k * 3 + l] * nsi[l * 3 + j];
      t += nd[i * 3 + k] * a;
    }
    T[tid] = t;
  }
  if (tid == 0) { bbox[0] = bbox[1] = INT_MAX; bbox[2] = bbox[3] = INT_MIN; }
  __syncthreads();
  if (tid < 9) {
    const int i = tid / 3, j = tid - 3 * i;
    auto cof = [&](int r, int c) {   // cofactor of T[r][c] (cyclic form, sign included)
      const int r1 = (r + 1) % 3, r2 = (r + 2) % 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      return T[r1 * 3 + c1] * T[r2 * 3 + c2] - T[r1 * 3 + c2] * T[r2 * 3 + c1];
    };
    const double det = T[0] * cof(0, 0) + T[1] * cof(0, 1) + T[2] * cof(0, 2);
    S[tid] = cof(j, i) / det;       // inverse = adjugate / det
  } else if (tid == 9) {
    S[9] = dst.W > 1 ? 2.0 / (double)(dst.W - 1) : 0.0;
  } else if (tid == 10) {
    S[10] = dst.H > 1 ? 2.0 / (double)(dst.H - 1) : 0.0;
  }
  __syncthreads();
  const int x = blockIdx.x * WARP_TILE + threadIdx.x;
  const double gx = fma((double)x, S[9], -1.0);
  const double hw = align_corners ? 0.5 * (double)(src.W - 1) : 0.5 * (double)src.W;
  const double hh = align_corners ? 0.5 * (double)(src.H - 1) : 0.5 * (double)src.H;
  int x0[WARP_ROWS], y0[WARP_ROWS];
  float ax[WARP_ROWS], ay[WARP_ROWS];
  bool fin[WARP_ROWS];
  unsigned minx = 0xffffu, miny = 0xffffu, maxx = 0u, maxy = 0u;
  const unsigned bias = 4u;   // x0, y0 >= -2 for finite pixels
#pragma unroll
  for (int k = 0; k < WARP_ROWS; ++k) {
    const int y = blockIdx.y * (8 * WARP_ROWS) + threadIdx.y + 8 * k;
    const double gy = fma((double)y, S[10], -1.0);
    double u = fma(gx, S[0], fma(gy, S[1], S[2]));
    double v = fma(gx, S[3], fma(gy, S[4], S[5]));
    const double z = fma(gx, S[6], fma(gy, S[7], S[8]));
    if (fabs(z) > 1e-8) {
      double r = (double)(1.0f / (float)z);      // 24-bit seed, two Newton steps -> full double precision
      r = r * (2.0 - z * r);
      r = r * (2.0 - z * r);
      u *= r; v *= r;
    }
    // align_corners: ((u+1)/2)*(W-1);  else ((u+1)*W - 1)/2
    const double ix = align_corners ? (u + 1.0) * hw : fma(u + 1.0, hw, -0.5);
    const double iy = align_corners ? (v + 1.0) * hh : fma(v + 1.0, hh, -0.5);
    // samples further than one pixel outside the source contribute nothing; the range test keeps the ints small
    fin[k] = x < dst.W && y < dst.H && ix > -2.0 && iy > -2.0 && ix < (double)src.W + 1.0 && iy < (double)src.H + 1.0;
    const double fx = fin[k] ? floor(ix) : 0.0, fy = fin[k] ? floor(iy) : 0.0;
    x0[k] = (int)fx; y0[k] = (int)fy;
    ax[k] = (float)(ix - fx); ay[k] = (float)(iy - fy);
    if (fin[k]) {
      minx = min(minx, (unsigned)(x0[k] + bias)); maxx = max(maxx, (unsigned)(x0[k] + bias));
      miny = min(miny, (unsigned)(y0[k] + bias)); maxy = max(maxy, (unsigned)(y0[k] + bias));
    }
  }
  // source window of the block: the (x, y) extremes are packed two to a word and reduced with warp shuffles
  // (min of the low halves via 0xffff - v), then merged across the 8 warps with shared-memory atomics
  {
    unsigned lo = (minx << 16) | miny, hi = (maxx << 16) | maxy;
    const unsigned minx_w = __reduce_min_sync(0xffffffffu, lo) >> 16, maxx_w = __reduce_max_sync(0xffffffffu, hi) >> 16;
    const unsigned miny_w = __reduce_min_sync(0xffffffffu, (miny << 16) | minx) >> 16;
    const unsigned maxy_w = __reduce_max_sync(0xffffffffu, (maxy << 16) | maxx) >> 16;
    if (threadIdx.x == 0 && maxx_w >= minx_w && minx_w != 0xffffu) {
      atomicMin(&bbox[0], (int)minx_w - (int)bias); atomicMin(&bbox[1], (int)miny_w - (int)bias);
      atomicMax(&bbox[2], (int)maxx_w - (int)bias); atomicMax(&bbox[3], (int)maxy_w - (int)bias);
    }
  }
  __syncthreads();
  // window clipped to the image: columns [wx0, wx0 + ww), rows [wy0, wy0 + wh)
  int wx0 = 0, wy0 = 0, ww = 0, wh = 0;
  if (bbox[2] >= bbox[0]) {   // at least one finite pixel
    wx0 = max(bbox[0], 0); wy0 = max(bbox[1], 0);
    ww = min(bbox[2] + 1, src.W - 1) - wx0 + 1; wh = min(bbox[3] + 1, src.H - 1) - wy0 + 1;
  }
  const bool staged = ww > 0 && wh > 0 && src.fmt == HESIC_FMT_NCHW_F32 &&
                      (size_t)ww * wh * src.C <= (size_t)WARP_STAGE_FLOATS;
  if (staged) {
    const int rows = wh * src.C;
    for (int r = threadIdx.y; r < rows; r += 8) {
      const int c = r / wh, yy = wy0 + (r - c * wh);
      const float *g = (const float *)src.p0 + (((size_t)b * src.Cs + c) * src.H + yy) * src.W + wx0;
      for (int col = threadIdx.x; col < ww; col += 32) cp_async<4>(stage + r * ww + col, g + col, true);
    }
    cp_async_wait_all();
    __syncthreads();
  }
  if (x >= dst.W) return;
  const int nc = src.C, plane = wh * ww;
  const bool nchw_out = dst.fmt == HESIC_FMT_NCHW_F32;
  const size_t dplane = (size_t)dst.H * dst.W;
  if (staged && nc == 3 && nchw_out) {
    // Fast path of the forward pass (RGB, window staged): branch-free taps.  A tap outside the image gets weight 0
    // and its address clamped into the staged window (finite data), which leaves grid_sample's result and its
    // nw + ne + sw + se summation order unchanged.
#pragma unroll
    for (int k = 0; k < WARP_ROWS; ++k) {
      const int y = blockIdx.y * (8 * WARP_ROWS) + threadIdx.y + 8 * k;
      if (y >= dst.H) break;
      const int x1 = x0[k] + 1, y1 = y0[k] + 1;
      const bool vx0 = fin[k] && x0[k] >= 0 && x0[k] < src.W, vx1 = fin[k] && x1 >= 0 && x1 < src.W;
      const bool vy0 = y0[k] >= 0 && y0[k] < src.H, vy1 = y1 >= 0 && y1 < src.H;
      const float bx0 = vx0 ? 1.f - ax[k] : 0.f, bx1 = vx1 ? ax[k] : 0.f;
      const float by0 = vy0 ? 1.f - ay[k] : 0.f, by1 = vy1 ? ay[k] : 0.f;
      // weights exactly as grid_sample forms them: (1-ax)(1-ay), ax(1-ay), (1-ax)ay, ax*ay -- or 0
      const float wnw = bx0 * by0, wne = bx1 * by0, wsw = bx0 * by1, wse = bx1 * by1;
      const int cx0 = min(max(x0[k] - wx0, 0), ww - 1), cx1 = min(max(x1 - wx0, 0), ww - 1);
      const int r0 = min(max(y0[k] - wy0, 0), wh - 1) * ww, r1 = min(max(y1 - wy0, 0), wh - 1) * ww;
      float out[3];
      float *dp = (float *)dst.p0 + ((size_t)b * dst.Cs * dst.H + y) * dst.W + x;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float *t = stage + c * plane;
        float o = t[r0 + cx0] * wnw;
        o += t[r0 + cx1] * wne;
        o += t[r1 + cx0] * wsw;
        o += t[r1 + cx1] * wse;
        out[c] = o;
        dp[c * dplane] = o;
      }
      if (dst2.p0) store_rowpad_pixel(dst2, b, y, x, out);
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < WARP_ROWS; ++k) {
    const int y = blockIdx.y * (8 * WARP_ROWS) + threadIdx.y + 8 * k;
    if (y >= dst.H) break;
    const int x1 = x0[k] + 1, y1 = y0[k] + 1;
    const float wnw = (1.f - ax[k]) * (1.f - ay[k]), wne = ax[k] * (1.f - ay[k]);
    const float wsw = (1.f - ax[k]) * ay[k], wse = ax[k] * ay[k];
    const bool vx0 = x0[k] >= 0 && x0[k] < src.W, vx1 = x1 >= 0 && x1 < src.W;
    const bool vy0 = y0[k] >= 0 && y0[k] < src.H, vy1 = y1 >= 0 && y1 < src.H;
    const bool t00 = fin[k] && vy0 && vx0, t01 = fin[k] && vy0 && vx1, t10 = fin[k] && vy1 && vx0, t11 = fin[k] && vy1 && vx1;
    const int o00 = (y0[k] - wy0) * ww + (x0[k] - wx0);   // tap offset inside the staged window
    auto sample = [&](int c) {
      float o = 0.f;
      if (staged) {
        const float *t = stage + c * plane + o00;
        if (t00) o += t[0] * wnw;
        if (t01) o += t[1] * wne;
        if (t10) o += t[ww] * wsw;
        if (t11) o += t[ww + 1] * wse;
      } else {
        if (t00) o += tload(src, b, c, y0[k], x0[k]) * wnw;
        if (t01) o += tload(src, b, c, y0[k], x1) * wne;
        if (t10) o += tload(src, b, c, y1, x0[k]) * wsw;
        if (t11) o += tload(src, b, c, y1, x1) * wse;
      }
      return o;
    };
    float *dp = nchw_out ? (float *)dst.p0 + ((size_t)b * dst.Cs * dst.H + y) * dst.W + x : nullptr;
    float out[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      out[c] = 0.f;
      if (c < nc) {
        out[c] = sample(c);
        if (nchw_out) dp[c * dplane] = out[c];
        else tstore(dst, b, c, y, x, out[c]);
      }
    }
    for (int c = 8; c < nc; ++c) tstore(dst, b, c, y, x, sample(c));
    if (dst2.p0) store_rowpad_pixel(dst2, b, y, x, out);
  }
}

// ---------------------------------------------------------------------------------------------
// Forward-path form of the warp (RGB, NCHW fp32 in and out): every WARP stages its own source window.
//
// r01 / r02 profiles of warp_kernel: 90 us for 100 MB (0.17 of the HBM roofline), DRAM 9 % busy, instruction- and
// latency-bound -- the block-wide protocol (two __syncthreads, shared-memory atomics for the bounding box, a staging
// loop of 4-byte cp.async over the whole block, a third __syncthreads) leaves the SM idle between phases.  Here a warp
// owns a strip of 32 x WR destination pixels: bounding box by warp shuffles only (redux.sync on packed words), its own
// slice of shared memory filled with 16-byte cp.async (source rows start on a 16-byte boundary: window x origin
// rounded down to 4 pixels), __syncwarp, taps from shared memory, one coalesced 128-byte store per channel row.  No
// block-level synchronisation after the prologue; the other warps' copies hide a warp's latency.  Same arithmetic as
// warp_kernel (fp64 coordinates, fp32 blend in grid_sample's order), so results are bit-identical to it.
constexpr int WRGB_ROWS = 8;                  // destination rows per warp
constexpr int WRGB_WARPS = 8;                 // warps per block: block = 32 x 64 destination pixels
constexpr int WRGB_STAGE = 1408;              // floats per warp (5.5 KB, a multiple of 128 B)
// r03: the warp's window arrives as ONE TMA box of fixed size (3 channels x 11 rows x 40 pixels, origin rounded down to 4
// pixels = 16 bytes; out-of-image elements zero-filled) instead of a loop of 16-byte cp.async whose index arithmetic (two
// integer divisions per chunk) and variable window pitch made integer instructions the bulk of the kernel (ncu source page:
// IMAD / ISETP / LEA / VIADD 42 % of the executed instructions, DFMA 4 %).
constexpr int WRGB_BW = 40, WRGB_BH = 11, WRGB_PLANE = WRGB_BW * WRGB_BH;
static_assert(3 * WRGB_PLANE <= WRGB_STAGE, "warp window does not fit the warp's staging slice");

// r04: capped at 80 registers (3 blocks = 24 warps per SM; 98 registers and 2 blocks without the cap): the kernel is bound by the
// latency of each warp's window load, which only other resident warps hide -- 54.5 -> 50.2 us; a cap of 64 (4 blocks, 60 bytes
// of spills) 51.6 us.
__global__ void __launch_bounds__(32 * WRGB_WARPS, 3) warp_rgb_kernel(const __grid_constant__ CUtensorMap map_src, const TView src,
                                                                   const float *__restrict__ Mx, const TView dst, const TView dst2,
                                                                   int align_corners) {
  __shared__ double T[9], S[12];
  __shared__ __align__(128) float stage_all[WRGB_WARPS * WRGB_STAGE];
  __shared__ __align__(8) unsigned long long wbar[WRGB_WARPS];
  const int b = blockIdx.z;
  const int lane = threadIdx.x, wid = threadIdx.y;
  const int tid = wid * 32 + lane;
  if (lane == 0) {
    tc::mbar_init(tc::smem_u32(&wbar[wid]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 9) {
    const int i = tid / 3, j = tid - 3 * i;
    const float *m = Mx + b * 9;
    const double h = src.H, w = src.W, ho = dst.H, wo = dst.W;
    const double nsi[9] = {(w - 1) / 2, 0, (w - 1) / 2, 0, (h - 1) / 2, (h - 1) / 2, 0, 0, 1};
    const double nd[9] = {2 / (wo - 1), 0, -1, 0, 2 / (ho - 1), -1, 0, 0, 1};
    double t = 0;
    for (int k = 0; k < 3; ++k) {
      double a = 0;
      for (int l = 0; l < 3; ++l) a += (double)m[k * 3 + l] * nsi[l * 3 + j];
      t += nd[i * 3 + k] * a;
    }
    T[tid] = t;
  }
  __syncthreads();
  if (tid < 9) {
    const int i = tid / 3, j = tid - 3 * i;
    auto cof = [&](int r, int c) {
      const int r1 = (r + 1) % 3, r2 = (r + 2) % 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      return T[r1 * 3 + c1] * T[r2 * 3 + c2] - T[r1 * 3 + c2] * T[r2 * 3 + c1];
    };
    const double det = T[0] * cof(0, 0) + T[1] * cof(0, 1) + T[2] * cof(0, 2);
    S[tid] = cof(j, i) / det;
  } else if (tid == 9) {
    S[9] = dst.W > 1 ? 2.0 / (double)(dst.W - 1) : 0.0;
  } else if (tid == 10) {
    S[10] = dst.H > 1 ? 2.0 / (double)(dst.H - 1) : 0.0;
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + lane;
  const int yb = (blockIdx.y * WRGB_WARPS + wid) * WRGB_ROWS;
  if (yb >= dst.H) return;                       // whole warp
  const double gx = fma((double)x, S[9], -1.0);
  const double hw = align_corners ? 0.5 * (double)(src.W - 1) : 0.5 * (double)src.W;
  const double hh = align_corners ? 0.5 * (double)(src.H - 1) : 0.5 * (double)src.H;
  int x0[WRGB_ROWS], y0[WRGB_ROWS];
  float ax[WRGB_ROWS], ay[WRGB_ROWS];
  unsigned finmask = 0;
  unsigned minx = 0xffffu, miny = 0xffffu, maxx = 0u, maxy = 0u;
  const unsigned bias = 4u;
#pragma unroll
  for (int k = 0; k < WRGB_ROWS; ++k) {
    const int y = yb + k;
    const double gy = fma((double)y, S[10], -1.0);
    double u = fma(gx, S[0], fma(gy, S[1], S[2]));
    double v = fma(gx, S[3], fma(gy, S[4], S[5]));
    const double z = fma(gx, S[6], fma(gy, S[7], S[8]));
    if (fabs(z) > 1e-8) {
      double r = (double)(1.0f / (float)z);
      r = r * (2.0 - z * r);
      r = r * (2.0 - z * r);
      u *= r; v *= r;
    }
    const double ix = align_corners ? (u + 1.0) * hw : fma(u + 1.0, hw, -0.5);
    const double iy = align_corners ? (v + 1.0) * hh : fma(v + 1.0, hh, -0.5);
    const bool fin = x < dst.W && y < dst.H && ix > -2.0 && iy > -2.0 && ix < (double)src.W + 1.0 && iy < (double)src.H + 1.0;
    const double fx = fin ? floor(ix) : 0.0, fy = fin ? floor(iy) : 0.0;
    x0[k] = (int)fx; y0[k] = (int)fy;
    ax[k] = (float)(ix - fx); ay[k] = (float)(iy - fy);
    if (fin) {
      finmask |= 1u << k;
      minx = min(minx, (unsigned)(x0[k] + bias)); maxx = max(maxx, (unsigned)(x0[k] + bias));
      miny = min(miny, (unsigned)(y0[k] + bias)); maxy = max(maxy, (unsigned)(y0[k] + bias));
    }
  }
  // the warp's source window: extremes packed two to a word, reduced with redux.sync
  const unsigned minx_w = __reduce_min_sync(0xffffffffu, (minx << 16) | miny) >> 16;
  const unsigned maxx_w = __reduce_max_sync(0xffffffffu, (maxx << 16) | maxy) >> 16;
  const unsigned miny_w = __reduce_min_sync(0xffffffffu, (miny << 16) | minx) >> 16;
  const unsigned maxy_w = __reduce_max_sync(0xffffffffu, (maxy << 16) | maxx) >> 16;
  const bool any = minx_w != 0xffffu;
  int wx0 = 0, wy0 = 0;
  bool staged = false;
  if (any) {
    wx0 = max((int)minx_w - (int)bias, 0) & ~3;
    wy0 = max((int)miny_w - (int)bias, 0);
    // columns / rows the taps can touch: up to x0 + 1 and y0 + 1, clipped to the image
    const int need_w = min((int)maxx_w - (int)bias + 1, src.W - 1) - wx0 + 1;
    const int need_h = min((int)maxy_w - (int)bias + 1, src.H - 1) - wy0 + 1;
    staged = need_w > 0 && need_h > 0 && need_w <= WRGB_BW && need_h <= WRGB_BH;
  }
  constexpr int ww = WRGB_BW, wh = WRGB_BH;
  float *stage = stage_all + wid * WRGB_STAGE;
  if (staged) {
    const uint32_t bar = tc::smem_u32(&wbar[wid]);
    if (tc::elect_one()) {
      tc::mbar_expect_tx(bar, (uint32_t)(3 * WRGB_PLANE * sizeof(float)));
      tc::tma_load_4d(&map_src, tc::smem_u32(stage), bar, wx0, wy0, 0, b);
    }
    tc::mbar_wait(bar, 0, 21);
  }
  if (x >= dst.W) return;
  constexpr int plane = WRGB_PLANE;
  const size_t dplane = (size_t)dst.H * dst.W;
  const float *gsrc = (const float *)src.p0 + (size_t)b * src.Cs * src.H * src.W;
  float prev[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < WRGB_ROWS; ++k) {
    const int y = yb + k;
    if (y >= dst.H) break;
    const bool fin = (finmask >> k) & 1u;
    const int x1 = x0[k] + 1, y1 = y0[k] + 1;
    const bool vx0 = fin && x0[k] >= 0 && x0[k] < src.W, vx1 = fin && x1 >= 0 && x1 < src.W;
    const bool vy0 = y0[k] >= 0 && y0[k] < src.H, vy1 = y1 >= 0 && y1 < src.H;
    const float bx0 = vx0 ? 1.f - ax[k] : 0.f, bx1 = vx1 ? ax[k] : 0.f;
    const float by0 = vy0 ? 1.f - ay[k] : 0.f, by1 = vy1 ? ay[k] : 0.f;
    // weights exactly as grid_sample forms them: (1-ax)(1-ay), ax(1-ay), (1-ax)ay, ax*ay -- or 0 for a tap outside
    const float wnw = bx0 * by0, wne = bx1 * by0, wsw = bx0 * by1, wse = bx1 * by1;
    float out[3];
    float *dp = (float *)dst.p0 + ((size_t)b * dst.Cs * dst.H + y) * dst.W + x;
    if (staged) {
      // a tap outside the image has weight 0 and its address clamped into the window (finite data)
      const int cx0 = min(max(x0[k] - wx0, 0), ww - 1), cx1 = min(max(x1 - wx0, 0), ww - 1);
      const int r0 = min(max(y0[k] - wy0, 0), wh - 1) * ww, r1 = min(max(y1 - wy0, 0), wh - 1) * ww;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float *t = stage + c * plane;
        float o = t[r0 + cx0] * wnw;
        o += t[r0 + cx1] * wne;
        o += t[r1 + cx0] * wsw;
        o += t[r1 + cx1] * wse;
        out[c] = o;
        dp[c * dplane] = o;
      }
    } else {
      // window too large for the warp's slice (strong perspective / minification): gather from global memory
      const int cx0 = min(max(x0[k], 0), src.W - 1), cx1 = min(max(x1, 0), src.W - 1);
      const size_t r0 = (size_t)min(max(y0[k], 0), src.H - 1) * src.W, r1 = (size_t)min(max(y1, 0), src.H - 1) * src.W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float *t = gsrc + (size_t)c * src.H * src.W;
        float o = __ldg(t + r0 + cx0) * wnw;
        o += __ldg(t + r0 + cx1) * wne;
        o += __ldg(t + r1 + cx0) * wsw;
        o += __ldg(t + r1 + cx1) * wse;
        out[c] = o;
        dp[c * dplane] = o;
      }
    }
    if (dst2.p0) {
      if (dst2.Cs == 4) {
        // rows 2j, 2j + 1 of a column share 16 contiguous bytes per plane of the pair-interleaved format: one full store
        if ((k & 1) == 0) {
          prev[0] = out[0]; prev[1] = out[1]; prev[2] = out[2];
        } else {
          __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            split_bf16(c < 3 ? prev[c < 3 ? c : 0] : 0.f, hi[c], lo[c]);
            split_bf16(c < 3 ? out[c < 3 ? c : 0] : 0.f, hi[4 + c], lo[4 + c]);
          }
          const size_t o = rowpad_off(dst2, b, y - 1, x);
          *reinterpret_cast<uint4 *>((__nv_bfloat16 *)dst2.p0 + o) = *reinterpret_cast<const uint4 *>(hi);
          *reinterpret_cast<uint4 *>((__nv_bfloat16 *)dst2.p1 + o) = *reinterpret_cast<const uint4 *>(lo);
        }
      } else {
        store_rowpad_pixel(dst2, b, y, x, out);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// EntropyBottleneck forward (eval).  params: 60 floats per channel, see hesic_b200.h.
__device__ __forceinline__ float eb_logits(const float *__restrict__ p, float v) {
  float h[3], g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t = p[j] * v + p[3 + j];
    h[j] = t + p[6 + j] * tanhf(t);
  }
  const float *q = p + 9;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float t = q[j * 3 + 0] * h[0];
      t += q[j * 3 + 1] * h[1];
      t += q[j * 3 + 2] * h[2];
      t += q[9 + j];
      g[j] = t + q[12 + j] * tanhf(t);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) h[j] = g[j];
    q += 15;
  }
  float t = q[0] * h[0];
  t += q[1] * h[1];
  t += q[2] * h[2];
  return t + q[3];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) eb_kernel(const TView z, const float *__restrict__ params, float bound,
                                                const TView z_hat, const TView lik, double *log2_sum, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  double lg = 0.0;
  if (i < n) {
    int b, c, y, x;
    unflatten(z, i, b, c, y, x);
    const float *p = params + (size_t)c * HESIC_EB_PARAMS_PER_CHANNEL;
    float med = p[58];
    float q = rintf(tload(z, b, c, y, x) - med) + med;
    float lower = eb_logits(p, q - 0.5f), upper = eb_logits(p, q + 0.5f);
    float t = lower + upper;
    float sg = t > 0.f ? -1.f : (t < 0.f ? 1.f : 0.f);
    float l = fabsf(sigmoidf_(sg * upper) - sigmoidf_(sg * lower));
    if (bound > 0.f) l = fmaxf(l, bound);
    if (z_hat.p0) tstore(z_hat, b, c, y, x, q);
    if (lik.p0) tstore(lik, b, c, y, x, l);
    lg = (double)log2f(l);
  }
  if (log2_sum) block_add_double(lg, log2_sum);
}

__global__ void eb_pack_kernel(const float *m0, const float *m1, const float *m2, const float *m3, const float *m4,
                               const float *b0, const float *b1, const float *b2, const float *b3, const float *b4,
                               const float *f0, const float *f1, const float *f2, const float *f3,
                               const float *quantiles, int C, float *out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float *p = out + (size_t)c * HESIC_EB_PARAMS_PER_CHANNEL;
  auto sp = [](float v) { return v > 20.f ? v : log1pf(expf(v)); };  // F.softplus (beta=1, threshold=20)
  for (int j = 0; j < 3; ++j) { p[j] = sp(m0[c * 3 + j]); p[3 + j] = b0[c * 3 + j]; p[6 + j] = tanhf(f0[c * 3 + j]); }
  const float *ms[3] = {m1, m2, m3}, *bs[3] = {b1, b2, b3}, *fs[3] = {f1, f2, f3};
  float *q = p + 9;
  for (int l = 0; l < 3; ++l) {
    for (int j = 0; j < 9; ++j) q[j] = sp(ms[l][c * 9 + j]);
    for (int j = 0; j < 3; ++j) { q[9 + j] = bs[l][c * 3 + j]; q[12 + j] = tanhf(fs[l][c * 3 + j]); }
    q += 15;
  }
  for (int j = 0; j < 3; ++j) q[j] = sp(m4[c * 3 + j]);
  q[3] = b4[c];
  p[58] = quantiles[c * 3 + 1];
  p[59] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// GaussianMixtureConditional / GaussianConditional forward (eval)
// Phi((0.5 - d) / s) - Phi((-0.5 - d) / s), d = |y_hat - mu| >= 0: the probability mass of one quantisation bin
// (entropy_models.py:546-554,693-702: _standardized_cumulative(u) - _standardized_cumulative(l), 0.5 erfc(-x / sqrt 2)).
// r03: the likelihood kernels were erfc-bound (0.27 of the HBM roofline: two libm erfcf and two IEEE divisions per mixture
// component, ~230 instructions).  Two cheaper forms, both MORE accurate against the exact value than the reference's own fp32
// erfc difference (<= 1.2e-5 relative down to the 1e-9 floor, the reference: 2.5e-5 at sigma < 32, 1e-4 at sigma ~ 128, where the
// two erfc values cancel), so the distance to the reference is the reference's own rounding noise (tools/erfc_eval.py):
//   * narrow bins (1 / s <= 1/8): the integral of the density over the bin as a series in h = 1 / s around its centre m = d h,
//       h phi(m) (1 + h^2 (m^2 - 1) / 24 + h^4 (m^4 - 6 m^2 + 3) / 1920)         -- no cancellation at all;
//   * otherwise 0.5 (erfc(a) - erfc(b)), a = (d - 0.5) h / sqrt 2, b = (d + 0.5) h / sqrt 2 > 0, with erfc(z >= 0) =
//     t exp(-z^2 + P(t)), t = 1 / (1 + z / 2) (the classic Chebyshev fit with 1.2e-7 RELATIVE error on the whole axis, so the
//     tails keep their relative accuracy) -- 9 FMAs, one MUFU.RCP, one MUFU.EX2.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {     // 1 ulp; the arguments here are >= 1 or scale-bounded
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float erfc_nonneg(float z) {
  const float t = rcp_approx(fmaf(0.5f, z, 1.0f));
  float p = 0.17087277f;
  p = fmaf(p, t, -0.82215223f); p = fmaf(p, t, 1.48851587f); p = fmaf(p, t, -1.13520398f); p = fmaf(p, t, 0.27886807f);
  p = fmaf(p, t, -0.18628806f); p = fmaf(p, t, 0.09678418f); p = fmaf(p, t, 0.37409196f); p = fmaf(p, t, 1.00002368f);
  const float e = fmaf(-z, z, fmaf(t, p, -1.26551223f));
  return t * ex2_approx(e * 1.4426950408889634f);
}
__device__ __forceinline__ float bin_mass(float d, float s) {
  const float h = rcp_approx(s);
  if (h <= 0.125f) {
    const float m = d * h, m2 = m * m, h2 = h * h;
    const float phi = 0.3989422804014327f * ex2_approx(m2 * -0.7213475204444817f);
    const float c = fmaf(h2, fmaf(h2, fmaf(m2, m2 - 6.0f, 3.0f) * (1.0f / 1920.0f), (m2 - 1.0f) * (1.0f / 24.0f)), 1.0f);
    return h * phi * c;
  }
  const float a = (d - 0.5f) * h * 0.70710678118654752440f, b = (d + 0.5f) * h * 0.70710678118654752440f;
  const float ea = erfc_nonneg(fabsf(a)), eb = erfc_nonneg(b);
  return 0.5f * ((a >= 0.f ? ea : 2.0f - ea) - eb);
}

__global__ void __launch_bounds__(256) gaussian_kernel(const TView y, const TView scales, const TView means,
                                                      const float *__restrict__ weights, int K, int mixture,
                                                      float scale_bound, float lik_bound, const TView y_hat,
                                                      const TView lik, double *log2_sum, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  double lg = 0.0;
  if (i < n) {
    int b, c, yy, xx;
    unflatten(y, i, b, c, yy, xx);
    const int M = y.C;
    float v = tload(y, b, c, yy, xx);
    float q, l;
    if (mixture) {
      q = rintf(v);
      l = 0.f;
      for (int k = 0; k < K; ++k) {
        int ck = k * M + c;
        float d = fabsf(q - tload(means, b, ck, yy, xx));
        float s = fmaxf(tload(scales, b, ck, yy, xx), scale_bound);
        float term = bin_mass(d, s) * weights[(size_t)b * K * M + ck];
        l = k == 0 ? term : l + term;
      }
    } else {
      float d;
      if (means.p0) {
        float mu = tload(means, b, c, yy, xx);
        float r = rintf(v - mu);
        q = r + mu;
        d = fabsf(q - mu);
      } else {
        q = rintf(v);
        d = fabsf(q);
      }
      float s = fmaxf(tload(scales, b, c, yy, xx), scale_bound);
      l = bin_mass(d, s);
    }
    if (lik_bound > 0.f) l = fmaxf(l, lik_bound);
    if (y_hat.p0) tstore(y_hat, b, c, yy, xx, q);
    if (lik.p0) tstore(lik, b, c, yy, xx, l);
    lg = (double)log2f(l);
  }
  if (log2_sum) block_add_double(lg, log2_sum);
}

// Tiled form for the layout the forward path uses: y / scales / means channels-last fp32 (read once with
// 128-bit loads, 4 channels per thread), y_hat / likelihood returned NCHW fp32 like the reference.  A block
// owns 32 pixels x 64 channels; results cross from channel-major registers to pixel-major stores through a
// padded shared-memory tile, so both the reads (256 B per pixel) and the writes (128 B per channel row) are
// whole lines.  Optionally also emits y_hat as bf16 (hi, lo) channels-last planes -- the synthesis stack's
// input -- saving a separate conversion pass.  Arithmetic identical to gaussian_kernel.
template <bool MIX>
__global__ void __launch_bounds__(256) gaussian_tile_kernel(const TView y, const TView scales, const TView means,
                                                           const float *__restrict__ weights, int K, float scale_bound,
                                                           float lik_bound, const TView y_hat, const TView lik,
                                                           const TView y_hat2, double *log2_sum) {
  __shared__ float s_q[64][33], s_l[64][33];
  const int HW = y.H * y.W, M = y.C;
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  const int cg = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const int c = c0 + cg * 4;
  double lg = 0.0;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int px = p0 + pl + 16 * it;
    if (px < HW && c < M) {
      const size_t pix = (size_t)b * HW + px;
      const float4 v4 = *reinterpret_cast<const float4 *>((const float *)y.p0 + pix * y.Cs + c);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      float q[4], l[4];
      if (MIX) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { q[j] = rintf(v[j]); l[j] = 0.f; }
        for (int k = 0; k < K; ++k) {
          const int ck = k * M + c;
          const float4 m4 = *reinterpret_cast<const float4 *>((const float *)means.p0 + pix * means.Cs + ck);
          const float4 s4 = *reinterpret_cast<const float4 *>((const float *)scales.p0 + pix * scales.Cs + ck);
          const float4 w4 = *reinterpret_cast<const float4 *>(weights + (size_t)b * K * M + ck);
          const float mu[4] = {m4.x, m4.y, m4.z, m4.w}, sg[4] = {s4.x, s4.y, s4.z, s4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float d = fabsf(q[j] - mu[j]);
            const float sc = fmaxf(sg[j], scale_bound);
            const float term = bin_mass(d, sc) * w[j];
            l[j] = k == 0 ? term : l[j] + term;
          }
        }
      } else {
        const float4 s4 = *reinterpret_cast<const float4 *>((const float *)scales.p0 + pix * scales.Cs + c);
        const float sg[4] = {s4.x, s4.y, s4.z, s4.w};
        float mu[4] = {0.f, 0.f, 0.f, 0.f};
        if (means.p0) {
          const float4 m4 = *reinterpret_cast<const float4 *>((const float *)means.p0 + pix * means.Cs + c);
          mu[0] = m4.x; mu[1] = m4.y; mu[2] = m4.z; mu[3] = m4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d;
          if (means.p0) {
            const float r = rintf(v[j] - mu[j]);
            q[j] = r + mu[j];
            d = fabsf(q[j] - mu[j]);
          } else {
            q[j] = rintf(v[j]);
            d = fabsf(q[j]);
          }
          const float sc = fmaxf(sg[j], scale_bound);
          l[j] = bin_mass(d, sc);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (lik_bound > 0.f) l[j] = fmaxf(l[j], lik_bound);
        s_q[cg * 4 + j][pl + 16 * it] = q[j];
        s_l[cg * 4 + j][pl + 16 * it] = l[j];
        lg += (double)log2f(l[j]);
      }
      if (y_hat2.p0) {
        __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16(q[j], hi[j], lo[j]);
        const size_t o = pix * y_hat2.Cs + c;
        *reinterpret_cast<uint2 *>((__nv_bfloat16 *)y_hat2.p0 + o) = *reinterpret_cast<const uint2 *>(hi);
        *reinterpret_cast<uint2 *>((__nv_bfloat16 *)y_hat2.p1 + o) = *reinterpret_cast<const uint2 *>(lo);
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int px = p0 + lane;
  if (px < HW) {
    for (int r = wrp; r < 64; r += 8) {
      const int ch = c0 + r;
      if (ch >= M) break;
      const size_t o = ((size_t)b * y_hat.Cs + ch) * HW + px;
      if (y_hat.p0) ((float *)y_hat.p0)[((size_t)b * y_hat.Cs + ch) * HW + px] = s_q[r][lane];
      if (lik.p0) ((float *)lik.p0)[((size_t)b * lik.Cs + ch) * HW + px] = s_l[r][lane];
      (void)o;
    }
  }
  if (log2_sum) block_add_double(lg, log2_sum);
}

// ---------------------------------------------------------------------------------------------
// spatial_pool2d: global max per (b, c).
// NHWC: block = 32 channels x 8 pixel lanes (a warp reads 32 consecutive channels of one pixel).
// Vector form (channels-last fp32, C % 4 == 0, 16-byte aligned): a thread owns one channel QUAD and every 32nd pixel; a warp
// instruction reads 4 pixels x 32 channels as 16-byte loads and four of them are in flight per thread (16 KB per block), running
// maxima advance by pointer increments.  r03: the scalar form below (one channel per thread, 64-bit index arithmetic per load:
// IMAD + LEA were 54 % of its instructions, 81 % of its stall samples waiting for loads) took 27 us for 63 MB.
__global__ void __launch_bounds__(256) spatial_max_nhwc4_kernel(const TView x, float *__restrict__ out) {
  __shared__ float4 red[32][8];
  const int b = blockIdx.y;
  const int q = threadIdx.x & 7, pl = threadIdx.x >> 3;          // channel quad, pixel lane (0..31)
  const int c = blockIdx.x * 32 + 4 * q;
  const int P = x.H * x.W;
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  if (c < x.C) {
    const size_t step = (size_t)32 * x.Cs;
    const float *ptr = (const float *)x.p0 + ((size_t)b * P + pl) * x.Cs + c;
    float4 m1 = m, m2 = m, m3 = m;
    int p = pl;
    for (; p + 96 < P; p += 128, ptr += 4 * step) {
      const float4 v0 = __ldg(reinterpret_cast<const float4 *>(ptr)), v1 = __ldg(reinterpret_cast<const float4 *>(ptr + step));
      const float4 v2 = __ldg(reinterpret_cast<const float4 *>(ptr + 2 * step)), v3 = __ldg(reinterpret_cast<const float4 *>(ptr + 3 * step));
      m = make_float4(fmaxf(m.x, v0.x), fmaxf(m.y, v0.y), fmaxf(m.z, v0.z), fmaxf(m.w, v0.w));
      m1 = make_float4(fmaxf(m1.x, v1.x), fmaxf(m1.y, v1.y), fmaxf(m1.z, v1.z), fmaxf(m1.w, v1.w));
      m2 = make_float4(fmaxf(m2.x, v2.x), fmaxf(m2.y, v2.y), fmaxf(m2.z, v2.z), fmaxf(m2.w, v2.w));
      m3 = make_float4(fmaxf(m3.x, v3.x), fmaxf(m3.y, v3.y), fmaxf(m3.z, v3.z), fmaxf(m3.w, v3.w));
    }
    for (; p < P; p += 32, ptr += step) {
      const float4 v0 = __ldg(reinterpret_cast<const float4 *>(ptr));
      m = make_float4(fmaxf(m.x, v0.x), fmaxf(m.y, v0.y), fmaxf(m.z, v0.z), fmaxf(m.w, v0.w));
    }
    m = make_float4(fmaxf(fmaxf(m.x, m1.x), fmaxf(m2.x, m3.x)), fmaxf(fmaxf(m.y, m1.y), fmaxf(m2.y, m3.y)),
                    fmaxf(fmaxf(m.z, m1.z), fmaxf(m2.z, m3.z)), fmaxf(fmaxf(m.w, m1.w), fmaxf(m2.w, m3.w)));
  }
  red[pl][q] = m;
  __syncthreads();
  if (pl == 0 && c < x.C) {
#pragma unroll 4
    for (int j = 1; j < 32; ++j) {
      const float4 v = red[j][q];
      m = make_float4(fmaxf(m.x, v.x), fmaxf(m.y, v.y), fmaxf(m.z, v.z), fmaxf(m.w, v.w));
    }
    *reinterpret_cast<float4 *>(out + (size_t)b * x.C + c) = m;
  }
}

__global__ void __launch_bounds__(256) spatial_max_nhwc_kernel(const TView x, float *__restrict__ out) {
  __shared__ float red[8][33];
  const int b = blockIdx.y;
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float m = -INFINITY;
  const int P = x.H * x.W;
  if (c < x.C) {
    if (x.fmt == HESIC_FMT_NHWC_F32) {
      // four independent running maxima per thread: the loads of a thread are in flight together
      const float *base = (const float *)x.p0 + (size_t)b * P * x.Cs + c;
      float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      int p = pl;
      for (; p + 24 < P; p += 32) {
        m = fmaxf(m, __ldg(base + (size_t)p * x.Cs));
        m1 = fmaxf(m1, __ldg(base + (size_t)(p + 8) * x.Cs));
        m2 = fmaxf(m2, __ldg(base + (size_t)(p + 16) * x.Cs));
        m3 = fmaxf(m3, __ldg(base + (size_t)(p + 24) * x.Cs));
      }
      for (; p < P; p += 8) m = fmaxf(m, __ldg(base + (size_t)p * x.Cs));
      m = fmaxf(fmaxf(m, m1), fmaxf(m2, m3));
    } else {
      for (int p = pl; p < P; p += 8) m = fmaxf(m, tload(x, b, c, p / x.W, p % x.W));
    }
  }
  red[pl][cl] = m;
  __syncthreads();
  if (pl == 0 && c < x.C) {
#pragma unroll
    for (int j = 1; j < 8; ++j) m = fmaxf(m, red[j][cl]);
    out[(size_t)b * x.C + c] = m;
  }
}

// NCHW: one warp per (b, c) plane, lanes walk pixels.
__global__ void __launch_bounds__(256) spatial_max_nchw_kernel(const TView x, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int plane = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (plane >= x.B * x.C) return;
  const int b = plane / x.C, c = plane % x.C;
  const int P = x.H * x.W;
  float m = -INFINITY;
  for (int p = lane; p < P; p += 32) m = fmaxf(m, tload(x, b, c, p / x.W, p % x.W));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) out[plane] = m;
}

// LeakyReLU -> conv1x1 -> softmax over the K mixture components (newnet1.py:498-512)
__global__ void mixture_weights_kernel(const float *__restrict__ pooled, const float *__restrict__ w,
                                       const float *__restrict__ bias, int K, int M, float *__restrict__ out) {
  extern __shared__ float logit[];  // [K]
  const int b = blockIdx.y, m = blockIdx.x;
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KM = K * M;
  const int o = k * M + m;
  const float *wr = w + (size_t)o * KM;
  const float *pin = pooled + (size_t)b * KM;
  float s = 0.f;
  for (int i = lane; i < KM; i += 32) {
    float v = pin[i];
    v = v > 0.f ? v : 0.01f * v;
    s = fmaf(wr[i], v, s);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) logit[k] = s + (bias ? bias[o] : 0.f);
  __syncthreads();
  if (threadIdx.x < K) {
    float mx = -INFINITY;
    for (int j = 0; j < K; ++j) mx = fmaxf(mx, logit[j]);
    float den = 0.f;
    for (int j = 0; j < K; ++j) den += expf(logit[j] - mx);
    out[(size_t)b * KM + threadIdx.x * M + m] = expf(logit[threadIdx.x] - mx) / den;
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_kernel(const TView x, const TView y, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b, c, oy, ox;
  unflatten(y, i, b, c, oy, ox);
  // at::native area_pixel_compute_scale / upsample_bilinear2d, align_corners=True
  float sh = y.H > 1 ? (float)(x.H - 1) / (float)(y.H - 1) : 0.f;
  float sw = y.W > 1 ? (float)(x.W - 1) / (float)(y.W - 1) : 0.f;
  float fy = sh * oy, fx = sw * ox;
  int y0 = (int)fy, x0 = (int)fx;
  int yp = y0 < x.H - 1 ? 1 : 0, xp = x0 < x.W - 1 ? 1 : 0;
  float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  float v = hy * (hx * tload(x, b, c, y0, x0) + lx * tload(x, b, c, y0, x0 + xp)) +
            ly * (hx * tload(x, b, c, y0 + yp, x0) + lx * tload(x, b, c, y0 + yp, x0 + xp));
  tstore(y, b, c, oy, ox, v);
}

// nn.UpsamplingBilinear2d on channels-last data, four channels per thread: NHWC fp32 or SPLIT in (channel slices
// allowed), SPLIT or NHWC fp32 out.  Same arithmetic as upsample_kernel (at::native upsample_bilinear2d, align_corners).
__device__ __forceinline__ float4 ld4_cl(const TView &t, size_t pix, int c) {
  const size_t o = pix * t.Cs + c;
  if (t.fmt == HESIC_FMT_NHWC_F32) return __ldg(reinterpret_cast<const float4 *>((const float *)t.p0 + o));
  const uint2 h = __ldg(reinterpret_cast<const uint2 *>((const __nv_bfloat16 *)t.p0 + o));
  const uint2 l = __ldg(reinterpret_cast<const uint2 *>((const __nv_bfloat16 *)t.p1 + o));
  return make_float4(__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16),
                     __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u),
                     __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16),
                     __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u));
}

__global__ void __launch_bounds__(256) upsample_cl_kernel(const TView x, const TView y, size_t n4) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int C4 = y.C >> 2;
  const int cq = (int)(i % C4);
  size_t r = i / C4;
  const int ox = (int)(r % y.W); r /= y.W;
  const int oy = (int)(r % y.H);
  const int b = (int)(r / y.H);
  const float sh = y.H > 1 ? (float)(x.H - 1) / (float)(y.H - 1) : 0.f;
  const float sw = y.W > 1 ? (float)(x.W - 1) / (float)(y.W - 1) : 0.f;
  const float fy = sh * oy, fx = sw * ox;
  const int y0 = (int)fy, x0 = (int)fx;
  const int yp = y0 < x.H - 1 ? 1 : 0, xp = x0 < x.W - 1 ? 1 : 0;
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const size_t row0 = ((size_t)b * x.H + y0) * x.W, row1 = ((size_t)b * x.H + y0 + yp) * x.W;
  const float4 a = ld4_cl(x, row0 + x0, 4 * cq), bq = ld4_cl(x, row0 + x0 + xp, 4 * cq);
  const float4 c = ld4_cl(x, row1 + x0, 4 * cq), d = ld4_cl(x, row1 + x0 + xp, 4 * cq);
  const float o[4] = {hy * (hx * a.x + lx * bq.x) + ly * (hx * c.x + lx * d.x), hy * (hx * a.y + lx * bq.y) + ly * (hx * c.y + lx * d.y),
                      hy * (hx * a.z + lx * bq.z) + ly * (hx * c.z + lx * d.z), hy * (hx * a.w + lx * bq.w) + ly * (hx * c.w + lx * d.w)};
  const size_t off = (((size_t)b * y.H + oy) * y.W + ox) * y.Cs + 4 * cq;
  if (y.fmt == HESIC_FMT_NHWC_SPLIT) {
    __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) split_bf16(o[k], hi[k], lo[k]);
    *reinterpret_cast<uint2 *>((__nv_bfloat16 *)y.p0 + off) = *reinterpret_cast<const uint2 *>(hi);
    *reinterpret_cast<uint2 *>((__nv_bfloat16 *)y.p1 + off) = *reinterpret_cast<const uint2 *>(lo);
  } else {
    *reinterpret_cast<float4 *>((float *)y.p0 + off) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(256) convert_kernel(const TView x, const TView y, int op, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b, c, yy, xx;
  unflatten(y, i, b, c, yy, xx);
  float v = tload(x, b, c, yy, xx);
  if (op == HESIC_OP_ABS) v = fabsf(v);
  else if (op == HESIC_OP_ROUND) v = rintf(v);
  tstore(y, b, c, yy, xx, v);
}

// NHWC fp32 -> SPLIT planes (channel slices allowed), four channels per thread: the y -> |y| / round(y) hand-offs
// between the analysis stack and the hyper path (newnet1.py:435,755)
__global__ void __launch_bounds__(256) convert_nhwc_split_kernel(const TView x, const TView y, int op, size_t n4) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int C4 = x.C >> 2;
  const int cq = (int)(i % C4);
  const size_t pix = i / C4;
  const float4 v = __ldg(reinterpret_cast<const float4 *>((const float *)x.p0 + pix * x.Cs + 4 * cq));
  float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (op == HESIC_OP_ABS) o[k] = fabsf(o[k]);
    else if (op == HESIC_OP_ROUND) o[k] = rintf(o[k]);
  }
  __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) split_bf16(o[k], hi[k], lo[k]);
  const size_t off = pix * y.Cs + 4 * cq;
  *reinterpret_cast<uint2 *>((__nv_bfloat16 *)y.p0 + off) = *reinterpret_cast<const uint2 *>(hi);
  *reinterpret_cast<uint2 *>((__nv_bfloat16 *)y.p1 + off) = *reinterpret_cast<const uint2 *>(lo);
}

// NCHW fp32 (C <= 8) -> ROWPAD split planes: thread = pixel, one 16- or 8-byte store per plane (all channel
// slots, zeros above C), reads coalesced per source plane.
__global__ void __launch_bounds__(256) rowpad_kernel(const TView x, const TView y, int op, size_t npix) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  int xx = i % x.W; size_t r = i / x.W;
  int yy = r % x.H; int b = r / x.H;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    v[c] = c < x.C ? ((const float *)x.p0)[(((size_t)b * x.Cs + c) * x.H + yy) * x.W + xx] : 0.f;
    if (op == HESIC_OP_ABS) v[c] = fabsf(v[c]);
    else if (op == HESIC_OP_ROUND) v[c] = rintf(v[c]);
  }
  store_rowpad_pixel(y, b, yy, xx, v);
}

// The 4-slot (row-pair interleaved) form, C <= 4: thread = one pixel column of a ROW PAIR, so that the two rows' slots
// (2 x 4 bf16 = 16 bytes per plane) leave as ONE full 16-byte store per plane -- the per-pixel form above writes 8 of every
// 16 bytes and leaves the other half of each sector to a different warp (r02: 72 us for 117 MB).
__global__ void __launch_bounds__(256) rowpad4_pair_kernel(const TView x, const TView y, int op, size_t ncols) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= ncols) return;
  const int xx = (int)(i % x.W);
  size_t r = i / x.W;
  const int H2 = x.H >> 1;
  const int yp = (int)(r % H2), b = (int)(r / H2);
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = k & 3, yy = 2 * yp + (k >> 2);
    float v = c < x.C ? __ldg((const float *)x.p0 + (((size_t)b * x.Cs + c) * x.H + yy) * x.W + xx) : 0.f;
    if (op == HESIC_OP_ABS) v = fabsf(v);
    else if (op == HESIC_OP_ROUND) v = rintf(v);
    split_bf16(v, hi[k], lo[k]);
  }
  const size_t o = rowpad_off(y, b, 2 * yp, xx);          // slot 0 of the even row; the odd row's slots follow
  *reinterpret_cast<uint4 *>((__nv_bfloat16 *)y.p0 + o) = *reinterpret_cast<const uint4 *>(hi);
  *reinterpret_cast<uint4 *>((__nv_bfloat16 *)y.p1 + o) = *reinterpret_cast<const uint4 *>(lo);
}

// ---------------------------------------------------------------------------------------------
// integer preparation for the host rANS coder (bit-exact with the reference)
__global__ void __launch_bounds__(256) symbols_kernel(const TView x, const float *__restrict__ cmeans, const TView means,
                                                     int32_t *__restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int xx = i % x.W; size_t r = i / x.W;
  int yy = r % x.H; r /= x.H;
  int c = r % x.C; int b = r / x.C;
  float v = tload(x, b, c, yy, xx);
  if (cmeans) v = v - cmeans[c];
  else if (means.p0) v = v - tload(means, b, c, yy, xx);
  out[i] = (int32_t)rintf(v);
}

__global__ void __launch_bounds__(256) indexes_channel_kernel(int C, int HW, int32_t *__restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)((i / HW) % C);
}

__global__ void __launch_bounds__(256) indexes_scale_kernel(const TView sc, const float *__restrict__ table, int nt,
                                                           float bound, int32_t *__restrict__ out, size_t n) {
  extern __shared__ float tab[];
  for (int j = threadIdx.x; j < nt; j += blockDim.x) tab[j] = table[j];
  __syncthreads();
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int xx = i % sc.W; size_t r = i / sc.W;
  int yy = r % sc.H; r /= sc.H;
  int c = r % sc.C; int b = r / sc.C;
  float s = fmaxf(tload(sc, b, c, yy, xx), bound);
  int idx = nt - 1;
  for (int j = 0; j < nt - 1; ++j) idx -= (s <= tab[j]) ? 1 : 0;
  out[i] = idx;
}

// ---------------------------------------------------------------------------------------------
// File codec (HSIC.compress / decompress, newnet1.py:934-978): per latent element the integer cumulative-frequency
// row of its K-component mixture.  Thread = element; the arithmetic follows the reference's op sequence in fp32
// (torch elementwise ops for the pmf, then numpy: clip, pairwise-summed normalisation, round half even, running sum).
//
// These integers ARE the code: encoder and decoder must derive identical rows.  The one operation of that sequence
// whose result is implementation-defined is erfc -- the reference calls torch.erfc on whatever device it runs on (CUDA's
// erfcf on 'cuda:0', SLEEF / libm on a CPU), and those differ in the last bit.  The table arithmetic therefore uses the
// CORRECTLY ROUNDED fp32 erfc: evaluated in fp64 on the fp32 argument and rounded once (cdf_std_cumulative).  Every
// other step is a single IEEE fp32 operation, so the rows are reproducible bit for bit by any implementation -- the
// CPU checker of the test suite restates exactly that and tests/test_gpu_codec.py holds equality.
constexpr int CDF_MAX_S = 129;   // rows of up to 129 samples (minmax <= 64) are built in registers / local memory;
                                 // longer ones in place in the output row (any minmax)

__device__ __forceinline__ float cdf_std_cumulative(float v) {
  const float t = __fmul_rn(-0.70710678118654752440f, v);      // float(-(2 ** -0.5)) * inputs   (newnet1.py:795-797)
  return __fmul_rn(0.5f, (float)erfc((double)t));
}

// numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src: pairwise_sum) for a contiguous fp32 vector
__device__ float np_pairwise_sum(const float *a, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

__global__ void __launch_bounds__(128) gmm_cdf_kernel(const TView scales, const TView means, const float *__restrict__ weights,
                                                     int K, int M, const int32_t *__restrict__ channels, int n_channels,
                                                     int minmax, float scale_bound, int32_t *__restrict__ out) {
  const int HW = scales.H * scales.W;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_channels * HW) return;
  const int ci = e / HW, pos = e - ci * HW;
  const int c = channels[ci], yy = pos / scales.W, xx = pos - yy * scales.W;
  const int S = 2 * minmax + 1;
  int32_t *row = out + (size_t)e * (S + 1);
  float local[CDF_MAX_S];
  // pmf[s] of a long row lives in the slot row[s + 1] it is finally replaced by (read, then written, by this thread only)
  float *pmf = S <= CDF_MAX_S ? local : reinterpret_cast<float *>(row + 1);
  for (int k = 0; k < K; ++k) {
    const int ck = k * M + c;
    const float mu = __fadd_rn(tload(means, 0, ck, yy, xx), (float)minmax);      // means + minmax  (newnet1.py:950)
    const float sc = fmaxf(tload(scales, 0, ck, yy, xx), scale_bound);           // lower_bound_scale
    const float w = weights[ck];
    for (int sidx = 0; sidx < S; ++sidx) {
      const float v = fabsf(__fsub_rn((float)sidx, mu));
      const float up = cdf_std_cumulative(__fdiv_rn(__fsub_rn(0.5f, v), sc));
      const float lo = cdf_std_cumulative(__fdiv_rn(__fsub_rn(-0.5f, v), sc));
      const float term = __fmul_rn(__fsub_rn(up, lo), w);
      pmf[sidx] = k == 0 ? term : __fadd_rn(pmf[sidx], term);
    }
  }
  for (int sidx = 0; sidx < S; ++sidx) pmf[sidx] = fminf(fmaxf(pmf[sidx], 1.0f / 65536.0f), 1.0f);   // np.clip
  const float tot = np_pairwise_sum(pmf, S);
  float acc = 0.f;
  row[0] = 0;
  for (int sidx = 0; sidx < S; ++sidx) {
    const float q = rintf(__fmul_rn(__fdiv_rn(pmf[sidx], tot), 65536.0f));       // np.round(pmf / sum * 65536)
    acc = __fadd_rn(acc, q);                                                     // np.add.accumulate (fp32)
    row[sidx + 1] = (int32_t)acc;
  }
}

// both tensors dense fp32 in the same layout: 128-bit loads, fp32 squares folded into fp64 four at a time
__global__ void __launch_bounds__(256) sse_dense_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, double *acc,
                                                       size_t n4) {
  double s = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 u = __ldg(a + i), v = __ldg(b + i);
    const float dx = u.x - v.x, dy = u.y - v.y, dz = u.z - v.z, dw = u.w - v.w;
    s += ((double)dx * (double)dx + (double)dy * (double)dy) + ((double)dz * (double)dz + (double)dw * (double)dw);
  }
  block_add_double(s, acc);
}

__global__ void __launch_bounds__(256) sse_kernel(const TView a, const TView b_, double *acc, size_t n) {
  double s = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int b, c, y, x;
    unflatten(a, i, b, c, y, x);
    float d = tload(a, b, c, y, x) - tload(b_, b, c, y, x);
    s += (double)d * (double)d;
  }
  block_add_double(s, acc);
}

static inline size_t numel(const hesic_tensor *t) { return (size_t)t->B * t->C * t->H * t->W; }
static inline unsigned nblk(size_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace hesic

using namespace hesic;

extern "C" int hesic_warp_perspective(const hesic_tensor *src, const float *M, const hesic_tensor *dst,
                                      const hesic_tensor *dst_rowpad, int align_corners, void *stream) {
  int r;
  if ((r = check_tensor(src, "warp src")) != HESIC_OK) return r;
  if ((r = check_tensor(dst, "warp dst")) != HESIC_OK) return r;
  TView d2;
  memset(&d2, 0, sizeof(d2));
  if (dst_rowpad && dst_rowpad->p0) {
    if ((r = check_tensor(dst_rowpad, "warp dst_rowpad")) != HESIC_OK) return r;
    HESIC_REQUIRE(dst_rowpad->fmt == HESIC_FMT_ROWPAD8_SPLIT && same_shape(dst, dst_rowpad) && dst->C <= dst_rowpad->Cs,
                  "warp: dst_rowpad must be a ROWPAD tensor shaped like dst");
    HESIC_REQUIRE((((uintptr_t)dst_rowpad->p0 | (uintptr_t)dst_rowpad->p1) & 15u) == 0, "warp: dst_rowpad is misaligned");
    d2 = view(dst_rowpad);
  }
  HESIC_REQUIRE(M != nullptr, "warp: null homography");
  HESIC_REQUIRE(src->B == dst->B && src->C == dst->C, "warp: batch/channel mismatch");
  HESIC_REQUIRE(src->p0 != dst->p0, "warp: in-place is not supported");
  if (numel(dst) == 0) return HESIC_OK;
  HESIC_REQUIRE(src->H < 60000 && src->W < 60000, "warp: source image too large");
  // forward-path form: RGB, NCHW fp32 on both sides, rows 16-byte aligned -> per-warp staging (warp_rgb_kernel)
  if (src->fmt == HESIC_FMT_NCHW_F32 && dst->fmt == HESIC_FMT_NCHW_F32 && src->C == 3 && (src->W & 3) == 0 &&
      ((uintptr_t)src->p0 & 15u) == 0) {
    dim3 blk(32, WRGB_WARPS), grid((dst->W + 31) / 32, (dst->H + WRGB_WARPS * WRGB_ROWS - 1) / (WRGB_WARPS * WRGB_ROWS), dst->B);
    CUtensorMap msrc;
    {
      const int sCs = src->Cs > 0 ? src->Cs : src->C;
      const uint64_t dims[4] = {(uint64_t)src->W, (uint64_t)src->H, 3, (uint64_t)src->B};
      const uint64_t strides[3] = {(uint64_t)src->W * 4, (uint64_t)src->H * src->W * 4, (uint64_t)sCs * src->H * src->W * 4};
      const uint32_t box[4] = {(uint32_t)WRGB_BW, (uint32_t)WRGB_BH, 3u, 1u};
      if ((r = tc::make_tensor_map(&msrc, src->p0, 4, dims, strides, box, true, false)) != HESIC_OK) return r;
    }
    warp_rgb_kernel<<<grid, blk, 0, as_stream(stream)>>>(msrc, view(src), M, view(dst), d2, align_corners);
    HESIC_LAUNCHED("warp_rgb_kernel");
    return HESIC_OK;
  }
  static const int rows = diag_env("HESIC_WARP_ROWS") ? atoi(diag_env("HESIC_WARP_ROWS")) : 4;
  const int ty = 8 * (rows == 1 ? 1 : (rows == 2 ? 2 : 4));
  dim3 blk(32, 8), grid((dst->W + WARP_TILE - 1) / WARP_TILE, (dst->H + ty - 1) / ty, dst->B);
  if (rows == 1) warp_kernel<1><<<grid, blk, 0, as_stream(stream)>>>(view(src), M, view(dst), d2, align_corners);
  else if (rows == 2) warp_kernel<2><<<grid, blk, 0, as_stream(stream)>>>(view(src), M, view(dst), d2, align_corners);
  else warp_kernel<4><<<grid, blk, 0, as_stream(stream)>>>(view(src), M, view(dst), d2, align_corners);
  HESIC_LAUNCHED("warp_kernel");
  return HESIC_OK;
}

extern "C" int hesic_eb_pack(const float *const *m, const float *const *b, const float *const *f, const float *quantiles,
                             int C, float *params_out, void *stream) {
  HESIC_REQUIRE(m && b && f && quantiles && params_out && C > 0, "eb_pack: null argument");
  eb_pack_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(m[0], m[1], m[2], m[3], m[4], b[0], b[1], b[2], b[3], b[4],
                                                                f[0], f[1], f[2], f[3], quantiles, C, params_out);
  HESIC_LAUNCHED("eb_pack_kernel");
  return HESIC_OK;
}

static hesic_tensor null_tensor() {
  hesic_tensor t;
  t.p0 = nullptr; t.p1 = nullptr; t.fmt = 0; t.B = t.C = t.H = t.W = t.Cs = 0;
  return t;
}

extern "C" int hesic_entropy_bottleneck(const hesic_tensor *z, const float *params, float likelihood_bound,
                                        const hesic_tensor *z_hat, const hesic_tensor *lik, double *log2_sum,
                                        void *stream) {
  int r;
  if ((r = check_tensor(z, "eb input")) != HESIC_OK) return r;
  HESIC_REQUIRE(params != nullptr, "eb: null params");
  hesic_tensor nt = null_tensor();
  const hesic_tensor *zh = z_hat ? z_hat : &nt, *lk = lik ? lik : &nt;
  if (zh->p0) HESIC_REQUIRE(same_shape(z, zh), "eb: z_hat shape mismatch");
  if (lk->p0) HESIC_REQUIRE(same_shape(z, lk), "eb: likelihood shape mismatch");
  size_t n = numel(z);
  if (n == 0) return HESIC_OK;
  eb_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(z), params, likelihood_bound, view(zh), view(lk), log2_sum, n);
  HESIC_LAUNCHED("eb_kernel");
  return HESIC_OK;
}

static bool vec4_ok(const hesic_tensor *t) {
  const int Cs = t->Cs > 0 ? t->Cs : t->C;
  return t->fmt == HESIC_FMT_NHWC_F32 && (Cs & 3) == 0 && ((uintptr_t)t->p0 & 15u) == 0;
}

extern "C" int hesic_gaussian_conditional(const hesic_tensor *y, const hesic_tensor *scales, const hesic_tensor *means,
                                          const float *weights, int K, int mixture, float scale_bound,
                                          float likelihood_bound, const hesic_tensor *y_hat, const hesic_tensor *lik,
                                          const hesic_tensor *y_hat_split, double *log2_sum, void *stream) {
  int r;
  if ((r = check_tensor(y, "gaussian input")) != HESIC_OK) return r;
  if ((r = check_tensor(scales, "gaussian scales")) != HESIC_OK) return r;
  hesic_tensor nt = null_tensor();
  const hesic_tensor *mu = means ? means : &nt, *yh = y_hat ? y_hat : &nt, *lk = lik ? lik : &nt;
  HESIC_REQUIRE(K >= 1, "gaussian: K must be >= 1");
  if (mixture) {
    HESIC_REQUIRE(weights != nullptr && mu->p0 != nullptr, "gaussian mixture: means and weights are required");
    HESIC_REQUIRE(scales->C == K * y->C && mu->C == K * y->C, "gaussian mixture: scales/means need K*M channels");
  } else {
    HESIC_REQUIRE(K == 1 && scales->C == y->C, "gaussian: scales channel mismatch");
    if (mu->p0) HESIC_REQUIRE(mu->C == y->C, "gaussian: means channel mismatch");
  }
  HESIC_REQUIRE(scales->B == y->B && scales->H == y->H && scales->W == y->W, "gaussian: scales shape mismatch");
  if (mu->p0) HESIC_REQUIRE(mu->B == y->B && mu->H == y->H && mu->W == y->W, "gaussian: means shape mismatch");
  if (yh->p0) HESIC_REQUIRE(same_shape(y, yh), "gaussian: y_hat shape mismatch");
  if (lk->p0) HESIC_REQUIRE(same_shape(y, lk), "gaussian: likelihood shape mismatch");
  const hesic_tensor *y2 = y_hat_split ? y_hat_split : &nt;
  if (y2->p0) {
    HESIC_REQUIRE(same_shape(y, y2) && y2->fmt == HESIC_FMT_NHWC_SPLIT, "gaussian: y_hat_split must be a SPLIT tensor like y");
    int r2;
    if ((r2 = check_tensor(y2, "gaussian y_hat_split")) != HESIC_OK) return r2;
  }
  size_t n = numel(y);
  if (n == 0) return HESIC_OK;
  // the forward path's layout: channels-last inputs read with 128-bit loads, NCHW outputs through a smem transpose
  const bool tiled = vec4_ok(y) && vec4_ok(scales) && (!mu->p0 || vec4_ok(mu)) && (y->C & 3) == 0 &&
                     (!yh->p0 || yh->fmt == HESIC_FMT_NCHW_F32) && (!lk->p0 || lk->fmt == HESIC_FMT_NCHW_F32) &&
                     (!mixture || ((uintptr_t)weights & 15u) == 0) && y->B <= 65535 &&
                     (!y2->p0 || (((y2->Cs > 0 ? y2->Cs : y2->C) & 3) == 0 && (((uintptr_t)y2->p0 | (uintptr_t)y2->p1) & 7u) == 0));
  if (tiled) {
    dim3 grid((y->H * y->W + 31) / 32, (y->C + 63) / 64, y->B);
    if (mixture)
      gaussian_tile_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(view(y), view(scales), view(mu), weights, K, scale_bound,
                                                                    likelihood_bound, view(yh), view(lk), view(y2), log2_sum);
    else
      gaussian_tile_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(view(y), view(scales), view(mu), weights, K, scale_bound,
                                                                     likelihood_bound, view(yh), view(lk), view(y2), log2_sum);
    HESIC_LAUNCHED("gaussian_tile_kernel");
    return HESIC_OK;
  }
  gaussian_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(y), view(scales), view(mu), weights, K, mixture,
                                                         scale_bound, likelihood_bound, view(yh), view(lk), log2_sum, n);
  HESIC_LAUNCHED("gaussian_kernel");
  if (y2->p0) {
    // generic path: second copy of y_hat by conversion
    HESIC_REQUIRE(yh->p0 != nullptr, "gaussian: y_hat_split needs y_hat on the generic path");
    return hesic_convert(yh, y2, HESIC_OP_COPY, stream);
  }
  return HESIC_OK;
}

extern "C" int hesic_spatial_max(const hesic_tensor *x, float *out_max, void *stream) {
  int r;
  if ((r = check_tensor(x, "spatial_max input")) != HESIC_OK) return r;
  HESIC_REQUIRE(out_max != nullptr, "spatial_max: null output");
  HESIC_REQUIRE(x->H * x->W > 0, "spatial_max: empty spatial extent");
  if (x->B * x->C == 0) return HESIC_OK;
  if (x->fmt == HESIC_FMT_NCHW_F32) {
    spatial_max_nchw_kernel<<<(x->B * x->C + 7) / 8, 256, 0, as_stream(stream)>>>(view(x), out_max);
  } else {
    dim3 grid((x->C + 31) / 32, x->B);
    const int xCs = x->Cs > 0 ? x->Cs : x->C;
    if (x->fmt == HESIC_FMT_NHWC_F32 && (x->C & 3) == 0 && (xCs & 3) == 0 && ((uintptr_t)x->p0 & 15u) == 0 && ((uintptr_t)out_max & 15u) == 0)
      spatial_max_nhwc4_kernel<<<grid, 256, 0, as_stream(stream)>>>(view(x), out_max);
    else
      spatial_max_nhwc_kernel<<<grid, 256, 0, as_stream(stream)>>>(view(x), out_max);
  }
  HESIC_LAUNCHED("spatial_max_kernel");
  return HESIC_OK;
}

extern "C" int hesic_mixture_weights(const float *pooled, const float *w1x1, const float *bias, int B, int K, int M,
                                     float *out, void *stream) {
  HESIC_REQUIRE(pooled && w1x1 && out, "mixture_weights: null argument");
  HESIC_REQUIRE(K >= 1 && K <= 32 && M >= 1 && B >= 0, "mixture_weights: bad sizes");
  if (B == 0) return HESIC_OK;
  mixture_weights_kernel<<<dim3(M, B), K * 32, K * sizeof(float), as_stream(stream)>>>(pooled, w1x1, bias, K, M, out);
  HESIC_LAUNCHED("mixture_weights_kernel");
  return HESIC_OK;
}

extern "C" int hesic_upsample_bilinear(const hesic_tensor *x, const hesic_tensor *y, int scale, void *stream) {
  int r;
  if ((r = check_tensor(x, "upsample input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "upsample output")) != HESIC_OK) return r;
  HESIC_REQUIRE(scale >= 1 && y->H == x->H * scale && y->W == x->W * scale && y->B == x->B && y->C == x->C,
                "upsample: output shape mismatch");
  size_t n = numel(y);
  if (n == 0) return HESIC_OK;
  {
    const int xCs = x->Cs > 0 ? x->Cs : x->C, yCs = y->Cs > 0 ? y->Cs : y->C;
    const bool xin = x->fmt == HESIC_FMT_NHWC_F32 || x->fmt == HESIC_FMT_NHWC_SPLIT;
    const bool yout = y->fmt == HESIC_FMT_NHWC_F32 || y->fmt == HESIC_FMT_NHWC_SPLIT;
    const uintptr_t xa = x->fmt == HESIC_FMT_NHWC_F32 ? 15 : 7, ya = y->fmt == HESIC_FMT_NHWC_F32 ? 15 : 7;
    const bool al = ((uintptr_t)x->p0 & xa) == 0 && ((uintptr_t)y->p0 & ya) == 0 &&
                    (x->fmt != HESIC_FMT_NHWC_SPLIT || ((uintptr_t)x->p1 & 7) == 0) &&
                    (y->fmt != HESIC_FMT_NHWC_SPLIT || ((uintptr_t)y->p1 & 7) == 0);
    if (xin && yout && al && (x->C & 3) == 0 && (xCs & 3) == 0 && (yCs & 3) == 0) {
      upsample_cl_kernel<<<nblk(n / 4), 256, 0, as_stream(stream)>>>(view(x), view(y), n / 4);
      HESIC_LAUNCHED("upsample_cl_kernel");
      return HESIC_OK;
    }
  }
  upsample_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(x), view(y), n);
  HESIC_LAUNCHED("upsample_kernel");
  return HESIC_OK;
}

extern "C" int hesic_convert(const hesic_tensor *x, const hesic_tensor *y, int op, void *stream) {
  int r;
  if ((r = check_tensor(x, "convert input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "convert output")) != HESIC_OK) return r;
  HESIC_REQUIRE(same_shape(x, y), "convert: shape mismatch");
  HESIC_REQUIRE(op >= 0 && op <= 2, "convert: bad op");
  size_t n = numel(y);
  if (n == 0) return HESIC_OK;
  if (y->fmt == HESIC_FMT_ROWPAD8_SPLIT && x->fmt == HESIC_FMT_NCHW_F32 && ((uintptr_t)y->p0 & 15) == 0 &&
      ((uintptr_t)y->p1 & 15) == 0) {
    // whole-pixel fast path (the view starts at channel slot 0 and owns all 8 slots)
    size_t npix = (size_t)y->B * y->H * y->W;
    if (y->Cs == 4 && (y->H & 1) == 0 && x->C <= 4) {
      rowpad4_pair_kernel<<<nblk(npix / 2), 256, 0, as_stream(stream)>>>(view(x), view(y), op, npix / 2);
      HESIC_LAUNCHED("rowpad4_pair_kernel");
      return HESIC_OK;
    }
    rowpad_kernel<<<nblk(npix), 256, 0, as_stream(stream)>>>(view(x), view(y), op, npix);
    HESIC_LAUNCHED("rowpad_kernel");
    return HESIC_OK;
  }
  if (x->fmt == HESIC_FMT_NHWC_F32 && y->fmt == HESIC_FMT_NHWC_SPLIT && (x->C & 3) == 0 && ((x->Cs > 0 ? x->Cs : x->C) & 3) == 0 &&
      ((y->Cs > 0 ? y->Cs : y->C) & 3) == 0 && ((uintptr_t)x->p0 & 15) == 0 && (((uintptr_t)y->p0 | (uintptr_t)y->p1) & 7) == 0) {
    convert_nhwc_split_kernel<<<nblk(n / 4), 256, 0, as_stream(stream)>>>(view(x), view(y), op, n / 4);
    HESIC_LAUNCHED("convert_nhwc_split_kernel");
    return HESIC_OK;
  }
  convert_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(x), view(y), op, n);
  HESIC_LAUNCHED("convert_kernel");
  return HESIC_OK;
}

// uint8 [B][H][W][C] -> NCHW fp32 / 255.  Thread = 4 consecutive pixels of one row: C x 4 bytes in (three 32-bit loads for
// RGB), one float4 per channel plane out -- both sides coalesced.
template <int CH>
__global__ void __launch_bounds__(256) images_u8_kernel(const uint8_t *__restrict__ src, float *__restrict__ dst, int Cs, size_t HW,
                                                        size_t nquads) {
  const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (q >= nquads) return;
  const size_t pix = q * 4, b = pix / HW, r = pix - b * HW;
  uint32_t w[CH];
  const uint32_t *s = reinterpret_cast<const uint32_t *>(src + pix * CH);
#pragma unroll
  for (int i = 0; i < CH; ++i) w[i] = __ldg(s + i);
  float v[CH][4];
#pragma unroll
  for (int i = 0; i < 4 * CH; ++i) {
    const uint32_t byte = (w[i >> 2] >> (8 * (i & 3))) & 0xffu;      // byte i = (pixel i / CH, channel i % CH)
    v[i % CH][i / CH] = __fdiv_rn((float)byte, 255.0f);
  }
#pragma unroll
  for (int c = 0; c < CH; ++c)
    *reinterpret_cast<float4 *>(dst + (b * Cs + c) * HW + r) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
}

__global__ void __launch_bounds__(256) images_u8_generic_kernel(const uint8_t *__restrict__ src, const TView dst, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % dst.C);
  size_t p = i / dst.C;
  const int x = (int)(p % dst.W); p /= dst.W;
  const int y = (int)(p % dst.H);
  tstore(dst, (int)(p / dst.H), c, y, x, __fdiv_rn((float)src[i], 255.0f));
}

extern "C" int hesic_images_from_u8(const uint8_t *src, int B, int H, int W, int C, const hesic_tensor *dst, void *stream) {
  int r;
  if ((r = check_tensor(dst, "images_from_u8 dst")) != HESIC_OK) return r;
  HESIC_REQUIRE(B >= 0 && H >= 0 && W >= 0 && C >= 1, "images_from_u8: bad size");
  HESIC_REQUIRE(dst->B == B && dst->C == C && dst->H == H && dst->W == W, "images_from_u8: dst is %dx%dx%dx%d, source %dx%dx%dx%d", dst->B,
                dst->C, dst->H, dst->W, B, C, H, W);
  const size_t n = numel(dst);
  if (n == 0) return HESIC_OK;
  HESIC_REQUIRE(src != nullptr, "images_from_u8: null source");
  const size_t HW = (size_t)H * W;
  if (dst->fmt == HESIC_FMT_NCHW_F32 && C == 3 && (W & 3) == 0 && ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst->p0 & 15) == 0) {
    const size_t nq = (size_t)B * HW / 4;
    images_u8_kernel<3><<<nblk(nq), 256, 0, as_stream(stream)>>>(src, (float *)dst->p0, dst->Cs > 0 ? dst->Cs : C, HW, nq);
    HESIC_LAUNCHED("images_u8_kernel");
    return HESIC_OK;
  }
  images_u8_generic_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(src, view(dst), n);
  HESIC_LAUNCHED("images_u8_generic_kernel");
  return HESIC_OK;
}

extern "C" int hesic_prepare_symbols(const hesic_tensor *x, const float *channel_means, const hesic_tensor *means,
                                     int32_t *out_symbols, void *stream) {
  int r;
  if ((r = check_tensor(x, "symbols input")) != HESIC_OK) return r;
  HESIC_REQUIRE(out_symbols != nullptr || numel(x) == 0, "symbols: null output");
  hesic_tensor nt = null_tensor();
  const hesic_tensor *mu = means ? means : &nt;
  if (mu->p0) HESIC_REQUIRE(same_shape(x, mu), "symbols: means shape mismatch");
  size_t n = numel(x);
  if (n == 0) return HESIC_OK;
  symbols_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(view(x), channel_means, view(mu), out_symbols, n);
  HESIC_LAUNCHED("symbols_kernel");
  return HESIC_OK;
}

extern "C" int hesic_build_indexes_channel(int B, int C, int H, int W, int32_t *out, void *stream) {
  HESIC_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0, "indexes: negative size");
  size_t n = (size_t)B * C * H * W;
  if (n == 0) return HESIC_OK;
  HESIC_REQUIRE(out != nullptr, "indexes: null output");
  indexes_channel_kernel<<<nblk(n), 256, 0, as_stream(stream)>>>(C, H * W, out, n);
  HESIC_LAUNCHED("indexes_channel_kernel");
  return HESIC_OK;
}

extern "C" int hesic_build_indexes_scale(const hesic_tensor *scales, const float *table, int n_table, float scale_bound,
                                         int32_t *out, void *stream) {
  int r;
  if ((r = check_tensor(scales, "indexes scales")) != HESIC_OK) return r;
  HESIC_REQUIRE(table && n_table >= 1 && n_table <= 4096, "indexes: bad scale table");
  size_t n = numel(scales);
  if (n == 0) return HESIC_OK;
  HESIC_REQUIRE(out != nullptr, "indexes: null output");
  indexes_scale_kernel<<<nblk(n), 256, n_table * sizeof(float), as_stream(stream)>>>(view(scales), table, n_table,
                                                                                    scale_bound, out, n);
  HESIC_LAUNCHED("indexes_scale_kernel");
  return HESIC_OK;
}

extern "C" int hesic_gmm_cdf_tables(const hesic_tensor *scales, const hesic_tensor *means, const float *weights, int K, int M,
                                    const int32_t *channels, int n_channels, int minmax, float scale_bound,
                                    int32_t *out_cdf, void *stream) {
  int r;
  if ((r = check_tensor(scales, "cdf tables scales")) != HESIC_OK) return r;
  if ((r = check_tensor(means, "cdf tables means")) != HESIC_OK) return r;
  HESIC_REQUIRE(weights && out_cdf && (channels || n_channels == 0), "cdf tables: null argument");
  HESIC_REQUIRE(K >= 1 && M >= 1 && scales->C == K * M && means->C == K * M, "cdf tables: scales/means need K*M channels");
  HESIC_REQUIRE(scales->B == 1 && means->B == 1 && means->H == scales->H && means->W == scales->W,
                "cdf tables: one image at a time (as the reference's codec), scales and means of the same size");
  HESIC_REQUIRE(minmax >= 1 && minmax <= 32767, "cdf tables: minmax must be in [1, 32767]");
  HESIC_REQUIRE(n_channels >= 0 && n_channels <= M, "cdf tables: bad channel count");
  const int n = n_channels * scales->H * scales->W;
  if (n == 0) return HESIC_OK;
  gmm_cdf_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(view(scales), view(means), weights, K, M, channels, n_channels,
                                                                 minmax, scale_bound, out_cdf);
  HESIC_LAUNCHED("gmm_cdf_kernel");
  return HESIC_OK;
}

extern "C" int hesic_sum_squared_error(const hesic_tensor *a, const hesic_tensor *b, double *acc, void *stream) {
  int r;
  if ((r = check_tensor(a, "sse a")) != HESIC_OK) return r;
  if ((r = check_tensor(b, "sse b")) != HESIC_OK) return r;
  HESIC_REQUIRE(same_shape(a, b) && acc, "sse: shape mismatch or null accumulator");
  size_t n = numel(a);
  if (n == 0) return HESIC_OK;
  unsigned blocks = nblk(n) < 148 * 8 ? nblk(n) : 148 * 8;
  const bool dense = a->fmt == b->fmt && (a->fmt == HESIC_FMT_NCHW_F32 || a->fmt == HESIC_FMT_NHWC_F32) &&
                     (a->Cs == 0 || a->Cs == a->C) && (b->Cs == 0 || b->Cs == b->C) && (n & 3) == 0 &&
                     (((uintptr_t)a->p0 | (uintptr_t)b->p0) & 15u) == 0;
  if (dense) {
    sse_dense_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const float4 *)a->p0, (const float4 *)b->p0, acc, n / 4);
    HESIC_LAUNCHED("sse_dense_kernel");
    return HESIC_OK;
  }
  sse_kernel<<<blocks, 256, 0, as_stream(stream)>>>(view(a), view(b), acc, n);
  HESIC_LAUNCHED("sse_kernel");
  return HESIC_OK;
}

// ---------------------------------------------------------------------------------------------
// Homography front-end glue: kornia.get_perspective_transform (+ torch.inverse) and nn.MaxPool2d(2, 2).
namespace hesic {

// One thread per pair: the 4-point direct linear transform.  Unknowns h0..h7 of H = [[h0 h1 h2] [h3 h4 h5] [h6 h7 1]] from
//   [x y 1 0 0 0 -x u -y u] h = u,  [0 0 0 x y 1 -x v -y v] h = v      (the system the reference's call sites solve through
// kornia, ywz/mywork/test3real.py:179, udh/udh/model.py:27), Gaussian elimination with partial pivoting in fp64 (the fp32
// LAPACK / cuSOLVER solves the reference runs differ from each other in the last digits; fp64 is closer to the exact answer
// than either), optionally followed by the 3 x 3 inverse (torch.inverse, test3real.py:180) through the adjugate.
__global__ void perspective_transform_kernel(const float *__restrict__ src, const float *__restrict__ dst, int B, int invert,
                                             float *__restrict__ H) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double A[8][9];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double x = src[(b * 4 + i) * 2], y = src[(b * 4 + i) * 2 + 1];
    const double u = dst[(b * 4 + i) * 2], v = dst[(b * 4 + i) * 2 + 1];
    double *r0 = A[i], *r1 = A[4 + i];
    r0[0] = x; r0[1] = y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -x * u; r0[7] = -y * u; r0[8] = u;
    r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x; r1[4] = y; r1[5] = 1; r1[6] = -x * v; r1[7] = -y * v; r1[8] = v;
  }
  for (int c = 0; c < 8; ++c) {
    int piv = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < 8; ++r)
      if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (piv != c)
      for (int k = c; k < 9; ++k) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
    const double inv = 1.0 / A[c][c];      // singular system (degenerate corners): inf / nan propagate to H, as documented
    for (int r = c + 1; r < 8; ++r) {
      const double f = A[r][c] * inv;
      for (int k = c; k < 9; ++k) A[r][k] -= f * A[c][k];
    }
  }
  double h[9];
  for (int c = 7; c >= 0; --c) {
    double s = A[c][8];
    for (int k = c + 1; k < 8; ++k) s -= A[c][k] * h[k];
    h[c] = s / A[c][c];
  }
  h[8] = 1.0;
  // the reference path rounds H to fp32 before inverting it (kornia returns fp32, torch.inverse takes it)
  for (int k = 0; k < 9; ++k) h[k] = (double)(float)h[k];
  if (invert) {
    const double a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7], i = h[8];
    const double c00 = e * i - f * hh, c01 = f * g - d * i, c02 = d * hh - e * g;
    const double idet = 1.0 / (a * c00 + bb * c01 + c * c02);
    const double o[9] = {c00 * idet, (c * hh - bb * i) * idet, (bb * f - c * e) * idet,
                         c01 * idet, (a * i - c * g) * idet, (c * d - a * f) * idet,
                         c02 * idet, (bb * g - a * hh) * idet, (a * e - bb * d) * idet};
    for (int k = 0; k < 9; ++k) h[k] = o[k];
  }
  for (int k = 0; k < 9; ++k) H[b * 9 + k] = (float)h[k];
}

// nn.MaxPool2d(2, 2) on dense NCHW fp32 (floor mode: a trailing odd row / column is dropped); thread = two output pixels
__global__ void max_pool2x2_kernel(const float *__restrict__ x, float *__restrict__ y, int H, int W, int Ho, int Wo, size_t n2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const int Wo2 = (Wo + 1) / 2;
  const int ox = (int)(i % Wo2) * 2;
  const size_t r = i / Wo2;
  const int oy = (int)(r % Ho);
  const size_t plane = r / Ho;
  const float *p = x + (plane * H + 2 * (size_t)oy) * W + 2 * ox;
  float *o = y + (plane * Ho + oy) * Wo + ox;
  if (ox + 1 < Wo && (((uintptr_t)p | (uintptr_t)(p + W)) & 15u) == 0) {
    const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + W);
    o[0] = fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y));
    o[1] = fmaxf(fmaxf(a.z, a.w), fmaxf(b.z, b.w));
  } else {
    o[0] = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[W], p[W + 1]));
    if (ox + 1 < Wo) o[1] = fmaxf(fmaxf(p[2], p[3]), fmaxf(p[W + 2], p[W + 3]));
  }
}

}  // namespace hesic

extern "C" int hesic_perspective_transform(const float *src, const float *dst, int B, int invert, float *H, void *stream) {
  HESIC_REQUIRE(src && dst && H && B >= 0, "perspective_transform: null argument");
  if (B == 0) return HESIC_OK;
  perspective_transform_kernel<<<(B + 63) / 64, 64, 0, as_stream(stream)>>>(src, dst, B, invert, H);
  HESIC_LAUNCHED("perspective_transform_kernel");
  return HESIC_OK;
}

extern "C" int hesic_max_pool2x2(const hesic_tensor *x, const hesic_tensor *y, void *stream) {
  int r;
  if ((r = check_tensor(x, "max_pool input")) != HESIC_OK) return r;
  if ((r = check_tensor(y, "max_pool output")) != HESIC_OK) return r;
  HESIC_REQUIRE(x->fmt == HESIC_FMT_NCHW_F32 && y->fmt == HESIC_FMT_NCHW_F32 && (x->Cs == 0 || x->Cs == x->C) &&
                    (y->Cs == 0 || y->Cs == y->C),
                "max_pool2x2: dense NCHW fp32 tensors only");
  HESIC_REQUIRE(y->B == x->B && y->C == x->C && y->H == x->H / 2 && y->W == x->W / 2, "max_pool2x2: output shape mismatch");
  const size_t n2 = (size_t)y->B * y->C * y->H * ((y->W + 1) / 2);
  if (n2 == 0) return HESIC_OK;
  max_pool2x2_kernel<<<nblk(n2), 256, 0, as_stream(stream)>>>((const float *)x->p0, (float *)y->p0, x->H, x->W, y->H, y->W, n2);
  HESIC_LAUNCHED("max_pool2x2_kernel");
  return HESIC_OK;
}
