/* CPU oracle for the host-side entropy-coder boundary.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of
 *   - compressai/cpp_exts/ops/ops.cpp:24-81         pmf_to_quantized_cdf
 *   - compressai/cpp_exts/rans/rans_interface.cpp:99-204  encode_with_indexes + flush
 *   - compressai/cpp_exts/rans/rans_interface.cpp:206-275 decode_with_indexes
 *   - third_party/ryg_rans/rans64.h:59-142          Rans64 primitives
 * Pinned against the reference's own C++ (compiled into oracle/_ref by the
 * Makefile next to this file) in tests/test_oracle_coder.py, and against the
 * committed byte-stream fixtures in tests/golden/.  The product never links
 * or loads this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PRECISION 16
#define ORC_BYPASS_BITS 4
#define ORC_BYPASS_MAX ((1 << ORC_BYPASS_BITS) - 1)
#define ORC_RANS_L (1ull << 31)

/* ops.cpp:24-81.  cdf must hold n+1 entries.  returns 0, or -1 if no symbol can donate. */
int orc_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf) {
  int size = n + 1;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
  /* std::accumulate(..., 0): int accumulator, then stored to uint32 */
  int acc = 0;
  for (int i = 0; i < size; ++i) acc = (int)((uint32_t)acc + cdf[i]);
  uint32_t total = (uint32_t)acc;
  for (int i = 0; i < size; ++i) cdf[i] = (uint32_t)((((uint64_t)(1 << precision)) * cdf[i]) / total);
  for (int i = 1; i < size; ++i) cdf[i] += cdf[i - 1];
  cdf[size - 1] = 1u << precision;
  for (int i = 0; i < size - 1; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best_freq = ~0u;
    int best = -1;
    for (int j = 0; j < size - 1; ++j) {
      uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best_freq) { best_freq = f; best = j; }
    }
    if (best < 0) return -1;
    if (best < i) {
      for (int j = best + 1; j <= i; ++j) cdf[j]--;
    } else {
      for (int j = i + 1; j <= best; ++j) cdf[j]++;
    }
  }
  return 0;
}

typedef struct { uint16_t start, range; uint8_t bypass; } orc_sym;

/* rans_interface.cpp:99-204.  cdfs is a dense [ncdf x pitch] int32 table.
 * out must hold at least 4*(2 + 9*n) bytes.  returns byte count. */
long orc_rans_encode(const int32_t *symbols, const int32_t *indexes, long n, const int32_t *cdfs, int pitch,
                     const int32_t *cdf_sizes, const int32_t *offsets, uint8_t *out) {
  long cap = 12 * n + 16, cnt = 0;
  orc_sym *syms = (orc_sym *)malloc(sizeof(orc_sym) * (size_t)cap);
  for (long i = 0; i < n; ++i) {
    int ci = indexes[i];
    const int32_t *cdf = cdfs + (long)ci * pitch;
    int32_t max_value = cdf_sizes[ci] - 2;
    int32_t value = symbols[i] - offsets[ci];
    uint32_t raw = 0;
    if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
    else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
    syms[cnt++] = (orc_sym){(uint16_t)cdf[value], (uint16_t)(cdf[value + 1] - cdf[value]), 0};
    if (value == max_value) {
      int32_t nb = 0;
      while ((raw >> (nb * ORC_BYPASS_BITS)) != 0) ++nb;
      int32_t v = nb;
      while (v >= ORC_BYPASS_MAX) { syms[cnt++] = (orc_sym){ORC_BYPASS_MAX, ORC_BYPASS_MAX + 1, 1}; v -= ORC_BYPASS_MAX; }
      syms[cnt++] = (orc_sym){(uint16_t)v, (uint16_t)(v + 1), 1};
      for (int32_t j = 0; j < nb; ++j) {
        int32_t d = (raw >> (j * ORC_BYPASS_BITS)) & ORC_BYPASS_MAX;
        syms[cnt++] = (orc_sym){(uint16_t)d, (uint16_t)(d + 1), 1};
      }
    }
  }
  long words = cnt + 2;
  uint32_t *buf = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)words);
  uint32_t *ptr = buf + words;
  uint64_t x = ORC_RANS_L; /* rans64.h:66-69 */
  for (long i = cnt - 1; i >= 0; --i) {
    orc_sym s = syms[i];
    if (!s.bypass) { /* rans64.h:77-93 */
      uint64_t x_max = ((ORC_RANS_L >> ORC_PRECISION) << 32) * s.range;
      if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
      x = ((x / s.range) << ORC_PRECISION) + (x % s.range) + s.start;
    } else { /* rans_interface.cpp:60-77 */
      uint32_t freq = 1u << (16 - ORC_BYPASS_BITS);
      uint64_t x_max = ((ORC_RANS_L >> 16) << 32) * freq;
      if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
      x = (x << ORC_BYPASS_BITS) | s.start;
    }
  }
  ptr -= 2; /* rans64.h:96-103 */
  ptr[0] = (uint32_t)x;
  ptr[1] = (uint32_t)(x >> 32);
  long nbytes = (long)((buf + words) - ptr) * 4;
  memcpy(out, ptr, (size_t)nbytes);
  free(buf);
  free(syms);
  return nbytes;
}

static uint32_t orc_get_bits(uint64_t *x, const uint32_t **pp, uint32_t nbits) { /* rans_interface.cpp:79-96 */
  uint32_t v = (uint32_t)(*x & ((1u << nbits) - 1));
  *x >>= nbits;
  if (*x < ORC_RANS_L) { *x = (*x << 32) | **pp; *pp += 1; }
  return v;
}

/* rans_interface.cpp:206-275 */
void orc_rans_decode(const uint8_t *stream, const int32_t *indexes, long n, const int32_t *cdfs, int pitch,
                     const int32_t *cdf_sizes, const int32_t *offsets, int32_t *out) {
  const uint32_t *ptr = (const uint32_t *)stream;
  uint64_t x = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
  ptr += 2;
  for (long i = 0; i < n; ++i) {
    int ci = indexes[i];
    const int32_t *cdf = cdfs + (long)ci * pitch;
    int32_t max_value = cdf_sizes[ci] - 2;
    uint32_t cum = (uint32_t)(x & ((1u << ORC_PRECISION) - 1));
    int s = 0;
    while (s < cdf_sizes[ci] && !((uint32_t)cdf[s] > cum)) ++s;
    s -= 1;
    uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
    x = (uint64_t)freq * (x >> ORC_PRECISION) + (x & ((1ull << ORC_PRECISION) - 1)) - start;
    if (x < ORC_RANS_L) { x = (x << 32) | *ptr++; }
    int32_t value = s;
    if (value == max_value) {
      int32_t v = (int32_t)orc_get_bits(&x, &ptr, ORC_BYPASS_BITS);
      int32_t nb = v;
      while (v == ORC_BYPASS_MAX) { v = (int32_t)orc_get_bits(&x, &ptr, ORC_BYPASS_BITS); nb += v; }
      int32_t raw = 0;
      for (int32_t j = 0; j < nb; ++j) raw |= (int32_t)orc_get_bits(&x, &ptr, ORC_BYPASS_BITS) << (j * ORC_BYPASS_BITS);
      value = raw >> 1;
      if (raw & 1) value = -value - 1; else value += max_value;
    }
    out[i] = value + offsets[ci];
  }
}
