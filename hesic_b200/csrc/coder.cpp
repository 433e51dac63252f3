// Host-side entropy coder of the drop-in boundary: the `compressai.ans` and `compressai._CXX`
// surface (compressai/cpp_exts/rans/rans_interface.cpp:99-350, compressai/cpp_exts/ops/ops.cpp:24-81),
// re-implemented on flat arrays so the device-side symbol/index preparation kernels can feed it
// without Python-list marshalling.  Bit-stream format: 64-bit rANS state, 32-bit little-endian word
// renormalisation, 16-bit probabilities, 4-bit bypass digits for out-of-table symbols.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/hesic_b200.h"

namespace hesic {
void set_error(const char *fmt, ...);
}
using hesic::set_error;

namespace {
constexpr int kPrecision = 16;
constexpr int kBypassBits = 4;
constexpr uint32_t kBypassMax = (1u << kBypassBits) - 1;
constexpr uint64_t kLow = 1ull << 31;  // lower bound of the normalised state interval

struct Tok {  // one coding step: a table interval or a raw 4-bit digit
  uint16_t start, width;
  uint16_t raw;
};

inline void put_interval(uint64_t &x, std::vector<uint32_t> &rev, uint32_t start, uint32_t width, int bits) {
  const uint64_t limit = ((kLow >> bits) << 32) * width;
  if (x >= limit) {
    rev.push_back(static_cast<uint32_t>(x));
    x >>= 32;
  }
  x = ((x / width) << bits) + (x % width) + start;
}
}  // namespace

struct hesic_rans_encoder {
  std::vector<Tok> toks;
};

struct hesic_rans_decoder {
  std::vector<uint32_t> words;
  size_t pos = 0;
  uint64_t x = 0;
  bool ready = false;
  uint32_t next() { return pos < words.size() ? words[pos++] : 0u; }
  uint32_t digits(int bits) {
    uint32_t v = static_cast<uint32_t>(x & ((1u << bits) - 1));
    x >>= bits;
    if (x < kLow) x = (x << 32) | next();
    return v;
  }
};

extern "C" int hesic_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf) {
  if (!pmf || !cdf || n < 1 || precision < 1 || precision > 16) {
    set_error("pmf_to_quantized_cdf: invalid argument");
    return HESIC_E_INVALID;
  }
  const int len = n + 1;
  const float scale = static_cast<float>(1 << precision);
  cdf[0] = 0;
  uint32_t total = 0;
  for (int i = 0; i < n; ++i) {
    cdf[i + 1] = static_cast<uint32_t>(std::round(pmf[i] * scale));
    total += cdf[i + 1];
  }
  if (total == 0) {
    set_error("pmf_to_quantized_cdf: pmf sums to zero");
    return HESIC_E_INVALID;
  }
  // renormalise so the frequencies sum to <= 2^precision, then prefix-sum
  uint32_t run = 0;
  for (int i = 0; i < len; ++i) {
    run += static_cast<uint32_t>((static_cast<uint64_t>(1u << precision) * cdf[i]) / total);
    cdf[i] = run;
  }
  cdf[len - 1] = 1u << precision;
  // every symbol needs a non-empty interval: take one count from the smallest donor with freq > 1
  for (int i = 0; i + 1 < len; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    int donor = -1;
    uint32_t donor_freq = ~0u;
    for (int j = 0; j + 1 < len; ++j) {
      const uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < donor_freq) { donor_freq = f; donor = j; }
    }
    if (donor < 0) {
      set_error("pmf_to_quantized_cdf: no symbol can donate frequency");
      return HESIC_E_INVALID;
    }
    if (donor < i) {
      for (int j = donor + 1; j <= i; ++j) --cdf[j];
    } else {
      for (int j = i + 1; j <= donor; ++j) ++cdf[j];
    }
  }
  return HESIC_OK;
}

extern "C" hesic_rans_encoder *hesic_rans_encoder_create(void) { return new hesic_rans_encoder(); }
extern "C" void hesic_rans_encoder_destroy(hesic_rans_encoder *e) { delete e; }

static int check_tables(const int32_t *cdfs, int n_cdfs, int pitch, const int32_t *sizes, const int32_t *offsets) {
  if (!cdfs || !sizes || !offsets || n_cdfs < 1 || pitch < 2) {
    set_error("rans: invalid CDF tables");
    return HESIC_E_INVALID;
  }
  for (int i = 0; i < n_cdfs; ++i)
    if (sizes[i] < 2 || sizes[i] > pitch) {
      set_error("rans: cdf %d has size %d (pitch %d)", i, sizes[i], pitch);
      return HESIC_E_INVALID;
    }
  return HESIC_OK;
}

extern "C" int hesic_rans_encoder_push(hesic_rans_encoder *e, const int32_t *symbols, const int32_t *indexes, int64_t n,
                                       const int32_t *cdfs, int n_cdfs, int pitch, const int32_t *sizes,
                                       const int32_t *offsets) {
  if (!e || n < 0 || (n > 0 && (!symbols || !indexes))) {
    set_error("rans encode: invalid argument");
    return HESIC_E_INVALID;
  }
  int r = check_tables(cdfs, n_cdfs, pitch, sizes, offsets);
  if (r != HESIC_OK) return r;
  e->toks.reserve(e->toks.size() + static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) {
    const int t = indexes[i];
    if (t < 0 || t >= n_cdfs) {
      set_error("rans encode: index %d out of range at position %lld", t, (long long)i);
      return HESIC_E_INVALID;
    }
    const int32_t *cdf = cdfs + static_cast<size_t>(t) * pitch;
    const int32_t escape = sizes[t] - 2;  // last table slot = "value follows in bypass digits"
    int32_t v = symbols[i] - offsets[t];
    uint32_t raw = 0;
    if (v < 0) {
      raw = static_cast<uint32_t>(-2 * v - 1);
      v = escape;
    } else if (v >= escape) {
      raw = static_cast<uint32_t>(2 * (v - escape));
      v = escape;
    }
    e->toks.push_back({static_cast<uint16_t>(cdf[v]), static_cast<uint16_t>(cdf[v + 1] - cdf[v]), 0});
    if (v == escape) {
      int nd = 0;
      while ((raw >> (nd * kBypassBits)) != 0) ++nd;
      int left = nd;  // digit count, unary-ish in base 15
      while (left >= (int)kBypassMax) {
        e->toks.push_back({static_cast<uint16_t>(kBypassMax), 0, 1});
        left -= kBypassMax;
      }
      e->toks.push_back({static_cast<uint16_t>(left), 0, 1});
      for (int j = 0; j < nd; ++j)
        e->toks.push_back({static_cast<uint16_t>((raw >> (j * kBypassBits)) & kBypassMax), 0, 1});
    }
  }
  return HESIC_OK;
}

extern "C" int64_t hesic_rans_encoder_flush(hesic_rans_encoder *e, uint8_t *out, int64_t out_cap) {
  if (!e) {
    set_error("rans flush: null encoder");
    return HESIC_E_INVALID;
  }
  // rANS is last-in first-out: walk the tokens backwards, emit words in reverse
  std::vector<uint32_t> rev;
  rev.reserve(e->toks.size() / 2 + 4);
  uint64_t x = kLow;
  for (size_t i = e->toks.size(); i-- > 0;) {
    const Tok &t = e->toks[i];
    if (!t.raw) {
      put_interval(x, rev, t.start, t.width, kPrecision);
    } else {
      const uint64_t limit = ((kLow >> 16) << 32) * (1u << (16 - kBypassBits));
      if (x >= limit) {
        rev.push_back(static_cast<uint32_t>(x));
        x >>= 32;
      }
      x = (x << kBypassBits) | t.start;
    }
  }
  rev.push_back(static_cast<uint32_t>(x >> 32));
  rev.push_back(static_cast<uint32_t>(x));
  const int64_t nbytes = static_cast<int64_t>(rev.size()) * 4;
  if (!out || out_cap < nbytes) return nbytes;  // caller retries with a larger buffer; nothing consumed
  uint32_t *w = reinterpret_cast<uint32_t *>(out);
  for (size_t i = 0; i < rev.size(); ++i) {
    uint32_t v = rev[rev.size() - 1 - i];
    std::memcpy(w + i, &v, 4);
  }
  e->toks.clear();
  return nbytes;
}

extern "C" hesic_rans_decoder *hesic_rans_decoder_create(void) { return new hesic_rans_decoder(); }
extern "C" void hesic_rans_decoder_destroy(hesic_rans_decoder *d) { delete d; }

extern "C" int hesic_rans_decoder_set_stream(hesic_rans_decoder *d, const uint8_t *stream, int64_t nbytes) {
  if (!d || !stream || nbytes < 8 || (nbytes & 3)) {
    set_error("rans decode: stream must be a multiple of 4 bytes and at least 8");
    return HESIC_E_INVALID;
  }
  d->words.resize(static_cast<size_t>(nbytes / 4));
  std::memcpy(d->words.data(), stream, static_cast<size_t>(nbytes));
  d->x = static_cast<uint64_t>(d->words[0]) | (static_cast<uint64_t>(d->words[1]) << 32);
  d->pos = 2;
  d->ready = true;
  return HESIC_OK;
}

extern "C" int hesic_rans_decoder_decode(hesic_rans_decoder *d, const int32_t *indexes, int64_t n, const int32_t *cdfs,
                                         int n_cdfs, int pitch, const int32_t *sizes, const int32_t *offsets,
                                         int32_t *out) {
  if (!d || !d->ready || n < 0 || (n > 0 && (!indexes || !out))) {
    set_error("rans decode: invalid argument or no stream set");
    return HESIC_E_INVALID;
  }
  int r = check_tables(cdfs, n_cdfs, pitch, sizes, offsets);
  if (r != HESIC_OK) return r;
  const uint32_t mask = (1u << kPrecision) - 1;
  for (int64_t i = 0; i < n; ++i) {
    const int t = indexes[i];
    if (t < 0 || t >= n_cdfs) {
      set_error("rans decode: index %d out of range at position %lld", t, (long long)i);
      return HESIC_E_INVALID;
    }
    const int32_t *cdf = cdfs + static_cast<size_t>(t) * pitch;
    const int32_t escape = sizes[t] - 2;
    const uint32_t target = static_cast<uint32_t>(d->x & mask);
    // first entry strictly greater than target, minus one (tables are short: linear/upper_bound both fine)
    const int32_t *hit = std::upper_bound(cdf, cdf + sizes[t], static_cast<int32_t>(target));
    int32_t s = static_cast<int32_t>(hit - cdf) - 1;
    if (s < 0) s = 0;
    if (s > escape) s = escape;
    const uint32_t start = static_cast<uint32_t>(cdf[s]), width = static_cast<uint32_t>(cdf[s + 1] - cdf[s]);
    d->x = static_cast<uint64_t>(width) * (d->x >> kPrecision) + (d->x & mask) - start;
    if (d->x < kLow) d->x = (d->x << 32) | d->next();
    int32_t v = s;
    if (s == escape) {
      uint32_t dg = d->digits(kBypassBits);
      int32_t nd = static_cast<int32_t>(dg);
      while (dg == kBypassMax) {
        dg = d->digits(kBypassBits);
        nd += static_cast<int32_t>(dg);
      }
      int32_t raw = 0;
      for (int32_t j = 0; j < nd; ++j) raw |= static_cast<int32_t>(d->digits(kBypassBits)) << (j * kBypassBits);
      v = raw >> 1;
      v = (raw & 1) ? -v - 1 : v + escape;
    }
    out[i] = v + offsets[t];
  }
  return HESIC_OK;
}

// ---------------------------------------------------------------------------------------------
// Range coder for the file codec of the stereo models (HSIC.compress / decompress, newnet1.py:823-1273; SURVEY.md
// 8f rank 2).  The reference calls the un-vendored, un-pinned PyPI `range_coder` there
// (RangeEncoder(path).encode([symbol], cdf) once per latent element, newnet1.py:983-1040), so the byte stream of
// that package cannot be pinned; this is a self-contained carry-less range coder (64-bit low, 48-bit minimum
// range, byte renormalisation) with the same calling pattern: one symbol per call against an arbitrary integer
// cumulative-frequency row whose total need not be a power of two (the reference's rows sum to 65536 only
// approximately, newnet1.py:974-977).
namespace {
constexpr uint64_t kTop = 1ull << 56, kBot = 1ull << 48;
}

struct hesic_range_encoder {
  uint64_t low = 0, range = ~0ull;
  std::vector<uint8_t> out;
  void put(uint32_t cum, uint32_t freq, uint32_t total) {
    range /= total;
    low += cum * range;
    range *= freq;
    while ((low ^ (low + range)) < kTop || (range < kBot && ((range = (0 - low) & (kBot - 1)), true))) {
      out.push_back(static_cast<uint8_t>(low >> 56));
      low <<= 8;
      range <<= 8;
    }
  }
};

struct hesic_range_decoder {
  uint64_t low = 0, range = ~0ull, code = 0;
  std::vector<uint8_t> in;
  size_t pos = 0;
  uint8_t next() { return pos < in.size() ? in[pos++] : 0; }
};

extern "C" hesic_range_encoder *hesic_range_encoder_create(void) { return new hesic_range_encoder(); }
extern "C" void hesic_range_encoder_destroy(hesic_range_encoder *e) { delete e; }

extern "C" int hesic_range_encoder_push(hesic_range_encoder *e, const int32_t *symbols, int64_t n, const int32_t *cdfs,
                                        int cdf_pitch, int cdf_len) {
  if (!e || n < 0 || (n > 0 && (!symbols || !cdfs)) || cdf_len < 2 || cdf_pitch < cdf_len) {
    set_error("range_encoder_push: invalid argument");
    return HESIC_E_INVALID;
  }
  for (int64_t i = 0; i < n; ++i) {
    const int32_t *row = cdfs + (size_t)i * cdf_pitch;
    const int32_t s = symbols[i];
    if (s < 0 || s >= cdf_len - 1 || row[s + 1] <= row[s] || row[0] != 0 || row[cdf_len - 1] <= 0) {
      set_error("range_encoder_push: symbol %d of element %lld has no probability mass (row length %d)", (int)s, (long long)i,
                cdf_len);
      return HESIC_E_INVALID;
    }
    e->put((uint32_t)row[s], (uint32_t)(row[s + 1] - row[s]), (uint32_t)row[cdf_len - 1]);
  }
  return HESIC_OK;
}

extern "C" int64_t hesic_range_encoder_finish(hesic_range_encoder *e, uint8_t *out, int64_t out_cap) {
  if (!e) { set_error("range_encoder_finish: null encoder"); return HESIC_E_INVALID; }
  const int64_t need = (int64_t)e->out.size() + 8;
  if (!out || out_cap < need) return need;
  std::copy(e->out.begin(), e->out.end(), out);
  uint64_t low = e->low;
  for (int i = 0; i < 8; ++i) { out[e->out.size() + i] = static_cast<uint8_t>(low >> 56); low <<= 8; }
  e->out.clear();
  e->low = 0;
  e->range = ~0ull;
  return need;
}

extern "C" hesic_range_decoder *hesic_range_decoder_create(const uint8_t *stream, int64_t nbytes) {
  if (nbytes < 0 || (nbytes > 0 && !stream)) { set_error("range_decoder_create: invalid stream"); return nullptr; }
  hesic_range_decoder *d = new hesic_range_decoder();
  d->in.assign(stream, stream + nbytes);
  for (int i = 0; i < 8; ++i) d->code = (d->code << 8) | d->next();
  return d;
}
extern "C" void hesic_range_decoder_destroy(hesic_range_decoder *d) { delete d; }

extern "C" int hesic_range_decoder_decode(hesic_range_decoder *d, int64_t n, const int32_t *cdfs, int cdf_pitch, int cdf_len,
                                          int32_t *out_symbols) {
  if (!d || n < 0 || (n > 0 && (!cdfs || !out_symbols)) || cdf_len < 2 || cdf_pitch < cdf_len) {
    set_error("range_decoder_decode: invalid argument");
    return HESIC_E_INVALID;
  }
  for (int64_t i = 0; i < n; ++i) {
    const int32_t *row = cdfs + (size_t)i * cdf_pitch;
    const uint32_t total = (uint32_t)row[cdf_len - 1];
    if (row[0] != 0 || row[cdf_len - 1] <= 0) { set_error("range_decoder_decode: malformed cumulative row"); return HESIC_E_INVALID; }
    d->range /= total;
    uint64_t v = (d->code - d->low) / d->range;
    if (v >= total) v = total - 1;                 // corrupt stream: stay inside the table
    // last index with row[idx] <= v
    const int32_t *hi = std::upper_bound(row, row + cdf_len, (int32_t)v);
    int s = (int)(hi - row) - 1;
    if (s > cdf_len - 2) s = cdf_len - 2;
    while (s > 0 && row[s + 1] <= row[s]) --s;     // never land on an empty interval
    out_symbols[i] = s;
    d->low += (uint64_t)(uint32_t)row[s] * d->range;
    d->range *= (uint32_t)(row[s + 1] - row[s]);
    while ((d->low ^ (d->low + d->range)) < kTop || (d->range < kBot && ((d->range = (0 - d->low) & (kBot - 1)), true))) {
      d->code = (d->code << 8) | d->next();
      d->low <<= 8;
      d->range <<= 8;
    }
  }
  return HESIC_OK;
}
