/*
 * oracle/lbic_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * CPU restatement of the two native functions the reference's hot path calls through
 * pybind11 into CompressAI (a third-party dependency that is NOT vendored in the
 * reference tree and NOT version-pinned: README.md:24-31 says "clone CompressAI master,
 * pip install -e ."):
 *
 *   compressai._CXX.pmf_to_quantized_cdf           call sites  graphs/layers/entropy_layers_cai.py:13,61-64
 *   compressai.ans.BufferedRansEncoder / RansDecoder call sites graphs/models/BlockBasedImgCompLossy_net.py:9,328,359-360,409-410,439
 *
 * The arithmetic is the published CompressAI algorithm (cpp_exts/ops/ops.cpp and
 * cpp_exts/rans/rans_interface.cpp over ryg_rans' rans64.h: 64-bit state, RANS64_L = 2^31,
 * 32-bit renormalisation words, 16-bit probabilities, 4-bit bypass escape), restated here
 * from its public description.  PARITY UNPINNED for these two functions: the reference holds
 * no golden bitstreams or stored CDF tables and CompressAI cannot be installed offline, so
 * byte-level agreement with a real CompressAI build cannot be checked in this environment.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RANS64_L (1ull << 31)
#define PRECISION 16
#define BYPASS_PRECISION 4
#define MAX_BYPASS_VAL ((1 << BYPASS_PRECISION) - 1)

/* ---- pmf_to_quantized_cdf (entropy_layers_cai.py:61-64 -> compressai._CXX) ------------------
 * cdf has n+1 entries.  Returns 0 on success, <0 on invalid input. */
int oracle_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf)
{
    for (int i = 0; i < n; ++i)
        if (pmf[i] < 0.0f || !isfinite(pmf[i])) return -1;
    cdf[0] = 0;
    for (int i = 0; i < n; ++i) {
        /* std::round on a float product: half away from zero */
        float scaled = pmf[i] * (float)(1 << precision);
        cdf[i + 1] = (uint32_t)roundf(scaled);
    }
    uint32_t total = 0;
    for (int i = 0; i <= n; ++i) total += cdf[i];
    if (total == 0) return -2;
    for (int i = 0; i <= n; ++i)
        cdf[i] = (uint32_t)((((uint64_t)1 << precision) * (uint64_t)cdf[i]) / total);
    for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
    cdf[n] = 1u << precision;
    for (int i = 0; i < n; ++i) {
        if (cdf[i] == cdf[i + 1]) {
            uint32_t best_freq = ~0u;
            int best_steal = -1;
            for (int j = 0; j < n; ++j) {
                uint32_t freq = cdf[j + 1] - cdf[j];
                if (freq > 1 && freq < best_freq) { best_freq = freq; best_steal = j; }
            }
            if (best_steal < 0) return -3;
            if (best_steal < i) {
                for (int j = best_steal + 1; j <= i; ++j) cdf[j]--;
            } else {
                for (int j = i + 1; j <= best_steal; ++j) cdf[j]++;
            }
        }
    }
    return 0;
}

/* ---- rANS64 encoder (BlockBasedImgCompLossy_net.py:359-360) -------------------------------- */
typedef struct { uint16_t start; uint16_t range; uint8_t bypass; } rsym_t;

/* Encodes n (symbol, index) pairs into a single-state rANS64 stream.
 * cdf: row-major [n_tables][cdf_stride].  Returns the byte count (>= 8), or <0 on error.
 * The bytes are written to out[0..ret).  */
long oracle_rans_encode(const int32_t *symbols, const int32_t *indexes, long n,
                        const int32_t *cdf, int cdf_stride, const int32_t *cdf_sizes,
                        const int32_t *offsets, uint8_t *out, long out_cap)
{
    long cap = n * 2 + 16, cnt = 0;
    rsym_t *syms = (rsym_t *)malloc((size_t)cap * sizeof(rsym_t));
    if (!syms) return -1;
    for (long i = 0; i < n; ++i) {
        const int32_t ci = indexes[i];
        const int32_t *c = cdf + (long)ci * cdf_stride;
        const int32_t max_value = cdf_sizes[ci] - 2;
        int32_t value = symbols[i] - offsets[ci];
        uint32_t raw_val = 0;
        if (value < 0) {
            raw_val = (uint32_t)(-2 * value - 1);
            value = max_value;
        } else if (value >= max_value) {
            raw_val = (uint32_t)(2 * (value - max_value));
            value = max_value;
        }
        if (cnt + 24 > cap) {
            cap *= 2;
            syms = (rsym_t *)realloc(syms, (size_t)cap * sizeof(rsym_t));
            if (!syms) return -1;
        }
        syms[cnt].start = (uint16_t)c[value];
        syms[cnt].range = (uint16_t)(c[value + 1] - c[value]);
        syms[cnt].bypass = 0;
        cnt++;
        if (value == max_value) {
            int32_t n_bypass = 0;
            while (n_bypass < 8 && (raw_val >> (n_bypass * BYPASS_PRECISION)) != 0) ++n_bypass;
            int32_t val = n_bypass;
            while (val >= MAX_BYPASS_VAL) {
                syms[cnt].start = MAX_BYPASS_VAL; syms[cnt].range = MAX_BYPASS_VAL + 1; syms[cnt].bypass = 1; cnt++;
                val -= MAX_BYPASS_VAL;
            }
            syms[cnt].start = (uint16_t)val; syms[cnt].range = (uint16_t)(val + 1); syms[cnt].bypass = 1; cnt++;
            for (int32_t j = 0; j < n_bypass; ++j) {
                int32_t v = (raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL;
                syms[cnt].start = (uint16_t)v; syms[cnt].range = (uint16_t)(v + 1); syms[cnt].bypass = 1; cnt++;
            }
        }
    }
    /* flush(): walk the pushed symbols backwards, words are written from the buffer end */
    long nwords = cnt + 4;
    uint32_t *buf = (uint32_t *)malloc((size_t)nwords * sizeof(uint32_t));
    if (!buf) { free(syms); return -1; }
    uint32_t *ptr = buf + nwords;
    uint64_t x = RANS64_L;
    for (long k = cnt - 1; k >= 0; --k) {
        const rsym_t s = syms[k];
        if (!s.bypass) {
            uint64_t x_max = ((RANS64_L >> PRECISION) << 32) * (uint64_t)s.range;
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = ((x / s.range) << PRECISION) + (x % s.range) + s.start;
        } else {
            uint32_t freq = 1u << (PRECISION - BYPASS_PRECISION);
            uint64_t x_max = ((RANS64_L >> PRECISION) << 32) * (uint64_t)freq;
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = (x << BYPASS_PRECISION) | s.start;
        }
    }
    ptr -= 2;
    ptr[0] = (uint32_t)(x >> 0);
    ptr[1] = (uint32_t)(x >> 32);
    long nbytes = (long)((buf + nwords) - ptr) * 4;
    long ret = nbytes;
    if (nbytes > out_cap) ret = -2; else memcpy(out, ptr, (size_t)nbytes);
    free(buf);
    free(syms);
    return ret;
}

/* ---- rANS64 decoder (BlockBasedImgCompLossy_net.py:409-410,439) ---------------------------- */
typedef struct {
    uint64_t x;
    const uint32_t *ptr;
    const uint32_t *end;
} oracle_rans_dec;

static inline uint32_t dec_next_word(oracle_rans_dec *d)
{
    /* A well-formed stream never reads past its end; reading zeros keeps a corrupt one finite. */
    if (d->ptr < d->end) return *d->ptr++;
    d->ptr++;
    return 0;
}

void oracle_rans_dec_init(oracle_rans_dec *d, const uint8_t *stream, long nbytes)
{
    d->ptr = (const uint32_t *)stream;
    d->end = d->ptr + nbytes / 4;
    uint64_t lo = dec_next_word(d);
    uint64_t hi = dec_next_word(d);
    d->x = lo | (hi << 32);
}

static inline int32_t dec_get_bits(oracle_rans_dec *d, int nbits)
{
    uint64_t x = d->x;
    int32_t val = (int32_t)(x & ((1u << nbits) - 1));
    x >>= nbits;
    if (x < RANS64_L) x = (x << 32) | dec_next_word(d);
    d->x = x;
    return val;
}

/* decode_stream(): n symbols with the given indexes, continuing from the decoder state */
int oracle_rans_dec_stream(oracle_rans_dec *d, const int32_t *indexes, long n,
                           const int32_t *cdf, int cdf_stride, const int32_t *cdf_sizes,
                           const int32_t *offsets, int32_t *out)
{
    for (long i = 0; i < n; ++i) {
        const int32_t ci = indexes[i];
        const int32_t *c = cdf + (long)ci * cdf_stride;
        const int32_t max_value = cdf_sizes[ci] - 2;
        const int32_t offset = offsets[ci];
        const uint32_t cum_freq = (uint32_t)(d->x & ((1u << PRECISION) - 1));
        int32_t k = 0;
        const int32_t sz = cdf_sizes[ci];
        while (k < sz && !((uint32_t)c[k] > cum_freq)) ++k; /* linear find_if */
        const int32_t s = k - 1;
        {
            uint64_t x = d->x;
            uint32_t start = (uint32_t)c[s], freq = (uint32_t)(c[s + 1] - c[s]);
            x = (uint64_t)freq * (x >> PRECISION) + (x & ((1u << PRECISION) - 1)) - start;
            if (x < RANS64_L) x = (x << 32) | dec_next_word(d);
            d->x = x;
        }
        int32_t value = s;
        if (value == max_value) {
            int32_t val = dec_get_bits(d, BYPASS_PRECISION);
            int32_t n_bypass = val;
            while (val == MAX_BYPASS_VAL) {
                val = dec_get_bits(d, BYPASS_PRECISION);
                n_bypass += val;
            }
            int32_t raw_val = 0;
            for (int32_t j = 0; j < n_bypass; ++j) {
                val = dec_get_bits(d, BYPASS_PRECISION);
                raw_val |= val << (j * BYPASS_PRECISION);
            }
            value = raw_val >> 1;
            if (raw_val & 1) value = -value - 1; else value += max_value;
        }
        out[i] = value + offset;
    }
    return 0;
}

/* One-shot decode of a whole stream. */
int oracle_rans_decode(const uint8_t *stream, long nbytes, const int32_t *indexes, long n,
                       const int32_t *cdf, int cdf_stride, const int32_t *cdf_sizes,
                       const int32_t *offsets, int32_t *out)
{
    oracle_rans_dec d;
    oracle_rans_dec_init(&d, stream, nbytes);
    return oracle_rans_dec_stream(&d, indexes, n, cdf, cdf_stride, cdf_sizes, offsets, out);
}

/* number of stream bytes consumed so far by a decoder (for tests) */
long oracle_rans_dec_consumed(const oracle_rans_dec *d, const uint8_t *stream)
{
    return (long)((const uint8_t *)d->ptr - stream);
}
