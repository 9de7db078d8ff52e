// Device-side rANS64 decoder pieces shared by rans.cu (stand-alone decode-step kernels) and gemm_wave.cu (decode step
// fused into the persistent wavefront kernel).  Bit-compatible with CompressAI's RansDecoder (see rans.cu).
#pragma once
#include "epilogue.cuh"

namespace {

constexpr unsigned long long RANS_L = 1ull << 31;
constexpr int PREC = 16;
constexpr int BYPASS = 4;
constexpr int MAX_BYPASS = 15;

// ---------------------------------------------------------------------------------------------
// decoder
// ---------------------------------------------------------------------------------------------
struct DecCursor {
    unsigned long long x;
    const uint32_t *words;
    uint32_t pos, nwords;
};

__device__ __forceinline__ uint32_t dec_word(DecCursor &d) {
    const uint32_t w = d.pos < d.nwords ? __ldg(d.words + d.pos) : 0u;   // a corrupt stream stays finite
    d.pos++;
    return w;
}

__device__ __forceinline__ int dec_bits(DecCursor &d) {
    const int val = (int)(d.x & MAX_BYPASS);
    d.x >>= BYPASS;
    if (d.x < RANS_L) d.x = (d.x << 32) | dec_word(d);
    return val;
}

// Decodes one symbol; warp-cooperative CDF search (all 32 lanes hold identical cursor state).
__device__ __forceinline__ int dec_symbol_warp(DecCursor &d, const int32_t *__restrict__ row, int len, int off,
                                               int lane) {
    const uint32_t cf = (uint32_t)(d.x & 0xFFFFu);
    const int max_value = len - 2;
    // first k with row[k] > cf, searched in a 32-wide window around the distribution centre (value of symbol 0)
    int g = -off - 15;
    g = g < 0 ? 0 : g;
    g = g > len - 32 ? (len - 32 < 0 ? 0 : len - 32) : g;
    const int k = g + lane;
    const bool gt = (k < len) && ((uint32_t)__ldg(row + k) > cf);
    const unsigned ball = __ballot_sync(0xffffffffu, gt);
    int s;
    if ((ball & 1u) == 0 && ball != 0) {
        s = g + (__ffs(ball) - 1) - 1;
    } else {
        // outside the window: upper_bound by bisection (identical result to the reference's linear find_if)
        int lo = 0, hi = len - 1;   // row[len-1] = 65536 > cf always
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((uint32_t)__ldg(row + mid) > cf) hi = mid; else lo = mid + 1;
        }
        s = lo - 1;
    }
    const uint32_t start = (uint32_t)__ldg(row + s);
    const uint32_t freq = (uint32_t)__ldg(row + s + 1) - start;
    d.x = (unsigned long long)freq * (d.x >> PREC) + cf - start;
    if (d.x < RANS_L) d.x = (d.x << 32) | dec_word(d);
    int value = s;
    if (value == max_value) {
        int val = dec_bits(d);
        int nb = val;
        while (val == MAX_BYPASS) {
            val = dec_bits(d);
            nb += val;
        }
        int raw = 0;
        for (int j = 0; j < nb; ++j) {
            val = dec_bits(d);
            raw |= val << (j * BYPASS);
        }
        value = raw >> 1;
        if (raw & 1) value = -value - 1; else value += max_value;
    }
    return value + off;
}

}  // namespace
