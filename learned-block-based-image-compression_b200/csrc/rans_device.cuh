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

// ---------------------------------------------------------------------------------------------
// Lean warp-per-row decoder on SHARED-memory tables (the decode step of small batches / single images, where the
// serial symbol chain of a row is the latency of the whole step; rans.cu's rans_dec_step_smem_kernel and the rANS tiles
// of gemm_wave.cu).  A warp running alone issues one dependent instruction every ~5 cycles, so what counts is the
// NUMBER of instructions on the chain: everything that does not depend on the coder state is done before or beside it
//   * per channel, in parallel: CDF index -> one 64-bit record {row address + window start, length | window start,
//     symbol-0 position} in the warp's scratch;
//   * in the loop, one symbol ahead: the record (a broadcast 8-byte shared load) and the 32-wide window of the row;
//   * the next 64 words of the stream sit in registers (two per lane), fetched by shuffles.
// On the chain: slot = x & 0xFFFF, compare, vote, find-first-set, two shuffles (start, next), the 64-bit state
// update and the renormalisation test.  Same symbols as dec_symbol_warp: (first k with cdf[k] > slot) - 1.
// Tables: Tables::cdf16 (16-bit rows without their final 65536) and int[3][64] = row offsets, lengths, symbol offsets.
// ---------------------------------------------------------------------------------------------
struct DecCursorW {
    unsigned long long x;
    const uint32_t *words;
    uint32_t pos, nwords, base, w0, w1;    // lane i holds words base + i and base + 32 + i
};
__device__ __forceinline__ void dec_fill_w(DecCursorW &d, int lane) {
    d.base = d.pos;
    d.w0 = d.base + lane < d.nwords ? __ldg(d.words + d.base + lane) : 0u;
    d.w1 = d.base + 32 + lane < d.nwords ? __ldg(d.words + d.base + 32 + lane) : 0u;
}
__device__ __forceinline__ uint32_t dec_word_w(DecCursorW &d, int lane) {
    if (d.pos - d.base >= 64) dec_fill_w(d, lane);                 // warp-uniform
    const uint32_t rel = d.pos - d.base;
    const uint32_t v0 = __shfl_sync(0xffffffffu, d.w0, rel & 31), v1 = __shfl_sync(0xffffffffu, d.w1, rel & 31);
    d.pos++;
    return rel < 32 ? v0 : v1;                                      // past the end of the stream: zeros (a corrupt stream stays finite)
}
__device__ __forceinline__ int dec_bits_w(DecCursorW &d, int lane) {
    const int val = (int)(d.x & MAX_BYPASS);
    d.x >>= BYPASS;
    if (d.x < RANS_L) d.x = (d.x << 32) | dec_word_w(d, lane);
    return val;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
constexpr int RANS_ROW_SCRATCH(int M) { return 12 * M; }   // bytes of shared memory per decoding warp: records + symbols

// window value of lane `lane` for a record: entry k = g + lane of the row (65536 for the implicit last entry, 0 beyond
// the row so that it never compares greater than a slot)
__device__ __forceinline__ uint32_t rans_window(uint32_t s_cdf16, uint2 rec, int lane) {
    const int len = (int)(rec.x >> 16), g = (int)(rec.y & 0xFFFFu);
    const int k = g + lane;
    uint32_t v = k < len - 1 ? lds_u16(s_cdf16 + 2u * ((rec.x & 0xFFFFu) + (uint32_t)lane)) : 0u;
    v = k == len - 1 ? 65536u : v;
    return v;
}

// Decodes the M symbols of one row (scales -> CDF indexes from krow[0..M), coder state d).  s_cdf16 / s_meta / s_scr:
// 32-bit shared addresses of the rows, of int[3][64] {row offset, length, symbol offset} and of this warp's scratch
// (RANS_ROW_SCRATCH(M) bytes).  Leaves symbol c at s_scr + 8 M + 4 c; the caller reads it after __syncwarp().
__device__ __forceinline__ void rans_decode_row_warp(DecCursorW &d, uint32_t s_cdf16, uint32_t s_meta, uint32_t s_scr,
                                                     const float *__restrict__ krow, const float *stab, int M, int lane) {
    for (int c = lane; c < M; c += 32) {
        const int ci = scale_to_index(__ldcg(krow + c), stab);
        const int off16 = lds_s32(s_meta + 4u * ci), len = lds_s32(s_meta + 256u + 4u * ci);
        const int center = -lds_s32(s_meta + 512u + 4u * ci);
        int g = center - 15;
        g = g < 0 ? 0 : g;
        g = g > len - 32 ? (len - 32 < 0 ? 0 : len - 32) : g;
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(s_scr + 8u * c), "r"((uint32_t)(off16 + g) | ((uint32_t)len << 16)),
                     "r"((uint32_t)g | ((uint32_t)center << 16)) : "memory");
    }
    __syncwarp();
    uint2 rec = lds_u64(s_scr);
    uint32_t val = rans_window(s_cdf16, rec, lane);
#pragma unroll 1
    for (int c = 0; c < M; ++c) {
        // one symbol ahead, off the state chain
        const uint2 rec_n = lds_u64(s_scr + 8u * (c + 1 < M ? c + 1 : c));
        const uint32_t val_n = rans_window(s_cdf16, rec_n, lane);
        const int len = (int)(rec.x >> 16), g = (int)(rec.y & 0xFFFFu), center = (int)(rec.y >> 16);
        const int max_value = len - 2;
        const uint32_t cf = (uint32_t)(d.x & 0xFFFFu);
        const unsigned ball = __ballot_sync(0xffffffffu, val > cf);
        int sidx;
        uint32_t start, next;
        if ((ball & 1u) == 0 && ball != 0) {
            const int j = __ffs(ball) - 1;
            sidx = g + j - 1;
            start = __shfl_sync(0xffffffffu, val, j - 1);
            next = __shfl_sync(0xffffffffu, val, j);
        } else {
            // outside the window: upper_bound by bisection over the whole row (warp-uniform)
            const uint32_t row = s_cdf16 + 2u * ((rec.x & 0xFFFFu) - (uint32_t)g);
            int lo = 0, hi = len - 1;          // entry len-1 (= 65536) always exceeds the slot
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (lds_u16(row + 2u * mid) > cf) hi = mid; else lo = mid + 1;
            }
            sidx = lo - 1;
            start = lds_u16(row + 2u * sidx);
            next = sidx + 1 < len - 1 ? lds_u16(row + 2u * (sidx + 1)) : 65536u;
        }
        d.x = (unsigned long long)(next - start) * (d.x >> PREC) + cf - start;
        if (d.x < RANS_L) d.x = (d.x << 32) | dec_word_w(d, lane);
        int value = sidx;
        if (value == max_value) {
            int v = dec_bits_w(d, lane);
            int nb = v;
            while (v == MAX_BYPASS) {
                v = dec_bits_w(d, lane);
                nb += v;
            }
            int raw = 0;
            for (int j = 0; j < nb; ++j) {
                v = dec_bits_w(d, lane);
                raw |= v << (j * BYPASS);
            }
            value = raw >> 1;
            if (raw & 1) value = -value - 1; else value += max_value;
        }
        if (lane == 0) asm volatile("st.shared.s32 [%0], %1;" ::"r"(s_scr + 8u * M + 4u * c), "r"(value - center) : "memory");
        rec = rec_n;
        val = val_n;
    }
    __syncwarp();
}

}  // namespace
