// Fused layer epilogues, shared by the tcgen05 core (gemm_tc.cu) and its SIMT twin (gemm_simt.cu).
// Each call handles NV consecutive output columns [c, c+NV) of one row r (NV in {4,16}); the
// caller guarantees r < p.R and c + NV <= p.cout (all channel counts are multiples of 16).
#pragma once
#include "lbic_internal.h"

#define LBIC_SCALES_MIN 0.11f   // NET:13, ENT:553 LowerBound(scale_bound)

__device__ __forceinline__ void split_bf16(float v, bf16 &hi, bf16 &lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// build_indexes (ENT:649-654): idx = 63 - #{i<63 : max(s,0.11) <= T[i]} = #{i<63 : T[i] < s'}
__device__ __forceinline__ int scale_to_index(float scale, const float *__restrict__ tab) {
    const float s = fmaxf(scale, LBIC_SCALES_MIN);
    int lo = 0, hi = 63;   // first i in [0,63] with T[i] >= s  (63 = none)
#pragma unroll
    for (int it = 0; it < 6; ++it) {
        const int mid = (lo + hi) >> 1;
        const bool lt = (mid < 63) && (__ldg(tab + mid) < s);
        lo = lt ? mid + 1 : lo;
        hi = lt ? hi : mid;
    }
    return lo;
}

template <int NV>
__device__ __forceinline__ void store_hilo(bf16 *__restrict__ ph, bf16 *__restrict__ pl, const float (&v)[NV]) {
    static_assert(NV % 4 == 0, "NV");
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
        bf16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16(v[i + j], h[j], l[j]);
        uint2 uh, ul;
        uh.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
        uh.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
        ul.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
        ul.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
        *reinterpret_cast<uint2 *>(ph + i) = uh;
        *reinterpret_cast<uint2 *>(pl + i) = ul;
    }
}

template <int NV>
__device__ __forceinline__ void store_f32(float *__restrict__ p, const float (&v)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; i += 4)
        *reinterpret_cast<float4 *>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

template <int NV>
__device__ __forceinline__ void load_f32(const float *__restrict__ p, float (&v)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
        const float4 t = *reinterpret_cast<const float4 *>(p + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
}

// Operands an epilogue needs besides the accumulator, fetched ahead of the TMEM load they are combined with
// (the tcgen05 core software-pipelines: prefetch chunk i+1 while chunk i is being finished).
template <int NV>
struct EpiPre {
    float b[NV];    // bias / beta
    float a[NV];    // pre-GDN activation (GDN modes) or predicted scale (QUANT)
    float a2[NV];   // predicted mean (QUANT)
};

// bias: pointer to the bias of column c (global memory, or the CTA's shared-memory copy of its tile slice)
template <int NV>
__device__ __forceinline__ void epi_prefetch(const EpiParams &p, const float *__restrict__ bias, int r, int c,
                                             EpiPre<NV> &pre) {
    if (p.mode == EPI_RAW) return;
    load_f32<NV>(bias, pre.b);
    if (p.mode == EPI_GDN || p.mode == EPI_IGDN) {
        load_f32<NV>(p.aux + (size_t)r * p.ld_aux + c, pre.a);
    } else if (p.mode == EPI_QUANT) {
        load_f32<NV>(p.aux + (size_t)r * p.ld_aux + c, pre.a);            // scales = ksi[:, :M]   (NET:369)
        load_f32<NV>(p.aux + (size_t)r * p.ld_aux + p.M + c, pre.a2);     // means  = ksi[:, M:]
    }
}

template <int NV>
__device__ __forceinline__ void epi_apply(const EpiParams &p, int r, int c, const float (&acc)[NV],
                                          const EpiPre<NV> &pre) {
    float v[NV];
    if (p.mode == EPI_RAW) {
        store_f32<NV>(p.out_f32 + (size_t)r * p.ld_f32 + c, acc);
        return;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = acc[i] + pre.b[i];

    switch (p.mode) {
    case EPI_LRELU: {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = v[i] > 0.0f ? v[i] : v[i] * 0.01f;   // nn.LeakyReLU() default slope
        size_t orow = (size_t)r;
        if (p.out_pos) {
            int img, bv, bh;
            step_row_to_block(p.step, r, img, bv, bh);
            orow = g0_pos_index(img, bv, bh, p.step.Hb, p.step.Wb);
        }
        store_hilo<NV>(p.out_hi + orow * p.ld_out + c, p.out_lo + orow * p.ld_out + c, v);
    } break;
    case EPI_PREGDN: {
        store_f32<NV>(p.out_f32 + (size_t)r * p.ld_f32 + c, v);
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = v[i] * v[i];
        store_hilo<NV>(p.out_hi + (size_t)r * p.ld_out + c, p.out_lo + (size_t)r * p.ld_out + c, v);
    } break;
    case EPI_GDN:
    case EPI_IGDN: {
        const bool inv = (p.mode == EPI_IGDN);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float sq = __fsqrt_rn(v[i]);
            v[i] = pre.a[i] * (inv ? sq : __fdiv_rn(1.0f, sq));   // torch.sqrt / torch.rsqrt (GDNF:73-76)
        }
        store_hilo<NV>(p.out_hi + (size_t)r * p.ld_out + c, p.out_lo + (size_t)r * p.ld_out + c, v);
    } break;
    case EPI_KSI: {
        store_f32<NV>(p.out_f32 + (size_t)r * p.ld_f32 + c, v);
    } break;
    case EPI_QUANT: {
        int img, bv, bh;
        step_row_to_block(p.step, r, img, bv, bh);
        const size_t o = (((size_t)img * p.step.Hb + bv) * p.step.Wb + bh) * p.M + c;
        int32_t s[NV];
        uint32_t packed_idx[NV / 4];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float q = rintf(v[i] - pre.a2[i]);                   // torch.round: half to even (ENT:143)
            s[i] = (int32_t)q;
            v[i] = q + pre.a2[i];                                      // y_qnt = y_sym + means (NET:374)
            const int k = scale_to_index(pre.a[i], p.scale_tab);
            if ((i & 3) == 0) packed_idx[i >> 2] = 0;
            packed_idx[i >> 2] |= (uint32_t)k << (8 * (i & 3));
        }
        store_hilo<NV>(p.out_hi + (size_t)r * p.ld_out + c, p.out_lo + (size_t)r * p.ld_out + c, v);
        if (p.sym) {
#pragma unroll
            for (int i = 0; i < NV; i += 4)
                *reinterpret_cast<int4 *>(p.sym + o + i) = make_int4(s[i], s[i + 1], s[i + 2], s[i + 3]);
        }
        if (p.idx) {
#pragma unroll
            for (int i = 0; i < NV / 4; ++i) *reinterpret_cast<uint32_t *>(p.idx + o + 4 * i) = packed_idx[i];
        }
    } break;
    case EPI_RECON: {
        int img, bv, bh;
        step_row_to_block(p.step, r, img, bv, bh);
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = fminf(fmaxf(v[i], -0.5f), 0.5f);   // clamp_(-0.5, 0.5) NET:357
        store_f32<NV>(p.zhat + (((size_t)img * p.step.Hb + bv) * p.step.Wb + bh) * p.cout + c, v);
    } break;
    default: break;
    }
}

template <int NV>
__device__ __forceinline__ void epilogue_store(const EpiParams &p, int r, int c, const float (&acc)[NV]) {
    EpiPre<NV> pre;
    epi_prefetch<NV>(p, p.bias + c, r, c, pre);
    epi_apply<NV>(p, r, c, acc, pre);
}
