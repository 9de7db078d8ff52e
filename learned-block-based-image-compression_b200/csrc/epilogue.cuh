// Fused layer epilogues, shared by the tcgen05 core (gemm_tc.cu) and its SIMT twin (gemm_simt.cu).
// Each call handles NV consecutive output columns [c, c+NV) of one row r (NV in {4,16}); the
// caller guarantees r < p.R and c + NV <= p.cout (all channel counts are multiples of 16).
#pragma once
#include "lbic_internal.h"

#define LBIC_SCALES_MIN 0.11f   // NET:13, ENT:553 LowerBound(scale_bound)

#define LBIC_H16_MAX 65504.0f

// v ~= hi + lo with hi = fp16(v), lo = fp16(v - hi): 22 significand bits.  |v| is saturated at the fp16 maximum
// (activations of this codec are O(1..100); a saturated value loses precision but never becomes inf/NaN).
__device__ __forceinline__ void split_h16(float v, h16 &hi, h16 &lo) {
    v = fminf(fmaxf(v, -LBIC_H16_MAX), LBIC_H16_MAX);
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// build_indexes (ENT:649-654): idx = 63 - #{i<63 : max(s,0.11) <= T[i]} = #{i<63 : T[i] < s'}
// `tab` may point to global memory or to a shared-memory copy of the table (plain generic loads).  The tensor-core
// kernels configure almost the whole L1 as shared memory, so a table left in global memory misses L1 and every probe
// costs an L2 round trip (profiles/r1_quant_epilogue.md); they pass a shared-memory copy.
__device__ __forceinline__ int scale_to_index_bisect(float s, const float *tab) {
    int lo = 0, hi = 63;   // first i in [0,63] with T[i] >= s  (63 = none)
#pragma unroll
    for (int it = 0; it < 6; ++it) {
        const int mid = (lo + hi) >> 1;
        const bool lt = (mid < 63) && (tab[mid] < s);
        lo = lt ? mid + 1 : lo;
        hi = lt ? hi : mid;
    }
    return lo;
}
static __device__ __noinline__ int scale_to_index_slow(float s, const float *tab) { return scale_to_index_bisect(s, tab); }

// Hint for tables that are geometric, like the reference's (ENT:22-28: exp(linspace(log 0.11, log 256, 64))):
// log(s) then lands within a fraction of an entry of the answer.
struct ScaleHint {
    float t0, log_t0, inv_step;
};
__device__ __forceinline__ ScaleHint scale_hint(const float *tab) {
    ScaleHint h;
    h.t0 = tab[0];
    h.log_t0 = __logf(h.t0);
    h.inv_step = 63.0f / (__logf(tab[63]) - h.log_t0);
    return h;
}

// Candidate index from the hint; *ok says whether the table itself confirms it (T[c-1] < s <= T[c], two independent
// probes instead of six dependent ones).  A confirmed candidate IS the bisection's answer for any increasing table.
__device__ __forceinline__ int scale_to_index_hinted(float s, const float *tab, const ScaleHint &h, bool *ok) {
    int c = (int)ceilf((__logf(s) - h.log_t0) * h.inv_step);
    c = c < 0 ? 0 : (c > 63 ? 63 : c);
    c = h.t0 < s ? c : 0;   // scales clamped to the lower bound equal T[0] exactly: rounding in the line above may say 1
    const float below = tab[c > 0 ? c - 1 : 0];
    const float at = tab[c < 63 ? c : 62];
    *ok = (c == 0 || below < s) && (c == 63 || !(at < s));
    return c;
}
__device__ __forceinline__ int scale_to_index(float scale, const float *tab) {
    return scale_to_index_bisect(fmaxf(scale, LBIC_SCALES_MIN), tab);
}

// MUFU.RSQ alone.  rsqrtf() wraps the same instruction in a range fix-up for subnormal arguments (four more instructions
// per element); the GDN norm beta + gamma x^2 is bounded below by the reparametrised beta (GDNF:52-58), a normal number,
// for which both give the same bits.
__device__ __forceinline__ float rsqrt_mufu(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NV>
__device__ __forceinline__ void store_hilo(h16 *__restrict__ ph, h16 *__restrict__ pl, const float (&v)[NV]) {
    static_assert(NV % 4 == 0, "NV");
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
        h16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_h16(v[i + j], h[j], l[j]);
        uint2 uh, ul;
        uh.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
        uh.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
        ul.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
        ul.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
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

// Same, bypassing L1 (ld.global.cg): for buffers another CTA of the SAME launch has just written (the persistent
// dataflow / wavefront kernels reuse the step's compact activation rows launch-long, so an L1 line may be stale).
template <int NV>
__device__ __forceinline__ void load_f32_cg(const float *__restrict__ p, float (&v)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
        const float4 t = __ldcg(reinterpret_cast<const float4 *>(p + i));
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
}

// Operands an epilogue needs besides the accumulator, fetched ahead of the TMEM load they are combined with
// (the tcgen05 core software-pipelines: prefetch chunk i+1 while chunk i is being finished).
template <int NV>
struct EpiPre {
    float a[NV];    // pre-GDN activation (GDN modes) or predicted scale (QUANT)
    float a2[NV];   // predicted mean (QUANT)
};

template <int NV>
__device__ __forceinline__ void epi_prefetch(const EpiParams &p, int r, int c, EpiPre<NV> &pre) {
    if (p.mode == EPI_GDN || p.mode == EPI_IGDN) {
        load_f32_cg<NV>(p.aux + (size_t)r * p.ld_aux + c, pre.a);
    } else if (p.mode == EPI_RESID) {
        load_f32_cg<NV>(p.aux + (size_t)r * p.ld_aux + c, pre.a);            // the block the residual is added to
    } else if (p.mode == EPI_QUANT) {
        load_f32_cg<NV>(p.aux + (size_t)r * p.ld_aux + c, pre.a);            // scales = ksi[:, :M]   (NET:369)
        load_f32_cg<NV>(p.aux + (size_t)r * p.ld_aux + p.M + c, pre.a2);     // means  = ksi[:, M:]
    }
}

// Results of one epilogue chunk, still in registers.  Which planes are meaningful depends on the mode:
//   f32 plane : RAW, PREGDN (a), KSI, RECON (zhat), QUANT (symbols, as int32 bit patterns)
//   hi/lo     : LRELU, PREGDN (a*a), GDN, IGDN, QUANT (y_qnt)
//   idx       : QUANT (CDF indexes, 4 per word)
template <int NV>
struct EpiOut {
    float f[NV];
    uint32_t hi[NV / 2], lo[NV / 2];   // h16 pairs, element 2j in the low half
    uint32_t idx[NV / 4];
};

__host__ __device__ __forceinline__ bool epi_has_f32(int mode) {
    return mode == EPI_RAW || mode == EPI_PREGDN || mode == EPI_KSI || mode == EPI_RECON || mode == EPI_QUANT || mode == EPI_RESID;
}
__host__ __device__ __forceinline__ bool epi_has_hilo(int mode) {
    return mode == EPI_LRELU || mode == EPI_PREGDN || mode == EPI_GDN || mode == EPI_IGDN || mode == EPI_QUANT;
}

template <int NV>
__device__ __forceinline__ void pack_hilo(const float (&v)[NV], EpiOut<NV> &o) {
#pragma unroll
    for (int i = 0; i < NV; i += 2) {
        // two elements at a time: hi = fp16(v), lo = fp16(v - hi)   (same values as split_h16)
        const float a = fminf(fmaxf(v[i], -LBIC_H16_MAX), LBIC_H16_MAX);
        const float b = fminf(fmaxf(v[i + 1], -LBIC_H16_MAX), LBIC_H16_MAX);
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
        o.hi[i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
        o.lo[i >> 1] = *reinterpret_cast<const uint32_t *>(&l);
    }
}

// The arithmetic of every fused epilogue.  bias: pointer to the bias / beta of the chunk's first column (global
// memory, or the CTA's shared-memory copy of its tile slice).
template <int NV>
__device__ __forceinline__ void epi_compute(const EpiParams &p, const float *__restrict__ bias, const float (&acc)[NV],
                                            const EpiPre<NV> &pre, EpiOut<NV> &o, const float *stab = nullptr) {
    if (p.mode == EPI_RAW) {
#pragma unroll
        for (int i = 0; i < NV; ++i) o.f[i] = acc[i];
        return;
    }
    float v[NV];
    load_f32<NV>(bias, v);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = fmaf(acc[i], p.acc_scale, v[i]);   // exact power-of-two rescale, one rounding
    switch (p.mode) {
    case EPI_LRELU: {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = fmaxf(v[i], v[i] * 0.01f);   // nn.LeakyReLU() default slope: x > 0 ? x : 0.01 x, in two instructions
        pack_hilo<NV>(v, o);
    } break;
    case EPI_PREGDN: {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            o.f[i] = v[i];
            v[i] = v[i] * v[i];                                                   // x ** 2 (GDNF:71)
        }
        pack_hilo<NV>(v, o);
    } break;
    case EPI_GDN:
    case EPI_IGDN: {
        // GDNF:73-78: x * rsqrt(norm) (forward) or x * sqrt(norm) (inverse).  rsqrtf is the 2-ulp hardware
        // approximation (the IEEE sqrt + divide sequences cost ~50 instructions per element and made these
        // epilogues 3x longer, profiles/r1_chain_trace.md); its error (2^-22) is far below the h16 hi/lo operand
        // split (2^-22 per operand, more after K accumulation) and it is the same instruction in encoder and decoder.
        const bool inv = (p.mode == EPI_IGDN);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float rs = rsqrt_mufu(v[i]);
            v[i] = pre.a[i] * (inv ? v[i] * rs : rs);
        }
        pack_hilo<NV>(v, o);
    } break;
    case EPI_KSI: {
#pragma unroll
        for (int i = 0; i < NV; ++i) o.f[i] = v[i];
    } break;
    case EPI_RESID: {
        // x + res_net(x) (NET:475-476), then the caller's clamp_(-0.5, 0.5) (AGENT:606) unless no_clamp
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float s = pre.a[i] + v[i];
            o.f[i] = p.no_clamp ? s : fminf(fmaxf(s, -0.5f), 0.5f);
        }
    } break;
    case EPI_QUANT: {
        const float *tab = stab ? stab : p.scale_tab;
        const ScaleHint hint = scale_hint(tab);
        int k[NV];
        bool all_ok = true;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float q = rintf(v[i] - pre.a2[i]);                   // torch.round: half to even (ENT:143)
            o.f[i] = __int_as_float((int32_t)q);
            v[i] = q + pre.a2[i];                                      // y_qnt = y_sym + means (NET:374)
            bool ok;
            k[i] = scale_to_index_hinted(fmaxf(pre.a[i], LBIC_SCALES_MIN), tab, hint, &ok);
            all_ok = all_ok && ok;
        }
        if (!all_ok) {   // a table that is not geometric (or a candidate off by one): bisect
#pragma unroll
            for (int i = 0; i < NV; ++i) k[i] = scale_to_index_slow(fmaxf(pre.a[i], LBIC_SCALES_MIN), tab);
        }
#pragma unroll
        for (int i = 0; i < NV; i += 4)
            o.idx[i >> 2] = (uint32_t)k[i] | ((uint32_t)k[i + 1] << 8) | ((uint32_t)k[i + 2] << 16) | ((uint32_t)k[i + 3] << 24);
        pack_hilo<NV>(v, o);
    } break;
    case EPI_RECON: {
#pragma unroll
        for (int i = 0; i < NV; ++i) o.f[i] = p.no_clamp ? v[i] : fminf(fmaxf(v[i], -0.5f), 0.5f);   // clamp_(-0.5, 0.5) NET:357
    } break;
    default: break;
    }
}

// Where row r of the step's compact matrices lands for each output plane (element offsets, column 0).
struct EpiRowDst {
    size_t f32, hilo, blk;   // blk: block index (img,v,h) for the scattered QUANT / RECON outputs
};

__device__ __forceinline__ EpiRowDst epi_row_dst(const EpiParams &p, int r) {
    EpiRowDst d;
    d.f32 = (size_t)r * p.ld_f32;
    d.hilo = (size_t)r * p.ld_out;
    d.blk = 0;
    if (p.mode == EPI_QUANT || p.mode == EPI_RECON || (p.mode == EPI_LRELU && p.out_pos)) {
        int img, bv, bh;
        step_row_to_block(p.step, r, img, bv, bh);
        d.blk = ((size_t)img * p.step.Hb + bv) * p.step.Wb + bh;
        if (p.mode == EPI_LRELU) d.hilo = g0_pos_index(img, bv, bh, p.step.Hb, p.step.Wb) * p.ld_out;
    }
    return d;
}

// f32-plane destination pointer of (row, column c); nullptr if the mode has none / it is not requested
__device__ __forceinline__ float *epi_f32_ptr(const EpiParams &p, const EpiRowDst &d, int c) {
    switch (p.mode) {
    case EPI_RAW: case EPI_PREGDN: case EPI_KSI: case EPI_RESID: return p.out_f32 + d.f32 + c;
    case EPI_RECON: return p.zhat + d.blk * p.cout + c;
    case EPI_QUANT: return p.sym ? reinterpret_cast<float *>(p.sym) + d.blk * p.M + c : nullptr;
    default: return nullptr;
    }
}

// Direct (uncoalesced, one row per thread) store of a chunk: used by the SIMT twin.
template <int NV>
__device__ __forceinline__ void epi_store_direct(const EpiParams &p, int r, int c, const EpiOut<NV> &o) {
    const EpiRowDst d = epi_row_dst(p, r);
    float *pf = epi_f32_ptr(p, d, c);
    if (pf) {
#pragma unroll
        for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4 *>(pf + i) = make_float4(o.f[i], o.f[i + 1], o.f[i + 2], o.f[i + 3]);
    }
    if (epi_has_hilo(p.mode)) {
#pragma unroll
        for (int i = 0; i < NV / 2; i += 2) {
            *reinterpret_cast<uint2 *>(p.out_hi + d.hilo + c + 2 * i) = make_uint2(o.hi[i], o.hi[i + 1]);
            *reinterpret_cast<uint2 *>(p.out_lo + d.hilo + c + 2 * i) = make_uint2(o.lo[i], o.lo[i + 1]);
        }
    }
    if (p.mode == EPI_QUANT && p.idx) {
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) *reinterpret_cast<uint32_t *>(p.idx + d.blk * p.M + c + 4 * i) = o.idx[i];
    }
}

template <int NV>
__device__ __forceinline__ void epilogue_store(const EpiParams &p, int r, int c, const float (&acc)[NV]) {
    EpiPre<NV> pre;
    EpiOut<NV> o;
    epi_prefetch<NV>(p, r, c, pre);
    epi_compute<NV>(p, p.bias + c, acc, pre, o);
    epi_store_direct<NV>(p, r, c, o);
}
