// tcgen05 / TMEM / TMA GEMM core with fused layer epilogues (sm_100a only).
//
// D[R, cout] = sum over K segments of A_s[R, K_s] * W_s[cout, K_s]^T, fp32-grade accuracy from
// bf16 tensor-core passes: every operand is stored as two bf16 planes (hi = bf16(v),
// lo = bf16(v - hi)) and each 16-wide k-step issues three MMAs into one TMEM accumulator,
//     Ahi*Whi + Ahi*Wlo + Alo*Whi          (the lo*lo term, ~2^-32 relative, is dropped),
// in a fixed order, with no split-K: a row's result depends only on that row's inputs, never on
// the tile it shares or on whether the encoder or the decoder computes it.
//
// CTA = one 128 x bn output tile (bn a multiple of 16, <= 256).  Warp roles:
//   warp 0 lane 0   TMA producer: per 64-wide k-block loads A hi/lo [128x64] and W hi/lo [bn x 64]
//                   (SWIZZLE_128B, K-major) into a multi-stage ring, mbarrier complete_tx
//   warp 1 lane 0   single-thread tcgen05.mma issuer (warp 1 also owns the TMEM allocation);
//                   tcgen05.commit frees stages and finally signals the accumulator
//   all 8 warps     epilogue (two warps per TMEM lane quarter): software-pipelined tcgen05.ld (32 lanes x 16
//                   columns) + operand prefetch -> fused epilogue -> staged in the idle ring -> coalesced stores
#include "epilogue.cuh"

#include <cstdio>
#include <cstring>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_PLANE = BM * BK * 2;     // 16 KiB per bf16 plane
constexpr int MAX_STAGES = 4;
constexpr int NUM_THREADS = 256;          // 8 warps = 2 per SM sub-partition (255 registers available)
constexpr int SMEM_LIMIT = 232448;       // 227 KiB opt-in maximum per CTA
constexpr int BAR_BLOCK = 128;           // full[4] | empty[4] | tmem_full | tmem base address
constexpr int SMEM_SLACK = 1024 + BAR_BLOCK + 1024;   // ring alignment + barrier block + bias slice (<= 256 floats)

struct TcParams {
    int kb[2];          // k-blocks per segment
    int bn;
    int stages;
    uint32_t tmem_cols;
    uint32_t idesc;
    uint32_t ring_bytes;   // max(stages * stage_bytes, epilogue staging), multiple of 1024
    EpiParams ep;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a descriptor/transaction-count bug must fail the launch, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: 8-row atoms of 1024 B (SBO), LBO unused (1),
// descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Asynchronous TMEM load of 16 consecutive fp32 columns of this warp's 32 lanes; the registers may only be
// read after tmem_ld_wait (the "+r" operands there make that a data dependence the compiler must respect).
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// ---- epilogue staging ---------------------------------------------------------------------------
// Once the accumulator is complete every pipeline stage is idle, so the ring is reused to stage one 128-row x
// (<=128)-column group of outputs: one thread per row writes its chunks (padded rows -> conflict-free 16-byte
// shared stores), then whole warps copy each row to global memory as contiguous 16-byte-per-lane stores.  The
// direct alternative (one row per lane, 16 B per store) costs one L1 tag lookup per lane per instruction and made
// the epilogue 60 % of the kernel (profiles/r1_gemm_epilogue.md).
constexpr int GROUP_COLS = 128;
constexpr int STG_F_STRIDE = GROUP_COLS * 4 + 16;     // fp32 plane row stride (bytes)
constexpr int STG_H_STRIDE = GROUP_COLS * 2 + 16;     // bf16 plane row stride
constexpr int STG_I_STRIDE = GROUP_COLS + 16;         // uint8 index plane row stride
constexpr int STG_F_OFF = 0;
constexpr int STG_H_OFF = STG_F_OFF + BM * STG_F_STRIDE;
constexpr int STG_L_OFF = STG_H_OFF + BM * STG_H_STRIDE;
constexpr int STG_I_OFF = STG_L_OFF + BM * STG_H_STRIDE;
constexpr int STG_BYTES = STG_I_OFF + BM * STG_I_STRIDE;   // 155,648 B

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// thread = row `rl` of the tile; writes the 16 columns starting at group-local column gc
__device__ __forceinline__ void stage_chunk(uint32_t stg, int mode, int rl, int gc, const EpiOut<16> &o) {
    if (epi_has_f32(mode)) {
        const uint32_t a = stg + STG_F_OFF + rl * STG_F_STRIDE + gc * 4;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
            sts128(a + i * 4, __float_as_uint(o.f[i]), __float_as_uint(o.f[i + 1]), __float_as_uint(o.f[i + 2]),
                   __float_as_uint(o.f[i + 3]));
    }
    if (epi_has_hilo(mode)) {
        const uint32_t h = stg + STG_H_OFF + rl * STG_H_STRIDE + gc * 2;
        const uint32_t l = stg + STG_L_OFF + rl * STG_H_STRIDE + gc * 2;
        sts128(h, o.hi[0], o.hi[1], o.hi[2], o.hi[3]);
        sts128(h + 16, o.hi[4], o.hi[5], o.hi[6], o.hi[7]);
        sts128(l, o.lo[0], o.lo[1], o.lo[2], o.lo[3]);
        sts128(l + 16, o.lo[4], o.lo[5], o.lo[6], o.lo[7]);
    }
    if (mode == EPI_QUANT) sts128(stg + STG_I_OFF + rl * STG_I_STRIDE + gc, o.idx[0], o.idx[1], o.idx[2], o.idx[3]);
}

// one warp copies row `rl` (global row r) of the staged group: `ncols` valid columns starting at tile column c0
__device__ __forceinline__ void store_row(const EpiParams &ep, uint32_t stg, int rl, int r, int c0, int ncols, int lane) {
    const EpiRowDst d = epi_row_dst(ep, r);
    float *pf = epi_f32_ptr(ep, d, c0);
    if (pf) {
        const uint32_t a = stg + STG_F_OFF + rl * STG_F_STRIDE;
        for (int i = lane; i < (ncols >> 2); i += 32) {
            const uint4 v = lds128(a + i * 16);
            *reinterpret_cast<uint4 *>(pf + i * 4) = v;
        }
    }
    if (epi_has_hilo(ep.mode)) {
        // lanes 0..15 move the hi plane, lanes 16..31 the lo plane (ncols*2 bytes each, <= 256 B)
        const int sub = lane & 15;
        const bool is_lo = lane >= 16;
        const uint32_t a = stg + (is_lo ? STG_L_OFF : STG_H_OFF) + rl * STG_H_STRIDE;
        bf16 *dst = (is_lo ? ep.out_lo : ep.out_hi) + d.hilo + c0;
        if (sub < (ncols >> 3)) {
            const uint4 v = lds128(a + sub * 16);
            *reinterpret_cast<uint4 *>(dst + sub * 8) = v;
        }
    }
    if (ep.mode == EPI_QUANT && ep.idx) {
        if (lane < (ncols >> 4)) {
            const uint4 v = lds128(stg + STG_I_OFF + rl * STG_I_STRIDE + lane * 16);
            *reinterpret_cast<uint4 *>(ep.idx + d.blk * ep.M + c0 + lane * 16) = v;
        }
    }
}

__device__ __forceinline__ void epi_chunk_stage(const EpiParams &ep, const float *bias, uint32_t stg, int rl, int gc,
                                                const uint32_t (&raw)[16], const EpiPre<16> &pre) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
    EpiOut<16> o;
    epi_compute<16>(ep, bias, v, pre, o);
    stage_chunk(stg, ep.mode, rl, gc, o);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0h, const __grid_constant__ CUtensorMap tmA0l,
               const __grid_constant__ CUtensorMap tmW0h, const __grid_constant__ CUtensorMap tmW0l,
               const __grid_constant__ CUtensorMap tmA1h, const __grid_constant__ CUtensorMap tmA1l,
               const __grid_constant__ CUtensorMap tmW1h, const __grid_constant__ CUtensorMap tmW1l,
               const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t ring = (raw + 1023u) & ~1023u;                     // SWIZZLE_128B needs 1024-B alignment
    const uint32_t w_plane = (uint32_t)p.bn * (BK * 2);
    const uint32_t stage_bytes = 2 * A_PLANE + 2 * w_plane;
    const uint32_t bars = ring + p.ring_bytes;                        // after the ring / staging area
    // barrier block: full[4] | empty[4] | tmem_full | tmem base address
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (MAX_STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * MAX_STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * MAX_STAGES + 1);
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * p.bn;
    const int nkb = p.kb[0] + p.kb[1];

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA0h)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW0h)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // every thread helps staging this tile's bias slice (read back by the epilogue from shared memory)
    float *sbias = reinterpret_cast<float *>(smem_raw + (bars + BAR_BLOCK - raw));
    if (p.ep.mode != EPI_RAW)
        for (int i = threadIdx.x; i < p.bn; i += NUM_THREADS) sbias[i] = (n0 + i < p.ep.cout) ? p.ep.bias[n0 + i] : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_expect_tx(full_bar(s), stage_bytes);
                const uint32_t sa = ring + s * stage_bytes;
                const bool seg1 = kb >= p.kb[0];
                const int kk = (seg1 ? kb - p.kb[0] : kb) * BK;
                tma_load_2d(sa, seg1 ? &tmA1h : &tmA0h, full_bar(s), kk, m0);
                tma_load_2d(sa + A_PLANE, seg1 ? &tmA1l : &tmA0l, full_bar(s), kk, m0);
                tma_load_2d(sa + 2 * A_PLANE, seg1 ? &tmW1h : &tmW0h, full_bar(s), kk, n0);
                tma_load_2d(sa + 2 * A_PLANE + w_plane, seg1 ? &tmW1l : &tmW0l, full_bar(s), kk, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = ring + s * stage_bytes;
                const uint64_t a_hi = make_smem_desc(sa);
                const uint64_t a_lo = make_smem_desc(sa + A_PLANE);
                const uint64_t w_hi = make_smem_desc(sa + 2 * A_PLANE);
                const uint64_t w_lo = make_smem_desc(sa + 2 * A_PLANE + w_plane);
                // each k16 step advances the start address by 32 B (= 2 in 16-B units)
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(tmem_acc, a_hi + 2 * k, w_hi + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_acc, a_hi + 2 * k, w_lo + 2 * k, p.idesc, 1u);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_acc, a_lo + 2 * k, w_hi + 2 * k, p.idesc, 1u);
                umma_commit(empty_bar(s));   // frees the stage once these MMAs have read it
            }
            umma_commit(tmem_full_bar);      // accumulator complete
        }
    }
    __syncwarp();
    {
        const int q = warp & 3;                  // TMEM lane quarter this warp may access (warp id % 4)
        const int ew = warp;                     // epilogue warp 0..7
        const int sub = ew >> 2;                 // the two warps of a lane quarter split each group's chunks
        const int rl = q * 32 + lane;
        const int r = m0 + rl;
        const bool row_ok = r < p.ep.R;
        const uint32_t lane_base = tmem_acc + ((uint32_t)(q * 32) << 16);
        const uint32_t stg = ring;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        for (int g0 = 0; g0 < p.bn; g0 += GROUP_COLS) {
            const int gcols = (p.bn - g0) < GROUP_COLS ? (p.bn - g0) : GROUP_COLS;
            const int gch = gcols >> 4;
            const int ch_end = sub ? gch : (gch + 1) >> 1;
            int ch = sub ? (gch + 1) >> 1 : 0;
            // chunk `k` of this group covers tile columns g0 + 16k
            auto ok = [&](int k) { return row_ok && (n0 + g0 + k * 16 < p.ep.cout); };
            EpiPre<16> preA, preB;
            uint32_t accA[16], accB[16];
            if (ch < ch_end) {
                if (ok(ch)) epi_prefetch<16>(p.ep, r, n0 + g0 + ch * 16, preA);
                tmem_ld_issue(lane_base + (uint32_t)(g0 + ch * 16), accA);
                tmem_ld_wait(accA);
            }
            for (; ch < ch_end; ch += 2) {
                const bool hasB = ch + 1 < ch_end;
                if (hasB) {
                    tmem_ld_issue(lane_base + (uint32_t)(g0 + (ch + 1) * 16), accB);
                    if (ok(ch + 1)) epi_prefetch<16>(p.ep, r, n0 + g0 + (ch + 1) * 16, preB);
                }
                if (ok(ch)) epi_chunk_stage(p.ep, sbias + g0 + ch * 16, stg, rl, ch * 16, accA, preA);
                if (hasB) {
                    tmem_ld_wait(accB);
                    const bool hasA = ch + 2 < ch_end;
                    if (hasA) {
                        tmem_ld_issue(lane_base + (uint32_t)(g0 + (ch + 2) * 16), accA);
                        if (ok(ch + 2)) epi_prefetch<16>(p.ep, r, n0 + g0 + (ch + 2) * 16, preA);
                    }
                    if (ok(ch + 1)) epi_chunk_stage(p.ep, sbias + g0 + (ch + 1) * 16, stg, rl, (ch + 1) * 16, accB, preB);
                    if (hasA) tmem_ld_wait(accA);
                }
            }
            __syncthreads();                                      // the group is staged
            int nvalid = p.ep.cout - (n0 + g0);
            nvalid = nvalid < 0 ? 0 : (nvalid > gcols ? gcols : nvalid);
            if (nvalid > 0) {
                for (int row = ew; row < BM; row += 8) {
                    const int rr = m0 + row;
                    if (rr < p.ep.R) store_row(p.ep, stg, row, rr, n0 + g0, nvalid, lane);
                }
            }
            if (g0 + GROUP_COLS < p.bn) __syncthreads();          // before restaging
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(p.tmem_cols)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;

}  // namespace

int gemm_tc_init() {
    if (g_encode_tiled) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    LBIC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
        return lbic_fail(LBIC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    LBIC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    return 0;
}

// bf16 row-major [outer][ld_elems] matrix, logical inner extent `inner`; out-of-bounds box elements
// read as zero.
int make_tmap_2d(CUtensorMap *tm, const void *base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                 uint32_t box_inner, uint32_t box_outer) {
    LBIC_TRY(gemm_tc_init());
    if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld_elems * 2) & 15))
        return lbic_fail(LBIC_ERR_INVALID, "TMA operand must be 16-byte aligned (base %p, ld %llu)", base,
                         (unsigned long long)ld_elems);
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstride[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return lbic_fail(LBIC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
    return 0;
}

int gemm_tc_launch(const GemmCall &g, cudaStream_t st) {
    if (g.R <= 0) return 0;
    LBIC_TRY(gemm_tc_init());
    if (g.bn % 16 || g.bn < 16 || g.bn > 256) return lbic_fail(LBIC_ERR_INVALID, "bad tile N %d", g.bn);
    TcParams p;
    p.kb[0] = (g.K[0] + BK - 1) / BK;
    p.kb[1] = g.nseg > 1 ? (g.K[1] + BK - 1) / BK : 0;
    p.bn = g.bn;
    const int stage_bytes = 2 * A_PLANE + 2 * g.bn * BK * 2;
    int stages = (SMEM_LIMIT - SMEM_SLACK) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return lbic_fail(LBIC_ERR_INVALID, "tile does not fit shared memory");
    p.stages = stages;
    uint32_t cols = 32;
    while ((int)cols < g.bn) cols <<= 1;
    p.tmem_cols = cols;
    // kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=bn, M=128
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    p.ep = g.ep;
    const int s1 = g.nseg > 1 ? 1 : 0;
    dim3 grid((g.R + BM - 1) / BM, (g.cout + g.bn - 1) / g.bn);
    size_t ring_bytes = (size_t)stages * stage_bytes;
    if (ring_bytes < (size_t)STG_BYTES) ring_bytes = (STG_BYTES + 1023) / 1024 * 1024;   // epilogue staging lives in the ring
    p.ring_bytes = (uint32_t)ring_bytes;
    const size_t smem = ring_bytes + SMEM_SLACK;
    gemm_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(*g.A[0].tm_hi, *g.A[0].tm_lo, *g.W[0].tm_hi, *g.W[0].tm_lo,
                                                    *g.A[s1].tm_hi, *g.A[s1].tm_lo, *g.W[s1].tm_hi, *g.W[s1].tm_lo, p);
    count_launch(0);
    LBIC_CUDA(cudaGetLastError());
    return 0;
}
