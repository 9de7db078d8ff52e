// Device-side building blocks shared by the tcgen05 kernels (gemm_tc.cu: one tile per CTA; gemm_ws.cu: persistent
// per-layer and dataflow kernels; gemm_wave.cu: the latency kernel): mbarrier / TMA / UMMA / TMEM wrappers and the
// shared-memory staged epilogue.
#pragma once
#include "epilogue.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_PLANE = BM * BK * 2;     // 16 KiB per h16 plane
constexpr int MAX_STAGES = 4;
constexpr int NUM_THREADS = 256;          // 8 warps = 2 per SM sub-partition (255 registers available)
constexpr int SMEM_LIMIT = 232448;       // 227 KiB opt-in maximum per CTA
constexpr int BAR_BLOCK = 128;           // full[4] | empty[4] | tmem_full | tmem base address
constexpr int STAB_BYTES = 256;          // shared-memory copy of the 64-entry scale table (QUANT epilogue), right after the row table
constexpr int SMEM_SLACK = 1024 + BAR_BLOCK + 1024 + 3072 + STAB_BYTES;   // ring alignment + barrier block + bias slice (<= 256 floats) + row table + scale table


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
// non-blocking probe of a phase (try_wait may suspend the thread for a while)
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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

// One lane of a CONVERGED warp.  The single-thread instructions (tcgen05.mma, tcgen05.commit) are issued under this
// predicate by a warp whose 32 lanes all walk the role loop: control flow and operands are then warp-uniform, ptxas keeps
// the descriptors in uniform registers and emits the UTCHMMAs of a k-block back to back.  Under `if (lane == 0)` (a
// divergent region) it wraps EVERY UTCHMMA in an ELECT / R2UR / BRA.U.ANY sequence that costs ~150 cycles per MMA --
// more than a 256 x 192 x 16 MMA takes (96) -- scripts/mma_chain_bench.cu, profiles/r2_mma_issue.md.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
constexpr int STG_H_STRIDE = GROUP_COLS * 2 + 16;     // h16 plane row stride
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

// Per-row destinations of a tile (shared memory, filled once per tile by the row's owner thread so that the
// store loops below carry no index arithmetic): fp32-plane pointer, hi/lo element offset, index-plane pointer,
// each already advanced to the tile's first column n0.
struct RowTab {
    unsigned long long f32[BM];
    unsigned long long hilo[BM];
    unsigned long long idx[BM];
};
constexpr int ROWTAB_BYTES = (int)sizeof(RowTab);   // 3 KiB

// copy of the scale table for the QUANT epilogue; the caller synchronises before the epilogue reads it
__device__ __forceinline__ void fill_scale_tab(float *stab, const EpiParams &ep, int tid) {
    if (ep.mode == EPI_QUANT && tid < 64) stab[tid] = ep.scale_tab[tid];
}

__device__ __forceinline__ void epi_chunk_stage(const EpiParams &ep, const float *bias, uint32_t stg, int rl, int gc,
                                                const uint32_t (&raw)[16], const EpiPre<16> &pre, const float *stab) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
    EpiOut<16> o;
    epi_compute<16>(ep, bias, v, pre, o, stab);
    stage_chunk(stg, ep.mode, rl, gc, o);
}

// Epilogue of one 128 x bn tile, executed by all 8 warps of the CTA after the accumulator is complete
// (the caller has waited on the accumulator barrier and issued tcgen05.fence::after_thread_sync).
// sbias: shared-memory copy of the tile's bias slice; stg: shared address of the (idle) ring used for staging;
// rt: shared-memory row table.
__device__ __forceinline__ void tile_epilogue(const EpiParams &ep, const float *sbias, uint32_t stg, RowTab *rt,
                                              uint32_t tmem_acc, int m0, int n0, int bn, int warp, int lane,
                                              unsigned long long *trace = nullptr) {
    const int q = warp & 3;                  // TMEM lane quarter this warp may access (warp id % 4)
    const int ew = warp;                     // epilogue warp 0..7
    const int sub = ew >> 2;                 // the two warps of a lane quarter split each group's chunks
    const int rl = q * 32 + lane;
    const int r = m0 + rl;
    const bool row_ok = r < ep.R;
    const int rows_valid = (ep.R - m0) < BM ? (ep.R - m0) : BM;
    const int mode = ep.mode;
    const bool gdn = (mode == EPI_GDN || mode == EPI_IGDN);
    const uint32_t lane_base = tmem_acc + ((uint32_t)(q * 32) << 16);
    const float *stab = reinterpret_cast<const float *>(rt + 1);   // filled by the caller (fill_scale_tab)
    if (sub == 0 && row_ok) {
        const EpiRowDst d = epi_row_dst(ep, r);
        rt->f32[rl] = reinterpret_cast<unsigned long long>(epi_f32_ptr(ep, d, n0));
        rt->hilo[rl] = (unsigned long long)(d.hilo + n0);
        rt->idx[rl] = reinterpret_cast<unsigned long long>(mode == EPI_QUANT && ep.idx ? ep.idx + d.blk * ep.M + n0 : nullptr);
    }
    for (int g0 = 0; g0 < bn; g0 += GROUP_COLS) {
        const int gcols = (bn - g0) < GROUP_COLS ? (bn - g0) : GROUP_COLS;
        int nvalid = ep.cout - (n0 + g0);
        nvalid = nvalid < 0 ? 0 : (nvalid > gcols ? gcols : nvalid);
        if (gdn) {
            // the pre-GDN activations of this group, loaded row by row with coalesced 16-byte lanes into the
            // (otherwise unused) fp32 staging plane; each thread then reads its own row from shared memory
            const int nv4 = nvalid >> 2;
            if (lane < nv4) {
                const float *src = ep.aux + (size_t)m0 * ep.ld_aux + n0 + g0 + lane * 4;
                const uint32_t dst = stg + STG_F_OFF + lane * 16;
                // 16 rows per warp, in two batches of 8 independent 16-byte loads (all in flight before the first store)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int row = ew + 8 * (half * 8 + j);
                        v[j] = row < rows_valid ? *reinterpret_cast<const uint4 *>(src + (size_t)row * ep.ld_aux)
                                                : make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int row = ew + 8 * (half * 8 + j);
                        sts128(dst + row * STG_F_STRIDE, v[j].x, v[j].y, v[j].z, v[j].w);
                    }
                }
            }
            __syncthreads();
        }
        const int gch = gcols >> 4;
        const int ch_end = sub ? gch : (gch + 1) >> 1;
        int ch = sub ? (gch + 1) >> 1 : 0;
        // chunk `k` of this group covers tile columns g0 + 16k
        auto ok = [&](int k) { return row_ok && (n0 + g0 + k * 16 < ep.cout); };
        auto fetch = [&](int k, EpiPre<16> &pre) {
            if (gdn) {
                const uint32_t a = stg + STG_F_OFF + rl * STG_F_STRIDE + k * 64;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 v = lds128(a + i * 16);
                    pre.a[4 * i] = __uint_as_float(v.x); pre.a[4 * i + 1] = __uint_as_float(v.y);
                    pre.a[4 * i + 2] = __uint_as_float(v.z); pre.a[4 * i + 3] = __uint_as_float(v.w);
                }
            } else {
                epi_prefetch<16>(ep, r, n0 + g0 + k * 16, pre);
            }
        };
        EpiPre<16> preA, preB;
        uint32_t accA[16], accB[16];
        if (ch < ch_end) {
            if (ok(ch)) fetch(ch, preA);
            tmem_ld_issue(lane_base + (uint32_t)(g0 + ch * 16), accA);
            tmem_ld_wait(accA);
        }
        for (; ch < ch_end; ch += 2) {
            const bool hasB = ch + 1 < ch_end;
            if (hasB) {
                tmem_ld_issue(lane_base + (uint32_t)(g0 + (ch + 1) * 16), accB);
                if (ok(ch + 1)) fetch(ch + 1, preB);
            }
            if (ok(ch)) epi_chunk_stage(ep, sbias + g0 + ch * 16, stg, rl, ch * 16, accA, preA, stab);
            if (hasB) {
                tmem_ld_wait(accB);
                const bool hasA = ch + 2 < ch_end;
                if (hasA) {
                    tmem_ld_issue(lane_base + (uint32_t)(g0 + (ch + 2) * 16), accA);
                    if (ok(ch + 2)) fetch(ch + 2, preA);
                }
                if (ok(ch + 1)) epi_chunk_stage(ep, sbias + g0 + (ch + 1) * 16, stg, rl, (ch + 1) * 16, accB, preB, stab);
                if (hasA) tmem_ld_wait(accA);
            }
        }
        if (trace && threadIdx.x == 0) trace[10] = clock64();
        __syncthreads();                                      // the group (and the row table) is staged
        if (trace && threadIdx.x == 0) trace[11] = clock64();
        if (nvalid > 0) {
            // lean store loops: one plane at a time, one row per warp and iteration, 16 bytes per lane
            if (epi_has_f32(mode) && (mode != EPI_QUANT || ep.sym)) {
                const int nv4 = nvalid >> 2;
                const uint32_t src = stg + STG_F_OFF + lane * 16;
                if (lane < nv4) {
#pragma unroll 2
                    for (int row = ew; row < rows_valid; row += 8) {
                        const uint4 v = lds128(src + row * STG_F_STRIDE);
                        float *dst = reinterpret_cast<float *>(rt->f32[row]) + g0 + lane * 4;
                        *reinterpret_cast<uint4 *>(dst) = v;
                    }
                }
            }
            if (epi_has_hilo(mode)) {
                const int l16 = lane & 15;
                const bool is_lo = lane >= 16;
                const uint32_t src = stg + (is_lo ? STG_L_OFF : STG_H_OFF) + l16 * 16;
                h16 *base = (is_lo ? ep.out_lo : ep.out_hi) + g0 + l16 * 8;
                if (l16 < (nvalid >> 3)) {
#pragma unroll 2
                    for (int row = ew; row < rows_valid; row += 8) {
                        const uint4 v = lds128(src + row * STG_H_STRIDE);
                        *reinterpret_cast<uint4 *>(base + rt->hilo[row]) = v;
                    }
                }
            }
            if (mode == EPI_QUANT && ep.idx) {
                if (lane < (nvalid >> 4)) {
                    for (int row = ew; row < rows_valid; row += 8) {
                        const uint4 v = lds128(stg + STG_I_OFF + row * STG_I_STRIDE + lane * 16);
                        *reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(rt->idx[row]) + g0 + lane * 16) = v;
                    }
                }
            }
        }
        if (g0 + GROUP_COLS < bn) __syncthreads();          // before restaging
    }
}

}  // namespace
