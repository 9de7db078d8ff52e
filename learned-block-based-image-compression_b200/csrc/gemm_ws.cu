// Warp-specialised, persistent tcgen05 GEMM: the throughput kernel for large wavefront steps.
//
// profiles/r1_chain_trace.md showed that in gemm_tc_kernel the epilogue (TMEM -> fused math -> staged, coalesced
// stores) takes about as long as the mainloop, and that every tile also pays prologue / pipeline fill / drain.  Here a
// CTA stays resident, loops over output tiles, and runs the three stages of consecutive tiles concurrently:
//
//   warp 0 lane 0   TMA producer: runs ahead over tile boundaries, gated only by the stage ring
//   warp 1 lane 0   tcgen05.mma issuer: accumulates tile i into TMEM buffer (i & 1) as soon as the epilogue has
//                   drained that buffer (acc_empty), tcgen05.commit -> acc_full
//   warps 2..9      epilogue of tile i-1 from the other TMEM buffer, in 32-column groups staged through a small
//                   dedicated shared-memory area (so the ring keeps streaming), then coalesced row stores
//
// Tiles are 128 x bn with bn <= 192 (2 stages of 80 KiB + 44 KiB staging fit the 227 KiB of shared memory).  Same
// operand planes, same k order and same epilogue arithmetic as gemm_tc_kernel: results are bit-identical.
//
// PAIR = true is the CTA-pair form (cluster of 2, tcgen05 cta_group::2).  The single-CTA kernel turned out to be bound
// by L2 -> SM bandwidth, not by the tensor pipe: a 128 x 192 tile moves (128 + 192) x 64 x 4 B per k-block for 12 MMAs,
// i.e. 71 B/clk/SM against a chip-wide L2 limit of ~43 B/clk/SM (profiles/r1_l2_bound.md).  A pair computes a 256 x bn
// tile: each CTA loads its own 128 activation rows but only HALF of the weight tile, and the tensor cores of both SMs
// read both halves, which cuts the bytes per flop by 30 %.  The leader CTA (rank 0) issues every MMA; full barriers live
// in the leader, empty / acc_full barriers are signalled in both CTAs by multicast commits, and both epilogues report
// to the leader's acc_empty barrier.
#include "tc_common.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int WS_THREADS = 320;
constexpr int WS_EPI_THREADS = 256;
constexpr int WS_MAX_BN = 192;
constexpr int WS_ACC_STRIDE = 256;          // TMEM columns between the two accumulators
constexpr int GC = 32;                      // columns per staged group
// One staging region, reused by two passes per group so that it stays small and the operand ring gets the space
// (the kernel is bound by L2 -> SM bandwidth; a third 64 KiB stage for the 256-wide pair tiles is worth more than a
// barrier):  pass F  = the fp32 plane (128 B per row),  pass HL = hi | lo | CDF-index planes (64 + 64 + 32 B per row).
constexpr int WF_STRIDE = GC * 4 + 16;      // 144 B: rows 16 B apart modulo 128 -> conflict-free 16-byte accesses
constexpr int WHL_STRIDE = 176;             // 128 + 48: 3 * 16 B apart modulo 128, also conflict-free
constexpr int WHL_LO = 64, WHL_IDX = 128;   // byte offsets of the lo / index planes within a pass-HL row
constexpr int WSTG_BYTES = BM * WHL_STRIDE; // 22,528 B
// GDN modes: pass HL needs no index plane (row stride 144 B) and the group's pre-activations (fp32, 128 x 128 B) get
// two cp.async buffers behind it, XOR-swizzled in 16-byte chunks instead of padded.
constexpr int WGDN_STRIDE = 144;
constexpr int WAUX_BUF = BM * GC * 4;       // 16,384 B
constexpr int WGDN_BYTES = BM * WGDN_STRIDE + 2 * WAUX_BUF;   // 51,200 B
constexpr int WS_BAR_BLOCK = 128;           // full[4] empty[4] acc_full[2] acc_empty[2] tmem slot
constexpr int WS_TAIL = WS_BAR_BLOCK + 2 * 1024 + ROWTAB_BYTES + STAB_BYTES;   // barriers, two bias slices, row table, scale table

struct WsParams {
    int kb[2];
    int bn, ntiles_n, total_tiles;
    int stages;
    uint32_t slot_bytes, ring_bytes, stg_bytes;   // stg_bytes: staging region (+ the GDN pre-activation buffer), multiple of 128
    uint32_t idesc;
    EpiParams ep;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---- CTA-pair helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(bar), "r"(cta) : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes are credited to the LEADER's barrier (the shared
// window of the odd CTA differs from the even one in bit 24 of the cluster address).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all prior MMAs of the pair -> arrive on the barrier at this offset in both CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(WS_THREADS, 1)
gemm_ws_kernel(const __grid_constant__ CUtensorMap tmA0h, const __grid_constant__ CUtensorMap tmA0l,
               const __grid_constant__ CUtensorMap tmW0h, const __grid_constant__ CUtensorMap tmW0l,
               const __grid_constant__ CUtensorMap tmA1h, const __grid_constant__ CUtensorMap tmA1l,
               const __grid_constant__ CUtensorMap tmW1h, const __grid_constant__ CUtensorMap tmW1l,
               const WsParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t ring = (raw + 1023u) & ~1023u;
    const uint32_t stg = ring + p.ring_bytes;                 // dedicated epilogue staging (1024-aligned)
    const uint32_t bars = stg + p.stg_bytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (MAX_STAGES + s); };
    auto acc_full = [&](int a) { return bars + 8u * (2 * MAX_STAGES + a); };
    auto acc_empty = [&](int a) { return bars + 8u * (2 * MAX_STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * MAX_STAGES + 4);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));
    float *sbias = reinterpret_cast<float *>(smem_raw + (bars + WS_BAR_BLOCK - raw));        // [2][256]
    RowTab *rt = reinterpret_cast<RowTab *>(smem_raw + (bars + WS_BAR_BLOCK + 2048 - raw));
    float *stab = reinterpret_cast<float *>(rt + 1);
    fill_scale_tab(stab, p.ep, threadIdx.x);                       // visible after the setup barrier below

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.kb[0] + p.kb[1];
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;            // 0 = leader
    const int w_rows = PAIR ? p.bn / 2 : p.bn;                      // weight rows this CTA loads per stage
    const uint32_t w_plane = (uint32_t)w_rows * (BK * 2);
    const uint32_t stage_tx = (2 * A_PLANE + 2 * w_plane) * (PAIR ? 2u : 1u);
    const int tile_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int tile_rows = PAIR ? 2 * BM : BM;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), PAIR ? 2 : 1);          // pair: one arrival per CTA's producer, on the leader
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(acc_full(a), 1);
            mbar_init(acc_empty(a), (PAIR ? 2 : 1) * (WS_EPI_THREADS / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA0h)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW0h)) : "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    asm volatile("griddepcontrol.wait;" ::: "memory");             // PDL: see gemm_tc.cu
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = tile_first; t < p.total_tiles; t += tile_stride) {
                const int m0 = (t / p.ntiles_n) * tile_rows + (int)rank * BM;
                const int n0 = (t % p.ntiles_n) * p.bn + (int)rank * w_rows;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sa = ring + s * p.slot_bytes;
                    const bool seg1 = kb >= p.kb[0];
                    const int kk = (seg1 ? kb - p.kb[0] : kb) * BK;
                    if (PAIR) {
                        if (rank == 0) mbar_expect_tx(full_bar(s), stage_tx);
                        else mbar_arrive_remote(full_bar(s), 0);
                        tma_load_2d_pair(sa, seg1 ? &tmA1h : &tmA0h, full_bar(s), kk, m0);
                        tma_load_2d_pair(sa + A_PLANE, seg1 ? &tmA1l : &tmA0l, full_bar(s), kk, m0);
                        tma_load_2d_pair(sa + 2 * A_PLANE, seg1 ? &tmW1h : &tmW0h, full_bar(s), kk, n0);
                        tma_load_2d_pair(sa + 2 * A_PLANE + w_plane, seg1 ? &tmW1l : &tmW0l, full_bar(s), kk, n0);
                    } else {
                        mbar_expect_tx(full_bar(s), stage_tx);
                        tma_load_2d(sa, seg1 ? &tmA1h : &tmA0h, full_bar(s), kk, m0);
                        tma_load_2d(sa + A_PLANE, seg1 ? &tmA1l : &tmA0l, full_bar(s), kk, m0);
                        tma_load_2d(sa + 2 * A_PLANE, seg1 ? &tmW1h : &tmW0h, full_bar(s), kk, n0);
                        tma_load_2d(sa + 2 * A_PLANE + w_plane, seg1 ? &tmW1l : &tmW0l, full_bar(s), kk, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            uint32_t it = 0, ti = 0;
            for (int t = tile_first; t < p.total_tiles; t += tile_stride, ++ti) {
                const uint32_t a = ti & 1u;
                mbar_wait(acc_empty(a), ((ti >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + a * WS_ACC_STRIDE;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1u;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t sa = ring + s * p.slot_bytes;
                    const uint64_t a_hi = make_smem_desc(sa);
                    const uint64_t a_lo = make_smem_desc(sa + A_PLANE);
                    const uint64_t w_hi = make_smem_desc(sa + 2 * A_PLANE);
                    const uint64_t w_lo = make_smem_desc(sa + 2 * A_PLANE + w_plane);
                    auto mma = [&](uint64_t ad, uint64_t wd, uint32_t acc) {
                        if (PAIR) umma_f16_pair(tmem_acc, ad, wd, p.idesc, acc);
                        else umma_f16(tmem_acc, ad, wd, p.idesc, acc);
                    };
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) mma(a_hi + 2 * k, w_hi + 2 * k, (kb | k) != 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) mma(a_hi + 2 * k, w_lo + 2 * k, 1u);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) mma(a_lo + 2 * k, w_hi + 2 * k, 1u);
                    if (PAIR) umma_commit_pair(empty_bar(s)); else umma_commit(empty_bar(s));
                }
                if (PAIR) umma_commit_pair(acc_full(a)); else umma_commit(acc_full(a));
            }
        }
    } else {
        // ---- epilogue warps 2..9 ----------------------------------------------------------------------
        const int ew = warp - 2;                 // 0..7
        const int q = warp & 3;                  // TMEM lane quarter (warp id % 4)
        const int sub = ew >> 2;                 // which 16-column chunk of a 32-column group
        const int et = threadIdx.x - 64;         // 0..255
        const int rl = q * 32 + lane;
        const EpiParams &ep = p.ep;
        const int mode = ep.mode;
        const bool gdn = (mode == EPI_GDN || mode == EPI_IGDN);
        uint32_t ti = 0;
        for (int t = tile_first; t < p.total_tiles; t += tile_stride, ++ti) {
            const uint32_t a = ti & 1u;
            const int m0 = (t / p.ntiles_n) * tile_rows + (int)rank * BM, n0 = (t % p.ntiles_n) * p.bn;
            const int r = m0 + rl;
            const bool row_ok = r < ep.R;
            const int rows_valid = (ep.R - m0) < BM ? (ep.R - m0) : BM;
            float *sb = sbias + a * 256;
            // per-tile setup (overlaps the mainloop of this tile): bias slice, row table
            epi_bar();                           // previous tile's stores have finished reading rt / sbias
            if (mode != EPI_RAW)
                for (int i = et; i < p.bn; i += WS_EPI_THREADS) sb[i] = (n0 + i < ep.cout) ? ep.bias[n0 + i] : 0.0f;
            if (sub == 0 && row_ok) {
                const EpiRowDst d = epi_row_dst(ep, r);
                rt->f32[rl] = reinterpret_cast<unsigned long long>(epi_f32_ptr(ep, d, n0));
                rt->hilo[rl] = (unsigned long long)(d.hilo + n0);
                rt->idx[rl] = reinterpret_cast<unsigned long long>(mode == EPI_QUANT && ep.idx ? ep.idx + d.blk * ep.M + n0 : nullptr);
            }
            epi_bar();
            const uint32_t lane_base = tmem_base + a * WS_ACC_STRIDE + ((uint32_t)(q * 32) << 16);
            const int ngroups = (p.bn + GC - 1) / GC;   // the last group may hold a single 16-column chunk
            // Software pipeline over the 32-column groups: the TMEM load of group g+1 is issued before group g is
            // finished; in the GDN modes each thread also requests its own row's pre-activations of group g+1 (64
            // contiguous bytes) one group ahead, so that global-load latency hides behind a whole group of work.
            const int rsub = lane >> 3, c16 = lane & 7;
            const bool has_f32 = epi_has_f32(mode) && (mode != EPI_QUANT || ep.sym);
            const bool has_hilo = epi_has_hilo(mode);
            auto group_valid = [&](int g) {
                int nv = ep.cout - (n0 + g * GC);
                nv = nv < 0 ? 0 : (nv > GC ? GC : nv);
                return nv > p.bn - g * GC ? p.bn - g * GC : nv;
            };
            auto chunk_ok = [&](int g) {
                return row_ok && (n0 + g * GC + sub * 16) < ep.cout && (g * GC + sub * 16) < p.bn;
            };
            // GDN modes: the group's pre-activations arrive by cp.async with coalesced 16-byte lanes (4 rows x 128 B per
            // warp instruction; a row-per-thread read costs four times the L1 tag lookups and was measurably slower),
            // two groups ahead, into two swizzled buffers behind the staging region: under a saturated L2 a load
            // takes longer than one group of epilogue work.
            const int hl_stride = gdn ? WGDN_STRIDE : WHL_STRIDE;
            const uint32_t aux_base = stg + BM * WGDN_STRIDE;
            auto aux_issue = [&](int g) {
                const int nv = group_valid(g);
                const uint32_t buf = aux_base + (uint32_t)(g & 1) * WAUX_BUF;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int row = ew * 16 + j * 4 + rsub;
                    const bool valid = row < rows_valid && c16 * 4 < nv;
                    const float *src = valid ? ep.aux + (size_t)(m0 + row) * ep.ld_aux + n0 + g * GC + c16 * 4 : ep.aux;
                    const uint32_t dst = buf + row * 128 + ((uint32_t)(c16 ^ (row & 7)) << 4);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            uint32_t accA[16], accB[16];
            if (gdn) {                                       // overlaps the wait for the accumulator
                aux_issue(0);
                if (ngroups > 1) {
                    aux_issue(1);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                epi_bar();
            }
            mbar_wait(acc_full(a), (ti >> 1) & 1u);
            tc_fence_after();
            tmem_ld_issue(lane_base + (uint32_t)(sub * 16), accA);
            for (int g = 0; g < ngroups; ++g) {
                const int g0 = g * GC;
                const int nvalid = group_valid(g);
                const bool even = (g & 1) == 0;
                const bool ok = chunk_ok(g);
                EpiOut<16> o;
                // phase A: this warp's 16-column chunk of the group
                {
                    const int c = n0 + g0 + sub * 16;
                    EpiPre<16> pre;
                    if (ok) {
                        if (gdn) {
                            const uint32_t src = aux_base + (uint32_t)(g & 1) * WAUX_BUF + rl * 128;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint4 v = lds128(src + ((uint32_t)((sub * 4 + i) ^ (rl & 7)) << 4));
                                pre.a[4 * i] = __uint_as_float(v.x); pre.a[4 * i + 1] = __uint_as_float(v.y);
                                pre.a[4 * i + 2] = __uint_as_float(v.z); pre.a[4 * i + 3] = __uint_as_float(v.w);
                            }
                        } else {
                            epi_prefetch<16>(ep, r, c, pre);
                        }
                    }
                    if (even) tmem_ld_wait(accA); else tmem_ld_wait(accB);
                    if (g + 1 < ngroups) {
                        if (even) tmem_ld_issue(lane_base + (uint32_t)(g0 + GC + sub * 16), accB);
                        else tmem_ld_issue(lane_base + (uint32_t)(g0 + GC + sub * 16), accA);
                    } else {
                        // all TMEM reads of this accumulator are done: hand it back to the MMA issuer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (PAIR) mbar_arrive_remote(acc_empty(a), 0);
                            else mbar_arrive(acc_empty(a));
                        }
                    }
                    if (ok) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(even ? accA[i] : accB[i]);
                        epi_compute<16>(ep, sb + g0 + sub * 16, v, pre, o, stab);
                    }
                }
                const int gc = sub * 16;
                auto stage_hl = [&]() {
                    const uint32_t hrow = stg + rl * hl_stride + gc * 2;
                    sts128(hrow, o.hi[0], o.hi[1], o.hi[2], o.hi[3]);
                    sts128(hrow + 16, o.hi[4], o.hi[5], o.hi[6], o.hi[7]);
                    sts128(hrow + WHL_LO, o.lo[0], o.lo[1], o.lo[2], o.lo[3]);
                    sts128(hrow + WHL_LO + 16, o.lo[4], o.lo[5], o.lo[6], o.lo[7]);
                    if (mode == EPI_QUANT) sts128(stg + rl * WHL_STRIDE + WHL_IDX + gc, o.idx[0], o.idx[1], o.idx[2], o.idx[3]);
                };
                auto store_hl = [&]() {
                    const int c8 = lane & 3;                 // 16-byte chunk within the 64-byte plane row
                    const bool is_lo = (lane >> 2) & 1;
                    if (c8 * 8 < nvalid) {
                        h16 *base = (is_lo ? ep.out_lo : ep.out_hi) + g0 + c8 * 8;
                        const uint32_t src = stg + (is_lo ? WHL_LO : 0) + c8 * 16;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int row = ew * 16 + j * 4 + rsub;
                            if (row < rows_valid) {
                                const uint4 v = lds128(src + row * hl_stride);
                                *reinterpret_cast<uint4 *>(base + rt->hilo[row]) = v;
                            }
                        }
                    }
                    if (mode == EPI_QUANT && ep.idx && c16 < 2 && c16 * 16 < nvalid) {      // 2 x 16 B per row
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int row = ew * 16 + j * 4 + rsub;
                            if (row < rows_valid) {
                                const uint4 v = lds128(stg + row * WHL_STRIDE + WHL_IDX + c16 * 16);
                                *reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(rt->idx[row]) + g0 + c16 * 16) = v;
                            }
                        }
                    }
                };
                if (has_f32) {
                    // pass F: stage the fp32 plane, then coalesced stores, 4 rows (128 B each) per warp instruction
                    if (ok) {
                        const uint32_t d = stg + rl * WF_STRIDE + gc * 4;
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            sts128(d + i * 4, __float_as_uint(o.f[i]), __float_as_uint(o.f[i + 1]),
                                   __float_as_uint(o.f[i + 2]), __float_as_uint(o.f[i + 3]));
                    }
                    epi_bar();
                    if (c16 * 4 < nvalid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int row = ew * 16 + j * 4 + rsub;
                            if (row < rows_valid) {
                                const uint4 v = lds128(stg + row * WF_STRIDE + c16 * 16);
                                float *dst = reinterpret_cast<float *>(rt->f32[row]) + g0 + c16 * 4;
                                *reinterpret_cast<uint4 *>(dst) = v;
                            }
                        }
                    }
                    if (has_hilo) {
                        epi_bar();               // the fp32 stores have read the staging area
                        if (ok) stage_hl();
                        epi_bar();
                        if (nvalid > 0) store_hl();
                    }
                } else {
                    if (ok) stage_hl();
                    epi_bar();                   // (GDN: every thread has also read its pre-activations of this group)
                    if (gdn && g + 2 < ngroups) aux_issue(g + 2);   // into the buffer this group has just released
                    if (nvalid > 0) store_hl();
                    if (gdn && g + 1 < ngroups) {                   // this thread's share of group g+1 has landed
                        if (g + 2 < ngroups) asm volatile("cp.async.wait_group 1;" ::: "memory");
                        else asm volatile("cp.async.wait_group 0;" ::: "memory");
                    }
                }
                if (g + 1 < ngroups) epi_bar();   // stores have read the staging area; next pre-activations visible
            }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();   // pair: the peer's shared memory / barriers stay live until both are done
    if (warp == 1) {
        if (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

bool g_ws_attr[2] = {false, false};

}  // namespace

int gemm_ws_max_bn() { return WS_MAX_BN; }
int gemm_pair_max_bn() { return WS_ACC_STRIDE; }

// pair != 0: CTA-pair kernel; g.W[*] tensor maps must then have a box of bn / 2 rows.
int gemm_ws_launch(const GemmCall &g, cudaStream_t st, int pair) {
    if (g.R <= 0) return 0;
    LBIC_TRY(gemm_tc_init());
    pair = pair ? 1 : 0;
    if (g.bn % 16 || g.bn < 16 || g.bn > (pair ? WS_ACC_STRIDE : WS_MAX_BN))
        return lbic_fail(LBIC_ERR_INVALID, "ws kernel: bad tile N %d", g.bn);
    if (!g_ws_attr[pair]) {
        if (pair) LBIC_CUDA(cudaFuncSetAttribute(gemm_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        else LBIC_CUDA(cudaFuncSetAttribute(gemm_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        g_ws_attr[pair] = true;
    }
    WsParams p;
    const int tile_rows = pair ? 2 * BM : BM;
    const int w_rows = pair ? g.bn / 2 : g.bn;
    p.kb[0] = (g.K[0] + BK - 1) / BK;
    p.kb[1] = g.nseg > 1 ? (g.K[1] + BK - 1) / BK : 0;
    p.bn = g.bn;
    p.ntiles_n = (g.cout + g.bn - 1) / g.bn;
    p.total_tiles = ((g.R + tile_rows - 1) / tile_rows) * p.ntiles_n;
    p.slot_bytes = 2 * A_PLANE + 2 * (uint32_t)w_rows * BK * 2;
    const bool gdn_mode = g.ep.mode == EPI_GDN || g.ep.mode == EPI_IGDN;
    p.stg_bytes = (uint32_t)(gdn_mode ? WGDN_BYTES : WSTG_BYTES);
    const int avail = SMEM_LIMIT - 1024 - (int)p.stg_bytes - 128 - WS_TAIL;
    int stages = avail / (int)p.slot_bytes;
    stages = stages > MAX_STAGES ? MAX_STAGES : stages;
    {
        static int cap = -1;   // tuning hook: LBIC_WS_STAGES caps the pipeline depth
        if (cap < 0) { const char *e = getenv("LBIC_WS_STAGES"); cap = e ? atoi(e) : 0; }
        if (cap >= 2 && stages > cap) stages = cap;
    }
    if (stages < 2) return lbic_fail(LBIC_ERR_INVALID, "ws kernel: tile does not fit shared memory");
    p.stages = stages;
    p.ring_bytes = (uint32_t)stages * p.slot_bytes;
    p.idesc = (1u << 4) | ((uint32_t)(g.bn >> 3) << 17) | ((uint32_t)(tile_rows >> 4) << 24);
    p.ep = g.ep;
    int n_sm = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    int grid = p.total_tiles < n_sm ? p.total_tiles : n_sm;
    if (pair) grid = 2 * (p.total_tiles < n_sm / 2 ? p.total_tiles : n_sm / 2);
    const int s1 = g.nseg > 1 ? 1 : 0;
    const size_t smem = 1024 + (size_t)p.ring_bytes + p.stg_bytes + WS_TAIL;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(WS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pair) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (gemm_get_pdl()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (pair)
        LBIC_CUDA(cudaLaunchKernelEx(&cfg, gemm_ws_kernel<true>, *g.A[0].tm_hi, *g.A[0].tm_lo, *g.W[0].tm_hi, *g.W[0].tm_lo,
                                     *g.A[s1].tm_hi, *g.A[s1].tm_lo, *g.W[s1].tm_hi, *g.W[s1].tm_lo, p));
    else
        LBIC_CUDA(cudaLaunchKernelEx(&cfg, gemm_ws_kernel<false>, *g.A[0].tm_hi, *g.A[0].tm_lo, *g.W[0].tm_hi, *g.W[0].tm_lo,
                                     *g.A[s1].tm_hi, *g.A[s1].tm_lo, *g.W[s1].tm_hi, *g.W[s1].tm_lo, p));
    count_launch(0);
    LBIC_CUDA(cudaGetLastError());
    return 0;
}
