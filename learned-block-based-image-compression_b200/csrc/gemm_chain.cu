// Persistent chain kernel: one launch runs a whole sequence of layers (e.g. the 19 GEMMs of an encode wavefront
// step) instead of one launch per layer.
//
// A thread-block cluster of S CTAs owns one 128-row tile of the step's compact activation matrices at a time and
// walks it through layers [l0, l1): in every layer CTA `rank` computes output tiles n = rank, rank + S, ... (the same
// 128 x bn tcgen05 tile as gemm_tc.cu, same k order, so results are bit-identical to the per-layer path), the
// epilogue writes the next layer's operands to global memory (they stay in L2), and a cluster barrier separates the
// layers.  Row tiles never depend on each other inside a step, so there is no grid-wide synchronisation; clusters
// stride over the row tiles.  This removes 18 of the 19 launches of a step together with their fixed cost (launch gap,
// barrier init, TMEM allocation, descriptor fetch), which profiles/r1_gemm_epilogue.md showed to be ~2/3 of a step.
#include "tc_common.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

struct ChainParams {
    const ChainLayer *layers;
    int l0, l1;
    int R;
    StepDesc step;
    int S;               // cluster size
    int vi;              // tile-width variant = index of S among the split factors
    int n_clusters;
    int stages;
    uint32_t slot_bytes;   // pipeline slot: A hi/lo (32 KiB) + W hi/lo for the widest tile of this launch
    uint32_t ring_bytes;
    uint32_t tmem_cols;
    unsigned long long *trace;   // debug: per-phase clock64() stamps of CTA 0 (LBIC_CHAIN_TRACE), else nullptr
};

struct LayerScalars {
    int kb[2];
    int cout, bn, ntiles, vi;
    uint32_t idesc;
    EpiParams ep;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

#define TRACE0(slot) do { if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[(l - p.l0) * 16 + (slot)] = clock64(); } while (0)
#define TRACE1(slot) do { if (p.trace && blockIdx.x == 0 && threadIdx.x == 32) p.trace[(l - p.l0) * 16 + (slot)] = clock64(); } while (0)

constexpr int CHAIN_SLACK = 1024 + BAR_BLOCK + 1024 + 512 + 3072 + STAB_BYTES;   // ring alignment, barriers, bias slice, layer scalars, row table

__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_chain_kernel(const ChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t ring = (raw + 1023u) & ~1023u;
    const uint32_t bars = ring + p.ring_bytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (MAX_STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * MAX_STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * MAX_STAGES + 1);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));
    float *sbias = reinterpret_cast<float *>(smem_raw + (bars + BAR_BLOCK - raw));
    LayerScalars *sl = reinterpret_cast<LayerScalars *>(smem_raw + (bars + BAR_BLOCK + 1024 - raw));
    RowTab *rowtab = reinterpret_cast<RowTab *>(smem_raw + (bars + BAR_BLOCK + 1024 + 512 - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = p.S > 1 ? (int)cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / p.S;
    const int row_tiles = (p.R + BM - 1) / BM;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot_ptr;

    uint32_t it = 0;         // running k-block counter (producer and MMA issuer each advance their own copy)
    uint32_t tile_cnt = 0;   // tiles this CTA has finished (accumulator barrier parity)

    for (int m = cluster_id; m < row_tiles; m += p.n_clusters) {
        const int m0 = m * BM;
        for (int l = p.l0; l < p.l1; ++l) {
            const ChainLayer *L = p.layers + l;
            TRACE0(0);
            // layer scalars + epilogue parameters -> shared memory (one copy per CTA)
            if (threadIdx.x == 0) {
                const int vi = p.vi;
                const int bn = L->bn_v[vi];
                sl->kb[0] = L->kb[0];
                sl->kb[1] = L->nseg > 1 ? L->kb[1] : 0;
                sl->cout = L->cout;
                sl->bn = bn;
                sl->vi = vi;
                sl->ntiles = (L->cout + bn - 1) / bn;
                sl->idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            }
            {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(&L->ep);
                uint32_t *dst = reinterpret_cast<uint32_t *>(&sl->ep);
                for (int i = threadIdx.x; i < (int)(sizeof(EpiParams) / 4); i += NUM_THREADS) dst[i] = src[i];
            }
            __syncthreads();
            EpiParams ep = sl->ep;
            ep.R = p.R;
            ep.step = p.step;
            const int bn = sl->bn, vi = sl->vi, ntiles = sl->ntiles;
            const int nkb = sl->kb[0] + sl->kb[1], kb0 = sl->kb[0];
            const uint32_t idesc = sl->idesc;
            const uint32_t w_plane = (uint32_t)bn * (BK * 2);
            const uint32_t stage_tx = 2 * A_PLANE + 2 * w_plane;
            TRACE0(1);

            for (int n = rank; n < ntiles; n += p.S) {
                const int n0 = n * bn;
                if (ep.mode != EPI_RAW)
                    for (int i = threadIdx.x; i < bn; i += NUM_THREADS)
                        sbias[i] = (n0 + i < ep.cout) ? ep.bias[n0 + i] : 0.0f;
                fill_scale_tab(reinterpret_cast<float *>(rowtab + 1), ep, threadIdx.x);
                __syncthreads();
                TRACE0(2);
                if (warp == 0) {
                    if (lane == 0) {
                        for (int kb = 0; kb < nkb; ++kb, ++it) {
                            const int s = it % p.stages;
                            const uint32_t ph = (it / p.stages) & 1u;
                            mbar_wait(empty_bar(s), ph ^ 1u);
                            mbar_expect_tx(full_bar(s), stage_tx);
                            const uint32_t sa = ring + s * p.slot_bytes;
                            const int sg = kb >= kb0 ? 1 : 0;
                            const int kk = (sg ? kb - kb0 : kb) * BK;
                            tma_load_2d(sa, &L->tmA[sg][0], full_bar(s), kk, m0);
                            tma_load_2d(sa + A_PLANE, &L->tmA[sg][1], full_bar(s), kk, m0);
                            tma_load_2d(sa + 2 * A_PLANE, &L->tmW[vi][sg][0], full_bar(s), kk, n0);
                            tma_load_2d(sa + 2 * A_PLANE + w_plane, &L->tmW[vi][sg][1], full_bar(s), kk, n0);
                        }
                    }
                } else if (warp == 1) {
                    if (lane == 0) {
                        tc_fence_after();
                        for (int kb = 0; kb < nkb; ++kb, ++it) {
                            const int s = it % p.stages;
                            const uint32_t ph = (it / p.stages) & 1u;
                            mbar_wait(full_bar(s), ph);
                            tc_fence_after();
                            if (kb == 0) TRACE1(8);
                            const uint32_t sa = ring + s * p.slot_bytes;
                            const uint64_t a_hi = make_smem_desc(sa);
                            const uint64_t a_lo = make_smem_desc(sa + A_PLANE);
                            const uint64_t w_hi = make_smem_desc(sa + 2 * A_PLANE);
                            const uint64_t w_lo = make_smem_desc(sa + 2 * A_PLANE + w_plane);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_f16(tmem_acc, a_hi + 2 * k, w_hi + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_acc, a_hi + 2 * k, w_lo + 2 * k, idesc, 1u);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_acc, a_lo + 2 * k, w_hi + 2 * k, idesc, 1u);
                            umma_commit(empty_bar(s));
                        }
                        umma_commit(tmem_full_bar);
                        TRACE1(9);
                    }
                }
                __syncwarp();
                TRACE0(3);
                mbar_wait(tmem_full_bar, tile_cnt & 1u);
                tc_fence_after();
                TRACE0(4);
                tile_epilogue(ep, sbias, ring, rowtab, tmem_acc, m0, n0, bn, warp, lane,
                              (p.trace && blockIdx.x == 0) ? p.trace + (l - p.l0) * 16 : nullptr);
                ++tile_cnt;
                TRACE0(5);
                // the staging area (generic proxy) becomes a TMA destination (async proxy) again, and the accumulator
                // is about to be overwritten: order both before the next tile
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();
                __syncthreads();
            }
            // layer boundary: this layer's global stores (generic proxy) are read by the next layer's TMA loads
            // (async proxy), possibly issued by another CTA of the cluster
            TRACE0(6);
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            if (p.S > 1) cluster_sync_all(); else __syncthreads();
            asm volatile("fence.proxy.async;" ::: "memory");
            TRACE0(7);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(p.tmem_cols)
                     : "memory");
    }
}

unsigned long long g_chain_attr_mask = 0;

// per-tile time model (us) used only to pick the cluster size; from profiles/r1_gemm_sweep_v2_staged_epilogue.log
double tile_us(int kblocks, int bn) {
    const double per_kb = bn > 128 ? 0.87 : (bn > 64 ? 0.61 : 0.45);
    const double epi = 2.0 + bn / 64.0;
    return kblocks * per_kb + epi + 1.0;
}

}  // namespace

int chain_variant(const ChainLayer &L, int S) {
    for (int i = 0; i < LBIC_NBN; ++i)
        if (lbic_split(i) == S) return i;
    (void)L;
    return 0;
}

int gemm_chain_launch(const ChainLayer *d_layers, const ChainLayer *h_layers, int l0, int l1, int R,
                      const StepDesc &step, int force_S, cudaStream_t st) {
    if (R <= 0 || l1 <= l0) return 0;
    LBIC_TRY(gemm_tc_init());
    if (lbic_first_use_on_device(g_chain_attr_mask)) {
        LBIC_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        LBIC_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
    }
    int n_sm = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    const int row_tiles = (R + BM - 1) / BM;
    // choose the cluster size by a small cost model: rounds x sum over layers of (tiles per CTA x tile time)
    static const int cand[] = {1, 2, 3, 4, 6, 8};
    int best_S = 1;
    double best_cost = 1e30;
    for (int S : cand) {
        if (force_S && S != force_S) continue;
        const int ncl = (n_sm / S) < row_tiles ? (n_sm / S) : row_tiles;
        if (ncl < 1) continue;
        const int rounds = (row_tiles + ncl - 1) / ncl;
        double chain = 0;
        for (int l = l0; l < l1; ++l) {
            const ChainLayer &L = h_layers[l];
            const int bn = L.bn_v[chain_variant(L, S)];
            const int ntiles = (L.cout + bn - 1) / bn;
            const int per_cta = (ntiles + S - 1) / S;
            chain += per_cta * tile_us(L.kb[0] + (L.nseg > 1 ? L.kb[1] : 0), bn) + (S > 1 ? 0.6 : 0.2);
        }
        const double cost = rounds * chain;
        if (cost < best_cost) { best_cost = cost; best_S = S; }
    }
    ChainParams p;
    p.layers = d_layers; p.l0 = l0; p.l1 = l1; p.R = R; p.step = step;
    p.S = best_S;
    p.vi = chain_variant(h_layers[l0], best_S);
    p.n_clusters = (n_sm / best_S) < row_tiles ? (n_sm / best_S) : row_tiles;
    int max_bn = 16;
    for (int l = l0; l < l1; ++l) {
        const int bn = h_layers[l].bn_v[chain_variant(h_layers[l], best_S)];
        max_bn = bn > max_bn ? bn : max_bn;
    }
    p.slot_bytes = 2 * A_PLANE + 2 * (uint32_t)max_bn * BK * 2;
    int stages = (SMEM_LIMIT - CHAIN_SLACK) / (int)p.slot_bytes;
    stages = stages > MAX_STAGES ? MAX_STAGES : stages;
    p.stages = stages;
    size_t ring_bytes = (size_t)stages * p.slot_bytes;
    if (ring_bytes < (size_t)STG_BYTES) ring_bytes = (STG_BYTES + 1023) / 1024 * 1024;
    p.ring_bytes = (uint32_t)ring_bytes;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < max_bn) p.tmem_cols <<= 1;

    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(p.n_clusters * p.S, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = ring_bytes + CHAIN_SLACK;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    p.trace = nullptr;
    static int trace_left = getenv("LBIC_CHAIN_TRACE") ? atoi(getenv("LBIC_CHAIN_TRACE")) : 0;
    const bool tracing = trace_left > 0 && row_tiles >= 20;
    if (tracing) {
        --trace_left;
        LBIC_CUDA(cudaMalloc(&p.trace, sizeof(unsigned long long) * 16 * 32));
        LBIC_CUDA(cudaMemsetAsync(p.trace, 0, sizeof(unsigned long long) * 16 * 32, st));
    }
    LBIC_CUDA(cudaLaunchKernelEx(&cfg, gemm_chain_kernel, p));
    count_launch(0);
    if (tracing) {
        unsigned long long h[16 * 32];
        LBIC_CUDA(cudaStreamSynchronize(st));
        LBIC_CUDA(cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(p.trace);
        fprintf(stderr, "[chain trace] R=%d S=%d clusters=%d stages=%d layers %d..%d (cycles: setup, bias, issue, mma-wait, epi, tail, sync | first-data, mma-done rel. to tile start)\n",
                R, p.S, p.n_clusters, p.stages, l0, l1);
        for (int l = 0; l < l1 - l0; ++l) {
            const unsigned long long *t = h + l * 16;
            fprintf(stderr, "  L%02d bn=%3d kb=%2d: %6lld %6lld %6lld %6lld %6lld %6lld %6lld | %6lld %6lld  total %lld\n", l0 + l,
                    h_layers[l0 + l].bn_v[chain_variant(h_layers[l0 + l], p.S)], h_layers[l0 + l].kb[0] + h_layers[l0 + l].kb[1],
                    (long long)(t[1] - t[0]), (long long)(t[2] - t[1]), (long long)(t[3] - t[2]), (long long)(t[4] - t[3]),
                    (long long)(t[5] - t[4]), (long long)(t[6] - t[5]), (long long)(t[7] - t[6]), (long long)(t[8] - t[2]),
                    (long long)(t[9] - t[2]), (long long)(t[7] - t[0]));
            fprintf(stderr, "        epilogue: phaseA %lld  barrier %lld  phaseB %lld\n", (long long)(t[10] - t[4]),
                    (long long)(t[11] - t[10]), (long long)(t[5] - t[11]));
        }
    }
    return 0;
}
