// fp32 SIMT evaluation of the same h16 hi/lo split operands the tcgen05 core consumes.
// Bring-up / cross-check twin only (LBIC_OPT_GEMM_CORE = 1): identical buffers, identical fused
// epilogues, plain FFMA accumulation.  D[r, c] = sum_seg sum_k (Ah+Al)[r,k] * (Wh+Wl)[c,k].
#include "epilogue.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtSeg {
    const h16 *a_hi, *a_lo, *w_hi, *w_lo;
    int lda, ldw, K;
};

struct SimtParams {
    int nseg;
    SimtSeg seg[2];
    EpiParams ep;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p) {
    __shared__ float sA[TK][TM + 4];
    __shared__ float sW[TK][TN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4x4 outputs each
    const int row0 = blockIdx.x * TM, col0 = blockIdx.y * TN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int s = 0; s < p.nseg; ++s) {
        const SimtSeg sg = p.seg[s];
        for (int k0 = 0; k0 < sg.K; k0 += TK) {
            // each thread loads 4 elements of A and 4 of W: element (m = e / TK, k = e % TK)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = tid + i * 256;
                const int m = e / TK, k = e % TK;
                float a = 0.0f, w = 0.0f;
                if (k0 + k < sg.K) {
                    const int r = row0 + m;
                    if (r < p.ep.R) {
                        const size_t o = (size_t)r * sg.lda + k0 + k;
                        a = __half2float(sg.a_hi[o]) + __half2float(sg.a_lo[o]);
                    }
                    const int c = col0 + m;
                    if (c < p.ep.cout) {
                        const size_t o = (size_t)c * sg.ldw + k0 + k;
                        w = __half2float(sg.w_hi[o]) + __half2float(sg.w_lo[o]);
                    }
                }
                sA[k][m] = a;
                sW[k][m] = w;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                float a[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = sW[k][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    const int c = col0 + tx * 4;
    if (c < p.ep.cout) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = row0 + ty * 4 + i;
            if (r < p.ep.R) epilogue_store<4>(p.ep, r, c, acc[i]);
        }
    }
}

}  // namespace

int gemm_simt_launch(const GemmCall &g, cudaStream_t st) {
    if (g.R <= 0) return 0;
    SimtParams p;
    p.nseg = g.nseg;
    for (int s = 0; s < g.nseg; ++s) {
        p.seg[s].a_hi = g.A[s].hi; p.seg[s].a_lo = g.A[s].lo; p.seg[s].lda = g.A[s].ld;
        p.seg[s].w_hi = g.W[s].hi; p.seg[s].w_lo = g.W[s].lo; p.seg[s].ldw = g.W[s].ld;
        p.seg[s].K = g.K[s];
    }
    p.ep = g.ep;
    dim3 grid((g.R + TM - 1) / TM, (g.cout + TN - 1) / TN);
    gemm_simt_kernel<<<grid, 256, 0, st>>>(p);
    count_launch(0);
    LBIC_CUDA(cudaGetLastError());
    return 0;
}
