// Micro-benchmark: how long does ONE tcgen05.mma.kind::f16 (K = 16, operands in shared memory) take as a function of
// the tile width N, of M, and of whether consecutive MMAs accumulate into the SAME TMEM accumulator?
//
// Why: a layer of the latency (wave) kernel is a chain of 3 K / 16 MMAs into one accumulator (gemm_wave.cu); its
// mainloop time is proportional to that count (46 ns per MMA at N = 32, profiles/r2_wave_trace_*_v4.txt) although the
// tensor pipe would need 16 cycles for such an MMA.  This program separates issue rate from dependent-accumulate latency.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_chain_bench scripts/mma_chain_bench.cu && /tmp/mma_chain_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}

// n_mma MMAs of M x N x 16, round-robin over n_acc accumulators (TMEM column offset a * acc_stride), operands from
// n_slab different 32-byte k-slices of a 64-wide SWIZZLE_128B k-block (as the real mainloop does); clocks of the whole chain
// (first issue -> commit observed) into out[0], of the issue loop alone into out[1].
// UNIFORM = false: the issue loop runs under `if (threadIdx.x == 0)` (a divergent region: ptxas wraps every UTCHMMA in an
// ELECT / R2UR / BRA.U.ANY sequence); true: the whole warp runs the loop and the MMAs sit under elect.sync, unrolled by 12.
template <bool UNIFORM>
__global__ void __launch_bounds__(128, 1) mma_chain(int M, int N, int n_mma, int n_acc, int acc_stride, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem + (base - smem_u32(smem)))[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (UNIFORM) {
        if (threadIdx.x < 32) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
            const uint64_t a = make_smem_desc(base), b = make_smem_desc(base + 32 * 1024);
            for (int rep = 0; rep < 3; ++rep) {
                const long long t0 = clock64();
                for (int i = 0; i < n_mma; i += 12) {
                    const uint32_t d = tmem + ((i / 12) % n_acc) * acc_stride;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 12; ++k) umma_f16(d, a + 2 * (k & 3), b + 2 * (k & 3), idesc, (i | k) >= n_acc * 12 ? 1u : (k ? 1u : 0u));
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(smem_u32(&bar));
                __syncwarp();
                const long long t1 = clock64();
                while (!mbar_try_wait(smem_u32(&bar), rep & 1)) {}
                const long long t2 = clock64();
                if (threadIdx.x == 0) { out[0] = t2 - t0; out[1] = t1 - t0; }
            }
        }
    } else if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t a = make_smem_desc(base), b = make_smem_desc(base + 32 * 1024);
        for (int rep = 0; rep < 3; ++rep) {      // rep 0, 1 warm up
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                const int acc = i % n_acc;
                umma_f16(tmem + acc * acc_stride, a + 2 * (i & 3), b + 2 * (i & 3), idesc, i >= n_acc ? 1u : 0u);
            }
            umma_commit(smem_u32(&bar));
            const long long t1 = clock64();
            while (!mbar_try_wait(smem_u32(&bar), rep & 1)) {}
            const long long t2 = clock64();
            out[0] = t2 - t0;
            out[1] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long *d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(mma_chain<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024);
    cudaFuncSetAttribute(mma_chain<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("# tcgen05.mma kind::f16 K=16, A and B in shared memory (SWIZZLE_128B K-major); clocks per MMA over a chain of 288\n");
    printf("# %-8s %-4s %-4s %-5s %14s %14s %10s\n", "issue", "M", "N", "accs", "clk/MMA chain", "clk/MMA issue", "floor M*N/256");
    const int n_mma = 288;
    const int Ms[] = {128, 64};
    const int Ns[] = {16, 32, 64, 96, 128, 176, 192, 256};
    for (int uni = 0; uni < 2; ++uni)
    for (int M : Ms)
        for (int N : Ns)
            for (int n_acc : {1, 2, 4}) {
                const int stride = N < 32 ? 32 : N;
                if (n_acc * stride > 512) continue;
                if (uni) mma_chain<true><<<1, 128, 68 * 1024>>>(M, N, n_mma, n_acc, stride, d);
                else mma_chain<false><<<1, 128, 68 * 1024>>>(M, N, n_mma, n_acc, stride, d);
                long long h[2];
                if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) {
                    printf("M=%d N=%d accs=%d: %s\n", M, N, n_acc, cudaGetErrorString(cudaGetLastError()));
                    return 1;
                }
                printf("  %-8s %-4d %-4d %-5d %14.1f %14.1f %10.1f\n", uni ? "uniform" : "lane0", M, N, n_acc, (double)h[0] / n_mma, (double)h[1] / n_mma, 128.0 * N / 256);
            }
    return 0;
}
