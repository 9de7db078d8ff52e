// Which TMEM lane holds row r of the accumulator of a tcgen05.mma.cta_group::1.kind::f16 with M = 64 (and M = 128)?
// A[r][0] = r + 1 (other k zero), B[n][0] = 1  ->  D[r][n] = r + 1; every warp reads its 32 lanes of column 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_m64_layout scripts/mma_m64_layout.cu && /tmp/mma_m64_layout
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
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

__global__ void __launch_bounds__(128, 1) probe(int M, float *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    uint8_t *b0 = smem + (base - smem_u32(smem));
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(b0)[i] = 0;
    __syncthreads();
    // K-major SWIZZLE_128B: row r at r * 128 B, 16-byte chunk c at ((c ^ (r & 7)) << 4); element k = 0 is in chunk 0
    if (threadIdx.x < 128) {
        const int r = threadIdx.x;
        __half *a = reinterpret_cast<__half *>(b0 + r * 128 + ((0 ^ (r & 7)) << 4));
        a[0] = __float2half((float)(r + 1));
        __half *b = reinterpret_cast<__half *>(b0 + 32 * 1024 + r * 128 + ((0 ^ (r & 7)) << 4));
        if (r < 32) b[0] = __float2half(1.0f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t a = make_smem_desc(base), b = make_smem_desc(base + 32 * 1024);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[threadIdx.x] = __uint_as_float(v);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
    float *d, h[128];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024);
    for (int M : {128, 64}) {
        cudaMemset(d, 0, sizeof(h));
        probe<<<1, 128, 68 * 1024>>>(M, d);
        if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("M=%d: %s\n", M, cudaGetErrorString(cudaGetLastError())); return 1; }
        printf("M = %d: TMEM lane -> accumulator row + 1 (column 0)\n", M);
        for (int l = 0; l < 128; ++l) printf("%s%3d:%-4g", l % 16 == 0 ? "\n  " : " ", l, h[l]);
        printf("\n");
    }
    return 0;
}
