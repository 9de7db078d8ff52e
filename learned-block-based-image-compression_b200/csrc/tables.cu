// K4: Gaussian-conditional entropy tables built on the GPU.
//
// Restates GaussianConditional.update (graphs/layers/entropy_layers_cai.py:590-613), _pmf_to_cdf (:175-183)
// and the CompressAI native pmf_to_quantized_cdf it calls (:61-64): for each of the 64 scale levels the
// pmf of the integer-quantised zero-mean Gaussian over [-c, c], c = ceil(scale * 6.1094...), plus the
// tail mass as a final "escape" bin, quantised to a strictly increasing 16-bit CDF with the
// steal-from-the-smallest-frequency fix-up.  One CTA per scale level; the fix-up keeps the reference's
// sequential semantics (outer loop in order, argmin with first-index tie break) and parallelises the
// O(n) inner scans across the CTA.
#include <math.h>

#include "lbic_internal.h"

namespace {

constexpr int MAX_CDF = 4096;   // >= max cdf length (3133 for the reference's scale table)
constexpr int THREADS = 256;

__device__ __forceinline__ float std_cumulative(float v) {
    // _standardized_cumulative (ENT:569-573): 0.5 * erfc(-(2**-0.5) * v) in fp32.  erfc is evaluated in
    // fp64 and rounded once so the result is the correctly rounded fp32 erfc.
    const float c = (float)(-0.70710678118654752440);
    const float arg = c * v;
    const float e = (float)erfc((double)arg);
    return 0.5f * e;
}

__global__ void __launch_bounds__(THREADS) build_cdf_kernel(const float *__restrict__ scales,
                                                            const int *__restrict__ centers, int stride,
                                                            int32_t *__restrict__ cdf_out,
                                                            int32_t *__restrict__ len_out,
                                                            int32_t *__restrict__ off_out) {
    __shared__ uint32_t cdf[MAX_CDF];
    __shared__ unsigned long long red[THREADS / 32];
    __shared__ unsigned long long bcast;
    const int lvl = blockIdx.x, tid = threadIdx.x;
    const float scale = scales[lvl];
    const int center = centers[lvl];
    const int pmf_len = 2 * center + 1;
    const int n = pmf_len + 1;           // + tail bin; cdf has n + 1 entries

    // pmf -> round(p * 2^16)   (std::round: half away from zero)
    for (int j = tid; j < n; j += THREADS) {
        float p;
        if (j < pmf_len) {
            const float samp = fabsf((float)(j - center));
            const float upper = std_cumulative((0.5f - samp) / scale);
            const float lower = std_cumulative((-0.5f - samp) / scale);
            p = upper - lower;
        } else {
            const float samp = fabsf((float)(0 - center));
            p = 2.0f * std_cumulative((-0.5f - samp) / scale);   // tail_mass = 2 * lower[:, :1]  (ENT:607)
        }
        cdf[j + 1] = (uint32_t)roundf(p * 65536.0f);
    }
    if (tid == 0) cdf[0] = 0;
    __syncthreads();
    // total
    unsigned long long part = 0;
    for (int j = tid; j <= n; j += THREADS) part += cdf[j];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < THREADS / 32; ++w) t += red[w];
        bcast = t;
    }
    __syncthreads();
    const unsigned long long total = bcast & 0xffffffffull;   // the reference accumulates in uint32
    for (int j = tid; j <= n; j += THREADS) cdf[j] = (uint32_t)(((1ull << 16) * (unsigned long long)cdf[j]) / total);
    __syncthreads();
    if (tid == 0) {   // partial_sum; n <= 3132, once per model load
        uint32_t run = 0;
        for (int j = 0; j <= n; ++j) {
            run += cdf[j];
            cdf[j] = run;
        }
        cdf[n] = 1u << 16;
    }
    __syncthreads();
    // zero-frequency fix-up
    for (int i = 0; i < n; ++i) {
        if (cdf[i] != cdf[i + 1]) continue;   // uniform: every thread reads the same shared values
        unsigned long long best = ~0ull;      // key = freq << 32 | j  -> min freq, first j on ties
        for (int j = tid; j < n; j += THREADS) {
            const uint32_t freq = cdf[j + 1] - cdf[j];
            if (freq > 1) {
                const unsigned long long key = ((unsigned long long)freq << 32) | (unsigned)j;
                best = key < best ? key : best;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
        if ((tid & 31) == 0) red[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            unsigned long long b = red[0];
            for (int w = 1; w < THREADS / 32; ++w) b = red[w] < b ? red[w] : b;
            bcast = b;
        }
        __syncthreads();
        const int steal = (int)(bcast & 0xffffffffull);
        if (bcast != ~0ull) {
            if (steal < i) {
                for (int j = steal + 1 + tid; j <= i; j += THREADS) cdf[j]--;
            } else {
                for (int j = i + 1 + tid; j <= steal; j += THREADS) cdf[j]++;
            }
        }
        __syncthreads();
    }
    for (int j = tid; j < stride; j += THREADS) cdf_out[(size_t)lvl * stride + j] = (j <= n) ? (int32_t)cdf[j] : 0;
    if (tid == 0) {
        len_out[lvl] = pmf_len + 2;   // _cdf_length = pmf_length + 2 (ENT:613)
        off_out[lvl] = -center;       // _offset = -pmf_center      (ENT:612)
    }
}

// -Phi^-1(q): solve 0.5*erfc(z/sqrt(2)) = q for z (Newton on the fp64 erfc); scipy.stats.norm.ppf in ENT:575-577
double neg_norm_ppf(double q) {
    double z = 6.0;
    for (int it = 0; it < 100; ++it) {
        const double f = 0.5 * erfc(z * 0.70710678118654752440) - q;
        const double d = -exp(-0.5 * z * z) * 0.39894228040143267794;
        const double step = f / d;
        z -= step;
        if (fabs(step) < 1e-15 * fabs(z)) break;
    }
    return z;
}

// One CTA per level: 16-bit copy of the row (without its final 65536) and the 256-bucket search table.
__global__ void __launch_bounds__(THREADS) compact_cdf_kernel(const int32_t *__restrict__ cdf, int stride,
                                                              const int32_t *__restrict__ cdf_len,
                                                              const int32_t *__restrict__ off16, int total,
                                                              uint16_t *__restrict__ out) {
    const int l = blockIdx.x;
    const int32_t *row = cdf + (size_t)l * stride;
    const int len = cdf_len[l];
    uint16_t *dst = out + off16[l];
    for (int k = threadIdx.x; k < len - 1; k += blockDim.x) dst[k] = (uint16_t)row[k];
    uint16_t *lut = out + total + (size_t)l * 257;
    for (int b = threadIdx.x; b <= 256; b += blockDim.x) {
        int lo = 0, hi = len - 1;                 // first k with row[k] > 256 b; row[len-1] = 65536 always qualifies
        if (b < 256) {
            const int32_t t = 256 * b;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (row[mid] > t) hi = mid; else lo = mid + 1;
            }
        } else {
            lo = len - 1;
        }
        lut[b] = (uint16_t)lo;
    }
}

}  // namespace

void tables_free(Tables &T) {
    if (T.cdf) { cudaFree(T.cdf); cudaFree(T.cdf_length); cudaFree(T.offset); T.cdf = nullptr; }
    if (T.cdf16) { cudaFree(T.cdf16); cudaFree(T.cdf16_off); T.cdf16 = nullptr; T.cdf16_off = nullptr; }
    T.cdf16_total = 0;
}

int tables_compact(Tables &T, cudaStream_t st) {
    if (T.cdf16) { cudaFree(T.cdf16); cudaFree(T.cdf16_off); T.cdf16 = nullptr; T.cdf16_off = nullptr; }
    T.cdf16_total = 0;
    if (!T.cdf || T.n_levels != 64) return 0;
    int32_t len[64], off[65];
    LBIC_CUDA(cudaMemcpyAsync(len, T.cdf_length, sizeof(int32_t) * 64, cudaMemcpyDeviceToHost, st));
    LBIC_CUDA(cudaStreamSynchronize(st));
    int total = 0;
    for (int i = 0; i < 64; ++i) {
        if (len[i] < 3 || len[i] > T.stride) return 0;   // not a table this kernel understands: keep the warp kernel
        off[i] = total;
        total += len[i] - 1;
    }
    off[64] = total;
    total = (total + 1) & ~1;                            // keep the bucket tables 4-byte aligned
    if (total > 60000) return 0;                         // 2 B each + 33 KB of bucket tables must fit shared memory
    LBIC_CUDA(cudaMalloc(&T.cdf16, sizeof(uint16_t) * ((size_t)total + 64 * 257 + 8)));   // + padding: copied in 16-byte units
    LBIC_CUDA(cudaMalloc(&T.cdf16_off, sizeof(int32_t) * 65));
    LBIC_CUDA(cudaMemcpyAsync(T.cdf16_off, off, sizeof(int32_t) * 65, cudaMemcpyHostToDevice, st));
    compact_cdf_kernel<<<64, THREADS, 0, st>>>(T.cdf, T.stride, T.cdf_length, T.cdf16_off, total, T.cdf16);
    count_launch(1);
    LBIC_CUDA(cudaGetLastError());
    LBIC_CUDA(cudaStreamSynchronize(st));                // off[] is a stack buffer; once per table update
    T.cdf16_total = total;
    return 0;
}

int tables_build(Tables &T, const float *scale_table_host, int n_levels, double tail_mass, cudaStream_t st) {
    if (n_levels <= 0 || n_levels > 64) return lbic_fail(LBIC_ERR_INVALID, "n_levels must be in 1..64");
    const float mult = (float)neg_norm_ppf(tail_mass / 2);   // fp32 tensor * python float  (ENT:591-592)
    int centers[64], max_len = 0;
    for (int i = 0; i < n_levels; ++i) {
        centers[i] = (int)ceilf(scale_table_host[i] * mult);
        const int len = 2 * centers[i] + 1;
        max_len = len > max_len ? len : max_len;
    }
    const int stride = max_len + 2;
    if (stride > MAX_CDF) return lbic_fail(LBIC_ERR_INVALID, "cdf length %d exceeds the kernel's capacity", stride);
    tables_free(T);
    LBIC_CUDA(cudaMalloc(&T.cdf, sizeof(int32_t) * (size_t)n_levels * stride));
    LBIC_CUDA(cudaMalloc(&T.cdf_length, sizeof(int32_t) * 64));
    LBIC_CUDA(cudaMalloc(&T.offset, sizeof(int32_t) * 64));
    float *d_scales;
    int *d_centers;
    LBIC_CUDA(cudaMalloc(&d_scales, sizeof(float) * 64));
    LBIC_CUDA(cudaMalloc(&d_centers, sizeof(int) * 64));
    LBIC_CUDA(cudaMemcpyAsync(d_scales, scale_table_host, sizeof(float) * n_levels, cudaMemcpyHostToDevice, st));
    LBIC_CUDA(cudaMemcpyAsync(d_centers, centers, sizeof(int) * n_levels, cudaMemcpyHostToDevice, st));
    build_cdf_kernel<<<n_levels, THREADS, 0, st>>>(d_scales, d_centers, stride, T.cdf, T.cdf_length, T.offset);
    count_launch(1);
    LBIC_CUDA(cudaGetLastError());
    LBIC_CUDA(cudaStreamSynchronize(st));   // centers[] is a stack buffer; once per model load
    cudaFree(d_scales);
    cudaFree(d_centers);
    T.n_levels = n_levels;
    T.stride = stride;
    for (int i = 0; i < 64; ++i) T.scale_table[i] = i < n_levels ? scale_table_host[i] : 3.0e38f;
    if (!T.d_scale_table) LBIC_CUDA(cudaMalloc(&T.d_scale_table, sizeof(float) * 64));
    LBIC_CUDA(cudaMemcpy(T.d_scale_table, T.scale_table, sizeof(float) * 64, cudaMemcpyHostToDevice));
    return tables_compact(T, st);
}
