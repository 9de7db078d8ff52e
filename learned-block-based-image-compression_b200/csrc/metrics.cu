// Image quality metrics of eval_model on the GPU (agents/blkbsdimgcomp_agent.py:611-619): per-image MSE (-> PSNR) and
// MS-SSIM.  The reference computes MSE with F.mse_loss and MS-SSIM with pytorch_msssim.ms_ssim(x + 0.5, xhat + 0.5,
// data_range=1.0): 11-tap Gaussian window (sigma 1.5), valid filtering, five scales, 2x2 average pooling between scales
// (zero padded on odd sizes, padding counted), per-channel product of relu(cs_s)^w_s (s < 4) and relu(ssim_4)^w_4,
// mean over channels.  pytorch_msssim is not vendored by the reference and not installable offline: this follows its
// published algorithm and is checked against the torch restatement in lbic_b200/codec.py (PARITY UNPINNED against the
// package itself).  HBM-bound: every scale reads its two planes once; sums are reduced in a fixed order (two passes),
// so results are bit-reproducible.
#include "lbic_internal.h"

namespace {

constexpr int WIN = 11;
constexpr int TILE = 16;                 // outputs per block edge
constexpr int IN = TILE + WIN - 1;       // 26 inputs per block edge
__constant__ float c_win[WIN];

__device__ __forceinline__ double block_sum(double v, double *sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    return t;    // valid in thread 0
}

// partial[img][blockIdx.x] = sum over this block's slice of (x - y)^2
__global__ void sqdiff_partial_kernel(const float *__restrict__ x, const float *__restrict__ y, size_t per_img,
                                      double *__restrict__ partial) {
    __shared__ double sh[8];
    const size_t base = (size_t)blockIdx.y * per_img;
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += (size_t)gridDim.x * blockDim.x) {
        const float d = x[base + i] - y[base + i];
        s += (double)d * (double)d;
    }
    const double t = block_sum(s, sh);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// out[g] = sum of n_part partials of group g, in index order
__global__ void sum_partials_kernel(const double *__restrict__ partial, int n_part, double scale, double *__restrict__ out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (int)gridDim.x * (int)blockDim.x) return;
    double s = 0.0;
    for (int i = 0; i < n_part; ++i) s += partial[(size_t)g * n_part + i];
    out[g] = s * scale;
}

// One scale of SSIM for one (image, channel) plane: a block computes TILE x TILE outputs of the valid 11 x 11 Gaussian
// filtering of x, y, x^2, y^2, x y, the ssim / cs maps, and their partial sums.
__global__ void __launch_bounds__(256)
ssim_scale_kernel(const float *__restrict__ x, const float *__restrict__ y, int H, int W, float offset, float c1, float c2,
                  double *__restrict__ part_ssim, double *__restrict__ part_cs) {
    __shared__ float sx[IN][IN + 1], sy[IN][IN + 1];
    __shared__ float hz[5][IN][TILE + 1];          // horizontally filtered: mu1, mu2, xx, yy, xy
    __shared__ double sh[8];
    const int Ho = H - WIN + 1, Wo = W - WIN + 1;
    const int tx = blockIdx.x * TILE, ty = blockIdx.y * TILE;
    const size_t plane = (size_t)blockIdx.z * H * W;
    for (int i = threadIdx.x; i < IN * IN; i += blockDim.x) {
        const int r = i / IN, c = i - r * IN;
        const int yy = ty + r, xx = tx + c;
        const bool in = yy < H && xx < W;
        sx[r][c] = in ? x[plane + (size_t)yy * W + xx] + offset : 0.0f;
        sy[r][c] = in ? y[plane + (size_t)yy * W + xx] + offset : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IN * TILE; i += blockDim.x) {
        const int r = i / TILE, c = i - r * TILE;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const float w = c_win[k], u = sx[r][c + k], v = sy[r][c + k];
            a += w * u; b += w * v; aa += w * (u * u); bb += w * (v * v); ab += w * (u * v);
        }
        hz[0][r][c] = a; hz[1][r][c] = b; hz[2][r][c] = aa; hz[3][r][c] = bb; hz[4][r][c] = ab;
    }
    __syncthreads();
    double s_ssim = 0.0, s_cs = 0.0;
    {
        const int r = threadIdx.x / TILE, c = threadIdx.x % TILE;      // 256 threads = one output each
        if (ty + r < Ho && tx + c < Wo) {
            float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
            for (int k = 0; k < WIN; ++k) {
                const float w = c_win[k];
                m1 += w * hz[0][r + k][c]; m2 += w * hz[1][r + k][c];
                xx += w * hz[2][r + k][c]; yy += w * hz[3][r + k][c]; xy += w * hz[4][r + k][c];
            }
            const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
            const float s1 = xx - mu1_sq, s2 = yy - mu2_sq, s12 = xy - mu12;
            const float cs = (2.0f * s12 + c2) / (s1 + s2 + c2);
            const float ssim = ((2.0f * mu12 + c1) / (mu1_sq + mu2_sq + c1)) * cs;
            s_ssim = ssim; s_cs = cs;
        }
    }
    const double t1 = block_sum(s_ssim, sh);
    const double t2 = block_sum(s_cs, sh);
    if (threadIdx.x == 0) {
        const size_t o = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        part_ssim[o] = t1;
        part_cs[o] = t2;
    }
}

// F.avg_pool2d(x, kernel_size=2, padding=(H % 2, W % 2)): zero padding, padded elements counted (divide by 4 always)
__global__ void avgpool2_kernel(const float *__restrict__ src, float *__restrict__ dst, int planes, int H, int W, int ph,
                                int pw, int Ho, int Wo, float offset) {
    const size_t total = (size_t)planes * Ho * Wo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % Wo);
        const size_t t = i / Wo;
        const int yo = (int)(t % Ho);
        const size_t p = t / Ho;
        float s = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int yy = 2 * yo + dy - ph, xx = 2 * xo + dx - pw;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += src[p * H * W + (size_t)yy * W + xx] + offset;
            }
        dst[i] = s * 0.25f;
    }
}

}  // namespace

// mse_out[n] (host): mean squared difference per image.  msssim_out[n] (host, nullable): MS-SSIM of (x + offset,
// y + offset) per image (mean over channels), data range `range`.  scratch: device, at least metrics_scratch_bytes().
size_t metrics_scratch_bytes(int n, int C, int H, int W) {
    const size_t planes = (size_t)n * C;
    size_t pyr = 0;
    int h = H, w = W;
    for (int s = 1; s < 5; ++s) {
        h = (h + 2 * (h % 2) - 2) / 2 + 1; w = (w + 2 * (w % 2) - 2) / 2 + 1;
        pyr += planes * (size_t)h * w;
    }
    const size_t tiles = (size_t)((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE);
    return 2 * pyr * sizeof(float) + (2 * planes * tiles + 2 * planes * 5 + 1024 * (size_t)n + n + 64) * sizeof(double);
}

int launch_image_metrics(const float *x, const float *y, int n, int C, int H, int W, float offset, float range, void *scratch,
                         double *mse_out, double *msssim_out, cudaStream_t st) {
    if (n < 1 || C < 1 || H < 1 || W < 1) return lbic_fail(LBIC_ERR_INVALID, "bad image shape");
    if (msssim_out && (H < W ? H : W) <= (WIN - 1) * 16)
        return lbic_fail(LBIC_ERR_INVALID, "image too small for five-scale MS-SSIM (smaller side must exceed 160)");
    static bool win_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!win_set[dev & 63]) {
        float g[WIN], sum = 0.f;
        for (int i = 0; i < WIN; ++i) { const float c = (float)(i - WIN / 2); g[i] = expf(-(c * c) / (2.0f * 1.5f * 1.5f)); sum += g[i]; }
        for (int i = 0; i < WIN; ++i) g[i] /= sum;
        LBIC_CUDA(cudaMemcpyToSymbol(c_win, g, sizeof(g)));
        win_set[dev & 63] = true;
    }
    const size_t planes = (size_t)n * C;
    const size_t per_img = (size_t)C * H * W;
    // scratch layout: [pyramid x | pyramid y] floats, then doubles: partials, per-plane sums (5 scales x 2), mse partials, mse
    size_t pyr = 0;
    {
        int h = H, w = W;
        for (int s = 1; s < 5; ++s) { h = (h + 2 * (h % 2) - 2) / 2 + 1; w = (w + 2 * (w % 2) - 2) / 2 + 1; pyr += planes * (size_t)h * w; }
    }
    float *px = (float *)scratch, *py = px + pyr;
    double *dbl = (double *)(((uintptr_t)(py + pyr) + 15) & ~(uintptr_t)15);
    const size_t tiles0 = (size_t)((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE);
    double *part_a = dbl, *part_b = part_a + planes * tiles0, *sums = part_b + planes * tiles0;   // sums[5][2][planes]
    double *mse_part = sums + 10 * planes, *mse = mse_part + 1024 * (size_t)n;
    const int nb = 1024;
    sqdiff_partial_kernel<<<dim3(nb, n), 256, 0, st>>>(x, y, per_img, mse_part);
    count_launch(1);
    sum_partials_kernel<<<n, 1, 0, st>>>(mse_part, nb, 1.0 / (double)per_img, mse);
    count_launch(1);
    LBIC_CUDA(cudaGetLastError());
    std::vector<double> h_sums;
    if (msssim_out) {
        const float c1 = (0.01f * range) * (0.01f * range), c2 = (0.03f * range) * (0.03f * range);
        const float *cx = x, *cy = y;
        float *nx = px, *ny = py;
        int h = H, w = W;
        float off = offset;
        for (int s = 0; s < 5; ++s) {
            const int Ho = h - WIN + 1, Wo = w - WIN + 1;
            dim3 grid((Wo + TILE - 1) / TILE, (Ho + TILE - 1) / TILE, (unsigned)planes);
            ssim_scale_kernel<<<grid, 256, 0, st>>>(cx, cy, h, w, off, c1, c2, part_a, part_b);
            count_launch(1);
            const int n_part = (int)(grid.x * grid.y);
            const double inv = 1.0 / ((double)Ho * (double)Wo);
            sum_partials_kernel<<<(unsigned)planes, 1, 0, st>>>(part_a, n_part, inv, sums + (size_t)(2 * s) * planes);
            sum_partials_kernel<<<(unsigned)planes, 1, 0, st>>>(part_b, n_part, inv, sums + (size_t)(2 * s + 1) * planes);
            count_launch(1);
            if (s < 4) {
                const int ph = h % 2, pw = w % 2;
                const int h2 = (h + 2 * ph - 2) / 2 + 1, w2 = (w + 2 * pw - 2) / 2 + 1;
                const size_t tot = planes * (size_t)h2 * w2;
                const int g = (int)((tot + 255) / 256 > 148 * 16 ? 148 * 16 : (tot + 255) / 256);
                avgpool2_kernel<<<g, 256, 0, st>>>(cx, nx, (int)planes, h, w, ph, pw, h2, w2, off);
                avgpool2_kernel<<<g, 256, 0, st>>>(cy, ny, (int)planes, h, w, ph, pw, h2, w2, off);
                count_launch(1);
                cx = nx; cy = ny;
                nx += tot; ny += tot;
                h = h2; w = w2;
                off = 0.0f;                       // the offset is folded into the pyramid from scale 1 on
            }
        }
        LBIC_CUDA(cudaGetLastError());
        h_sums.resize(10 * planes);
    }
    LBIC_CUDA(cudaMemcpyAsync(mse_out, mse, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (msssim_out) LBIC_CUDA(cudaMemcpyAsync(h_sums.data(), sums, sizeof(double) * 10 * planes, cudaMemcpyDeviceToHost, st));
    LBIC_CUDA(cudaStreamSynchronize(st));      // a metric is a host value (the reference calls .item())
    if (msssim_out) {
        static const double wts[5] = {0.0448, 0.2856, 0.3001, 0.2363, 0.1333};
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int c = 0; c < C; ++c) {
                const size_t p = (size_t)i * C + c;
                double prod = 1.0;
                for (int s = 0; s < 5; ++s) {
                    // scales 0..3 contribute cs, scale 4 ssim; both through relu (pytorch_msssim), in fp32 like the package
                    double v = s < 4 ? h_sums[(size_t)(2 * s + 1) * planes + p] : h_sums[(size_t)(2 * s) * planes + p];
                    v = v > 0.0 ? v : 0.0;
                    prod *= pow(v, wts[s]);
                }
                acc += prod;
            }
            msssim_out[i] = acc / C;
        }
    }
    return 0;
}
