// HBM-bound reductions next to the spectral path:
//   VideoNormalizer (/root/reference/data_processor.py:201-212): per-pixel mean / std over (slices, frames) of the
//   mouth-crop tensor [N][H][W][F] and the in-place normalisation; the MSE-on-log-mel that network.evaluate reports
//   (/root/reference/network.py:214-220, loss = mean squared error over every element).
#include <cuda_runtime.h>
#include <string>

#include "../../include/avse_b200.h"
#include "avse_ctx.h"

namespace {

// partial sums of one chunk of slices: thread = pixel, its F frame values are contiguous (coalesced 4 F-byte runs)
__global__ void __launch_bounds__(256) avse_video_stats_kernel(const float* __restrict__ video, long long n_slices, int hw, int frames,
                                                               int slices_per_block, double* __restrict__ acc /* [hw][2] */) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hw) return;
    const long long s0 = (long long)blockIdx.y * slices_per_block;
    long long s1 = s0 + slices_per_block;
    if (s1 > n_slices) s1 = n_slices;
    double sum = 0.0, sq = 0.0;
    const size_t stride = (size_t)hw * frames;
    const float* q = video + (size_t)s0 * stride + (size_t)p * frames;
    for (long long s = s0; s < s1; ++s, q += stride)
        for (int f = 0; f < frames; ++f) {
            const double v = (double)q[f];
            sum += v;
            sq = fma(v, v, sq);
        }
    atomicAdd(acc + 2 * (size_t)p, sum);
    atomicAdd(acc + 2 * (size_t)p + 1, sq);
}

__global__ void __launch_bounds__(256) avse_video_finalize_kernel(const double* __restrict__ acc, int hw, double count,
                                                                  float* __restrict__ mean, float* __restrict__ stdv) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hw) return;
    const double m = acc[2 * (size_t)p] / count;
    double var = acc[2 * (size_t)p + 1] / count - m * m;      // np.std: population (ddof = 0)
    if (var < 0.0) var = 0.0;
    mean[p] = (float)m;
    stdv[p] = (float)sqrt(var);
}

__global__ void __launch_bounds__(256) avse_video_normalize_kernel(float* __restrict__ video, long long total, int hw, int frames,
                                                                   const float* __restrict__ mean, const float* __restrict__ stdv) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const int p = (int)((i / frames) % hw);
        video[i] = (video[i] - mean[p]) / stdv[p];             // dp:211-212: no epsilon, like the reference
    }
}

__global__ void __launch_bounds__(256) avse_mse_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                       double* __restrict__ acc) {
    double s = 0.0;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const double d = (double)a[i] - (double)b[i];
        s = fma(d, d, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(acc, t);
    }
}

__global__ void avse_mse_finalize_kernel(const double* __restrict__ acc, double n, float* __restrict__ out) { *out = (float)(*acc / n); }

}  // namespace

extern "C" int avse_video_stats(avse_ctx* ctx, const float* video, long long n_slices, int hw, int frames, double* scratch,
                                float* mean_out, float* std_out, void* stream) {
    if (!ctx || !video || !scratch || !mean_out || !std_out) return avse_fail(AVSE_E_ARG, "avse_video_stats: NULL argument");
    if (n_slices <= 0 || hw <= 0 || frames <= 0) return avse_fail(AVSE_E_ARG, "avse_video_stats: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * (size_t)hw, st));
    const int bx = (hw + 255) / 256;
    long long by = (8LL * ctx->num_sms + bx - 1) / bx;        // ~8 CTAs per SM in flight
    if (by > n_slices) by = n_slices;
    if (by > 65535) by = 65535;
    const int per = (int)((n_slices + by - 1) / by);
    by = (n_slices + per - 1) / per;
    avse_video_stats_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, st>>>(video, n_slices, hw, frames, per, scratch);
    CUDA_TRY(cudaGetLastError());
    avse_video_finalize_kernel<<<bx, 256, 0, st>>>(scratch, hw, (double)n_slices * frames, mean_out, std_out);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int avse_video_normalize(avse_ctx* ctx, float* video, long long n_slices, int hw, int frames, const float* mean,
                                    const float* stdv, void* stream) {
    if (!ctx || !video || !mean || !stdv) return avse_fail(AVSE_E_ARG, "avse_video_normalize: NULL argument");
    if (n_slices <= 0 || hw <= 0 || frames <= 0) return avse_fail(AVSE_E_ARG, "avse_video_normalize: bad sizes");
    const long long total = n_slices * hw * frames;
    long long bx = (total + 256 * 8 - 1) / (256 * 8);
    const long long cap = 16LL * ctx->num_sms;
    if (bx > cap) bx = cap;
    avse_video_normalize_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream>>>(video, total, hw, frames, mean, stdv);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int avse_mse(avse_ctx* ctx, const float* a, const float* b, long long n, double* scratch, float* out, void* stream) {
    if (!ctx || !a || !b || !scratch || !out) return avse_fail(AVSE_E_ARG, "avse_mse: NULL argument");
    if (n <= 0) return avse_fail(AVSE_E_ARG, "avse_mse: n must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double), st));
    long long bx = (n + 256 * 8 - 1) / (256 * 8);
    const long long cap = 16LL * ctx->num_sms;
    if (bx > cap) bx = cap;
    avse_mse_kernel<<<(unsigned)bx, 256, 0, st>>>(a, b, n, scratch);
    CUDA_TRY(cudaGetLastError());
    avse_mse_finalize_kernel<<<1, 1, 0, st>>>(scratch, (double)n, out);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
