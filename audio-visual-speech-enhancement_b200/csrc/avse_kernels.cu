// CUDA kernels (sm_100a) + C ABI of the spectral front/back end.  See include/avse_b200.h.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <new>

#include "../../include/avse_b200.h"
#include "avse_common.h"
#include "avse_tables.h"
#include "avse_fwd_stages.cuh"

using namespace avse;

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct avse_ctx {
    int device = 0;
    int num_sms = 148;
    HostTables host;
    // device tables (one allocation)
    void* dbase = nullptr;
    FwdTables fwd{};
    const float* d_tri_w = nullptr;
    const float* d_tri_ipiv = nullptr;
    const float* d_tri_sup = nullptr;
    const int* d_col_band = nullptr;
    const float* d_col_w = nullptr;
};

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return (int)e;
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_, #x); } while (0)

extern "C" const char* avse_last_error(void) { return g_err.c_str(); }
extern "C" const char* avse_version(void) { return "avse_b200 0.1 (sm_100a)"; }

extern "C" int avse_create(int sample_rate, double fmin, double fmax, int device, avse_ctx** out) {
    if (out == nullptr) return fail(AVSE_E_ARG, "avse_create: out is NULL");
    *out = nullptr;
    avse_ctx* c = new (std::nothrow) avse_ctx();
    if (c == nullptr) return fail(AVSE_E_ARG, "avse_create: out of host memory");
    if (!build_tables(c->host, sample_rate, fmin, fmax)) {
        std::string e = c->host.error;
        delete c;
        return fail(AVSE_E_CONFIG, "avse_create: " + e);
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { delete c; return fail(AVSE_E_NOCUDA, "avse_create: no CUDA device (this library has no CPU path)"); }
    if (device < 0 || device >= ndev) { delete c; return fail(AVSE_E_ARG, "avse_create: bad device index"); }
    c->device = device;
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);

    const HostTables& h = c->host;
    // pack everything into one buffer, 256-byte aligned sections
    struct Sec { const void* src; size_t bytes; size_t off; };
    std::vector<Sec> secs = {
        {h.window.data(), h.window.size() * 4, 0},     {h.tw1t.data(), h.tw1t.size() * 4, 0},
        {h.mel_w.data(), h.mel_w.size() * 4, 0},       {h.mel_lo.data(), h.mel_lo.size() * 4, 0},
        {h.mel_roundw.data(), h.mel_roundw.size() * 4, 0}, {h.tri_w.data(), h.tri_w.size() * 4, 0},
        {h.tri_ipiv.data(), h.tri_ipiv.size() * 4, 0}, {h.tri_sup.data(), h.tri_sup.size() * 4, 0},
        {h.col_band.data(), h.col_band.size() * 4, 0}, {h.col_w.data(), h.col_w.size() * 4, 0},
    };
    size_t total = 0;
    for (auto& s : secs) { s.off = total; total += (s.bytes + 255) / 256 * 256; }
    std::vector<char> stage(total, 0);
    for (auto& s : secs) memcpy(stage.data() + s.off, s.src, s.bytes);
    e = cudaMalloc(&c->dbase, total);
    if (e != cudaSuccess) { delete c; cudaSetDevice(prev); return cuda_fail(e, "cudaMalloc(tables)"); }
    e = cudaMemcpy(c->dbase, stage.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(c->dbase); delete c; cudaSetDevice(prev); return cuda_fail(e, "cudaMemcpy(tables)"); }
    char* b = (char*)c->dbase;
    c->fwd.window = (const float*)(b + secs[0].off);
    c->fwd.tw1t = (const float*)(b + secs[1].off);
    c->fwd.mel_w = (const float*)(b + secs[2].off);
    c->fwd.mel_lo = (const int*)(b + secs[3].off);
    c->fwd.mel_roundw = (const int*)(b + secs[4].off);
    c->d_tri_w = (const float*)(b + secs[5].off);
    c->d_tri_ipiv = (const float*)(b + secs[6].off);
    c->d_tri_sup = (const float*)(b + secs[7].off);
    c->d_col_band = (const int*)(b + secs[8].off);
    c->d_col_w = (const float*)(b + secs[9].off);
    cudaSetDevice(prev);
    *out = c;
    return 0;
}

extern "C" void avse_destroy(avse_ctx* ctx) {
    if (ctx == nullptr) return;
    if (ctx->dbase) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(ctx->device);
        cudaFree(ctx->dbase);
        cudaSetDevice(prev);
    }
    delete ctx;
}

extern "C" int avse_get_filterbank(const avse_ctx* ctx, double* host_out) {
    if (ctx == nullptr || host_out == nullptr) return fail(AVSE_E_ARG, "avse_get_filterbank: NULL argument");
    memcpy(host_out, ctx->host.fb.data(), sizeof(double) * NMEL * NBINS);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// SNR factor (dp:130) -- one CTA per utterance, float64 accumulation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) avse_snr_factor_kernel(const float* __restrict__ speech, const float* __restrict__ noise,
                                                              long long stride, const int* __restrict__ lengths, int L,
                                                              const float* __restrict__ snr_db, float* __restrict__ factor_out,
                                                              int* __restrict__ max_key) {
    const int u = blockIdx.x;
    const int n = lengths ? lengths[u] : L;
    const float* s = speech + (size_t)u * stride;
    const float* z = noise + (size_t)u * stride;
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    const int n4 = ((((size_t)s | (size_t)z) & 15) == 0) ? (n >> 2) : 0;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 sv = __ldg(reinterpret_cast<const float4*>(s) + i);
        const float4 zv = __ldg(reinterpret_cast<const float4*>(z) + i);
        a0 += (double)sv.x + (double)sv.y + (double)sv.z + (double)sv.w;
        a1 += (double)sv.x * sv.x + (double)sv.y * sv.y + (double)sv.z * sv.z + (double)sv.w * sv.w;
        b0 += (double)zv.x + (double)zv.y + (double)zv.z + (double)zv.w;
        b1 += (double)zv.x * zv.x + (double)zv.y * zv.y + (double)zv.z * zv.z + (double)zv.w * zv.w;
    }
    for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
        const double sv = s[i], zv = z[i];
        a0 += sv; a1 += sv * sv; b0 += zv; b1 += zv * zv;
    }
    __shared__ double red[4][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        b0 += __shfl_xor_sync(0xffffffffu, b0, o);
        b1 += __shfl_xor_sync(0xffffffffu, b1, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = a0; red[1][w] = a1; red[2][w] = b0; red[3][w] = b1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        for (int i = 0; i < 8; ++i) { t0 += red[0][i]; t1 += red[1][i]; t2 += red[2][i]; t3 += red[3][i]; }
        const double inv = 1.0 / (double)n;
        const double ms = t0 * inv, mn = t2 * inv;
        const double vs = t1 * inv - ms * ms, vn = t3 * inv - mn * mn;
        const double db = snr_db ? (double)snr_db[u] : 0.0;
        factor_out[u] = (float)(sqrt(vs / vn) * pow(10.0, -db / 20.0));
        if (max_key) { max_key[3 * u] = (int)0x80000000; max_key[3 * u + 1] = (int)0x80000000; max_key[3 * u + 2] = (int)0x80000000; }
    }
}

extern "C" int avse_snr_factor(avse_ctx* ctx, const float* speech, const float* noise, long long stride, const int* lengths,
                               int B, int L, const float* snr_db, float* factor_out, int* max_key, void* stream) {
    if (!ctx || !speech || !noise || !factor_out) return fail(AVSE_E_ARG, "avse_snr_factor: NULL argument");
    if (B <= 0 || L <= 0 || stride < L) return fail(AVSE_E_ARG, "avse_snr_factor: bad sizes");
    avse_snr_factor_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(speech, noise, stride, lengths, L, snr_db, factor_out, max_key);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// fused forward kernel
// ---------------------------------------------------------------------------------------------
constexpr int FWD_WARPS = 8;
constexpr int FWD_THREADS = FWD_WARPS * 32;
// shared memory: per-warp frame buffers, then CTA-shared tables
constexpr int FWD_SM_MELW = FWD_WARPS * WARP_SMEM_F;              // [80][MEL_WROW]
constexpr int FWD_SM_MELLO = FWD_SM_MELW + NMEL * MEL_WROW;       // [80] int
constexpr int FWD_SM_ROUNDW = FWD_SM_MELLO + NMEL;                // [16] int
constexpr int FWD_SM_WIN = FWD_SM_ROUNDW + 16;                    // [640]
constexpr int FWD_SM_TW = FWD_SM_WIN + NFFT;                      // [16][40] vec2
constexpr int FWD_SMEM_F = FWD_SM_TW + N1 * N2 * 2;
constexpr int FWD_SMEM_BYTES = FWD_SMEM_F * 4;
static_assert((FWD_SM_MELW % 4) == 0 && (FWD_SM_TW % 2) == 0, "table alignment");
static_assert(2 * (FWD_SMEM_BYTES + 1024) <= 233472, "two CTAs per SM must fit");

struct FwdParams {
    avse_forward_args a;
    FwdTables tb;
    int T;   // frames per utterance
    int G;   // groups of 4 frames per utterance
};

template <bool STD>
__global__ void __launch_bounds__(FWD_THREADS, 2) avse_forward_kernel(const __grid_constant__ FwdParams P) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_melw = smem + FWD_SM_MELW;
    int* s_mello = reinterpret_cast<int*>(smem + FWD_SM_MELLO);
    int* s_roundw = reinterpret_cast<int*>(smem + FWD_SM_ROUNDW);
    float* s_win = smem + FWD_SM_WIN;
    vec2* s_tw = reinterpret_cast<vec2*>(smem + FWD_SM_TW);
    for (int i = threadIdx.x; i < NMEL * MEL_WROW; i += FWD_THREADS) s_melw[i] = P.tb.mel_w[i];
    for (int i = threadIdx.x; i < NMEL; i += FWD_THREADS) s_mello[i] = P.tb.mel_lo[i];
    if (threadIdx.x < MEL_ROUNDS) s_roundw[threadIdx.x] = P.tb.mel_roundw[threadIdx.x];
    for (int i = threadIdx.x; i < NFFT; i += FWD_THREADS) s_win[i] = P.tb.window[i];
    for (int i = threadIdx.x; i < N1 * N2 * 2; i += FWD_THREADS) smem[FWD_SM_TW + i] = P.tb.tw1t[i];
    float* frames = smem + warp * WARP_SMEM_F;
    // keep never-written pad slots finite (they are multiplied by exact-zero weights)
    for (int i = lane; i < WARP_SMEM_F; i += 32) frames[i] = 0.0f;
    __syncthreads();

    const avse_forward_args& A = P.a;
    const long long total = (long long)A.B * P.G;
    const long long nwarps = (long long)gridDim.x * FWD_WARPS;
    const long long per = (total + nwarps - 1) / nwarps;
    const long long gw = (long long)blockIdx.x * FWD_WARPS + warp;
    long long tile = gw * per;
    const long long tile_end = (tile + per < total) ? tile + per : total;

    int cur_u = -1;
    float mx[3] = {neg_inf(), neg_inf(), neg_inf()};
    FwdTile tl{};
    FwdOut out{};
    const bool have_noise = A.noise != nullptr;

    auto flush_max = [&](int u) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            float v = mx[s];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) atomicMax(A.max_key + 3 * u + s, float_to_key(v));
            mx[s] = neg_inf();
        }
    };

    for (; tile < tile_end; ++tile) {
        const int u = (int)(tile / P.G);
        const int g = (int)(tile - (long long)u * P.G);
        if (u != cur_u) {
            if (cur_u >= 0) flush_max(cur_u);
            cur_u = u;
            tl.sp = A.speech + (size_t)u * A.in_stride;
            tl.nz = have_noise ? A.noise + (size_t)u * A.in_stride : nullptr;
            tl.L = A.L;
            tl.T = P.T;
            int vs = A.len_speech ? A.len_speech[u] : A.L;
            int vn = A.len_noise ? A.len_noise[u] : vs;
            tl.valid_s = vs < A.L ? vs : A.L;
            tl.valid_n = vn < A.L ? vn : A.L;
            tl.vmin = have_noise ? (tl.valid_s < tl.valid_n ? tl.valid_s : tl.valid_n) : 0;
            tl.factor = have_noise ? (A.factor ? A.factor[u] : 1.0f) : 0.0f;
            tl.mixed_pcm = A.mixed_pcm ? A.mixed_pcm + (size_t)u * A.pcm_stride : nullptr;
            out.dst[0] = A.out_speech ? A.out_speech + (size_t)u * A.out_stride : nullptr;
            out.dst[1] = A.out_noise ? A.out_noise + (size_t)u * A.out_stride : nullptr;
            out.dst[2] = A.out_mixed ? A.out_mixed + (size_t)u * A.out_stride : nullptr;
            out.layout = A.layout;
            out.n_slices = A.n_slices;
            out.ld_t = A.ld_t;
        }
        tl.t0 = g * FPG;

        // ---- pass 1 ----
        stage_pass1(tl, lane, s_win, s_tw, frames);
        __syncwarp();

        // ---- pass 2 ----
        {
            float yr[40], yi[40];
            pass2_compute(lane, frames, yr, yi);
            __syncwarp();
            pass2_store(lane, frames, yr, yi);
        }
        __syncwarp();

        // ---- post ----
        {
            vec2* srow = nullptr;
            const int tf = tl.t0 + (lane >> 4);
            if (A.stft_speech != nullptr && tf < tl.T)
                srow = reinterpret_cast<vec2*>(A.stft_speech) + ((size_t)u * P.T + tf) * NBINS;
            stage_post(lane, tl.factor, frames, srow);
        }
        __syncwarp();

        // ---- mel (results staged over the now-dead frame buffer 0) ----
        {
            float acc[MEL_ROUNDS][3];
            stage_mel<STD>(lane, s_roundw, s_melw, s_mello, frames, acc);
            __syncwarp();
            stage_mel_store(lane, acc, frames);
        }
        __syncwarp();

        // ---- dB + stores ----
#pragma unroll
        for (int q = 0; q < 8; ++q) stage_db(lane, q, tl.factor, have_noise, frames, out, tl.t0, tl.T, mx);
        __syncwarp();
    }
    if (cur_u >= 0) flush_max(cur_u);
}

static bool tables_are_std(const HostTables& h) {
    const int std_w[MEL_ROUNDS] = AVSE_STD_ROUNDW;
    for (int r = 0; r < MEL_ROUNDS; ++r)
        if (h.mel_roundw[r] != std_w[r]) return false;
    return true;
}

extern "C" int avse_forward(avse_ctx* ctx, const avse_forward_args* args, void* stream) {
    if (!ctx || !args) return fail(AVSE_E_ARG, "avse_forward: NULL argument");
    const avse_forward_args& a = *args;
    if (!a.speech || !a.max_key) return fail(AVSE_E_ARG, "avse_forward: speech and max_key are required");
    if (a.B <= 0 || a.L <= HALF) return fail(AVSE_E_ARG, "avse_forward: need B > 0 and L > 320 (reflect padding)");
    if (!a.len_speech && a.in_stride < a.L) return fail(AVSE_E_ARG, "avse_forward: in_stride < L needs len_speech");
    if (a.layout != AVSE_LAYOUT_SLICES && a.layout != AVSE_LAYOUT_SPEC) return fail(AVSE_E_ARG, "avse_forward: bad layout");
    FwdParams P;
    P.a = a;
    P.tb = ctx->fwd;
    P.T = 1 + a.L / HOP;
    P.G = (P.T + FPG - 1) / FPG;
    if (a.layout == AVSE_LAYOUT_SLICES) {
        if (a.n_slices < 0 || (long long)a.n_slices * AVSE_SPSS > P.T) return fail(AVSE_E_ARG, "avse_forward: n_slices exceeds int(T/20) (dp:50)");
        if (a.out_stride < (long long)a.n_slices * NMEL * AVSE_SPSS) return fail(AVSE_E_ARG, "avse_forward: out_stride too small");
    } else {
        if (a.ld_t < P.T) return fail(AVSE_E_ARG, "avse_forward: ld_t < T");
        if (a.out_stride < (long long)NMEL * a.ld_t) return fail(AVSE_E_ARG, "avse_forward: out_stride too small");
    }
    if (a.mixed_pcm && a.pcm_stride < a.L) return fail(AVSE_E_ARG, "avse_forward: pcm_stride < L");
    float* outs[3] = {a.out_speech, a.out_noise, a.out_mixed};
    for (int s = 0; s < 3; ++s)
        if (outs[s] && (((size_t)outs[s] & 15) || (a.out_stride & 3))) return fail(AVSE_E_ARG, "avse_forward: outputs must be 16-byte aligned with out_stride % 4 == 0");

    static thread_local int configured_dev = -1;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != ctx->device) return fail(AVSE_E_ARG, "avse_forward: current device differs from the context's device");
    if (configured_dev != dev) {
        CUDA_TRY(cudaFuncSetAttribute(avse_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(avse_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES));
        configured_dev = dev;
    }
    const long long total = (long long)a.B * P.G;
    long long blocks = 2LL * ctx->num_sms;
    const long long need = (total + FWD_WARPS - 1) / FWD_WARPS;
    if (blocks > need) blocks = need;
    if (tables_are_std(ctx->host)) avse_forward_kernel<true><<<(unsigned)blocks, FWD_THREADS, FWD_SMEM_BYTES, (cudaStream_t)stream>>>(P);
    else avse_forward_kernel<false><<<(unsigned)blocks, FWD_THREADS, FWD_SMEM_BYTES, (cudaStream_t)stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// top_db floor (dp:94) in place, and floor + segment gather (dp:49-57)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) avse_floor_inplace_kernel(float* __restrict__ data, long long stride, long long n_per_utt,
                                                                 const int* __restrict__ max_key, int which) {
    const int u = blockIdx.y;
    const float thr = key_to_float(max_key[3 * u + which]) - TOP_DB;
    float* p = data + (size_t)u * stride;
    const long long n4 = n_per_utt >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4*>(p)[i];
        v.x = fmaxf(v.x, thr); v.y = fmaxf(v.y, thr); v.z = fmaxf(v.z, thr); v.w = fmaxf(v.w, thr);
        reinterpret_cast<float4*>(p)[i] = v;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_utt; i += (long long)gridDim.x * blockDim.x)
        p[i] = fmaxf(p[i], thr);
}

extern "C" int avse_floor_inplace(avse_ctx* ctx, float* data, long long stride, long long n_per_utt, int B, const int* max_key,
                                  int which, void* stream) {
    if (!ctx || !data || !max_key) return fail(AVSE_E_ARG, "avse_floor_inplace: NULL argument");
    if (B <= 0 || n_per_utt <= 0 || stride < n_per_utt || which < 0 || which > 2) return fail(AVSE_E_ARG, "avse_floor_inplace: bad sizes");
    if (((size_t)data & 15) || (stride & 3)) return fail(AVSE_E_ARG, "avse_floor_inplace: data must be 16-byte aligned with stride % 4 == 0");
    long long bx = (n_per_utt / 4 + 255) / 256;
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    dim3 grid((unsigned)bx, (unsigned)B);
    avse_floor_inplace_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(data, stride, n_per_utt, max_key, which);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

__global__ void __launch_bounds__(256) avse_floor_gather_kernel(const float* __restrict__ spec, long long spec_stride, int ld_t,
                                                                float* __restrict__ slices, long long slices_stride, int n_slices,
                                                                const int* __restrict__ max_key, int which) {
    const int u = blockIdx.y;
    const float thr = key_to_float(max_key[3 * u + which]) - TOP_DB;
    const float* src = spec + (size_t)u * spec_stride;
    float* dst = slices + (size_t)u * slices_stride;
    const int n = n_slices * NMEL * AVSE_SPSS;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = i % AVSE_SPSS;
        const int m = (i / AVSE_SPSS) % NMEL;
        const int s = i / (AVSE_SPSS * NMEL);
        dst[i] = fmaxf(src[(size_t)m * ld_t + s * AVSE_SPSS + j], thr);
    }
}

extern "C" int avse_floor_gather(avse_ctx* ctx, const float* spec, long long spec_stride, int ld_t, float* slices,
                                 long long slices_stride, int n_slices, int B, const int* max_key, int which, void* stream) {
    if (!ctx || !spec || !slices || !max_key) return fail(AVSE_E_ARG, "avse_floor_gather: NULL argument");
    if (B <= 0 || n_slices <= 0 || ld_t < n_slices * AVSE_SPSS || which < 0 || which > 2) return fail(AVSE_E_ARG, "avse_floor_gather: bad sizes");
    const int n = n_slices * NMEL * AVSE_SPSS;
    int bx = (n + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid((unsigned)bx, (unsigned)B);
    avse_floor_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(spec, spec_stride, ld_t, slices, slices_stride, n_slices, max_key, which);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

__global__ void avse_reset_max_kernel(int* __restrict__ max_key, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) max_key[i] = (int)0x80000000;
}

extern "C" int avse_reset_max(avse_ctx* ctx, int* max_key, int n, void* stream) {
    if (!ctx || !max_key || n <= 0) return fail(AVSE_E_ARG, "avse_reset_max: bad argument");
    avse_reset_max_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(max_key, n);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

__global__ void avse_max_db_kernel(const int* __restrict__ max_key, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = key_to_float(max_key[i]);
}

extern "C" int avse_max_db(avse_ctx* ctx, const int* max_key, int n, float* out_db, void* stream) {
    if (!ctx || !max_key || !out_db || n <= 0) return fail(AVSE_E_ARG, "avse_max_db: bad argument");
    avse_max_db_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(max_key, n, out_db);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
