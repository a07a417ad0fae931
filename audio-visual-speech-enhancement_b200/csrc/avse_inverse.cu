// Inverse path (predict time): dB log-mel + mixture PCM -> PCM.  See include/avse_b200.h, avse_inv_stages.cuh.
#include <cuda_runtime.h>
#include <cstdlib>
#include <string>

#include "../../include/avse_b200.h"
#include "avse_common.h"
#include "avse_ctx.h"
#include "avse_inv_stages.cuh"
#include "avse_inv8_stages.cuh"

using namespace avse;

// ---------------------------------------------------------------------------------------------
// kernel 1: c = (F F^T)^-1 10^(dB/20) per frame (Thomas, one thread per frame), coefficient layout
// work[u][band][t_pad] (time minor, zero for t >= T_use so that padded groups contribute nothing).
// Replaces np.linalg.pinv + np.dot of dp:112 together with db_to_amplitude (dp:101).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) avse_mel_to_coef_kernel(const float* __restrict__ mel_db, int layout, int ld_t,
                                                               long long mel_stride, int T_use, int T_pad,
                                                               const float* __restrict__ tri_w, const float* __restrict__ tri_ipiv,
                                                               const float* __restrict__ tri_sup, float* __restrict__ work,
                                                               long long work_stride) {
    const int u = blockIdx.x;                               // utterance on grid.x: B is not capped at 65 535
    const int t = blockIdx.y * blockDim.x + threadIdx.x;
    if (t >= T_pad) return;
    float* dst = work + (size_t)u * work_stride + t;
    if (t >= T_use) {
#pragma unroll 8
        for (int m = 0; m < NMEL; ++m) dst[(size_t)m * T_pad] = 0.0f;
        return;
    }
    const float* src = mel_db + (size_t)u * mel_stride;
    size_t base, mstride;
    if (layout == AVSE_LAYOUT_SLICES) {
        const int s = t / SPSS, tt = t - s * SPSS;
        base = (size_t)s * NMEL * SPSS + tt;
        mstride = SPSS;
    } else {
        base = t;
        mstride = ld_t;
    }
    constexpr float K = 0.16609640474436813f;   // log2(10) / 20 :  10^(dB/20) = 2^(K dB)   (librosa.db_to_amplitude)
    float d[NMEL];
#pragma unroll
    for (int m = 0; m < NMEL; ++m) d[m] = exp2f(K * src[base + (size_t)m * mstride]);
#pragma unroll
    for (int m = 1; m < NMEL; ++m) d[m] -= __ldg(tri_w + m) * d[m - 1];
    d[NMEL - 1] *= __ldg(tri_ipiv + NMEL - 1);
#pragma unroll
    for (int m = NMEL - 2; m >= 0; --m) d[m] = (d[m] - __ldg(tri_sup + m) * d[m + 1]) * __ldg(tri_ipiv + m);
#pragma unroll
    for (int m = 0; m < NMEL; ++m) dst[(size_t)m * T_pad] = d[m];
}

// ---------------------------------------------------------------------------------------------
// kernel 2: phase of the mixture, lin * phase, packed inverse FFT, overlap-add in registers
// ---------------------------------------------------------------------------------------------
// Code-size switches (A/B-measured on B200, profiles/README.md round 2): the hot loop exceeds the 32 KB L1.5 instruction cache.
#if !defined(AVSE_INV_SHARED_DFT40)
#define AVSE_INV_SHARED_DFT40 1     // pass 2 and pass A share one DFT-40 copy (rolled two-phase loop)
#endif
#if !defined(AVSE_INV_ROLLED_PASSB)
#define AVSE_INV_ROLLED_PASSB 1     // pass B main / side rounds share one column (DFT-16 + window) copy
#endif
constexpr int INV_WARPS = 6;
constexpr int INV_THREADS = INV_WARPS * 32;
constexpr int INV_SM_WIN = INV_WARPS * INV_WARP_SMEM_F;      // [640]
constexpr int INV_SM_TW = INV_SM_WIN + 2 * NFFT;             // [16][40] vec2   W^{n2 k1}, n2 minor  (window: (w, w) pairs)
constexpr int INV_SM_TWT = INV_SM_TW + N1 * N2 * 2;          // [40][16] vec2   W^{n1' k2'}, n1' minor
constexpr int INV_SM_COL = INV_SM_TWT + N1 * N2 * 2;         // [SCAN_BINS] ivec4
constexpr int INV_SMEM_F = INV_SM_COL + SCAN_BINS * 4;
constexpr int INV_SMEM_BYTES = INV_SMEM_F * 4;
static_assert((INV_SM_COL % 4) == 0 && (INV_SM_TW % 2) == 0, "table alignment");
static_assert(2 * (INV_SMEM_BYTES + 1024) <= 233472, "two CTAs per SM must fit");

struct InvParams {
    avse_inverse_args a;
    const float* window;
    const float* tw1t;
    const int* col_band;
    const float* col_w;
    const float* spike;
    const unsigned* post_mask;   // I8 post stage, walk form (avse_tables.h)
    const float* post_w;
    int T;            // mixture STFT frames
    int T_use;        // frames reconstructed
    int T_pad;        // coefficient row length (multiple of 4, >= 4 (G + 1))
    int G;            // groups of 4 frames
    int chunks;       // chunks per utterance
    int cg;           // groups per chunk
    int out_len;      // 160 (T_use - 1)
    long long total_groups;   // I8 kernel: B * G groups, cut into one contiguous range per warp
    WarpSplit split;          // I8 kernel: balanced contiguous ranges (avse_common.h)
};

// EXT: explicit phase array (dp:99 signature) instead of the recomputed mixture STFT; a separate instantiation keeps each
// kernel's code (instruction-cache footprint) small.
template <bool EXT, typename O>
__global__ void __launch_bounds__(INV_THREADS, 2) avse_inverse_kernel(const __grid_constant__ InvParams P) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * NFFT; i += INV_THREADS) smem[INV_SM_WIN + i] = P.window[i];
    for (int i = threadIdx.x; i < N1 * N2; i += INV_THREADS) {
        const int k1 = i / N2, n2 = i - k1 * N2;
        const float re = P.tw1t[2 * i], im = P.tw1t[2 * i + 1];
        smem[INV_SM_TW + 2 * i] = re; smem[INV_SM_TW + 2 * i + 1] = im;
        smem[INV_SM_TWT + 2 * (n2 * N1 + k1)] = re; smem[INV_SM_TWT + 2 * (n2 * N1 + k1) + 1] = im;   // symmetric in (n2, k1)
    }
    for (int i = threadIdx.x; i < SCAN_BINS; i += INV_THREADS) {
        int* e = reinterpret_cast<int*>(smem + INV_SM_COL) + 4 * i;
        if (i < NBINS) {
            e[0] = P.col_band[2 * i]; e[1] = P.col_band[2 * i + 1];
            e[2] = __float_as_int(P.col_w[2 * i]); e[3] = __float_as_int(P.col_w[2 * i + 1]);
        } else { e[0] = 0; e[1] = 0; e[2] = 0; e[3] = 0; }
    }
    float* frames = smem + warp * INV_WARP_SMEM_F;
    float* ybuf = frames + 2 * FRAME_F;
    float* side = ybuf + INV_Y_F;
    for (int i = lane; i < INV_WARP_SMEM_F; i += 32) frames[i] = 0.0f;
    __syncthreads();

    const float* s_win = smem + INV_SM_WIN;
    const vec2* s_tw = reinterpret_cast<const vec2*>(smem + INV_SM_TW);
    const vec2* s_twT = reinterpret_cast<const vec2*>(smem + INV_SM_TWT);
    const ivec4* s_col = reinterpret_cast<const ivec4*>(smem + INV_SM_COL);
    const avse_inverse_args& A = P.a;

    const int n_items = A.B * P.chunks;
    const int n_warps = gridDim.x * INV_WARPS;
#pragma unroll 1
    for (int item = blockIdx.x * INV_WARPS + warp; item < n_items; item += n_warps) {
        const int u = item / P.chunks;
        const int c = item - u * P.chunks;
        const int g0 = c * P.cg;
        if (g0 >= P.G && !(c == 0)) continue;
        int g1 = g0 + P.cg;
        const bool last_chunk = g1 >= P.G;
        if (last_chunk) g1 = P.G;
        const int g_first = g0 > 0 ? g0 - 1 : 0;          // warm-up group rebuilds the overlap-add carry
        const int g_last = last_chunk ? P.G : g1 - 1;     // the last chunk also drains the carry (group G)

        InvTile tl;
        tl.pcm = A.mixed_pcm + (size_t)u * A.pcm_stride;
        tl.L = A.L;
        int valid = A.len_pcm ? A.len_pcm[u] : A.L;
        tl.valid = valid < A.L ? valid : A.L;
        tl.T = P.T;
        tl.T_use = P.T_use;
        const float* ycoef = A.work + (size_t)u * A.work_stride;
        O* out = static_cast<O*>(A.out_pcm) + (size_t)u * A.out_stride;

        float acc[INV_SIDE_ROWS];
#pragma unroll
        for (int J = 0; J < INV_SIDE_ROWS; ++J) acc[J] = 0.0f;
        for (int i = lane; i < INV_SIDE_F; i += 32) side[i] = 0.0f;
        __syncwarp();

#pragma unroll 1
        for (int g = g_first; g <= g_last; ++g) {
            tl.t0 = g * INV_FPG;
            if (tl.t0 < P.T_use) {
                // coefficients of the 4 frames -> ybuf[band][4]: loaded now, stored after pass 1 so that their latency is
                // covered by pass 1 instead of stalling the warp at the top of the group
                float4 cf[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int m = lane + 32 * q;
                    if (m < NMEL) cf[q] = *reinterpret_cast<const float4*>(ycoef + (size_t)m * P.T_pad + tl.t0);
                }
                if (!EXT) {
                    // L2 prefetch of the mixture samples first touched by the next group (one 128-byte line per lane)
                    if (g < g_last) {
                        const int i0 = (((tl.t0 + INV_FPG) * HOP + HALF - HOP) & ~31) + 32 * lane;
                        if (lane < 22 && i0 < tl.valid) asm volatile("prefetch.global.L2 [%0];" ::"l"(tl.pcm + i0));
                    }
                    inv_stage_pass1(tl, lane, s_win, s_tw, frames);
                }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int m = lane + 32 * q;
                    if (m < NMEL) *reinterpret_cast<float4*>(ybuf + 4 * m) = cf[q];
                }
#if AVSE_INV_SHARED_DFT40
                __syncwarp();
                if (EXT) {
                    const int tA = tl.t0 + 2 * (lane & 1);
                    const vec2* ph = reinterpret_cast<const vec2*>(A.phase) + (size_t)u * A.phase_stride;
                    const vec2* phA = tA < P.T_use ? ph + (size_t)tA * NBINS : nullptr;
                    const vec2* phB = tA + 1 < P.T_use ? ph + (size_t)(tA + 1) * NBINS : nullptr;
                    inv_stage_post<true>(lane, s_col, ybuf, frames, phA, phB);
                    __syncwarp();
                }
                // pass 2 (phase 0: rows -> Z, then the post stage) and pass A (phase 1: V -> rows) through ONE copy of the
                // DFT-40 codelet
#pragma unroll 1
                for (int phs = EXT ? 1 : 0; phs < 2; ++phs) {
                    cpx x[40];
                    if (phs == 0) pass2_load(lane, frames, x);
                    else inv_passA_load(lane, frames, x);
                    dft40_inplace(x);
                    if (phs != 0) inv_passA_twiddle(lane, s_twT, x);
                    __syncwarp();
                    if (phs == 0) pass2_store(lane, frames, x);
                    else inv_passA_store(lane, frames, x);
                    __syncwarp();
                    if (phs == 0) {
                        inv_stage_post<false>(lane, s_col, ybuf, frames, nullptr, nullptr);
                        __syncwarp();
                    }
                }
#else
                if (!EXT) {
                    __syncwarp();
                    {
                        cpx x[40];
                        pass2_compute(lane, frames, x);
                        __syncwarp();
                        pass2_store(lane, frames, x);
                    }
                    __syncwarp();
                    inv_stage_post<false>(lane, s_col, ybuf, frames, nullptr, nullptr);
                } else {
                    __syncwarp();
                    const int tA = tl.t0 + 2 * (lane & 1);
                    const vec2* ph = reinterpret_cast<const vec2*>(A.phase) + (size_t)u * A.phase_stride;
                    const vec2* phA = tA < P.T_use ? ph + (size_t)tA * NBINS : nullptr;
                    const vec2* phB = tA + 1 < P.T_use ? ph + (size_t)(tA + 1) * NBINS : nullptr;
                    inv_stage_post<true>(lane, s_col, ybuf, frames, phA, phB);
                }
                __syncwarp();
                {
                    cpx x[40];
                    inv_passA_compute(lane, s_twT, frames, x);
                    __syncwarp();
                    inv_passA_store(lane, frames, x);
                }
                __syncwarp();
#endif
#if AVSE_INV_ROLLED_PASSB
#pragma unroll 1
                for (int r = 0; r < 4; ++r) {
                    inv_stage_passB_round(lane, r, s_win, frames, acc, side);
                    if (r >= 2) __syncwarp();
                }
#else
                inv_stage_passB_main(lane, s_win, frames, acc);
#pragma unroll 1
                for (int ph = 0; ph < 2; ++ph) {      // rolled: one copy of the column code for both side phases
                    inv_stage_passB_side(lane, ph, s_win, frames, side);
                    __syncwarp();
                }
#endif
            }
            const bool write = g >= g0;
            inv_stage_emit_main(lane, tl.t0, P.T_use, P.out_len, write, s_win, out, acc);
            {
                float keep[4], carry[4];
                inv_stage_emit_side(lane, tl.t0, P.T_use, P.out_len, write, s_win, out, side, keep, carry);
                inv_stage_rotate_side(lane, side, carry);
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// I8 kernel: eight real frames (four packed FFTs) per group, 8 warps per SM.  See avse_inv8_stages.cuh.
// ---------------------------------------------------------------------------------------------
#if !defined(AVSE_I8_PIPE_MEL)
#define AVSE_I8_PIPE_MEL 0       // 1: the next group's dB values are loaded a whole group ahead (20 more loop-carried registers)
#endif
constexpr int I8_WARPS = 8;
constexpr int I8_THREADS = I8_WARPS * 32;
constexpr int I8_SM_WIN = I8_WARPS * I8_WARP_SMEM_F;         // [640] window
constexpr int I8_SM_TW = I8_SM_WIN + NFFT;                   // [16][40] vec2   W^{n2 k1}, n2 minor
constexpr int I8_SM_TWT = I8_SM_TW + N1 * N2 * 2;            // [40][16] vec2   W^{n1' k2'}, n1' minor
constexpr int I8_SM_COL = I8_SM_TWT + N1 * N2 * 2;           // [SCAN4_BINS] ivec4 (b0, b1, w0 / 640, w1 / 640)
constexpr int I8_SM_SPK = I8_SM_COL + SCAN4_BINS * 4;        // [4][SPIKE_ROW] partitioned tridiagonal solve (avse_tables.h)
constexpr int I8_SMEM_F = I8_SM_SPK + SPIKE_P * SPIKE_ROW;
constexpr int I8_SMEM_BYTES = I8_SMEM_F * 4;
static_assert((I8_SM_COL % 4) == 0 && (I8_SM_TW % 2) == 0 && (I8_WARP_SMEM_F % 4) == 0, "table alignment");
static_assert(I8_SMEM_BYTES + 1024 <= 232448, "I8 shared memory must fit in one SM");

#ifndef AVSE_I8_PROLOGUE_OUT_OF_LINE
#define AVSE_I8_PROLOGUE_OUT_OF_LINE 1
#endif
#if AVSE_I8_PROLOGUE_OUT_OF_LINE
#define AVSE_I8_PROLOGUE_Q __device__ __noinline__
#else
#define AVSE_I8_PROLOGUE_Q __device__ __forceinline__
#endif
// Kernel start: the CTA's tables, every global load in flight before the first store (table_fetch, avse_common.h), the warp's
// buffers zeroed meanwhile.
struct I8TablePtrs {      // by value: taking the address of the kernel's parameter block would make every later access of it a load
    const float* window; const float* tw1t; const float* spike; const float* post_w; const unsigned* post_mask;
    const int* col_band; const float* col_w;
};
AVSE_I8_PROLOGUE_Q void i8_fill_tables(const I8TablePtrs P, float* smem, float* frames) {
    const int tid = threadIdx.x, lane = threadIdx.x & 31;
    static_assert((I8_SM_WIN % 4) == 0 && (I8_SM_TW % 4) == 0 && (I8_SM_COL % 4) == 0 && (I8_SM_SPK % 4) == 0 && (I8_WARP_SMEM_F % 4) == 0 &&
                  ((SPIKE_P * SPIKE_ROW) % 4) == 0, "16-byte table copies");
    TableRegs<NFFT * 2 / 4, I8_THREADS> r_win;              // P.window holds (w, w) pairs
    TableRegs<N1 * N2 * 2 / 4, I8_THREADS> r_tw;
    TableRegs<SPIKE_P * SPIKE_ROW / 4, I8_THREADS> r_spk;
    table_fetch(P.window, tid, r_win);
    table_fetch(P.tw1t, tid, r_tw);
    table_fetch(P.spike, tid, r_spk);
#if AVSE_I8_POST_WALK
    TableRegs<SCAN4_BINS * 2 / 4, I8_THREADS> r_pw;         // (w0, w1) per bin; then per chunk (mask lo, mask hi, first band, -)
    TableRegs<8, I8_THREADS> r_pm;
    table_fetch(P.post_w, tid, r_pw);
    table_fetch(P.post_mask, tid, r_pm);
#endif
    for (int i = lane; i < I8_WARP_SMEM_F / 4; i += 32) reinterpret_cast<float4*>(frames)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < TableRegs<NFFT * 2 / 4, I8_THREADS>::K; ++k) {
        const int i = tid + k * I8_THREADS;
        if (i < NFFT * 2 / 4) { smem[I8_SM_WIN + 2 * i] = r_win.v[k].x; smem[I8_SM_WIN + 2 * i + 1] = r_win.v[k].z; }
    }
    table_put(smem + I8_SM_TW, tid, r_tw);
    table_put(smem + I8_SM_SPK, tid, r_spk);
#if !AVSE_I8_TW_IN_B
#pragma unroll
    for (int k = 0; k < TableRegs<N1 * N2 * 2 / 4, I8_THREADS>::K; ++k) {      // transposed copy [40][16] for pass A's own twiddles
        const int i4 = tid + k * I8_THREADS;
        if (i4 < N1 * N2 * 2 / 4) {
            const int i = 2 * i4, k1 = i / N2, n2 = i - k1 * N2;             // entries i and i + 1 (same k1: N2 is even)
            smem[I8_SM_TWT + 2 * (n2 * N1 + k1)] = r_tw.v[k].x; smem[I8_SM_TWT + 2 * (n2 * N1 + k1) + 1] = r_tw.v[k].y;
            smem[I8_SM_TWT + 2 * ((n2 + 1) * N1 + k1)] = r_tw.v[k].z; smem[I8_SM_TWT + 2 * ((n2 + 1) * N1 + k1) + 1] = r_tw.v[k].w;
        }
    }
#endif
#if AVSE_I8_POST_WALK
    table_put_scaled(smem + I8_SM_COL, tid, r_pw, INV_SCALE);       // irfft's 1 / 640 folded in
    table_put(smem + I8_SM_COL + 2 * SCAN4_BINS, tid, r_pm);
#else
    for (int i = tid; i < SCAN4_BINS; i += I8_THREADS) {
        int* e = reinterpret_cast<int*>(smem + I8_SM_COL) + 4 * i;
        if (i < NBINS) {
            e[0] = P.col_band[2 * i]; e[1] = P.col_band[2 * i + 1];
            e[2] = __float_as_int(P.col_w[2 * i] * INV_SCALE); e[3] = __float_as_int(P.col_w[2 * i + 1] * INV_SCALE);
        } else { e[0] = 0; e[1] = 0; e[2] = 0; e[3] = 0; }
    }
#endif
}

template <bool EXT, typename O>
__global__ void __launch_bounds__(I8_THREADS, 1) avse_inverse8_kernel(const __grid_constant__ InvParams P) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* frames = smem + warp * I8_WARP_SMEM_F;
    float* ybuf = frames + I8_NC * FRAME4_F;
    float* side = ybuf + I8_Y_F;
    float* xch = side + 2 * I8_SIDE_F;
    i8_fill_tables(I8TablePtrs{P.window, P.tw1t, P.spike, P.post_w, P.post_mask, P.col_band, P.col_w}, smem, frames);
    __syncthreads();

    const float* s_win = smem + I8_SM_WIN;
    const vec2* s_tw = reinterpret_cast<const vec2*>(smem + I8_SM_TW);
#if !AVSE_I8_TW_IN_B
    const vec2* s_twT = reinterpret_cast<const vec2*>(smem + I8_SM_TWT);
#endif
    const ivec4* s_col = reinterpret_cast<const ivec4*>(smem + I8_SM_COL);
    const float* s_spk = smem + I8_SM_SPK;
    const avse_inverse_args& A = P.a;
    Lane4Const lc;
    lane4_const_init(lane, s_win, s_tw, lc);

    // Work distribution: the B * G groups of the launch, in (utterance, group) order, are cut into ONE contiguous range per warp
    // (like the forward kernel's tiles).  A range that starts inside an utterance first recomputes the group before it (no
    // stores) to rebuild the overlap-add carry; a range that ends an utterance also drains the carry (group G).  Against
    // whole-utterance items this removes the idle SMs of a 1 000-utterance launch (1 000 items on 1 184 warps: 23 SMs had no work).
    long long tile, n_groups;
    warp_split_range(P.split, (int)blockIdx.x, warp, I8_WARPS, tile, n_groups);
    const long long tile_end = tile + n_groups;
#pragma unroll 1
    while (tile < tile_end) {
        const int u = (int)(tile / P.G);
        const int g0 = (int)(tile - (long long)u * P.G);
        int g1 = g0 + (int)(tile_end - tile);
        const bool last_chunk = g1 >= P.G;
        if (last_chunk) g1 = P.G;
        tile += g1 - g0;
        const int g_first = g0 > 0 ? g0 - 1 : 0;          // warm-up group rebuilds the overlap-add carry
        const int g_last = last_chunk ? P.G : g1 - 1;     // the end of an utterance also drains the carry (group G)

        InvTile tl;
        tl.pcm = A.mixed_pcm + (size_t)u * A.pcm_stride;
        tl.L = A.L;
        int valid = A.len_pcm ? A.len_pcm[u] : A.L;
        tl.valid = valid < 0 ? 0 : (valid < A.L ? valid : A.L);
        tl.T = P.T;
        tl.T_use = P.T_use;
        const float* mel = A.mel_db + (size_t)u * A.mel_stride;
        O* out = static_cast<O*>(A.out_pcm) + (size_t)u * A.out_stride;

        float acc[I8_ACC];
#pragma unroll
        for (int J = 0; J < I8_ACC; ++J) acc[J] = 0.0f;
        for (int i = lane; i < 2 * I8_SIDE_F; i += 32) side[i] = 0.0f;
        __syncwarp();

        // software pipeline: the samples of the next computed group are loaded before pass B of the current one
        float raw[I8_RAW], rt[20], cd[SPIKE_Q];
        bool interior = false;
        tl.t0 = g_first * I8_FPG;
        if (tl.t0 < P.T_use) {
#if AVSE_I8_PIPE_MEL
            i8_coef_load(lane, mel, A.layout, A.ld_t, tl.t0, P.T_use, cd);
#endif
            if (!EXT) {
                interior = i8_group_interior(tl);
                if (interior) { i8_load_raw(tl, lane, raw); i8_load_tail_raw(tl, lane, rt); }
                else if (AVSE_I8_REFLECT_FAST && i8_group_reflect_only(tl)) {      // mirrored loads, interior pass 1 (avse_inv8_stages.cuh)
                    i8_load_raw_reflect(tl, lane, raw); i8_load_tail_raw_reflect(tl, lane, rt);
                    interior = true;
                }
            }
        }

#pragma unroll 1
        for (int g = g_first; g <= g_last; ++g) {
            tl.t0 = g * I8_FPG;
            const bool have = tl.t0 < P.T_use;
            const bool write = g >= g0;
            const float* side_in = side + I8_SIDE_F * (g & 1);
            float* side_out = side + I8_SIDE_F * ((g & 1) ^ 1);
            // FFTs [c_lo, c_hi) of the group are needed: a warm-up group only rebuilds the carry, which FFTs 0 and 1 (rows 0..27 of the
            // group) cannot reach; FFTs whose frames lie beyond T_use have zero coefficients.  Skipped FFTs add exact zeros.
#if AVSE_I8_SKIP_FFTS
            const int c_lo = write ? 0 : 2;
            const int c_left = (P.T_use - tl.t0 + 1) >> 1;
            const int c_hi = c_left < I8_NC ? c_left : I8_NC;
#else
            const int c_lo = 0, c_hi = I8_NC;
#endif
            if (have) {
                // coefficients of the 8 frames -> ybuf[band][I8_YS] (i8_coef_*): the 20 dB values of this lane are loaded ahead of
                // pass 1 (AVSE_I8_PIPE_MEL: one whole group ahead, with the samples); the partitioned solve runs after pass 1
#if !AVSE_I8_PIPE_MEL
                i8_coef_load(lane, mel, A.layout, A.ld_t, tl.t0, P.T_use, cd);
#endif
                if (!EXT) {
                    if (interior) {
                        i8_pass1_main(lane, raw, lc, frames, c_lo, c_hi);
                        i8_pass1_tail(lane, rt, s_win, s_tw, frames);
                    } else {
                        i8_pass1_edge(tl, lane, s_win, s_tw, frames, c_lo, c_hi);
                    }
                }
                i8_coef_local(lane, s_spk, cd, xch);
                __syncwarp();
                i8_coef_finish(lane, s_spk, cd, xch, ybuf);
                __syncwarp();
                if (EXT) {
                    const int tA = tl.t0 + 2 * (lane >> 3);
                    const vec2* ph = reinterpret_cast<const vec2*>(A.phase) + (size_t)u * A.phase_stride;
                    const vec2* phA = tA < P.T_use ? ph + (size_t)tA * NBINS : nullptr;
                    const vec2* phB = tA + 1 < P.T_use ? ph + (size_t)(tA + 1) * NBINS : nullptr;
                    i8_stage_post<true>(lane, s_col, ybuf, frames, phA, phB);
                    __syncwarp();
                }
                // pass 2 (iterations 0, 1: rows -> Z), the post stage, pass A (iterations 2, 3: V -> rows): ONE DFT-40 copy
#pragma unroll 1
                for (int it = EXT ? 2 : 0; it < 4; ++it) {
                    const int r = it & 1;
                    if (it == 2 && !EXT) {
                        i8_stage_post<false>(lane, s_col, ybuf, frames, nullptr, nullptr);
                        __syncwarp();
                    }
                    if (2 * r + 1 < c_lo || 2 * r >= c_hi) continue;      // neither FFT of this round is needed
                    cpx x[40];
                    if (it < 2) p4_pass2_load(lane, r, frames, x);
                    else i8_passA_load(lane, r, frames, x);
                    dft40_inplace(x);
#if !AVSE_I8_TW_IN_B
                    if (it >= 2) inv_passA_twiddle(lane, s_twT, x);
#endif
                    __syncwarp();
                    if (it < 2) p4_pass2_store(lane, r, frames, x);
                    else i8_passA_store(lane, r, frames, x);
                    __syncwarp();
                }
                if (c_lo > 0 || c_hi < I8_NC) i8_rearm_flags(lane, frames);
            }
            // ---- next computed group: issue its loads now (they land during pass B / emit) ----
            bool interior2 = false;
            if (g < g_last) {
                InvTile tn = tl;
                tn.t0 = (g + 1) * I8_FPG;
                if (tn.t0 < P.T_use) {
#if AVSE_I8_PIPE_MEL
                    i8_coef_load(lane, mel, A.layout, A.ld_t, tn.t0, P.T_use, cd);
#endif
                    if (!EXT) {
                        interior2 = i8_group_interior(tn);
                        if (interior2) { i8_load_raw(tn, lane, raw); i8_load_tail_raw(tn, lane, rt); }
                        else if (AVSE_I8_REFLECT_FAST && i8_group_reflect_only(tn)) {
                            i8_load_raw_reflect(tn, lane, raw); i8_load_tail_raw_reflect(tn, lane, rt);
                            interior2 = true;
                        }
                    }
                }
            }
            // ---- pass B + emit: one FFT at a time (rolled), two finished hops leave after each ----
#pragma unroll 1
            for (int cc = 0; cc < I8_NC; ++cc) {
                if (have && cc >= c_lo && cc < c_hi) i8_passB_add(lane, cc, lc, frames, acc);
                i8_emit_main(lane, tl.t0 + 2 * cc, P.T_use, P.out_len, write, s_win, out, acc);
            }
            if (have) i8_passB_tail(lane, s_win, s_tw, frames, ybuf, c_lo, c_hi);
            __syncwarp();
            i8_tail_reduce_emit(lane, tl.t0, P.T_use, P.out_len, write, have, s_win, out, ybuf, side_in, side_out);
            __syncwarp();
            interior = interior2;
        }
    }
}

// Scratch per utterance: the coefficient rows [80][T_pad], T_pad = frames padded to whole groups plus one drain group.  Sized for
// the 8-frame groups of the I8 kernel, which also covers the 4-frame kernel's padding.
extern "C" int avse_inverse_work_elems(int n_frames_use, long long* per_utterance) {
    if (n_frames_use <= 0 || per_utterance == nullptr) return avse_fail(AVSE_E_ARG, "avse_inverse_work_elems: bad argument");
    const int G = (n_frames_use + I8_FPG - 1) / I8_FPG;
    *per_utterance = (long long)NMEL * (I8_FPG * (G + 1));
    return 0;
}

// Scratch floats per utterance for `ctx`'s geometry: the generic path keeps every windowed output frame ([T][n_inv]).
extern "C" int avse_inverse_work_elems_ctx(const avse_ctx* ctx, int n_frames_use, long long* per_utterance) {
    if (!ctx) return avse_fail(AVSE_E_ARG, "avse_inverse_work_elems_ctx: NULL context");
    if (!ctx->generic) return avse_inverse_work_elems(n_frames_use, per_utterance);
    if (n_frames_use <= 0 || per_utterance == nullptr) return avse_fail(AVSE_E_ARG, "avse_inverse_work_elems_ctx: bad argument");
    *per_utterance = ((long long)n_frames_use * ctx->gen.geo.n_inv + 3) / 4 * 4;
    return 0;
}

extern "C" int avse_inverse(avse_ctx* ctx, const avse_inverse_args* args, void* stream) {
    if (!ctx || !args) return avse_fail(AVSE_E_ARG, "avse_inverse: NULL argument");
    if (ctx->generic) return avse_generic_inverse(ctx, args, stream);
    const avse_inverse_args& a = *args;
    if (!a.mel_db || (!a.mixed_pcm && !a.phase) || !a.out_pcm) return avse_fail(AVSE_E_ARG, "avse_inverse: NULL buffer");
    if (a.B <= 0 || (!a.phase && a.L <= HALF)) return avse_fail(AVSE_E_ARG, "avse_inverse: need B > 0 and L > 320");
    if (a.layout != AVSE_LAYOUT_SLICES && a.layout != AVSE_LAYOUT_SPEC) return avse_fail(AVSE_E_ARG, "avse_inverse: bad layout");
    if (a.out_format != AVSE_SAMPLE_F32 && a.out_format != AVSE_SAMPLE_I16) return avse_fail(AVSE_E_ARG, "avse_inverse: bad out_format");
    InvParams P;
    P.a = a;
    P.T = a.phase ? a.phase_frames : 1 + a.L / HOP;
    const int t_mel = a.layout == AVSE_LAYOUT_SLICES ? a.n_slices * AVSE_SPSS : a.n_frames;
    if (t_mel <= 0) return avse_fail(AVSE_E_ARG, "avse_inverse: no mel frames");
    P.T_use = t_mel < P.T ? t_mel : P.T;                       // dp:68
    if (P.T_use < 2) return avse_fail(AVSE_E_ARG, "avse_inverse: fewer than 2 frames gives an empty signal");
    // kernel choice: the I8 kernel (8 frames per group, 8 warps / SM); AVSE_INV4=1 keeps the round-1 4-frame kernel (A/B runs, tests)
    static const bool force4 = [] { const char* e = getenv("AVSE_INV4"); return e != nullptr && e[0] == '1'; }();
    const bool use8 = !force4;
    const int fpg = use8 ? I8_FPG : INV_FPG;
    P.G = (P.T_use + fpg - 1) / fpg;
    P.T_pad = fpg * (P.G + 1);
    P.out_len = HOP * (P.T_use - 1);
    if (a.layout == AVSE_LAYOUT_SPEC && a.ld_t < P.T_use) return avse_fail(AVSE_E_ARG, "avse_inverse: ld_t < frames used");
    if (!use8) {      // scratch is only used by the 4-frame kernel (the I8 kernel keeps the coefficients on chip)
        if (!a.work) return avse_fail(AVSE_E_ARG, "avse_inverse: NULL buffer");
        if (a.work_stride < (long long)NMEL * P.T_pad) return avse_fail(AVSE_E_ARG, "avse_inverse: work_stride too small (see avse_inverse_work_elems)");
        if (((size_t)a.work & 15) || (a.work_stride & 3)) return avse_fail(AVSE_E_ARG, "avse_inverse: work must be 16-byte aligned with work_stride % 4 == 0");
    }
    if (a.out_stride < P.out_len) return avse_fail(AVSE_E_ARG, "avse_inverse: out_stride < 160 (T_use - 1)");
    if (a.phase && a.phase_stride < (long long)P.T_use * NBINS) return avse_fail(AVSE_E_ARG, "avse_inverse: phase_stride too small");
    if (!a.phase && !a.len_pcm && a.pcm_stride < a.L) return avse_fail(AVSE_E_ARG, "avse_inverse: pcm_stride < L needs len_pcm");
    P.window = ctx->fwd.window2;   // (w, w) pairs
    P.tw1t = ctx->fwd.tw1t;
    P.col_band = ctx->d_col_band;
    P.col_w = ctx->d_col_w;
    P.spike = ctx->d_spike;
    P.post_mask = ctx->d_post_mask;
    P.post_w = ctx->d_post_w;

    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != ctx->device) return avse_fail(AVSE_E_ARG, "avse_inverse: current device differs from the context's device");
    static thread_local int configured_dev = -1;
    if (configured_dev != dev) {
        CUDA_TRY(cudaFuncSetAttribute(avse_inverse_kernel<false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, INV_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(avse_inverse_kernel<true, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, INV_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(avse_inverse_kernel<false, short>, cudaFuncAttributeMaxDynamicSharedMemorySize, INV_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(avse_inverse_kernel<true, short>, cudaFuncAttributeMaxDynamicSharedMemorySize, INV_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(avse_inverse8_kernel<false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(avse_inverse8_kernel<true, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(avse_inverse8_kernel<false, short>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(avse_inverse8_kernel<true, short>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
        configured_dev = dev;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (!use8) {      // the I8 kernel solves for the coefficients itself (i8_coef_*); only the 4-frame kernel needs this pass
        dim3 grid((unsigned)a.B, (unsigned)((P.T_pad + 127) / 128));
        avse_mel_to_coef_kernel<<<grid, 128, 0, st>>>(a.mel_db, a.layout, a.ld_t, a.mel_stride, P.T_use, P.T_pad, ctx->d_tri_w,
                                                      ctx->d_tri_ipiv, ctx->d_tri_sup, a.work, a.work_stride);
        CUDA_TRY(cudaGetLastError());
    }
    // chunking: a warp streams through one (utterance, chunk) item at a time; every chunk but the first pays one
    // recomputed warm-up group, and the launch ends when the warp with the most items finishes.  Pick the chunk count
    // that minimises  rounds x (groups per chunk + warm-up)  with rounds = ceil(items / resident warps).
    const long long n_warps = use8 ? (long long)ctx->num_sms * I8_WARPS : 2LL * ctx->num_sms * INV_WARPS;
    const int min_cg = use8 ? 2 : 4;                                       // smallest chunk worth its warm-up group
    const long long max_chunks = P.G / min_cg > 0 ? P.G / min_cg : 1;
    long long best_cost = -1;
    int best_cg = P.G;
    for (long long c = 1; c <= max_chunks; ++c) {
        const int cg = (int)((P.G + c - 1) / c);
        const long long chunks = (P.G + cg - 1) / cg;
        const long long rounds = ((long long)a.B * chunks + n_warps - 1) / n_warps;
        const long long cost = rounds * (cg + (chunks > 1 ? 2 : 1));   // + warm-up group, + drain / partial group
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_cg = cg; }
    }
    P.cg = best_cg;
    P.chunks = (P.G + P.cg - 1) / P.cg;
    const bool o16 = a.out_format == AVSE_SAMPLE_I16;
    if (use8) {
        P.total_groups = (long long)a.B * P.G;
        long long blocks = ctx->num_sms;
        const long long need = (P.total_groups + I8_WARPS - 1) / I8_WARPS;
        if (blocks > need) blocks = need;
        P.split = make_warp_split(P.total_groups, blocks, I8_WARPS);
        if (a.phase) {
            if (o16) avse_inverse8_kernel<true, short><<<(unsigned)blocks, I8_THREADS, I8_SMEM_BYTES, st>>>(P);
            else avse_inverse8_kernel<true, float><<<(unsigned)blocks, I8_THREADS, I8_SMEM_BYTES, st>>>(P);
        } else {
            if (o16) avse_inverse8_kernel<false, short><<<(unsigned)blocks, I8_THREADS, I8_SMEM_BYTES, st>>>(P);
            else avse_inverse8_kernel<false, float><<<(unsigned)blocks, I8_THREADS, I8_SMEM_BYTES, st>>>(P);
        }
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    long long blocks = 2LL * ctx->num_sms;
    const long long need = ((long long)a.B * P.chunks + INV_WARPS - 1) / INV_WARPS;
    if (blocks > need) blocks = need;
    if (a.phase) {
        if (o16) avse_inverse_kernel<true, short><<<(unsigned)blocks, INV_THREADS, INV_SMEM_BYTES, st>>>(P);
        else avse_inverse_kernel<true, float><<<(unsigned)blocks, INV_THREADS, INV_SMEM_BYTES, st>>>(P);
    } else {
        if (o16) avse_inverse_kernel<false, short><<<(unsigned)blocks, INV_THREADS, INV_SMEM_BYTES, st>>>(P);
        else avse_inverse_kernel<false, float><<<(unsigned)blocks, INV_THREADS, INV_SMEM_BYTES, st>>>(P);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
