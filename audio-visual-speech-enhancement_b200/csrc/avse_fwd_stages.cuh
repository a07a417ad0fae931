// Warp-stage functions of the fused forward kernel (mix at SNR -> STFT -> mel -> dB).
//
// One warp owns a group of FPG = 2 consecutive STFT frames of one utterance and walks them
// through five warp-synchronous stages that exchange data only through that warp's private
// shared-memory region (no block-level barrier anywhere in the frame loop):
//
//   pass 1  lane = (frame, n2):   strided load of speech+noise (hop = 4 strides of 40 samples, so the
//                                  two frames of a group share one batch of 20 loads per signal),
//                                  Hann, packed z = s + i*n, DFT-16 over n1, twiddle W_640^{n2 k1}
//                                                                          -> rows [k1][n2]
//   pass 2  lane = (frame, k1):   DFT-40 over n2 (5 x 8 PFA)              -> Z[k] natural order
//   post    lane = (frame, chunk): unpack S' = Z_k + conj Z_{N-k}, N' = (Z_k - conj Z_{N-k})/i,
//                                  M' = S' + f N', three magnitudes       -> in place
//   mel     lane = (frame, band):  banded Slaney filterbank sums           -> raw mel [sig][band][frame]
//   dB      lane = band:           20 log10(max(1e-5, .)), running max, 8-byte stores
//
// Reference semantics: /root/reference/data_processor.py:77-96 (signal_to_spectrogram),
// :130-133 (SNR mix), :35-57 (slice layout); librosa/mediaio semantics per SURVEY.md App. A.
//
// Every function is __host__ __device__: tests/emul builds the same code with g++ and runs a
// warp as a loop over 32 lanes per stage (the stage boundaries are the __syncwarp points).
#pragma once
#include "avse_common.h"
#include "avse_dft.cuh"

#if !defined(__CUDACC__)
#include <cmath>
#endif

// Tuning switches (measured on B200, profiles/README.md): the rolled pass 1 with twiddles loaded ahead
// of the DFT and the post scan unrolled by 7 gave the best kernel time (0.69 ms vs 0.78 ms per 1000 x 3 s).
#if !defined(AVSE_PASS1_UNROLLED) && !defined(AVSE_PASS1_ROLLED)
#define AVSE_PASS1_ROLLED 1
#endif
#if !defined(AVSE_NO_TW_PRELOAD) && !defined(AVSE_TW_PRELOAD)
#define AVSE_TW_PRELOAD 1
#endif
#if !defined(AVSE_POST_UNROLL)
#define AVSE_POST_UNROLL 7
#endif

namespace avse {

struct alignas(8) vec2 { float x, y; };
struct alignas(16) vec4 { float x, y, z, w; };

AVSE_HD float fast_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}

AVSE_HD float fast_log2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // x >= 1e-5: never denormal
    return y;
#else
    return log2f(x);
#endif
}

AVSE_HD float amp_to_db(float a) {
    // librosa.amplitude_to_db with ref=1, amin=1e-5 before the top_db floor (dp:94)
    constexpr float K = 6.0205999132796239f;  // 20 / log2(10)
    return K * fast_log2(a > AMIN ? a : AMIN);
}

// Constant tables.  window / tw1t / mel_w / mel_lo are staged in shared memory by the kernel.
struct FwdTables {
    const float* window;      // [640]
    const float* tw1t;        // [16][40][2]
    const float* mel_w;       // [80][MEL_WROW]
    const int* mel_lo;        // [80]
    const int* mel_roundw;    // [MEL_ROUNDS]
    const float* scan_w;      // [SCAN_BINS][2]  fused post+mel scan weights (avse_tables.h)
    const int* scan_loc;      // [80][4]
};

// Mel round widths of the reference configuration (sr 16 kHz, fmin 0, fmax 8 kHz): max band width
// over bands 16r..16r+15.  The specialised kernel unrolls the band loops with these constants; any
// other table uses the generic (runtime-width) kernel.
#define AVSE_STD_ROUNDW {3, 4, 7, 13, 23}

// One group of FPG frames of one utterance.
struct FwdTile {
    const float* sp;     // speech samples of this utterance
    const float* nz;     // noise samples (already fitted to the speech length), may be nullptr
    int L;               // signal length after pad/truncate (dp:37-42); reflect domain
    int valid_s;         // samples present in sp (zeros beyond, dp:40)
    int valid_n;         // samples present in nz
    int vmin;            // min(valid_s, valid_n) (0 when nz == nullptr): interior test
    int T;               // STFT frames: 1 + L / hop
    int t0;              // first frame of the group (multiple of FPG)
    float factor;        // SNR factor (dp:130); 0 when nz == nullptr
    float* mixed_pcm;    // [L] or nullptr: s + f*n (dp:133), zero-padded / truncated to L
};

AVSE_HD float load_sample_edge(const float* p, int i, int L, int valid) {
    // np.pad(y, n_fft//2, mode='reflect') on the length-L (zero padded) signal
    i = i < 0 ? -i : i;
    i = i >= L ? 2 * (L - 1) - i : i;
    i = i < 0 ? 0 : i;
    return (p != nullptr && i < valid) ? p[i] : 0.0f;
}

// A group is "interior" when both of its frames exist and all their samples are present in both
// signals without reflection or zero padding.
AVSE_HD bool group_interior(const FwdTile& tl) {
    return tl.nz != nullptr && tl.t0 * HOP - HALF >= 0 && (tl.t0 + 1) * HOP + HALF <= tl.vmin && tl.t0 + 1 < tl.T;
}

// ---------------------------------------------------------------------------------------
// pass 1.  80 (frame, n2) columns per group over 32 lanes in three rounds:
//   round 0: (frame 0, n2 = lane)   round 1: (frame 1, n2 = lane)   round 2: lanes 0..15 take
//   (frame lane/8, n2 = 32 + lane%8).
// ---------------------------------------------------------------------------------------
// DFT-16 over n1, twiddle W_640^{n2 k1}, store column n2 of frame f as rows [k1][n2].
AVSE_HD void pass1_column(float (&xr)[16], float (&xi)[16], int f, int n2, const vec2* s_tw, float* frames) {
#if defined(AVSE_TW_PRELOAD)
    // issue the 15 twiddle loads before the DFT so their latency hides behind its arithmetic
    vec2 twv[16];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) twv[k1] = s_tw[k1 * N2 + n2];
    dft16(xr, xi);
    float* rowp = frames + f * FRAME_F + 2 * n2;
    {
        vec2 v; v.x = xr[0]; v.y = xi[0];
        *reinterpret_cast<vec2*>(rowp) = v;
    }
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {
        vec2 v;
        v.x = xr[k1] * twv[k1].x - xi[k1] * twv[k1].y;
        v.y = xr[k1] * twv[k1].y + xi[k1] * twv[k1].x;
        *reinterpret_cast<vec2*>(rowp + k1 * ROW_F) = v;
    }
    return;
#endif
    dft16(xr, xi);
    float* row = frames + f * FRAME_F + 2 * n2;
    const vec2* tw = s_tw + n2;
    {
        vec2 v; v.x = xr[0]; v.y = xi[0];
        *reinterpret_cast<vec2*>(row) = v;
    }
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {
        const vec2 t = tw[k1 * N2];
        vec2 v;
        v.x = xr[k1] * t.x - xi[k1] * t.y;
        v.y = xr[k1] * t.y + xi[k1] * t.x;
        *reinterpret_cast<vec2*>(row + k1 * ROW_F) = v;
    }
}

// Interior fast path: no reflection, no bounds checks, rounds unrolled with static register indices.
// Rounds 0 and 1 share one batch of 20 strided loads per signal (frame 1's column is frame 0's
// shifted by 4 strides) and the 16 window values of residue n2 = lane.
AVSE_HD void stage_pass1_interior(const FwdTile& tl, int lane, const float* s_win, const vec2* s_tw, float* frames) {
    const int o0 = tl.t0 * HOP - HALF + lane;
    float xr[16], xi[16];
    {
        const float* ps = tl.sp + o0;
        const float* pn = tl.nz + o0;
        float rs[20], rn[20], wv[16];
#pragma unroll
        for (int j = 0; j < 20; ++j) { rs[j] = ps[N2 * j]; rn[j] = pn[N2 * j]; }
#pragma unroll
        for (int j = 0; j < 16; ++j) wv[j] = s_win[N2 * j + lane];
        if (tl.mixed_pcm != nullptr) {
            // own hops: frame 0 -> strides 8..11, frame 1 -> strides 12..15 of the shared batch
            float* pm = tl.mixed_pcm + o0 + HALF;
#pragma unroll
            for (int j = 0; j < 8; ++j) pm[N2 * j] = rs[8 + j] + tl.factor * rn[8 + j];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) { xr[j] = rs[j] * wv[j]; xi[j] = rn[j] * wv[j]; }
        pass1_column(xr, xi, 0, lane, s_tw, frames);
#pragma unroll
        for (int j = 0; j < 16; ++j) { xr[j] = rs[4 + j] * wv[j]; xi[j] = rn[4 + j] * wv[j]; }
        pass1_column(xr, xi, 1, lane, s_tw, frames);
    }
    if (lane < 16) {
        const int f = (lane >> 3) & 1, n2 = 32 + (lane & 7);
        const int o2 = (tl.t0 + f) * HOP - HALF + n2;
        const float* ps = tl.sp + o2;
        const float* pn = tl.nz + o2;
#pragma unroll
        for (int j = 0; j < 16; ++j) { xr[j] = ps[N2 * j]; xi[j] = pn[N2 * j]; }
        if (tl.mixed_pcm != nullptr) {
            float* pm = tl.mixed_pcm + o2 + HALF;
#pragma unroll
            for (int j = 0; j < 4; ++j) pm[N2 * j] = xr[8 + j] + tl.factor * xi[8 + j];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float w = s_win[N2 * j + n2]; xr[j] *= w; xi[j] *= w; }
        pass1_column(xr, xi, f, n2, s_tw, frames);
    }
}

// Edge / generic path (first and last frames, short or zero-padded signals, single-signal mode):
// every sample goes through the reflect + zero-pad loader.  Cold code, rolled.
AVSE_HD void stage_pass1_edge(const FwdTile& tl, int lane, const float* s_win, const vec2* s_tw, float* frames) {
#pragma unroll 1
    for (int round = 0; round < 3; ++round) {
        const int f = round < 2 ? round : (lane >> 3) & 1;
        const int n2 = round < 2 ? lane : 32 + (lane & 7);
        if (round == 2 && lane >= 16) continue;
        const int t_raw = tl.t0 + f;
        const int t = t_raw < tl.T ? t_raw : tl.T - 1;   // frames past the end duplicate the last one (never stored)
        const int base = t * HOP - HALF + n2;
        float xr[16], xi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            xr[j] = load_sample_edge(tl.sp, base + N2 * j, tl.L, tl.valid_s);
            xi[j] = load_sample_edge(tl.nz, base + N2 * j, tl.L, tl.valid_n);
        }
        if (tl.mixed_pcm != nullptr && t_raw < tl.T) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = t * HOP + N2 * j + n2;    // this frame's own hop: strides 8..11
                if (i < tl.L) tl.mixed_pcm[i] = xr[8 + j] + tl.factor * xi[8 + j];
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float w = s_win[N2 * j + n2]; xr[j] *= w; xi[j] *= w; }
        pass1_column(xr, xi, f, n2, s_tw, frames);
    }
}

// Interior path, rolled: one copy of the column code, every round loads its own 16 strides per
// signal (the 12 shared with the previous round come from L1/L2).  Smaller code and fewer live
// registers than the unrolled variant at the price of 24 more loads per group.
AVSE_HD void stage_pass1_interior_rolled(const FwdTile& tl, int lane, const float* s_win, const vec2* s_tw, float* frames) {
#pragma unroll 1
    for (int round = 0; round < 3; ++round) {
        if (round == 2 && lane >= 16) break;
        const int f = round < 2 ? round : (lane >> 3) & 1;
        const int n2 = round < 2 ? lane : 32 + (lane & 7);
        const int o = (tl.t0 + f) * HOP - HALF + n2;
        const float* ps = tl.sp + o;
        const float* pn = tl.nz + o;
        float xr[16], xi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { xr[j] = ps[N2 * j]; xi[j] = pn[N2 * j]; }
        if (tl.mixed_pcm != nullptr) {
            float* pm = tl.mixed_pcm + o + HALF;
#pragma unroll
            for (int j = 0; j < 4; ++j) pm[N2 * j] = xr[8 + j] + tl.factor * xi[8 + j];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float w = s_win[N2 * j + n2]; xr[j] *= w; xi[j] *= w; }
        pass1_column(xr, xi, f, n2, s_tw, frames);
    }
}

AVSE_HD void stage_pass1(const FwdTile& tl, int lane, const float* s_win, const vec2* s_tw, float* frames) {
#if defined(AVSE_PASS1_ROLLED)
    if (group_interior(tl)) stage_pass1_interior_rolled(tl, lane, s_win, s_tw, frames);
#else
    if (group_interior(tl)) stage_pass1_interior(tl, lane, s_win, s_tw, frames);
#endif
    else stage_pass1_edge(tl, lane, s_win, s_tw, frames);
}

// ---------------------------------------------------------------------------------------
// pass 2: lane = (f = lane/16, k1 = lane%16): load row, in-place DFT-40; after a warp sync the
// results are stored in natural order Z[k1 + 16 k2] over the same frame buffer.
// ---------------------------------------------------------------------------------------
AVSE_HD void pass2_compute(int lane, const float* frames, float (&xr)[40], float (&xi)[40]) {
    const int f = lane >> 4, k1 = lane & 15;
    const float* row = frames + f * FRAME_F + k1 * ROW_F;
#pragma unroll
    for (int q = 0; q < 20; ++q) {
        const vec4 v = *reinterpret_cast<const vec4*>(row + 4 * q);
        xr[2 * q] = v.x; xi[2 * q] = v.y; xr[2 * q + 1] = v.z; xi[2 * q + 1] = v.w;
    }
    dft40_inplace(xr, xi);
}

AVSE_HD void pass2_store(int lane, float* frames, const float (&xr)[40], const float (&xi)[40]) {
    const int f = lane >> 4, k1 = lane & 15;
    float* z = frames + f * FRAME_F + 2 * k1;
#pragma unroll
    for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int idx = (8 * c + 5 * d) % 40;     // register holding output (c, d)
            const int k2 = (16 * c + 25 * d) % 40;    // its frequency index within the DFT-40
            vec2 v; v.x = xr[idx]; v.y = xi[idx];
            *reinterpret_cast<vec2*>(z + 2 * N1 * k2) = v;   // Z[k1 + 16 k2]
        }
}

// ---------------------------------------------------------------------------------------
// post: lane = (f = lane/16, chunk p = lane%16), bins k = 21p .. 21p+20 (p = 15: 315..319 only).
// In place: slot k <- (|S'|, |N'|), slot 640-k <- (|M'|, 0).  Bin 0 is processed like any other
// (its "magnitudes" are garbage but finite and carry zero mel weight).
// STFT: also emit X_speech[k] (dp:79 D) to stft_row[0..320].
// ---------------------------------------------------------------------------------------
template <bool STFT>
AVSE_HD void stage_post(int lane, float factor, float* frames, vec2* stft_row) {
    const int f = lane >> 4, p = lane & 15;
    float* za = frames + f * FRAME_F + 2 * POST_CHUNK * p;            // slot k      = za + 2 i
    float* zc = frames + f * FRAME_F + 2 * (NFFT - POST_CHUNK * p);   // slot 640-k  = zc - 2 i
    const bool last = p == 15;
    constexpr int LAST_N = (NBINS - 1) - 15 * POST_CHUNK;              // 5 bins in the last chunk
    if (STFT && stft_row != nullptr && last) {
        vec2 d; d.x = za[2 * LAST_N]; d.y = 0.0f;                      // Nyquist bin 320 (real for real input)
        stft_row[NBINS - 1] = d;
    }
#pragma unroll
    for (int i = 0; i < POST_CHUNK; ++i) {
        if (i >= LAST_N && last) continue;
        const vec2 a = *reinterpret_cast<const vec2*>(za + 2 * i);
        const vec2 c = *reinterpret_cast<const vec2*>(zc - 2 * i);
        const float sr = a.x + c.x, si = a.y - c.y;     // 2 * X_speech[k]
        const float nr = a.y + c.y, ni = c.x - a.x;     // 2 * X_noise[k]
        const float mr = sr + factor * nr, mi = si + factor * ni;
        vec2 o1, o2;
        o1.x = fast_sqrt(sr * sr + si * si);
        o1.y = fast_sqrt(nr * nr + ni * ni);
        o2.x = fast_sqrt(mr * mr + mi * mi);
        o2.y = 0.0f;
        *reinterpret_cast<vec2*>(za + 2 * i) = o1;
        *reinterpret_cast<vec2*>(zc - 2 * i) = o2;
        if (STFT && stft_row != nullptr) {
            vec2 d;
            if (i == 0 && p == 0) { d.x = a.x; d.y = 0.0f; }    // DC bin (real for real input)
            else { d.x = 0.5f * sr; d.y = 0.5f * si; }
            stft_row[POST_CHUNK * p + i] = d;
        }
    }
}

// ---------------------------------------------------------------------------------------
// mel: round r, lane = (f = lane/16, band m = 16r + lane%16); W = number of bin iterations
// (compile-time for the specialised kernel).  Accumulates into as/an/am.
// ---------------------------------------------------------------------------------------
template <int W>
AVSE_HD void mel_round_fixed(int lane, int r, const float* s_melw, const int* s_mello, const float* frames,
                             float& as, float& an, float& am) {
    const int f = lane >> 4, m = 16 * r + (lane & 15);
    const int lo = s_mello[m];
    const float* wrow = s_melw + m * MEL_WROW;
    const float* zs = frames + f * FRAME_F + 2 * lo;
    const float* zm = frames + f * FRAME_F + 2 * (NFFT - lo);
    as = 0.0f; an = 0.0f; am = 0.0f;
    constexpr int W4 = (W + 3) / 4;
#pragma unroll
    for (int q = 0; q < W4; ++q) {
        const vec4 wv = *reinterpret_cast<const vec4*>(wrow + 4 * q);
        const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            if (j < W) {
                const vec2 sn = *reinterpret_cast<const vec2*>(zs + 2 * j);
                const float mm = zm[-2 * j];
                as += ww[e] * sn.x;
                an += ww[e] * sn.y;
                am += ww[e] * mm;
            }
        }
    }
}

AVSE_HD void mel_round_generic(int lane, int r, int roundw, const float* s_melw, const int* s_mello, const float* frames,
                               float& as, float& an, float& am) {
    const int f = lane >> 4, m = 16 * r + (lane & 15);
    const int lo = s_mello[m];
    const float* wrow = s_melw + m * MEL_WROW;
    const float* zs = frames + f * FRAME_F + 2 * lo;
    const float* zm = frames + f * FRAME_F + 2 * (NFFT - lo);
    as = 0.0f; an = 0.0f; am = 0.0f;
#pragma unroll 1
    for (int j = 0; j < roundw; ++j) {
        const float w = wrow[j];
        const vec2 sn = *reinterpret_cast<const vec2*>(zs + 2 * j);
        const float mm = zm[-2 * j];
        as += w * sn.x;
        an += w * sn.y;
        am += w * mm;
    }
}

// All five rounds; results stay in registers (acc[r][sig]) until every lane has finished reading
// the frame buffers, then stage_mel_store writes them over frame buffer 0.
template <bool STD>
AVSE_HD void stage_mel(int lane, const int* roundw, const float* s_melw, const int* s_mello, const float* frames,
                       float (&acc)[MEL_ROUNDS][3]) {
    if (STD) {
        mel_round_fixed<3>(lane, 0, s_melw, s_mello, frames, acc[0][0], acc[0][1], acc[0][2]);
        mel_round_fixed<4>(lane, 1, s_melw, s_mello, frames, acc[1][0], acc[1][1], acc[1][2]);
        mel_round_fixed<7>(lane, 2, s_melw, s_mello, frames, acc[2][0], acc[2][1], acc[2][2]);
        mel_round_fixed<13>(lane, 3, s_melw, s_mello, frames, acc[3][0], acc[3][1], acc[3][2]);
        mel_round_fixed<23>(lane, 4, s_melw, s_mello, frames, acc[4][0], acc[4][1], acc[4][2]);
    } else {
#pragma unroll
        for (int r = 0; r < MEL_ROUNDS; ++r)
            mel_round_generic(lane, r, roundw[r], s_melw, s_mello, frames, acc[r][0], acc[r][1], acc[r][2]);
    }
}

// raw mel layout in (dead) frame buffer 0: melst[(sig * 80 + band) * 2 + frame]
AVSE_HD void stage_mel_store(int lane, const float (&acc)[MEL_ROUNDS][3], float* melst) {
    const int f = lane >> 4;
#pragma unroll
    for (int r = 0; r < MEL_ROUNDS; ++r) {
        const int m = 16 * r + (lane & 15);
        melst[(0 * NMEL + m) * FPG + f] = acc[r][0];   // speech
        melst[(1 * NMEL + m) * FPG + f] = acc[r][1];   // noise (unscaled)
        melst[(2 * NMEL + m) * FPG + f] = acc[r][2];   // mixture
    }
}

// ---------------------------------------------------------------------------------------
// dB: lane = band m = 32 q + lane (q = 0..2), all three signals of both frames.
// Output layouts: slices [n_slices][80][20] (dp:49-57) or spectrogram [80][ld_t].
// Folds the max dB over the valid frames into mx[sig] (0 speech, 1 noise, 2 mixture).
// ---------------------------------------------------------------------------------------
struct FwdOut {
    float* dst[3];    // base of this utterance's output per signal (nullptr: skip stores)
    int layout;       // 0: slices [ns][80][20], 1: spectrogram [80][ld_t]
    int n_slices;     // slices kept (dp:164: min(video, audio))
    int ld_t;         // leading dimension for layout 1
};

AVSE_HD float neg_inf() {
#if defined(__CUDA_ARCH__)
    return __int_as_float(0xff800000);
#else
    return -INFINITY;
#endif
}

AVSE_HD void stage_db(int lane, int q, float factor, bool have_noise, const float* melst, const FwdOut& out, int t0, int T,
                      float (&mx)[3]) {
    const int m = 32 * q + lane;
    if (m >= NMEL) return;
    const bool v1 = t0 + 1 < T;    // frame t0 itself is always < T
    int off;                       // element offset of (band m, frame t0) inside this utterance's output
    bool store = true;
    if (out.layout == 0) {
        const int sl = t0 / SPSS, tt = t0 - sl * SPSS;   // 2 | t0 and 2 | 20: a group never straddles slices
        off = (sl * NMEL + m) * SPSS + tt;
        store = sl < out.n_slices;
    } else {
        off = m * out.ld_t + t0;
    }
#pragma unroll
    for (int sig = 0; sig < 3; ++sig) {
        if (sig > 0 && !have_noise) break;
        const vec2 v = *reinterpret_cast<const vec2*>(melst + (sig * NMEL + m) * FPG);
        const float scale = sig == 1 ? factor : 1.0f;
        const float d0 = amp_to_db(v.x * scale);
        const float d1 = amp_to_db(v.y * scale);
        const float lm = (v1 && d1 > d0) ? d1 : d0;
        mx[sig] = lm > mx[sig] ? lm : mx[sig];
        float* dst = out.dst[sig];
        if (dst != nullptr && store) {
            if (v1 && !(off & 1)) {
                vec2 o; o.x = d0; o.y = d1;
                *reinterpret_cast<vec2*>(dst + off) = o;    // off is even: 8-byte aligned
            } else {
                dst[off] = d0;
                if (v1) dst[off + 1] = d1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Fused post + mel "scan" (used when HostTables::scan_ok): lane = (chunk p = lane/2, frame f = lane%2)
// walks its 21 bins in order, unpacks the three magnitudes in registers and accumulates them straight
// into two running band sums: A = band seg-1 (falling edge), B = band seg (rising edge).  When the
// segment index advances (sign bit of the table's wA), A is complete for this lane: it is written over
// the bin's own, already consumed, slots and the accumulators rotate.  The two sums left at the end of
// the chunk go to the frame's flush area.  The magnitudes never touch shared memory, and the banded
// gather of the generic path (the dominant source of bank conflicts) disappears.
// s_scan: [SCAN_BINS] (wA | emit flag, wB).
// ---------------------------------------------------------------------------------------
template <bool STFT>
AVSE_HD void stage_post_scan(int lane, float factor, const vec2* s_scan, float* frames, vec2* stft_row) {
    const int f = lane & 1, p = lane >> 1;
    float* fr = frames + f * FRAME_F;
    float* za = fr + 2 * POST_CHUNK * p;            // slot k      = za + 2 i
    float* zc = fr + 2 * (NFFT - POST_CHUNK * p);   // slot 640-k  = zc - 2 i
    const vec2* tab = s_scan + POST_CHUNK * p;
    const bool last = p == 15;
    constexpr int LAST_N = (NBINS - 1) - 15 * POST_CHUNK;              // 5 bins in the last chunk
    if (STFT && stft_row != nullptr && last) {
        vec2 d; d.x = za[2 * LAST_N]; d.y = 0.0f;                      // Nyquist bin 320 (real for real input)
        stft_row[NBINS - 1] = d;
    }
    float As = 0.0f, An = 0.0f, Am = 0.0f, Bs = 0.0f, Bn = 0.0f, Bm = 0.0f;
#if defined(AVSE_POST_UNROLL)
#define AVSE_STR_(x) #x
#define AVSE_UNROLL_N_(n) _Pragma(AVSE_STR_(unroll n))
    AVSE_UNROLL_N_(AVSE_POST_UNROLL)
#else
#pragma unroll
#endif
    for (int i = 0; i < POST_CHUNK; ++i) {
        if (i >= LAST_N && last) continue;
        const vec2 a = *reinterpret_cast<const vec2*>(za + 2 * i);
        const vec2 c = *reinterpret_cast<const vec2*>(zc - 2 * i);
        const vec2 w = tab[i];
#if defined(__CUDA_ARCH__)
        const bool emit = __float_as_int(w.x) < 0;
#else
        const bool emit = std::signbit(w.x);
#endif
        if (emit) {
            vec2 o1, o2;
            o1.x = As; o1.y = An; o2.x = Am; o2.y = 0.0f;
            *reinterpret_cast<vec2*>(za + 2 * i) = o1;
            *reinterpret_cast<vec2*>(zc - 2 * i) = o2;
        }
        const float sr = a.x + c.x, si = a.y - c.y;     // 2 * X_speech[k]
        const float nr = a.y + c.y, ni = c.x - a.x;     // 2 * X_noise[k]
        const float mr = sr + factor * nr, mi = si + factor * ni;
        const float ms = fast_sqrt(sr * sr + si * si);
        const float mn = fast_sqrt(nr * nr + ni * ni);
        const float mm = fast_sqrt(mr * mr + mi * mi);
        const float wa = fabsf(w.x), wb = w.y;
        const float a_s = emit ? Bs : As, a_n = emit ? Bn : An, a_m = emit ? Bm : Am;
        const float b_s = emit ? 0.0f : Bs, b_n = emit ? 0.0f : Bn, b_m = emit ? 0.0f : Bm;
        As = a_s + wa * ms; An = a_n + wa * mn; Am = a_m + wa * mm;
        Bs = b_s + wb * ms; Bn = b_n + wb * mn; Bm = b_m + wb * mm;
        if (STFT && stft_row != nullptr) {
            vec2 d;
            if (i == 0 && p == 0) { d.x = a.x; d.y = 0.0f; }    // DC bin (real for real input)
            else { d.x = 0.5f * sr; d.y = 0.5f * si; }
            stft_row[POST_CHUNK * p + i] = d;
        }
    }
    float* fl = fr + FRAME_FLUSH_F + 6 * p;
    { vec2 o; o.x = As; o.y = An; *reinterpret_cast<vec2*>(fl + 0) = o; }
    { vec2 o; o.x = Bs; o.y = Bn; *reinterpret_cast<vec2*>(fl + 2) = o; }
    { vec2 o; o.x = Am; o.y = Bm; *reinterpret_cast<vec2*>(fl + 4) = o; }
}

struct alignas(16) ivec4 { int x, y, z, w; };

// dB stage of the scan path: lane = band m = 32 q + lane; each band's mel sum is the sum of its two
// partial-sum locations (s_loc[m] = frame-relative offsets SN0, M0, SN1, M1; absent parts point at zeros).
AVSE_HD void stage_db_scan(int lane, int q, float factor, bool have_noise, const ivec4* s_loc, const float* frames,
                           const FwdOut& out, int t0, int T, float (&mx)[3]) {
    const int m = 32 * q + lane;
    if (m >= NMEL) return;
    const ivec4 loc = s_loc[m];
    float mel[3][2];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
        const float* fr = frames + f * FRAME_F;
        const vec2 sn0 = *reinterpret_cast<const vec2*>(fr + loc.x);
        const vec2 sn1 = *reinterpret_cast<const vec2*>(fr + loc.z);
        mel[0][f] = sn0.x + sn1.x;
        mel[1][f] = (sn0.y + sn1.y) * factor;
        mel[2][f] = fr[loc.y] + fr[loc.w];
    }
    const bool v1 = t0 + 1 < T;    // frame t0 itself is always < T
    int off;
    bool store = true;
    if (out.layout == 0) {
        const int sl = t0 / SPSS, tt = t0 - sl * SPSS;   // 2 | t0 and 2 | 20: a group never straddles slices
        off = (sl * NMEL + m) * SPSS + tt;
        store = sl < out.n_slices;
    } else {
        off = m * out.ld_t + t0;
    }
#pragma unroll
    for (int sig = 0; sig < 3; ++sig) {
        if (sig > 0 && !have_noise) break;
        const float d0 = amp_to_db(mel[sig][0]);
        const float d1 = amp_to_db(mel[sig][1]);
        const float lm = (v1 && d1 > d0) ? d1 : d0;
        mx[sig] = lm > mx[sig] ? lm : mx[sig];
        float* dst = out.dst[sig];
        if (dst != nullptr && store) {
            if (v1 && !(off & 1)) {
                vec2 o; o.x = d0; o.y = d1;
                *reinterpret_cast<vec2*>(dst + off) = o;
            } else {
                dst[off] = d0;
                if (v1) dst[off + 1] = d1;
            }
        }
    }
}

// order-preserving float <-> int key for atomicMax on floats of either sign
AVSE_HD int float_to_key(float x) {
#if defined(__CUDA_ARCH__)
    const int b = __float_as_int(x);
#else
    int b; { union { float f; int i; } u; u.f = x; b = u.i; }
#endif
    return b >= 0 ? b : (b ^ 0x7fffffff);
}
AVSE_HD float key_to_float(int k) {
    const int b = k >= 0 ? k : (k ^ 0x7fffffff);
#if defined(__CUDA_ARCH__)
    return __int_as_float(b);
#else
    union { float f; int i; } u; u.i = b; return u.f;
#endif
}

}  // namespace avse
