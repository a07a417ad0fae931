// Warp-stage functions of the fused forward kernel (mix at SNR -> STFT -> mel -> dB).
//
// One warp owns a group of FPG = 4 consecutive STFT frames of one utterance and walks them
// through five warp-synchronous stages that exchange data only through that warp's private
// shared-memory region (no block-level barrier anywhere in the frame loop):
//
//   pass 1  lane = (frame, n2):   strided load of speech+noise, Hann, packed z = s + i*n,
//                                  DFT-16 over n1, twiddle W_640^{n2 k1}  -> rows [k1][n2]
//   pass 2  lane = (frame, k1):   DFT-40 over n2 (5 x 8 PFA)              -> Z[k] natural order
//   post    lane = (frame, chunk): unpack S' = Z_k + conj Z_{N-k}, N' = (Z_k - conj Z_{N-k})/i,
//                                  M' = S' + f N', three magnitudes       -> in place
//   mel     lane = (frame, band):  banded Slaney filterbank sums           -> raw mel [sig][band][frame]
//   dB      lane = band:           20 log10(max(1e-5, .)), running max, 16-byte stores
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
    return __log2f(x);
#else
    return log2f(x);
#endif
}

AVSE_HD float amp_to_db(float a) {
    // librosa.amplitude_to_db with ref=1, amin=1e-5 before the top_db floor (dp:94)
    constexpr float K = 6.0205999132796239f;  // 20 / log2(10)
    return K * fast_log2(a > AMIN ? a : AMIN);
}

// Constant tables, device-resident (or host arrays in the emulation).
struct FwdTables {
    const float* window;      // [640]
    const float* tw1t;        // [16][40][2]
    const float* mel_w;       // [80][MEL_WROW]   (copied to shared memory by the kernel)
    const int* mel_lo;        // [80]
    const int* mel_roundw;    // [10]
};

// One group of 4 frames of one utterance.
struct FwdTile {
    const float* sp;     // speech samples of this utterance
    const float* nz;     // noise samples (already fitted to the speech length), may be nullptr
    int L;               // signal length after pad/truncate (dp:37-42); reflect domain
    int valid_s;         // samples present in sp (zeros beyond, dp:40)
    int valid_n;         // samples present in nz
    int T;               // STFT frames: 1 + L / hop
    int t0;              // first frame of the group (multiple of 4)
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

// ---------------------------------------------------------------------------------------
// pass 1: one (frame f, residue n2) task
// ---------------------------------------------------------------------------------------
template <bool EDGE>
AVSE_HD void pass1_task(const FwdTile& tl, int f, int n2, const float (&w)[16], const float (&twr)[16],
                        const float (&twi)[16], float* frames) {
    const int t_raw = tl.t0 + f;
    const int t = t_raw < tl.T ? t_raw : tl.T - 1;
    const int base = t * HOP - HALF + n2;
    float xr[16], xi[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const int i = base + N2 * n1;
        float s, n;
        if (EDGE) {
            s = load_sample_edge(tl.sp, i, tl.L, tl.valid_s);
            n = load_sample_edge(tl.nz, i, tl.L, tl.valid_n);
        } else {
            s = tl.sp[i];
            n = tl.nz[i];
        }
        if (n1 >= 8 && n1 < 12) {
            // this frame's own hop: original samples [160 t, 160 t + 160)
            if (tl.mixed_pcm != nullptr && (!EDGE || (t_raw < tl.T && i < tl.L))) tl.mixed_pcm[i] = s + tl.factor * n;
        }
        xr[n1] = s * w[n1];
        xi[n1] = n * w[n1];
    }
    dft16(xr, xi);
    float* row = frames + f * FRAME_F + 2 * n2;
    {
        vec2 v; v.x = xr[0]; v.y = xi[0];
        *reinterpret_cast<vec2*>(row) = v;
    }
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {
        vec2 v;
        v.x = xr[k1] * twr[k1] - xi[k1] * twi[k1];
        v.y = xr[k1] * twi[k1] + xi[k1] * twr[k1];
        *reinterpret_cast<vec2*>(row + k1 * ROW_F) = v;
    }
}

AVSE_HD void load_lane_consts(const FwdTables& tb, int n2, float (&w)[16], float (&twr)[16], float (&twi)[16]) {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) w[n1] = tb.window[N2 * n1 + n2];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const vec2 v = *reinterpret_cast<const vec2*>(tb.tw1t + (k1 * N2 + n2) * 2);
        twr[k1] = v.x;
        twi[k1] = v.y;
    }
}

template <bool EDGE>
AVSE_HD void stage_pass1(const FwdTables& tb, const FwdTile& tl, int lane, float* frames) {
    float w[16], twr[16], twi[16];
    load_lane_consts(tb, lane, w, twr, twi);
#pragma unroll 1
    for (int f = 0; f < FPG; ++f) pass1_task<EDGE>(tl, f, lane, w, twr, twi, frames);
    const int n2b = 32 + (lane & 7);
    load_lane_consts(tb, n2b, w, twr, twi);
    pass1_task<EDGE>(tl, lane >> 3, n2b, w, twr, twi, frames);
}

// ---------------------------------------------------------------------------------------
// pass 2: lane = (f = 2j + lane/16, k1 = lane%16); load + DFT-40, then (after a warp sync) store
// ---------------------------------------------------------------------------------------
AVSE_HD void pass2_compute(int lane, int j, const float* frames, float (&yr)[40], float (&yi)[40]) {
    const int f = 2 * j + (lane >> 4), k1 = lane & 15;
    const float* row = frames + f * FRAME_F + k1 * ROW_F;
    float xr[40], xi[40];
#pragma unroll
    for (int q = 0; q < 20; ++q) {
        const vec4 v = *reinterpret_cast<const vec4*>(row + 4 * q);
        xr[2 * q] = v.x; xi[2 * q] = v.y; xr[2 * q + 1] = v.z; xi[2 * q + 1] = v.w;
    }
    dft40(xr, xi, yr, yi);
}

AVSE_HD void pass2_store(int lane, int j, float* frames, const float (&yr)[40], const float (&yi)[40]) {
    const int f = 2 * j + (lane >> 4), k1 = lane & 15;
    float* z = frames + f * FRAME_F + 2 * k1;
#pragma unroll
    for (int k2 = 0; k2 < 40; ++k2) {
        vec2 v; v.x = yr[k2]; v.y = yi[k2];
        *reinterpret_cast<vec2*>(z + 2 * N1 * k2) = v;   // Z[k1 + 16 k2]
    }
}

// ---------------------------------------------------------------------------------------
// post: lane = (f = lane/8, chunk p = lane%8), bins k = 41p .. 41p+40 clipped to [1, 319]
// in place: slot k <- (|S'|, |N'|), slot 640-k <- (|M'|, 0)
// ---------------------------------------------------------------------------------------
AVSE_HD void stage_post(int lane, float factor, float* frames, vec2* stft_row) {
    // stft_row: optional [321] complex row of this lane's frame receiving X_speech (dp:79 D), or nullptr
    const int f = lane >> 3, p = lane & 7;
    float* zb = frames + f * FRAME_F;
#pragma unroll 4
    for (int i = 0; i < POST_CHUNK; ++i) {
        const int k = POST_CHUNK * p + i;
        if (k >= 1 && k <= NBINS - 2) {
            const vec2 a = *reinterpret_cast<const vec2*>(zb + 2 * k);
            const vec2 c = *reinterpret_cast<const vec2*>(zb + 2 * (NFFT - k));
            const float sr = a.x + c.x, si = a.y - c.y;     // 2 * X_speech[k]
            const float nr = a.y + c.y, ni = c.x - a.x;     // 2 * X_noise[k]
            const float mr = sr + factor * nr, mi = si + factor * ni;
            vec2 o1, o2;
            o1.x = fast_sqrt(sr * sr + si * si);
            o1.y = fast_sqrt(nr * nr + ni * ni);
            o2.x = fast_sqrt(mr * mr + mi * mi);
            o2.y = 0.0f;
            *reinterpret_cast<vec2*>(zb + 2 * k) = o1;
            *reinterpret_cast<vec2*>(zb + 2 * (NFFT - k)) = o2;
            if (stft_row != nullptr) { vec2 d; d.x = 0.5f * sr; d.y = 0.5f * si; stft_row[k] = d; }
        } else if (stft_row != nullptr && (k == 0 || k == NBINS - 1)) {
            vec2 d; d.x = zb[2 * k]; d.y = 0.0f;   // DC / Nyquist of the real part of the packed input
            stft_row[k] = d;
        }
    }
}

// ---------------------------------------------------------------------------------------
// mel: round r, lane = (f = lane/8, band m = 8r + lane%8)
// mel_w / mel_lo may live in shared memory
// ---------------------------------------------------------------------------------------
AVSE_HD void stage_mel_round(int lane, int r, int roundw, const float* mel_w, const int* mel_lo,
                             const float* frames, float* melst) {
    const int f = lane >> 3, m = 8 * r + (lane & 7);
    const int lo = mel_lo[m];
    const float* wrow = mel_w + m * MEL_WROW;
    const float* zs = frames + f * FRAME_F + 2 * lo;
    const float* zm = frames + f * FRAME_F + 2 * (NFFT - lo);
    float as = 0.0f, an = 0.0f, am = 0.0f;
#pragma unroll 1
    for (int j = 0; j < roundw; ++j) {
        const float w = wrow[j];
        const vec2 sn = *reinterpret_cast<const vec2*>(zs + 2 * j);
        const float mm = zm[-2 * j];
        as += w * sn.x;
        an += w * sn.y;
        am += w * mm;
    }
    melst[(0 * NMEL + m) * FPG + f] = as;   // speech
    melst[(1 * NMEL + m) * FPG + f] = an;   // noise (unscaled)
    melst[(2 * NMEL + m) * FPG + f] = am;   // mixture
}

// ---------------------------------------------------------------------------------------
// dB: signal sig (0 speech, 1 noise, 2 mixture), sub-round q in 0..2, band m = 32q + lane
// Output layouts: slices [n_slices][80][20] (dp:49-57) or spectrogram [80][ld_t].
// Returns the max dB over the valid frames of this task (or -inf).
// ---------------------------------------------------------------------------------------
struct FwdOut {
    float* dst;       // base of this utterance's output for signal sig (nullptr: skip stores)
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

AVSE_HD float stage_db(int lane, int q, float scale, const float* melst_sig, const FwdOut& out, int t0, int T) {
    const int m = 32 * q + lane;
    float mx = neg_inf();
    if (m < NMEL) {
        const vec4 v = *reinterpret_cast<const vec4*>(melst_sig + m * FPG);
        float d[4];
        d[0] = amp_to_db(v.x * scale);
        d[1] = amp_to_db(v.y * scale);
        d[2] = amp_to_db(v.z * scale);
        d[3] = amp_to_db(v.w * scale);
#pragma unroll
        for (int f = 0; f < 4; ++f)
            if (t0 + f < T) mx = d[f] > mx ? d[f] : mx;
        if (out.dst != nullptr) {
            if (out.layout == 0) {
                const int spss = 20;
                const int sl = t0 / spss, tt = t0 - sl * spss;    // 4 | t0 and 4 | 20: group never straddles
                if (sl < out.n_slices) {
                    float* p = out.dst + ((size_t)sl * NMEL + m) * spss + tt;
                    if (t0 + 3 < T) {
                        vec4 o; o.x = d[0]; o.y = d[1]; o.z = d[2]; o.w = d[3];
                        *reinterpret_cast<vec4*>(p) = o;
                    } else {
#pragma unroll
                        for (int f = 0; f < 4; ++f)
                            if (t0 + f < T) p[f] = d[f];
                    }
                }
            } else {
                float* p = out.dst + (size_t)m * out.ld_t + t0;
#pragma unroll
                for (int f = 0; f < 4; ++f)
                    if (t0 + f < T) p[f] = d[f];
            }
        }
    }
    return mx;
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
