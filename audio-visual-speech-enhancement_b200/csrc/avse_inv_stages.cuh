// Warp-stage functions of the inverse kernel: dB log-mel + mixture PCM -> enhanced PCM
// (reconstruct_speech_signal, /root/reference/data_processor.py:60-74 and :99-116).
//
// Per frame the reference computes  lin = pinv(F) 10^(dB/20),  Y = lin * phase(STFT(mixture)),
// y = istft(Y)  (irfft, Hann, overlap-add, / window sum-square, centre trim).  Here:
//   * pinv(F) = F^T (F F^T)^-1 with F F^T tridiagonal: a Thomas solve per frame (separate tiny kernel,
//     avse_mel_to_coef_kernel) gives c = (F F^T)^-1 a; lin[k] = sum of <= 2 taps of F^T c.
//   * Two real frames ride one complex 640-point FFT in both directions: the mixture frames (t, t+1)
//     are packed as re/im for the phase, and V = conj(Y_t + i Y_{t+1}) (Hermitian-extended) is sent
//     through the same forward-FFT codelets, giving 640 (y_t - i y_{t+1}).
//   * One warp streams through consecutive groups of 4 frames of one utterance chunk and keeps the
//     overlap-add in registers: lane l owns output samples == l (mod 40) (the 8 residues 32..39 live
//     in a small shared side buffer), so no atomics and no inter-warp exchange are needed.
//
// Stage order per group (warp-synchronous, private shared memory):
//   pass 1 (mixture -> rows) | pass 2 (-> Z) | post (phase * lin -> V, in place) |
//   pass A (DFT-40 over n2', twiddle -> rows) | pass B (DFT-16 over n1', window, overlap-add) | emit 4 hops
#pragma once
#include "avse_common.h"
#include "avse_dft.cuh"
#include "avse_fwd_stages.cuh"

#if !defined(AVSE_INV_POST_UNROLL)
#define AVSE_INV_POST_UNROLL 3      // unroll factor of the post stage's bin loop (21 bins): code size vs loads in flight
#endif

namespace avse {

constexpr int INV_FPG = 4;                 // real frames per group (2 packed complex FFTs)
constexpr int INV_Y_F = NMEL * INV_FPG;    // coefficient buffer [80][4]
constexpr int INV_SIDE_ROWS = 28;          // overlap-add rows (40 samples each) alive per group
constexpr int INV_SIDE_F = INV_SIDE_ROWS * 8;
constexpr int INV_WARP_SMEM_F = 2 * FRAME_F + INV_Y_F + INV_SIDE_F;   // 2912 + 320 + 224 = 3456 floats
constexpr float INV_SCALE = 1.0f / NFFT;   // irfft normalisation

struct InvTile {
    const float* pcm;   // mixture samples of this utterance
    int L;              // signal length (reflect domain of librosa.stft)
    int valid;          // samples present (zeros beyond)
    int T;              // STFT frames of the mixture: 1 + L / hop
    int T_use;          // frames reconstructed: min(20 * n_slices, T)  (dp:68)
    int t0;             // first frame of the group (multiple of 4)
};

AVSE_HD bool inv_group_interior(const InvTile& tl) {
    return tl.t0 * HOP - HALF >= 0 && (tl.t0 + 3) * HOP + HALF <= tl.valid && tl.t0 + 3 < tl.T;
}

// An all-zero windowed frame (zero padding dp:40, digital silence) has D == 0 exactly in the reference and
// therefore phase 1 + 0j (librosa.magphase).  Packed with a non-zero partner frame its unpacked spectrum
// would be rounding noise with a random phase, so pass 1 records per packed frame whether any windowed
// sample is non-zero (all writers store the same value: no race) and the post stage forces 1 + 0j otherwise.
// Flags live in the frame region's otherwise unused floats [FRAME_ZERO_F, FRAME_ZERO_F + 2).
AVSE_HD void inv_mark_nonzero(const cpx (&x)[16], float* frame_base) {
    bool nr = false, ni = false;
#pragma unroll
    for (int j = 0; j < 16; ++j) { nr = nr || (cre(x[j]) != 0.0f); ni = ni || (cim(x[j]) != 0.0f); }
    if (nr) frame_base[FRAME_ZERO_F] = 1.0f;
    if (ni) frame_base[FRAME_ZERO_F + 1] = 1.0f;
}

AVSE_HD int float_bits(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(x);
#else
    union { float f; int i; } u; u.f = x; return u.i;
#endif
}

// Interior groups: the same flags from the RAW samples (3-input integer ORs of the float bit patterns, ~4x fewer
// instructions than comparing the 32 windowed values).  w[n] != 0 for every n except n = 0 (periodic Hann), so the
// windowed frame is all-zero iff every raw sample except the one at n = 0 (j = 0 of column n2 = 0) is zero.
AVSE_HD void inv_mark_nonzero_raw(const float (&raw)[20], int n2, float* frame_base) {
    int mid = 0;
#pragma unroll
    for (int j = 5; j < 16; ++j) mid |= float_bits(raw[j]);
    const int r0 = n2 != 0 ? float_bits(raw[0]) : 0;     // frame A's n = 0 sample meets w[0] = 0
    const int r4 = float_bits(raw[4]);                   // frame A: n = 160 + n2; frame B: n = n2 (its n = 0 sample in column 0)
    const int a = mid | r0 | float_bits(raw[1]) | float_bits(raw[2]) | float_bits(raw[3]) | r4;
    const int b = mid | (n2 != 0 ? r4 : 0) | float_bits(raw[16]) | float_bits(raw[17]) | float_bits(raw[18]) | float_bits(raw[19]);
    if ((a & 0x7fffffff) != 0) frame_base[FRAME_ZERO_F] = 1.0f;
    if ((b & 0x7fffffff) != 0) frame_base[FRAME_ZERO_F + 1] = 1.0f;
}

// ---------------------------------------------------------------------------------------
// pass 1: complex FFT f packs mixture frames tA = t0 + 2f (real part) and tA + 1 (imaginary part);
// the second is the first shifted by 4 strides of 40 samples, so one batch of 20 loads serves both.
// s_win2: [640] (w, w) pairs.
// ---------------------------------------------------------------------------------------
AVSE_HD void inv_stage_pass1(const InvTile& tl, int lane, const float* s_win2, const vec2* s_tw, float* frames) {
    const bool interior = inv_group_interior(tl);
#pragma unroll 1
    for (int round = 0; round < 3; ++round) {
        if (round == 2 && lane >= 16) break;
        const int f = round < 2 ? round : (lane >> 3) & 1;
        const int n2 = round < 2 ? lane : 32 + (lane & 7);
        const int tA = tl.t0 + 2 * f;
        cpx x[16];
        if (interior) {
            const float* p = tl.pcm + tA * HOP - HALF + n2;
            float raw[20];
#pragma unroll
            for (int j = 0; j < 20; ++j) raw[j] = p[N2 * j];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = cmake(raw[j], raw[j + 4]);
            inv_mark_nonzero_raw(raw, n2, frames + f * FRAME_F);
        } else {
            // frames beyond the last one are clamped (their coefficients are zero, so they contribute nothing);
            // reflection breaks the shift relation: load both frames explicitly
            const int ta = tA < tl.T ? tA : tl.T - 1;
            const int tb = tA + 1 < tl.T ? tA + 1 : tl.T - 1;
            const int ba = ta * HOP - HALF + n2, bb = tb * HOP - HALF + n2;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                x[j] = cmake(load_sample_edge(tl.pcm, ba + N2 * j, tl.L, tl.valid), load_sample_edge(tl.pcm, bb + N2 * j, tl.L, tl.valid));
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = cmul_pp(x[j], cload(s_win2 + 2 * (N2 * j + n2)));
        if (!interior) inv_mark_nonzero(x, frames + f * FRAME_F);
        pass1_column(x, f, n2, s_tw, frames);
    }
}

// ---------------------------------------------------------------------------------------
// post: lane = (chunk p = lane/2, complex FFT f = lane%2).  For each bin k of the chunk:
//   X_A = (Z_k + conj Z_{N-k})/2, X_B = (Z_k - conj Z_{N-k})/(2i)        (frames tA, tA+1)
//   phase = X/|X| (1 + 0j where X == 0, librosa.magphase dp:80)
//   lin_t[k] = w0 c_t[b0] + w1 c_t[b1]                                    (F^T c, dp:112)
//   Y = lin * phase;  V_k = conj(Y_A + i Y_B), V_{N-k} = conj(conj(Y_A) + i conj(Y_B))   (in place)
// s_col: [SCAN_BINS] (b0, b1, w0, w1) as ivec4 bit patterns; ybuf: [80][4] coefficients of the group.
// EXT: the phase is read from a caller-supplied array (reconstruct_signal_from_spectrogram, dp:99) instead of
// the recomputed mixture STFT; phA / phB point at the [321] complex rows of frames tA and tA + 1 (or nullptr).
// ---------------------------------------------------------------------------------------
AVSE_HD float inv_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}

// 2^x for the dB -> amplitude conversion (x = dB log2(10) / 20, a few tens in magnitude; the -1e30 dB sentinel of absent frames
// gives exactly 0).  exp2f() wraps the same MUFU.EX2 in a range test and two scalings for results below 2^-126 that cannot occur.
AVSE_HD float inv_exp2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}

AVSE_HD float bits_to_float(int b) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(b);
#else
    union { float f; int i; } u; u.i = b; return u.f;
#endif
}

template <bool EXT>
AVSE_HD void inv_stage_post(int lane, const ivec4* s_col, const float* ybuf, float* frames, const vec2* phA, const vec2* phB) {
    const int f = lane & 1, p = lane >> 1;
    float* fr = frames + f * FRAME_F;
    float* za = fr + 2 * POST_CHUNK * p;
    float* zc = fr + 2 * (NFFT - POST_CHUNK * p);
    const ivec4* tab = s_col + POST_CHUNK * p;
    const float* yb = ybuf + 2 * f;     // (c_tA, c_tA+1) of band b at yb[4 b .. 4 b + 1]
    const bool last = p == 15;
    const bool liveA = EXT || fr[FRAME_ZERO_F] != 0.0f;       // see inv_mark_nonzero
    const bool liveB = EXT || fr[FRAME_ZERO_F + 1] != 0.0f;
    constexpr int LAST_N = (NBINS - 1) - 15 * POST_CHUNK;
    AVSE_UNROLL_N_(AVSE_INV_POST_UNROLL)
    for (int i = 0; i < POST_CHUNK; ++i) {
        // chunk 15 holds only bins 315..320: the tail iterations compute on harmless in-range slots and keep their stores
        // predicated off (a branch here diverges in every iteration of the hot loop); the Nyquist bin (lin = 0) stores 0
        const bool tail = last && i >= LAST_N;
        const bool nyq = last && i == LAST_N;
        const ivec4 t = tab[i];
        const cpx y0 = cload(yb + 4 * t.x), y1 = cload(yb + 4 * t.y);
        // (lin_A, lin_B) = w0 (c_A, c_B)[b0] + w1 (c_A, c_B)[b1]
        const cpx lin = cfma_s(y1, bits_to_float(t.w), cmul_s(y0, bits_to_float(t.z)));
        float yar, yai, ybr, ybi;      // Y_A = lin_A * phase_A, Y_B = lin_B * phase_B
        if (EXT) {
            const int k = POST_CHUNK * p + i;
            const vec2 qa = (phA != nullptr && !tail) ? phA[k] : vec2{1.0f, 0.0f};
            const vec2 qb = (phB != nullptr && !tail) ? phB[k] : vec2{1.0f, 0.0f};
            yar = cre(lin) * qa.x; yai = cre(lin) * qa.y;
            ybr = cim(lin) * qb.x; ybi = cim(lin) * qb.y;
        } else {
            const cpx a = cload(za + 2 * i);
            const cpx c = cload(zc - 2 * i);
            const cpx xa = cfma_pp(c, cmake(1.0f, -1.0f), a);             // 2 X_A = Z_k + conj Z_{N-k}
            const cpx xn = cfma_pp(c, cmake(-1.0f, 1.0f), a);             // 2 i X_B = Z_k - conj Z_{N-k}:  2 X_B = (xn.im, -xn.re)
            const float na = cre(xa) * cre(xa) + cim(xa) * cim(xa), nb = cre(xn) * cre(xn) + cim(xn) * cim(xn);
            // lin * X / |X| with the scale lin / |X| formed once per frame; 1 + 0j where X == 0 (librosa.magphase, dp:80)
            const float sa = cre(lin) * inv_rsqrt(na), sb = cim(lin) * inv_rsqrt(nb);
            const bool okA = na > 0.0f && liveA, okB = nb > 0.0f && liveB;
            yar = okA ? cre(xa) * sa : cre(lin); yai = okA ? cim(xa) * sa : 0.0f;
            ybr = okB ? cim(xn) * sb : cim(lin); ybi = okB ? -cre(xn) * sb : 0.0f;
        }
        if (nyq) cstore(za + 2 * i, cmake(0.0f, 0.0f));
        if (!tail) {
            cstore(za + 2 * i, cmake(yar - ybi, -(yai + ybr)));   // conj(Y_A + i Y_B)
            cstore(zc - 2 * i, cmake(yar + ybi, yai - ybr));      // conj(conj(Y_A) + i conj(Y_B))
        }
    }
}

// ---------------------------------------------------------------------------------------
// pass A: lane = (f = lane/16, n1' = lane%16): gather V[n1' + 16 n2'], DFT-40 over n2', multiply by
// W_640^{n1' k2'}; after a warp sync the 40 results are written as row n1' = [k2'].
// s_twT: [40][16] vec2, W_640^{n1' k2'} with n1' minor (conflict-free for lane = n1').
// ---------------------------------------------------------------------------------------
// The three pieces of pass A (gather, DFT-40, twiddle) are separate functions so that the kernel can run pass 2 and pass A
// through ONE copy of the DFT-40 codelet (a rolled two-phase loop): the kernel's hot loop is larger than the 32 KB L1.5
// instruction cache and "no instruction" was its top stall reason (profiles/inverse_kernels_r1.txt).
AVSE_HD void inv_passA_load(int lane, const float* frames, cpx (&x)[40]) {
    const int f = lane >> 4, n1 = lane & 15;
    const float* z = frames + f * FRAME_F + 2 * n1;
#pragma unroll
    for (int n2 = 0; n2 < 40; ++n2) x[n2] = cload(z + 2 * N1 * n2);
}

AVSE_HD void inv_passA_twiddle(int lane, const vec2* s_twT, cpx (&x)[40]) {
    const int n1 = lane & 15;
#pragma unroll
    for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int idx = (8 * c + 5 * d) % 40, k2 = (16 * c + 25 * d) % 40;
            if (k2 == 0) continue;
            const vec2 t = s_twT[k2 * N1 + n1];
            x[idx] = cmul(x[idx], t.x, t.y);
        }
}

AVSE_HD void inv_passA_compute(int lane, const vec2* s_twT, const float* frames, cpx (&x)[40]) {
    inv_passA_load(lane, frames, x);
    dft40_inplace(x);
    inv_passA_twiddle(lane, s_twT, x);
}

AVSE_HD void inv_passA_store(int lane, float* frames, const cpx (&x)[40]) {
    const int f = lane >> 4, n1 = lane & 15;
    float* row = frames + f * FRAME_F + n1 * ROW_F;
    if (n1 == 0) { frames[f * FRAME_F + FRAME_ZERO_F] = 0.0f; frames[f * FRAME_F + FRAME_ZERO_F + 1] = 0.0f; }   // re-arm inv_mark_nonzero
#pragma unroll
    for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int idx = (8 * c + 5 * d) % 40, k2 = (16 * c + 25 * d) % 40;
            cstore(row + 2 * k2, x[idx]);
        }
}

// ---------------------------------------------------------------------------------------
// pass B: column k2' of complex FFT f: DFT-16 over n1' gives 640 (y_A - i y_B) at n = 40 k1' + k2'.
// Window (x 1/640) and overlap-add.  Rounds 0/1 (k2' = lane): acc[J], J = j + 8 f, row J = samples
// 160 t0 + 40 J + lane.  Round 2 (k2' = 32 + lane%8, f = lane/8, lanes 0..15): shared side buffer
// side[J][r], two ordered phases (f = 0 then f = 1) so the adds never race.
// ---------------------------------------------------------------------------------------
AVSE_HD void inv_passB_column(int f, int k2, const float* s_win2, const float* frames, float (&c)[20]) {
    const float* col = frames + f * FRAME_F + 2 * k2;
    cpx x[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) x[n1] = cload(col + n1 * ROW_F);
    dft16(x);
#pragma unroll
    for (int j = 0; j < 20; ++j) c[j] = 0.0f;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const cpx y = cmul_pp(x[k1], cmul_pp(cload(s_win2 + 2 * (N2 * k1 + k2)), cmake(INV_SCALE, -INV_SCALE)));
        c[k1] += cre(y);           // frame tA sample 40 k1 + k2
        c[k1 + 4] += cim(y);       // frame tA + 1 (= -Im / 640), one hop (4 rows) later
    }
}

// Pass B as ONE rolled loop over four column rounds (r = 0, 1: the main rounds f = r with k2' = lane; r = 2, 3: the two ordered
// side phases) so that the column code (DFT-16 + window) exists once in the instruction stream instead of three times.
// Only the 20 accumulations differ per round.  Must be followed by a __syncwarp() for r >= 2 (the caller's loop does it).
AVSE_HD void inv_stage_passB_round(int lane, int r, const float* s_win2, const float* frames, float (&acc)[INV_SIDE_ROWS], float* side) {
    const bool main = r < 2;
    const int f = main ? r : (lane >> 3);
    const int k2 = main ? lane : 32 + (lane & 7);
    if (!main && (lane >= 16 || f != r - 2)) return;
    float c[20];
    inv_passB_column(f, k2, s_win2, frames, c);
    if (main) {
        if (r == 0) {
#pragma unroll
            for (int j = 0; j < 20; ++j) acc[j] += c[j];
        } else {
#pragma unroll
            for (int j = 0; j < 20; ++j) acc[j + 8] += c[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 20; ++j) side[(j + 8 * f) * 8 + (lane & 7)] += c[j];
    }
}

AVSE_HD void inv_stage_passB_main(int lane, const float* s_win2, const float* frames, float (&acc)[INV_SIDE_ROWS]) {
#pragma unroll
    for (int f = 0; f < 2; ++f) {
        float c[20];
        inv_passB_column(f, lane, s_win2, frames, c);
#pragma unroll
        for (int j = 0; j < 20; ++j) acc[j + 8 * f] += c[j];
    }
}

// side-buffer phase ph (0 or 1): lanes with f == ph add their column into side[J][r]
AVSE_HD void inv_stage_passB_side(int lane, int ph, const float* s_win2, const float* frames, float* side) {
    if (lane >= 16) return;
    const int f = lane >> 3, r = lane & 7;
    if (f != ph) return;
    float c[20];
    inv_passB_column(f, 32 + r, s_win2, frames, c);
#pragma unroll
    for (int j = 0; j < 20; ++j) side[(j + 8 * f) * 8 + r] += c[j];
}

// Last store of the path: float32 PCM, or int16 like AudioSignal.save_to_wav_file (se:176-177):
// np.clip(x, -32768, 32767).astype(int16), i.e. clip then truncate toward zero.
AVSE_HD void store_pcm(float* p, float v) { *p = v; }
AVSE_HD void store_pcm(short* p, float v) {
    v = v < -32768.0f ? -32768.0f : (v > 32767.0f ? 32767.0f : v);
    *p = (short)(int)v;      // C conversion truncates toward zero like numpy's astype
}

// librosa.istft window sum-square at padded position P for T_use frames (Appendix A.1): sum over the
// (<= 4) frames t = P/160 - q that exist.  Interior value is exactly 1.5 for the periodic Hann at hop N/4.
AVSE_HD float inv_wss_recip(int P, int T_use, const float* s_win2) {
    const int h = P / HOP, r = P - h * HOP;
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int t = h - q;
        const float w = s_win2[2 * (r + HOP * q)];
        if (t >= 0 && t < T_use) s += w * w;
    }
    return s > 1.17549435e-38f ? 1.0f / s : 1.0f;   // "> tiny(float32)" guard of librosa.istft
}

// Emit the 4 finished hops of the group (rows J = 0..15) and rotate the overlap-add state.
// out: trimmed PCM of this utterance (index o = P - 320), out_len = 160 (T_use - 1).
template <typename O>
AVSE_HD void inv_stage_emit_main(int lane, int t0, int T_use, int out_len, bool write, const float* s_win2, O* out,
                                 float (&acc)[INV_SIDE_ROWS]) {
    if (write) {
        // hot case: every row has its 4 frames and lies inside the trimmed output -> straight-line stores, no cold code
        // in the instruction stream (the per-row edge branches showed up as instruction-fetch stalls in ncu)
        const bool interior = t0 >= 3 && t0 + 3 < T_use && t0 * HOP - HALF >= 0 && t0 * HOP + N2 * 15 + 31 - HALF < out_len;
        if (interior) {
            O* o = out + (t0 * HOP - HALF + lane);
#pragma unroll
            for (int J = 0; J < 16; ++J) store_pcm(o + N2 * J, acc[J] * (1.0f / 1.5f));
        } else {
#pragma unroll 1
            for (int J = 0; J < 16; ++J) {
                const int P = t0 * HOP + N2 * J + lane;
                const int o = P - HALF;
                float a = acc[0];
#pragma unroll
                for (int q = 1; q < 16; ++q) a = J == q ? acc[q] : a;      // register select (rolled cold loop)
                if (o >= 0 && o < out_len) store_pcm(out + o, a * inv_wss_recip(P, T_use, s_win2));
            }
        }
    }
#pragma unroll
    for (int J = 0; J < 12; ++J) acc[J] = acc[J + 16];
#pragma unroll
    for (int J = 12; J < INV_SIDE_ROWS; ++J) acc[J] = 0.0f;
}

// side buffer: lane = (row J = lane/2, half = lane%2): 4 samples each, then rotate rows.
template <typename O>
AVSE_HD void inv_stage_emit_side(int lane, int t0, int T_use, int out_len, bool write, const float* s_win2, O* out,
                                 const float* side, float (&keep)[4], float (&carry)[4]) {
    const int J = lane >> 1, hf = lane & 1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        keep[e] = side[J * 8 + 4 * hf + e];
        carry[e] = (J + 16 < INV_SIDE_ROWS) ? side[(J + 16) * 8 + 4 * hf + e] : 0.0f;
    }
    if (write) {
        const bool interior = t0 >= 3 && t0 + 3 < T_use && t0 * HOP - HALF >= 0 && t0 * HOP + N2 * 15 + 39 - HALF < out_len;
        const int P0 = t0 * HOP + N2 * J + 32 + 4 * hf;
        if (interior) {
#pragma unroll
            for (int e = 0; e < 4; ++e) store_pcm(out + (P0 - HALF + e), keep[e] * (1.0f / 1.5f));
        } else {
#pragma unroll 1
            for (int e = 0; e < 4; ++e) {
                const int o = P0 + e - HALF;
                const float k = e == 0 ? keep[0] : (e == 1 ? keep[1] : (e == 2 ? keep[2] : keep[3]));
                if (o >= 0 && o < out_len) store_pcm(out + o, k * inv_wss_recip(P0 + e, T_use, s_win2));
            }
        }
    }
}

AVSE_HD void inv_stage_rotate_side(int lane, float* side, const float (&carry)[4]) {
    const int J = lane >> 1, hf = lane & 1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        side[J * 8 + 4 * hf + e] = carry[e];                 // rows 0..15 <- rows 16..31 (rows >= 28 are zero)
        if (J + 16 < INV_SIDE_ROWS) side[(J + 16) * 8 + 4 * hf + e] = 0.0f;
    }
}

}  // namespace avse
