// Warp-stage functions of the fast inverse kernel ("I8": one warp owns EIGHT consecutive real frames = four packed complex
// FFTs per group), the predict-time path reconstruct_speech_signal (/root/reference/data_processor.py:60-74, :99-116).
//
// Same arithmetic as avse_inv_stages.cuh (mixture frames (t, t+1) packed as re / im for the phase, lin = F^T c, V = conj(Y_t +
// i Y_{t+1}) through the forward codelets, overlap-add in registers), re-mapped onto the SM the way the forward F4 kernel is
// (avse_fwd4_stages.cuh), because the round-1 ncu profile of the 4-frame kernel showed where its 1 119 instructions per real frame
// went (profiles/inverse_kernel_by_stage_r1.txt):
//
//   * pass 1 / pass B: lane = column n2 (k2') for ALL four complex FFTs, so the 16 window values and 15 twiddles of a lane are
//     loop-invariant REGISTERS (the 4-frame kernel re-loaded 31 table values per column from shared memory), and the four FFTs
//     share one batch of 44 strided sample loads per lane;
//   * the 8 tail columns (n2 = 32..39) of the four FFTs fill ONE full round of 32 lanes -- the 4-frame kernel ran its tail
//     rounds with 16 (pass 1) and 8 + 8 (pass B side phases) active lanes: 120 of its instructions per frame;
//   * the next group's samples are loaded into registers before pass B (software pipelining: long-scoreboard was 0.5 stall
//     cycles per issue);
//   * pass B is a rolled loop over the four FFTs: add, emit the two finished hops, rotate the 20-row accumulator -- one copy of
//     the column code, and pass 2 / pass A run through one copy of the DFT-40 (the hot loop must fit the 32 KB L1.5 I-cache);
//   * the irfft's 1/640 is folded into the F^T tap weights when the table is staged.
//
// 8 warps per SM, up to 255 registers.  Every function is __host__ __device__ (tests/emul runs a warp as a loop over lanes).
#pragma once
#include "avse_common.h"
#include "avse_dft.cuh"
#include "avse_fwd_stages.cuh"
#include "avse_fwd4_stages.cuh"
#include "avse_inv_stages.cuh"
#include "avse_tables.h"

#if !defined(AVSE_I8_SKIP_FFTS)
#define AVSE_I8_SKIP_FFTS 1      // skip the FFTs of a group that cannot matter (warm-up groups: 0 and 1; frames beyond T_use)
#endif
#if !defined(AVSE_I8_POST_WALK)
#define AVSE_I8_POST_WALK 1      // post stage: the coefficient pair of a lane's current band walks in registers (mask-driven advance)
#endif
#if !defined(AVSE_I8_POST_UNROLL)
#define AVSE_I8_POST_UNROLL 2    // bins per unrolled block of the post stage (independent load -> rsqrt -> store chains in flight)
#endif
#if !defined(AVSE_I8_EDGE_OUT_OF_LINE)
#define AVSE_I8_EDGE_OUT_OF_LINE 0   // 1: the cold reflect / zero-pad pass-1 stage as a real function call (one copy, outside the hot loop)
#endif
#if !defined(AVSE_I8_TW_IN_B)
#define AVSE_I8_TW_IN_B 1        // inter-pass twiddles applied on pass B's column loads (lane-resident registers) instead of in pass A
#endif
#if !defined(AVSE_I8_PREFETCH_MEL)
#define AVSE_I8_PREFETCH_MEL 0
#endif
#if !defined(AVSE_I8_ROLL_P1)
#define AVSE_I8_ROLL_P1 1        // pass 1: one copy of the column code, the 20-stride window slides through the raw registers
#endif

namespace avse {

constexpr int I8_FPG = 8;                          // real frames per group
constexpr int I8_NC = 4;                           // packed complex FFTs per group
constexpr int I8_RAW = 16 + 4 * (I8_FPG - 1);      // 44 strides of 40 samples cover the eight frames of a group
#if !defined(AVSE_I8_YS)
#define AVSE_I8_YS 10            // floats per band row of the coefficient buffer (8 used): 10 spreads the post stage's band gathers
#endif                           // over 16 bank groups instead of 4 (8 = dense rows)
constexpr int I8_YS = AVSE_I8_YS;
constexpr int I8_Y_F = NMEL * I8_YS;               // coefficient buffer [80][I8_YS]; re-used as the tail-column staging [4][20][8]
constexpr int I8_ST_C = I8_YS == 8 ? 160 : 168;    // staging stride between FFTs: 168 = 8 (mod 32) keeps the four lanes groups apart
constexpr int I8_ACC = 20;                         // accumulator rows (40 samples each) alive while one FFT is added
constexpr int I8_SIDE_ROWS = 12;                   // tail-column rows carried from one group to the next
constexpr int I8_SIDE_F = I8_SIDE_ROWS * 8;        // 96 floats, double-buffered
constexpr int I8_FLAG_F = N1 * ROW_F;              // per frame buffer: floats [1344, 1346) = "windowed frame A / B is non-zero"
constexpr int I8_XCH_F = 2 * SPIKE_P * I8_FPG;     // 64 floats: (top, bottom) of every partition's local solve, per frame
constexpr int I8_WARP_SMEM_F = I8_NC * FRAME4_F + I8_Y_F + 2 * I8_SIDE_F + I8_XCH_F;     // 5440 + 800 + 192 + 64 = 6496 floats
static_assert(3 * I8_ST_C + 160 <= I8_Y_F && (I8_YS % 2) == 0, "the tail staging re-uses the coefficient buffer");
static_assert(I8_FLAG_F + 2 <= FRAME4_F, "flags live in the frame buffer's pad");

// librosa.istft window sum-square at padded position P (see inv_wss_recip), plain [640] window table.
AVSE_HD float inv_wss_recip_w(int P, int T_use, const float* s_win) {
    const int h = P / HOP, r = P - h * HOP;
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int t = h - q;
        const float w = s_win[r + HOP * q];
        if (t >= 0 && t < T_use) s += w * w;
    }
    return s > 1.17549435e-38f ? 1.0f / s : 1.0f;   // "> tiny(float32)" guard of librosa.istft
}

// A group is "interior" when its eight frames exist and need neither reflection nor zero padding.
AVSE_HD bool i8_group_interior(const InvTile& tl) {
    return tl.t0 * HOP - HALF >= 0 && (tl.t0 + I8_FPG - 1) * HOP + HALF <= tl.valid && tl.t0 + I8_FPG - 1 < tl.T;
}

// The 44 strided samples a main lane needs for the four FFTs, and the 20 of a tail lane (c = lane / 8, n2 = 32 + lane % 8).
AVSE_HD void i8_load_raw(const InvTile& tl, int lane, float (&raw)[I8_RAW]) {
    const float* p = tl.pcm + (tl.t0 * HOP - HALF + lane);
#pragma unroll
    for (int j = 0; j < I8_RAW; ++j) raw[j] = p[N2 * j];
}

AVSE_HD void i8_load_tail_raw(const InvTile& tl, int lane, float (&rt)[20]) {
    const float* p = tl.pcm + ((tl.t0 + 2 * (lane >> 3)) * HOP - HALF + 32 + (lane & 7));
#pragma unroll
    for (int j = 0; j < 20; ++j) rt[j] = p[N2 * j];
}

// Reflect-only edge groups (AVSE_I8_REFLECT_FAST, like the forward kernel's AVSE_F4_REFLECT_FAST): with a full-length mixture
// (no zero padding) the first and the last group of an utterance differ from an interior one only in the mirrored sample index of
// librosa's centre padding (dp:79), so they take the interior pass 1 on mirrored loads instead of the per-sample edge loader.
// Frames past the last mixture frame read mirrored samples too: their coefficients are zero (or their FFT is skipped).
#if !defined(AVSE_I8_REFLECT_FAST)
#define AVSE_I8_REFLECT_FAST 1
#endif
#if !defined(AVSE_I8_PAD_FAST)
#define AVSE_I8_PAD_FAST 1        // zero-padded mixtures too: guarded variant of the mirrored loader (indices at or past `valid` read 0)
#endif
AVSE_HD bool i8_group_reflect_only(const InvTile& tl) {
    return (AVSE_I8_PAD_FAST || tl.valid >= tl.L) && tl.L >= 4 * NFFT;      // every mirrored index (phantom frames included) stays inside [0, L)
}

AVSE_HD int i8_reflect_index(int i, int L) {
    i = i < 0 ? -i : i;
    const int m = 2 * (L - 1) - i;
    return i < m ? i : m;
}

AVSE_HD void i8_load_raw_reflect(const InvTile& tl, int lane, float (&raw)[I8_RAW]) {
    const int o = tl.t0 * HOP - HALF + lane;
    if (!AVSE_I8_PAD_FAST || tl.valid >= tl.L) {
#pragma unroll
        for (int j = 0; j < I8_RAW; ++j) raw[j] = tl.pcm[i8_reflect_index(o + N2 * j, tl.L)];
    } else {
#pragma unroll
        for (int j = 0; j < I8_RAW; ++j) { const int i = i8_reflect_index(o + N2 * j, tl.L); raw[j] = i < tl.valid ? tl.pcm[i] : 0.0f; }
    }
}

AVSE_HD void i8_load_tail_raw_reflect(const InvTile& tl, int lane, float (&rt)[20]) {
    const int o = (tl.t0 + 2 * (lane >> 3)) * HOP - HALF + 32 + (lane & 7);
    if (!AVSE_I8_PAD_FAST || tl.valid >= tl.L) {
#pragma unroll
        for (int j = 0; j < 20; ++j) rt[j] = tl.pcm[i8_reflect_index(o + N2 * j, tl.L)];
    } else {
#pragma unroll
        for (int j = 0; j < 20; ++j) { const int i = i8_reflect_index(o + N2 * j, tl.L); rt[j] = i < tl.valid ? tl.pcm[i] : 0.0f; }
    }
}

// Non-zero flags of the two real frames packed in one FFT, from the raw sample bits (see inv_mark_nonzero_raw): frame A uses
// strides 0..15, frame B strides 4..19 of r[]; w[n] != 0 except n = 0, which only column n2 = 0 holds (stride 0 of its frame).
AVSE_HD void i8_mark(const float* r, int n2, float* frame_base) {
    int mid = 0;
#pragma unroll
    for (int j = 5; j < 16; ++j) mid |= float_bits(r[j]);
    const int r4 = float_bits(r[4]);
    const int a = mid | (n2 != 0 ? float_bits(r[0]) : 0) | float_bits(r[1]) | float_bits(r[2]) | float_bits(r[3]) | r4;
    const int b = mid | (n2 != 0 ? r4 : 0) | float_bits(r[16]) | float_bits(r[17]) | float_bits(r[18]) | float_bits(r[19]);
    if ((a & 0x7fffffff) != 0) frame_base[I8_FLAG_F] = 1.0f;
    if ((b & 0x7fffffff) != 0) frame_base[I8_FLAG_F + 1] = 1.0f;
}

// The same flags for the eight frames of a main lane (column n2 = lane) in one go, from the 44 raw strides: frame t uses strides
// [4t, 4t + 16); its n = 0 sample (w[0] = 0) is stride 4t of column 0.  ORs of 4-stride blocks are shared between the frames
// (~50 logic instructions per group; marking inside the rolled column loop cost 55 instructions per frame).
AVSE_HD void i8_mark_group(const float (&raw)[I8_RAW], int lane, float* frames) {
    int rest[11], blk[11];       // rest[b] = OR of strides 4b+1 .. 4b+3, blk[b] = rest[b] | stride 4b
#pragma unroll
    for (int b = 0; b < 11; ++b) {
        rest[b] = float_bits(raw[4 * b + 1]) | float_bits(raw[4 * b + 2]) | float_bits(raw[4 * b + 3]);
        blk[b] = rest[b] | float_bits(raw[4 * b]);
    }
#pragma unroll
    for (int t = 0; t < I8_FPG; ++t) {
        const int first = lane != 0 ? float_bits(raw[4 * t]) : 0;
        const int any = rest[t] | first | blk[t + 1] | blk[t + 2] | blk[t + 3];
        if ((any & 0x7fffffff) != 0) frames[(t >> 1) * FRAME4_F + I8_FLAG_F + (t & 1)] = 1.0f;
    }
}

// pass 1, interior groups, columns n2 = lane of the four FFTs.  FFT c packs frames (t0 + 2c, t0 + 2c + 1): strides [8c, 8c+16)
// and [8c+4, 8c+20) of the batch.  Consumes (rotates) raw[] when rolled.
// [c_lo, c_hi): the FFTs that are needed (a warm-up group: 2, 3 -- FFTs 0 and 1 cannot reach the carry; a last group: those with
// frames below T_use; see the kernel).
AVSE_HD void i8_pass1_main(int lane, float (&raw)[I8_RAW], const Lane4Const& lc, float* frames, int c_lo, int c_hi) {
    i8_mark_group(raw, lane, frames);
#if AVSE_I8_ROLL_P1 == 2      // two FFTs per iteration: half the register rotation (28 instead of 2 x 36 moves per pair)
    float* dst = frames + 2 * lane;
#pragma unroll 1
    for (int c = 0; c < I8_NC; c += 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (c + h >= c_lo && c + h < c_hi) {
                cpx x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = cmake(raw[8 * h + j] * lc.win[j], raw[8 * h + j + 4] * lc.win[j]);
                p4_column(x, lc.tw, dst + h * FRAME4_F);
            }
        }
        dst += 2 * FRAME4_F;
#pragma unroll
        for (int j = 0; j + 16 < I8_RAW; ++j) raw[j] = raw[j + 16];
    }
#elif AVSE_I8_ROLL_P1
    float* dst = frames + 2 * lane;
#pragma unroll 1
    for (int c = 0; c < I8_NC; ++c) {
        if (c >= c_lo && c < c_hi) {
            cpx x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = cmake(raw[j] * lc.win[j], raw[j + 4] * lc.win[j]);
            p4_column(x, lc.tw, dst);
        }
        dst += FRAME4_F;
#pragma unroll
        for (int j = 0; j + 8 < I8_RAW; ++j) raw[j] = raw[j + 8];
    }
#else
#pragma unroll
    for (int c = 0; c < I8_NC; ++c) {
        if (c < c_lo || c >= c_hi) continue;
        cpx x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = cmake(raw[8 * c + j] * lc.win[j], raw[8 * c + j + 4] * lc.win[j]);
        p4_column(x, lc.tw, frames + c * FRAME4_F + 2 * lane);
    }
#endif
}

// pass 1, interior groups, the fifth round: lane = (c, r), column n2 = 32 + r of FFT c; window / twiddles from the CTA's tables.
AVSE_HD void i8_pass1_tail(int lane, const float (&rt)[20], const float* s_win, const vec2* s_tw, float* frames) {
    const int c = lane >> 3, n2 = 32 + (lane & 7);
    vec2 tw[16];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) tw[k1] = s_tw[k1 * N2 + n2];
    tw[0].x = 1.0f; tw[0].y = 0.0f;
    cpx x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { const float w = s_win[N2 * j + n2]; x[j] = cmake(rt[j] * w, rt[j + 4] * w); }
    i8_mark(rt, n2, frames + c * FRAME4_F);
    p4_column(x, tw, frames + c * FRAME4_F + 2 * n2);
}

// pass 1, edge groups (first / last frames of an utterance, zero padding): every sample through the reflect + pad loader; frames
// beyond the last one are clamped (their coefficients are zero).  Cold code, rolled over the five rounds.
#if AVSE_I8_EDGE_OUT_OF_LINE
#define AVSE_I8_EDGE_Q AVSE_HD_COLD
#else
#define AVSE_I8_EDGE_Q AVSE_HD
#endif
AVSE_I8_EDGE_Q void i8_pass1_edge(const InvTile tl, int lane, const float* s_win, const vec2* s_tw, float* frames, int c_lo, int c_hi) {
#pragma unroll 1
    for (int round = 0; round < 5; ++round) {
        if (round < 4 && (round < c_lo || round >= c_hi)) continue;     // FFTs outside [c_lo, c_hi) are not needed (see the kernel)
        const int c = round < 4 ? round : lane >> 3;
        const int n2 = round < 4 ? lane : 32 + (lane & 7);
        const int tA = tl.t0 + 2 * c;
        const int ta = tA < tl.T ? tA : tl.T - 1;
        const int tb = tA + 1 < tl.T ? tA + 1 : tl.T - 1;
        const int ba = ta * HOP - HALF + n2, bb = tb * HOP - HALF + n2;
        cpx x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float w = s_win[N2 * j + n2];
            x[j] = cmake(load_sample_edge(tl.pcm, ba + N2 * j, tl.L, tl.valid) * w, load_sample_edge(tl.pcm, bb + N2 * j, tl.L, tl.valid) * w);
        }
        bool nr = false, ni = false;
#pragma unroll
        for (int j = 0; j < 16; ++j) { nr = nr || (cre(x[j]) != 0.0f); ni = ni || (cim(x[j]) != 0.0f); }
        if (nr) frames[c * FRAME4_F + I8_FLAG_F] = 1.0f;
        if (ni) frames[c * FRAME4_F + I8_FLAG_F + 1] = 1.0f;
        vec2 tw[16];
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) tw[k1] = s_tw[k1 * N2 + n2];
        tw[0].x = 1.0f; tw[0].y = 0.0f;
        p4_column(x, tw, frames + c * FRAME4_F + 2 * n2);
    }
}

// ---------------------------------------------------------------------------------------
// Coefficients c = (F F^T)^-1 10^(dB/20) of the group's eight frames, inside the kernel (round 1 and the 4-frame kernel run a
// separate kernel over all frames: 48 us plus a 96 MB write + read through HBM).  Replaces np.linalg.pinv + np.dot of dp:112
// together with db_to_amplitude (dp:101).  lane = (frame f = lane / 4, partition p = lane % 4): the 80-band tridiagonal system
// is solved in "SPIKE" form -- four independent 20-band Thomas solves per frame, the 8 x 8 interface system (a constant matrix,
// inverted on the host in float64) applied to the partitions' (top, bottom) values, which the lanes of a frame exchange through
// 64 floats of shared memory, then one rank-2 correction per band with the precomputed spikes.  All 32 lanes busy, two dependent
// chains of 20 steps instead of 80.  s_spk: [4][SPIKE_ROW] (avse_tables.h).
// Three stages: load (issued early, the values ride through pass 1), local solve, correction + store into ybuf[band][I8_YS].
// ---------------------------------------------------------------------------------------
AVSE_HD void i8_coef_load(int lane, const float* mel, int layout, int ld_t, int t0, int T_use, float (&d)[SPIKE_Q]) {
    const int f = lane >> 2, p = lane & 3, t = t0 + f;
    if (t >= T_use) {                   // frames of the padded last group: 10^(-inf / 20) = 0 amplitude, zero coefficients
#pragma unroll
        for (int i = 0; i < SPIKE_Q; ++i) d[i] = -1.0e30f;
        return;
    }
    const float* q;
    int mstride;
    if (layout == 0) {                  // AVSE_LAYOUT_SLICES [n_slices][80][20]
        const int sl = t / SPSS, tt = t - sl * SPSS;
        q = mel + ((size_t)sl * NMEL + SPIKE_Q * p) * SPSS + tt;
        mstride = SPSS;
    } else {                            // AVSE_LAYOUT_SPEC [80][ld_t]
        q = mel + (size_t)(SPIKE_Q * p) * ld_t + t;
        mstride = ld_t;
    }
#pragma unroll
    for (int i = 0; i < SPIKE_Q; ++i) d[i] = q[(size_t)i * mstride];
#if defined(__CUDA_ARCH__) && AVSE_I8_PREFETCH_MEL
    // L2 prefetch of the same bands two groups ahead (SPEC layout: 16 floats further; SLICES: the frame 16 later, possibly in the
    // next slice).  One lane per (partition, band-row sector): the frame lanes f = 0 of each partition issue them.
    if ((lane >> 2) == 0) {
        const int t2 = t + 2 * I8_FPG;
        if (t2 < T_use) {
            const float* q2;
            if (layout == 0) { const int sl = t2 / SPSS, tt = t2 - sl * SPSS; q2 = mel + ((size_t)sl * NMEL + SPIKE_Q * p) * SPSS + tt; }
            else q2 = q + 2 * I8_FPG;
#pragma unroll
            for (int i = 0; i < SPIKE_Q; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(q2 + (size_t)i * mstride));
        }
    }
#endif
}

AVSE_HD void i8_coef_local(int lane, const float* s_spk, float (&d)[SPIKE_Q], float* xch) {
    constexpr float K = 0.16609640474436813f;   // log2(10) / 20 :  10^(dB/20) = 2^(K dB)   (librosa.db_to_amplitude)
    const int f = lane >> 2, p = lane & 3;
    const float* tb = s_spk + p * SPIKE_ROW;
#pragma unroll
    for (int i = 0; i < SPIKE_Q; ++i) d[i] = inv_exp2(K * d[i]);
#pragma unroll
    for (int i = 1; i < SPIKE_Q; ++i) d[i] = fmaf(-tb[i], d[i - 1], d[i]);
    d[SPIKE_Q - 1] *= tb[SPIKE_Q + SPIKE_Q - 1];
#pragma unroll
    for (int i = SPIKE_Q - 2; i >= 0; --i) d[i] = fmaf(-tb[2 * SPIKE_Q + i], d[i + 1], d[i]) * tb[SPIKE_Q + i];
    xch[8 * f + 2 * p] = d[0];
    xch[8 * f + 2 * p + 1] = d[SPIKE_Q - 1];
}

AVSE_HD void i8_coef_finish(int lane, const float* s_spk, const float (&d)[SPIKE_Q], const float* xch, float* ybuf) {
    const int f = lane >> 2, p = lane & 3;
    const float* tb = s_spk + p * SPIKE_ROW;
    const float* rb = tb + 5 * SPIKE_Q;
    const float* rt = rb + 2 * SPIKE_P;
    float bprev = 0.0f, tnext = 0.0f;           // x[20 p - 1] and x[20 p + 20] of the full solution
#pragma unroll
    for (int k = 0; k < 2 * SPIKE_P; ++k) {
        const float y = xch[8 * f + k];
        bprev = fmaf(rb[k], y, bprev);
        tnext = fmaf(rt[k], y, tnext);
    }
    float* dst = ybuf + I8_YS * (SPIKE_Q * p) + f;
#pragma unroll
    for (int i = 0; i < SPIKE_Q; ++i)
        dst[I8_YS * i] = fmaf(-tb[4 * SPIKE_Q + i], tnext, fmaf(-tb[3 * SPIKE_Q + i], bprev, d[i]));
}

// ---------------------------------------------------------------------------------------
// post: lane = (FFT c = lane / 8, chunk p = lane % 8), bins k = 41 p + i.  See inv_stage_post for the arithmetic; here the tap
// weights carry the irfft's 1/640.  s_col: [SCAN4_BINS] (b0, b1, w0, w1) bit patterns; ybuf: [80][8] coefficients of the group.
// ---------------------------------------------------------------------------------------
template <bool EXT>
AVSE_HD void i8_post_core(int i, int p, cpx lin, float* za, float* zc, bool liveA, bool liveB, const vec2* phA, const vec2* phB) {
    const int k = CHUNK4 * p + i;
    const bool tail = k > NBINS - 1;            // chunk 7 ends with slots that mirror other bins: compute harmlessly, do not store
    const bool nyq = k == NBINS - 1;            // lin = 0 at the Nyquist bin (the filterbank's last column is empty)
    float yar, yai, ybr, ybi;
    if (EXT) {
        const vec2 qa = (phA != nullptr && !tail) ? phA[k] : vec2{1.0f, 0.0f};
        const vec2 qb = (phB != nullptr && !tail) ? phB[k] : vec2{1.0f, 0.0f};
        yar = cre(lin) * qa.x; yai = cre(lin) * qa.y;
        ybr = cim(lin) * qb.x; ybi = cim(lin) * qb.y;
    } else {
        const cpx a = cload(za + 2 * i);
        const cpx c = cload(zc - 2 * i);
        const cpx xa = cfma_pp(c, cmake(1.0f, -1.0f), a);             // 2 X_A = Z_k + conj Z_{N-k}
        const cpx xn = cfma_pp(c, cmake(-1.0f, 1.0f), a);             // 2 i X_B = Z_k - conj Z_{N-k}
        const float na = cre(xa) * cre(xa) + cim(xa) * cim(xa), nb = cre(xn) * cre(xn) + cim(xn) * cim(xn);
        const float sa = cre(lin) * inv_rsqrt(na), sb = cim(lin) * inv_rsqrt(nb);
        const bool okA = na > 0.0f && liveA, okB = nb > 0.0f && liveB;   // 1 + 0j where X == 0 (librosa.magphase, dp:80)
        yar = okA ? cre(xa) * sa : cre(lin); yai = okA ? cim(xa) * sa : 0.0f;
        ybr = okB ? cim(xn) * sb : cim(lin); ybi = okB ? -cre(xn) * sb : 0.0f;
    }
    if (nyq) cstore(za + 2 * i, cmake(0.0f, 0.0f));
    if (!tail) {
        cstore(za + 2 * i, cmake(yar - ybi, -(yai + ybr)));   // conj(Y_A + i Y_B)
        cstore(zc - 2 * i, cmake(yar + ybi, yai - ybr));      // conj(conj(Y_A) + i conj(Y_B))
    }
}

#if AVSE_I8_POST_WALK
// Walk form: s_col = [SCAN4_BINS] (w0, w1) / 640, followed by [8] (mask lo, mask hi, first band, -) per chunk.  Bit i of the
// chunk's mask: the pair (y[b], y[b + 1]) moves up one band before bin i.  The pair lives in registers; an advance is one
// predicated 8-byte load instead of two gathers and a 16-byte table entry per bin.
template <bool EXT>
AVSE_HD void i8_stage_post(int lane, const ivec4* s_col, const float* ybuf, float* frames, const vec2* phA, const vec2* phB) {
    const int c = lane >> 3, p = lane & 7;
    float* fr = frames + c * FRAME4_F;
    float* za = fr + 2 * CHUNK4 * p;
    float* zc = fr + 2 * (NFFT - CHUNK4 * p);
    const vec2* tab = reinterpret_cast<const vec2*>(s_col) + CHUNK4 * p;
    const ivec4 ch = *(reinterpret_cast<const ivec4*>(reinterpret_cast<const vec2*>(s_col) + SCAN4_BINS) + p);
    unsigned mlo = (unsigned)ch.x, mhi = (unsigned)ch.y;
    const float* yb = ybuf + 2 * c + I8_YS * ch.z;            // (c_A, c_B) of band b at ybuf[I8_YS b + 2 c]
    cpx y0 = cload(yb), y1 = cload(yb + I8_YS);
    const bool liveA = EXT || fr[I8_FLAG_F] != 0.0f;
    const bool liveB = EXT || fr[I8_FLAG_F + 1] != 0.0f;
    static_assert(CHUNK4 == 41 && (40 % AVSE_I8_POST_UNROLL) == 0 && AVSE_I8_POST_UNROLL <= 8, "blocks of AVSE_I8_POST_UNROLL bins + 1");
#pragma unroll 1
    for (int ib = 0; ib < 40; ib += AVSE_I8_POST_UNROLL) {
#pragma unroll
        for (int j = 0; j < AVSE_I8_POST_UNROLL; ++j) {
            if ((mlo >> j) & 1u) { yb += I8_YS; y0 = y1; y1 = cload(yb + I8_YS); }
            const vec2 w = tab[ib + j];
            const cpx lin = cfma_s(y1, w.y, cmul_s(y0, w.x));         // (lin_A, lin_B) / 640
            i8_post_core<EXT>(ib + j, p, lin, za, zc, liveA, liveB, phA, phB);
        }
        mlo = (mlo >> AVSE_I8_POST_UNROLL) | (mhi << (32 - AVSE_I8_POST_UNROLL));
        mhi >>= AVSE_I8_POST_UNROLL;
    }
    if (mlo & 1u) { yb += I8_YS; y0 = y1; y1 = cload(yb + I8_YS); }
    const vec2 w = tab[40];
    i8_post_core<EXT>(40, p, cfma_s(y1, w.y, cmul_s(y0, w.x)), za, zc, liveA, liveB, phA, phB);
}
#else
template <bool EXT>
AVSE_HD void i8_post_bin(int i, int p, const ivec4* tab, const float* yb, float* za, float* zc, bool liveA, bool liveB, const vec2* phA,
                         const vec2* phB) {
    const ivec4 t = tab[i];
    const cpx y0 = cload(yb + I8_YS * t.x), y1 = cload(yb + I8_YS * t.y);
    const cpx lin = cfma_s(y1, bits_to_float(t.w), cmul_s(y0, bits_to_float(t.z)));   // (lin_A, lin_B) / 640
    i8_post_core<EXT>(i, p, lin, za, zc, liveA, liveB, phA, phB);
}

template <bool EXT>
AVSE_HD void i8_stage_post(int lane, const ivec4* s_col, const float* ybuf, float* frames, const vec2* phA, const vec2* phB) {
    const int c = lane >> 3, p = lane & 7;
    float* fr = frames + c * FRAME4_F;
    float* za = fr + 2 * CHUNK4 * p;
    float* zc = fr + 2 * (NFFT - CHUNK4 * p);
    const ivec4* tab = s_col + CHUNK4 * p;
    const float* yb = ybuf + 2 * c;            // (c_A, c_B) of band b at yb[I8_YS b]
    const bool liveA = EXT || fr[I8_FLAG_F] != 0.0f;
    const bool liveB = EXT || fr[I8_FLAG_F + 1] != 0.0f;
    static_assert(CHUNK4 == 41 && (40 % AVSE_I8_POST_UNROLL) == 0, "blocks of AVSE_I8_POST_UNROLL bins + 1");
#pragma unroll 1
    for (int ib = 0; ib < 40; ib += AVSE_I8_POST_UNROLL) {
#pragma unroll
        for (int j = 0; j < AVSE_I8_POST_UNROLL; ++j) i8_post_bin<EXT>(ib + j, p, tab, yb, za, zc, liveA, liveB, phA, phB);
    }
    i8_post_bin<EXT>(40, p, tab, yb, za, zc, liveA, liveB, phA, phB);
}
#endif

// ---------------------------------------------------------------------------------------
// pass A, round r: lane = (c = 2 r + lane / 16, n1' = lane % 16): gather V[n1' + 16 n2'], (DFT-40 by the caller), twiddle
// W_640^{n1' k2'}, store as row n1' = [k2'].  s_twT: [40][16] vec2.
// ---------------------------------------------------------------------------------------
AVSE_HD void i8_passA_load(int lane, int r, const float* frames, cpx (&x)[40]) {
    const int c = 2 * r + (lane >> 4), n1 = lane & 15;
    const float* z = frames + c * FRAME4_F + 2 * n1;
#pragma unroll
    for (int n2 = 0; n2 < 40; ++n2) x[n2] = cload(z + 2 * N1 * n2);
}

AVSE_HD void i8_passA_store(int lane, int r, float* frames, const cpx (&x)[40]) {
    const int c = 2 * r + (lane >> 4), n1 = lane & 15;
    float* row = frames + c * FRAME4_F + n1 * ROW_F;
    if (n1 == 0) { frames[c * FRAME4_F + I8_FLAG_F] = 0.0f; frames[c * FRAME4_F + I8_FLAG_F + 1] = 0.0f; }   // re-arm the flags
#pragma unroll
    for (int cc = 0; cc < 5; ++cc)
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int idx = (8 * cc + 5 * d) % 40, k2 = (16 * cc + 25 * d) % 40;
            cstore(row + 2 * k2, x[idx]);
        }
}

// ---------------------------------------------------------------------------------------
// pass B.  Column k2' of FFT c: DFT-16 over n1' gives (y_A - i y_B) at n = 40 k1' + k2' (already / 640).
// acc[J]: row J = samples 160 (t0 + 2c) + 40 J + lane, J = 0..19, while FFT c is being added.
// ---------------------------------------------------------------------------------------
// AVSE_I8_TW_IN_B: the inter-pass twiddle W_640^{n1' k2'} is applied HERE, on the column loads, instead of at the end of pass A:
// with lane = k2' the 15 factors are exactly the lane's pass-1 twiddle registers (the table is symmetric in its two indices), so
// pass A loses its 39 table loads + 39 complex multiplies per lane and pass B gains 15 multiplies per column -- same products,
// same rounding, bit-identical results.
AVSE_HD void i8_passB_add(int lane, int c, const Lane4Const& lc, const float* frames, float (&acc)[I8_ACC]) {
    const float* col = frames + c * FRAME4_F + 2 * lane;
    cpx x[16];
#if AVSE_I8_TW_IN_B
    x[0] = cload(col);
#pragma unroll
    for (int n1 = 1; n1 < 16; ++n1) x[n1] = cmul(cload(col + n1 * ROW_F), lc.tw[n1].x, lc.tw[n1].y);
#else
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) x[n1] = cload(col + n1 * ROW_F);
#endif
    dft16(x);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        acc[k1] = fmaf(cre(x[k1]), lc.win[k1], acc[k1]);              // frame A sample 40 k1 + lane
        acc[k1 + 4] = fmaf(-cim(x[k1]), lc.win[k1], acc[k1 + 4]);     // frame B, one hop (4 rows) later
    }
}

// Emit the two hops (rows 0..7) that are complete once FFT c has been added, then rotate the accumulator by 8 rows.
// h0 = t0 + 2c is the first of the two hops; out: trimmed PCM of this utterance (index o = P - 320).
template <typename O>
AVSE_HD void i8_emit_main(int lane, int h0, int T_use, int out_len, bool write, const float* s_win, O* out, float (&acc)[I8_ACC]) {
    if (write) {
        // hot case: both hops have their four frames and lie inside the trimmed output
        if (h0 >= 3 && h0 + 1 <= T_use - 1) {
            O* o = out + (h0 * HOP - HALF + lane);
#pragma unroll
            for (int J = 0; J < 8; ++J) store_pcm(o + N2 * J, acc[J] * (1.0f / 1.5f));
        } else {
#pragma unroll 1
            for (int J = 0; J < 8; ++J) {
                const int P = h0 * HOP + N2 * J + lane;
                const int o = P - HALF;
                float a = acc[0];
#pragma unroll
                for (int q = 1; q < 8; ++q) a = J == q ? acc[q] : a;      // register select (rolled cold loop)
                if (o >= 0 && o < out_len) store_pcm(out + o, a * inv_wss_recip_w(P, T_use, s_win));
            }
        }
    }
#pragma unroll
    for (int J = 0; J < 12; ++J) acc[J] = acc[J + 8];
#pragma unroll
    for (int J = 12; J < I8_ACC; ++J) acc[J] = 0.0f;
}

// Tail columns k2' = 32 + r: lane = (c, r) computes its column and parks the 20 row contributions in the staging buffer
// st[c][j][r] (the dead coefficient buffer); window from the CTA's table.
// FFTs outside [c_lo, c_hi) were not computed in this group: their (stale) rows contribute exact zeros.
AVSE_HD void i8_passB_tail(int lane, const float* s_win, const vec2* s_tw, const float* frames, float* st, int c_lo, int c_hi) {
    const int c = lane >> 3, r = lane & 7, k2 = 32 + r;
    const bool act = c >= c_lo && c < c_hi;
    const float* col = frames + c * FRAME4_F + 2 * k2;
    cpx x[16];
#if AVSE_I8_TW_IN_B
    x[0] = cload(col);
#pragma unroll
    for (int n1 = 1; n1 < 16; ++n1) { const vec2 t = s_tw[n1 * N2 + k2]; x[n1] = cmul(cload(col + n1 * ROW_F), t.x, t.y); }
#else
    (void)s_tw;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) x[n1] = cload(col + n1 * ROW_F);
#endif
    dft16(x);
    float cc[20];
#pragma unroll
    for (int j = 0; j < 20; ++j) cc[j] = 0.0f;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const float w = s_win[N2 * k1 + k2];
        cc[k1] = fmaf(cre(x[k1]), w, cc[k1]);
        cc[k1 + 4] = fmaf(-cim(x[k1]), w, cc[k1 + 4]);
    }
    float* d = st + c * I8_ST_C + r;
#pragma unroll
    for (int j = 0; j < 20; ++j) d[8 * j] = act ? cc[j] : 0.0f;
}

// A group that skipped some FFTs leaves their "frame is non-zero" flags set by pass 1's marks without the re-arm of pass A's
// store: clear all eight.
AVSE_HD void i8_rearm_flags(int lane, float* frames) {
    if (lane < I8_FPG) frames[(lane >> 1) * FRAME4_F + I8_FLAG_F + (lane & 1)] = 0.0f;
}

// Tail columns, second half: the group's 44 tail rows x 8 columns are summed and leave.  lane = (q = lane / 8, r) owns column
// r of the rows R = 8 m + q (m = 0..5) and R = 8 m + q + 4 (m = 0..4): with that split the set of FFTs that touch a row --
// c = m - 2 (only rows with R % 8 < 4), m - 1, m, clipped to 0..3 -- is known at compile time, so the whole stage is
// straight-line code of <= 3 loads + adds per row (a rolled loop with four predicated terms per row cost 77 instructions per
// frame).  Terms are added in a FIXED order (carry, then c ascending): bit-reproducible.  Rows R < 32 (the group's own eight
// hops) leave through the output, rows 32..43 become the carry of the next group (side_out row R - 32).
// have: the group was computed (false for the drain group: only the carry is emitted).
template <typename O>
AVSE_HD void i8_tail_reduce_emit(int lane, int t0, int T_use, int out_len, bool write, bool have, const float* s_win, O* out, const float* st,
                                 const float* side_in, float* side_out) {
    const int q = lane >> 3, r = lane & 7;
    const bool fast = t0 >= 3 && t0 + I8_FPG - 1 <= T_use - 1;      // every hop of the group has its four frames
    O* obase = out + (t0 * HOP - HALF + 32 + r);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int e = q + 4 * half;
#pragma unroll
        for (int m = 0; m < 6 - half; ++m) {
            const int R = 8 * m + e;
            float v = (m == 0 || (m == 1 && half == 0)) ? side_in[R * 8 + r] : 0.0f;          // rows < 12 carry over
            if (have) {
                if (half == 0 && m >= 2 && m - 2 < I8_NC) v += st[(m - 2) * I8_ST_C + (e + 16) * 8 + r];
                if (m >= 1 && m - 1 < I8_NC) v += st[(m - 1) * I8_ST_C + (e + 8) * 8 + r];
                if (m < I8_NC) v += st[m * I8_ST_C + e * 8 + r];
            }
            if (m < 4) {
                if (write) {
                    if (fast) {
                        store_pcm(obase + N2 * R, v * (1.0f / 1.5f));
                    } else {
                        const int h = t0 + 2 * m + half;
                        const int P = t0 * HOP + N2 * R + 32 + r;
                        const int o = P - HALF;
                        if (h >= 3 && h <= T_use - 1) store_pcm(out + o, v * (1.0f / 1.5f));
                        else if (o >= 0 && o < out_len) store_pcm(out + o, v * inv_wss_recip_w(P, T_use, s_win));
                    }
                }
            } else {
                side_out[(R - 32) * 8 + r] = v;
            }
        }
    }
}

}  // namespace avse
