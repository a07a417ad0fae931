// Host emulation of the forward kernel's warp stages (TEST INFRASTRUCTURE, not a product path).
// Builds the SAME __host__ __device__ stage functions with g++ and runs one warp as a loop over
// 32 lanes per stage; the loop boundaries are exactly the kernel's __syncwarp() points.  This
// lets the CPU-only test suite check the index maps, twiddles, tables and float32 arithmetic of
// csrc/avse_fwd_stages.cuh against the float64 oracle without a GPU.
#include <vector>
#include <cstring>
#include <cmath>
#include "../../audio-visual-speech-enhancement_b200/csrc/avse_common.h"
#include "../../audio-visual-speech-enhancement_b200/csrc/avse_tables.h"
#include "../../audio-visual-speech-enhancement_b200/csrc/avse_fwd_stages.cuh"

using namespace avse;

extern "C" int emul_filterbank(int sample_rate, double fmin, double fmax, double* fb_out) {
    HostTables h;
    if (!build_tables(h, sample_rate, fmin, fmax)) return -2;
    memcpy(fb_out, h.fb.data(), sizeof(double) * NMEL * NBINS);
    return 0;
}

extern "C" int emul_tables_info(int sample_rate, double fmin, double fmax, int* lo, int* width, float* tri /*3*80*/) {
    HostTables h;
    if (!build_tables(h, sample_rate, fmin, fmax)) return -2;
    memcpy(lo, h.mel_lo.data(), sizeof(int) * NMEL);
    memcpy(width, h.mel_width.data(), sizeof(int) * NMEL);
    memcpy(tri, h.tri_w.data(), sizeof(float) * NMEL);
    memcpy(tri + NMEL, h.tri_ipiv.data(), sizeof(float) * NMEL);
    memcpy(tri + 2 * NMEL, h.tri_sup.data(), sizeof(float) * NMEL);
    return 0;
}

// 640-point FFT of one complex frame through pass 1 (with unit window) + pass 2: checks the
// DFT-16 / twiddle / DFT-40 codelets and the shared-memory index maps in isolation.
extern "C" int emul_fft640(const float* re, const float* im, float* out_re, float* out_im) {
    HostTables h;
    if (!build_tables(h, 16000, 0.0, 8000.0)) return -2;
    std::vector<float> ones(NFFT, 1.0f);
    FwdTables tb{ones.data(), h.tw1t.data(), h.mel_w.data(), h.mel_lo.data(), h.mel_roundw.data()};
    alignas(16) static float frames[WARP_SMEM_F];
    memset(frames, 0, sizeof(frames));
    // feed the frame as "speech = re, noise = im" of a signal whose frame 2 is interior
    const int L = 4 * NFFT;
    std::vector<float> s(L, 0.0f), n(L, 0.0f);
    const int t = 4;  // frame 4 covers original samples [4*160-320, 4*160+320) = [320, 960)
    for (int i = 0; i < NFFT; ++i) { s[t * HOP - HALF + i] = re[i]; n[t * HOP - HALF + i] = im[i]; }
    FwdTile tl{};
    tl.sp = s.data(); tl.nz = n.data(); tl.L = L; tl.valid_s = L; tl.valid_n = L; tl.T = 1 + L / HOP; tl.t0 = 4;
    tl.factor = 0.0f; tl.mixed_pcm = nullptr;
    for (int lane = 0; lane < 32; ++lane) stage_pass1<true>(tb, tl, lane, frames);
    static float yr[32][40], yi[32][40];
    for (int j = 0; j < 2; ++j) {
        for (int lane = 0; lane < 32; ++lane) pass2_compute(lane, j, frames, yr[lane], yi[lane]);
        for (int lane = 0; lane < 32; ++lane) pass2_store(lane, j, frames, yr[lane], yi[lane]);
    }
    for (int k = 0; k < NFFT; ++k) { out_re[k] = frames[2 * k]; out_im[k] = frames[2 * k + 1]; }
    return 0;
}

extern "C" int emul_forward(const float* speech, const float* noise, int L, int valid_s, int valid_n, float factor,
                            int layout, int n_slices, int ld_t, float* out_sp, float* out_nz, float* out_mix,
                            float* mixed_pcm, float* max3, int sample_rate, double fmin, double fmax) {
    HostTables h;
    if (!build_tables(h, sample_rate, fmin, fmax)) return -2;
    FwdTables tb{h.window.data(), h.tw1t.data(), h.mel_w.data(), h.mel_lo.data(), h.mel_roundw.data()};
    alignas(16) static float smem[WARP_SMEM_F];
    memset(smem, 0, sizeof(smem));
    float* frames = smem;
    float* melst = smem + FPG * FRAME_F;
    const int T = 1 + L / HOP, G = (T + FPG - 1) / FPG;
    FwdTile tl{};
    tl.sp = speech; tl.nz = noise; tl.L = L;
    tl.valid_s = valid_s < L ? valid_s : L;
    tl.valid_n = valid_n < L ? valid_n : L;
    tl.T = T; tl.factor = noise ? factor : 0.0f; tl.mixed_pcm = mixed_pcm;
    FwdOut out[3];
    float* dsts[3] = {out_sp, out_nz, out_mix};
    for (int s = 0; s < 3; ++s) { out[s].dst = dsts[s]; out[s].layout = layout; out[s].n_slices = n_slices; out[s].ld_t = ld_t; }
    float mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    static float yr[32][40], yi[32][40];
    for (int g = 0; g < G; ++g) {
        tl.t0 = g * FPG;
        const int vmin = tl.valid_s < tl.valid_n ? tl.valid_s : tl.valid_n;
        const bool interior = (tl.nz != nullptr) && (tl.t0 * HOP - HALF >= 0) && ((tl.t0 + FPG - 1) * HOP + HALF <= vmin) &&
                              (tl.t0 + FPG - 1 < tl.T);
        for (int lane = 0; lane < 32; ++lane) {
            if (interior) stage_pass1<false>(tb, tl, lane, frames);
            else stage_pass1<true>(tb, tl, lane, frames);
        }
        for (int j = 0; j < 2; ++j) {
            for (int lane = 0; lane < 32; ++lane) pass2_compute(lane, j, frames, yr[lane], yi[lane]);
            for (int lane = 0; lane < 32; ++lane) pass2_store(lane, j, frames, yr[lane], yi[lane]);
        }
        for (int lane = 0; lane < 32; ++lane) stage_post(lane, tl.factor, frames, nullptr);
        for (int r = 0; r < MEL_ROUNDS; ++r)
            for (int lane = 0; lane < 32; ++lane) stage_mel_round(lane, r, tb.mel_roundw[r], tb.mel_w, tb.mel_lo, frames, melst);
        for (int s = 0; s < 3; ++s) {
            if (s >= 1 && noise == nullptr) continue;
            const float scale = (s == 1) ? tl.factor : 1.0f;
            for (int q = 0; q < 3; ++q)
                for (int lane = 0; lane < 32; ++lane) {
                    const float v = stage_db(lane, q, scale, melst + s * NMEL * FPG, out[s], tl.t0, T);
                    if (v > mx[s]) mx[s] = v;
                }
        }
    }
    for (int s = 0; s < 3; ++s) max3[s] = key_to_float(float_to_key(mx[s]));
    return 0;
}
