// Test hooks for the generic-geometry host tables (csrc/avse_generic_tables.cpp): plain C++, compiled with g++ by
// tests/test_generic_tables.py.  Test infrastructure only.
#include <cstring>
#include "../../audio-visual-speech-enhancement_b200/csrc/avse_generic.h"

using namespace avse;

// geo_out: n_fft, hop, bins, n_mels, spss, n1, n2, n_inv, i1, i2.  fb_out [n_mels][bins] f64, pinv_out [bins][n_mels] f32,
// band_out [n_mels][2] (lo, cnt).  Returns 0 ok, 1 unsupported configuration.
extern "C" int emul_generic_tables(int sample_rate, int n_fft, int hop, int n_mels, int spss, double fmin, double fmax, int* geo_out,
                                   double* fb_out, float* pinv_out, int* band_out, float* window_out, double* tw_out) {
    GenericHost g;
    if (!build_generic(g, sample_rate, n_fft, hop, n_mels, spss, fmin, fmax)) return 1;
    const GenericGeo& q = g.geo;
    const int geo[10] = {q.n_fft, q.hop, q.bins, q.n_mels, q.spss, q.n1, q.n2, q.n_inv, q.i1, q.i2};
    std::memcpy(geo_out, geo, sizeof(geo));
    if (fb_out) std::memcpy(fb_out, g.fb.data(), g.fb.size() * sizeof(double));
    if (pinv_out) std::memcpy(pinv_out, g.pinv.data(), g.pinv.size() * sizeof(float));
    if (band_out)
        for (int m = 0; m < n_mels; ++m) { band_out[2 * m] = g.band_lo[m]; band_out[2 * m + 1] = g.band_cnt[m]; }
    if (window_out) std::memcpy(window_out, g.window.data(), g.window.size() * sizeof(float));
    if (tw_out) std::memcpy(tw_out, g.tw.data(), g.tw.size() * sizeof(double));
    return 0;
}
