// See avse_generic.h.  Float64 throughout; rounded to float32 once at the end.  Plain C++, no CUDA.
#include "avse_generic.h"

#include <algorithm>
#include <cmath>

namespace avse {

namespace {

const double kPi = 3.14159265358979323846;

// librosa.hz_to_mel / mel_to_hz, htk=False (Slaney): linear below 1 kHz, log above.
double g_hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0;
    const double min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double g_mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0;
    const double min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

void factor_pair(int n, int& a, int& b) {
    a = 1;
    for (int d = 1; (long long)d * d <= n; ++d)
        if (n % d == 0) a = d;
    b = n / a;
}

void hann_and_twiddles(int n, std::vector<float>& win, std::vector<double>& tw) {
    win.resize(n);
    tw.resize(2 * (size_t)n);
    for (int j = 0; j < n; ++j) {
        win[j] = (float)(0.5 - 0.5 * std::cos(2.0 * kPi * j / n));      // scipy get_window('hann', n, fftbins=True)
        const double a = -2.0 * kPi * j / n;
        tw[2 * j] = std::cos(a);
        tw[2 * j + 1] = std::sin(a);
    }
}

}  // namespace

void mel_filterbank_dense(int sample_rate, int n_fft, int n_mels, double fmin, double fmax, std::vector<double>& fb) {
    const int bins = 1 + n_fft / 2;
    std::vector<double> mel_f(n_mels + 2);
    const double m0 = g_hz_to_mel(fmin), m1 = g_hz_to_mel(fmax);
    for (int j = 0; j < n_mels + 2; ++j) mel_f[j] = g_mel_to_hz(m0 + (m1 - m0) * j / (n_mels + 1));
    fb.assign((size_t)n_mels * bins, 0.0);
    for (int m = 0; m < n_mels; ++m) {
        const double fd0 = mel_f[m + 1] - mel_f[m], fd1 = mel_f[m + 2] - mel_f[m + 1];
        const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
        double rmax = 0.0;
        for (int k = 0; k < bins; ++k) {
            const double fk = bins > 1 ? (double)sample_rate / 2.0 * k / (bins - 1) : 0.0;   // np.linspace(0, sr/2, bins)
            const double lower = -(mel_f[m] - fk) / fd0;
            const double upper = (mel_f[m + 2] - fk) / fd1;
            const double w = std::max(0.0, std::min(lower, upper)) * enorm;
            fb[(size_t)m * bins + k] = w;
            rmax = std::max(rmax, w);
        }
        // a bin sitting exactly on a mel point has weight 0 in exact arithmetic; libm rounding can leave ~1e-16
        for (int k = 0; k < bins; ++k)
            if (fb[(size_t)m * bins + k] < 1e-10 * rmax) fb[(size_t)m * bins + k] = 0.0;
    }
}

void pinv_dense(const std::vector<double>& F, int rows, int cols, std::vector<double>& P) {
    // One-sided (Hestenes) Jacobi on the rows: G = R F with orthogonal rows g_i = sigma_i v_i^T, F = R^T Sigma V^T,
    // pinv(F) = V Sigma^+ R  ->  P[k][m] = sum_i g_i[k] / sigma_i^2 * R[i][m] over sigma_i > rcond * sigma_max.
    std::vector<double> G(F), R((size_t)rows * rows, 0.0);
    for (int i = 0; i < rows; ++i) R[(size_t)i * rows + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < rows - 1; ++p)
            for (int q = p + 1; q < rows; ++q) {
                double al = 0.0, be = 0.0, ga = 0.0;
                const double* gp = &G[(size_t)p * cols];
                const double* gq = &G[(size_t)q * cols];
                for (int k = 0; k < cols; ++k) { al += gp[k] * gp[k]; be += gq[k] * gq[k]; ga += gp[k] * gq[k]; }
                if (ga == 0.0 || std::fabs(ga) <= 1e-17 * std::sqrt(al * be)) continue;
                off = std::max(off, std::fabs(ga) / std::sqrt(al * be));
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                double* wp = &G[(size_t)p * cols];
                double* wq = &G[(size_t)q * cols];
                for (int k = 0; k < cols; ++k) { const double x = wp[k], y = wq[k]; wp[k] = c * x - s * y; wq[k] = s * x + c * y; }
                double* rp = &R[(size_t)p * rows];
                double* rq = &R[(size_t)q * rows];
                for (int k = 0; k < rows; ++k) { const double x = rp[k], y = rq[k]; rp[k] = c * x - s * y; rq[k] = s * x + c * y; }
            }
        if (off < 1e-15) break;
    }
    std::vector<double> s2(rows, 0.0);
    double smax2 = 0.0;
    for (int i = 0; i < rows; ++i) {
        for (int k = 0; k < cols; ++k) s2[i] += G[(size_t)i * cols + k] * G[(size_t)i * cols + k];
        smax2 = std::max(smax2, s2[i]);
    }
    P.assign((size_t)cols * rows, 0.0);
    const double cut2 = 1e-30 * smax2;     // (rcond = 1e-15)^2, np.linalg.pinv default
    for (int i = 0; i < rows; ++i) {
        if (!(s2[i] > cut2)) continue;
        const double inv = 1.0 / s2[i];
        for (int k = 0; k < cols; ++k) {
            const double gk = G[(size_t)i * cols + k] * inv;
            if (gk == 0.0) continue;
            for (int m = 0; m < rows; ++m) P[(size_t)k * rows + m] += gk * R[(size_t)i * rows + m];
        }
    }
}

bool build_generic(GenericHost& g, int sample_rate, int n_fft, int hop, int n_mels, int spss, double fmin, double fmax) {
    g.error.clear();
    if (sample_rate <= 0 || n_fft < 4 || n_fft > 4096) { g.error = "n_fft must be in [4, 4096]"; return false; }
    if (hop < 1 || hop > n_fft) { g.error = "hop must be in [1, n_fft]"; return false; }
    if (n_mels < 1 || n_mels > 256) { g.error = "n_mels must be in [1, 256]"; return false; }
    if (spss < 1) { g.error = "spss (spectrogram frames per slice) must be >= 1"; return false; }
    if (!(fmin >= 0.0) || !(fmax > fmin)) { g.error = "need 0 <= fmin < fmax"; return false; }
    GenericGeo& q = g.geo;
    q.n_fft = n_fft;
    q.hop = hop;
    q.bins = 1 + n_fft / 2;
    q.n_mels = n_mels;
    q.spss = spss;
    factor_pair(n_fft, q.n1, q.n2);
    q.n_inv = 2 * (q.bins - 1);
    factor_pair(q.n_inv, q.i1, q.i2);
    hann_and_twiddles(n_fft, g.window, g.tw);
    hann_and_twiddles(q.n_inv, g.window_inv, g.tw_inv);

    mel_filterbank_dense(sample_rate, n_fft, n_mels, fmin, fmax, g.fb);
    g.band_lo.assign(n_mels, 0);
    g.band_cnt.assign(n_mels, 0);
    g.band_off.assign(n_mels, 0);
    g.band_w.clear();
    for (int m = 0; m < n_mels; ++m) {
        int lo = -1, hi = -1;
        for (int k = 0; k < q.bins; ++k)
            if (g.fb[(size_t)m * q.bins + k] != 0.0) { if (lo < 0) lo = k; hi = k; }
        g.band_off[m] = (int)g.band_w.size();
        if (lo < 0) continue;      // empty filter (librosa only warns): the band's mel value is 0 -> -100 dB
        g.band_lo[m] = lo;
        g.band_cnt[m] = hi - lo + 1;
        for (int k = lo; k <= hi; ++k) g.band_w.push_back((float)g.fb[(size_t)m * q.bins + k]);
    }
    if (g.band_w.empty()) g.band_w.push_back(0.0f);

    std::vector<double> P;
    pinv_dense(g.fb, n_mels, q.bins, P);
    g.pinv.resize(P.size());
    for (size_t i = 0; i < P.size(); ++i) g.pinv[i] = (float)P[i];
    return true;
}

}  // namespace avse
