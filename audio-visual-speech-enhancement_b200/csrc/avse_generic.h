// Generic-geometry path: any n_fft the reference derives from (sample rate, video frame rate) (dp:44-45, dp:61-62),
// e.g. 320 @ 50 fps, 666 @ 24 fps, 533 @ 30 fps, 1764 @ 44.1 kHz / 25 fps.  The specialised 640 / 160 kernels stay the
// hot path; these kernels are the correct-but-slower fallback behind the same C ABI.
//
// Forward transform: two-level Cooley-Tukey with run-time factors n_fft = n1 * n2 (n1 <= n2 the divisor pair closest
// to sqrt(n_fft); prime sizes degenerate to the plain DFT), cost n_fft (n1 + n2) complex MACs per frame, twiddles from
// one table W[j] = exp(-2 pi i j / n_fft) built in float64.  Host tables restate librosa.filters.mel / get_window /
// np.linalg.pinv as used at /root/reference/data_processor.py:79-89, :104-112.
#pragma once
#include <string>
#include <vector>

namespace avse {

struct GenericGeo {
    int n_fft = 0;     // forward STFT size (dp:44)
    int hop = 0;       // dp:45
    int bins = 0;      // 1 + n_fft / 2
    int n_mels = 0;    // dp:86
    int spss = 0;      // spectrogram frames per slice (dp:49)
    int n1 = 1, n2 = 1;   // n_fft = n1 * n2
    int n_inv = 0;     // librosa.istft's inferred size 2 * (bins - 1) (== n_fft when n_fft is even)
    int i1 = 1, i2 = 1;   // n_inv = i1 * i2
};

struct GenericHost {
    GenericGeo geo;
    std::vector<float> window;       // [n_fft]  periodic Hann
    std::vector<double> tw;          // [n_fft][2]  exp(-2 pi i j / n_fft), float64
    std::vector<float> window_inv;   // [n_inv]
    std::vector<double> tw_inv;      // [n_inv][2]
    std::vector<double> fb;          // [n_mels][bins] dense filterbank (float64)
    std::vector<int> band_lo;        // [n_mels] first non-zero bin
    std::vector<int> band_cnt;       // [n_mels] bins from the first to the last non-zero (0 for an empty band)
    std::vector<int> band_off;       // [n_mels] offset of the band's weights in band_w
    std::vector<float> band_w;       // weights, band after band
    std::vector<float> pinv;         // [bins][n_mels]  np.linalg.pinv(fb) (one-sided Jacobi SVD in float64, rcond 1e-15)
    std::string error;
};

// Device copies of the tables (one allocation, owned by the context).  tw / tw_inv hold (re, im) pairs.
struct GenericDev {
    const float* window = nullptr;
    const double* tw = nullptr;
    const float* window_inv = nullptr;
    const double* tw_inv = nullptr;
    const int* band_lo = nullptr;
    const int* band_cnt = nullptr;
    const int* band_off = nullptr;
    const float* band_w = nullptr;
    const float* pinv = nullptr;
};

// librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) (htk=False, Slaney norm), dense float64 [n_mels][1 + n_fft/2].
void mel_filterbank_dense(int sample_rate, int n_fft, int n_mels, double fmin, double fmax, std::vector<double>& fb);

// np.linalg.pinv(F) for F [rows][cols] (rows <= cols), returned as [cols][rows].
void pinv_dense(const std::vector<double>& F, int rows, int cols, std::vector<double>& P);

bool build_generic(GenericHost& g, int sample_rate, int n_fft, int hop, int n_mels, int spss, double fmin, double fmax);

}  // namespace avse
