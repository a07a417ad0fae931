// Shared constants and host/device qualifiers for the B200 spectral front/back end.
//
// The path mirrors /root/reference/data_processor.py:35-139 at the reference's hard-coded
// geometry (dp:44-45, dp:83-89): 16 kHz / 25 fps -> n_fft = 640 (= 2^7 * 5), hop = 160,
// 321 bins, 80 Slaney mel bands, 200 ms slices = 20 spectrogram frames.
#pragma once

#if defined(__CUDACC__)
#define AVSE_HD __host__ __device__ __forceinline__
#define AVSE_HD_COLD __host__ __device__ __noinline__     // cold helpers: ONE copy, kept out of the hot instruction stream
#else
#define AVSE_HD inline
#define AVSE_HD_COLD inline
#endif

namespace avse {

constexpr int NFFT = 640;            // dp:44   int(16000 / 25)
constexpr int HOP = 160;             // dp:45   int(n_fft / 4)
constexpr int NBINS = NFFT / 2 + 1;  // 321
constexpr int NMEL = 80;             // dp:86
constexpr int SPSS = 20;             // dp:49   spectrogram frames per 200 ms slice
constexpr int HALF = NFFT / 2;       // reflect-pad width of librosa.stft(center=True)

// FFT-640 factorisation: n = 40*n1 + n2, k = k1 + 16*k2  (DFT-16, twiddle, DFT-40 = 5 x 8 PFA)
constexpr int N1 = 16;
constexpr int N2 = 40;

// One warp owns a group of 2 consecutive STFT frames (10.9 KB of shared memory per warp -> 16 warps / SM).
constexpr int FPG = 2;
// Shared-memory frame buffer: 16 rows (k1) x 84 floats (40 complex + 2 pad) + 16 floats skew.
constexpr int ROW_F = 84;
// Layout of one frame region (floats): [0, 1344) rows / natural-order Z (Z uses [0, 1282)),
// [1344, 1360) never written (kept zero: "null" partial sums), [1360, 1456) chunk-end partial sums
// of the fused post+mel scan (16 chunks x 6 floats).
constexpr int FRAME_ZERO_F = N1 * ROW_F;        // 1344
constexpr int FRAME_FLUSH_F = N1 * ROW_F + 16;  // 1360
constexpr int FRAME_F = FRAME_FLUSH_F + 96;     // 1456 floats, == 16 (mod 32), multiple of 4
constexpr int SCAN_BINS = 336;                  // 16 chunks x POST_CHUNK bins (bins >= 320 unused)
constexpr int MEL_STAGE_F = 3 * NMEL * FPG;  // raw mel sums [sig][band][frame], staged over frame buffer 0
constexpr int WARP_SMEM_F = FPG * FRAME_F;   // 2720 floats = 10880 B per warp
static_assert(MEL_STAGE_F <= FRAME_F, "mel staging must fit in one frame buffer");

// mel tables
constexpr int MEL_WMAX = 24;    // max bins per band supported (reference config: 23)
constexpr int MEL_WROW = 28;    // padded row stride of the weight table (16-byte rows, bank spread)
constexpr int MEL_ROUNDS = NMEL / 16;   // one round = 16 consecutive bands x 2 frames
constexpr int POST_CHUNK = 21;  // bins per lane in the pointwise post stage (16 lanes x 21 >= 321; odd: bank spread)

constexpr float AMIN = 1e-5f;   // librosa.amplitude_to_db amin (dp:94)
constexpr float TOP_DB = 80.0f; // librosa.amplitude_to_db top_db (dp:94)

}  // namespace avse
