/* Plain-C client of the C ABI (include/avse_b200.h): no Python, no torch -- device buffers from cudaMalloc.
 *
 *   preprocess_audio_pair (dp:119-139):  avse_snr_factor -> avse_forward -> avse_floor_inplace3
 *   reconstruct_speech_signal (dp:60-74): avse_inverse_work_elems_ctx -> avse_inverse
 *
 * Build (see tests/test_c_abi_demo.py):
 *   gcc examples/c_abi_demo.c -Iinclude -I/usr/local/cuda/include -Laudio-visual-speech-enhancement_b200 -lavse_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/audio-visual-speech-enhancement_b200 -o build/c_abi_demo
 * Prints one line of checksums that the test compares with the Python engine on the same inputs. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "avse_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_AVSE(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, avse_last_error()); return 3; } } while (0)

int main(void) {
    enum { B = 2, NVS = 5, L = 3200 * NVS, NS = NVS, ROW = AVSE_N_MELS * AVSE_SPSS };
    const size_t n_in = (size_t)B * L, n_out = (size_t)B * NS * ROW;
    float* h_s = (float*)malloc(n_in * sizeof(float));
    float* h_n = (float*)malloc(n_in * sizeof(float));
    unsigned state = 12345u;     /* deterministic inputs: two tones + LCG noise */
    for (int u = 0; u < B; ++u)
        for (int i = 0; i < L; ++i) {
            state = state * 1664525u + 1013904223u;
            const double r = ((state >> 8) & 0xffff) / 65536.0 - 0.5;
            h_s[(size_t)u * L + i] = (float)(0.3 * sin(2.0 * 3.14159265358979 * (220.0 + 110.0 * u) * i / 16000.0) * (0.5 + 0.5 * sin(i / 1500.0)));
            h_n[(size_t)u * L + i] = (float)(0.1 * r);
        }

    avse_ctx* ctx = NULL;
    CHECK_AVSE(avse_create(16000, 0.0, 8000.0, 0, &ctx));
    float *d_s, *d_n, *d_factor, *d_eq, *d_sp, *d_nz, *d_mx, *d_pcm, *d_rec, *d_work;
    int* d_keys;
    CHECK_CUDA(cudaMalloc((void**)&d_s, n_in * 4));
    CHECK_CUDA(cudaMalloc((void**)&d_n, n_in * 4));
    CHECK_CUDA(cudaMalloc((void**)&d_factor, B * 4));
    CHECK_CUDA(cudaMalloc((void**)&d_eq, B * 4));
    CHECK_CUDA(cudaMalloc((void**)&d_keys, 2 * B * 3 * 4));
    CHECK_CUDA(cudaMalloc((void**)&d_sp, n_out * 4));
    CHECK_CUDA(cudaMalloc((void**)&d_nz, n_out * 4));
    CHECK_CUDA(cudaMalloc((void**)&d_mx, n_out * 4));
    CHECK_CUDA(cudaMalloc((void**)&d_pcm, n_in * 4));
    CHECK_CUDA(cudaMemcpy(d_s, h_s, n_in * 4, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(d_n, h_n, n_in * 4, cudaMemcpyHostToDevice));
    int* d_max = d_keys;
    int* d_min = d_keys + B * 3;

    CHECK_AVSE(avse_snr_factor(ctx, d_s, d_n, AVSE_SAMPLE_F32, L, NULL, NULL, B, L, NULL, d_factor, d_eq, d_max, d_min, NULL));
    avse_forward_args fa = {0};
    fa.speech = d_s; fa.noise = d_n; fa.in_stride = L; fa.factor = d_factor; fa.equalizer = d_eq;
    fa.B = B; fa.L = L; fa.layout = AVSE_LAYOUT_SLICES; fa.n_slices = NS;
    fa.out_speech = d_sp; fa.out_noise = d_nz; fa.out_mixed = d_mx; fa.out_stride = (long long)NS * ROW;
    fa.mixed_pcm = d_pcm; fa.pcm_stride = L; fa.max_key = d_max; fa.min_key = d_min; fa.sample_format = AVSE_SAMPLE_F32;
    CHECK_AVSE(avse_forward(ctx, &fa, NULL));
    CHECK_AVSE(avse_floor_inplace3(ctx, d_sp, d_nz, d_mx, (long long)NS * ROW, (long long)NS * ROW, B, d_max, d_min, NULL));

    const int t_use = AVSE_SPSS * NS;                   /* min(20 n, T) with T = 1 + L / 160 = 101 */
    const int out_len = AVSE_HOP * (t_use - 1);
    long long per = 0;
    CHECK_AVSE(avse_inverse_work_elems_ctx(ctx, t_use, &per));
    CHECK_CUDA(cudaMalloc((void**)&d_work, (size_t)B * per * 4));
    CHECK_CUDA(cudaMalloc((void**)&d_rec, (size_t)B * out_len * 4));
    avse_inverse_args ia = {0};
    ia.mel_db = d_sp; ia.layout = AVSE_LAYOUT_SLICES; ia.n_slices = NS; ia.mel_stride = (long long)NS * ROW;
    ia.mixed_pcm = d_pcm; ia.pcm_stride = L; ia.B = B; ia.L = L;
    ia.out_pcm = d_rec; ia.out_stride = out_len; ia.work = d_work; ia.work_stride = per; ia.out_format = AVSE_SAMPLE_F32;
    CHECK_AVSE(avse_inverse(ctx, &ia, NULL));
    CHECK_CUDA(cudaDeviceSynchronize());

    float* h_mx = (float*)malloc(n_out * 4);
    float* h_rec = (float*)malloc((size_t)B * out_len * 4);
    float h_factor[B];
    CHECK_CUDA(cudaMemcpy(h_mx, d_mx, n_out * 4, cudaMemcpyDeviceToHost));
    CHECK_CUDA(cudaMemcpy(h_rec, d_rec, (size_t)B * out_len * 4, cudaMemcpyDeviceToHost));
    CHECK_CUDA(cudaMemcpy(h_factor, d_factor, B * 4, cudaMemcpyDeviceToHost));
    double sum_mx = 0.0, sum_rec = 0.0;
    for (size_t i = 0; i < n_out; ++i) sum_mx += h_mx[i];
    for (size_t i = 0; i < (size_t)B * out_len; ++i) sum_rec += fabs((double)h_rec[i]);
    printf("avse_c_abi_demo %s factor0=%.7g factor1=%.7g sum_mixed_db=%.6f sum_abs_recon=%.6f\n", avse_version(), h_factor[0], h_factor[1], sum_mx, sum_rec);

    /* error behaviour: bad arguments return a negative status and a message, nothing throws */
    fa.n_slices = 99;
    if (avse_forward(ctx, &fa, NULL) != AVSE_E_ARG) { fprintf(stderr, "expected AVSE_E_ARG\n"); return 4; }
    avse_destroy(ctx);
    return 0;
}
