// Register-resident small DFT codelets (forward transform, e^{-2 pi i nk/N}) used to build the
// 640-point FFT = DFT-16 x twiddle x DFT-40 (40 = 5 x 8 Good-Thomas, twiddle-free).
// All loops are compile-time unrolled; arrays live in registers.  Compiles for host (g++)
// as well, so the same arithmetic is exercised by the CPU-side emulation tests.
#pragma once
#include "avse_common.h"

namespace avse {

// --- radix-4 butterfly, in place, forward -------------------------------------------------
AVSE_HD void dft4(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3) {
    const float t0r = r0 + r2, t0i = i0 + i2;
    const float t1r = r0 - r2, t1i = i0 - i2;
    const float t2r = r1 + r3, t2i = i1 + i3;
    const float t3r = r1 - r3, t3i = i1 - i3;
    r0 = t0r + t2r; i0 = t0i + t2i;
    r2 = t0r - t2r; i2 = t0i - t2i;
    r1 = t1r + t3i; i1 = t1i - t3r;   // t1 - i*t3
    r3 = t1r - t3i; i3 = t1i + t3r;   // t1 + i*t3
}

// --- DFT-5 (in place) ---------------------------------------------------------------------
AVSE_HD void dft5(float (&r)[5], float (&i)[5]) {
    constexpr float C1 = 0.30901699437494745f;   // cos(2pi/5)
    constexpr float C2 = -0.80901699437494745f;  // cos(4pi/5)
    constexpr float S1 = 0.95105651629515353f;   // sin(2pi/5)
    constexpr float S2 = 0.58778525229247314f;   // sin(4pi/5)
    const float a1r = r[1] + r[4], a1i = i[1] + i[4];
    const float a2r = r[2] + r[3], a2i = i[2] + i[3];
    const float b1r = r[1] - r[4], b1i = i[1] - i[4];
    const float b2r = r[2] - r[3], b2i = i[2] - i[3];
    const float x0r = r[0], x0i = i[0];
    r[0] = x0r + a1r + a2r;
    i[0] = x0i + a1i + a2i;
    const float p1r = x0r + C1 * a1r + C2 * a2r, p1i = x0i + C1 * a1i + C2 * a2i;
    const float p2r = x0r + C2 * a1r + C1 * a2r, p2i = x0i + C2 * a1i + C1 * a2i;
    const float q1r = S1 * b1r + S2 * b2r, q1i = S1 * b1i + S2 * b2i;
    const float q2r = S2 * b1r - S1 * b2r, q2i = S2 * b1i - S1 * b2i;
    r[1] = p1r + q1i; i[1] = p1i - q1r;   // p1 - i*q1
    r[4] = p1r - q1i; i[4] = p1i + q1r;   // p1 + i*q1
    r[2] = p2r + q2i; i[2] = p2i - q2r;
    r[3] = p2r - q2i; i[3] = p2i + q2r;
}

// --- DFT-8 (in place, natural order out) ---------------------------------------------------
AVSE_HD void dft8(float (&r)[8], float (&i)[8]) {
    constexpr float H = 0.70710678118654752f;
    // even / odd DFT-4
    float e0r = r[0], e0i = i[0], e1r = r[2], e1i = i[2], e2r = r[4], e2i = i[4], e3r = r[6], e3i = i[6];
    float o0r = r[1], o0i = i[1], o1r = r[3], o1i = i[3], o2r = r[5], o2i = i[5], o3r = r[7], o3i = i[7];
    dft4(e0r, e0i, e1r, e1i, e2r, e2i, e3r, e3i);
    dft4(o0r, o0i, o1r, o1i, o2r, o2i, o3r, o3i);
    // twiddles W8^k on the odd half
    const float w1r = (o1r + o1i) * H, w1i = (o1i - o1r) * H;     // * (1 - i)/sqrt2
    const float w2r = o2i, w2i = -o2r;                              // * (-i)
    const float w3r = (o3i - o3r) * H, w3i = -(o3r + o3i) * H;    // * (-1 - i)/sqrt2
    r[0] = e0r + o0r; i[0] = e0i + o0i; r[4] = e0r - o0r; i[4] = e0i - o0i;
    r[1] = e1r + w1r; i[1] = e1i + w1i; r[5] = e1r - w1r; i[5] = e1i - w1i;
    r[2] = e2r + w2r; i[2] = e2i + w2i; r[6] = e2r - w2r; i[6] = e2i - w2i;
    r[3] = e3r + w3r; i[3] = e3i + w3i; r[7] = e3r - w3r; i[7] = e3i - w3i;
}

// --- DFT-16 (in place, natural order out): radix-4 x radix-4 -------------------------------
AVSE_HD void cmul_const(float& r, float& i, const float c, const float s) {
    // (r + i i) * (c - i s)  where the twiddle is e^{-i theta} = c - i s
    const float tr = r * c + i * s;
    const float ti = i * c - r * s;
    r = tr; i = ti;
}

AVSE_HD void dft16(float (&r)[16], float (&i)[16]) {
    // cos/sin(2 pi m / 16)
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f;
    constexpr float C2 = 0.70710678118654752f, S2 = 0.70710678118654752f;
    constexpr float C3 = 0.38268343236508977f, S3 = 0.92387953251128674f;
    // stage 1: for each b, DFT-4 over a of x[4a + b]  -> y[b][k'] stored back at x[4k' + b]
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(r[b], i[b], r[4 + b], i[4 + b], r[8 + b], i[8 + b], r[12 + b], i[12 + b]);
    // twiddle y[b][k'] *= W16^{b k'}  (element at index 4k' + b)
    cmul_const(r[4 + 1], i[4 + 1], C1, S1);          // b=1,k'=1 : m=1
    cmul_const(r[8 + 1], i[8 + 1], C2, S2);          // b=1,k'=2 : m=2
    cmul_const(r[12 + 1], i[12 + 1], C3, S3);        // b=1,k'=3 : m=3
    cmul_const(r[4 + 2], i[4 + 2], C2, S2);          // b=2,k'=1 : m=2
    { const float t = r[8 + 2]; r[8 + 2] = i[8 + 2]; i[8 + 2] = -t; }   // b=2,k'=2 : m=4 -> * (-i)
    cmul_const(r[12 + 2], i[12 + 2], -C2, S2);       // b=2,k'=3 : m=6 -> cos=-C2, sin=S2
    cmul_const(r[4 + 3], i[4 + 3], C3, S3);          // b=3,k'=1 : m=3
    cmul_const(r[8 + 3], i[8 + 3], -C2, S2);         // b=3,k'=2 : m=6
    cmul_const(r[12 + 3], i[12 + 3], -C1, -S1);      // b=3,k'=3 : m=9 -> cos=-C1, sin=-S1
    // stage 2: for each k', DFT-4 over b of y[b][k'] (indices 4k' + b) -> X[k' + 4k''] at index 4k' + k''
#pragma unroll
    for (int k = 0; k < 4; ++k) dft4(r[4 * k], i[4 * k], r[4 * k + 1], i[4 * k + 1], r[4 * k + 2], i[4 * k + 2], r[4 * k + 3], i[4 * k + 3]);
    // now index 4k' + k'' holds X[k' + 4k'']: transpose the 4x4 to natural order
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            float t = r[4 * a + b]; r[4 * a + b] = r[4 * b + a]; r[4 * b + a] = t;
            t = i[4 * a + b]; i[4 * a + b] = i[4 * b + a]; i[4 * b + a] = t;
        }
}

// --- DFT-40 = 5 x 8 prime-factor (Good-Thomas), out of place --------------------------------
// in index n = (8a + 5b) mod 40, out index k = (16c + 25d) mod 40; no twiddles.
AVSE_HD void dft40(const float (&xr)[40], const float (&xi)[40], float (&yr)[40], float (&yi)[40]) {
    float ur[8][5], ui[8][5];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        float tr[5], ti[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) { tr[a] = xr[(8 * a + 5 * b) % 40]; ti[a] = xi[(8 * a + 5 * b) % 40]; }
        dft5(tr, ti);
#pragma unroll
        for (int c = 0; c < 5; ++c) { ur[b][c] = tr[c]; ui[b][c] = ti[c]; }
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        float tr[8], ti[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) { tr[b] = ur[b][c]; ti[b] = ui[b][c]; }
        dft8(tr, ti);
#pragma unroll
        for (int d = 0; d < 8; ++d) { yr[(16 * c + 25 * d) % 40] = tr[d]; yi[(16 * c + 25 * d) % 40] = ti[d]; }
    }
}

// In-place variant: element (8a + 5b) % 40 holds input n = (8a + 5b) % 40 on entry and, on exit,
// element (8c + 5d) % 40 holds output k = (16c + 25d) % 40.  80 live registers instead of 160.
AVSE_HD void dft40_inplace(float (&xr)[40], float (&xi)[40]) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        float tr[5], ti[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) { tr[a] = xr[(8 * a + 5 * b) % 40]; ti[a] = xi[(8 * a + 5 * b) % 40]; }
        dft5(tr, ti);
#pragma unroll
        for (int c = 0; c < 5; ++c) { xr[(8 * c + 5 * b) % 40] = tr[c]; xi[(8 * c + 5 * b) % 40] = ti[c]; }
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        float tr[8], ti[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) { tr[b] = xr[(8 * c + 5 * b) % 40]; ti[b] = xi[(8 * c + 5 * b) % 40]; }
        dft8(tr, ti);
#pragma unroll
        for (int d = 0; d < 8; ++d) { xr[(8 * c + 5 * d) % 40] = tr[d]; xi[(8 * c + 5 * d) % 40] = ti[d]; }
    }
}

}  // namespace avse
