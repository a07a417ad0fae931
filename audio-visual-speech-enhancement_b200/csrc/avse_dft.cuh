// Register-resident small DFT codelets (forward transform, e^{-2 pi i nk/N}) used to build the
// 640-point FFT = DFT-16 x twiddle x DFT-40 (40 = 5 x 8 Good-Thomas, twiddle-free).
//
// Complex values are kept as (re, im) pairs in ONE 64-bit register pair so that the butterflies'
// complex additions map onto Blackwell's packed FP32 instructions (PTX add/sub/mul/fma.rn.f32x2 ->
// SASS FADD2 / FMUL2 / FFMA2, sm_100+): one instruction per complex add instead of two.  Measured on
// B200 (tools/ubench_f32x2.cu): FADD2 sustains 36.6 T lane-op/s against 28.0 for scalar FADD at the
// same occupancy, with half the issue slots.  The packed ops round each half exactly like the scalar
// ones (IEEE rn), so results are bit-identical to the scalar formulation.
//
// All loops are compile-time unrolled; arrays live in registers.  The same source compiles for the
// host (g++, plain float pairs) so the CPU-side emulation tests exercise the same arithmetic order.
#pragma once
#include "avse_common.h"

namespace avse {

#if defined(__CUDA_ARCH__)
struct cpx { unsigned long long v; };
AVSE_HD cpx cmake(float r, float i) { cpx c; asm("mov.b64 %0, {%1, %2};" : "=l"(c.v) : "f"(r), "f"(i)); return c; }
AVSE_HD float cre(cpx a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }   // register-pair halves: no instruction
AVSE_HD float cim(cpx a) { return __uint_as_float((unsigned)(a.v >> 32)); }
AVSE_HD cpx cadd(cpx a, cpx b) { cpx c; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c.v) : "l"(a.v), "l"(b.v)); return c; }
AVSE_HD cpx csub(cpx a, cpx b) { cpx c; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c.v) : "l"(a.v), "l"(b.v)); return c; }
// element-wise products with a (pr, pi) pair held in a register pair
AVSE_HD cpx cmul_pp(cpx a, cpx p) { cpx c; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c.v) : "l"(a.v), "l"(p.v)); return c; }
AVSE_HD cpx cfma_pp(cpx a, cpx p, cpx b) { cpx c; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c.v) : "l"(a.v), "l"(p.v), "l"(b.v)); return c; }
AVSE_HD cpx cload(const float* p) { cpx c; c.v = *reinterpret_cast<const unsigned long long*>(p); return c; }
AVSE_HD void cstore(float* p, cpx a) { *reinterpret_cast<unsigned long long*>(p) = a.v; }
AVSE_HD void cload2(const float* p, cpx& a, cpx& b) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    a.v = v.x; b.v = v.y;
}
AVSE_HD void cstore2(float* p, cpx a, cpx b) {
    ulonglong2 v; v.x = a.v; v.y = b.v;
    *reinterpret_cast<ulonglong2*>(p) = v;
}
#else
struct cpx { float r, i; };
AVSE_HD cpx cmake(float r, float i) { cpx c; c.r = r; c.i = i; return c; }
AVSE_HD float cre(cpx a) { return a.r; }
AVSE_HD float cim(cpx a) { return a.i; }
AVSE_HD cpx cadd(cpx a, cpx b) { return cmake(a.r + b.r, a.i + b.i); }
AVSE_HD cpx csub(cpx a, cpx b) { return cmake(a.r - b.r, a.i - b.i); }
AVSE_HD cpx cmul_pp(cpx a, cpx p) { return cmake(a.r * p.r, a.i * p.i); }
AVSE_HD cpx cfma_pp(cpx a, cpx p, cpx b) { return cmake(a.r * p.r + b.r, a.i * p.i + b.i); }
AVSE_HD cpx cload(const float* p) { return cmake(p[0], p[1]); }
AVSE_HD void cstore(float* p, cpx a) { p[0] = a.r; p[1] = a.i; }
AVSE_HD void cload2(const float* p, cpx& a, cpx& b) { a = cmake(p[0], p[1]); b = cmake(p[2], p[3]); }
AVSE_HD void cstore2(float* p, cpx a, cpx b) { p[0] = a.r; p[1] = a.i; p[2] = b.r; p[3] = b.i; }
#endif

// scalar broadcast forms: a * s and a * s + b on both halves
AVSE_HD cpx cmul_s(cpx a, float s) { return cmul_pp(a, cmake(s, s)); }
AVSE_HD cpx cfma_s(cpx a, float s, cpx b) { return cfma_pp(a, cmake(s, s), b); }
// -i * a = (im, -re)   and   a * (tr + i ti)  (scalar: the halves are crossed)
AVSE_HD cpx cmul_negj(cpx a) { return cmake(cim(a), -cre(a)); }
AVSE_HD cpx cmul(cpx a, float tr, float ti) {
    const float ar = cre(a), ai = cim(a);
    return cmake(ar * tr - ai * ti, ar * ti + ai * tr);
}

// --- radix-4 butterfly, in place, forward: 7 packed + 2 scalar adds ----------------------------
AVSE_HD void dft4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
    const cpx t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3);
    const cpx t3j = cmake(cim(x1) - cim(x3), cre(x3) - cre(x1));   // -i (x1 - x3)
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = cadd(t1, t3j);   // t1 - i t3
    x3 = csub(t1, t3j);   // t1 + i t3
}

// --- DFT-5 (in place) ---------------------------------------------------------------------------
AVSE_HD void dft5(cpx (&x)[5]) {
    constexpr float C1 = 0.30901699437494745f;   // cos(2pi/5)
    constexpr float C2 = -0.80901699437494745f;  // cos(4pi/5)
    constexpr float S1 = 0.95105651629515353f;   // sin(2pi/5)
    constexpr float S2 = 0.58778525229247314f;   // sin(4pi/5)
    const cpx a1 = cadd(x[1], x[4]), a2 = cadd(x[2], x[3]);
    const cpx b1 = csub(x[1], x[4]), b2 = csub(x[2], x[3]);
    const cpx x0 = x[0];
    x[0] = cadd(cadd(x0, a1), a2);
    const cpx p1 = cfma_s(a2, C2, cfma_s(a1, C1, x0));
    const cpx p2 = cfma_s(a2, C1, cfma_s(a1, C2, x0));
    const cpx q1 = cfma_s(b2, S2, cmul_s(b1, S1));
    const cpx q2 = cfma_s(b2, -S1, cmul_s(b1, S2));
    const cpx q1j = cmul_negj(q1), q2j = cmul_negj(q2);   // -i q
    x[1] = cadd(p1, q1j);   // p1 - i q1
    x[4] = csub(p1, q1j);   // p1 + i q1
    x[2] = cadd(p2, q2j);
    x[3] = csub(p2, q2j);
}

// --- DFT-8 (in place, natural order out) ---------------------------------------------------------
AVSE_HD void dft8(cpx (&x)[8]) {
    constexpr float H = 0.70710678118654752f;
    cpx e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    cpx o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    // twiddles W8^k on the odd half
    const cpx w1 = cmul_pp(cmake(cre(o1) + cim(o1), cim(o1) - cre(o1)), cmake(H, H));     // * (1 - i)/sqrt2
    const cpx w2 = cmul_negj(o2);                                                          // * (-i)
    const cpx w3 = cmul_pp(cmake(cim(o3) - cre(o3), cre(o3) + cim(o3)), cmake(H, -H));    // * (-1 - i)/sqrt2
    x[0] = cadd(e0, o0); x[4] = csub(e0, o0);
    x[1] = cadd(e1, w1); x[5] = csub(e1, w1);
    x[2] = cadd(e2, w2); x[6] = csub(e2, w2);
    x[3] = cadd(e3, w3); x[7] = csub(e3, w3);
}

// --- DFT-16 (in place, natural order out): radix-4 x radix-4 ------------------------------------
// twiddle e^{-i theta} = c - i s
AVSE_HD cpx cmul_const(cpx a, const float c, const float s) { return cmul(a, c, -s); }

AVSE_HD void dft16(cpx (&x)[16]) {
    // cos/sin(2 pi m / 16)
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f;
    constexpr float C2 = 0.70710678118654752f, S2 = 0.70710678118654752f;
    constexpr float C3 = 0.38268343236508977f, S3 = 0.92387953251128674f;
    // stage 1: for each b, DFT-4 over a of x[4a + b]  -> y[b][k'] stored back at x[4k' + b]
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(x[b], x[4 + b], x[8 + b], x[12 + b]);
    // twiddle y[b][k'] *= W16^{b k'}  (element at index 4k' + b)
    x[4 + 1] = cmul_const(x[4 + 1], C1, S1);          // b=1,k'=1 : m=1
    x[8 + 1] = cmul_const(x[8 + 1], C2, S2);          // b=1,k'=2 : m=2
    x[12 + 1] = cmul_const(x[12 + 1], C3, S3);        // b=1,k'=3 : m=3
    x[4 + 2] = cmul_const(x[4 + 2], C2, S2);          // b=2,k'=1 : m=2
    x[8 + 2] = cmul_negj(x[8 + 2]);                   // b=2,k'=2 : m=4 -> * (-i)
    x[12 + 2] = cmul_const(x[12 + 2], -C2, S2);       // b=2,k'=3 : m=6 -> cos=-C2, sin=S2
    x[4 + 3] = cmul_const(x[4 + 3], C3, S3);          // b=3,k'=1 : m=3
    x[8 + 3] = cmul_const(x[8 + 3], -C2, S2);         // b=3,k'=2 : m=6
    x[12 + 3] = cmul_const(x[12 + 3], -C1, -S1);      // b=3,k'=3 : m=9 -> cos=-C1, sin=-S1
    // stage 2: for each k', DFT-4 over b of y[b][k'] (indices 4k' + b) -> X[k' + 4k''] at index 4k' + k''
#pragma unroll
    for (int k = 0; k < 4; ++k) dft4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    // now index 4k' + k'' holds X[k' + 4k'']: transpose the 4x4 to natural order
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            const cpx t = x[4 * a + b]; x[4 * a + b] = x[4 * b + a]; x[4 * b + a] = t;
        }
}

// --- DFT-40 = 5 x 8 prime-factor (Good-Thomas), in place, no twiddles ----------------------------
// Element (8a + 5b) % 40 holds input n = (8a + 5b) % 40 on entry and, on exit, element (8c + 5d) % 40
// holds output k = (16c + 25d) % 40.
AVSE_HD void dft40_inplace(cpx (&x)[40]) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        cpx t[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) t[a] = x[(8 * a + 5 * b) % 40];
        dft5(t);
#pragma unroll
        for (int c = 0; c < 5; ++c) x[(8 * c + 5 * b) % 40] = t[c];
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        cpx t[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) t[b] = x[(8 * c + 5 * b) % 40];
        dft8(t);
#pragma unroll
        for (int d = 0; d < 8; ++d) x[(8 * c + 5 * d) % 40] = t[d];
    }
}

}  // namespace avse
