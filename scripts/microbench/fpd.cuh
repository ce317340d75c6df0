// FpD: the BLS12-381 base field on the FP64 pipe.
//
// B200 has a full-rate FP64 pipe (64 DFMA lanes per SM per clock) that the
// integer Montgomery path (field.cuh, IMAD.WIDE on the fma pipe) leaves idle.
// FpD runs the same Montgomery arithmetic (same R = 2^384, so values are
// interchangeable with `Fp` after a limb re-packing) on that pipe, so that
// FP64 warps and integer warps of one kernel add points side by side.
//
// Representation: x = sum v[i] * 2^(24 i), 16 limbs held as doubles, every
// limb an integer.  "Normalised" means |v[i]| <= 2^23 (balanced digits); the
// value itself is only kept in (-p, p), i.e. it is a redundant representative
// of its residue class.  All arithmetic is exact: a column of the 16 x 16
// product plus the 16 x 16 Montgomery correction is at most 32 products of
// two 24-bit integers, which is below 2^53, so chained DFMAs never round.
// Rounding to a multiple of 2^24 (carry extraction) is the usual magic-number
// addition in round-to-nearest.  Because everything is exact integer
// arithmetic, results converted back to canonical form are bit-identical to
// the integer path's (and to the reference's blst) by construction.
//
// Replaces, like field.cuh, the field layer the reference takes from
// nam-blstrs / nam-blst (reference Cargo.lock:1385-1411).
#pragma once
#include <cmath>

#include "../../masp_b200/csrc/field.cuh"

namespace mb {

#if defined(__CUDA_ARCH__)
MB_D double d_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
MB_D double d_add(double a, double b) { return __dadd_rn(a, b); }
MB_D double d_mul(double a, double b) { return __dmul_rn(a, b); }
MB_D double d_from_u24(uint32_t x) { return __dadd_rn(__hiloint2double(0x43300000, (int)x), -4503599627370496.0); }
MB_D uint32_t d_to_u32(double x) { return (uint32_t)__double2loint(__dadd_rn(x, 4503599627370496.0)); }
#else
inline double d_fma(double a, double b, double c) { return std::fma(a, b, c); }
inline double d_add(double a, double b) {
    volatile double r = a + b;
    return r;
}
inline double d_mul(double a, double b) { return a * b; }
inline double d_from_u24(uint32_t x) { return (double)x; }
inline uint32_t d_to_u32(double x) { return (uint32_t)(int64_t)x; }
#endif

struct FpD {
    static constexpr int N = 16;
    double v[N];

    static constexpr double M24 = 113336795588871485128704.0;     // 3 * 2^75: x + M24 has ulp 2^24
    static constexpr double TWO24 = 16777216.0;
    static constexpr double INV24 = 1.0 / 16777216.0;
    static constexpr double PINV = -196611.0;                      // -p^-1 mod 2^24, balanced

    MB_HD static constexpr double P(int i) {  // balanced 24-bit digits of p
        constexpr double t[16] = {-21845.0,   0.0,        -17921.0,   -5155840.0, -5505025.0, -646113.0,
                                  -6228303.0, 6762707.0,  -8056129.0, 4949236.0,  -2661257.0, 4410285.0,
                                  1812406.0,  -1664437.0, -1427072.0, 1704210.0};
        return t[i];
    }
    // nearest multiple of 2^24 (x + 3*2^75 lands where the ulp is 2^24)
    MB_HD static double rnd24(double x) { return d_add(d_add(x, M24), -M24); }

    MB_HD static FpD zero() {
        FpD r;
        MB_UNROLL
        for (int i = 0; i < N; ++i) r.v[i] = 0.0;
        return r;
    }

    // serial carry pass: balanced digits, top limb absorbs
    MB_HD void carry() {
        MB_UNROLL
        for (int k = 0; k < N - 1; ++k) {
            double c = rnd24(v[k]);
            v[k] = d_add(v[k], -c);
            v[k + 1] = d_fma(c, INV24, v[k + 1]);
        }
    }
    // value reduction into about (-0.51 p, 0.51 p) followed by a carry pass;
    // input limbs may be a few times 2^24, the value a few times p
    MB_HD void reduce() {
        double k0 = d_mul(v[N - 1], 1.0 / 1704210.0);
        double k = d_add(d_add(k0, 6755399441055744.0), -6755399441055744.0);  // nearest integer
        MB_UNROLL
        for (int j = 0; j < N; ++j) v[j] = d_fma(-k, P(j), v[j]);
        carry();
    }

    MB_HD static FpD add(const FpD& a, const FpD& b) {
        FpD r;
        MB_UNROLL
        for (int i = 0; i < N; ++i) r.v[i] = d_add(a.v[i], b.v[i]);
        return r;
    }
    MB_HD static FpD sub(const FpD& a, const FpD& b) {
        FpD r;
        MB_UNROLL
        for (int i = 0; i < N; ++i) r.v[i] = d_add(a.v[i], -b.v[i]);
        return r;
    }
    // s * a - b with s = +-1
    MB_HD static FpD ssub(double s, const FpD& a, const FpD& b) {
        FpD r;
        MB_UNROLL
        for (int i = 0; i < N; ++i) r.v[i] = d_fma(s, a.v[i], -b.v[i]);
        return r;
    }

    // Montgomery step for column i of the running product t[]
    MB_HD static void redc_col(double* t, int i) {
        double c = rnd24(t[i]);
        double lo = d_add(t[i], -c);
        double m0 = d_mul(lo, PINV);
        double m = d_add(m0, -rnd24(m0));
        MB_UNROLL
        for (int j = 0; j < N; ++j) t[i + j] = d_fma(m, P(j), t[i + j]);
        t[i + 1] = d_fma(t[i], INV24, t[i + 1]);
    }
    MB_HD static FpD finish(double* t) {
        FpD r;
        MB_UNROLL
        for (int k = 0; k < N - 1; ++k) {
            double c = rnd24(t[N + k]);
            r.v[k] = d_add(t[N + k], -c);
            t[N + k + 1] = d_fma(c, INV24, t[N + k + 1]);
        }
        r.v[N - 1] = t[2 * N - 1];
        return r;
    }

    // a * b / 2^384 (mod p); |result| < |a||b| / 2^384 + p/2, limbs normalised.
    // Needs 16 * max|a_i| * max|b_i| <= 2^52.
    MB_HD static FpD mul(const FpD& a, const FpD& b) {
        double t[2 * N];
        MB_UNROLL
        for (int j = 0; j < N; ++j) t[j] = d_mul(a.v[j], b.v[0]);
        MB_UNROLL
        for (int j = N; j < 2 * N; ++j) t[j] = 0.0;
        redc_col(t, 0);
        MB_UNROLL
        for (int i = 1; i < N; ++i) {
            MB_UNROLL
            for (int j = 0; j < N; ++j) t[i + j] = d_fma(a.v[j], b.v[i], t[i + j]);
            redc_col(t, i);
        }
        return finish(t);
    }
    MB_HD static FpD sqr(const FpD& a) {
        double t[2 * N], a2[N];
        MB_UNROLL
        for (int j = 0; j < N; ++j) a2[j] = d_add(a.v[j], a.v[j]);
        MB_UNROLL
        for (int j = 0; j < 2 * N; ++j) t[j] = 0.0;
        MB_UNROLL
        for (int i = 0; i < N; ++i) {
            t[2 * i] = d_fma(a.v[i], a.v[i], t[2 * i]);
            MB_UNROLL
            for (int j = i + 1; j < N; ++j) t[i + j] = d_fma(a2[j], a.v[i], t[i + j]);
            redc_col(t, i);
        }
        return finish(t);
    }

    // canonical integer limbs (Montgomery form, < p) -> digits in [0, 2^24)
    MB_HD static FpD from_fp(const Fp& a) {
        FpD r;
        MB_UNROLL
        for (int g = 0; g < 4; ++g) {
            uint32_t w0 = a.v[3 * g], w1 = a.v[3 * g + 1], w2 = a.v[3 * g + 2];
            r.v[4 * g + 0] = d_from_u24(w0 & 0xffffffu);
            r.v[4 * g + 1] = d_from_u24((w0 >> 24) | ((w1 & 0xffffu) << 8));
            r.v[4 * g + 2] = d_from_u24((w1 >> 16) | ((w2 & 0xffu) << 16));
            r.v[4 * g + 3] = d_from_u24(w2 >> 8);
        }
        return r;
    }
    // any representative with |value| < p -> canonical integer limbs
    MB_HD static Fp to_fp(const FpD& x) {
        FpD y = x;
        y.reduce();                       // (-0.51p, 0.51p), balanced
        MB_UNROLL
        for (int j = 0; j < N; ++j) y.v[j] = d_add(y.v[j], P(j));  // (0.49p, 1.51p)
        // non-negative digits
        MB_UNROLL
        for (int k = 0; k < N - 1; ++k) {
            double c = rnd24(y.v[k]);
            double r = d_add(y.v[k], -c);
            if (r < 0.0) {
                r = d_add(r, TWO24);
                c = d_add(c, -TWO24);
            }
            y.v[k] = r;
            y.v[k + 1] = d_fma(c, INV24, y.v[k + 1]);
        }
        Fp s, t;
        MB_UNROLL
        for (int g = 0; g < 4; ++g) {
            uint32_t l0 = d_to_u32(y.v[4 * g]), l1 = d_to_u32(y.v[4 * g + 1]);
            uint32_t l2 = d_to_u32(y.v[4 * g + 2]), l3 = d_to_u32(y.v[4 * g + 3]);
            s.v[3 * g + 0] = l0 | (l1 << 24);
            s.v[3 * g + 1] = (l1 >> 8) | (l2 << 16);
            s.v[3 * g + 2] = (l2 >> 16) | (l3 << 8);
        }
        // s in (0, 2p): one conditional subtraction
        t.v[0] = sub_cc(s.v[0], FpCfg::mod(0));
        MB_UNROLL
        for (int i = 1; i < 12; ++i) t.v[i] = subc_cc(s.v[i], FpCfg::mod(i));
        uint32_t borrow = subc(0, 0);
        MB_UNROLL
        for (int i = 0; i < 12; ++i) s.v[i] = borrow ? s.v[i] : t.v[i];
        return s;
    }
    // the Montgomery one (2^384 mod p), balanced digits
    MB_HD static FpD one() {
        FpD r = from_fp(Fp::one());
        r.carry();
        return r;
    }
};

}  // namespace mb
