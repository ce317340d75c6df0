// Small utility kernels: canonical-scalar validation, device self-test of the
// register-level arithmetic, Fp multiplication micro-benchmark.
#pragma once
#include "synth.cuh"

namespace mb {

// ---------------------------------------------------------------------------
// validation, self-test and micro-benchmark kernels
// ---------------------------------------------------------------------------
struct ValidateArgs {
    size_t nthreads;  // rows_n * per_row
    const Fr* base;
    size_t per_row, row_stride;
    uint32_t* flag;
};
MB_HD void validate_body(const ValidateArgs& a, size_t tid) {
    size_t row = tid / a.per_row, i = tid - row * a.per_row;
    if (Fr::std_ge_mod(a.base[row * a.row_stride + i])) *a.flag = 1;
}
MB_K_MISC(validate_scalars, ValidateArgs, validate_body, 256)

struct SelfTestArgs {
    size_t nthreads;
    uint32_t* mismatches;
    G1Affine g1;
    G2Affine g2;
};
template <class F>
MB_HD F st_rand(uint64_t key, uint64_t i) {
    F t;
    for (int j = 0; j < F::N; ++j) t.v[j] = (uint32_t)mix64(key + i * 131 + j);
    t.v[F::N - 1] &= 0x0fffffffu;  // < modulus for both fields
    return t;
}
template <class F>
MB_HD uint32_t st_field(uint64_t key, size_t tid) {
    uint32_t bad = 0;
    F a = st_rand<F>(key, 2 * tid), b = st_rand<F>(key, 2 * tid + 1);
    if (tid % 5 == 0) a = F::sub(F::zero(), F::one());  // p - 1 (as a plain integer pattern)
    if (!F::mul(a, b).eq(F::mul_portable(a, b))) bad++;
    if (!F::sqr(a).eq(F::mul_portable(a, a))) bad++;
    F s = F::add(a, b);
    if (!F::sub(s, b).eq(a)) bad++;
    if (!F::add(a, F::neg(a)).is_zero()) bad++;
    if (!F::mul(F::add(a, b), a).eq(F::add(F::mul_portable(a, a), F::mul_portable(b, a)))) bad++;
    if (!F::to_std(F::from_std(a)).eq(a)) bad++;
    // binary-Euclid inverse against the Fermat power (a few threads: the power is ~570 multiplications),
    // and a * a^-1 = 1 everywhere; 1, 2, p - 1 and 0 among the operands
    {
        F x = a;
        if (tid % 11 == 1) x = F::one();
        if (tid % 11 == 2) x = F::dbl(F::one());
        if (tid % 11 == 3) x = F::zero();
        F xi = F::inv(x);
        if (x.is_zero() ? !xi.is_zero() : !F::mul(x, xi).eq(F::one())) bad++;
        if (tid < 24 && !xi.eq(F::inv_fermat(x))) bad++;
    }
    // fused a b + c d (+ ...) against separate multiplications and additions, the extreme case
    // (p-1)^2 + (p-1)^2 (+ ...) included; only where the modulus leaves the head-room (Fp)
    if constexpr (F::SOP4_OK) {
        if (!F::sqr_inline(a).eq(F::mul_portable(a, a))) bad++;  // the triangular squaring of the accumulate kernels
        if (!F::sqr_inline(b).eq(F::mul_portable(b, b))) bad++;
        F c = st_rand<F>(key ^ 0x9e37, 2 * tid), d = st_rand<F>(key ^ 0x9e37, 2 * tid + 1);
        if (tid % 7 == 0) c = F::sub(F::zero(), F::one());
        if (tid % 35 == 0) { b = a; d = c; }  // a = b = c = d = p - 1 on both cycles
        if (!F::sop2_inline(a, b, c, d).eq(F::add(F::mul_portable(a, b), F::mul_portable(c, d)))) bad++;
        if (!F::sop4_inline(a, b, c, d, b, c, d, a).eq(F::add(F::add(F::mul_portable(a, b), F::mul_portable(c, d)),
                                                             F::add(F::mul_portable(b, c), F::mul_portable(d, a))))) bad++;
    }
    return bad;
}
MB_HD void selftest_body(const SelfTestArgs& a, size_t tid) {
    uint32_t bad = st_field<Fp>(0x1234, tid) + st_field<Fr>(0x5678, tid);
    if (tid < 64) {
        // curve identities on small multiples of the generators
        uint32_t k1[8] = {(uint32_t)tid + 2, 0, 0, 0, 0, 0, 0, 0}, k2[8] = {(uint32_t)(3 * tid + 5), 0, 0, 0, 0, 0, 0, 0};
        uint32_t k3[8] = {(uint32_t)(4 * tid + 7), 0, 0, 0, 0, 0, 0, 0};
        G1XYZZ p = xyzz_mul_affine(a.g1, k1), q = xyzz_mul_affine(a.g1, k2), e = xyzz_mul_affine(a.g1, k3);
        G1XYZZ t = p;
        xyzz_add_cold(t, q);
        G1Affine ta = xyzz_to_affine(t), ea = xyzz_to_affine(e);
        if (!ta.x.eq(ea.x) || !ta.y.eq(ea.y)) bad++;
        // y^2 = x^3 + 4
        Fp four = Fp::dbl(Fp::dbl(Fp::one()));
        if (!Fp::sqr(ta.y).eq(Fp::add(Fp::mul(Fp::sqr(ta.x), ta.x), four))) bad++;
        // mixed add of the affine image, doubling and cancellation branches
        G1XYZZ u = p;
        xyzz_madd_cold(u, xyzz_to_affine(q), false);
        G1Affine ua = xyzz_to_affine(u);
        if (!ua.x.eq(ea.x) || !ua.y.eq(ea.y)) bad++;
        G1XYZZ d = p;
        xyzz_madd_cold(d, xyzz_to_affine(p), false);
        G1XYZZ d2 = xyzz_dbl_cold(p);
        G1Affine da = xyzz_to_affine(d), d2a = xyzz_to_affine(d2);
        if (!da.x.eq(d2a.x) || !da.y.eq(d2a.y)) bad++;
        G1XYZZ z = p;
        xyzz_madd_cold(z, xyzz_to_affine(p), true);
        if (!z.is_inf()) bad++;
        uint8_t buf[96];
        G1Affine back;
        g1_encode(ta, buf);
        if (!g1_decode(buf, back) || !back.x.eq(ta.x) || !back.y.eq(ta.y)) bad++;
        if (tid < 8) {
            G2XYZZ p2 = xyzz_mul_affine(a.g2, k1), q2 = xyzz_mul_affine(a.g2, k2), e2 = xyzz_mul_affine(a.g2, k3);
            xyzz_add_cold(p2, q2);
            G2Affine t2 = xyzz_to_affine(p2), e2a = xyzz_to_affine(e2);
            if (!t2.x.eq(e2a.x) || !t2.y.eq(e2a.y)) bad++;
            uint8_t b2[192];
            G2Affine back2;
            g2_encode(t2, b2);
            if (!g2_decode(b2, back2) || !back2.x.eq(t2.x) || !back2.y.eq(t2.y)) bad++;
        }
    }
    if (bad) MB_ATOMIC_ADD(a.mismatches, bad);
}
MB_K_MISC(selftest_kernel, SelfTestArgs, selftest_body, 64)

struct FpMulBenchArgs {
    size_t nthreads;
    Fp* sink;
    uint32_t iters;
};
MB_HD void fpmul_bench_body(const FpMulBenchArgs& a, size_t tid) {
    Fp x = st_rand<Fp>(1, tid), y = st_rand<Fp>(2, tid), z = st_rand<Fp>(3, tid), w = st_rand<Fp>(4, tid);
    for (uint32_t i = 0; i < a.iters; ++i) {
        x = Fp::mul_inline(x, y);  // the inlined multiplier of the hot kernel, not the out-of-line copy
        y = Fp::mul_inline(y, z);
        z = Fp::mul_inline(z, w);
        w = Fp::mul_inline(w, x);
    }
    if (x.v[0] == 0x12345 && y.v[1] == 7) a.sink[tid] = Fp::add(Fp::add(x, y), Fp::add(z, w));
}
MB_K_MISC(fpmul_bench, FpMulBenchArgs, fpmul_bench_body, 256)

// single-warp latency of dependent operations (mode 0: inlined multiply,
// 1: out-of-line multiply, 2: out-of-line XYZZ doubling, 3: out-of-line XYZZ add)
struct LatencyArgs {
    size_t nthreads;
    Fp* sink;
    uint32_t iters, mode;
    G1Affine g1;
};
MB_HD void latency_body(const LatencyArgs& a, size_t tid) {
    Fp x = st_rand<Fp>(1, tid), y = st_rand<Fp>(2, tid);
    G1XYZZ p = G1XYZZ::from_affine(a.g1), q = xyzz_dbl_cold(p);
    MB_NOUNROLL
    for (uint32_t i = 0; i < a.iters; ++i) {
        if (a.mode == 0) x = Fp::mul_inline(x, y);
        else if (a.mode == 1) x = Fp::mul(x, y);
        else if (a.mode == 2) p = xyzz_dbl_cold(p);
        else xyzz_add_cold(p, q);
    }
    a.sink[tid] = Fp::add(x, p.x);
}
MB_K_MISC(latency_kernel, LatencyArgs, latency_body, 32)

struct IotaArgs {
    size_t nthreads;
    uint32_t* out;
};
MB_HD void iota_body(const IotaArgs& a, size_t tid) { a.out[tid] = (uint32_t)tid; }
MB_K_MISC(iota_kernel, IotaArgs, iota_body, 256)

struct SumPartialsArgs {
    size_t nthreads;  // 1
    const G1XYZZ* parts;
    size_t count;
    uint8_t* out;
};
MB_HD void sum_partials_body(const SumPartialsArgs& a, size_t) {
    G1XYZZ acc = G1XYZZ::inf();
    for (size_t i = 0; i < a.count; ++i) xyzz_add(acc, a.parts[i]);
    g1_encode(xyzz_to_affine(acc), a.out);
}
MB_K_MISC(sum_partials, SumPartialsArgs, sum_partials_body, 32)


}  // namespace mb
