// Groth16 verification on the device: the check the reference runs on the CPU
// right after proving (`verify_proof(verifying_key, &proof, &public_input)` at
// masp_proofs/src/sapling/prover.rs:148 and :266; SURVEY.md §8 a-8 / NEXT-3),
//
//     e(A, B) = e(alpha, beta) * e(sum x_i IC_i, gamma) * e(C, delta),
//
// one thread per proof of a batch.  A batch check is a few hundred thousand
// dependent field multiplications per thread, i.e. latency-bound: it occupies a
// handful of warps and runs next to the bucket kernels of the following chunk.
//
// Tower: Fp2 = Fp[u]/(u^2 + 1), Fp6 = Fp2[v]/(v^3 - (1 + u)), Fp12 = Fp6[w]/(w^2 - v);
// optimal-ate Miller loop over |x| = 0xd201000000010000 with the G2 point kept
// in Jacobian coordinates (no inversions; every line is scaled by an Fp2 factor,
// which the final exponentiation kills); final exponentiation = easy part by
// conjugation, one inversion and the p^2-Frobenius, hard part (cubed) by the
// BLS12 x-chain: five powers by |x| and two Frobenius maps, pinned against the
// plain power (p^4 - p^2 + 1) / r by the device self-test.
#pragma once
#include "ec.cuh"

namespace mb {

struct Fp6 {
    Fp2 c0, c1, c2;
};
struct Fp12 {
    Fp6 c0, c1;
};

struct PairConst {
    // xi^((p^2 - 1) k / 6) for k = 1..5, all in Fp (Montgomery form)
    MB_HD static Fp w(int k) {
        constexpr uint32_t t[5][12] = {
            {0x798dba3au, 0xecfb361bu, 0x91865a2cu, 0xc100ddb8u, 0x232bda8eu, 0x0ec08ff1u, 0xf1ca4721u, 0xd5c13cc6u, 0xbf7b5c04u, 0x47222a47u, 0xe51c5f59u, 0x0110f184u},
            {0x798a64e8u, 0x30f1361bu, 0x7ece5a2au, 0xf3b8ddabu, 0xc61577f7u, 0x16a8ca3au, 0x74fd029bu, 0xc26a2ff8u, 0x60701c6eu, 0x3636b766u, 0x241b6160u, 0x051ba4abu},
            {0xfffcaaaeu, 0x43f5ffffu, 0xed47fffdu, 0x32b7fff2u, 0xa2e99d69u, 0x07e83a49u, 0x8332bb7au, 0xeca8f331u, 0xa0f4c069u, 0xef148d1eu, 0x3eff0206u, 0x040ab326u},
            {0x8671f071u, 0xcd03c9e4u, 0x1fcda5d2u, 0x5dab2246u, 0xd3851b95u, 0x587042afu, 0x01bacb9eu, 0x8eb60ebeu, 0x83d050d2u, 0x03f97d6eu, 0x54638741u, 0x18f02065u},
            {0x867545c3u, 0x890dc9e4u, 0x3285a5d5u, 0x2af32253u, 0x309b7e2cu, 0x50880866u, 0x7e881024u, 0xa20d1b8cu, 0xe2db9068u, 0x14e4f04fu, 0x1564853au, 0x14e56d3fu}};
        Fp r;
        for (int i = 0; i < 12; ++i) r.v[i] = t[k - 1][i];
        return r;
    }
    // xi^((p - 1) k / 6) for k = 1..5, in Fp2 (Montgomery form): the p-Frobenius of the basis w^k
    MB_HD static Fp2 g(int k) {
        constexpr uint32_t t[5][2][12] = {
            {{0xb319d465u, 0x07089552u, 0xb50a8313u, 0xc6695f92u, 0xd117228fu, 0x97e83cccu, 0xb2dc29eeu, 0xa35baecau, 0x5daace4du, 0x1ce393eau, 0xb0fb66ebu, 0x08f2220fu},
             {0x4ce5d646u, 0xb2f66aadu, 0xfc497cecu, 0x5842a06bu, 0x2599d394u, 0xcf4895d4u, 0x40a8e8d0u, 0xc11b9cbau, 0xe5a0de89u, 0x2e3813cbu, 0x88847fafu, 0x110eefdau}},
            {{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
             {0x8671f071u, 0xcd03c9e4u, 0x1fcda5d2u, 0x5dab2246u, 0xd3851b95u, 0x587042afu, 0x01bacb9eu, 0x8eb60ebeu, 0x83d050d2u, 0x03f97d6eu, 0x54638741u, 0x18f02065u}},
            {{0x5aa30fdau, 0x7bcfa7a2u, 0x2a927e7cu, 0xdc17dec1u, 0x6b4ebef1u, 0x2f088dd8u, 0xda74d4a7u, 0xd1ca2087u, 0x96cebc1du, 0x2da25966u, 0xbbfd87d2u, 0x0e2b7eedu},
             {0x5aa30fdau, 0x7bcfa7a2u, 0x2a927e7cu, 0xdc17dec1u, 0x6b4ebef1u, 0x2f088dd8u, 0xda74d4a7u, 0xd1ca2087u, 0x96cebc1du, 0x2da25966u, 0xbbfd87d2u, 0x0e2b7eedu}},
            {{0x867545c3u, 0x890dc9e4u, 0x3285a5d5u, 0x2af32253u, 0x309b7e2cu, 0x50880866u, 0x7e881024u, 0xa20d1b8cu, 0xe2db9068u, 0x14e4f04fu, 0x1564853au, 0x14e56d3fu},
             {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}},
            {{0x0dbce43fu, 0x82d83cf5u, 0xdf9d018fu, 0xa2813e53u, 0x3c65e181u, 0xc6f0caa5u, 0x8d50fe95u, 0x7525cf52u, 0xf4798a6bu, 0x4a85ed50u, 0x6cf8eebdu, 0x171da0fdu},
             {0xf242c66cu, 0x3726c30au, 0xd1b6fe70u, 0x7c2ac1aau, 0xba4b14a2u, 0xa04007fbu, 0x66341429u, 0xef517c32u, 0x4ed2226bu, 0x0095ba65u, 0xcc86f7ddu, 0x02e370ecu}}};
        Fp2 r;
        for (int i = 0; i < 12; ++i) {
            r.c0.v[i] = t[k - 1][0][i];
            r.c1.v[i] = t[k - 1][1][i];
        }
        return r;
    }
    // (p^4 - p^2 + 1) / r, 1268 bits, little-endian words (the plain-power form of the hard part; the
    // product path uses the x-chain below, this constant pins it in the self-test)
    MB_HD static uint32_t hard(int i) {
        constexpr uint32_t t[40] = {
            0x38e3ba79u, 0xe516c3f4u, 0xe208ccf1u, 0xfa9912aau, 0x335d5b68u, 0x905ce937u, 0xb0dea236u, 0xc71a2629u,
            0x996754c8u, 0x83774940u, 0xb6a1e799u, 0x21d160aeu, 0xed237db4u, 0x2ed0b283u, 0x6c6f1821u, 0x915c97f3u,
            0xde783765u, 0x67f17fcbu, 0x9096d1b7u, 0x2378b903u, 0x1bdc51dcu, 0x7988f876u, 0x03fc77a1u, 0x20769950u,
            0xa621315bu, 0x827eca0bu, 0x8d63cb9fu, 0xe5a72bceu, 0xc28b6f8au, 0xf68f7764u, 0xcf081517u, 0x2f230063u,
            0x528d6a9au, 0x94506632u, 0xeb996ca3u, 0xd3cde88eu, 0x195c899eu, 0xc0bd38c3u, 0x3d807d01u, 0x000f686bu};
        return t[i];
    }
    static constexpr int HARD_BITS = 1268;
    static constexpr unsigned long long X = 0xd201000000010000ull;  // |x|; x itself is negative
};

// ---------------------------------------------------------------------------
// tower arithmetic (out of line: this code is latency-bound, keep it small)
// ---------------------------------------------------------------------------
MB_HD Fp2 f2_mul_xi(const Fp2& a) { return {Fp::sub(a.c0, a.c1), Fp::add(a.c0, a.c1)}; }
MB_HD Fp2 f2_scale(const Fp2& a, const Fp& k) { return {Fp::mul(a.c0, k), Fp::mul(a.c1, k)}; }
MB_COLD Fp2 f2_mul(const Fp2& a, const Fp2& b) { return Fp2::mul(a, b); }
MB_COLD Fp2 f2_sqr(const Fp2& a) { return Fp2::sqr(a); }

MB_HD Fp6 f6_zero() { return {Fp2::zero(), Fp2::zero(), Fp2::zero()}; }
MB_HD Fp6 f6_add(const Fp6& a, const Fp6& b) { return {Fp2::add(a.c0, b.c0), Fp2::add(a.c1, b.c1), Fp2::add(a.c2, b.c2)}; }
MB_HD Fp6 f6_sub(const Fp6& a, const Fp6& b) { return {Fp2::sub(a.c0, b.c0), Fp2::sub(a.c1, b.c1), Fp2::sub(a.c2, b.c2)}; }
MB_HD Fp6 f6_neg(const Fp6& a) { return {Fp2::neg(a.c0), Fp2::neg(a.c1), Fp2::neg(a.c2)}; }
MB_HD Fp6 f6_mul_v(const Fp6& a) { return {f2_mul_xi(a.c2), a.c0, a.c1}; }
MB_COLD Fp6 f6_mul(const Fp6& a, const Fp6& b) {
    Fp2 t0 = f2_mul(a.c0, b.c0), t1 = f2_mul(a.c1, b.c1), t2 = f2_mul(a.c2, b.c2);
    Fp6 r;
    r.c0 = Fp2::add(t0, f2_mul_xi(Fp2::add(f2_mul(a.c1, b.c2), f2_mul(a.c2, b.c1))));
    r.c1 = Fp2::add(Fp2::add(f2_mul(a.c0, b.c1), f2_mul(a.c1, b.c0)), f2_mul_xi(t2));
    r.c2 = Fp2::add(Fp2::add(f2_mul(a.c0, b.c2), f2_mul(a.c2, b.c0)), t1);
    return r;
}
MB_COLD Fp6 f6_inv(const Fp6& a) {
    Fp2 c0 = Fp2::sub(f2_sqr(a.c0), f2_mul_xi(f2_mul(a.c1, a.c2)));
    Fp2 c1 = Fp2::sub(f2_mul_xi(f2_sqr(a.c2)), f2_mul(a.c0, a.c1));
    Fp2 c2 = Fp2::sub(f2_sqr(a.c1), f2_mul(a.c0, a.c2));
    Fp2 t = Fp2::add(f2_mul(a.c0, c0), f2_mul_xi(Fp2::add(f2_mul(a.c2, c1), f2_mul(a.c1, c2))));
    Fp2 ti = Fp2::inv(t);
    return {f2_mul(c0, ti), f2_mul(c1, ti), f2_mul(c2, ti)};
}

MB_HD Fp12 f12_one() { return {{Fp2::one(), Fp2::zero(), Fp2::zero()}, f6_zero()}; }
MB_COLD Fp12 f12_mul(const Fp12& a, const Fp12& b) {
    Fp6 t0 = f6_mul(a.c0, b.c0), t1 = f6_mul(a.c1, b.c1);
    Fp12 r;
    r.c1 = f6_sub(f6_sub(f6_mul(f6_add(a.c0, a.c1), f6_add(b.c0, b.c1)), t0), t1);
    r.c0 = f6_add(t0, f6_mul_v(t1));
    return r;
}
MB_HD Fp12 f12_conj(const Fp12& a) { return {a.c0, f6_neg(a.c1)}; }
MB_COLD Fp12 f12_inv(const Fp12& a) {
    Fp6 t = f6_sub(f6_mul(a.c0, a.c0), f6_mul_v(f6_mul(a.c1, a.c1)));
    Fp6 ti = f6_inv(t);
    return {f6_mul(a.c0, ti), f6_neg(f6_mul(a.c1, ti))};
}
// a^(p^2): the Fp2 coefficients are fixed, the basis element v^i w^j picks up xi^((p^2-1)(2i+j)/6)
MB_COLD Fp12 f12_frob2(const Fp12& a) {
    Fp12 r;
    r.c0.c0 = a.c0.c0;
    r.c0.c1 = f2_scale(a.c0.c1, PairConst::w(2));
    r.c0.c2 = f2_scale(a.c0.c2, PairConst::w(4));
    r.c1.c0 = f2_scale(a.c1.c0, PairConst::w(1));
    r.c1.c1 = f2_scale(a.c1.c1, PairConst::w(3));
    r.c1.c2 = f2_scale(a.c1.c2, PairConst::w(5));
    return r;
}
// a^p: conjugate the Fp2 coefficients, the basis element w^k picks up xi^((p-1) k / 6)
MB_HD Fp2 f2_conj(const Fp2& a) { return {a.c0, Fp::neg(a.c1)}; }
MB_COLD Fp12 f12_frob1(const Fp12& a) {
    Fp12 r;
    r.c0.c0 = f2_conj(a.c0.c0);
    r.c0.c1 = f2_mul(f2_conj(a.c0.c1), PairConst::g(2));
    r.c0.c2 = f2_mul(f2_conj(a.c0.c2), PairConst::g(4));
    r.c1.c0 = f2_mul(f2_conj(a.c1.c0), PairConst::g(1));
    r.c1.c1 = f2_mul(f2_conj(a.c1.c1), PairConst::g(3));
    r.c1.c2 = f2_mul(f2_conj(a.c1.c2), PairConst::g(5));
    return r;
}
// a^|x|
MB_COLD Fp12 f12_pow_x(const Fp12& a) {
    Fp12 r = a;
    MB_NOUNROLL
    for (int bit = 62; bit >= 0; --bit) {
        r = f12_mul(r, r);
        if ((PairConst::X >> bit) & 1) r = f12_mul(r, a);
    }
    return r;
}
MB_HD bool f12_is_one(const Fp12& a) {
    return a.c0.c0.eq(Fp2::one()) && a.c0.c1.is_zero() && a.c0.c2.is_zero() && a.c1.c0.is_zero() && a.c1.c1.is_zero() &&
           a.c1.c2.is_zero();
}

// ---------------------------------------------------------------------------
// Miller loop
// ---------------------------------------------------------------------------
struct G2Jac {
    Fp2 x, y, z;
};
// line through the untwisted T, scaled by an Fp2 factor, evaluated at P = (xp, yp):
//   c_w3 * yp * w^3  +  c_w2 * xp * w^2  +  c_1        (w^2 = v, w^3 = v w)
MB_HD Fp12 line_value(const Fp2& c_1, const Fp2& c_w2, const Fp2& c_w3, const G1Affine& p) {
    Fp12 l;
    l.c0 = {c_1, f2_scale(c_w2, p.x), Fp2::zero()};
    l.c1 = {Fp2::zero(), f2_scale(c_w3, p.y), Fp2::zero()};
    return l;
}
// T <- 2 T; returns the tangent line at (the old) T
MB_COLD Fp12 miller_double(G2Jac& t, const G1Affine& p) {
    Fp2 A = f2_sqr(t.x), B = f2_sqr(t.y), C = f2_sqr(B), Z2 = f2_sqr(t.z);
    Fp2 N = Fp2::add(Fp2::dbl(A), A);            // 3 X^2
    Fp2 D = Fp2::dbl(f2_mul(t.y, t.z));          // 2 Y Z
    // lambda = N / D; line * (D Z^2):  (D Z^2) yp w^3 - (N Z^2) xp w^2 + (N X - 2 Y^2)
    Fp12 l = line_value(Fp2::sub(f2_mul(N, t.x), Fp2::dbl(B)), Fp2::neg(f2_mul(N, Z2)), f2_mul(D, Z2), p);
    Fp2 S = Fp2::dbl(Fp2::dbl(f2_mul(t.x, B)));  // 4 X Y^2
    Fp2 X3 = Fp2::sub(f2_sqr(N), Fp2::dbl(S));
    Fp2 C8 = Fp2::dbl(Fp2::dbl(Fp2::dbl(C)));
    t.y = Fp2::sub(f2_mul(N, Fp2::sub(S, X3)), C8);
    t.x = X3;
    t.z = D;
    return l;
}
// T <- T + Q (Q affine); returns the chord through T and Q
MB_COLD Fp12 miller_add(G2Jac& t, const G2Affine& q, const G1Affine& p) {
    Fp2 Z2 = f2_sqr(t.z);
    Fp2 Z3 = f2_mul(Z2, t.z);
    Fp2 theta = Fp2::sub(t.y, f2_mul(q.y, Z3));
    Fp2 mu = Fp2::sub(t.x, f2_mul(q.x, Z2));
    Fp2 D = f2_mul(t.z, mu);
    // lambda = theta / D; line * D:  D yp w^3 - theta xp w^2 + (theta xq - D yq)
    Fp12 l = line_value(Fp2::sub(f2_mul(theta, q.x), f2_mul(D, q.y)), Fp2::neg(theta), D, p);
    Fp2 mu2 = f2_sqr(mu);
    Fp2 mu3 = f2_mul(mu2, mu);
    Fp2 xm2 = f2_mul(t.x, mu2);
    Fp2 X3 = Fp2::sub(Fp2::add(f2_sqr(theta), mu3), Fp2::dbl(xm2));
    t.y = Fp2::sub(f2_mul(theta, Fp2::sub(xm2, X3)), f2_mul(t.y, mu3));
    t.x = X3;
    t.z = D;
    return l;
}
// prod_i f_{|x|,Q_i}(P_i), conjugated (x < 0); pairs with an identity contribute 1
MB_COLD Fp12 miller_multi(const G1Affine* ps, const G2Affine* qs, int n) {
    G2Jac t[3];
    bool live[3];
    for (int i = 0; i < n; ++i) {
        live[i] = !ps[i].is_inf() && !qs[i].is_inf();
        t[i] = {qs[i].x, qs[i].y, Fp2::one()};
    }
    Fp12 f = f12_one();
    MB_NOUNROLL
    for (int bit = 62; bit >= 0; --bit) {
        f = f12_mul(f, f);
        MB_NOUNROLL
        for (int i = 0; i < n; ++i)
            if (live[i]) f = f12_mul(f, miller_double(t[i], ps[i]));
        if ((PairConst::X >> bit) & 1) {
            MB_NOUNROLL
            for (int i = 0; i < n; ++i)
                if (live[i]) f = f12_mul(f, miller_add(t[i], qs[i], ps[i]));
        }
    }
    return f12_conj(f);
}
MB_COLD Fp12 final_exp_easy(const Fp12& f) {
    Fp12 f1 = f12_mul(f12_conj(f), f12_inv(f));  // f^(p^6 - 1)
    return f12_mul(f12_frob2(f1), f1);           // ^(p^2 + 1): unitary from here on, inverse = conjugate
}
// The hard part as a plain power (p^4 - p^2 + 1) / r: 1268 squarings.  Reference form, used by
// the self-test to pin the chain below.
MB_COLD Fp12 final_exp_hard_plain(const Fp12& e) {
    Fp12 r = f12_one();
    MB_NOUNROLL
    for (int i = PairConst::HARD_BITS - 1; i >= 0; --i) {
        r = f12_mul(r, r);
        if ((PairConst::hard(i >> 5) >> (i & 31)) & 1) r = f12_mul(r, e);
    }
    return r;
}
// The hard part cubed, by the BLS12 identity
//     3 (p^4 - p^2 + 1) / r = (x - 1)^2 (x + p) (x^2 + p^2 - 1) + 3,
// five powers by |x| (x < 0: a^x = conj(a^|x|) on unitary a) and two Frobenius maps.  Cubing is
// harmless for the comparison with one: the result lies in the order-r subgroup and 3 does not divide r.
MB_COLD Fp12 final_exp_hard_cubed(const Fp12& e) {
    Fp12 a = f12_conj(f12_mul(f12_pow_x(e), e));      // e^(x - 1)
    a = f12_conj(f12_mul(f12_pow_x(a), a));           // e^((x - 1)^2)
    Fp12 b = f12_mul(f12_conj(f12_pow_x(a)), f12_frob1(a));                               // a^(x + p)
    Fp12 c = f12_mul(f12_mul(f12_pow_x(f12_pow_x(b)), f12_frob2(b)), f12_conj(b));        // b^(x^2 + p^2 - 1)
    return f12_mul(c, f12_mul(f12_mul(e, e), e));
}
MB_COLD Fp12 final_exponentiation(const Fp12& f) { return final_exp_hard_cubed(final_exp_easy(f)); }

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// Verifying key on the device (from the vk prefix of the Parameters bytes)
struct VkDev {
    const G1Affine* ic;      // n_inputs points
    const G2Affine* gamma;   // gamma_g2
    const G2Affine* delta;   // delta_g2
    const Fp12* alpha_beta;  // Miller value of (-alpha_g1, beta_g2), computed once per key
};

struct PairPrepArgs {
    size_t nthreads;  // 1
    const G1Affine* alpha;
    const G2Affine* beta;
    Fp12* out;
};
MB_HD void pair_prep_body(const PairPrepArgs& a, size_t) {
    G1Affine p = *a.alpha;
    p.y = Fp::neg(p.y);
    G2Affine q = *a.beta;
    *a.out = miller_multi(&p, &q, 1);
}

struct VerifyArgs {
    size_t nthreads;  // proofs
    const G1Affine* pa;  // proof.A
    const G2Affine* pb;  // proof.B
    const G1Affine* pc;  // proof.C
    const uint32_t* inputs;   // per proof: n_inputs scalars (plain, 8 words each), inputs[0] = 1
    size_t input_stride;      // in scalars
    uint32_t n_inputs;
    VkDev vk;
    uint32_t* ok;  // per proof: 1 = the equation holds
};
MB_HD void verify_body(const VerifyArgs& a, size_t tid) {
    // acc = IC_0 + sum_{i >= 1} x_i IC_i
    G1XYZZ acc = G1XYZZ::from_affine(a.vk.ic[0]);
    const uint32_t* x = a.inputs + tid * a.input_stride * 8;
    MB_NOUNROLL
    for (uint32_t i = 1; i < a.n_inputs; ++i) {
        G1XYZZ t = xyzz_mul(G1XYZZ::from_affine(a.vk.ic[i]), x + 8 * i);
        xyzz_add_cold(acc, t);
    }
    G1Affine ps[3];
    G2Affine qs[3];
    ps[0] = a.pa[tid];
    qs[0] = a.pb[tid];
    ps[1] = xyzz_to_affine(acc);
    ps[1].y = Fp::neg(ps[1].y);
    qs[1] = *a.vk.gamma;
    ps[2] = a.pc[tid];
    ps[2].y = Fp::neg(ps[2].y);
    qs[2] = *a.vk.delta;
    Fp12 f = f12_mul(miller_multi(ps, qs, 3), *a.vk.alpha_beta);
    a.ok[tid] = f12_is_one(final_exponentiation(f)) ? 1u : 0u;
}

// ---------------------------------------------------------------------------
// compressed proof points (Proof::read: zkcrypto from_compressed, i.e. on-curve AND in the
// prime-order subgroup), for verifying proofs that arrive as their 192 wire bytes
// (masp_proofs/src/sapling/verifier/single.rs:60,77,93 read the zkproof this way)
// ---------------------------------------------------------------------------
MB_COLD Fp fp_pow_limbs(const Fp& a, const uint32_t* e, int nbits) {
    Fp r = Fp::one();
    MB_NOUNROLL
    for (int i = nbits - 1; i >= 0; --i) {
        r = Fp::sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1) r = Fp::mul(r, a);
    }
    return r;
}
MB_COLD Fp2 fp2_pow_limbs(const Fp2& a, const uint32_t* e, int nbits) {
    Fp2 r = Fp2::one();
    MB_NOUNROLL
    for (int i = nbits - 1; i >= 0; --i) {
        r = f2_sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1) r = f2_mul(r, a);
    }
    return r;
}
// (p + k) >> s as limbs, for the small k, s the square roots need
MB_HD void p_shifted(int add, int sub, int shift, uint32_t out[12]) {
    uint32_t t[12];
    uint64_t c = (uint64_t)add;
    for (int i = 0; i < 12; ++i) {
        c += FpCfg::mod(i);
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    uint64_t b = (uint64_t)sub;
    for (int i = 0; i < 12; ++i) {
        uint64_t d = (uint64_t)t[i] - b;
        t[i] = (uint32_t)d;
        b = (d >> 63) & 1;
    }
    for (int i = 0; i < 12; ++i) out[i] = (t[i] >> shift) | (shift && i + 1 < 12 ? t[i + 1] << (32 - shift) : 0);
}
// p = 3 mod 4: sqrt(a) = a^((p+1)/4) when a is a square
MB_COLD bool fp_sqrt(const Fp& a, Fp& out) {
    uint32_t e[12];
    p_shifted(1, 0, 2, e);
    out = fp_pow_limbs(a, e, 384);
    return Fp::sqr(out).eq(a);
}
// Fp2 = Fp[u]/(u^2+1), p = 3 mod 4 (Adj & Rodriguez-Henriquez, Algorithm 9)
MB_COLD bool fp2_sqrt(const Fp2& a, Fp2& out) {
    if (a.is_zero()) {
        out = Fp2::zero();
        return true;
    }
    uint32_t e[12];
    p_shifted(0, 3, 2, e);                      // (p - 3) / 4
    Fp2 a1 = fp2_pow_limbs(a, e, 384);
    Fp2 alpha = f2_mul(f2_sqr(a1), a);          // a^((p-1)/2)
    Fp2 x0 = f2_mul(a1, a);
    Fp2 minus_one = Fp2::neg(Fp2::one());
    Fp2 a0 = f2_mul(f2_conj(alpha), alpha);     // alpha^(p+1)
    if (a0.eq(minus_one)) return false;
    if (alpha.eq(minus_one)) {
        out = {Fp::neg(x0.c1), x0.c0};          // u * x0
    } else {
        p_shifted(0, 1, 1, e);                  // (p - 1) / 2
        Fp2 b = fp2_pow_limbs(Fp2::add(Fp2::one(), alpha), e, 384);
        out = f2_mul(b, x0);
    }
    return f2_sqr(out).eq(a);
}
MB_HD void fr_modulus_words(uint32_t r[8]) {
    for (int i = 0; i < 8; ++i) r[i] = FrCfg::mod(i);
}
// false: malformed, not on the curve, or outside the prime-order subgroup
MB_COLD bool g1_decode_compressed(const uint8_t* b, G1Affine& out) {
    if (!(b[0] & 0x80)) return false;
    if (b[0] & 0x40) {
        uint32_t o = b[0] & 0x3f & ~0x40u;
        for (int i = 1; i < 48; ++i) o |= b[i];
        out = G1Affine::inf();
        return o == 0;
    }
    uint8_t t[48];
    for (int i = 0; i < 48; ++i) t[i] = b[i];
    t[0] &= 0x1f;
    Fp x;
    fp_limbs_from_be(t, x);
    if (Fp::std_ge_mod(x)) return false;
    out.x = Fp::from_std(x);
    Fp four = Fp::dbl(Fp::dbl(Fp::one()));
    Fp y;
    if (!fp_sqrt(Fp::add(Fp::mul(Fp::sqr(out.x), out.x), four), y)) return false;
    bool larger = Fp::std_gt_half(Fp::to_std(y));
    if (larger != ((b[0] & 0x20) != 0)) y = Fp::neg(y);
    out.y = y;
    uint32_t r[8];
    fr_modulus_words(r);
    return xyzz_mul(G1XYZZ::from_affine(out), r).is_inf();
}
MB_COLD bool g2_decode_compressed(const uint8_t* b, G2Affine& out) {
    if (!(b[0] & 0x80)) return false;
    if (b[0] & 0x40) {
        uint32_t o = b[0] & 0x3f & ~0x40u;
        for (int i = 1; i < 96; ++i) o |= b[i];
        out = G2Affine::inf();
        return o == 0;
    }
    uint8_t t[48];
    for (int i = 0; i < 48; ++i) t[i] = b[i];
    t[0] &= 0x1f;
    Fp c1, c0;
    fp_limbs_from_be(t, c1);
    fp_limbs_from_be(b + 48, c0);
    if (Fp::std_ge_mod(c1) || Fp::std_ge_mod(c0)) return false;
    out.x = {Fp::from_std(c0), Fp::from_std(c1)};
    Fp four = Fp::dbl(Fp::dbl(Fp::one()));
    Fp2 bb = {four, four};  // 4 (1 + u)
    Fp2 y;
    if (!fp2_sqrt(Fp2::add(f2_mul(f2_sqr(out.x), out.x), bb), y)) return false;
    bool larger = y.c1.is_zero() ? Fp::std_gt_half(Fp::to_std(y.c0)) : Fp::std_gt_half(Fp::to_std(y.c1));
    if (larger != ((b[0] & 0x20) != 0)) y = Fp2::neg(y);
    out.y = y;
    uint32_t r[8];
    fr_modulus_words(r);
    return xyzz_mul(G2XYZZ::from_affine(out), r).is_inf();
}
struct ProofReadArgs {
    size_t nthreads;  // 3 * proofs
    const uint8_t* proofs;  // 192 bytes each: A (48) | B (96) | C (48), compressed
    G1Affine* pa;
    G2Affine* pb;
    G1Affine* pc;
    uint32_t* bad;  // per proof: set when any of its points is rejected
};
MB_HD void proof_read_body(const ProofReadArgs& a, size_t tid) {
    size_t proof = tid / 3;
    uint32_t which = (uint32_t)(tid - proof * 3);
    const uint8_t* p = a.proofs + 192 * proof;
    bool ok;
    if (which == 0) ok = g1_decode_compressed(p, a.pa[proof]);
    else if (which == 1) ok = g2_decode_compressed(p + 48, a.pb[proof]);
    else ok = g1_decode_compressed(p + 144, a.pc[proof]);
    // bellman's Proof::read: "point at infinity" is an error for A, B and C alike
    if (ok) ok = which == 1 ? !a.pb[proof].is_inf() : (which == 0 ? !a.pa[proof].is_inf() : !a.pc[proof].is_inf());
    if (!ok) {
        if (which == 0) a.pa[proof] = G1Affine::inf();
        else if (which == 1) a.pb[proof] = G2Affine::inf();
        else a.pc[proof] = G1Affine::inf();
        a.bad[proof] = 1;
    }
}

// ---------------------------------------------------------------------------
// randomised batch check (bellman groth16::batch::Verifier as used by
// masp_proofs/src/sapling/verifier/batch.rs:24-31, 85-160): with random 128-bit z_i,
//   prod_i e(z_i A_i, B_i) * e(-(sum_i z_i) alpha, beta) * e(-sum_j (sum_i z_i x_ij) IC_j, gamma)
//                          * e(-sum_i z_i C_i, delta) = 1
// holds for a batch of valid proofs and fails for any invalid one except with probability
// ~2^-128: n + 3 Miller loops and ONE final exponentiation instead of 3 n and n.
// ---------------------------------------------------------------------------
struct BatchMillerArgs {
    size_t nthreads;  // proofs
    const G1Affine* pa;
    const G2Affine* pb;
    const G1Affine* pc;
    const uint32_t* inputs;  // [n][n_inputs] plain scalars, inputs[.][0] = 1
    uint32_t n_inputs;
    const uint32_t* z;       // [n][4] random 128-bit coefficients
    Fp12* f;                 // per proof: Miller value of (z A, B)
    G1XYZZ* zc;              // per proof: z C
    Fr* zx;                  // [n][n_inputs]: z * x_j mod r (plain)
};
MB_HD void batch_miller_body(const BatchMillerArgs& a, size_t tid) {
    uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) k[i] = a.z[4 * tid + i];
    G1Affine za = xyzz_to_affine(xyzz_mul(G1XYZZ::from_affine(a.pa[tid]), k));
    G2Affine b = a.pb[tid];
    a.f[tid] = miller_multi(&za, &b, 1);
    a.zc[tid] = xyzz_mul(G1XYZZ::from_affine(a.pc[tid]), k);
    Fr zf;
    for (int i = 0; i < 8; ++i) zf.v[i] = k[i];
    Fr zm = Fr::from_std(zf);  // Montgomery form: montmul(x, z R) = x z on plain x
    const Fr* x = (const Fr*)(a.inputs + tid * (size_t)a.n_inputs * 8);
    // column 0 is the constant ONE whatever the caller stored there (verify_body ignores it too):
    // both entry points must give the same verdict on the same bytes
    a.zx[tid * (size_t)a.n_inputs] = zf;
    for (uint32_t j = 1; j < a.n_inputs; ++j) a.zx[tid * (size_t)a.n_inputs + j] = Fr::mul(x[j], zm);
}
struct BatchFinalArgs {
    size_t nthreads;  // 1
    size_t n;
    const Fp12* f;
    const G1XYZZ* zc;
    const Fr* zx;
    uint32_t n_inputs;
    const G1Affine* alpha;
    const G2Affine* beta;
    VkDev vk;
    uint32_t* ok;
};
MB_HD void batch_final_body(const BatchFinalArgs& a, size_t) {
    Fp12 F = f12_one();
    G1XYZZ sc = G1XYZZ::inf();
    MB_NOUNROLL
    for (size_t i = 0; i < a.n; ++i) {
        F = f12_mul(F, a.f[i]);
        xyzz_add_cold(sc, a.zc[i]);
    }
    // sum_j (sum_i z_i x_ij) IC_j; column 0 is sum_i z_i because x_i0 = 1
    G1XYZZ acc = G1XYZZ::inf();
    Fr sz = Fr::zero();
    MB_NOUNROLL
    for (uint32_t j = 0; j < a.n_inputs; ++j) {
        Fr sj = Fr::zero();
        MB_NOUNROLL
        for (size_t i = 0; i < a.n; ++i) sj = Fr::add(sj, a.zx[i * (size_t)a.n_inputs + j]);
        if (j == 0) sz = sj;
        G1XYZZ t = xyzz_mul(G1XYZZ::from_affine(a.vk.ic[j]), sj.v);
        xyzz_add_cold(acc, t);
    }
    G1Affine ps[3];
    G2Affine qs[3];
    ps[0] = xyzz_to_affine(acc);
    qs[0] = *a.vk.gamma;
    ps[1] = xyzz_to_affine(sc);
    qs[1] = *a.vk.delta;
    ps[2] = xyzz_to_affine(xyzz_mul(G1XYZZ::from_affine(*a.alpha), sz.v));
    qs[2] = *a.beta;
    for (int i = 0; i < 3; ++i) ps[i].y = Fp::neg(ps[i].y);
    Fp12 f = f12_mul(F, miller_multi(ps, qs, 3));
    *a.ok = f12_is_one(final_exponentiation(f)) ? 1u : 0u;
}

// self-test: the x-chain against the plain power, Frobenius maps against plain powers' structure
struct PairSelfTestArgs {
    size_t nthreads;  // 1
    const G1Affine* g1;
    const G2Affine* g2;
    uint32_t* mismatches;
};
MB_HD void pair_selftest_body(const PairSelfTestArgs& a, size_t) {
    uint32_t bad = 0;
    Fp12 f = miller_multi(a.g1, a.g2, 1);
    Fp12 e = final_exp_easy(f);
    Fp12 plain = final_exp_hard_plain(e);
    Fp12 cubed = f12_mul(f12_mul(plain, plain), plain);
    Fp12 chain = final_exp_hard_cubed(e);
    Fp6 d0 = f6_sub(cubed.c0, chain.c0), d1 = f6_sub(cubed.c1, chain.c1);
    if (!(d0.c0.is_zero() && d0.c1.is_zero() && d0.c2.is_zero() && d1.c0.is_zero() && d1.c1.is_zero() && d1.c2.is_zero())) bad++;
    if (f12_is_one(plain)) bad++;                       // e(G1, G2) is not degenerate
    // unitary after the easy part: e * conj(e) = 1
    if (!f12_is_one(f12_mul(e, f12_conj(e)))) bad++;
    // p^2-Frobenius twice more is the p^6-Frobenius = conjugation on Fp12
    Fp12 c = f12_frob2(f12_frob2(f12_frob2(f)));
    Fp12 cj = f12_conj(f);
    Fp6 e0 = f6_sub(c.c0, cj.c0), e1 = f6_sub(c.c1, cj.c1);
    if (!(e0.c0.is_zero() && e0.c1.is_zero() && e0.c2.is_zero() && e1.c0.is_zero() && e1.c1.is_zero() && e1.c2.is_zero())) bad++;
    // p-Frobenius twice is the p^2-Frobenius
    Fp12 u = f12_frob1(f12_frob1(f)), v = f12_frob2(f);
    Fp6 g0 = f6_sub(u.c0, v.c0), g1 = f6_sub(u.c1, v.c1);
    if (!(g0.c0.is_zero() && g0.c1.is_zero() && g0.c2.is_zero() && g1.c0.is_zero() && g1.c1.is_zero() && g1.c2.is_zero())) bad++;
    if (bad) MB_ATOMIC_ADD(a.mismatches, bad);
}

#ifdef MB_DEFINE_PAIR
MB_KERNEL_DEF(pair_selftest, PairSelfTestArgs, pair_selftest_body, 32)
MB_KERNEL_DEF(proof_read, ProofReadArgs, proof_read_body, 32)
MB_KERNEL_DEF(batch_miller, BatchMillerArgs, batch_miller_body, 32)
MB_KERNEL_DEF(batch_final, BatchFinalArgs, batch_final_body, 32)
MB_KERNEL_DEF(pair_prep, PairPrepArgs, pair_prep_body, 32)
MB_KERNEL_DEF(verify_proofs, VerifyArgs, verify_body, 32)
#else
MB_KERNEL_DECL(pair_selftest, PairSelfTestArgs)
MB_KERNEL_DECL(proof_read, ProofReadArgs)
MB_KERNEL_DECL(batch_miller, BatchMillerArgs)
MB_KERNEL_DECL(batch_final, BatchFinalArgs)
MB_KERNEL_DECL(pair_prep, PairPrepArgs)
MB_KERNEL_DECL(verify_proofs, VerifyArgs)
#endif

}  // namespace mb
