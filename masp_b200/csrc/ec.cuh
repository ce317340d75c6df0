// Short-Weierstrass arithmetic for BLS12-381 G1 (over Fp) and G2 (over Fp2),
// y^2 = x^3 + b with a = 0, in extended Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): the cheapest mixed addition (8M + 2S)
// for bucket accumulation, where one operand is always an affine key point.
//
// Replaces the G1/G2 add/double/mixed-add the reference takes from nam-blstrs
// (SURVEY.md §2 #3).  Every formula handles the exceptional inputs exactly
// (identity, P + P, P + (-P)), so the result is the same group element the
// reference computes for any input, not only for generic ones.
#pragma once
#include "field.cuh"

namespace mb {

// Affine point; the identity is encoded as x = y = 0 (not on either curve).
template <class F>
struct Affine {
    F x, y;
    MB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    MB_HD static Affine inf() { return {F::zero(), F::zero()}; }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;
    MB_HD bool is_inf() const { return zz.is_zero(); }
    MB_HD static XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    MB_HD static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return {p.x, p.y, F::one(), F::one()};
    }
};

template <class F>
MB_COLD XYZZ<F> xyzz_dbl_cold(const XYZZ<F>& p);
template <class F>
MB_COLD XYZZ<F> xyzz_dbl_affine_cold(const Affine<F>& p);

// 2 * (affine P)
template <class F>
MB_HD XYZZ<F> xyzz_dbl_affine(const Affine<F>& p) {
    // U = 2Y, V = U^2, W = U V, S = X V, M = 3 X^2
    if (p.is_inf() || p.y.is_zero()) return XYZZ<F>::inf();
    F U = F::dbl(p.y);
    F V = F::sqr(U);
    F W = F::mul(U, V);
    F S = F::mul(p.x, V);
    F X2 = F::sqr(p.x);
    F M = F::add(F::dbl(X2), X2);
    XYZZ<F> r;
    r.x = F::sub(F::sqr(M), F::dbl(S));
    r.y = F::sub(F::mul(M, F::sub(S, r.x)), F::mul(W, p.y));
    r.zz = V;
    r.zzz = W;
    return r;
}

// 2 * P
template <class F>
MB_HD XYZZ<F> xyzz_dbl(const XYZZ<F>& p) {
    if (p.is_inf() || p.y.is_zero()) return XYZZ<F>::inf();
    F U = F::dbl(p.y);
    F V = F::sqr(U);
    F W = F::mul(U, V);
    F S = F::mul(p.x, V);
    F X2 = F::sqr(p.x);
    F M = F::add(F::dbl(X2), X2);
    XYZZ<F> r;
    r.x = F::sub(F::sqr(M), F::dbl(S));
    r.y = F::sub(F::mul(M, F::sub(S, r.x)), F::mul(W, p.y));
    r.zz = F::mul(V, p.zz);
    r.zzz = F::mul(W, p.zzz);
    return r;
}

// Y3 = R (Q - X3) - Y1 PPP.  Over Fp (in units with inlined multiplications: the accumulate kernel) both
// products go through ONE Montgomery reduction, R T + (p - Y1) PPP: 144 of the ~2 900 wide multiplies of
// a mixed addition less.
template <class F>
MB_HD F madd_y3(const F& R, const F& T, const F& Y1, const F& PPP) {
    return F::sub(F::mul(R, T), F::mul(Y1, PPP));
}
#ifndef MB_COLD_MUL
template <>
MB_HD Fp madd_y3<Fp>(const Fp& R, const Fp& T, const Fp& Y1, const Fp& PPP) {
    return Fp::sop2_inline(R, T, Fp::neg(Y1), PPP);
}
#ifdef MB_FP2_SOP
// Fp2: c0 = R0 T0 - R1 T1 - Y0 P0 + Y1 P1, c1 = R0 T1 + R1 T0 - Y0 P1 - Y1 P0: eight products, two reductions
template <>
MB_HD Fp2 madd_y3<Fp2>(const Fp2& R, const Fp2& T, const Fp2& Y1, const Fp2& PPP) {
    const Fp nR1 = Fp::neg(R.c1), nY0 = Fp::neg(Y1.c0), nY1 = Fp::neg(Y1.c1);
    return {Fp::sop4_inline(R.c0, T.c0, nR1, T.c1, nY0, PPP.c0, Y1.c1, PPP.c1),
            Fp::sop4_inline(R.c0, T.c1, R.c1, T.c0, nY0, PPP.c1, nY1, PPP.c0)};
}
#endif
#endif

// acc += (affine q), q optionally negated
template <class F>
MB_HD void xyzz_madd(XYZZ<F>& acc, const Affine<F>& q_in, bool negate) {
    if (q_in.is_inf()) return;
    Affine<F> q = q_in;
    if (negate) q.y = F::neg(q.y);
    if (acc.is_inf()) {
        acc.x = q.x;
        acc.y = q.y;
        acc.zz = F::one();
        acc.zzz = F::one();
        return;
    }
    F U2 = F::mul(q.x, acc.zz);
    F S2 = F::mul(q.y, acc.zzz);
    F P = F::sub(U2, acc.x);
    F R = F::sub(S2, acc.y);
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl_affine_cold(q);
        else acc = XYZZ<F>::inf();
        return;
    }
    F PP = F::sqr(P);
    F PPP = F::mul(P, PP);
    F Q = F::mul(acc.x, PP);
    F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    F Y3 = madd_y3(R, F::sub(Q, X3), acc.y, PPP);
    acc.x = X3;
    acc.y = Y3;
    acc.zz = F::mul(acc.zz, PP);
    acc.zzz = F::mul(acc.zzz, PPP);
}

// acc += q
template <class F>
MB_HD void xyzz_add(XYZZ<F>& acc, const XYZZ<F>& q) {
    if (q.is_inf()) return;
    if (acc.is_inf()) {
        acc = q;
        return;
    }
    F U1 = F::mul(acc.x, q.zz);
    F U2 = F::mul(q.x, acc.zz);
    F S1 = F::mul(acc.y, q.zzz);
    F S2 = F::mul(q.y, acc.zzz);
    F P = F::sub(U2, U1);
    F R = F::sub(S2, S1);
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl_cold(acc);
        else acc = XYZZ<F>::inf();
        return;
    }
    F PP = F::sqr(P);
    F PPP = F::mul(P, PP);
    F Q = F::mul(U1, PP);
    F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    F Y3 = F::sub(F::mul(R, F::sub(Q, X3)), F::mul(S1, PPP));
    acc.x = X3;
    acc.y = Y3;
    acc.zz = F::mul(F::mul(acc.zz, q.zz), PP);
    acc.zzz = F::mul(F::mul(acc.zzz, q.zzz), PPP);
}

// Out-of-line copies for everything off the bucket-accumulation hot loop:
// one body per field instead of one per call site (keeps ptxas time and
// instruction-cache footprint down; the call overhead is noise next to the
// dozen multiplications inside).
template <class F>
MB_COLD void xyzz_add_cold(XYZZ<F>& acc, const XYZZ<F>& q) { xyzz_add(acc, q); }
template <class F>
MB_COLD void xyzz_madd_cold(XYZZ<F>& acc, const Affine<F>& q, bool negate) { xyzz_madd(acc, q, negate); }
template <class F>
MB_COLD XYZZ<F> xyzz_dbl_cold(const XYZZ<F>& p) { return xyzz_dbl(p); }
template <class F>
MB_COLD XYZZ<F> xyzz_dbl_affine_cold(const Affine<F>& p) { return xyzz_dbl_affine(p); }
template <class F>
MB_COLD F field_inv_cold(const F& a) { return F::inv(a); }

// affine image: x = X / ZZ, y = Y / ZZZ via one inversion of ZZ * ZZZ
template <class F>
MB_COLD Affine<F> xyzz_to_affine(const XYZZ<F>& p) {
    if (p.is_inf()) return Affine<F>::inf();
    F t = field_inv_cold(F::mul(p.zz, p.zzz));
    F izz = F::mul(t, p.zzz);
    F izzz = F::mul(t, p.zz);
    return {F::mul(p.x, izz), F::mul(p.y, izzz)};
}

// k * P for a 256-bit plain little-endian scalar (8 limbs): fixed 4-bit
// windows, MSB first (256 doublings + at most 64 additions, no divergence
// on the scalar's bits beyond skipping zero digits)
template <class F>
MB_COLD XYZZ<F> xyzz_mul(const XYZZ<F>& p, const uint32_t* k) {
    XYZZ<F> tab[16];
    tab[0] = XYZZ<F>::inf();
    tab[1] = p;
    MB_NOUNROLL
    for (int j = 2; j < 16; ++j) {
        tab[j] = tab[j - 1];
        xyzz_add_cold(tab[j], p);
    }
    XYZZ<F> acc = XYZZ<F>::inf();
    MB_NOUNROLL
    for (int w = 63; w >= 0; --w) {
        MB_NOUNROLL
        for (int d = 0; d < 4; ++d) acc = xyzz_dbl_cold(acc);
        uint32_t digit = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
        if (digit) xyzz_add_cold(acc, tab[digit]);
    }
    return acc;
}
template <class F>
MB_COLD XYZZ<F> xyzz_mul_affine(const Affine<F>& p, const uint32_t* k) {
    XYZZ<F> acc = XYZZ<F>::inf();
    MB_NOUNROLL
    for (int i = 255; i >= 0; --i) {
        acc = xyzz_dbl_cold(acc);
        if ((k[i >> 5] >> (i & 31)) & 1) xyzz_madd_cold(acc, p, false);
    }
    return acc;
}

// k * P on G1 with the GLV endomorphism phi(x, y) = (beta x, y) = lambda P, lambda = x^2 - 1
// (lambda^2 + lambda + 1 = 0 mod r, 128 bits): k = k1 + k2 lambda by plain Euclidean division
// (lambda is about sqrt(r), so both halves are below 2^128 without any lattice step), then
// k1 P + k2 phi(P) with shared doublings: 128 doublings + at most 64 additions instead of
// 256 + 64.  Used for the two scalar multiplications of the proof's C element (s A + r B1),
// which sit on the latency path of every proof.
MB_HD Fp glv_beta() {
    constexpr uint32_t t[12] = {0x8671f071u, 0xcd03c9e4u, 0x1fcda5d2u, 0x5dab2246u, 0xd3851b95u, 0x587042afu,
                                0x01bacb9eu, 0x8eb60ebeu, 0x83d050d2u, 0x03f97d6eu, 0x54638741u, 0x18f02065u};
    Fp r;
    for (int i = 0; i < 12; ++i) r.v[i] = t[i];
    return r;
}
// k (8 words, < 2^256) = q * lambda + rem, lambda = 0xac45a4010001a40200000000ffffffff
MB_HD void glv_split(const uint32_t* k, uint32_t rem[4], uint32_t q[4]) {
    const uint32_t lam[4] = {0xffffffffu, 0x00000000u, 0x0001a402u, 0xac45a401u};
    uint32_t r[5] = {0, 0, 0, 0, 0};  // running remainder, < 2 lambda < 2^129
    uint32_t qq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    MB_NOUNROLL
    for (int i = 255; i >= 0; --i) {
        // r = (r << 1) | bit
        r[4] = (r[4] << 1) | (r[3] >> 31);
        r[3] = (r[3] << 1) | (r[2] >> 31);
        r[2] = (r[2] << 1) | (r[1] >> 31);
        r[1] = (r[1] << 1) | (r[0] >> 31);
        r[0] = (r[0] << 1) | ((k[i >> 5] >> (i & 31)) & 1u);
        // if r >= lambda: r -= lambda, quotient bit = 1
        uint32_t d[5];
        uint64_t br = 0;
        for (int j = 0; j < 5; ++j) {
            uint64_t x = (uint64_t)r[j] - (j < 4 ? lam[j] : 0u) - br;
            d[j] = (uint32_t)x;
            br = (x >> 63) & 1;
        }
        if (!br) {
            for (int j = 0; j < 5; ++j) r[j] = d[j];
            qq[i >> 5] |= 1u << (i & 31);
        }
    }
    for (int j = 0; j < 4; ++j) {
        rem[j] = r[j];
        q[j] = qq[j];  // k < r < lambda^2 + ..., so the quotient fits 128 bits (checked by the caller's tests)
    }
}
MB_COLD XYZZ<Fp> xyzz_mul_glv(const XYZZ<Fp>& p, const uint32_t* k) {
    uint32_t k1[4], k2[4];
    glv_split(k, k1, k2);
    XYZZ<Fp> tab[16];
    tab[0] = XYZZ<Fp>::inf();
    tab[1] = p;
    MB_NOUNROLL
    for (int j = 2; j < 16; ++j) {
        tab[j] = tab[j - 1];
        xyzz_add_cold(tab[j], p);
    }
    const Fp beta = glv_beta();
    XYZZ<Fp> acc = XYZZ<Fp>::inf();
    MB_NOUNROLL
    for (int w = 31; w >= 0; --w) {
        MB_NOUNROLL
        for (int d = 0; d < 4; ++d) acc = xyzz_dbl_cold(acc);
        uint32_t d1 = (k1[w >> 3] >> ((w & 7) * 4)) & 15u, d2 = (k2[w >> 3] >> ((w & 7) * 4)) & 15u;
        if (d1) xyzz_add_cold(acc, tab[d1]);
        if (d2) {
            XYZZ<Fp> t = tab[d2];
            t.x = Fp::mul(t.x, beta);  // phi on XYZZ: x = X / ZZ
            xyzz_add_cold(acc, t);
        }
    }
    return acc;
}

typedef Affine<Fp> G1Affine;
typedef Affine<Fp2> G2Affine;
typedef XYZZ<Fp> G1XYZZ;
typedef XYZZ<Fp2> G2XYZZ;

// ---------------------------------------------------------------------------
// wire encodings; plain-integer big-endian bytes
// ---------------------------------------------------------------------------
// 48 big-endian bytes -> limbs (plain integer, not reduced, not Montgomery)
MB_HD void fp_limbs_from_be(const uint8_t* b, Fp& out) {
    for (int i = 0; i < 12; ++i) {
        const uint8_t* q = b + 4 * (11 - i);
        out.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
}
MB_HD void fp_limbs_to_be(const Fp& a, uint8_t* b) {
    for (int i = 0; i < 12; ++i) {
        uint8_t* q = b + 4 * (11 - i);
        q[0] = (uint8_t)(a.v[i] >> 24);
        q[1] = (uint8_t)(a.v[i] >> 16);
        q[2] = (uint8_t)(a.v[i] >> 8);
        q[3] = (uint8_t)a.v[i];
    }
}

// Uncompressed G1 (96 B).  Returns false when malformed: a flag bit other
// than "infinity", non-zero bytes in an infinity encoding, or a coordinate
// that is not canonical.  No curve / subgroup check: the reference reads
// with checked = false (masp_proofs/src/lib.rs:336-341).
MB_COLD bool g1_decode(const uint8_t* b, G1Affine& out) {
    uint32_t flags = b[0] >> 5;
    if (flags & 0x5) return false;
    if (flags & 0x2) {
        uint32_t o = b[0] & 0x1f;
        for (int i = 1; i < 96; ++i) o |= b[i];
        out = G1Affine::inf();
        return o == 0;
    }
    Fp x, y;
    fp_limbs_from_be(b, x);
    fp_limbs_from_be(b + 48, y);
    if (Fp::std_ge_mod(x) || Fp::std_ge_mod(y)) return false;
    out.x = Fp::from_std(x);
    out.y = Fp::from_std(y);
    return true;
}
MB_COLD bool g2_decode(const uint8_t* b, G2Affine& out) {
    uint32_t flags = b[0] >> 5;
    if (flags & 0x5) return false;
    if (flags & 0x2) {
        uint32_t o = b[0] & 0x1f;
        for (int i = 1; i < 192; ++i) o |= b[i];
        out = G2Affine::inf();
        return o == 0;
    }
    Fp xc1, xc0, yc1, yc0;
    fp_limbs_from_be(b, xc1);
    fp_limbs_from_be(b + 48, xc0);
    fp_limbs_from_be(b + 96, yc1);
    fp_limbs_from_be(b + 144, yc0);
    if (Fp::std_ge_mod(xc1) || Fp::std_ge_mod(xc0) || Fp::std_ge_mod(yc1) || Fp::std_ge_mod(yc0)) return false;
    out.x.c0 = Fp::from_std(xc0);
    out.x.c1 = Fp::from_std(xc1);
    out.y.c0 = Fp::from_std(yc0);
    out.y.c1 = Fp::from_std(yc1);
    return true;
}
MB_COLD void g1_encode(const G1Affine& p, uint8_t* b) {
    if (p.is_inf()) {
        for (int i = 0; i < 96; ++i) b[i] = 0;
        b[0] = 0x40;
        return;
    }
    fp_limbs_to_be(Fp::to_std(p.x), b);
    fp_limbs_to_be(Fp::to_std(p.y), b + 48);
}
MB_COLD void g2_encode(const G2Affine& p, uint8_t* b) {
    if (p.is_inf()) {
        for (int i = 0; i < 192; ++i) b[i] = 0;
        b[0] = 0x40;
        return;
    }
    fp_limbs_to_be(Fp::to_std(p.x.c1), b);
    fp_limbs_to_be(Fp::to_std(p.x.c0), b + 48);
    fp_limbs_to_be(Fp::to_std(p.y.c1), b + 96);
    fp_limbs_to_be(Fp::to_std(p.y.c0), b + 144);
}
// Compressed encodings: x with flag bits; sort flag iff y is the
// lexicographically larger of {y, -y} (G2: compare c1 first, then c0).
MB_COLD void g1_encode_compressed(const G1Affine& p, uint8_t* b) {
    if (p.is_inf()) {
        for (int i = 0; i < 48; ++i) b[i] = 0;
        b[0] = 0xc0;
        return;
    }
    fp_limbs_to_be(Fp::to_std(p.x), b);
    b[0] |= 0x80;
    if (Fp::std_gt_half(Fp::to_std(p.y))) b[0] |= 0x20;
}
MB_COLD void g2_encode_compressed(const G2Affine& p, uint8_t* b) {
    if (p.is_inf()) {
        for (int i = 0; i < 96; ++i) b[i] = 0;
        b[0] = 0xc0;
        return;
    }
    fp_limbs_to_be(Fp::to_std(p.x.c1), b);
    fp_limbs_to_be(Fp::to_std(p.x.c0), b + 48);
    b[0] |= 0x80;
    bool larger = p.y.c1.is_zero() ? Fp::std_gt_half(Fp::to_std(p.y.c0)) : Fp::std_gt_half(Fp::to_std(p.y.c1));
    if (larger) b[0] |= 0x20;
}

}  // namespace mb
