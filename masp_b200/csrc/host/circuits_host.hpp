// The MASP Spend / Output / Convert circuits on the host-side sink of
// r1cs_host.hpp, with native Jubjub arithmetic for the witness values.
//
// Follows, gadget by gadget and in the same emission order (the order fixes
// the variable numbering, hence which key element pairs with which value):
//   masp_proofs/src/circuit/ecc.rs          witness :130-143, interpret :250-276,
//       double :278-371, add :374-471, conditionally_select :147-198,
//       mul :203-248, repr :112-126, fixed_base_multiplication :27-73,
//       Montgomery add :543-617, into_edwards :483-531
//   masp_proofs/src/circuit/pedersen_hash.rs:19-103
//   masp_proofs/src/circuit/sapling.rs:71-137 (expose_value_commitment),
//       :139-417 (Spend::synthesize), :419-596 (Output::synthesize)
//   masp_proofs/src/circuit/convert.rs:29-128 (Convert::synthesize)
//   masp_proofs/src/constants.rs:10-38, 76-94, 100-173 (curve constants, window tables)
// Generator coordinates are data from masp_primitives/src/constants.rs:50-251.
#pragma once
#include "r1cs_host.hpp"

namespace mbh {

inline Fr fr_hex(const char* s) {  // 64 hex digits, big-endian
    uint64_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 64; ++i) {
        char ch = s[i];
        uint64_t d = ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10;
        int bit = (63 - i) * 4;
        w[bit >> 6] |= d << (bit & 63);
    }
    return Fr::from_words(w);
}

struct JPoint {
    Fr u, v;
};

struct Jubjub {
    Fr d, mont_a, mont_scale;
    JPoint proof_generation_key_generator, note_commitment_randomness_generator, nullifier_position_generator,
        value_commitment_randomness_generator, spending_key_generator;
    JPoint pedersen_generators[6];
    // circuit tables
    std::vector<Window8> fixed[5];  // 84 windows each; index = enum Fixed
    std::vector<Window4> pedersen[6];  // 63 windows each
    enum Fixed { PROOF_GENERATION_KEY = 0, NOTE_COMMITMENT_RANDOMNESS, NULLIFIER_POSITION, VALUE_COMMITMENT_RANDOMNESS, SPENDING_KEY };

    JPoint identity() const { return {Fr::zero(), Fr::one()}; }
    JPoint add(const JPoint& p, const JPoint& q) const {
        Fr uv = p.u * q.v, vu = p.v * q.u;
        Fr t = d * uv * vu;
        Fr one = Fr::one();
        Fr dp = one + t, dm = one - t;
        Fr i = (dp * dm).inverse();
        return {(uv + vu) * (i * dm), (p.v * q.v + p.u * q.u) * (i * dp)};
    }
    JPoint dbl(const JPoint& p) const { return add(p, p); }
    // ---- chains without per-step inversions ---------------------------------
    // The circuits walk long chains of point additions whose every intermediate
    // point is a witness value in affine form.  bellman's closures invert once
    // per step; here a chain is first run in projective coordinates and all its
    // denominators are inverted together (Montgomery's trick), which yields the
    // same affine values with one inversion per chain.
    struct Ext {  // extended twisted Edwards, a = -1: u = X/Z, v = Y/Z, T = XY/Z
        Fr X, Y, Z, T;
    };
    Fr d2;  // 2 d
    Ext ext_of(const JPoint& p) const { return {p.u, p.v, Fr::one(), p.u * p.v}; }
    Ext ext_add(const Ext& p, const Ext& q) const {
        Fr A = (p.Y - p.X) * (q.Y - q.X), B = (p.Y + p.X) * (q.Y + q.X);
        Fr C = p.T * d2 * q.T, D = (p.Z * q.Z).dbl();
        Fr E = B - A, F = D - C, G = D + C, H = B + A;
        return {E * F, G * H, F * G, E * H};
    }
    Ext ext_dbl(const Ext& p) const {
        Fr A = p.X.square(), B = p.Y.square(), C = p.Z.square().dbl();
        Fr D = -A;
        Fr E = (p.X + p.Y).square() - A - B, G = D + B;
        Fr F = G - C, H = D - B;
        return {E * F, G * H, F * G, E * H};
    }
    // v[i] <- 1 / v[i]; zeros stay zero and are reported
    static bool batch_inverse(Fr* v, size_t n, std::vector<Fr>& scratch) {
        bool any_zero = false;
        scratch.resize(n);
        Fr acc = Fr::one();
        for (size_t i = 0; i < n; ++i) {
            scratch[i] = acc;
            if (v[i].is_zero()) any_zero = true;
            else acc *= v[i];
        }
        acc = acc.inverse();
        for (size_t i = n; i-- > 0;) {
            if (v[i].is_zero()) continue;
            Fr t = acc * scratch[i];
            acc *= v[i];
            v[i] = t;
        }
        return any_zero;
    }
    void ext_to_affine(const Ext* in, JPoint* out, size_t n) const {
        std::vector<Fr> z(n), scratch;
        for (size_t i = 0; i < n; ++i) z[i] = in[i].Z;
        batch_inverse(z.data(), n, scratch);
        for (size_t i = 0; i < n; ++i) out[i] = {in[i].X * z[i], in[i].Y * z[i]};
    }
    // sums[k] = pts[0] + ... + pts[k]
    void chain_sums(const JPoint* pts, size_t n, JPoint* sums) const {
        if (!n) return;
        std::vector<Ext> e(n);
        e[0] = ext_of(pts[0]);
        for (size_t k = 1; k < n; ++k) e[k] = ext_add(e[k - 1], ext_of(pts[k]));
        ext_to_affine(e.data(), sums, n);
    }
    // Montgomery chain S_0 = T_0, S_k = T_k + S_(k-1): the slopes lam[k] (k >= 1) of each step.
    // The running sum is kept projective, (X : Y : Z); a step's slope is u / v with
    // u = ty Z - Y, v = tx Z - X, so only the v's have to be inverted, all together.
    // Returns false if a step is exceptional (equal x: the reference's DivisionByZero).
    bool montgomery_chain_slopes(const Fr* tx, const Fr* ty, size_t n, Fr* lam) const {
        if (n < 2) return true;
        std::vector<Fr> u(n), v(n), scratch;
        Fr X = tx[0], Y = ty[0], Z = Fr::one();
        v[0] = Fr::one();
        for (size_t k = 1; k < n; ++k) {  // projective chord: (X:Y:Z) + affine (tx, ty)
            Fr x2z = tx[k] * Z;
            Fr uu = ty[k] * Z - Y, vv1 = x2z - X;
            u[k] = uu;
            v[k] = vv1;
            if (k + 1 == n) break;  // the last sum itself is not needed
            Fr vv = vv1.square(), vvv = vv * vv1;
            Fr W = uu.square() * Z - vv * (mont_a * Z + x2z + X);
            Fr Xn = W * vv1;
            Y = uu * (X * vv - W) - Y * vvv;
            X = Xn;
            Z = vvv * Z;
        }
        bool bad = batch_inverse(v.data(), n, scratch);
        for (size_t k = 1; k < n; ++k) lam[k] = u[k] * v[k];
        return !bad;
    }

    // Everything a Pedersen hash needs from the field's division, for all of its segments at once:
    // per window the slope of its chain step, per segment the Edwards image of the segment sum and
    // the running Edwards sum.  Three inversions per hash (all slope denominators and segment Z's;
    // the Montgomery -> Edwards denominators; the running sums), however many segments it has.
    struct PedersenPlan {
        std::vector<Fr> lam;         // per window (first window of a segment: unused)
        std::vector<JPoint> seg_ed;  // per segment: into_edwards of the segment's sum
        std::vector<JPoint> run_ed;  // per segment: seg_ed[0] + ... + seg_ed[s]
        bool ok = true;
    };
    // tx, ty: the window points of the whole hash; seg_len[s] windows belong to segment s
    void pedersen_plan(const std::vector<Fr>& tx, const std::vector<Fr>& ty, const std::vector<size_t>& seg_len,
                       PedersenPlan& plan) const {
        const size_t nw = tx.size(), ns = seg_len.size();
        std::vector<Fr> u(nw), den(nw + ns), scratch;  // den: v per window, then Z per segment
        std::vector<Fr> fx(ns), fy(ns);                // projective numerators of the segment sums
        size_t base = 0;
        for (size_t sgm = 0; sgm < ns; ++sgm) {
            size_t n = seg_len[sgm];
            Fr X = tx[base], Y = ty[base], Z = Fr::one();
            den[base] = Fr::one();
            for (size_t k = 1; k < n; ++k) {
                const Fr &x2 = tx[base + k], &y2 = ty[base + k];
                Fr x2z = x2 * Z;
                Fr uu = y2 * Z - Y, v1 = x2z - X;
                u[base + k] = uu;
                den[base + k] = v1;
                Fr vv = v1.square(), vvv = vv * v1;
                Fr W = uu.square() * Z - vv * (mont_a * Z + x2z + X);
                Fr Xn = W * v1;
                Y = uu * (X * vv - W) - Y * vvv;
                X = Xn;
                Z = vvv * Z;
            }
            fx[sgm] = X;
            fy[sgm] = Y;
            den[nw + sgm] = Z;
            base += n;
        }
        if (batch_inverse(den.data(), nw + ns, scratch)) plan.ok = false;
        plan.lam.assign(nw, Fr::zero());
        base = 0;
        for (size_t sgm = 0; sgm < ns; ++sgm) {
            for (size_t k = 1; k < seg_len[sgm]; ++k) plan.lam[base + k] = u[base + k] * den[base + k];
            base += seg_len[sgm];
        }
        // Montgomery -> Edwards for every segment sum: u = scale x / y, v = (x - 1) / (x + 1)
        std::vector<Fr> sx(ns), sy(ns), d2(ns);
        Fr one = Fr::one();
        for (size_t sgm = 0; sgm < ns; ++sgm) {
            sx[sgm] = fx[sgm] * den[nw + sgm];
            sy[sgm] = fy[sgm] * den[nw + sgm];
            d2[sgm] = sy[sgm] * (sx[sgm] + one);
        }
        if (batch_inverse(d2.data(), ns, scratch)) plan.ok = false;
        plan.seg_ed.resize(ns);
        for (size_t sgm = 0; sgm < ns; ++sgm) {
            Fr xp = sx[sgm] + one;
            plan.seg_ed[sgm] = {sx[sgm] * mont_scale * (d2[sgm] * xp), (sx[sgm] - one) * (d2[sgm] * sy[sgm])};
        }
        plan.run_ed.resize(ns);
        chain_sums(plan.seg_ed.data(), ns, plan.run_ed.data());
    }

    bool to_montgomery(const JPoint& p, Fr& x, Fr& y) const {  // constants.rs:100-141
        Fr one = Fr::one();
        if (p.v == one) return false;
        if (p.u.is_zero()) {
            x = Fr::zero();
            y = Fr::zero();
            return true;
        }
        x = (one + p.v) * (one - p.v).inverse();
        y = x * p.u.inverse() * mont_scale;
        return true;
    }

    Jubjub() {
        d = fr_hex("2a9318e74bfa2b48f5fd9207e6bd7fd4292d7f6d37579d2601065fd6d6343eb1");
        d2 = d.dbl();
        mont_a = Fr::from_u64(0xa002);
        mont_scale = fr_hex("2762de61e862645e31de341e77d764e5ce4069703da88abd8f4535f7cf82b8d9");
        proof_generation_key_generator = {fr_hex("4caaeacaaf28ed4b4ba1f065e719fd031e24f83267f15abd5f3c723aa2531b66"),
                                          fr_hex("00930d67d6906365c654dfdd36004de936b49c71a2af0708fe6f96bec575bff8")};
        note_commitment_randomness_generator = {fr_hex("434c9be15267b091c6de7556abb84082cd80edf5fe44c7bffc033fa2bf88cb2e"),
                                                fr_hex("29e2926993d3bc736d277197e97af8f0690b295c66b85c64c6b8daa0ee22aeed")};
        nullifier_position_generator = {fr_hex("50de6d98fee5282f84678dc2d85293df1e09674f28a4b844aafee844265fc1e7"),
                                        fr_hex("03260f0bf1244050f3f70dc31afe799d226945aee96dfe0aed034e3ee13a1eb3")};
        value_commitment_randomness_generator = {fr_hex("1c6da0ce9a5e5fdbcfa86026b8d99be991cc3e3835675450dd93d364cb8cec7e"),
                                                 fr_hex("555f11f9b720d50bbc900cd4b8ae1150f94c2daa360302fe28e5fce99ce692d0")};
        spending_key_generator = {fr_hex("5b389522a9e81532f831c2b19fec602639f5b03380af6020ec75293d81248452"),
                                  fr_hex("0cbc5f9f1e52e0ab75defecff1f49ef22012d031f624fd5214b62623a186b4b1")};
        const char* ph[6][2] = {
            {"113de62be6e0d32398ba470b0d28801b5c22a82a281c91811010503570c3ebf6", "5059678472abb6ae15cea14bc9f6b04b2ba3032d7064d633f031edff274efb14"},
            {"08c02a4c57f7f2cffc7cbea3c311f67f0a0df10182a290fdb9efa2cb80331936", "2e560a50271fd3fc4dc07857131f22a0ec376560c925452ddaf19ac3ab182662"},
            {"210f22d61b65767d413bc3c44e7aabe0df0694e57c6cbc03c93573b98709291e", "3f46b3371cff7474fb33884c42727482c6262ed4231796594781e2656b1ddaad"},
            {"274e99b16d4af911a02f0d3f7aad771d2bcc52dbba0ebf3acf0bc7224a63d094", "31f5e34f0804a8746b15ec6e59478694fd0153cfe15ec653e82e9061620a1df4"},
            {"3ca8b98873e5d19e50aa77ad2f57d2f77058160b9afaafafc64e25ca51961b53", "10609ce821a5a292238af7c9376608d65eb152c4606beb7e9dab539b32327842"},
            {"1ab3fe2ac6b3ff8adb3ff866eaf1bc855bdd5c30d83781f0f0ef2a816469118e", "2031e442c4af8277d5681f2f5c740d19a6b5863148627619e7c079b4e48233f5"}};
        for (int i = 0; i < 6; ++i) pedersen_generators[i] = {fr_hex(ph[i][0]), fr_hex(ph[i][1])};

        // 3-bit window tables [0, g, ..., 7g], 84 windows 8x apart (constants.rs:76-94)
        const JPoint* gens[5] = {&proof_generation_key_generator, &note_commitment_randomness_generator,
                                 &nullifier_position_generator, &value_commitment_randomness_generator,
                                 &spending_key_generator};
        for (int k = 0; k < 5; ++k) {
            JPoint g0 = *gens[k];
            for (int w = 0; w < 84; ++w) {
                Window8 win;
                JPoint g = g0;
                win.u[0] = Fr::zero();
                win.v[0] = Fr::one();
                for (int j = 1; j < 8; ++j) {
                    win.u[j] = g.u;
                    win.v[j] = g.v;
                    g = add(g, g0);
                }
                synth_coeffs(3, win.u, win.uc);
                synth_coeffs(3, win.v, win.vc);
                fixed[k].push_back(win);
                g0 = g;  // 8 * g0
            }
        }
        // 2-bit window tables [g, 2g, 3g, 4g] in Montgomery form, 63 windows 16x apart (constants.rs:143-173)
        for (int k = 0; k < 6; ++k) {
            JPoint g0 = pedersen_generators[k];
            for (int w = 0; w < 63; ++w) {
                Window4 win;
                JPoint g = g0;
                for (int j = 0; j < 4; ++j) {
                    to_montgomery(g, win.x[j], win.y[j]);
                    g = add(g, g0);
                }
                synth_coeffs(2, win.x, win.xc);
                synth_coeffs(2, win.y, win.yc);
                pedersen[k].push_back(win);
                for (int j = 0; j < 4; ++j) g0 = dbl(g0);
            }
        }
    }
};
inline const Jubjub& JJ() {
    static const Jubjub j;
    return j;
}

// ---------------------------------------------------------------------------
// ecc gadgets
// ---------------------------------------------------------------------------
struct EdwardsPoint {
    AllocatedNum u, v;

    static EdwardsPoint interpret(CS& cs, const AllocatedNum& u, const AllocatedNum& v) {
        AllocatedNum u2 = u.square(cs), v2 = v.square(cs);
        AllocatedNum u2v2 = u2.mul(cs, v2);
        MBH_ENFORCE(cs, LC(u2.var, K().minus_one).add(v2.var), LC(ONE, K().one), LC(ONE, K().one).add(u2v2.var, JJ().d));
        return {u, v};
    }
    static EdwardsPoint witness(CS& cs, const JPoint& p) {
        AllocatedNum u = AllocatedNum::alloc(cs, p.u);
        AllocatedNum v = AllocatedNum::alloc(cs, p.v);
        return interpret(cs, u, v);
    }
    void inputize(CS& cs) const {
        u.inputize(cs);
        v.inputize(cs);
    }
    Bits repr(CS& cs) const {
        Bits ub = u.to_bits_le_strict(cs);
        Bits vb = v.to_bits_le_strict(cs);
        vb.push_back(ub[0]);
        return vb;
    }
    // `hint`: the result in affine form when the caller has already computed it (chain pre-pass)
    EdwardsPoint dbl(CS& cs, const JPoint* hint = nullptr) const {
        const Fr& d = JJ().d;
        Fr s = u.value + v.value;
        AllocatedNum t = AllocatedNum::alloc(cs, s.square());
        MBH_ENFORCE(cs, LC(u.var, K().one).add(v.var), LC(u.var, K().one).add(v.var), LC(t.var, K().one));
        AllocatedNum a = u.mul(cs, v);
        AllocatedNum c = AllocatedNum::alloc(cs, a.value.square() * d);
        MBH_ENFORCE(cs, LC(a.var, d), LC(a.var, K().one), LC(c.var, K().one));
        JPoint r;
        if (hint) {
            r = *hint;
        } else {
            Fr one = Fr::one();
            Fr dp = one + c.value, dm = one - c.value;
            Fr i = (dp * dm).inverse();
            if (i.is_zero()) cs.failed = true;
            Fr a2 = a.value.dbl();
            r = {a2 * (i * dm), (t.value - a2) * (i * dp)};
        }
        AllocatedNum u3 = AllocatedNum::alloc(cs, r.u);
        MBH_ENFORCE(cs, LC(ONE, K().one).add(c.var), LC(u3.var, K().one), LC(a.var, K().two));
        AllocatedNum v3 = AllocatedNum::alloc(cs, r.v);
        MBH_ENFORCE(cs, LC(ONE, K().one).sub(c.var), LC(v3.var, K().one), LC(t.var, K().one).add(a.var, -K().two));
        return {u3, v3};
    }
    EdwardsPoint add(CS& cs, const EdwardsPoint& o, const JPoint* hint = nullptr) const {
        const Fr& d = JJ().d;
        AllocatedNum big_u = AllocatedNum::alloc(cs, (u.value + v.value) * (o.u.value + o.v.value));
        MBH_ENFORCE(cs, LC(u.var, K().one).add(v.var), LC(o.u.var, K().one).add(o.v.var), LC(big_u.var, K().one));
        AllocatedNum a = o.v.mul(cs, u);
        AllocatedNum b = o.u.mul(cs, v);
        AllocatedNum c = AllocatedNum::alloc(cs, a.value * b.value * d);
        MBH_ENFORCE(cs, LC(a.var, d), LC(b.var, K().one), LC(c.var, K().one));
        JPoint r;
        if (hint) {
            r = *hint;
        } else {
            Fr one = Fr::one();
            Fr dp = one + c.value, dm = one - c.value;
            Fr i = (dp * dm).inverse();
            if (i.is_zero()) cs.failed = true;
            r = {(a.value + b.value) * (i * dm), (big_u.value - a.value - b.value) * (i * dp)};
        }
        AllocatedNum u3 = AllocatedNum::alloc(cs, r.u);
        MBH_ENFORCE(cs, LC(ONE, K().one).add(c.var), LC(u3.var, K().one), LC(a.var, K().one).add(b.var));
        AllocatedNum v3 = AllocatedNum::alloc(cs, r.v);
        MBH_ENFORCE(cs, LC(ONE, K().one).sub(c.var), LC(v3.var, K().one), LC(big_u.var, K().one).sub(a.var).sub(b.var));
        return {u3, v3};
    }
    EdwardsPoint conditionally_select(CS& cs, const Boolean& cond) const {
        AllocatedNum up = AllocatedNum::alloc(cs, cond.value() ? u.value : Fr::zero());
        MBH_ENFORCE(cs, LC(u.var, K().one), cond.lc(K().one), LC(up.var, K().one));
        AllocatedNum vp = AllocatedNum::alloc(cs, cond.value() ? v.value : Fr::one());
        LC c(vp.var, K().one);
        c.sub(cond.not_().lc(K().one));
        MBH_ENFORCE(cs, LC(v.var, K().one), cond.lc(K().one), c);
        return {up, vp};
    }
    EdwardsPoint mul(CS& cs, const Bits& by) const {
        // pre-pass: every 2^i P and every running sum, one inversion for all of them
        const Jubjub& J = JJ();
        size_t n = by.size();
        std::vector<Jubjub::Ext> e(2 * n);
        std::vector<JPoint> aff(2 * n);
        if (n) {
            e[0] = J.ext_of({u.value, v.value});
            for (size_t i = 1; i < n; ++i) e[i] = J.ext_dbl(e[i - 1]);
            Jubjub::Ext id = J.ext_of(J.identity());
            e[n] = by[0].value() ? e[0] : id;
            for (size_t i = 1; i < n; ++i) e[n + i] = by[i].value() ? J.ext_add(e[n + i - 1], e[i]) : e[n + i - 1];
            J.ext_to_affine(e.data(), aff.data(), 2 * n);
        }
        EdwardsPoint curbase = *this, result = *this;
        for (size_t i = 0; i < n; ++i) {
            if (i) curbase = curbase.dbl(cs, &aff[i]);
            EdwardsPoint thisbase = curbase.conditionally_select(cs, by[i]);
            result = i ? result.add(cs, thisbase, &aff[n + i]) : thisbase;
        }
        return result;
    }
};

inline EdwardsPoint fixed_base_multiplication(CS& cs, Jubjub::Fixed gen, const Bits& by) {
    const std::vector<Window8>& table = JJ().fixed[gen];
    EdwardsPoint result;
    Boolean f = Boolean::constant(false);
    // pre-pass: the looked-up points and their running sums
    size_t nw = (by.size() + 2) / 3;
    std::vector<JPoint> pts(nw), sums(nw);
    for (size_t k = 0; k < nw; ++k) {
        size_t i = 3 * k;
        int idx = (by[i].value() ? 1 : 0) | (i + 1 < by.size() && by[i + 1].value() ? 2 : 0) |
                  (i + 2 < by.size() && by[i + 2].value() ? 4 : 0);
        pts[k] = {table[k].u[idx], table[k].v[idx]};
    }
    JJ().chain_sums(pts.data(), nw, sums.data());
    for (size_t i = 0; i < by.size(); i += 3) {
        Boolean chunk[3] = {by[i], i + 1 < by.size() ? by[i + 1] : f, i + 2 < by.size() ? by[i + 2] : f};
        EdwardsPoint p;
        lookup3_xy(cs, chunk, table[i / 3], p.u, p.v);
        result = i ? result.add(cs, p, &sums[i / 3]) : p;
    }
    return result;
}

struct MontgomeryPoint {
    Num x, y;
    EdwardsPoint into_edwards(CS& cs, const JPoint* hint = nullptr) const {
        JPoint r;
        if (hint) {
            r = *hint;
        } else {
            Fr one = Fr::one();
            // one inversion for 1 / y and 1 / (x + 1)
            Fr xp = x.value + one;
            Fr i = (y.value * xp).inverse();
            if (i.is_zero()) cs.failed = true;
            r = {x.value * JJ().mont_scale * (i * xp), (x.value - one) * (i * y.value)};
        }
        AllocatedNum u = AllocatedNum::alloc(cs, r.u);
        MBH_ENFORCE(cs, y.lc, LC(u.var, K().one), x.lc.scaled(JJ().mont_scale));
        AllocatedNum v = AllocatedNum::alloc(cs, r.v);
        LC a = x.lc, c = x.lc;
        a.add(ONE, K().one);
        c.add(ONE, K().minus_one);
        MBH_ENFORCE(cs, a, LC(v.var, K().one), c);
        return {u, v};
    }
    MontgomeryPoint add(CS& cs, const MontgomeryPoint& o, const Fr* lam_hint = nullptr) const {
        Fr lv;
        if (lam_hint) {
            lv = *lam_hint;
        } else {
            Fr i = (o.x.value - x.value).inverse();
            if (i.is_zero()) cs.failed = true;
            lv = (o.y.value - y.value) * i;
        }
        AllocatedNum lam = AllocatedNum::alloc(cs, lv);
        {
            LC a = o.x.lc, c = o.y.lc;
            a.sub(x.lc);
            c.sub(y.lc);
            MBH_ENFORCE(cs, a, LC(lam.var, K().one), c);
        }
        AllocatedNum xprime = AllocatedNum::alloc(cs, lam.value.square() - JJ().mont_a - x.value - o.x.value);
        {
            LC c(ONE, JJ().mont_a);
            c.add(x.lc).add(o.x.lc).add(xprime.var);
            MBH_ENFORCE(cs, LC(lam.var, K().one), LC(lam.var, K().one), c);
        }
        AllocatedNum yprime = AllocatedNum::alloc(cs, -((xprime.value - x.value) * lam.value + y.value));
        {
            LC a = x.lc, c(yprime.var, K().one);
            a.sub(xprime.var);
            c.add(y.lc);
            MBH_ENFORCE(cs, a, LC(lam.var, K().one), c);
        }
        return {Num::from_allocated(xprime), Num::from_allocated(yprime)};
    }
};

inline EdwardsPoint pedersen_hash(CS& cs, const bool personalization[6], const Bits& bits) {
    Bits all;
    for (int i = 0; i < 6; ++i) all.push_back(Boolean::constant(personalization[i]));
    all.insert(all.end(), bits.begin(), bits.end());
    Boolean f = Boolean::constant(false);
    // pre-pass over the whole hash: window points from the tables, then every division it needs
    const size_t nw = (all.size() + 2) / 3;
    std::vector<Fr> tx(nw), ty(nw);
    std::vector<size_t> seg_len;
    for (size_t k = 0; k < nw; ++k) {
        size_t seg = k / 63, w = k % 63, q = 3 * k;
        if (w == 0) seg_len.push_back(0);
        ++seg_len.back();
        const Window4& win = JJ().pedersen[seg][w];
        int idx = (all[q].value() ? 1 : 0) | (q + 1 < all.size() && all[q + 1].value() ? 2 : 0);
        tx[k] = win.x[idx];
        ty[k] = (q + 2 < all.size() && all[q + 2].value()) ? -win.y[idx] : win.y[idx];
    }
    Jubjub::PedersenPlan plan;
    JJ().pedersen_plan(tx, ty, seg_len, plan);
    if (!plan.ok) cs.failed = true;

    EdwardsPoint edwards_result;
    size_t pos = 0, k = 0;
    for (size_t seg = 0; seg < seg_len.size(); ++seg) {
        MontgomeryPoint segment_result;
        const std::vector<Window4>& windows = JJ().pedersen[seg];
        for (size_t w = 0; w < seg_len[seg]; ++w, ++k) {
            Boolean chunk[3] = {all[pos], pos + 1 < all.size() ? all[pos + 1] : f, pos + 2 < all.size() ? all[pos + 2] : f};
            pos += 3;
            MontgomeryPoint tmp;
            lookup3_xy_with_conditional_negation(cs, chunk, windows[w], tmp.x, tmp.y);
            segment_result = w ? tmp.add(cs, segment_result, &plan.lam[k]) : tmp;
        }
        EdwardsPoint se = segment_result.into_edwards(cs, &plan.seg_ed[seg]);
        edwards_result = seg ? se.add(cs, edwards_result, &plan.run_ed[seg]) : se;
    }
    return edwards_result;
}

static const bool NOTE_COMMITMENT_PERSONALIZATION[6] = {true, true, true, true, true, true};
inline void merkle_personalization(int depth, bool out[6]) {
    for (int i = 0; i < 6; ++i) out[i] = (depth >> i) & 1;
}

// ---------------------------------------------------------------------------
// circuits
// ---------------------------------------------------------------------------
struct AuthNode {
    Fr sibling;
    bool is_right;
};

// sapling.rs:71-137
inline void expose_value_commitment(CS& cs, const JPoint& asset_generator, uint64_t value, const uint64_t rcv[4],
                                    Bits& asset_generator_bits, Bits& value_bits) {
    EdwardsPoint ag = EdwardsPoint::witness(cs, asset_generator);
    asset_generator_bits = ag.repr(cs);
    ag = ag.dbl(cs);
    ag = ag.dbl(cs);
    ag = ag.dbl(cs);
    ag.u.assert_nonzero(cs);
    value_bits = u64_into_boolean_vec_le(cs, value);
    EdwardsPoint val = ag.mul(cs, value_bits);
    Bits rcv_bits = words_into_boolean_vec_le(cs, rcv, 252);
    EdwardsPoint rcv_p = fixed_base_multiplication(cs, Jubjub::VALUE_COMMITMENT_RANDOMNESS, rcv_bits);
    EdwardsPoint cv = val.add(cs, rcv_p);
    cv.inputize(cs);
}

inline Num value_num_of(const Bits& value_bits) {
    Num n;
    for (size_t i = 0; i < value_bits.size(); ++i) n.add_bool_with_coeff(value_bits[i], K().pow2[i]);
    return n;
}

inline void assert_not_small_order(CS& cs, const EdwardsPoint& p) {
    EdwardsPoint t = p.dbl(cs);
    t = t.dbl(cs);
    t = t.dbl(cs);
    t.u.assert_nonzero(cs);
}

inline Bits merkle_and_anchor(CS& cs, AllocatedNum cur, const std::vector<AuthNode>& path, const Fr& anchor,
                              const Num& value_num) {
    Bits position_bits;
    for (size_t i = 0; i < path.size(); ++i) {
        Boolean cur_is_right = Boolean::from_bit(AllocatedBit::alloc(cs, path[i].is_right));
        position_bits.push_back(cur_is_right);
        AllocatedNum path_element = AllocatedNum::alloc(cs, path[i].sibling);
        AllocatedNum ul, ur;
        AllocatedNum::conditionally_reverse(cs, cur, path_element, cur_is_right, ul, ur);
        Bits preimage = ul.to_bits_le(cs);
        Bits r = ur.to_bits_le(cs);
        preimage.insert(preimage.end(), r.begin(), r.end());
        bool pers[6];
        merkle_personalization((int)i, pers);
        cur = pedersen_hash(cs, pers, preimage).u;
    }
    cs.root = cur.value;
    AllocatedNum rt = AllocatedNum::alloc(cs, anchor);
    MBH_ENFORCE(cs, LC(cur.var, K().one).sub(rt.var), value_num.lc, LC());
    rt.inputize(cs);
    return position_bits;
}

struct ConvertWitness {
    JPoint asset_generator;
    uint64_t value;
    uint64_t rcv[4];
    Fr anchor;
    std::vector<AuthNode> path;
};
inline void convert_circuit(CS& cs, const ConvertWitness& w) {
    Bits ag_bits, value_bits;
    expose_value_commitment(cs, w.asset_generator, w.value, w.rcv, ag_bits, value_bits);
    Num value_num = value_num_of(value_bits);
    EdwardsPoint cm = pedersen_hash(cs, NOTE_COMMITMENT_PERSONALIZATION, ag_bits);
    merkle_and_anchor(cs, cm.u, w.path, w.anchor, value_num);
}

struct SpendWitness {
    JPoint ak, g_d, asset_generator;
    uint64_t nsk[4], rcv[4], rcm[4], ar[4];
    uint64_t value;
    Fr anchor;
    std::vector<AuthNode> path;
};
inline void spend_circuit(CS& cs, const SpendWitness& w) {
    EdwardsPoint ak = EdwardsPoint::witness(cs, w.ak);
    assert_not_small_order(cs, ak);
    Bits ar_bits = words_into_boolean_vec_le(cs, w.ar, 252);
    EdwardsPoint ar_p = fixed_base_multiplication(cs, Jubjub::SPENDING_KEY, ar_bits);
    EdwardsPoint rk = ak.add(cs, ar_p);
    rk.inputize(cs);
    Bits nsk_bits = words_into_boolean_vec_le(cs, w.nsk, 252);
    EdwardsPoint nk = fixed_base_multiplication(cs, Jubjub::PROOF_GENERATION_KEY, nsk_bits);
    Bits ivk_preimage = ak.repr(cs);
    Bits repr_nk = nk.repr(cs);
    ivk_preimage.insert(ivk_preimage.end(), repr_nk.begin(), repr_nk.end());
    Bits nf_preimage = repr_nk;
    Bits ivk = blake2s(cs, ivk_preimage, "MASP_ivk");
    ivk.resize(251);
    EdwardsPoint g_d = EdwardsPoint::witness(cs, w.g_d);
    assert_not_small_order(cs, g_d);
    EdwardsPoint pk_d = g_d.mul(cs, ivk);
    Bits ag_bits, value_bits;
    expose_value_commitment(cs, w.asset_generator, w.value, w.rcv, ag_bits, value_bits);
    Num value_num = value_num_of(value_bits);
    Bits note = ag_bits;
    note.insert(note.end(), value_bits.begin(), value_bits.end());
    Bits t = g_d.repr(cs);
    note.insert(note.end(), t.begin(), t.end());
    t = pk_d.repr(cs);
    note.insert(note.end(), t.begin(), t.end());
    EdwardsPoint cm = pedersen_hash(cs, NOTE_COMMITMENT_PERSONALIZATION, note);
    Bits rcm_bits = words_into_boolean_vec_le(cs, w.rcm, 252);
    EdwardsPoint rcm_p = fixed_base_multiplication(cs, Jubjub::NOTE_COMMITMENT_RANDOMNESS, rcm_bits);
    cm = cm.add(cs, rcm_p);
    Bits position_bits = merkle_and_anchor(cs, cm.u, w.path, w.anchor, value_num);
    EdwardsPoint position = fixed_base_multiplication(cs, Jubjub::NULLIFIER_POSITION, position_bits);
    EdwardsPoint rho = cm.add(cs, position);
    t = rho.repr(cs);
    nf_preimage.insert(nf_preimage.end(), t.begin(), t.end());
    Bits nf = blake2s(cs, nf_preimage, "MASP__nf");
    pack_into_inputs(cs, nf);
}

struct OutputWitness {
    uint8_t asset_identifier[32];
    JPoint asset_generator, g_d, pk_d;
    uint64_t rcv[4], rcm[4], esk[4];
    uint64_t value;
};
inline void output_circuit(CS& cs, const OutputWitness& w) {
    Bits preimage;
    for (int i = 0; i < 256; ++i)
        preimage.push_back(Boolean::from_bit(AllocatedBit::alloc(cs, (w.asset_identifier[i >> 3] >> (i & 7)) & 1)));
    Bits image = blake2s(cs, preimage, "MASP__v_");
    Bits ag_bits, value_bits;
    expose_value_commitment(cs, w.asset_generator, w.value, w.rcv, ag_bits, value_bits);
    for (int i = 0; i < 256; ++i) Boolean::enforce_equal(cs, ag_bits[i], image[i]);
    Bits note = ag_bits;
    note.insert(note.end(), value_bits.begin(), value_bits.end());
    EdwardsPoint g_d = EdwardsPoint::witness(cs, w.g_d);
    assert_not_small_order(cs, g_d);
    Bits t = g_d.repr(cs);
    note.insert(note.end(), t.begin(), t.end());
    Bits esk_bits = words_into_boolean_vec_le(cs, w.esk, 252);
    EdwardsPoint epk = g_d.mul(cs, esk_bits);
    epk.inputize(cs);
    uint64_t pkv[4], pku[4];
    w.pk_d.v.to_words(pkv);
    w.pk_d.u.to_words(pku);
    Bits v_contents = words_into_boolean_vec_le(cs, pkv, 255);
    Boolean sign_bit = Boolean::from_bit(AllocatedBit::alloc(cs, pku[0] & 1));
    note.insert(note.end(), v_contents.begin(), v_contents.end());
    note.push_back(sign_bit);
    EdwardsPoint cm = pedersen_hash(cs, NOTE_COMMITMENT_PERSONALIZATION, note);
    Bits rcm_bits = words_into_boolean_vec_le(cs, w.rcm, 252);
    EdwardsPoint rcm_p = fixed_base_multiplication(cs, Jubjub::NOTE_COMMITMENT_RANDOMNESS, rcm_bits);
    cm = cm.add(cs, rcm_p);
    cm.u.inputize(cs);
}

}  // namespace mbh
