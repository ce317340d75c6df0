// Constraint-system sink, the bellman gadget set and the three MASP circuits for the host side of
// the proving path (SURVEY.md §8 a-2 / NEXT-1), written once over a value POLICY:
//
//   ScalarPolicy  one witness at a time on the 4 x 64-bit field of fr_host.hpp.  It also RECORDS:
//                 in a shape pass (recording() = true) every enforce() is stored as three sparse
//                 rows, which give the matrices for the GPU (r1cs_eval kernel), the density bitmaps
//                 and the structural hash;
//   SimdPolicy    (fr8_host.hpp, csrc/circuits_simd.cpp) eight witnesses side by side, one AVX-512
//                 lane each: field values are Fr8, booleans lane masks, every witness-dependent
//                 choice a select.  The circuits' structure does not depend on witness values, so
//                 the eight lanes walk the same allocation sequence; linear combinations are a null
//                 type and cost nothing.
//
// In a witness pass only alloc() / alloc_input() values are produced: the per-row evaluations
// <A_i,z>, <B_i,z>, <C_i,z> that bellman's ProvingAssignment computes on the CPU inside enforce()
// are a sparse matrix-vector product done on the device.
//
// Gadget semantics follow bellman (nam-bellperson / bellpepper-core, reference Cargo.lock:154-155,
// 1355-1358) as restated in SURVEY.md Appendix B; emission order is what fixes the variable
// numbering and hence the key layout, and it is pinned by the reference's cs.hash() strings
// (masp_proofs/src/circuit/convert.rs:218-224, sapling.rs:730-741, 1024-1045).  Gadget by gadget:
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
#include <algorithm>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "fr_host.hpp"

namespace mbh {

inline bool& recording() {
    static thread_local bool r = false;
    return r;
}

struct Var {
    uint32_t id;  // inputs: index; aux: index | AUX
    static constexpr uint32_t AUX = 0x80000000u;
};
static const Var ONE = {0};

struct Consts {  // coefficients of linear combinations (structure: always the scalar field)
    Fr one, minus_one, two;
    Fr pow2[256];
    Consts() {
        one = Fr::one();
        minus_one = -one;
        two = one + one;
        pow2[0] = one;
        for (int i = 1; i < 256; ++i) pow2[i] = pow2[i - 1].dbl();
    }
};
inline const Consts& K() {
    static const Consts k;
    return k;
}

// linear combination; only materialised while recording
struct LC {
    std::vector<std::pair<uint32_t, Fr>> t;
    LC() {}
    LC(Var v, const Fr& c) { add(v, c); }
    LC& add(Var v, const Fr& c) {
        if (recording()) t.emplace_back(v.id, c);
        return *this;
    }
    LC& add(Var v) { return add(v, K().one); }
    LC& sub(Var v) { return add(v, K().minus_one); }
    LC& add(const LC& o) {
        if (recording()) t.insert(t.end(), o.t.begin(), o.t.end());
        return *this;
    }
    LC& sub(const LC& o) {
        if (recording())
            for (auto& e : o.t) t.emplace_back(e.first, -e.second);
        return *this;
    }
    LC scaled(const Fr& k) const {
        LC r;
        if (recording())
            for (auto& e : t) r.t.emplace_back(e.first, e.second * k);
        return r;
    }
};
// what a policy that never records uses in its place: every operation is a no-op
struct NullLC {
    NullLC() {}
    NullLC(Var, const Fr&) {}
    NullLC& add(Var, const Fr&) { return *this; }
    NullLC& add(Var) { return *this; }
    NullLC& sub(Var) { return *this; }
    NullLC& add(const NullLC&) { return *this; }
    NullLC& sub(const NullLC&) { return *this; }
    NullLC scaled(const Fr&) const { return NullLC(); }
};

struct Matrix {  // CSR over the variable ids
    std::vector<uint32_t> rowptr{0};
    std::vector<uint32_t> col;
    std::vector<Fr> coef;
};

// the scalar sink: values in vectors, rows recorded in a shape pass
struct CS {
    std::vector<Fr> inputs, aux;
    size_t n_constraints = 0;
    bool failed = false;  // a witness closure hit a division by zero (bellman: SynthesisError)
    Fr root = Fr::zero(); // the Merkle root the circuit computed (Spend, Convert), for callers that need the anchor
    Matrix A, B, C;       // recorded rows (shape pass only)
    CS() { inputs.push_back(Fr::one()); }

    Var alloc(const Fr& v) {
        aux.push_back(v);
        return Var{(uint32_t)(aux.size() - 1) | Var::AUX};
    }
    Var alloc_bit(bool b) { return alloc(b ? K().one : Fr::zero()); }
    Var alloc_input(const Fr& v) {
        inputs.push_back(v);
        return Var{(uint32_t)(inputs.size() - 1)};
    }
    void fail_if(bool b) {
        if (b) failed = true;
    }
    static void push_row(Matrix& m, const LC& lc) {
        // canonical: merged per variable, zero coefficients dropped, inputs
        // before aux, each by index (the order TestConstraintSystem::hash uses)
        std::vector<std::pair<uint32_t, Fr>> t = lc.t;
        std::stable_sort(t.begin(), t.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
        for (size_t i = 0; i < t.size();) {
            Fr c = t[i].second;
            size_t j = i + 1;
            while (j < t.size() && t[j].first == t[i].first) c += t[j++].second;
            if (!c.is_zero()) {
                m.col.push_back(t[i].first);
                m.coef.push_back(c);
            }
            i = j;
        }
        m.rowptr.push_back((uint32_t)m.col.size());
    }
    void enforce(const LC& a, const LC& b, const LC& c) {
        ++n_constraints;
        if (recording()) {
            push_row(A, a);
            push_row(B, b);
            push_row(C, c);
        }
    }
};

template <int LANES>
struct WordsN {  // one 256-bit little-endian integer per lane (a Jubjub scalar, 32 identifier bytes)
    uint64_t w[LANES][4];
};

struct ScalarPolicy {
    typedef Fr F;         // a field value
    typedef bool B;       // a boolean value
    typedef uint64_t U;   // a small unsigned integer (u32 sums of BLAKE2s, the u64 value)
    typedef Fr C;         // a table constant used as a VALUE (coefficients of LCs are always Fr)
    typedef LC L;
    typedef CS Sink;
    typedef WordsN<1> W;
    static constexpr bool RECORDS = true;
    static constexpr int LANES = 1;
    static F from_words(const uint64_t w[1][4]) { return Fr::from_words(w[0]); }
    static U ufrom(const uint64_t v[1]) { return v[0]; }
    static B bfrom(const bool b[1]) { return b[0]; }
    static C konst(const Fr& f) { return f; }
    static F lift(const C& c) { return c; }
    static F zero() { return Fr::zero(); }
    static F one() { return Fr::one(); }
    static B ball(bool b) { return b; }
    static bool any(B b) { return b; }
    static F select(B c, const F& a, const F& b) { return c ? a : b; }
    static F mask(B c, const C& a) { return c ? a : Fr::zero(); }
    static B is_zero(const F& a) { return a.is_zero(); }
    static F inverse(const F& a) { return a.inverse(); }
    static void to_bits(const F& v, int n, B* out) {
        uint64_t w[4];
        v.to_words(w);
        for (int i = 0; i < n; ++i) out[i] = (w[i >> 6] >> (i & 63)) & 1;
    }
    static B wbit(const W& w, int i) { return (w.w[0][i >> 6] >> (i & 63)) & 1; }
    static U uzero() { return 0; }
    static U uset(U v, B b, int i) { return v | ((uint64_t)(b ? 1 : 0) << i); }
    static B ubit(U v, int i) { return (v >> i) & 1; }
    static U uadd(U a, U b) { return a + b; }
};

// ---------------------------------------------------------------------------
// window tables: LC coefficients (scalar) and values (per policy)
// ---------------------------------------------------------------------------
inline void synth_coeffs(int window, const Fr* consts, Fr* a) {
    int n = 1 << window;
    for (int i = 0; i < n; ++i) a[i] = Fr::zero();
    for (int i = 0; i < n; ++i) {
        Fr cur = consts[i] - a[i];
        a[i] = cur;
        for (int j = i + 1; j < n; ++j)
            if ((j & i) == i) a[j] += cur;
    }
}
struct Window8 {  // one 3-bit window of a fixed-base table: 8 (u, v) pairs + their interpolation coefficients
    Fr u[8], v[8], uc[8], vc[8];
};
struct Window4 {  // one Pedersen window: 4 Montgomery (x, y) pairs + coefficients
    Fr x[4], y[4], xc[4], yc[4];
};

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

struct JPointS {  // a Jubjub point over the scalar field (table construction, callers)
    Fr u, v;
};

// curve constants and the circuit tables, computed once on the scalar field
struct Jubjub {
    Fr d, d2, mont_a, mont_scale;
    JPointS proof_generation_key_generator, note_commitment_randomness_generator, nullifier_position_generator,
        value_commitment_randomness_generator, spending_key_generator;
    JPointS pedersen_generators[6];
    std::vector<Window8> fixed[5];     // 84 windows each; index = enum Fixed
    std::vector<Window4> pedersen[6];  // 63 windows each
    enum Fixed { PROOF_GENERATION_KEY = 0, NOTE_COMMITMENT_RANDOMNESS, NULLIFIER_POSITION, VALUE_COMMITMENT_RANDOMNESS, SPENDING_KEY };

    JPointS add(const JPointS& p, const JPointS& q) const {
        Fr uv = p.u * q.v, vu = p.v * q.u;
        Fr t = d * uv * vu;
        Fr one = Fr::one();
        Fr dp = one + t, dm = one - t;
        Fr i = (dp * dm).inverse();
        return {(uv + vu) * (i * dm), (p.v * q.v + p.u * q.u) * (i * dp)};
    }
    JPointS dbl(const JPointS& p) const { return add(p, p); }
    bool to_montgomery(const JPointS& p, Fr& x, Fr& y) const {  // constants.rs:100-141
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
        const JPointS* gens[5] = {&proof_generation_key_generator, &note_commitment_randomness_generator,
                                  &nullifier_position_generator, &value_commitment_randomness_generator,
                                  &spending_key_generator};
        for (int k = 0; k < 5; ++k) {
            JPointS g0 = *gens[k];
            for (int w = 0; w < 84; ++w) {
                Window8 win;
                JPointS g = g0;
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
            JPointS g0 = pedersen_generators[k];
            for (int w = 0; w < 63; ++w) {
                Window4 win;
                JPointS g = g0;
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

namespace blake2s_gadget {
static const uint32_t IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
static const uint8_t SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
}  // namespace blake2s_gadget

static const bool NOTE_COMMITMENT_PERSONALIZATION[6] = {true, true, true, true, true, true};
inline void merkle_personalization(int depth, bool out[6]) {
    for (int i = 0; i < 6; ++i) out[i] = (depth >> i) & 1;
}

// In a witness pass the three linear combinations are not even constructed.
#define MBH_ENFORCE(cs, a, b, c)                              \
    do {                                                      \
        if (P::RECORDS && ::mbh::recording()) (cs).enforce(a, b, c); \
        else ++(cs).n_constraints;                            \
    } while (0)

// ===========================================================================
// everything that touches witness VALUES, over a policy
// ===========================================================================
template <class P>
struct G {
    typedef typename P::F F;
    typedef typename P::B B;
    typedef typename P::U U;
    typedef typename P::C C;
    typedef typename P::L LC;
    typedef typename P::Sink CS;
    typedef typename P::W W;

    struct JPoint {
        F u, v;
    };

    // ---- value-side constants of this policy -------------------------------------------------
    struct Window8V {
        C u[8], v[8];
    };
    struct Window4V {
        C x[4], y[4], xc[4];
    };
    struct Tables {
        C d, d2, mont_a, mont_scale, one;
        C pow2[256];
        std::vector<Window8V> fixed[5];
        std::vector<Window4V> pedersen[6];
        Tables() {
            const Jubjub& J = JJ();
            d = P::konst(J.d);
            d2 = P::konst(J.d2);
            mont_a = P::konst(J.mont_a);
            mont_scale = P::konst(J.mont_scale);
            one = P::konst(Fr::one());
            for (int i = 0; i < 256; ++i) pow2[i] = P::konst(K().pow2[i]);
            for (int k = 0; k < 5; ++k)
                for (const Window8& w : J.fixed[k]) {
                    Window8V o;
                    for (int j = 0; j < 8; ++j) {
                        o.u[j] = P::konst(w.u[j]);
                        o.v[j] = P::konst(w.v[j]);
                    }
                    fixed[k].push_back(o);
                }
            for (int k = 0; k < 6; ++k)
                for (const Window4& w : J.pedersen[k]) {
                    Window4V o;
                    for (int j = 0; j < 4; ++j) {
                        o.x[j] = P::konst(w.x[j]);
                        o.y[j] = P::konst(w.y[j]);
                        o.xc[j] = P::konst(w.xc[j]);
                    }
                    pedersen[k].push_back(o);
                }
        }
    };
    static const Tables& T() {
        static const Tables t;
        return t;
    }
    // table entry selected by two / three booleans (bit 0 first)
    static F lookup4(B b0, B b1, const C* t) {
        return P::select(b1, P::select(b0, P::lift(t[3]), P::lift(t[2])), P::select(b0, P::lift(t[1]), P::lift(t[0])));
    }
    static F lookup8(B b0, B b1, B b2, const C* t) { return P::select(b2, lookup4(b0, b1, t + 4), lookup4(b0, b1, t)); }

    // ---- native Jubjub arithmetic on values ----------------------------------------------------
    // The circuits walk long chains of point additions whose every intermediate point is a witness
    // value in affine form.  bellman's closures invert once per step; here a chain is first run in
    // projective coordinates and all its denominators are inverted together (Montgomery's trick),
    // which yields the same affine values with one inversion per chain.
    struct Ext {  // extended twisted Edwards, a = -1: u = X/Z, v = Y/Z, T = XY/Z
        F X, Y, Z, T;
    };
    static JPoint identity() { return {P::zero(), P::one()}; }
    static Ext ext_of(const JPoint& p) { return {p.u, p.v, P::one(), p.u * p.v}; }
    static Ext ext_add(const Ext& p, const Ext& q) {
        F A = (p.Y - p.X) * (q.Y - q.X), Bq = (p.Y + p.X) * (q.Y + q.X);
        F Cq = p.T * P::lift(T().d2) * q.T, D = (p.Z * q.Z).dbl();
        F E = Bq - A, Fq = D - Cq, Gq = D + Cq, H = Bq + A;
        return {E * Fq, Gq * H, Fq * Gq, E * H};
    }
    static Ext ext_dbl(const Ext& p) {
        F A = p.X.square(), Bq = p.Y.square(), Cq = p.Z.square().dbl();
        F D = -A;
        F E = (p.X + p.Y).square() - A - Bq, Gq = D + Bq;
        F Fq = Gq - Cq, H = D - Bq;
        return {E * Fq, Gq * H, Fq * Gq, E * H};
    }
    static Ext ext_select(B c, const Ext& a, const Ext& b) {
        return {P::select(c, a.X, b.X), P::select(c, a.Y, b.Y), P::select(c, a.Z, b.Z), P::select(c, a.T, b.T)};
    }
    // v[i] <- 1 / v[i]; zeros stay zero and are reported
    static B batch_inverse(F* v, size_t n, std::vector<F>& scratch) {
        B any_zero = P::ball(false);
        scratch.resize(n);
        F acc = P::one();
        for (size_t i = 0; i < n; ++i) {
            scratch[i] = acc;
            B z = P::is_zero(v[i]);
            any_zero = any_zero | z;
            acc = acc * P::select(z, P::one(), v[i]);
        }
        acc = P::inverse(acc);
        for (size_t i = n; i-- > 0;) {
            B z = P::is_zero(v[i]);
            F t = acc * scratch[i];
            acc = acc * P::select(z, P::one(), v[i]);
            v[i] = P::select(z, P::zero(), t);
        }
        return any_zero;
    }
    static void ext_to_affine(const Ext* in, JPoint* out, size_t n) {
        std::vector<F> z(n), scratch;
        for (size_t i = 0; i < n; ++i) z[i] = in[i].Z;
        batch_inverse(z.data(), n, scratch);
        for (size_t i = 0; i < n; ++i) out[i] = {in[i].X * z[i], in[i].Y * z[i]};
    }
    // sums[k] = pts[0] + ... + pts[k]
    static void chain_sums(const JPoint* pts, size_t n, JPoint* sums) {
        if (!n) return;
        std::vector<Ext> e(n);
        e[0] = ext_of(pts[0]);
        for (size_t k = 1; k < n; ++k) e[k] = ext_add(e[k - 1], ext_of(pts[k]));
        ext_to_affine(e.data(), sums, n);
    }
    // Everything a Pedersen hash needs from the field's division, for all of its segments at once:
    // per window the slope of its chain step, per segment the Edwards image of the segment sum and
    // the running Edwards sum.  Three inversions per hash (all slope denominators and segment Z's;
    // the Montgomery -> Edwards denominators; the running sums), however many segments it has.
    // A segment runs as a projective Montgomery chord, (X : Y : Z) + affine (tx, ty): the step's
    // slope is u / v with u = ty Z - Y, v = tx Z - X, so only the v's have to be inverted.
    struct PedersenPlan {
        std::vector<F> lam;          // per window (first window of a segment: unused)
        std::vector<JPoint> seg_ed;  // per segment: into_edwards of the segment's sum
        std::vector<JPoint> run_ed;  // per segment: seg_ed[0] + ... + seg_ed[s]
        B bad;                       // an exceptional step (equal x: the reference's DivisionByZero)
    };
    static void pedersen_plan(const std::vector<F>& tx, const std::vector<F>& ty, const std::vector<size_t>& seg_len,
                              PedersenPlan& plan) {
        const size_t nw = tx.size(), ns = seg_len.size();
        const F mont_a = P::lift(T().mont_a), mont_scale = P::lift(T().mont_scale), one = P::one();
        std::vector<F> u(nw), den(nw + ns), scratch;  // den: v per window, then Z per segment
        std::vector<F> fx(ns), fy(ns);                // projective numerators of the segment sums
        size_t base = 0;
        for (size_t sgm = 0; sgm < ns; ++sgm) {
            size_t n = seg_len[sgm];
            F X = tx[base], Y = ty[base], Z = one;
            den[base] = one;
            for (size_t k = 1; k < n; ++k) {
                const F &x2 = tx[base + k], &y2 = ty[base + k];
                F x2z = x2 * Z;
                F uu = y2 * Z - Y, v1 = x2z - X;
                u[base + k] = uu;
                den[base + k] = v1;
                F vv = v1.square(), vvv = vv * v1;
                F Wq = uu.square() * Z - vv * (mont_a * Z + x2z + X);
                F Xn = Wq * v1;
                Y = uu * (X * vv - Wq) - Y * vvv;
                X = Xn;
                Z = vvv * Z;
            }
            fx[sgm] = X;
            fy[sgm] = Y;
            den[nw + sgm] = Z;
            base += n;
        }
        plan.bad = batch_inverse(den.data(), nw + ns, scratch);
        plan.lam.assign(nw, P::zero());
        base = 0;
        for (size_t sgm = 0; sgm < ns; ++sgm) {
            for (size_t k = 1; k < seg_len[sgm]; ++k) plan.lam[base + k] = u[base + k] * den[base + k];
            base += seg_len[sgm];
        }
        // Montgomery -> Edwards for every segment sum: u = scale x / y, v = (x - 1) / (x + 1)
        std::vector<F> sx(ns), sy(ns), dd(ns);
        for (size_t sgm = 0; sgm < ns; ++sgm) {
            sx[sgm] = fx[sgm] * den[nw + sgm];
            sy[sgm] = fy[sgm] * den[nw + sgm];
            dd[sgm] = sy[sgm] * (sx[sgm] + one);
        }
        plan.bad = plan.bad | batch_inverse(dd.data(), ns, scratch);
        plan.seg_ed.resize(ns);
        for (size_t sgm = 0; sgm < ns; ++sgm) {
            F xp = sx[sgm] + one;
            plan.seg_ed[sgm] = {sx[sgm] * mont_scale * (dd[sgm] * xp), (sx[sgm] - one) * (dd[sgm] * sy[sgm])};
        }
        plan.run_ed.resize(ns);
        chain_sums(plan.seg_ed.data(), ns, plan.run_ed.data());
    }

    // ---------------------------------------------------------------------------
    // booleans
    // ---------------------------------------------------------------------------
    struct AllocatedBit {
        Var var;
        B value;
        static AllocatedBit alloc(CS& cs, B value) {
            Var v = cs.alloc_bit(value);
            MBH_ENFORCE(cs, LC(ONE, K().one).sub(v), LC(v, K().one), LC());
            return {v, value};
        }
        static AllocatedBit alloc_conditionally(CS& cs, B value, const AllocatedBit& must_be_false) {
            Var v = cs.alloc_bit(value);
            MBH_ENFORCE(cs, LC(ONE, K().one).sub(must_be_false.var).sub(v), LC(v, K().one), LC());
            return {v, value};
        }
        static AllocatedBit and_(CS& cs, const AllocatedBit& a, const AllocatedBit& b) {
            B val = a.value & b.value;
            Var v = cs.alloc_bit(val);
            MBH_ENFORCE(cs, LC(a.var, K().one), LC(b.var, K().one), LC(v, K().one));
            return {v, val};
        }
        static AllocatedBit and_not(CS& cs, const AllocatedBit& a, const AllocatedBit& b) {
            B val = a.value & !b.value;
            Var v = cs.alloc_bit(val);
            MBH_ENFORCE(cs, LC(a.var, K().one), LC(ONE, K().one).sub(b.var), LC(v, K().one));
            return {v, val};
        }
        static AllocatedBit nor(CS& cs, const AllocatedBit& a, const AllocatedBit& b) {
            B val = !a.value & !b.value;
            Var v = cs.alloc_bit(val);
            MBH_ENFORCE(cs, LC(ONE, K().one).sub(a.var), LC(ONE, K().one).sub(b.var), LC(v, K().one));
            return {v, val};
        }
        static AllocatedBit xor_(CS& cs, const AllocatedBit& a, const AllocatedBit& b) {
            B val = a.value ^ b.value;
            Var v = cs.alloc_bit(val);
            MBH_ENFORCE(cs, LC(a.var, K().two), LC(b.var, K().one), LC(a.var, K().one).add(b.var).sub(v));
            return {v, val};
        }
    };

    struct Boolean {
        enum Kind : uint8_t { IS, NOT, CONST } kind;
        AllocatedBit bit;
        bool c;  // the constant: structure, not a witness value
        static Boolean constant(bool b) { return {CONST, {{0}, P::ball(false)}, b}; }
        static Boolean from_bit(const AllocatedBit& b) { return {IS, b, false}; }
        B value() const { return kind == CONST ? P::ball(c) : (kind == IS ? bit.value : !bit.value); }
        bool is_const() const { return kind == CONST; }
        Boolean not_() const {
            if (kind == CONST) return constant(!c);
            return {kind == IS ? NOT : IS, bit, false};
        }
        // lc += k * self
        void add_to(LC& lc, const Fr& k) const {
            if (!(P::RECORDS && recording())) return;
            if (kind == CONST) {
                if (c) lc.add(ONE, k);
            } else if (kind == IS) {
                lc.add(bit.var, k);
            } else {
                lc.add(ONE, k);
                lc.add(bit.var, -k);
            }
        }
        LC lc(const Fr& k) const {
            LC r;
            add_to(r, k);
            return r;
        }
        static Boolean and_(CS& cs, const Boolean& a, const Boolean& b) {
            if (a.kind == CONST) return a.c ? b : constant(false);
            if (b.kind == CONST) return b.c ? a : constant(false);
            if (a.kind == IS && b.kind == IS) return from_bit(AllocatedBit::and_(cs, a.bit, b.bit));
            if (a.kind == IS && b.kind == NOT) return from_bit(AllocatedBit::and_not(cs, a.bit, b.bit));
            if (a.kind == NOT && b.kind == IS) return from_bit(AllocatedBit::and_not(cs, b.bit, a.bit));
            return from_bit(AllocatedBit::nor(cs, a.bit, b.bit));
        }
        static Boolean xor_(CS& cs, const Boolean& a, const Boolean& b) {
            if (a.kind == CONST) return a.c ? b.not_() : b;
            if (b.kind == CONST) return b.c ? a.not_() : a;
            if (a.kind == IS && b.kind == NOT) return from_bit(AllocatedBit::xor_(cs, a.bit, b.bit)).not_();
            if (a.kind == NOT && b.kind == IS) return from_bit(AllocatedBit::xor_(cs, b.bit, a.bit)).not_();
            return from_bit(AllocatedBit::xor_(cs, a.bit, b.bit));
        }
        static void enforce_equal(CS& cs, const Boolean& a, const Boolean& b) {
            if (a.kind == CONST && b.kind == CONST) {
                if (a.c != b.c) cs.fail_if(P::ball(true));
                return;
            }
            LC c = a.lc(K().one);
            c.sub(b.lc(K().one));
            MBH_ENFORCE(cs, LC(), LC(), c);
        }
    };
    typedef std::vector<Boolean> Bits;

    static Bits u64_into_boolean_vec_le(CS& cs, U value) {
        Bits r;
        for (int i = 0; i < 64; ++i) r.push_back(Boolean::from_bit(AllocatedBit::alloc(cs, P::ubit(value, i))));
        return r;
    }
    // masp_proofs/src/circuit/gadgets.rs:6-50 on a 256-bit little-endian integer
    static Bits words_into_boolean_vec_le(CS& cs, const W& w, int num_bits) {
        Bits r;
        for (int i = 0; i < num_bits; ++i) r.push_back(Boolean::from_bit(AllocatedBit::alloc(cs, P::wbit(w, i))));
        return r;
    }

    // ---------------------------------------------------------------------------
    // numbers
    // ---------------------------------------------------------------------------
    struct AllocatedNum {
        Var var;
        F value;
        static AllocatedNum alloc(CS& cs, const F& v) { return {cs.alloc(v), v}; }
        AllocatedNum mul(CS& cs, const AllocatedNum& o) const {
            AllocatedNum out = alloc(cs, value * o.value);
            MBH_ENFORCE(cs, LC(var, K().one), LC(o.var, K().one), LC(out.var, K().one));
            return out;
        }
        AllocatedNum square(CS& cs) const {
            AllocatedNum out = alloc(cs, value.square());
            MBH_ENFORCE(cs, LC(var, K().one), LC(var, K().one), LC(out.var, K().one));
            return out;
        }
        void assert_nonzero(CS& cs) const {
            cs.fail_if(P::is_zero(value));
            Var inv = cs.alloc(P::inverse(value));
            MBH_ENFORCE(cs, LC(var, K().one), LC(inv, K().one), LC(ONE, K().one));
        }
        void inputize(CS& cs) const {
            Var inp = cs.alloc_input(value);
            MBH_ENFORCE(cs, LC(inp, K().one), LC(ONE, K().one), LC(var, K().one));
        }
        static void conditionally_reverse(CS& cs, const AllocatedNum& a, const AllocatedNum& b, const Boolean& cond,
                                          AllocatedNum& c, AllocatedNum& d) {
            B cv = cond.value();
            c = alloc(cs, P::select(cv, b.value, a.value));
            MBH_ENFORCE(cs, LC(a.var, K().one).sub(b.var), cond.lc(K().one), LC(a.var, K().one).sub(c.var));
            d = alloc(cs, P::select(cv, a.value, b.value));
            MBH_ENFORCE(cs, LC(b.var, K().one).sub(a.var), cond.lc(K().one), LC(b.var, K().one).sub(d.var));
        }
        Bits to_bits_le(CS& cs) const {
            B w[255];
            P::to_bits(value, 255, w);
            Bits bits;
            LC lc;
            for (int i = 0; i < 255; ++i) {
                AllocatedBit b = AllocatedBit::alloc(cs, w[i]);
                lc.add(b.var, K().pow2[i]);
                bits.push_back(Boolean::from_bit(b));
            }
            lc.sub(var);
            MBH_ENFORCE(cs, LC(), LC(), lc);
            return bits;
        }
        // bits of the value with the proof that they encode an integer <= r - 1
        Bits to_bits_le_strict(CS& cs) const {
            const uint64_t bound[4] = {Fr::M0 - 1, Fr::M1, Fr::M2, Fr::M3};
            auto bound_bit = [&](int i) { return (bound[i >> 6] >> (i & 63)) & 1; };
            B w[255];
            P::to_bits(value, 255, w);
            std::vector<AllocatedBit> result;  // big-endian
            std::vector<AllocatedBit> run;
            bool have_last = false;
            AllocatedBit last_run = {{0}, P::ball(false)};
            for (int i = 254; i >= 0; --i) {
                B a_bit = w[i];
                if (bound_bit(i)) {
                    AllocatedBit bit = AllocatedBit::alloc(cs, a_bit);
                    run.push_back(bit);
                    result.push_back(bit);
                } else {
                    if (!run.empty()) {
                        if (have_last) run.push_back(last_run);
                        AllocatedBit cur = run[0];
                        for (size_t k = 1; k < run.size(); ++k) cur = AllocatedBit::and_(cs, cur, run[k]);
                        last_run = cur;
                        have_last = true;
                        run.clear();
                    }
                    result.push_back(AllocatedBit::alloc_conditionally(cs, a_bit, last_run));
                }
            }
            LC lc;
            Bits out(255, Boolean::constant(false));
            for (int i = 0; i < 255; ++i) {
                const AllocatedBit& b = result[254 - i];
                lc.add(b.var, K().pow2[i]);
                out[i] = Boolean::from_bit(b);
            }
            lc.sub(var);
            MBH_ENFORCE(cs, LC(), LC(), lc);
            return out;
        }
    };

    struct Num {
        LC lc;
        F value = P::zero();
        static Num from_allocated(const AllocatedNum& n) {
            Num r;
            r.lc.add(n.var, K().one);
            r.value = n.value;
            return r;
        }
        // coeff twice: as an LC coefficient (scalar field) and as a value constant of the policy
        Num& add_bool_with_coeff(const Boolean& bit, const Fr& coeff, const C& coeff_v) {
            bit.add_to(lc, coeff);
            if (bit.is_const()) {
                if (bit.c) value += P::lift(coeff_v);
            } else {
                value += P::mask(bit.value(), coeff_v);
            }
            return *this;
        }
    };

    // ---------------------------------------------------------------------------
    // window lookups
    // ---------------------------------------------------------------------------
    static void lookup3_xy(CS& cs, const Boolean bits[3], const Window8& w, const Window8V& wv, AllocatedNum& res_x,
                           AllocatedNum& res_y) {
        B b0 = bits[0].value(), b1 = bits[1].value(), b2 = bits[2].value();
        res_x = AllocatedNum::alloc(cs, lookup8(b0, b1, b2, wv.u));
        res_y = AllocatedNum::alloc(cs, lookup8(b0, b1, b2, wv.v));
        Boolean precomp = Boolean::and_(cs, bits[1], bits[2]);
        for (int k = 0; k < 2; ++k) {
            const Fr* co = k ? w.vc : w.uc;
            const AllocatedNum& res = k ? res_y : res_x;
            LC a, c;
            if (P::RECORDS && recording()) {
                a.add(ONE, co[1]);
                bits[1].add_to(a, co[3]);
                bits[2].add_to(a, co[5]);
                precomp.add_to(a, co[7]);
                c.add(res.var, K().one);
                c.add(ONE, -co[0]);
                bits[1].add_to(c, -co[2]);
                bits[2].add_to(c, -co[4]);
                precomp.add_to(c, -co[6]);
            }
            MBH_ENFORCE(cs, a, bits[0].lc(K().one), c);
        }
    }

    static void lookup3_xy_with_conditional_negation(CS& cs, const Boolean bits[3], const Window4& w, const Window4V& wv,
                                                     Num& x, Num& y) {
        F yv = lookup4(bits[0].value(), bits[1].value(), wv.y);
        yv = P::select(bits[2].value(), -yv, yv);
        AllocatedNum ya = AllocatedNum::alloc(cs, yv);
        Boolean precomp = Boolean::and_(cs, bits[0], bits[1]);
        x = Num();
        x.add_bool_with_coeff(Boolean::constant(true), w.xc[0], wv.xc[0]);
        x.add_bool_with_coeff(bits[0], w.xc[1], wv.xc[1]);
        x.add_bool_with_coeff(bits[1], w.xc[2], wv.xc[2]);
        x.add_bool_with_coeff(precomp, w.xc[3], wv.xc[3]);
        LC y_lc;
        if (P::RECORDS && recording()) {
            precomp.add_to(y_lc, w.yc[3]);
            bits[1].add_to(y_lc, w.yc[2]);
            bits[0].add_to(y_lc, w.yc[1]);
            y_lc.add(ONE, w.yc[0]);
        }
        LC a = y_lc;
        a.add(y_lc);
        LC c = y_lc;
        c.sub(ya.var);
        MBH_ENFORCE(cs, a, bits[2].lc(K().one), c);
        y = Num::from_allocated(ya);
    }

    // multipack::pack_into_inputs: 254 bits per public input
    static void pack_into_inputs(CS& cs, const Bits& bits) {
        for (size_t k = 0; k < bits.size(); k += 254) {
            Num num;
            for (size_t j = k; j < bits.size() && j < k + 254; ++j)
                num.add_bool_with_coeff(bits[j], K().pow2[j - k], T().pow2[j - k]);
            Var inp = cs.alloc_input(num.value);
            MBH_ENFORCE(cs, num.lc, LC(ONE, K().one), LC(inp, K().one));
        }
    }

    // ---------------------------------------------------------------------------
    // UInt32, MultiEq, BLAKE2s
    // ---------------------------------------------------------------------------
    struct MultiEq {
        CS& cs;
        int bits_used = 0;
        LC lhs, rhs;
        explicit MultiEq(CS& c) : cs(c) {}
        void accumulate() {
            MBH_ENFORCE(cs, lhs, LC(ONE, K().one), rhs);
            lhs = LC();
            rhs = LC();
            bits_used = 0;
        }
        void enforce_equal(int num_bits, const LC& l, const LC& r) {
            if (254 <= bits_used + num_bits) accumulate();
            lhs.add(l.scaled(K().pow2[bits_used]));
            rhs.add(r.scaled(K().pow2[bits_used]));
            bits_used += num_bits;
        }
        void close() {
            if (bits_used > 0) accumulate();
        }
    };

    struct UInt32 {
        Boolean bits[32];  // LSB first
        static UInt32 constant(uint32_t v) {
            UInt32 r;
            for (int i = 0; i < 32; ++i) r.bits[i] = Boolean::constant((v >> i) & 1);
            return r;
        }
        U value() const {
            U v = P::uzero();
            for (int i = 0; i < 32; ++i) v = P::uset(v, bits[i].value(), i);
            return v;
        }
        UInt32 rotr(int k) const {
            UInt32 r;
            for (int i = 0; i < 32; ++i) r.bits[i] = bits[(i + k) % 32];
            return r;
        }
        UInt32 xor_(CS& cs, const UInt32& o) const {
            UInt32 r;
            for (int i = 0; i < 32; ++i) r.bits[i] = Boolean::xor_(cs, bits[i], o.bits[i]);
            return r;
        }
        static UInt32 addmany(MultiEq& meq, const UInt32* const* ops, int nops) {
            CS& cs = meq.cs;
            uint64_t max_value = (uint64_t)nops * 0xffffffffull;
            U total = P::uzero();
            LC lc;
            bool all_constants = true;
            uint64_t const_total = 0;
            for (int k = 0; k < nops; ++k) {
                total = P::uadd(total, ops[k]->value());
                uint32_t cv = 0;
                for (int i = 0; i < 32; ++i) {
                    ops[k]->bits[i].add_to(lc, K().pow2[i]);
                    all_constants &= ops[k]->bits[i].is_const();
                    if (ops[k]->bits[i].is_const() && ops[k]->bits[i].c) cv |= 1u << i;
                }
                const_total += cv;
            }
            if (all_constants) return constant((uint32_t)const_total);
            UInt32 r;
            LC result_lc;
            int i = 0;
            while (max_value) {
                AllocatedBit b = AllocatedBit::alloc(cs, P::ubit(total, i));
                result_lc.add(b.var, K().pow2[i]);
                if (i < 32) r.bits[i] = Boolean::from_bit(b);
                max_value >>= 1;
                ++i;
            }
            meq.enforce_equal(i, lc, result_lc);
            return r;
        }
    };

    static void mixing_g(MultiEq& meq, UInt32* v, int a, int b, int c, int d, const UInt32& x, const UInt32& y) {
        CS& cs = meq.cs;
        {
            const UInt32* ops[3] = {&v[a], &v[b], &x};
            v[a] = UInt32::addmany(meq, ops, 3);
        }
        v[d] = v[d].xor_(cs, v[a]).rotr(16);
        {
            const UInt32* ops[2] = {&v[c], &v[d]};
            v[c] = UInt32::addmany(meq, ops, 2);
        }
        v[b] = v[b].xor_(cs, v[c]).rotr(12);
        {
            const UInt32* ops[3] = {&v[a], &v[b], &y};
            v[a] = UInt32::addmany(meq, ops, 3);
        }
        v[d] = v[d].xor_(cs, v[a]).rotr(8);
        {
            const UInt32* ops[2] = {&v[c], &v[d]};
            v[c] = UInt32::addmany(meq, ops, 2);
        }
        v[b] = v[b].xor_(cs, v[c]).rotr(7);
    }

    static void compression(CS& cs, UInt32 h[8], const UInt32 m[16], uint64_t t, bool final) {
        using namespace blake2s_gadget;
        UInt32 v[16];
        for (int i = 0; i < 8; ++i) {
            v[i] = h[i];
            v[i + 8] = UInt32::constant(IV[i]);
        }
        v[12] = v[12].xor_(cs, UInt32::constant((uint32_t)t));
        v[13] = v[13].xor_(cs, UInt32::constant((uint32_t)(t >> 32)));
        if (final) v[14] = v[14].xor_(cs, UInt32::constant(0xffffffffu));
        MultiEq meq(cs);
        for (int i = 0; i < 10; ++i) {
            const uint8_t* s = SIGMA[i];
            mixing_g(meq, v, 0, 4, 8, 12, m[s[0]], m[s[1]]);
            mixing_g(meq, v, 1, 5, 9, 13, m[s[2]], m[s[3]]);
            mixing_g(meq, v, 2, 6, 10, 14, m[s[4]], m[s[5]]);
            mixing_g(meq, v, 3, 7, 11, 15, m[s[6]], m[s[7]]);
            mixing_g(meq, v, 0, 5, 10, 15, m[s[8]], m[s[9]]);
            mixing_g(meq, v, 1, 6, 11, 12, m[s[10]], m[s[11]]);
            mixing_g(meq, v, 2, 7, 8, 13, m[s[12]], m[s[13]]);
            mixing_g(meq, v, 3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        meq.close();
        for (int i = 0; i < 8; ++i) h[i] = h[i].xor_(cs, v[i]).xor_(cs, v[i + 8]);
    }

    // BLAKE2s-256 of a bit string (bytes in little-endian bit order), 8-byte personalization
    static Bits blake2s(CS& cs, const Bits& input, const char personalization[8]) {
        using namespace blake2s_gadget;
        auto le32 = [](const char* p) {
            return (uint32_t)(uint8_t)p[0] | ((uint32_t)(uint8_t)p[1] << 8) | ((uint32_t)(uint8_t)p[2] << 16) |
                   ((uint32_t)(uint8_t)p[3] << 24);
        };
        UInt32 h[8];
        for (int i = 0; i < 8; ++i) h[i] = UInt32::constant(IV[i]);
        h[0] = UInt32::constant(IV[0] ^ 0x01010000u ^ 32u);
        h[6] = UInt32::constant(IV[6] ^ le32(personalization));
        h[7] = UInt32::constant(IV[7] ^ le32(personalization + 4));
        std::vector<std::vector<UInt32>> blocks;
        for (size_t k = 0; k < input.size(); k += 512) {
            std::vector<UInt32> words;
            size_t end = std::min(input.size(), k + 512);
            for (size_t w = k; w < end; w += 32) {
                UInt32 u = UInt32::constant(0);
                for (size_t j = w; j < end && j < w + 32; ++j) u.bits[j - w] = input[j];
                words.push_back(u);
            }
            while (words.size() < 16) words.push_back(UInt32::constant(0));
            blocks.push_back(words);
        }
        if (blocks.empty()) blocks.push_back(std::vector<UInt32>(16, UInt32::constant(0)));
        for (size_t i = 0; i + 1 < blocks.size(); ++i) compression(cs, h, blocks[i].data(), (i + 1) * 64, false);
        compression(cs, h, blocks.back().data(), input.size() / 8, true);
        Bits out;
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 32; ++j) out.push_back(h[i].bits[j]);
        return out;
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
            F s = u.value + v.value;
            AllocatedNum t = AllocatedNum::alloc(cs, s.square());
            MBH_ENFORCE(cs, LC(u.var, K().one).add(v.var), LC(u.var, K().one).add(v.var), LC(t.var, K().one));
            AllocatedNum a = u.mul(cs, v);
            AllocatedNum c = AllocatedNum::alloc(cs, a.value.square() * P::lift(T().d));
            MBH_ENFORCE(cs, LC(a.var, d), LC(a.var, K().one), LC(c.var, K().one));
            JPoint r;
            if (hint) {
                r = *hint;
            } else {
                F one = P::one();
                F dp = one + c.value, dm = one - c.value;
                F i = P::inverse(dp * dm);
                cs.fail_if(P::is_zero(i));
                F a2 = a.value.dbl();
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
            AllocatedNum c = AllocatedNum::alloc(cs, a.value * b.value * P::lift(T().d));
            MBH_ENFORCE(cs, LC(a.var, d), LC(b.var, K().one), LC(c.var, K().one));
            JPoint r;
            if (hint) {
                r = *hint;
            } else {
                F one = P::one();
                F dp = one + c.value, dm = one - c.value;
                F i = P::inverse(dp * dm);
                cs.fail_if(P::is_zero(i));
                r = {(a.value + b.value) * (i * dm), (big_u.value - a.value - b.value) * (i * dp)};
            }
            AllocatedNum u3 = AllocatedNum::alloc(cs, r.u);
            MBH_ENFORCE(cs, LC(ONE, K().one).add(c.var), LC(u3.var, K().one), LC(a.var, K().one).add(b.var));
            AllocatedNum v3 = AllocatedNum::alloc(cs, r.v);
            MBH_ENFORCE(cs, LC(ONE, K().one).sub(c.var), LC(v3.var, K().one), LC(big_u.var, K().one).sub(a.var).sub(b.var));
            return {u3, v3};
        }
        EdwardsPoint conditionally_select(CS& cs, const Boolean& cond) const {
            B cv = cond.value();
            AllocatedNum up = AllocatedNum::alloc(cs, P::select(cv, u.value, P::zero()));
            MBH_ENFORCE(cs, LC(u.var, K().one), cond.lc(K().one), LC(up.var, K().one));
            AllocatedNum vp = AllocatedNum::alloc(cs, P::select(cv, v.value, P::one()));
            LC c(vp.var, K().one);
            c.sub(cond.not_().lc(K().one));
            MBH_ENFORCE(cs, LC(v.var, K().one), cond.lc(K().one), c);
            return {up, vp};
        }
        EdwardsPoint mul(CS& cs, const Bits& by) const {
            // pre-pass: every 2^i P and every running sum, one inversion for all of them
            size_t n = by.size();
            std::vector<Ext> e(2 * n);
            std::vector<JPoint> aff(2 * n);
            if (n) {
                e[0] = ext_of({u.value, v.value});
                for (size_t i = 1; i < n; ++i) e[i] = ext_dbl(e[i - 1]);
                Ext id = ext_of(identity());
                e[n] = ext_select(by[0].value(), e[0], id);
                for (size_t i = 1; i < n; ++i) {
                    if (std::is_same<B, bool>::value) {  // one witness: skip the additions its bits do not ask for
                        e[n + i] = P::any(by[i].value()) ? ext_add(e[n + i - 1], e[i]) : e[n + i - 1];
                    } else {
                        e[n + i] = ext_select(by[i].value(), ext_add(e[n + i - 1], e[i]), e[n + i - 1]);
                    }
                }
                ext_to_affine(e.data(), aff.data(), 2 * n);
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

    static EdwardsPoint fixed_base_multiplication(CS& cs, Jubjub::Fixed gen, const Bits& by) {
        const std::vector<Window8>& table = JJ().fixed[gen];
        const std::vector<Window8V>& tv = T().fixed[gen];
        EdwardsPoint result;
        Boolean f = Boolean::constant(false);
        // pre-pass: the looked-up points and their running sums
        size_t nw = (by.size() + 2) / 3;
        std::vector<JPoint> pts(nw), sums(nw);
        for (size_t k = 0; k < nw; ++k) {
            size_t i = 3 * k;
            B b0 = by[i].value(), b1 = i + 1 < by.size() ? by[i + 1].value() : P::ball(false),
              b2 = i + 2 < by.size() ? by[i + 2].value() : P::ball(false);
            pts[k] = {lookup8(b0, b1, b2, tv[k].u), lookup8(b0, b1, b2, tv[k].v)};
        }
        chain_sums(pts.data(), nw, sums.data());
        for (size_t i = 0; i < by.size(); i += 3) {
            Boolean chunk[3] = {by[i], i + 1 < by.size() ? by[i + 1] : f, i + 2 < by.size() ? by[i + 2] : f};
            EdwardsPoint p;
            lookup3_xy(cs, chunk, table[i / 3], tv[i / 3], p.u, p.v);
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
                F one = P::one();
                // one inversion for 1 / y and 1 / (x + 1)
                F xp = x.value + one;
                F i = P::inverse(y.value * xp);
                cs.fail_if(P::is_zero(i));
                r = {x.value * P::lift(T().mont_scale) * (i * xp), (x.value - one) * (i * y.value)};
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
        MontgomeryPoint add(CS& cs, const MontgomeryPoint& o, const F* lam_hint = nullptr) const {
            F lv;
            if (lam_hint) {
                lv = *lam_hint;
            } else {
                F i = P::inverse(o.x.value - x.value);
                cs.fail_if(P::is_zero(i));
                lv = (o.y.value - y.value) * i;
            }
            AllocatedNum lam = AllocatedNum::alloc(cs, lv);
            {
                LC a = o.x.lc, c = o.y.lc;
                a.sub(x.lc);
                c.sub(y.lc);
                MBH_ENFORCE(cs, a, LC(lam.var, K().one), c);
            }
            AllocatedNum xprime = AllocatedNum::alloc(cs, lam.value.square() - P::lift(T().mont_a) - x.value - o.x.value);
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

    static EdwardsPoint pedersen_hash(CS& cs, const bool personalization[6], const Bits& bits) {
        Bits all;
        for (int i = 0; i < 6; ++i) all.push_back(Boolean::constant(personalization[i]));
        all.insert(all.end(), bits.begin(), bits.end());
        Boolean f = Boolean::constant(false);
        // pre-pass over the whole hash: window points from the tables, then every division it needs
        const size_t nw = (all.size() + 2) / 3;
        std::vector<F> tx(nw), ty(nw);
        std::vector<size_t> seg_len;
        for (size_t k = 0; k < nw; ++k) {
            size_t seg = k / 63, w = k % 63, q = 3 * k;
            if (w == 0) seg_len.push_back(0);
            ++seg_len.back();
            const Window4V& win = T().pedersen[seg][w];
            B b0 = all[q].value(), b1 = q + 1 < all.size() ? all[q + 1].value() : P::ball(false),
              b2 = q + 2 < all.size() ? all[q + 2].value() : P::ball(false);
            tx[k] = lookup4(b0, b1, win.x);
            F y = lookup4(b0, b1, win.y);
            ty[k] = P::select(b2, -y, y);
        }
        PedersenPlan plan;
        pedersen_plan(tx, ty, seg_len, plan);
        cs.fail_if(plan.bad);

        EdwardsPoint edwards_result;
        size_t pos = 0, k = 0;
        for (size_t seg = 0; seg < seg_len.size(); ++seg) {
            MontgomeryPoint segment_result;
            const std::vector<Window4>& windows = JJ().pedersen[seg];
            const std::vector<Window4V>& windows_v = T().pedersen[seg];
            for (size_t w = 0; w < seg_len[seg]; ++w, ++k) {
                Boolean chunk[3] = {all[pos], pos + 1 < all.size() ? all[pos + 1] : f, pos + 2 < all.size() ? all[pos + 2] : f};
                pos += 3;
                MontgomeryPoint tmp;
                lookup3_xy_with_conditional_negation(cs, chunk, windows[w], windows_v[w], tmp.x, tmp.y);
                segment_result = w ? tmp.add(cs, segment_result, &plan.lam[k]) : tmp;
            }
            EdwardsPoint se = segment_result.into_edwards(cs, &plan.seg_ed[seg]);
            edwards_result = seg ? se.add(cs, edwards_result, &plan.run_ed[seg]) : se;
        }
        return edwards_result;
    }

    // ---------------------------------------------------------------------------
    // circuits
    // ---------------------------------------------------------------------------
    struct AuthNode {
        F sibling;
        B is_right;
    };

    // sapling.rs:71-137
    static void expose_value_commitment(CS& cs, const JPoint& asset_generator, U value, const W& rcv,
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

    static Num value_num_of(const Bits& value_bits) {
        Num n;
        for (size_t i = 0; i < value_bits.size(); ++i) n.add_bool_with_coeff(value_bits[i], K().pow2[i], T().pow2[i]);
        return n;
    }

    static void assert_not_small_order(CS& cs, const EdwardsPoint& p) {
        EdwardsPoint t = p.dbl(cs);
        t = t.dbl(cs);
        t = t.dbl(cs);
        t.u.assert_nonzero(cs);
    }

    static Bits merkle_and_anchor(CS& cs, AllocatedNum cur, const std::vector<AuthNode>& path, const F& anchor,
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
        U value;
        W rcv;
        F anchor;
        std::vector<AuthNode> path;
    };
    static void convert_circuit(CS& cs, const ConvertWitness& w) {
        Bits ag_bits, value_bits;
        expose_value_commitment(cs, w.asset_generator, w.value, w.rcv, ag_bits, value_bits);
        Num value_num = value_num_of(value_bits);
        EdwardsPoint cm = pedersen_hash(cs, NOTE_COMMITMENT_PERSONALIZATION, ag_bits);
        merkle_and_anchor(cs, cm.u, w.path, w.anchor, value_num);
    }

    struct SpendWitness {
        JPoint ak, g_d, asset_generator;
        W nsk, rcv, rcm, ar;
        U value;
        F anchor;
        std::vector<AuthNode> path;
    };
    static void spend_circuit(CS& cs, const SpendWitness& w) {
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
        ivk.resize(251, Boolean::constant(false));
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
        W asset_identifier;  // 32 bytes, little-endian bit order within each byte = bit i of the 256-bit integer
        JPoint asset_generator, g_d, pk_d;
        W rcv, rcm, esk;
        U value;
    };
    static void output_circuit(CS& cs, const OutputWitness& w) {
        Bits preimage;
        for (int i = 0; i < 256; ++i) preimage.push_back(Boolean::from_bit(AllocatedBit::alloc(cs, P::wbit(w.asset_identifier, i))));
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
        B pkv[255], pku[1];
        P::to_bits(w.pk_d.v, 255, pkv);
        P::to_bits(w.pk_d.u, 1, pku);
        Bits v_contents;
        for (int i = 0; i < 255; ++i) v_contents.push_back(Boolean::from_bit(AllocatedBit::alloc(cs, pkv[i])));
        Boolean sign_bit = Boolean::from_bit(AllocatedBit::alloc(cs, pku[0]));
        note.insert(note.end(), v_contents.begin(), v_contents.end());
        note.push_back(sign_bit);
        EdwardsPoint cm = pedersen_hash(cs, NOTE_COMMITMENT_PERSONALIZATION, note);
        Bits rcm_bits = words_into_boolean_vec_le(cs, w.rcm, 252);
        EdwardsPoint rcm_p = fixed_base_multiplication(cs, Jubjub::NOTE_COMMITMENT_RANDOMNESS, rcm_bits);
        cm = cm.add(cs, rcm_p);
        cm.u.inputize(cs);
    }
};

// ---------------------------------------------------------------------------
// witness decoding (layout documented in include/masp_b200.h), P::LANES witnesses at a time
// ---------------------------------------------------------------------------
enum { CIRCUIT_SPEND = 0, CIRCUIT_OUTPUT = 1, CIRCUIT_CONVERT = 2 };  // = MB200_CIRCUIT_*

template <class P>
struct WitnessReader {
    static constexpr int N = P::LANES;
    typedef G<P> Gp;
    const uint8_t* p[N];
    bool bad[N];  // per lane: a field is out of range
    explicit WitnessReader(const uint8_t* const* w) {
        for (int k = 0; k < N; ++k) {
            p[k] = w[k];
            bad[k] = false;
        }
    }
    void raw(uint64_t w[N][4]) {
        for (int k = 0; k < N; ++k) {
            memcpy(w[k], p[k], 32);
            p[k] += 32;
        }
    }
    typename P::F fr() {
        uint64_t w[N][4];
        raw(w);
        for (int k = 0; k < N; ++k)
            if (Fr::geq_mod(w[k])) {
                bad[k] = true;
                memset(w[k], 0, 32);
            }
        return P::from_words(w);
    }
    typename Gp::JPoint point() {
        typename Gp::JPoint q;
        q.u = fr();
        q.v = fr();
        return q;
    }
    typename P::W words() {
        typename P::W r;
        raw(r.w);
        return r;
    }
    // a Jubjub scalar: the circuits allocate 252 bits for it (JUBJUB_FR_BITS), anything above is an error
    typename P::W jscalar() {
        typename P::W r = words();
        for (int k = 0; k < N; ++k)
            if (r.w[k][3] >> 60) bad[k] = true;
        return r;
    }
    void u64_raw(uint64_t v[N]) {
        uint64_t w[N][4];
        raw(w);
        for (int k = 0; k < N; ++k) {
            if (w[k][1] | w[k][2] | w[k][3]) bad[k] = true;
            v[k] = w[k][0];
        }
    }
    typename P::U u64() {
        uint64_t v[N];
        u64_raw(v);
        return P::ufrom(v);
    }
    void path(uint32_t depth, std::vector<typename Gp::AuthNode>& out) {
        out.resize(depth);
        for (uint32_t i = 0; i < depth; ++i) {
            out[i].sibling = fr();
            uint64_t v[N];
            u64_raw(v);
            bool b[N];
            for (int k = 0; k < N; ++k) {
                if (v[k] > 1) bad[k] = true;
                b[k] = v[k] != 0;
            }
            out[i].is_right = P::bfrom(b);
        }
    }
};

// Runs the circuit over P::LANES witnesses; bad_out[k]: witness k has a field out of range (its lane
// ran on zeros in that field's place and its outputs mean nothing).
template <class P>
void run_circuit_lanes(typename P::Sink& cs, int kind, uint32_t depth, const uint8_t* const* w, bool* bad_out) {
    typedef G<P> Gp;
    WitnessReader<P> r(w);
    if (kind == CIRCUIT_SPEND) {
        typename Gp::SpendWitness s;
        s.ak = r.point();
        s.nsk = r.jscalar();
        s.g_d = r.point();
        s.asset_generator = r.point();
        s.value = r.u64();
        s.rcv = r.jscalar();
        s.rcm = r.jscalar();
        s.ar = r.jscalar();
        s.anchor = r.fr();
        r.path(depth, s.path);
        Gp::spend_circuit(cs, s);
    } else if (kind == CIRCUIT_OUTPUT) {
        typename Gp::OutputWitness o;
        o.asset_identifier = r.words();
        o.asset_generator = r.point();
        o.value = r.u64();
        o.rcv = r.jscalar();
        o.g_d = r.point();
        o.pk_d = r.point();
        o.rcm = r.jscalar();
        o.esk = r.jscalar();
        Gp::output_circuit(cs, o);
    } else {
        typename Gp::ConvertWitness c;
        c.asset_generator = r.point();
        c.value = r.u64();
        c.rcv = r.jscalar();
        c.anchor = r.fr();
        r.path(depth, c.path);
        Gp::convert_circuit(cs, c);
    }
    for (int k = 0; k < P::LANES; ++k) bad_out[k] = r.bad[k];
}

typedef G<ScalarPolicy> GS;

}  // namespace mbh
