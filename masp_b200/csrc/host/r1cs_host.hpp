// Constraint-system sink and the bellman gadget set the MASP circuits use,
// for the host side of the proving path (SURVEY.md §8 a-2 / NEXT-1).
//
// One sink, two uses:
//   * shape pass (record = true), once per circuit: every enforce() is stored
//     as three sparse rows; the matrices go to the GPU (r1cs_eval kernel) and
//     give the density bitmaps and the structural hash;
//   * witness pass (record = false), per proof: only alloc()/alloc_input()
//     values are produced.  Linear combinations are not even built: the per-row
//     evaluations <A_i,z>, <B_i,z>, <C_i,z> that bellman's ProvingAssignment
//     computes on the CPU inside enforce() are a sparse matrix-vector product
//     here, done on the device.
//
// Gadget semantics follow bellman (nam-bellperson / bellpepper-core, reference
// Cargo.lock:154-155, 1355-1358) as restated in SURVEY.md Appendix B; emission
// order is what fixes the variable numbering and hence the key layout, and it
// is pinned by the reference's cs.hash() strings
// (masp_proofs/src/circuit/convert.rs:218-224, sapling.rs:730-741, 1024-1045).
#pragma once
#include <algorithm>
#include <string>
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

struct Consts {
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

struct Matrix {  // CSR over the variable ids
    std::vector<uint32_t> rowptr{0};
    std::vector<uint32_t> col;
    std::vector<Fr> coef;
};

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
    Var alloc_input(const Fr& v) {
        inputs.push_back(v);
        return Var{(uint32_t)(inputs.size() - 1)};
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
// In a witness pass the three linear combinations are not even constructed.
#define MBH_ENFORCE(cs, a, b, c)          \
    do {                                  \
        if (::mbh::recording()) (cs).enforce(a, b, c); \
        else ++(cs).n_constraints;        \
    } while (0)

// ---------------------------------------------------------------------------
// booleans
// ---------------------------------------------------------------------------
struct AllocatedBit {
    Var var;
    bool value;
    static AllocatedBit alloc(CS& cs, bool value) {
        Var v = cs.alloc(value ? K().one : Fr::zero());
        MBH_ENFORCE(cs, LC(ONE, K().one).sub(v), LC(v, K().one), LC());
        return {v, value};
    }
    static AllocatedBit alloc_conditionally(CS& cs, bool value, const AllocatedBit& must_be_false) {
        Var v = cs.alloc(value ? K().one : Fr::zero());
        MBH_ENFORCE(cs, LC(ONE, K().one).sub(must_be_false.var).sub(v), LC(v, K().one), LC());
        return {v, value};
    }
    static AllocatedBit and_(CS& cs, const AllocatedBit& a, const AllocatedBit& b) {
        bool val = a.value && b.value;
        Var v = cs.alloc(val ? K().one : Fr::zero());
        MBH_ENFORCE(cs, LC(a.var, K().one), LC(b.var, K().one), LC(v, K().one));
        return {v, val};
    }
    static AllocatedBit and_not(CS& cs, const AllocatedBit& a, const AllocatedBit& b) {
        bool val = a.value && !b.value;
        Var v = cs.alloc(val ? K().one : Fr::zero());
        MBH_ENFORCE(cs, LC(a.var, K().one), LC(ONE, K().one).sub(b.var), LC(v, K().one));
        return {v, val};
    }
    static AllocatedBit nor(CS& cs, const AllocatedBit& a, const AllocatedBit& b) {
        bool val = !a.value && !b.value;
        Var v = cs.alloc(val ? K().one : Fr::zero());
        MBH_ENFORCE(cs, LC(ONE, K().one).sub(a.var), LC(ONE, K().one).sub(b.var), LC(v, K().one));
        return {v, val};
    }
    static AllocatedBit xor_(CS& cs, const AllocatedBit& a, const AllocatedBit& b) {
        bool val = a.value != b.value;
        Var v = cs.alloc(val ? K().one : Fr::zero());
        MBH_ENFORCE(cs, LC(a.var, K().two), LC(b.var, K().one), LC(a.var, K().one).add(b.var).sub(v));
        return {v, val};
    }
};

struct Boolean {
    enum Kind : uint8_t { IS, NOT, CONST } kind;
    AllocatedBit bit;
    bool c;
    static Boolean constant(bool b) { return {CONST, {{0}, false}, b}; }
    static Boolean from_bit(const AllocatedBit& b) { return {IS, b, false}; }
    bool value() const { return kind == CONST ? c : (kind == IS ? bit.value : !bit.value); }
    bool is_const() const { return kind == CONST; }
    Boolean not_() const {
        if (kind == CONST) return constant(!c);
        return {kind == IS ? NOT : IS, bit, false};
    }
    // lc += k * self
    void add_to(LC& lc, const Fr& k) const {
        if (!recording()) return;
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
            if (a.c != b.c) cs.failed = true;
            return;
        }
        LC c = a.lc(K().one);
        c.sub(b.lc(K().one));
        MBH_ENFORCE(cs, LC(), LC(), c);
    }
};
typedef std::vector<Boolean> Bits;

inline bool word_bit(const uint64_t w[4], int i) { return (w[i >> 6] >> (i & 63)) & 1; }

inline Bits u64_into_boolean_vec_le(CS& cs, uint64_t value) {
    Bits r;
    for (int i = 0; i < 64; ++i) r.push_back(Boolean::from_bit(AllocatedBit::alloc(cs, (value >> i) & 1)));
    return r;
}
// masp_proofs/src/circuit/gadgets.rs:6-50 on a 256-bit little-endian integer
inline Bits words_into_boolean_vec_le(CS& cs, const uint64_t w[4], int num_bits) {
    Bits r;
    for (int i = 0; i < num_bits; ++i) r.push_back(Boolean::from_bit(AllocatedBit::alloc(cs, word_bit(w, i))));
    return r;
}

// ---------------------------------------------------------------------------
// numbers
// ---------------------------------------------------------------------------
struct AllocatedNum {
    Var var;
    Fr value;
    static AllocatedNum alloc(CS& cs, const Fr& v) { return {cs.alloc(v), v}; }
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
        if (value.is_zero()) cs.failed = true;
        Var inv = cs.alloc(value.inverse());
        MBH_ENFORCE(cs, LC(var, K().one), LC(inv, K().one), LC(ONE, K().one));
    }
    void inputize(CS& cs) const {
        Var inp = cs.alloc_input(value);
        MBH_ENFORCE(cs, LC(inp, K().one), LC(ONE, K().one), LC(var, K().one));
    }
    static void conditionally_reverse(CS& cs, const AllocatedNum& a, const AllocatedNum& b, const Boolean& cond,
                                      AllocatedNum& c, AllocatedNum& d) {
        c = alloc(cs, cond.value() ? b.value : a.value);
        MBH_ENFORCE(cs, LC(a.var, K().one).sub(b.var), cond.lc(K().one), LC(a.var, K().one).sub(c.var));
        d = alloc(cs, cond.value() ? a.value : b.value);
        MBH_ENFORCE(cs, LC(b.var, K().one).sub(a.var), cond.lc(K().one), LC(b.var, K().one).sub(d.var));
    }
    Bits to_bits_le(CS& cs) const {
        uint64_t w[4];
        value.to_words(w);
        Bits bits;
        LC lc;
        for (int i = 0; i < 255; ++i) {
            AllocatedBit b = AllocatedBit::alloc(cs, word_bit(w, i));
            lc.add(b.var, K().pow2[i]);
            bits.push_back(Boolean::from_bit(b));
        }
        lc.sub(var);
        MBH_ENFORCE(cs, LC(), LC(), lc);
        return bits;
    }
    // bits of the value with the proof that they encode an integer <= r - 1
    Bits to_bits_le_strict(CS& cs) const {
        uint64_t w[4], bound[4] = {Fr::M0 - 1, Fr::M1, Fr::M2, Fr::M3};
        value.to_words(w);
        std::vector<AllocatedBit> result;  // big-endian
        std::vector<AllocatedBit> run;
        bool have_last = false;
        AllocatedBit last_run = {{0}, false};
        for (int i = 254; i >= 0; --i) {
            bool a_bit = word_bit(w, i);
            if (word_bit(bound, i)) {
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
    Fr value = Fr::zero();
    static Num from_allocated(const AllocatedNum& n) {
        Num r;
        r.lc.add(n.var, K().one);
        r.value = n.value;
        return r;
    }
    Num& add_bool_with_coeff(const Boolean& bit, const Fr& coeff) {
        bit.add_to(lc, coeff);
        if (bit.value()) value += coeff;
        return *this;
    }
};

// ---------------------------------------------------------------------------
// window lookups
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

inline void lookup3_xy(CS& cs, const Boolean bits[3], const Window8& w, AllocatedNum& res_x, AllocatedNum& res_y) {
    int i = (bits[0].value() ? 1 : 0) | (bits[1].value() ? 2 : 0) | (bits[2].value() ? 4 : 0);
    res_x = AllocatedNum::alloc(cs, w.u[i]);
    res_y = AllocatedNum::alloc(cs, w.v[i]);
    Boolean precomp = Boolean::and_(cs, bits[1], bits[2]);
    for (int k = 0; k < 2; ++k) {
        const Fr* co = k ? w.vc : w.uc;
        const AllocatedNum& res = k ? res_y : res_x;
        LC a, c;
        if (recording()) {
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

inline void lookup3_xy_with_conditional_negation(CS& cs, const Boolean bits[3], const Window4& w, Num& x, Num& y) {
    int i = (bits[0].value() ? 1 : 0) | (bits[1].value() ? 2 : 0);
    Fr yv = w.y[i];
    if (bits[2].value()) yv = -yv;
    AllocatedNum ya = AllocatedNum::alloc(cs, yv);
    Boolean precomp = Boolean::and_(cs, bits[0], bits[1]);
    x = Num();
    x.add_bool_with_coeff(Boolean::constant(true), w.xc[0]);
    x.add_bool_with_coeff(bits[0], w.xc[1]);
    x.add_bool_with_coeff(bits[1], w.xc[2]);
    x.add_bool_with_coeff(precomp, w.xc[3]);
    LC y_lc;
    if (recording()) {
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
inline void pack_into_inputs(CS& cs, const Bits& bits) {
    for (size_t k = 0; k < bits.size(); k += 254) {
        Num num;
        for (size_t j = k; j < bits.size() && j < k + 254; ++j) num.add_bool_with_coeff(bits[j], K().pow2[j - k]);
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
    uint32_t value() const {
        uint32_t v = 0;
        for (int i = 0; i < 32; ++i)
            if (bits[i].value()) v |= 1u << i;
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
        uint64_t total = 0;
        LC lc;
        bool all_constants = true;
        for (int k = 0; k < nops; ++k) {
            total += ops[k]->value();
            for (int i = 0; i < 32; ++i) {
                ops[k]->bits[i].add_to(lc, K().pow2[i]);
                all_constants &= ops[k]->bits[i].is_const();
            }
        }
        if (all_constants) return constant((uint32_t)total);
        UInt32 r;
        LC result_lc;
        int i = 0;
        while (max_value) {
            AllocatedBit b = AllocatedBit::alloc(cs, (total >> i) & 1);
            result_lc.add(b.var, K().pow2[i]);
            if (i < 32) r.bits[i] = Boolean::from_bit(b);
            max_value >>= 1;
            ++i;
        }
        meq.enforce_equal(i, lc, result_lc);
        return r;
    }
};

namespace blake2s_gadget {
static const uint32_t IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
static const uint8_t SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

inline void mixing_g(MultiEq& meq, UInt32* v, int a, int b, int c, int d, const UInt32& x, const UInt32& y) {
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

inline void compression(CS& cs, UInt32 h[8], const UInt32 m[16], uint64_t t, bool final) {
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
}  // namespace blake2s_gadget

// BLAKE2s-256 of a bit string (bytes in little-endian bit order), 8-byte personalization
inline Bits blake2s(CS& cs, const Bits& input, const char personalization[8]) {
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

}  // namespace mbh
