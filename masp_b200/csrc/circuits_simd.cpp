// Eight witnesses per thread: the gadget library of host/gadgets.hpp instantiated over AVX-512 IFMA
// lanes (host/fr8_host.hpp).  This is the only translation unit compiled with -mavx512f
// -mavx512ifma; csrc/circuits.cu checks the CPU before calling into it and otherwise keeps the
// one-witness-at-a-time scalar path.  Same allocation order, same values: tests compare the two
// byte for byte (tests/test_circuits_product.py::test_simd_witnesses_equal_scalar_witnesses).
//
// Why: Circuit::synthesize is part of create_random_proof in the reference
// (masp_proofs/src/sapling/prover.rs:96-117; benches/sapling.rs:71-85 time it), and one host
// thread of the scalar generator makes ~90 Spend witnesses a second -- 16 threads feed 2.5 B200s,
// not 8.  A field multiplication here costs 6.2 ns per lane against 32.6 ns scalar.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__AVX512F__) && defined(__AVX512IFMA__)
#include "host/fr8_host.hpp"
#include "host/gadgets.hpp"

namespace mbh {

// the sink of a SIMD pass: values leave at once, converted to the canonical 32 little-endian
// bytes the prover reads, one stream per lane; nothing is kept
struct SimdCS {
    uint8_t* inputs_out[8];
    uint8_t* aux_out[8];
    uint32_t n_aux = 0, n_inputs = 0;
    size_t n_constraints = 0;
    Mask8 failed;
    Fr8 root = Fr8::zero();

    static void emit(uint8_t* const out[8], uint32_t index, const Fr8& v) {
        uint64_t w[8][4];
        v.to_words(w);
        for (int k = 0; k < 8; ++k) memcpy(out[k] + 32 * (size_t)index, w[k], 32);
    }
    Var alloc(const Fr8& v) {
        emit(aux_out, n_aux, v);
        return Var{(n_aux++) | Var::AUX};
    }
    Var alloc_bit(Mask8 b) {  // 70 % of a Spend witness: no field arithmetic at all
        const __m256i z = _mm256_setzero_si256();
        for (int k = 0; k < 8; ++k) {
            uint8_t* o = aux_out[k] + 32 * (size_t)n_aux;
            _mm256_storeu_si256((__m256i*)o, z);
            o[0] = (uint8_t)((b.m >> k) & 1);
        }
        return Var{(n_aux++) | Var::AUX};
    }
    Var alloc_input(const Fr8& v) {
        emit(inputs_out, n_inputs, v);
        return Var{n_inputs++};
    }
    void fail_if(Mask8 b) { failed = failed | b; }
    void enforce(const NullLC&, const NullLC&, const NullLC&) { ++n_constraints; }
};

struct SimdPolicy {
    typedef Fr8 F;
    typedef Mask8 B;
    typedef U8x U;
    typedef FrC8 C;
    typedef NullLC L;
    typedef SimdCS Sink;
    typedef WordsN<8> W;
    static constexpr bool RECORDS = false;
    static constexpr int LANES = 8;
    static C konst(const Fr& f) { return frc8_of(f); }
    static F lift(const C& c) { return Fr8::splat(c); }
    static F zero() { return Fr8::zero(); }
    static F one() { return Fr8::one(); }
    static B ball(bool b) { return Mask8::all(b); }
    static bool any(B b) { return b.any(); }
    static F select(B c, const F& a, const F& b) { return Fr8::select(c, a, b); }
    static F mask(B c, const C& a) {
        Fr8 r;
        for (int i = 0; i < 5; ++i) r.l[i] = _mm512_maskz_set1_epi64(c.m, (long long)a.l[i]);
        return r;
    }
    static B is_zero(const F& a) { return a.is_zero(); }
    static F inverse(const F& a) { return a.inverse(); }
    static void to_bits(const F& v, int n, B* out) {
        Fr8 c = v.to_canonical();
        for (int i = 0; i < n; ++i) out[i] = c.canonical_bit(i);
    }
    static F from_words(const uint64_t w[8][4]) { return Fr8::from_words(w); }
    static B wbit(const W& w, int i) {
        unsigned m = 0;
        for (int k = 0; k < 8; ++k) m |= (unsigned)((w.w[k][i >> 6] >> (i & 63)) & 1) << k;
        return Mask8((__mmask8)m);
    }
    static U ufrom(const uint64_t v[8]) { return U8x(_mm512_loadu_si512(v)); }
    static B bfrom(const bool b[8]) {
        unsigned m = 0;
        for (int k = 0; k < 8; ++k) m |= (unsigned)(b[k] ? 1 : 0) << k;
        return Mask8((__mmask8)m);
    }
    static U uzero() { return U8x(); }
    static U uset(U v, B b, int i) { return v.with_bit(b, i); }
    static B ubit(U v, int i) { return v.bit(i); }
    static U uadd(U a, U b) { return a + b; }
};

}  // namespace mbh

extern "C" {

// Eight witnesses of one circuit -> their inputs and aux assignments.  status[k]: 0 ok, 1 a field of
// witness k is out of range, 2 its synthesis failed (bellman's SynthesisError), 3 the counts do not
// match the recorded circuit.  Returns 0, or -1 on an allocation failure.
int mbh_simd_synthesize8(int kind, uint32_t depth, const uint8_t* const witnesses[8], uint8_t* const inputs_out[8],
                         uint8_t* const aux_out[8], uint32_t expect_inputs, uint32_t expect_aux, uint8_t status[8]) {
    using namespace mbh;
    try {
        SimdCS cs;
        for (int k = 0; k < 8; ++k) {
            cs.inputs_out[k] = inputs_out[k];
            cs.aux_out[k] = aux_out[k];
        }
        cs.alloc_input(Fr8::one());  // ONE
        bool bad[8];
        run_circuit_lanes<SimdPolicy>(cs, kind, depth, witnesses, bad);
        for (int k = 0; k < 8; ++k) {
            status[k] = 0;
            if ((cs.failed.m >> k) & 1) status[k] = 2;
            if (bad[k]) status[k] = 1;
            if (cs.n_inputs != expect_inputs || cs.n_aux != expect_aux) status[k] = 3;
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

void mbh_simd_warm(void) {  // build the SIMD-domain tables before worker threads race for them
    (void)mbh::G<mbh::SimdPolicy>::T();
    (void)mbh::fr8k();
}

int mbh_simd_compiled(void) { return 1; }

}  // extern "C"

#else  // the compiler was not given AVX-512 IFMA: the scalar generator is the only one

extern "C" {
int mbh_simd_synthesize8(int, uint32_t, const uint8_t* const*, uint8_t* const*, uint8_t* const*, uint32_t, uint32_t, uint8_t*) {
    return -1;
}
void mbh_simd_warm(void) {}
int mbh_simd_compiled(void) { return 0; }
}
#endif
