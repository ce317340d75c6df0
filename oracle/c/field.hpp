// TEST INFRASTRUCTURE ONLY (oracle / CPU baseline).  Nothing under masp_b200/
// may include, link or load this.
//
// PARITY UNPINNED (SURVEY.md finding 3): restates the public BLS12-381 field
// arithmetic that the reference obtains from nam-blstrs 0.7.1-nam.0 over
// nam-blst 0.3.15-nam.0 (reference Cargo.lock:1385-1411; selected at
// masp_proofs/Cargo.toml:22).  64-bit limbs, Montgomery form, portable
// `unsigned __int128` (blst uses hand-written ADX/MULX assembly for the same
// operations).  Cross-checked byte-for-byte against oracle/py (Python ints).
#pragma once
#include <cstdint>
#include <cstring>

typedef unsigned __int128 u128;

template <int N>
struct BigN {
    uint64_t v[N];
};

template <int N>
static inline int bn_cmp(const uint64_t* a, const uint64_t* b) {
    for (int i = N - 1; i >= 0; --i) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
template <int N>
static inline uint64_t bn_add(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 c = 0;
    for (int i = 0; i < N; ++i) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
template <int N>
static inline uint64_t bn_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t borrow = 0;
    for (int i = 0; i < N; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}

// A prime field with an N-limb modulus supplied by Params::MOD (little-endian).
template <class Params>
struct Field {
    static constexpr int N = Params::N;
    uint64_t v[N];

    static inline const uint64_t* mod() { return Params::MOD; }
    static uint64_t INV;       // -p^{-1} mod 2^64
    static Field R1, R2;       // R mod p, R^2 mod p
    static bool ready;

    static void init() {
        if (ready) return;
        uint64_t inv = 1;
        for (int i = 0; i < 6; ++i) inv *= 2 - Params::MOD[0] * inv;  // Newton
        INV = (uint64_t)0 - inv;
        Field t;
        memset(t.v, 0, sizeof t.v);
        t.v[0] = 1;
        for (int i = 0; i < 64 * N; ++i) t.dbl_raw();
        R1 = t;
        for (int i = 0; i < 64 * N; ++i) t.dbl_raw();
        R2 = t;
        ready = true;
    }
    void dbl_raw() {  // value doubling mod p on plain integers
        uint64_t c = bn_add<N>(v, v, v);
        if (c || bn_cmp<N>(v, mod()) >= 0) bn_sub<N>(v, v, mod());
    }

    static Field zero() { Field r; memset(r.v, 0, sizeof r.v); return r; }
    static Field one() { return R1; }
    bool is_zero() const {
        uint64_t o = 0;
        for (int i = 0; i < N; ++i) o |= v[i];
        return o == 0;
    }
    bool operator==(const Field& o) const { return memcmp(v, o.v, sizeof v) == 0; }
    bool operator!=(const Field& o) const { return !(*this == o); }

    static inline Field add(const Field& a, const Field& b) {
        Field r;
        uint64_t c = bn_add<N>(r.v, a.v, b.v);
        if (c || bn_cmp<N>(r.v, mod()) >= 0) bn_sub<N>(r.v, r.v, mod());
        return r;
    }
    static inline Field sub(const Field& a, const Field& b) {
        Field r;
        if (bn_sub<N>(r.v, a.v, b.v)) bn_add<N>(r.v, r.v, mod());
        return r;
    }
    static inline Field neg(const Field& a) {
        if (a.is_zero()) return a;
        Field r;
        bn_sub<N>(r.v, mod(), a.v);
        return r;
    }
    static inline Field dbl(const Field& a) { return add(a, a); }

    // CIOS Montgomery multiplication.
    static inline Field mul(const Field& a, const Field& b) {
        uint64_t t[N + 2];
        memset(t, 0, sizeof t);
        for (int i = 0; i < N; ++i) {
            u128 c = 0;
            for (int j = 0; j < N; ++j) {
                c += (u128)a.v[j] * b.v[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N] = (uint64_t)c;
            t[N + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * INV;
            c = (u128)m * Params::MOD[0] + t[0];
            c >>= 64;
            for (int j = 1; j < N; ++j) {
                c += (u128)m * Params::MOD[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N - 1] = (uint64_t)c;
            t[N] = t[N + 1] + (uint64_t)(c >> 64);
        }
        Field r;
        memcpy(r.v, t, sizeof r.v);
        if (t[N] || bn_cmp<N>(r.v, mod()) >= 0) bn_sub<N>(r.v, r.v, mod());
        return r;
    }
    static inline Field sqr(const Field& a) { return mul(a, a); }

    Field operator+(const Field& o) const { return add(*this, o); }
    Field operator-(const Field& o) const { return sub(*this, o); }
    Field operator*(const Field& o) const { return mul(*this, o); }

    static Field from_raw(const uint64_t* limbs) {  // plain integer -> Montgomery
        Field t;
        memcpy(t.v, limbs, sizeof t.v);
        return mul(t, R2);
    }
    void to_raw(uint64_t* limbs) const {  // Montgomery -> plain integer
        Field o = zero();
        o.v[0] = 1;
        Field t = mul(*this, o);
        memcpy(limbs, t.v, sizeof t.v);
    }
    static Field from_u64(uint64_t x) {
        uint64_t l[N];
        memset(l, 0, sizeof l);
        l[0] = x;
        return from_raw(l);
    }
    // exponent: plain little-endian limbs
    static Field pow(const Field& a, const uint64_t* e, int nlimbs) {
        Field r = one();
        bool started = false;
        for (int i = nlimbs * 64 - 1; i >= 0; --i) {
            if (started) r = sqr(r);
            if ((e[i / 64] >> (i % 64)) & 1) {
                r = started ? mul(r, a) : a;
                started = true;
            }
        }
        return r;
    }
    static Field inv(const Field& a) {  // Fermat
        uint64_t e[N];
        uint64_t two[N];
        memset(two, 0, sizeof two);
        two[0] = 2;
        bn_sub<N>(e, mod(), two);
        return pow(a, e, N);
    }
    // canonical bytes
    bool from_be_bytes(const uint8_t* b) {  // N*8 bytes big-endian; false if >= p
        uint64_t l[N];
        for (int i = 0; i < N; ++i) {
            uint64_t w = 0;
            for (int k = 0; k < 8; ++k) w = (w << 8) | b[(N - 1 - i) * 8 + k];
            l[i] = w;
        }
        if (bn_cmp<N>(l, mod()) >= 0) return false;
        *this = from_raw(l);
        return true;
    }
    void to_be_bytes(uint8_t* b) const {
        uint64_t l[N];
        to_raw(l);
        for (int i = 0; i < N; ++i)
            for (int k = 0; k < 8; ++k) b[(N - 1 - i) * 8 + k] = (uint8_t)(l[i] >> (56 - 8 * k));
    }
    bool from_le_bytes(const uint8_t* b) {
        uint64_t l[N];
        memcpy(l, b, sizeof l);
        if (bn_cmp<N>(l, mod()) >= 0) return false;
        *this = from_raw(l);
        return true;
    }
    void to_le_bytes(uint8_t* b) const {
        uint64_t l[N];
        to_raw(l);
        memcpy(b, l, sizeof l);
    }
};
template <class P> uint64_t Field<P>::INV;
template <class P> Field<P> Field<P>::R1;
template <class P> Field<P> Field<P>::R2;
template <class P> bool Field<P>::ready = false;

struct FpParams {
    static constexpr int N = 6;
    static constexpr uint64_t MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                        0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
};
struct FrParams {
    static constexpr int N = 4;
    static constexpr uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                        0x73eda753299d7d48ULL};
};
typedef Field<FpParams> Fp;
typedef Field<FrParams> Fr;

// Fp2 = Fp[u]/(u^2+1)
struct Fp2 {
    Fp c0, c1;
    static Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
    static Fp2 one() { return {Fp::one(), Fp::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const Fp2& o) const { return !(*this == o); }
    static Fp2 add(const Fp2& a, const Fp2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
    static Fp2 sub(const Fp2& a, const Fp2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
    static Fp2 neg(const Fp2& a) { return {Fp::neg(a.c0), Fp::neg(a.c1)}; }
    static Fp2 dbl(const Fp2& a) { return add(a, a); }
    static Fp2 mul(const Fp2& a, const Fp2& b) {
        Fp t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
        Fp t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
        return {t0 - t1, t2 - t0 - t1};
    }
    static Fp2 sqr(const Fp2& a) {
        Fp t = a.c0 * a.c1;
        return {(a.c0 + a.c1) * (a.c0 - a.c1), t + t};
    }
    Fp2 operator+(const Fp2& o) const { return add(*this, o); }
    Fp2 operator-(const Fp2& o) const { return sub(*this, o); }
    Fp2 operator*(const Fp2& o) const { return mul(*this, o); }
    static Fp2 inv(const Fp2& a) {
        Fp t = Fp::inv(a.c0 * a.c0 + a.c1 * a.c1);
        return {a.c0 * t, Fp::neg(a.c1 * t)};
    }
};
