// Host-side BLS12-381 scalar field (= the Jubjub base field) for witness
// generation: 4 x 64-bit Montgomery limbs.
//
// Stands where the reference uses `bls12_381::Scalar` (nam-blstrs) inside
// circuit synthesis (masp_proofs/src/circuit/*.rs).  Product code: the oracle
// has its own, independent implementation.
#pragma once
#include <cstdint>
#include <cstring>

namespace mbh {

typedef unsigned __int128 u128;

struct Fr {
    uint64_t v[4];  // Montgomery form, < r

    static constexpr uint64_t M0 = 0xffffffff00000001ull, M1 = 0x53bda402fffe5bfeull, M2 = 0x3339d80809a1d805ull,
                              M3 = 0x73eda753299d7d48ull;
    static constexpr uint64_t INV = 0xfffffffeffffffffull;  // -r^-1 mod 2^64
    static const uint64_t* mod() {
        static const uint64_t m[4] = {M0, M1, M2, M3};
        return m;
    }
    static Fr raw(uint64_t a, uint64_t b, uint64_t c, uint64_t d) { return Fr{{a, b, c, d}}; }
    static Fr zero() { return raw(0, 0, 0, 0); }
    static Fr one() { return raw(0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full); }
    static Fr r2() { return raw(0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull); }
    static Fr r3() { return raw(0xc62c1807439b73afull, 0x1b3e0d188cf06990ull, 0x73d13c71c7b5f418ull, 0x6e2a5bb9c8db33e9ull); }

    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
    bool operator==(const Fr& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    bool operator!=(const Fr& o) const { return !(*this == o); }

    static bool geq_mod(const uint64_t* a) {
        const uint64_t* m = mod();
        for (int i = 3; i >= 0; --i) {
            if (a[i] > m[i]) return true;
            if (a[i] < m[i]) return false;
        }
        return true;
    }
    static void sub_mod(uint64_t* a) {
        const uint64_t* m = mod();
        u128 br = 0;
        for (int i = 0; i < 4; ++i) {
            u128 t = (u128)a[i] - m[i] - (uint64_t)br;
            a[i] = (uint64_t)t;
            br = (t >> 64) & 1;
        }
    }
    friend Fr operator+(const Fr& a, const Fr& b) {
        Fr r;
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)a.v[i] + b.v[i];
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
        if (geq_mod(r.v)) sub_mod(r.v);  // 2r < 2^256: no carry out
        return r;
    }
    friend Fr operator-(const Fr& a, const Fr& b) {
        Fr r;
        u128 br = 0;
        for (int i = 0; i < 4; ++i) {
            u128 t = (u128)a.v[i] - b.v[i] - (uint64_t)br;
            r.v[i] = (uint64_t)t;
            br = (t >> 64) & 1;
        }
        if (br) {
            const uint64_t* m = mod();
            u128 c = 0;
            for (int i = 0; i < 4; ++i) {
                c += (u128)r.v[i] + m[i];
                r.v[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return r;
    }
    Fr operator-() const { return zero() - *this; }
    static inline uint64_t mac(uint64_t acc, uint64_t x, uint64_t y, uint64_t& carry) {
        u128 t = (u128)x * y + acc + carry;
        carry = (uint64_t)(t >> 64);
        return (uint64_t)t;
    }
    friend Fr operator*(const Fr& a, const Fr& b) {
        // 4 x 4 schoolbook product, then four Montgomery rounds (fully unrolled by the compiler)
        uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma GCC unroll 4
        for (int i = 0; i < 4; ++i) {
            uint64_t c = 0;
#pragma GCC unroll 4
            for (int j = 0; j < 4; ++j) t[i + j] = mac(t[i + j], a.v[j], b.v[i], c);
            t[i + 4] = c;
        }
        uint64_t top = 0;
#pragma GCC unroll 4
        for (int i = 0; i < 4; ++i) {
            uint64_t q = t[i] * INV, c = 0;
            mac(t[i], q, M0, c);
            t[i + 1] = mac(t[i + 1], q, M1, c);
            t[i + 2] = mac(t[i + 2], q, M2, c);
            t[i + 3] = mac(t[i + 3], q, M3, c);
            u128 s = (u128)t[i + 4] + c + top;
            t[i + 4] = (uint64_t)s;
            top = (uint64_t)(s >> 64);
        }
        Fr r = raw(t[4], t[5], t[6], t[7]);
        if (top || geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    Fr& operator+=(const Fr& o) { return *this = *this + o; }
    Fr& operator-=(const Fr& o) { return *this = *this - o; }
    Fr& operator*=(const Fr& o) { return *this = *this * o; }
    Fr square() const { return *this * *this; }
    Fr dbl() const { return *this + *this; }

    static Fr from_u64(uint64_t x) { return raw(x, 0, 0, 0) * r2(); }
    // canonical little-endian integer
    void to_words(uint64_t out[4]) const {
        Fr t = *this * raw(1, 0, 0, 0);
        memcpy(out, t.v, 32);
    }
    void to_bytes(uint8_t out[32]) const {
        // most witness values are booleans: 0 and 1 need no Montgomery reduction
        if (is_zero()) {
            memset(out, 0, 32);
            return;
        }
        if (v[0] == 0x00000001fffffffeull && v[1] == 0x5884b7fa00034802ull && v[2] == 0x998c4fefecbc4ff5ull &&
            v[3] == 0x1824b159acc5056full) {
            memset(out, 0, 32);
            out[0] = 1;
            return;
        }
        uint64_t w[4];
        to_words(w);
        memcpy(out, w, 32);  // little-endian host
    }
    // false if the encoding is not canonical (>= r)
    static bool from_bytes(const uint8_t in[32], Fr& out) {
        uint64_t w[4];
        memcpy(w, in, 32);
        if (geq_mod(w)) return false;
        out = raw(w[0], w[1], w[2], w[3]) * r2();
        return true;
    }
    static Fr from_words(const uint64_t w[4]) { return raw(w[0], w[1], w[2], w[3]) * r2(); }

    // 2^k in Montgomery form, k < 256
    static Fr pow2_mont(unsigned k) {
        uint64_t w[4] = {0, 0, 0, 0};
        w[k >> 6] = 1ull << (k & 63);
        return raw(w[0], w[1], w[2], w[3]) * r2();
    }
    // a^-1 (0 -> 0).  Kaliski's almost-inverse on the Montgomery representative x = a R:
    // phase 1 yields x^-1 2^k (255 <= k <= 510) with shifts and additions only; the power of
    // two is then folded into the Montgomery factor: a^-1 R = x^-1 2^k * 2^(512 - k).
    Fr inverse() const {
        if (is_zero()) return zero();
        uint64_t u[4] = {M0, M1, M2, M3}, w[4], r[4] = {0, 0, 0, 0}, s[4] = {1, 0, 0, 0};
        memcpy(w, v, 32);
        unsigned k = 0;
        auto shr1 = [](uint64_t* a) {
            a[0] = (a[0] >> 1) | (a[1] << 63);
            a[1] = (a[1] >> 1) | (a[2] << 63);
            a[2] = (a[2] >> 1) | (a[3] << 63);
            a[3] >>= 1;
        };
        auto shl1 = [](uint64_t* a) {
            a[3] = (a[3] << 1) | (a[2] >> 63);
            a[2] = (a[2] << 1) | (a[1] >> 63);
            a[1] = (a[1] << 1) | (a[0] >> 63);
            a[0] <<= 1;
        };
        auto add = [](uint64_t* a, const uint64_t* b) {
            u128 c = 0;
            for (int i = 0; i < 4; ++i) {
                c += (u128)a[i] + b[i];
                a[i] = (uint64_t)c;
                c >>= 64;
            }
        };
        auto sub = [](uint64_t* out, const uint64_t* a, const uint64_t* b) {  // out = a - b, returns borrow
            u128 br = 0;
            for (int i = 0; i < 4; ++i) {
                u128 t = (u128)a[i] - b[i] - (uint64_t)br;
                out[i] = (uint64_t)t;
                br = (t >> 64) & 1;
            }
            return (uint64_t)br;
        };
        while ((w[0] | w[1] | w[2] | w[3]) != 0) {
            if (!(u[0] & 1)) {
                shr1(u);
                shl1(s);
            } else if (!(w[0] & 1)) {
                shr1(w);
                shl1(r);
            } else {
                uint64_t d[4];
                if (!sub(d, u, w)) {  // u >= w (u == w only when both are 1)
                    if ((d[0] | d[1] | d[2] | d[3]) == 0) {  // u == w: the v > u branch of the paper
                        sub(w, w, u);
                        shr1(w);
                        add(s, r);
                        shl1(r);
                    } else {
                        memcpy(u, d, 32);
                        shr1(u);
                        add(r, s);
                        shl1(s);
                    }
                } else {
                    sub(w, w, u);
                    shr1(w);
                    add(s, r);
                    shl1(r);
                }
            }
            ++k;
        }
        // r < 2 p
        if (geq_mod(r)) sub_mod(r);
        uint64_t x[4];
        sub(x, mod(), r);  // p - r = x^-1 2^k mod p
        Fr y = raw(x[0], x[1], x[2], x[3]);
        // multiply by 2^(512 - k): two Montgomery products with 2^j R
        unsigned e = 512 - k;  // 2 .. 257
        // y * (2^e) : mont(y, 2^e R) = y 2^e
        if (e >= 256) {
            y = y * pow2_mont(255);
            e -= 255;
        }
        return y * pow2_mont(e);
    }
    Fr pow(const uint64_t e[4]) const {
        Fr r = one();
        for (int i = 255; i >= 0; --i) {
            r = r.square();
            if ((e[i >> 6] >> (i & 63)) & 1) r = r * *this;
        }
        return r;
    }
};

}  // namespace mbh
