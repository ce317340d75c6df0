// BLAKE2b-512, unkeyed (RFC 7693), for the parameter-file check of
// masp_proofs/src/lib.rs:343-388: the reference hashes the WHOLE stream -- the Parameters encoding
// and the MPC transcript behind it -- through a HashReader (blake2b_simd, 64-byte digest) and
// compares the hex digest with MASP_{SPEND,OUTPUT,CONVERT}_HASH (lib.rs:70-72).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace mbh {

struct Blake2b {
    uint64_t h[8];
    uint8_t buf[128];
    size_t buflen = 0;
    uint64_t t0 = 0, t1 = 0;
    static const uint64_t* iv() {
        static const uint64_t v[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull,
                                      0xa54ff53a5f1d36f1ull, 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full,
                                      0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        return v;
    }
    Blake2b() {
        for (int i = 0; i < 8; ++i) h[i] = iv()[i];
        h[0] ^= 0x01010000ull ^ 64ull;  // digest length 64, no key, fanout = depth = 1
    }
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    void compress(const uint8_t* block, bool last) {
        static const uint8_t sigma[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        memcpy(m, block, 128);  // little-endian host
        for (int i = 0; i < 8; ++i) {
            v[i] = h[i];
            v[i + 8] = iv()[i];
        }
        v[12] ^= t0;
        v[13] ^= t1;
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x;
            v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y;
            v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; ++r) {
            const uint8_t* s = sigma[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);
            G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);
            G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);
            G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]);
            G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
    }
    void add_count(uint64_t n) {
        t0 += n;
        if (t0 < n) ++t1;
    }
    void update(const uint8_t* p, size_t n) {
        while (n) {
            if (buflen == 128) {  // a full buffer is only compressed once more data follows: the last block is special
                add_count(128);
                compress(buf, false);
                buflen = 0;
            }
            size_t k = n < 128 - buflen ? n : 128 - buflen;
            memcpy(buf + buflen, p, k);
            buflen += k;
            p += k;
            n -= k;
        }
    }
    void finish(uint8_t out[64]) {
        add_count(buflen);
        memset(buf + buflen, 0, 128 - buflen);
        compress(buf, true);
        memcpy(out, h, 64);  // little-endian host
    }
};

inline void blake2b512_hex(const uint8_t* p, size_t n, char out_hex[129]) {
    Blake2b b;
    b.update(p, n);
    uint8_t d[64];
    b.finish(d);
    static const char* hx = "0123456789abcdef";
    for (int i = 0; i < 64; ++i) {
        out_hex[2 * i] = hx[d[i] >> 4];
        out_hex[2 * i + 1] = hx[d[i] & 15];
    }
    out_hex[128] = 0;
}

}  // namespace mbh
