// Synthetic keys and points: every point is (PRNG scalar) * generator.
//
// Stands in for bellman's generate_random_parameters as the reference benches
// use it (masp_proofs/benches/sapling.rs:24-36, benches/convert.rs:21-30):
// a key of the right shape whose values do not matter for cost.  The discrete
// logs are a pure function of (seed, stream, index) -- the same SplitMix64
// counter derivation as masp_b200/synthetic.py -- so tests can check proofs
// in closed form in Fr.
#pragma once
#include "ec.cuh"

namespace mb {

MB_HD uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
MB_HD uint64_t stream_key(uint64_t seed, uint64_t stream) { return mix64(seed + stream * 0xD1342543DE82EF95ull); }

// 255 random bits, minus r if >= r: a plain scalar < r as 8 little-endian limbs
MB_HD void synth_scalar(uint64_t key, uint64_t index, uint32_t* out) {
    Fr t;
    for (int j = 0; j < 4; ++j) {
        uint64_t w = mix64(key + (4 * index + j) * 0x9E3779B97F4A7C15ull);
        if (j == 3) w &= 0x7FFFFFFFFFFFFFFFull;
        t.v[2 * j] = (uint32_t)w;
        t.v[2 * j + 1] = (uint32_t)(w >> 32);
    }
    if (Fr::std_ge_mod(t)) {
        t.v[0] = sub_cc(t.v[0], FrCfg::mod(0));
        for (int i = 1; i < 7; ++i) t.v[i] = subc_cc(t.v[i], FrCfg::mod(i));
        t.v[7] = subc(t.v[7], FrCfg::mod(7));
    }
    for (int i = 0; i < 8; ++i) out[i] = t.v[i];
}

struct SynthArgs {
    size_t nthreads;
    uint64_t key;      // stream key
    uint64_t start;    // first counter index
    uint8_t* out;      // uncompressed encodings, 96 / 192 bytes apart
    G1Affine g1;       // generators, Montgomery form
    G2Affine g2;
};
MB_HD void synth_g1_body(const SynthArgs& a, size_t tid) {
    uint32_t k[8];
    synth_scalar(a.key, a.start + tid, k);
    g1_encode(xyzz_to_affine(xyzz_mul_affine(a.g1, k)), a.out + 96 * tid);
}
MB_HD void synth_g2_body(const SynthArgs& a, size_t tid) {
    uint32_t k[8];
    synth_scalar(a.key, a.start + tid, k);
    g2_encode(xyzz_to_affine(xyzz_mul_affine(a.g2, k)), a.out + 192 * tid);
}
MB_K_G1(synth_g1, SynthArgs, synth_g1_body, 64)
MB_K_G2(synth_g2, SynthArgs, synth_g2_body, 32)

inline void fp_from_hex_host(const char* hex, Fp& out) {  // 96 hex digits, big-endian -> Montgomery
    uint8_t b[48];
    for (int i = 0; i < 48; ++i) {
        unsigned v = 0;
        sscanf(hex + 2 * i, "%2x", &v);
        b[i] = (uint8_t)v;
    }
    Fp t;
    fp_limbs_from_be(b, t);
    out = Fp::from_std(t);
}
inline G1Affine g1_generator_host() {
    G1Affine g;
    fp_from_hex_host("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb", g.x);
    fp_from_hex_host("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1", g.y);
    return g;
}
inline G2Affine g2_generator_host() {
    G2Affine g;
    fp_from_hex_host("024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8", g.x.c0);
    fp_from_hex_host("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e", g.x.c1);
    fp_from_hex_host("0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801", g.y.c0);
    fp_from_hex_host("0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be", g.y.c1);
    return g;
}

// stream ids (masp_b200/synthetic.py)
enum { STREAM_VK = 10, STREAM_IC = 11, STREAM_H = 12, STREAM_L = 13, STREAM_A = 14, STREAM_B = 15 };

// Writes a bellman Parameters encoding (Appendix D) of the given shape into
// the device buffer `out` (params_synth_size bytes).
inline size_t params_synth_size(uint32_t n_ic, uint32_t h_len, uint32_t l_len, uint32_t a_len, uint32_t b_len) {
    return 3 * 96 + 3 * 192 + 4 + 96 * (size_t)n_ic + 5 * 4 + 96 * ((size_t)h_len + l_len + a_len + b_len) +
           192 * (size_t)b_len;
}
inline void params_synthesize(uint64_t seed, uint32_t n_ic, uint32_t h_len, uint32_t l_len, uint32_t a_len,
                              uint32_t b_len, uint8_t* out_dev, cudaStream_t s) {
    SynthArgs a;
    a.g1 = g1_generator_host();
    a.g2 = g2_generator_host();
    size_t pos = 0;
    auto g1seg = [&](int stream, uint64_t start, size_t n) {
        a.nthreads = n; a.key = stream_key(seed, stream); a.start = start; a.out = out_dev + pos;
        launch_synth_g1(a, s);
        pos += 96 * n;
    };
    auto g2seg = [&](int stream, uint64_t start, size_t n) {
        a.nthreads = n; a.key = stream_key(seed, stream); a.start = start; a.out = out_dev + pos;
        launch_synth_g2(a, s);
        pos += 192 * n;
    };
    auto u32seg = [&](uint32_t v) {
        uint8_t b[4] = {(uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v};
        copy_h2d(out_dev + pos, b, 4, s);
        stream_sync(s);  // b is a stack temporary
        pos += 4;
    };
    g1seg(STREAM_VK, 0, 1);  // alpha_g1
    g1seg(STREAM_VK, 1, 1);  // beta_g1
    g2seg(STREAM_VK, 1, 1);  // beta_g2
    g2seg(STREAM_VK, 2, 1);  // gamma_g2
    g1seg(STREAM_VK, 3, 1);  // delta_g1
    g2seg(STREAM_VK, 3, 1);  // delta_g2
    u32seg(n_ic);
    g1seg(STREAM_IC, 0, n_ic);
    u32seg(h_len);
    g1seg(STREAM_H, 0, h_len);
    u32seg(l_len);
    g1seg(STREAM_L, 0, l_len);
    u32seg(a_len);
    g1seg(STREAM_A, 0, a_len);
    u32seg(b_len);
    g1seg(STREAM_B, 0, b_len);
    u32seg(b_len);
    g2seg(STREAM_B, 0, b_len);
}

}  // namespace mb
