// Device-resident proving key and the batched Groth16 prover.
//
// Drop-in for the work bellman::groth16::create_proof does after synthesis
// (Appendix A), behind the reference call sites
// masp_proofs/src/sapling/prover.rs:116-117 (Spend), 201-202 (Output),
// 251-252 (Convert), with the key parsed exactly as
// Parameters::<Bls12>::read(reader, false) does (masp_proofs/src/lib.rs:336-341).
//
// Per proof, four bucket MSMs over precomputed window tables:
//   HL  = sum h_i H_i + sum aux_i L_i                       (G1, joins C)
//   A   = sum_inputs + sum_dense-aux A_i  + r*delta1 + alpha1   = proof.A
//   B1  = sum B1_i                         + beta1
//   B2  = sum B2_i + s*delta2 + beta2                            = proof.B
//   C   = s*A + r*B1 + HL
// (delta, alpha, beta ride along as extra bases of the A / B tables, so the
// only variable-base scalar multiplications left are the two in C.)
#pragma once
#include <memory>
#include <vector>

#include "msm.cuh"
#include "ntt.cuh"
#include "pairing.cuh"
#include "r1cs.cuh"

namespace mb {


// ---------------------------------------------------------------------------
// kernels: ingest, table precompute, assembly
// ---------------------------------------------------------------------------
struct DecodeArgs {
    size_t nthreads;
    const uint8_t* src;   // nthreads encodings, 96 (G1) or 192 (G2) bytes each
    void* dst;            // Affine<F>[nthreads]
    uint32_t* bad;        // set to 1 on a malformed encoding
    uint32_t reject_inf = 0;  // 1: the identity is malformed too -- bellman's read_g1 / read_g2 closures in
                              // Parameters::read and VerifyingKey::read fail with "point at infinity" even unchecked
};
MB_HD void decode_g1_body(const DecodeArgs& a, size_t tid) {
    G1Affine p;
    if (!g1_decode(a.src + 96 * tid, p)) {
        *a.bad = 1;
        p = G1Affine::inf();
    }
    if (a.reject_inf && p.is_inf()) *a.bad = 1;
    ((G1Affine*)a.dst)[tid] = p;
}
MB_HD void decode_g2_body(const DecodeArgs& a, size_t tid) {
    G2Affine p;
    if (!g2_decode(a.src + 192 * tid, p)) {
        *a.bad = 1;
        p = G2Affine::inf();
    }
    if (a.reject_inf && p.is_inf()) *a.bad = 1;
    ((G2Affine*)a.dst)[tid] = p;
}
MB_K_G1(decode_g1, DecodeArgs, decode_g1_body, 128)
MB_K_G2(decode_g2, DecodeArgs, decode_g2_body, 64)

// table[w * n + k] = affine(2^(c w) * base_k), w < nwin
template <class F>
struct TableArgs {
    size_t nthreads;  // n bases
    const Affine<F>* bases;
    Affine<F>* table;
    uint32_t c, nwin;
};
template <class F>
MB_HD void table_body(const TableArgs<F>& a, size_t tid) {
    Affine<F> p = a.bases[tid];
    a.table[tid] = p;
    MB_NOUNROLL
    for (uint32_t w = 1; w < a.nwin; ++w) {
        XYZZ<F> j = xyzz_dbl_affine_cold(p);
        MB_NOUNROLL
        for (uint32_t d = 1; d < a.c; ++d) j = xyzz_dbl_cold(j);
        p = xyzz_to_affine(j);
        a.table[(size_t)w * a.nthreads + tid] = p;
    }
}
MB_HD void table_g1_body(const TableArgs<Fp>& a, size_t tid) { table_body<Fp>(a, tid); }
MB_HD void table_g2_body(const TableArgs<Fp2>& a, size_t tid) { table_body<Fp2>(a, tid); }
MB_K_G1(build_table_g1, TableArgs<Fp>, table_g1_body, 128)
MB_K_G2(build_table_g2, TableArgs<Fp2>, table_g2_body, 64)

struct FillOneArgs {
    size_t nthreads;  // proofs
    uint32_t* pool;
    size_t pool_stride;
    size_t one_index;
};
MB_HD void fill_one_body(const FillOneArgs& a, size_t tid) {
    uint32_t* p = a.pool + (tid * a.pool_stride + a.one_index) * 8;
    p[0] = 1;
    for (int i = 1; i < 8; ++i) p[i] = 0;
}
MB_K_G1(pool_fill_one, FillOneArgs, fill_one_body, 64)

// s * A and r * B1 (two threads per proof)
struct CmulArgs {
    size_t nthreads;  // 2 * proofs
    const G1XYZZ* a_res;
    const G1XYZZ* b1_res;
    const uint32_t* pool;
    size_t pool_stride, r_index, s_index;
    G1XYZZ* out;  // [proofs][2]
};
MB_HD void cmul_body(const CmulArgs& a, size_t tid) {
    size_t proof = tid >> 1;
    const uint32_t* base = a.pool + proof * a.pool_stride * 8;
    if (tid & 1) a.out[tid] = xyzz_mul_glv(a.b1_res[proof], base + a.r_index * 8);
    else a.out[tid] = xyzz_mul_glv(a.a_res[proof], base + a.s_index * 8);
}
MB_K_G1(proof_cmul, CmulArgs, cmul_body, 32)

// affine + compressed encoding of the three proof points (three threads per proof)
struct FinishArgs {
    size_t nthreads;  // 3 * proofs
    const G1XYZZ* a_res;
    const G2XYZZ* b2_res;
    const G1XYZZ* hl_res;
    const G1XYZZ* cmul;  // [proofs][2]
    uint8_t* proofs;     // 192 bytes each
    G1Affine* aff_a;     // optional (self-check): the three points in affine form
    G2Affine* aff_b;
    G1Affine* aff_c;
};
MB_HD void finish_body(const FinishArgs& a, size_t tid) {
    size_t proof = tid / 3;
    uint32_t which = (uint32_t)(tid - proof * 3);
    uint8_t* out = a.proofs + 192 * proof;
    if (which == 0) {
        G1Affine p = xyzz_to_affine(a.a_res[proof]);
        if (a.aff_a) a.aff_a[proof] = p;
        g1_encode_compressed(p, out);
    } else if (which == 1) {
        G2Affine p = xyzz_to_affine(a.b2_res[proof]);
        if (a.aff_b) a.aff_b[proof] = p;
        g2_encode_compressed(p, out + 48);
    } else {
        G1XYZZ c = a.hl_res[proof];
        xyzz_add_cold(c, a.cmul[2 * proof]);
        xyzz_add_cold(c, a.cmul[2 * proof + 1]);
        G1Affine p = xyzz_to_affine(c);
        if (a.aff_c) a.aff_c[proof] = p;
        g1_encode_compressed(p, out + 144);
    }
}
MB_K_G2(proof_finish, FinishArgs, finish_body, 32)

// uncompressed encoding of MSM results (standalone API)
struct EncodeArgs {
    size_t nthreads;
    const void* pts;  // XYZZ<F>
    uint8_t* out;
};
MB_HD void encode_g1_body(const EncodeArgs& a, size_t tid) {
    g1_encode(xyzz_to_affine(((const G1XYZZ*)a.pts)[tid]), a.out + 96 * tid);
}
MB_HD void encode_g2_body(const EncodeArgs& a, size_t tid) {
    g2_encode(xyzz_to_affine(((const G2XYZZ*)a.pts)[tid]), a.out + 192 * tid);
}
MB_K_G1(encode_g1, EncodeArgs, encode_g1_body, 32)
MB_K_G2(encode_g2, EncodeArgs, encode_g2_body, 32)

// ---------------------------------------------------------------------------
// window-size heuristic: minimise bucket additions + reduction additions
// ---------------------------------------------------------------------------
inline uint32_t choose_window(double n_full_width, uint32_t cmax = 16) {
    uint32_t best = 4;
    double best_cost = 1e300;
    for (uint32_t c = 4; c <= cmax; ++c) {
        double cost = n_full_width * msm_nwin(c) + 4.0 * (double)(1u << (c - 1));
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}
inline uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return (uint32_t)strtoul(v, nullptr, 10);
}

// ---------------------------------------------------------------------------
// the key
// ---------------------------------------------------------------------------
struct Params {
    uint32_t n_inputs = 0, n_aux = 0, h_len = 0, a_len = 0, b_len = 0, n_b_inputs = 0;
    uint32_t log_m = 0, m = 0;
    size_t consumed = 0;  // bytes of the Parameters encoding (the MPC transcript follows in real files)
    // scalar pool layout (indices in scalars)
    size_t pool_stride = 0, idx_aux = 0, idx_inputs = 0, idx_r = 0, idx_s = 0, idx_one = 0;
    DevBuf t_hl, t_a, t_b1, t_b2;          // window tables
    DevBuf sel_hl, sel_a, sel_b1, sel_b2;  // base -> pool index
    MsmClass k_hl, k_a, k_b1, k_b2;
    NttDomain dom;
    std::vector<uint8_t> vk_bytes;  // the VerifyingKey prefix, verbatim
    DevBuf vk_g1, vk_g2, vk_ic, vk_ab;  // alpha_g1 | beta, gamma, delta (G2) | IC | Miller(-alpha, beta): the self-check
    R1csDev r1cs;                   // the circuit's matrices, when one is bound (mb200_params_bind_circuit)
    size_t table_bytes = 0;
    // the density bitmaps the key was loaded with (empty = all dense): mb200_params_bind_circuit
    // compares them with the circuit's, position by position
    std::vector<uint8_t> a_aux_density, b_input_density, b_aux_density;
};
inline bool density_equal(const std::vector<uint8_t>& loaded, const std::vector<uint8_t>& want, size_t nbits) {
    for (size_t i = 0; i < nbits; ++i) {
        bool w = (want[i >> 3] >> (i & 7)) & 1;
        bool l = loaded.empty() ? true : ((loaded[i >> 3] >> (i & 7)) & 1);
        if (w != l) return false;
    }
    return true;
}

static inline bool bm_bit(const uint8_t* bm, size_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }
static inline uint32_t be32(const uint8_t* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

template <class F>
inline void launch_table(const TableArgs<F>& a, cudaStream_t s);
template <>
inline void launch_table<Fp>(const TableArgs<Fp>& a, cudaStream_t s) { launch_build_table_g1(a, s); }
template <>
inline void launch_table<Fp2>(const TableArgs<Fp2>& a, cudaStream_t s) { launch_build_table_g2(a, s); }

template <class F>
inline MsmClass build_class(DevBuf& table, const DevBuf& bases, const DevBuf& sel, uint32_t n, uint32_t c,
                            cudaStream_t s) {
    uint32_t nwin = msm_nwin(c);
    table.alloc((size_t)n * nwin * sizeof(Affine<F>));
    TableArgs<F> ta;
    ta.nthreads = n;
    ta.bases = bases.as<Affine<F>>();
    ta.table = table.as<Affine<F>>();
    ta.c = c;
    ta.nwin = nwin;
    launch_table<F>(ta, s);
    return msm_make_class(table.p, sel.as<uint32_t>(), n, c, true);
}

// Parameters::read(reader, checked = false) + densities -> device key.
inline Params* params_load(const uint8_t* buf, size_t len, const uint8_t* a_aux_density,
                           const uint8_t* b_input_density, const uint8_t* b_aux_density, cudaStream_t s) {
    const size_t VK_FIXED = 3 * 96 + 3 * 192;
    if (!buf || len < VK_FIXED + 4) fail(MB200_EPARSE, "parameters truncated (%s%ld bytes)", "", (long)len);
    size_t pos = VK_FIXED;
    auto need = [&](size_t n) {
        if (n > len - pos) fail(MB200_EPARSE, "parameters truncated at byte %s%ld", "", (long)pos);
    };
    auto vec = [&](size_t elem, uint32_t& count, size_t& off) {
        need(4);
        count = be32(buf + pos);
        pos += 4;
        if ((size_t)count > (len - pos) / elem) fail(MB200_EPARSE, "parameters truncated at byte %s%ld", "", (long)pos);
        off = pos;
        pos += (size_t)count * elem;
    };
    uint32_t n_ic, n_h, n_l, n_a, n_b1, n_b2;
    size_t o_ic, o_h, o_l, o_a, o_b1, o_b2;
    vec(96, n_ic, o_ic);
    size_t vk_end = pos;
    vec(96, n_h, o_h);
    vec(96, n_l, o_l);
    vec(96, n_a, o_a);
    vec(96, n_b1, o_b1);
    vec(192, n_b2, o_b2);

    std::unique_ptr<Params> P(new Params());
    P->consumed = pos;
    P->vk_bytes.assign(buf, buf + vk_end);
    P->n_inputs = n_ic;
    P->n_aux = n_l;
    P->h_len = n_h;
    P->a_len = n_a;
    P->b_len = n_b1;
    if (n_ic == 0) fail(MB200_EPARSE, "verifying key has no IC elements%s", "");
    if (n_b2 != n_b1) fail(MB200_EPARSE, "b_g1 / b_g2 lengths differ%s (%ld)", "", (long)n_b2);
    uint32_t m = n_h + 1;
    if (m < 2 || (m & (m - 1))) fail(MB200_EPARSE, "h query length + 1 is not a power of two%s (%ld)", "", (long)n_h);
    P->m = m;
    while ((1u << P->log_m) < m) P->log_m++;

    // density walk (Appendix A "MSMs"): k-th dense variable <-> k-th base
    std::vector<uint32_t> a_sel, b_in_sel, b_sel;
    for (uint32_t i = 0; i < n_l; ++i) {
        if (!a_aux_density || bm_bit(a_aux_density, i)) a_sel.push_back(i);
        if (!b_aux_density || bm_bit(b_aux_density, i)) b_sel.push_back(i);
    }
    for (uint32_t i = 0; i < n_ic; ++i)
        if (!b_input_density || bm_bit(b_input_density, i)) b_in_sel.push_back(i);
    if (n_a != n_ic + a_sel.size())
        fail(MB200_EINVAL, "a query length does not match a_aux_density%s (%ld bases)", "", (long)n_a);
    if (n_b1 != b_in_sel.size() + b_sel.size())
        fail(MB200_EINVAL, "b query length does not match the b densities%s (%ld bases)", "", (long)n_b1);
    P->n_b_inputs = (uint32_t)b_in_sel.size();
    if (a_aux_density) P->a_aux_density.assign(a_aux_density, a_aux_density + (n_l + 7) / 8);
    if (b_input_density) P->b_input_density.assign(b_input_density, b_input_density + (n_ic + 7) / 8);
    if (b_aux_density) P->b_aux_density.assign(b_aux_density, b_aux_density + (n_l + 7) / 8);

    P->idx_aux = m;
    P->idx_inputs = (size_t)m + n_l;
    P->idx_r = P->idx_inputs + n_ic;
    P->idx_s = P->idx_r + 1;
    P->idx_one = P->idx_s + 1;
    P->pool_stride = P->idx_one + 1;

    // base -> pool index maps
    std::vector<uint32_t> s_hl, s_a, s_b1, s_b2;
    for (uint32_t i = 0; i < n_h; ++i) s_hl.push_back(i);
    for (uint32_t i = 0; i < n_l; ++i) s_hl.push_back((uint32_t)P->idx_aux + i);
    for (uint32_t i = 0; i < n_ic; ++i) s_a.push_back((uint32_t)P->idx_inputs + i);
    for (uint32_t v : a_sel) s_a.push_back((uint32_t)P->idx_aux + v);
    s_a.push_back((uint32_t)P->idx_r);    // delta_g1
    s_a.push_back((uint32_t)P->idx_one);  // alpha_g1
    for (uint32_t v : b_in_sel) s_b1.push_back((uint32_t)P->idx_inputs + v);
    for (uint32_t v : b_sel) s_b1.push_back((uint32_t)P->idx_aux + v);
    s_b2 = s_b1;
    s_b1.push_back((uint32_t)P->idx_one);  // beta_g1
    s_b2.push_back((uint32_t)P->idx_s);    // delta_g2
    s_b2.push_back((uint32_t)P->idx_one);  // beta_g2
    auto upload = [&](DevBuf& d, const std::vector<uint32_t>& v) {
        d.alloc(v.size() * 4);
        copy_h2d(d.p, v.data(), v.size() * 4, s);
    };
    upload(P->sel_hl, s_hl);
    upload(P->sel_a, s_a);
    upload(P->sel_b1, s_b1);
    upload(P->sel_b2, s_b2);

    // raw bytes -> Montgomery affine bases
    DevBuf raw(pos), bad(4);
    copy_h2d(raw.p, buf, pos, s);
    dev_memset(bad.p, 0, 4, s);
    const uint8_t* rb = raw.as<uint8_t>();
    DevBuf b_hl((size_t)(n_h + n_l) * sizeof(G1Affine)), b_a((size_t)(n_a + 2) * sizeof(G1Affine)),
        b_b1((size_t)(n_b1 + 1) * sizeof(G1Affine)), b_b2((size_t)(n_b2 + 2) * sizeof(G2Affine));
    auto dec1 = [&](size_t off, uint32_t n, G1Affine* dst) {
        DecodeArgs a{n, rb + off, dst, bad.as<uint32_t>(), 1};
        launch_decode_g1(a, s);
    };
    auto dec2 = [&](size_t off, uint32_t n, G2Affine* dst) {
        DecodeArgs a{n, rb + off, dst, bad.as<uint32_t>(), 1};
        launch_decode_g2(a, s);
    };
    // VerifyingKey: alpha_g1 0, beta_g1 96, beta_g2 192, gamma_g2 384, delta_g1 576, delta_g2 672
    dec1(o_h, n_h, b_hl.as<G1Affine>());
    dec1(o_l, n_l, b_hl.as<G1Affine>() + n_h);
    dec1(o_a, n_a, b_a.as<G1Affine>());
    dec1(576, 1, b_a.as<G1Affine>() + n_a);
    dec1(0, 1, b_a.as<G1Affine>() + n_a + 1);
    dec1(o_b1, n_b1, b_b1.as<G1Affine>());
    dec1(96, 1, b_b1.as<G1Affine>() + n_b1);
    dec2(o_b2, n_b2, b_b2.as<G2Affine>());
    dec2(672, 1, b_b2.as<G2Affine>() + n_b2);
    dec2(192, 1, b_b2.as<G2Affine>() + n_b2 + 1);
    // the verifying key, kept for the post-proof self-check (pairing.cuh)
    P->vk_g1.alloc(sizeof(G1Affine));
    P->vk_g2.alloc(3 * sizeof(G2Affine));
    P->vk_ic.alloc((size_t)n_ic * sizeof(G1Affine));
    P->vk_ab.alloc(sizeof(Fp12));
    dec1(0, 1, P->vk_g1.as<G1Affine>());
    dec2(192, 1, P->vk_g2.as<G2Affine>());
    dec2(384, 1, P->vk_g2.as<G2Affine>() + 1);
    dec2(672, 1, P->vk_g2.as<G2Affine>() + 2);
    dec1(o_ic, n_ic, P->vk_ic.as<G1Affine>());
    PairPrepArgs pp{1, P->vk_g1.as<G1Affine>(), P->vk_g2.as<G2Affine>(), P->vk_ab.as<Fp12>()};
    launch_pair_prep(pp, s);

    // window sizes: full-width share guessed from the reference circuits
    // (SURVEY §8 scalar make-up: ~1/3 of L, ~1/5 of A and B are full width)
    uint32_t c_hl = env_u32("MB200_C_HL", choose_window(n_h + 0.33 * n_l));
    uint32_t c_a = env_u32("MB200_C_A", choose_window(0.25 * n_a + 16));
    uint32_t c_b1 = env_u32("MB200_C_B1", choose_window(0.25 * n_b1 + 16));
    uint32_t c_b2 = env_u32("MB200_C_B2", choose_window(0.25 * n_b2 + 16));
    P->k_hl = build_class<Fp>(P->t_hl, b_hl, P->sel_hl, n_h + n_l, c_hl, s);
    P->k_a = build_class<Fp>(P->t_a, b_a, P->sel_a, n_a + 2, c_a, s);
    P->k_b1 = build_class<Fp>(P->t_b1, b_b1, P->sel_b1, n_b1 + 1, c_b1, s);
    P->k_b2 = build_class<Fp2>(P->t_b2, b_b2, P->sel_b2, n_b2 + 2, c_b2, s);
    P->table_bytes = P->t_hl.bytes + P->t_a.bytes + P->t_b1.bytes + P->t_b2.bytes;
    P->dom.build(P->log_m, s);

    uint32_t bad_h = 0;
    copy_d2h(&bad_h, bad.p, 4, s);
    stream_sync(s);
    if (bad_h) fail(MB200_EPARSE, "malformed point encoding in parameters%s", "");
    return P.release();
}

// ---------------------------------------------------------------------------
// prove
// ---------------------------------------------------------------------------
struct ProveCtx {  // one in-flight chunk: its streams and scratch
    cudaStream_t stream = 0;      // copies, NTT, H+L MSM, assembly
    cudaStream_t side[3] = {0, 0, 0};  // A, B1, B2 MSMs: they need only the staged witness, not the NTT
    DevBuf abc, w0, w1, w2, w3, pool, flag, res_hl, res_a, res_b1, res_b2, cmul, proofs;
    MsmScratch msm, msm_side[3];
    bool have_stream = false;
#ifndef MB200_EMU
    cudaEvent_t ev_inputs = nullptr, ev_side[3] = {nullptr, nullptr, nullptr}, ev_tail = nullptr;
#endif
};

struct ProveInputs {  // host or device pointers, selected by `on_device`
    const uint8_t *a, *b, *c, *inputs, *aux, *r, *s;
    bool on_device;
};

inline void copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                      bool src_on_device, cudaStream_t s) {
#ifdef MB200_EMU
    (void)src_on_device;
    (void)s;
    for (size_t i = 0; i < height; ++i) memcpy((char*)dst + i * dpitch, (const char*)src + i * spitch, width);
#else
    MB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height,
                              src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
#endif
}

// Enqueue one chunk of proofs [first, first + count) on ctx.stream.  proofs_out:
// host memory, 192 bytes per proof.  No host synchronisation.
// Where a batch's self-check lives: per-batch device buffers (so the chunk context is free again
// as soon as the proofs are encoded) and the stream the check runs on.
struct VerifySink {
    G1Affine* a = nullptr;   // [n_proofs] affine proof points
    G2Affine* b = nullptr;
    G1Affine* c = nullptr;
    uint32_t* inputs = nullptr;  // [n_proofs][n_inputs] scalars
    uint32_t* ok_dev = nullptr;  // [n_proofs]
    uint32_t* ok_host = nullptr; // pinned
    cudaStream_t stream = 0;     // latency-bound verifier kernels run here, next to the following chunks
    bool on() const { return a != nullptr; }
};

inline void prove_chunk(const Params& P, ProveCtx& x, const ProveInputs& in, size_t first, uint32_t count,
                        size_t rows, uint8_t* proofs_out, const VerifySink& vs = VerifySink()) {
    cudaStream_t s = x.stream;
    const uint32_t m = P.m;
    size_t polys = (size_t)count * 3;
    x.abc.ensure(polys * rows * 32);
    x.w0.ensure(polys * m * 32);
    x.w1.ensure(polys * m * 32);
    x.w2.ensure(polys * m * 32);
    x.w3.ensure(polys * m * 32);
    x.pool.ensure((size_t)count * P.pool_stride * 32);
    x.res_hl.ensure(count * sizeof(G1XYZZ));
    x.res_a.ensure(count * sizeof(G1XYZZ));
    x.res_b1.ensure(count * sizeof(G1XYZZ));
    x.res_b2.ensure(count * sizeof(G2XYZZ));
    x.cmul.ensure((size_t)count * 2 * sizeof(G1XYZZ));
    x.proofs.ensure((size_t)count * 192);

    const size_t rb = rows * 32;
    uint8_t* abc = x.abc.as<uint8_t>();
    if (in.a) {
        copy_rows(abc, 3 * rb, in.a + first * rb, rb, rb, count, in.on_device, s);
        copy_rows(abc + rb, 3 * rb, in.b + first * rb, rb, rb, count, in.on_device, s);
        copy_rows(abc + 2 * rb, 3 * rb, in.c + first * rb, rb, rb, count, in.on_device, s);
    }
    uint8_t* pool = x.pool.as<uint8_t>();
    const size_t pitch = P.pool_stride * 32;
    copy_rows(pool + P.idx_aux * 32, pitch, in.aux + first * P.n_aux * 32, (size_t)P.n_aux * 32, (size_t)P.n_aux * 32,
              count, in.on_device, s);
    copy_rows(pool + P.idx_inputs * 32, pitch, in.inputs + first * P.n_inputs * 32, (size_t)P.n_inputs * 32,
              (size_t)P.n_inputs * 32, count, in.on_device, s);
    copy_rows(pool + P.idx_r * 32, pitch, in.r + first * 32, 32, 32, count, in.on_device, s);
    copy_rows(pool + P.idx_s * 32, pitch, in.s + first * 32, 32, 32, count, in.on_device, s);
    FillOneArgs fo{count, x.pool.as<uint32_t>(), P.pool_stride, P.idx_one};
    launch_pool_fill_one(fo, s);
    if (!in.a) {
        // witness-only call: the row evaluations are a sparse product over the staged witness
        const R1csDev& R = P.r1cs;
        R1csArgs ra;
        ra.nthreads = (size_t)count * rows;
        for (int k = 0; k < 3; ++k) {
            ra.rowptr[k] = R.rowptr[k].as<uint32_t>();
            ra.col[k] = R.col[k].as<uint32_t>();
            ra.cidx[k] = R.cidx[k].as<uint32_t>();
        }
        ra.dict = R.dict.as<Fr>();
        ra.ncons = R.ncons;
        ra.rows = (uint32_t)rows;
        ra.pool = x.pool.as<Fr>();
        ra.pool_stride = P.pool_stride;
        ra.idx_inputs = P.idx_inputs;
        ra.abc = x.abc.as<Fr>();
        ra.order = R.order.as<uint32_t>();
        launch_r1cs_eval(ra, s);
    }

    const uint32_t* pl = x.pool.as<uint32_t>();
    // fork: the A / B1 / B2 queries read aux, inputs, r, s only, so they run on
    // side streams next to the NTT pipeline and the H+L query; their
    // latency-bound reduction tails overlap the other streams' heavy kernels.
#ifndef MB200_EMU
    MB_CUDA(cudaEventRecord(x.ev_inputs, s));
    for (int i = 0; i < 3; ++i) MB_CUDA(cudaStreamWaitEvent(x.side[i], x.ev_inputs, 0));
#endif
    msm_run<Fp>(P.k_a, count, pl, P.pool_stride, x.res_a.as<G1XYZZ>(), x.msm_side[0], x.side[0]);
    msm_run<Fp>(P.k_b1, count, pl, P.pool_stride, x.res_b1.as<G1XYZZ>(), x.msm_side[1], x.side[1]);
    msm_run<Fp2>(P.k_b2, count, pl, P.pool_stride, x.res_b2.as<G2XYZZ>(), x.msm_side[2], x.side[2]);

    h_pipeline(P.dom, count, (uint32_t)rows, x.abc.as<Fr>(), rows, x.pool.as<Fr>(), P.pool_stride, x.w0.as<Fr>(),
               x.w1.as<Fr>(), x.w2.as<Fr>(), x.w3.as<Fr>(), s);
    // (Measured and rejected, profiles/r02_ab_hl_base_ranges.jsonl: the H+L query in 4 / 8 / 16 base ranges
    // whose table strips fit the L2.  Every launch of the accumulate kernel walks each bucket's whole entry
    // list, i.e. the whole 356 MB table, once per wave of resident blocks -- ~70 waves per 64-proof chunk,
    // 24.4 GB of DRAM reads, 13x the algorithmic bytes.  Ranges cut that to 14.9 GB and the L2 does hold
    // the strips, but the kernel is bound by the multiplier, not by those reads: 508 / 499 / 492 proofs/s
    // against 530 in one launch, because shorter tasks pay the accumulator load / store and the launch
    // tail more often.)
    msm_run<Fp>(P.k_hl, count, pl, P.pool_stride, x.res_hl.as<G1XYZZ>(), x.msm, s);
    // join: the assembly waits for the three side queries
#ifndef MB200_EMU
    for (int i = 0; i < 3; ++i) {
        MB_CUDA(cudaEventRecord(x.ev_side[i], x.side[i]));
        MB_CUDA(cudaStreamWaitEvent(s, x.ev_side[i], 0));
    }
#endif

    CmulArgs ca{(size_t)count * 2, x.res_a.as<G1XYZZ>(), x.res_b1.as<G1XYZZ>(), pl, P.pool_stride, P.idx_r, P.idx_s,
                x.cmul.as<G1XYZZ>()};
    launch_proof_cmul(ca, s);
    FinishArgs fa{(size_t)count * 3, x.res_a.as<G1XYZZ>(), x.res_b2.as<G2XYZZ>(), x.res_hl.as<G1XYZZ>(),
                  x.cmul.as<G1XYZZ>(), x.proofs.as<uint8_t>(), nullptr, nullptr, nullptr};
    if (vs.on()) {
        fa.aff_a = vs.a + first;
        fa.aff_b = vs.b + first;
        fa.aff_c = vs.c + first;
    }
    launch_proof_finish(fa, s);
    copy_d2h(proofs_out + first * 192, x.proofs.p, (size_t)count * 192, s);
    if (vs.on()) {
        // verify_proof(vk, proof, inputs) as at masp_proofs/src/sapling/prover.rs:148, :266.  The
        // public inputs are copied out of the chunk's scalar pool so that nothing the check reads
        // belongs to the chunk context any more.
        uint32_t* vin = vs.inputs + first * (size_t)P.n_inputs * 8;
        copy_rows(vin, (size_t)P.n_inputs * 32, x.pool.as<uint8_t>() + P.idx_inputs * 32, P.pool_stride * 32,
                  (size_t)P.n_inputs * 32, count, true, s);
        cudaStream_t v = vs.stream;
#ifndef MB200_EMU
        MB_CUDA(cudaEventRecord(x.ev_tail, s));
        MB_CUDA(cudaStreamWaitEvent(v, x.ev_tail, 0));
#endif
        VerifyArgs va;
        va.nthreads = count;
        va.pa = fa.aff_a;
        va.pb = fa.aff_b;
        va.pc = fa.aff_c;
        va.inputs = vin;
        va.input_stride = P.n_inputs;
        va.n_inputs = P.n_inputs;
        va.vk = {P.vk_ic.as<G1Affine>(), P.vk_g2.as<G2Affine>() + 1, P.vk_g2.as<G2Affine>() + 2, P.vk_ab.as<Fp12>()};
        va.ok = vs.ok_dev + first;
        launch_verify_proofs(va, v);
        copy_d2h(vs.ok_host + first, va.ok, (size_t)count * 4, v);
    }
}

}  // namespace mb
