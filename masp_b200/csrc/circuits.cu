// Host side of the proving path: circuit recording and witness generation for
// the MASP Spend / Output / Convert circuits (NEXT-1), behind
// the mb200_circuit_* entry points of include/masp_b200.h.  This unit holds no
// device code; the row evaluations run on the GPU (r1cs.cuh).
#include <atomic>
#include <memory>
#include <thread>

#include "host/circuit_obj.hpp"
#include "host/gadgets.hpp"
#include "rt.cuh"

using namespace mbh;

namespace {

// --- BLAKE2s-256, unkeyed, no personalization (for TestConstraintSystem::hash) ---
struct Blake2s {
    uint32_t h[8];
    uint8_t buf[64];
    size_t buflen = 0;
    uint64_t t = 0;
    Blake2s() {
        for (int i = 0; i < 8; ++i) h[i] = blake2s_gadget::IV[i];
        h[0] ^= 0x01010000u ^ 32u;
    }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void compress(const uint8_t* block, bool last) {
        uint32_t m[16], v[16];
        for (int i = 0; i < 16; ++i)
            m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) |
                   ((uint32_t)block[4 * i + 3] << 24);
        for (int i = 0; i < 8; ++i) {
            v[i] = h[i];
            v[i + 8] = blake2s_gadget::IV[i];
        }
        v[12] ^= (uint32_t)t;
        v[13] ^= (uint32_t)(t >> 32);
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
            v[a] = v[a] + v[b] + x;
            v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 12);
            v[a] = v[a] + v[b] + y;
            v[d] = rotr(v[d] ^ v[a], 8);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 7);
        };
        for (int r = 0; r < 10; ++r) {
            const uint8_t* s = blake2s_gadget::SIGMA[r];
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
    void update(const uint8_t* p, size_t n) {
        while (n) {
            if (buflen == 64) {
                t += 64;
                compress(buf, false);
                buflen = 0;
            }
            size_t k = std::min(n, 64 - buflen);
            memcpy(buf + buflen, p, k);
            buflen += k;
            p += k;
            n -= k;
        }
    }
    void u64be(uint64_t x) {
        uint8_t b[8];
        for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(x >> (56 - 8 * i));
        update(b, 8);
    }
    void finish(uint8_t out[32]) {
        t += buflen;
        memset(buf + buflen, 0, 64 - buflen);
        compress(buf, true);
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 4; ++j) out[4 * i + j] = (uint8_t)(h[i] >> (8 * j));
    }
};

std::string structural_hash(const CS& cs) {
    Blake2s h;
    h.u64be(cs.inputs.size());
    h.u64be(cs.aux.size());
    h.u64be(cs.n_constraints);
    const Matrix* ms[3] = {&cs.A, &cs.B, &cs.C};
    for (size_t row = 0; row < cs.n_constraints; ++row) {
        for (int k = 0; k < 3; ++k) {
            const Matrix& m = *ms[k];
            uint32_t e0 = m.rowptr[row], e1 = m.rowptr[row + 1];
            h.u64be(e1 - e0);
            for (uint32_t e = e0; e < e1; ++e) {
                uint32_t id = m.col[e];
                uint8_t kind = (id & Var::AUX) ? 'A' : 'I';
                h.update(&kind, 1);
                h.u64be(id & ~Var::AUX);
                uint8_t le[32], be[32];
                m.coef[e].to_bytes(le);
                for (int i = 0; i < 32; ++i) be[i] = le[31 - i];
                h.update(be, 32);
            }
        }
    }
    uint8_t d[32];
    h.finish(d);
    static const char* hx = "0123456789abcdef";
    std::string s;
    for (int i = 0; i < 32; ++i) {
        s.push_back(hx[d[i] >> 4]);
        s.push_back(hx[d[i] & 15]);
    }
    return s;
}

size_t witness_size(int kind, uint32_t depth) {
    switch (kind) {
        case MB200_CIRCUIT_SPEND: return 32 * (12 + 2 * (size_t)depth);
        case MB200_CIRCUIT_OUTPUT: return 32 * 11;
        case MB200_CIRCUIT_CONVERT: return 32 * (5 + 2 * (size_t)depth);
    }
    return 0;
}

// one witness on the scalar field; false: a field is out of range
bool run_circuit(CS& cs, int kind, uint32_t depth, const uint8_t* w) {
    bool bad = false;
    run_circuit_lanes<ScalarPolicy>(cs, kind, depth, &w, &bad);
    return !bad;
}

}  // namespace

extern "C" {

int mb200_circuit_new(int kind, uint32_t merkle_depth, mb200_circuit** out) {
    if (!out || kind < 0 || kind > 2 || merkle_depth > 62) return MB200_EINVAL;
    if (kind == MB200_CIRCUIT_OUTPUT) merkle_depth = 0;
    try {
        std::unique_ptr<mb200_circuit> c(new mb200_circuit());
        c->kind = kind;
        c->depth = merkle_depth;
        c->witness_bytes = witness_size(kind, merkle_depth);
        // shape pass on an all-zero witness: the structure does not depend on values
        std::vector<uint8_t> w(c->witness_bytes, 0);
        CS cs;
        recording() = true;
        run_circuit(cs, kind, merkle_depth, w.data());
        recording() = false;
        c->n_inputs = (uint32_t)cs.inputs.size();
        c->n_aux = (uint32_t)cs.aux.size();
        c->n_constraints = (uint32_t)cs.n_constraints;
        c->hash_hex = structural_hash(cs);
        c->a_aux_density.assign((c->n_aux + 7) / 8, 0);
        c->b_aux_density.assign((c->n_aux + 7) / 8, 0);
        c->b_input_density.assign((c->n_inputs + 7) / 8, 0);
        auto mark = [](std::vector<uint8_t>& bm, uint32_t i, uint32_t& ones) {
            if (!((bm[i >> 3] >> (i & 7)) & 1)) {
                bm[i >> 3] |= (uint8_t)(1u << (i & 7));
                ++ones;
            }
        };
        for (uint32_t id : cs.A.col)
            if (id & Var::AUX) mark(c->a_aux_density, id & ~Var::AUX, c->a_aux_ones);
        for (uint32_t id : cs.B.col) {
            if (id & Var::AUX) mark(c->b_aux_density, id & ~Var::AUX, c->b_aux_ones);
            else mark(c->b_input_density, id, c->b_input_ones);
        }
        c->A = std::move(cs.A);
        c->B = std::move(cs.B);
        c->C = std::move(cs.C);
        *out = c.release();
    } catch (const std::bad_alloc&) {
        recording() = false;
        return MB200_ENOMEM;
    }
    return MB200_OK;
}

void mb200_circuit_free(mb200_circuit* c) { delete c; }

int mb200_circuit_info(const mb200_circuit* c, uint64_t info[10]) {
    if (!c || !info) return MB200_EINVAL;
    info[0] = c->n_inputs;
    info[1] = c->n_aux;
    info[2] = c->n_constraints;
    info[3] = c->A.col.size();
    info[4] = c->B.col.size();
    info[5] = c->C.col.size();
    info[6] = c->witness_bytes;
    info[7] = c->a_aux_ones;
    info[8] = c->b_input_ones;
    info[9] = c->b_aux_ones;
    return MB200_OK;
}

int mb200_circuit_hash(const mb200_circuit* c, char out_hex[65]) {
    if (!c || !out_hex) return MB200_EINVAL;
    memcpy(out_hex, c->hash_hex.c_str(), 65);
    return MB200_OK;
}

int mb200_circuit_densities(const mb200_circuit* c, uint8_t* a_aux, uint8_t* b_input, uint8_t* b_aux) {
    if (!c || !a_aux || !b_input || !b_aux) return MB200_EINVAL;
    memcpy(a_aux, c->a_aux_density.data(), c->a_aux_density.size());
    memcpy(b_input, c->b_input_density.data(), c->b_input_density.size());
    memcpy(b_aux, c->b_aux_density.data(), c->b_aux_density.size());
    return MB200_OK;
}

int mb200_circuit_matrix(const mb200_circuit* c, int which, uint32_t* rowptr, uint32_t* col, uint8_t* coef) {
    if (!c || which < 0 || which > 2) return MB200_EINVAL;
    const Matrix& m = which == 0 ? c->A : which == 1 ? c->B : c->C;
    if (rowptr) memcpy(rowptr, m.rowptr.data(), m.rowptr.size() * 4);
    if (col) memcpy(col, m.col.data(), m.col.size() * 4);
    if (coef)
        for (size_t i = 0; i < m.coef.size(); ++i) m.coef[i].to_bytes(coef + 32 * i);
    return MB200_OK;
}

int mb200_pedersen_hash(const uint8_t* bits, size_t n_bits, uint8_t u_out[32], uint8_t v_out[32]) {
    if (!bits || !u_out || !v_out || n_bits < 6 || n_bits > 6 + 3 * 63 * 6) return MB200_EINVAL;
    try {
        CS cs;
        bool pers[6];
        for (int i = 0; i < 6; ++i) pers[i] = bits[i] != 0;
        GS::Bits in;
        for (size_t i = 6; i < n_bits; ++i) in.push_back(GS::Boolean::from_bit(GS::AllocatedBit::alloc(cs, bits[i] != 0)));
        GS::EdwardsPoint h = GS::pedersen_hash(cs, pers, in);
        if (cs.failed) return MB200_ESYNTH;
        h.u.value.to_bytes(u_out);
        h.v.value.to_bytes(v_out);
    } catch (const std::bad_alloc&) {
        return MB200_ENOMEM;
    }
    return MB200_OK;
}

int mb200_circuit_root(const mb200_circuit* c, const uint8_t* witness, uint8_t root_out[32]) {
    if (!c || !witness || !root_out || c->kind == MB200_CIRCUIT_OUTPUT) return MB200_EINVAL;
    try {
        CS cs;
        if (!run_circuit(cs, c->kind, c->depth, witness)) return MB200_ESCALAR;
        if (cs.failed) return MB200_ESYNTH;
        cs.root.to_bytes(root_out);
    } catch (const std::bad_alloc&) {
        return MB200_ENOMEM;
    }
    return MB200_OK;
}

// csrc/circuits_simd.cpp: eight witnesses per call on AVX-512 IFMA lanes
int mbh_simd_synthesize8(int kind, uint32_t depth, const uint8_t* const witnesses[8], uint8_t* const inputs_out[8],
                         uint8_t* const aux_out[8], uint32_t expect_inputs, uint32_t expect_aux, uint8_t status[8]);
void mbh_simd_warm(void);
int mbh_simd_compiled(void);

// 1: witnesses are generated eight at a time on AVX-512 IFMA lanes; 0: one at a time (CPU without
// IFMA, or MB200_WITNESS_SCALAR=1 for A/B measurements)
int mb200_circuit_simd(void) {
    static const int on = [] {
        const char* e = getenv("MB200_WITNESS_SCALAR");
        if (e && *e && *e != '0') return 0;
        if (!mbh_simd_compiled()) return 0;
#if defined(__x86_64__) && (defined(__GNUC__) || defined(__clang__))
        __builtin_cpu_init();
        return (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512ifma") &&
                __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq") &&
                __builtin_cpu_supports("avx512bw")) ? 1 : 0;
#else
        return 0;
#endif
    }();
    return on;
}

int mb200_circuit_synthesize(const mb200_circuit* c, size_t n, const uint8_t* witnesses, uint8_t* inputs_out,
                             uint8_t* aux_out, int n_threads) {
    if (!c || (n && (!witnesses || !inputs_out || !aux_out))) return MB200_EINVAL;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    const bool simd = mb200_circuit_simd() != 0 && n >= 2;
    // work items: groups of eight witnesses (SIMD; the last group is padded by repeating its last
    // witness into scratch outputs) or single witnesses (scalar)
    const size_t group = simd ? 8 : 1;
    const size_t n_items = (n + group - 1) / group;
    if ((size_t)n_threads > n_items) n_threads = (int)n_items;
    (void)JJ();  // build the window tables before the workers start
    (void)K();
    (void)GS::T();
    if (simd) mbh_simd_warm();
    std::atomic<size_t> next(0);
    std::atomic<int> status(MB200_OK);
    auto work = [&]() {
        try {
            std::vector<uint8_t> scratch_in, scratch_aux;  // outputs of padding lanes
            for (;;) {
                size_t item = next.fetch_add(1);
                if (item >= n_items) break;
                if (simd) {
                    const size_t first = item * 8, live = std::min<size_t>(8, n - first);
                    const uint8_t* w[8];
                    uint8_t *io[8], *ao[8];
                    if (live < 8) {
                        scratch_in.resize((size_t)c->n_inputs * 32);
                        scratch_aux.resize((size_t)c->n_aux * 32);
                    }
                    for (size_t k = 0; k < 8; ++k) {
                        size_t i = first + std::min(k, live - 1);
                        w[k] = witnesses + i * c->witness_bytes;
                        io[k] = k < live ? inputs_out + i * (size_t)c->n_inputs * 32 : scratch_in.data();
                        ao[k] = k < live ? aux_out + i * (size_t)c->n_aux * 32 : scratch_aux.data();
                    }
                    uint8_t st[8];
                    if (mbh_simd_synthesize8(c->kind, c->depth, w, io, ao, c->n_inputs, c->n_aux, st) != 0) {
                        status = MB200_ENOMEM;
                        continue;
                    }
                    for (size_t k = 0; k < live; ++k) {
                        if (st[k] == 1) status = MB200_ESCALAR;
                        else if (st[k]) status = MB200_ESYNTH;
                    }
                    continue;
                }
                const size_t i = item;
                CS cs;
                cs.aux.reserve(c->n_aux);
                if (!run_circuit(cs, c->kind, c->depth, witnesses + i * c->witness_bytes)) {
                    status = MB200_ESCALAR;
                    continue;
                }
                if (cs.failed || cs.inputs.size() != c->n_inputs || cs.aux.size() != c->n_aux) {
                    status = MB200_ESYNTH;
                    continue;
                }
                uint8_t* io = inputs_out + i * (size_t)c->n_inputs * 32;
                for (uint32_t k = 0; k < c->n_inputs; ++k) cs.inputs[k].to_bytes(io + 32 * k);
                uint8_t* ao = aux_out + i * (size_t)c->n_aux * 32;
                for (uint32_t k = 0; k < c->n_aux; ++k) cs.aux[k].to_bytes(ao + 32 * k);
            }
        } catch (const std::bad_alloc&) {
            status = MB200_ENOMEM;
        } catch (...) {
            status = MB200_ESYNTH;
        }
    };
    // nothing may unwind across the C boundary: a worker that cannot be spawned (std::system_error)
    // or a failed vector growth just means fewer workers -- the work queue is shared
    std::vector<std::thread> pool;
    try {
        pool.reserve((size_t)n_threads);
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    } catch (...) {
    }
    work();
    for (auto& t : pool) t.join();
    return status.load();
}

}  // extern "C"
