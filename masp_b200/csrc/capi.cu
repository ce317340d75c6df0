// C ABI of libmasp_b200 (include/masp_b200.h).  Host-side orchestration in
// C++ above hand-written sm_100a kernels; no CPU compute path exists here:
// every entry point needs mb200_init to have found a CUDA device.
#include <algorithm>
#include <array>
#include <map>
#include <mutex>

#include "host/circuit_obj.hpp"
#include "prover.cuh"
#include "synth.cuh"
#include "misc.cuh"

namespace mb {

unsigned long long g_launches = 0;
MsmProfile g_msm_profile;

// ---------------------------------------------------------------------------
// process state
// ---------------------------------------------------------------------------
struct NttCache {
    NttDomain dom;
    DevBuf gpow;  // g^i, Montgomery (standalone coset_fft)
};
struct State {
    bool inited = false;
    int device = 0;
    cudaStream_t main = 0;
    std::vector<ProveCtx> ctxs;
    uint32_t chunk = 64;
    int verify = 0;  // option "verify": 1 = check every proof before returning it (failure = MB200_EVERIFY),
                     // 2 = check and only count the failures (counter "verify_failed"; for measuring the cost)
    unsigned long long verify_failed = 0, verified = 0;
    cudaStream_t vstream[4] = {0, 0, 0, 0};  // the self-check's streams (round robin over chunks)
    size_t next_v = 0;
    std::map<unsigned, NttCache*> ntt;
    MsmScratch msm;
    std::mutex mu;
    double last_batch_ms = 0;  // device time of the last prove call (CUDA events on the chunk streams)
};
static State g;

static void require_init() {
    if (!g.inited) fail(MB200_ESTATE, "mb200_init has not been called%s", "");
}
static void set_ctx_count(size_t n) {
#ifndef MB200_EMU
    for (auto& c : g.ctxs)
        if (c.have_stream) {
            cudaStreamSynchronize(c.stream);
            cudaStreamDestroy(c.stream);
            cudaEventDestroy(c.ev_inputs);
            for (int i = 0; i < 3; ++i) {
                cudaStreamSynchronize(c.side[i]);
                cudaStreamDestroy(c.side[i]);
                cudaEventDestroy(c.ev_side[i]);
            }
            for (int i = 0; i < 4; ++i) {
                cudaStreamSynchronize(c.tail[i]);
                cudaStreamDestroy(c.tail[i]);
                cudaEventDestroy(c.ev_acc[i]);
            }
            cudaEventDestroy(c.ev_tail);
        }
#endif
    g.ctxs.clear();
    g.ctxs.resize(n);
#ifndef MB200_EMU
    // Throughput kernels (copies, NTT, digit sort, bucket accumulation) run on low-priority
    // streams; the latency-bound tails (bucket reduction, s*A and r*B1, encoding, self-check) on
    // high-priority ones, so that their few blocks are placed as soon as any SM has room instead
    // of queueing behind the thousands of pending accumulation blocks of the next chunk.
    int pr_least = 0, pr_greatest = 0;
    MB_CUDA(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
    // measured on B200 (profiles/r01_tail_streams_ab.jsonl): no gain over the single-stream tails,
    // the accumulation blocks hold the register file either way -- kept as an opt-in experiment
    const bool split = getenv("MB200_TAIL_STREAMS") != nullptr;
    for (auto& c : g.ctxs) {
        MB_CUDA(cudaStreamCreateWithPriority(&c.stream, cudaStreamNonBlocking, pr_least));
        MB_CUDA(cudaEventCreateWithFlags(&c.ev_inputs, cudaEventDisableTiming));
        for (int i = 0; i < 3; ++i) {
            MB_CUDA(cudaStreamCreateWithPriority(&c.side[i], cudaStreamNonBlocking, pr_least));
            MB_CUDA(cudaEventCreateWithFlags(&c.ev_side[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 4; ++i) {
            MB_CUDA(cudaStreamCreateWithPriority(&c.tail[i], cudaStreamNonBlocking, pr_greatest));
            MB_CUDA(cudaEventCreateWithFlags(&c.ev_acc[i], cudaEventDisableTiming));
        }
        MB_CUDA(cudaEventCreateWithFlags(&c.ev_tail, cudaEventDisableTiming));
        c.split_tail = split;
        c.have_stream = true;
    }
#endif
}

static uint32_t standalone_window(size_t n) {
    uint32_t best = 4;
    double best_cost = 1e300;
    for (uint32_t c = 4; c <= 18; ++c) {
        double nw = msm_nwin(c);
        double cost = (double)n * nw + 4.0 * nw * (double)(1u << (c - 1));
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return env_u32("MB200_C_MSM", best);
}

template <class F>
static void msm_on_device_bases(const Affine<F>* bases, const uint8_t* scalars_host, size_t n, XYZZ<F>* out_dev) {
    if (n == 0) {
        dev_memset(out_dev, 0, sizeof(XYZZ<F>), g.main);
        return;
    }
    if (n >= (1ull << 31)) fail(MB200_EINVAL, "MSM of %s%ld bases is too large", "", (long)n);
    DevBuf sel(n * 4), pool(n * 32), flag(4);
    IotaArgs ia{n, sel.as<uint32_t>()};
    launch_iota_kernel(ia, g.main);
    copy_h2d(pool.p, scalars_host, n * 32, g.main);
    dev_memset(flag.p, 0, 4, g.main);
    ValidateArgs va{n, pool.as<Fr>(), n, n, flag.as<uint32_t>()};
    launch_validate_scalars(va, g.main);
    MsmClass k = msm_make_class(bases, sel.as<uint32_t>(), (uint32_t)n, standalone_window(n), false);
    msm_run<F>(k, 1, pool.as<uint32_t>(), n, out_dev, g.msm, g.main);
    uint32_t bad = 0;
    copy_d2h(&bad, flag.p, 4, g.main);
    stream_sync(g.main);
    if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
}

static NttCache& ntt_cache(unsigned log_n) {
    auto it = g.ntt.find(log_n);
    if (it != g.ntt.end()) return *it->second;
    NttCache* c = new NttCache();
    c->dom.build(log_n, g.main);
    size_t n = (size_t)1 << log_n;
    c->gpow.alloc(n * sizeof(Fr));
    PowArgs pa;
    pa.nthreads = n;
    pa.out = c->gpow.as<Fr>();
    pa.base = fr_from_u64_host(7);
    pa.scale = Fr::one();
    launch_fr_powers(pa, g.main);
    g.ntt[log_n] = c;
    return *c;
}

static void check_scalars_dev(const Fr* base, size_t per_row, size_t row_stride, size_t nrows, uint32_t* flag,
                              cudaStream_t s) {
    ValidateArgs va{per_row * nrows, base, per_row, row_stride, flag};
    launch_validate_scalars(va, s);
}

// Buffers a batch needs for its lifetime.  They are recycled across batches: cudaMalloc /
// cudaFree / cudaFreeHost synchronise the whole device, which would serialise batches that are
// meant to overlap (submit k+1 while k is still running, wait k while k+1 runs).
struct TicketRes {
    uint8_t* staged = nullptr;     // pinned: proofs land here
    size_t staged_cap = 0;
    uint32_t* verdicts = nullptr;  // pinned: per-proof result of the self-check (option "verify")
    size_t verdicts_cap = 0;
    DevBuf v_a, v_b, v_c, v_in, v_ok;  // self-check inputs / verdicts of this batch (device)
    DevBuf flag;                   // set by the canonical-scalar checks of this batch
    ~TicketRes() {
        host_free_pinned(staged);
        host_free_pinned(verdicts);
    }
};
static std::vector<TicketRes*> g_res_free;
static TicketRes* res_acquire() {
    if (g_res_free.empty()) return new TicketRes();
    TicketRes* r = g_res_free.back();
    g_res_free.pop_back();
    return r;
}

// One submitted batch: everything needed to finish it later.
struct Ticket {
    TicketRes* res = nullptr;
    uint8_t* staged = nullptr;     // = res->staged
    uint32_t* verdicts = nullptr;  // = res->verdicts when the self-check is on
#ifndef MB200_EMU
    cudaEvent_t ev_v[4] = {nullptr, nullptr, nullptr, nullptr};
#endif
    uint8_t* out = nullptr;     // caller's buffer
    size_t n_proofs = 0;
    uint32_t* flag = nullptr;   // = res->flag
#ifndef MB200_EMU
    cudaEvent_t ev0 = nullptr;
    std::vector<cudaEvent_t> ev1;
#endif
};
static std::map<uint64_t, Ticket*> g_tickets;
static uint64_t g_next_ticket = 1;
static size_t g_next_ctx = 0;  // chunks go round the contexts across batches, so consecutive batches overlap

static void ticket_destroy(Ticket* t) {
    if (!t) return;
#ifndef MB200_EMU
    if (t->ev0) cudaEventDestroy(t->ev0);
    for (auto& e : t->ev_v)
        if (e) cudaEventDestroy(e);
    for (auto& e : t->ev1)
        if (e) cudaEventDestroy(e);
#endif
    if (t->res) g_res_free.push_back(t->res);
    delete t;
}

// Enqueue a whole batch on the chunk contexts; no host synchronisation.
static uint64_t prove_submit(const Params& P, size_t n_proofs, size_t rows, const ProveInputs& in, uint8_t* proofs_out) {
    require_init();
    if (!in.inputs || !in.aux || !in.r || !in.s || !proofs_out) fail(MB200_EINVAL, "null buffer%s", "");
    if (!in.a && !in.b && !in.c) {  // witness-only: rows come from the bound circuit
        if (!P.r1cs.bound) fail(MB200_EINVAL, "no circuit bound to these parameters%s", "");
        if (rows != (size_t)P.r1cs.ncons + P.n_inputs) fail(MB200_EINVAL, "rows do not match the bound circuit%s", "");
    } else if (!in.a || !in.b || !in.c) {
        fail(MB200_EINVAL, "null buffer%s", "");
    }
    size_t m = 1;
    while (m < rows) m <<= 1;
    if (rows == 0 || m != P.m)
        fail(MB200_EINVAL, "rows does not match the key's domain%s (key m = %ld)", "", (long)P.m);
    Ticket* t = new Ticket();
    try {
        t->n_proofs = n_proofs;
        t->out = proofs_out;
        TicketRes* R = t->res = res_acquire();
        if (R->staged_cap < n_proofs * 192) {
            host_free_pinned(R->staged);
            R->staged = nullptr;
            R->staged = (uint8_t*)host_alloc_pinned(n_proofs * 192);
            R->staged_cap = n_proofs * 192;
        }
        t->staged = R->staged;
        VerifySink sink;
        if (g.verify) {
            if (R->verdicts_cap < n_proofs * 4) {
                host_free_pinned(R->verdicts);
                R->verdicts = nullptr;
                R->verdicts = (uint32_t*)host_alloc_pinned(n_proofs * 4);
                R->verdicts_cap = n_proofs * 4;
            }
            t->verdicts = R->verdicts;
            R->v_a.ensure(n_proofs * sizeof(G1Affine));
            R->v_b.ensure(n_proofs * sizeof(G2Affine));
            R->v_c.ensure(n_proofs * sizeof(G1Affine));
            R->v_in.ensure(n_proofs * (size_t)P.n_inputs * 32);
            R->v_ok.ensure(n_proofs * 4);
            sink.a = R->v_a.as<G1Affine>();
            sink.b = R->v_b.as<G2Affine>();
            sink.c = R->v_c.as<G1Affine>();
            sink.inputs = R->v_in.as<uint32_t>();
            sink.ok_dev = R->v_ok.as<uint32_t>();
            sink.ok_host = t->verdicts;
        }
        R->flag.ensure(4);
        t->flag = R->flag.as<uint32_t>();
        dev_memset(t->flag, 0, 4, g.main);
#ifndef MB200_EMU
        // every chunk stream starts after `ev0` (and so after the flag reset)
        t->ev1.assign(g.ctxs.size(), nullptr);
        MB_CUDA(cudaEventCreate(&t->ev0));
        for (auto& e : t->ev1) MB_CUDA(cudaEventCreate(&e));
        MB_CUDA(cudaEventRecord(t->ev0, g.main));
        for (auto& x : g.ctxs) MB_CUDA(cudaStreamWaitEvent(x.stream, t->ev0, 0));
#endif
        for (size_t first = 0; first < n_proofs; first += g.chunk) {
            uint32_t count = (uint32_t)std::min<size_t>(g.chunk, n_proofs - first);
            ProveCtx& x = g.ctxs[g_next_ctx++ % g.ctxs.size()];
#ifndef MB200_EMU
            sink.stream = g.vstream[g.next_v++ % 4];
#endif
            prove_chunk(P, x, in, first, count, rows, t->staged, sink);
            // canonical-scalar check on what was just staged (abc and aux..s of the pool)
            check_scalars_dev(x.abc.as<Fr>(), (size_t)count * 3 * rows, 0, 1, t->flag, x.stream);
            check_scalars_dev(x.pool.as<Fr>() + P.idx_aux, P.idx_one - P.idx_aux, P.pool_stride, count,
                              t->flag, x.stream);
        }
#ifndef MB200_EMU
        for (size_t i = 0; i < g.ctxs.size(); ++i) MB_CUDA(cudaEventRecord(t->ev1[i], g.ctxs[i].stream));
        if (g.verify)
            for (int i = 0; i < 4; ++i) {
                MB_CUDA(cudaEventCreateWithFlags(&t->ev_v[i], cudaEventDisableTiming));
                MB_CUDA(cudaEventRecord(t->ev_v[i], g.vstream[i]));
            }
#endif
    } catch (...) {
#ifndef MB200_EMU
        // nothing of this batch may still be running when its buffers go back to the pool
        for (auto& x : g.ctxs) cudaStreamSynchronize(x.stream);
        for (auto& v : g.vstream) cudaStreamSynchronize(v);
#endif
        ticket_destroy(t);
        throw;
    }
    uint64_t id = g_next_ticket++;
    g_tickets[id] = t;
    return id;
}

// Wait for a submitted batch, check its flag, hand the proofs to the caller.
static void prove_wait(uint64_t id) {
    auto it = g_tickets.find(id);
    if (it == g_tickets.end()) fail(MB200_EINVAL, "unknown ticket%s (%ld)", "", (long)id);
    Ticket* t = it->second;
    g_tickets.erase(it);
    uint32_t bad = 0;
    try {
#ifndef MB200_EMU
        g.last_batch_ms = 0;
        for (auto& e : t->ev1) {
            MB_CUDA(cudaEventSynchronize(e));
            float ms = 0;
            MB_CUDA(cudaEventElapsedTime(&ms, t->ev0, e));
            if (ms > g.last_batch_ms) g.last_batch_ms = ms;
        }
#endif
#ifndef MB200_EMU
        for (auto& e : t->ev_v)
            if (e) MB_CUDA(cudaEventSynchronize(e));
#endif
        copy_d2h(&bad, t->flag, 4, g.main);
        stream_sync(g.main);
        if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
        memcpy(t->out, t->staged, t->n_proofs * 192);
        if (t->verdicts)
            for (size_t i = 0; i < t->n_proofs; ++i) {
                g.verified++;
                if (!t->verdicts[i]) {
                    g.verify_failed++;
                    if (g.verify == 1)
                        fail(MB200_EVERIFY, "proof %s%ld does not satisfy the verification equation", "", (long)i);
                }
            }
    } catch (...) {
        ticket_destroy(t);
        throw;
    }
    ticket_destroy(t);
}

// CSR matrices of a recorded circuit -> device, columns rewritten to scalar-pool
// indices, coefficients interned into a dictionary (Montgomery form; 0 -> +1, 1 -> -1).
static void r1cs_upload(R1csDev& R, const mb200_circuit& c, size_t idx_aux, size_t idx_inputs) {
    std::map<std::array<uint64_t, 4>, uint32_t> dict;
    std::vector<mbh::Fr> dict_vals;
    auto intern = [&](const mbh::Fr& f) {
        std::array<uint64_t, 4> k = {f.v[0], f.v[1], f.v[2], f.v[3]};
        auto it = dict.find(k);
        if (it != dict.end()) return it->second;
        uint32_t id = (uint32_t)dict_vals.size();
        dict[k] = id;
        dict_vals.push_back(f);
        return id;
    };
    intern(mbh::Fr::one());
    intern(-mbh::Fr::one());
    const mbh::Matrix* ms[3] = {&c.A, &c.B, &c.C};
    for (int k = 0; k < 3; ++k) {
        const mbh::Matrix& M = *ms[k];
        std::vector<uint32_t> col(M.col.size()), cidx(M.col.size());
        for (size_t e = 0; e < M.col.size(); ++e) {
            uint32_t id = M.col[e];
            col[e] = (id & mbh::Var::AUX) ? (uint32_t)idx_aux + (id & ~mbh::Var::AUX) : (uint32_t)idx_inputs + id;
            cidx[e] = intern(M.coef[e]);
        }
        R.rowptr[k].alloc(M.rowptr.size() * 4);
        R.col[k].alloc(col.size() * 4);
        R.cidx[k].alloc(cidx.size() * 4);
        copy_h2d(R.rowptr[k].p, M.rowptr.data(), M.rowptr.size() * 4, g.main);
        copy_h2d(R.col[k].p, col.data(), col.size() * 4, g.main);
        copy_h2d(R.cidx[k].p, cidx.data(), cidx.size() * 4, g.main);
        stream_sync(g.main);  // the staging vectors die at the end of this iteration
    }
    {  // rows by descending total non-zero count (stable: equal rows stay in constraint order)
        std::vector<uint32_t> order(c.n_constraints);
        for (uint32_t i = 0; i < c.n_constraints; ++i) order[i] = i;
        auto cost = [&](uint32_t r) {
            return (c.A.rowptr[r + 1] - c.A.rowptr[r]) + (c.B.rowptr[r + 1] - c.B.rowptr[r]) +
                   (c.C.rowptr[r + 1] - c.C.rowptr[r]);
        };
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cost(x) > cost(y); });
        R.order.alloc(order.size() * 4 + 4);
        copy_h2d(R.order.p, order.data(), order.size() * 4, g.main);
        stream_sync(g.main);
    }
    // mbh::Fr (4 x u64 Montgomery, R = 2^256) has the same memory image as the device Fr (8 x u32)
    R.dict.alloc(dict_vals.size() * 32);
    copy_h2d(R.dict.p, dict_vals.data(), dict_vals.size() * 32, g.main);
    stream_sync(g.main);
    R.ncons = c.n_constraints;
    R.n_inputs = c.n_inputs;
    R.n_aux = c.n_aux;
    R.bound = true;
}

static int prove_impl(const Params& P, size_t n_proofs, size_t rows, const ProveInputs& in, uint8_t* proofs_out) {
    require_init();
    if (n_proofs == 0) return MB200_OK;
    prove_wait(prove_submit(P, n_proofs, rows, in, proofs_out));
    return MB200_OK;
}

}  // namespace mb

using namespace mb;

#define MB_API_BEGIN                          \
    std::lock_guard<std::mutex> _lk(g.mu);    \
    try {
#define MB_API_END                            \
    }                                         \
    catch (const Exc& e) { return e.code; }   \
    catch (const std::bad_alloc&) {           \
        last_error().code = MB200_ENOMEM;     \
        snprintf(last_error().msg, sizeof last_error().msg, "host allocation failed"); \
        return MB200_ENOMEM;                  \
    }                                         \
    return MB200_OK;

struct mb200_params {
    Params* p;
};

extern "C" {

int mb200_init(const int* device_ids, int n_devices) {
    MB_API_BEGIN
    if (g.inited) return MB200_OK;
#ifndef MB200_EMU
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        fail(MB200_ECUDA, "no CUDA device: %s (this library has no CPU path)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    int dev = 0;
    if (device_ids && n_devices > 0) dev = device_ids[0];
    else MB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= count) fail(MB200_EINVAL, "device id out of range%s (%ld)", "", (long)dev);
    MB_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    MB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) fail(MB200_ECUDA, "device %s is not sm_100 (compute capability %ld.x)", prop.name, (long)prop.major);
    g.device = dev;
    MB_CUDA(cudaStreamCreateWithFlags(&g.main, cudaStreamNonBlocking));
    for (auto& v : g.vstream) MB_CUDA(cudaStreamCreateWithFlags(&v, cudaStreamNonBlocking));
#else
    (void)device_ids;
    (void)n_devices;
#endif
    g.chunk = env_u32("MB200_CHUNK", 64);
    set_ctx_count(env_u32("MB200_STREAMS", 3));
    g.inited = true;
    MB_API_END
}

int mb200_shutdown(void) {
    MB_API_BEGIN
    if (!g.inited) return MB200_OK;
    for (auto& kv : g.ntt) delete kv.second;
    g.ntt.clear();
    set_ctx_count(0);
    for (auto* r : g_res_free) delete r;
    g_res_free.clear();
    g.msm = MsmScratch();
#ifndef MB200_EMU
    cudaStreamDestroy(g.main);
    for (auto& v : g.vstream) {
        cudaStreamSynchronize(v);
        cudaStreamDestroy(v);
    }
#endif
    g.inited = false;
    MB_API_END
}

int mb200_params_load(const uint8_t* bytes, size_t len, const uint8_t* a_aux_density, const uint8_t* b_input_density,
                      const uint8_t* b_aux_density, mb200_params** out) {
    MB_API_BEGIN
    require_init();
    if (!out) fail(MB200_EINVAL, "null out pointer%s", "");
    *out = nullptr;
    Params* p = params_load(bytes, len, a_aux_density, b_input_density, b_aux_density, g.main);
    *out = new mb200_params{p};
    MB_API_END
}

int mb200_params_info(const mb200_params* p, uint64_t info[10]) {
    MB_API_BEGIN
    if (!p || !p->p || !info) fail(MB200_EINVAL, "null argument%s", "");
    const Params& P = *p->p;
    info[0] = P.n_inputs; info[1] = P.n_aux; info[2] = P.h_len; info[3] = P.a_len; info[4] = P.b_len;
    info[5] = P.m; info[6] = P.consumed; info[7] = P.table_bytes; info[8] = P.k_hl.c; info[9] = P.k_a.c;
    MB_API_END
}

void mb200_params_free(mb200_params* p) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!p) return;
#ifndef MB200_EMU
    cudaDeviceSynchronize();
#endif
    delete p->p;
    delete p;
}

size_t mb200_params_synth_size(uint32_t n_inputs, uint32_t h_len, uint32_t l_len, uint32_t a_len, uint32_t b_len) {
    return params_synth_size(n_inputs, h_len, l_len, a_len, b_len);
}

int mb200_params_synthesize(uint64_t seed, uint32_t n_inputs, uint32_t h_len, uint32_t l_len, uint32_t a_len,
                            uint32_t b_len, uint8_t* out, size_t out_len) {
    MB_API_BEGIN
    require_init();
    size_t need = params_synth_size(n_inputs, h_len, l_len, a_len, b_len);
    if (!out || out_len < need) fail(MB200_EINVAL, "output buffer too small%s (need %ld bytes)", "", (long)need);
    DevBuf d(need);
    params_synthesize(seed, n_inputs, h_len, l_len, a_len, b_len, d.as<uint8_t>(), g.main);
    copy_d2h(out, d.p, need, g.main);
    stream_sync(g.main);
    MB_API_END
}

int mb200_synth_points(uint64_t seed, uint32_t stream, uint64_t start, size_t n, int group, uint8_t* out) {
    MB_API_BEGIN
    require_init();
    if (!out || (group != 1 && group != 2)) fail(MB200_EINVAL, "bad argument%s", "");
    size_t bytes = n * (group == 1 ? 96 : 192);
    DevBuf d(bytes);
    SynthArgs a;
    a.nthreads = n;
    a.key = stream_key(seed, stream);
    a.start = start;
    a.out = d.as<uint8_t>();
    a.g1 = g1_generator_host();
    a.g2 = g2_generator_host();
    if (group == 1) launch_synth_g1(a, g.main);
    else launch_synth_g2(a, g.main);
    copy_d2h(out, d.p, bytes, g.main);
    stream_sync(g.main);
    MB_API_END
}

int mb200_prove_batch(const mb200_params* p, size_t n_proofs, size_t rows, const uint8_t* a_evals,
                      const uint8_t* b_evals, const uint8_t* c_evals, const uint8_t* inputs, const uint8_t* aux,
                      const uint8_t* r, const uint8_t* s, uint8_t* proofs_out) {
    MB_API_BEGIN
    if (!p || !p->p) fail(MB200_EINVAL, "null parameters%s", "");
    ProveInputs in{a_evals, b_evals, c_evals, inputs, aux, r, s, false};
    return prove_impl(*p->p, n_proofs, rows, in, proofs_out);
    MB_API_END
}

int mb200_prove_batch_device(const mb200_params* p, size_t n_proofs, size_t rows, const void* a_evals,
                             const void* b_evals, const void* c_evals, const void* inputs, const void* aux,
                             const void* r, const void* s, uint8_t* proofs_out) {
    MB_API_BEGIN
    if (!p || !p->p) fail(MB200_EINVAL, "null parameters%s", "");
    ProveInputs in{(const uint8_t*)a_evals, (const uint8_t*)b_evals, (const uint8_t*)c_evals, (const uint8_t*)inputs,
                   (const uint8_t*)aux, (const uint8_t*)r, (const uint8_t*)s, true};
    return prove_impl(*p->p, n_proofs, rows, in, proofs_out);
    MB_API_END
}

int mb200_prove_submit(const mb200_params* p, size_t n_proofs, size_t rows, const void* a_evals, const void* b_evals,
                       const void* c_evals, const void* inputs, const void* aux, const void* r, const void* s,
                       int on_device, uint8_t* proofs_out, uint64_t* ticket) {
    MB_API_BEGIN
    if (!p || !p->p || !ticket || n_proofs == 0) fail(MB200_EINVAL, "bad argument%s", "");
    ProveInputs in{(const uint8_t*)a_evals, (const uint8_t*)b_evals, (const uint8_t*)c_evals, (const uint8_t*)inputs,
                   (const uint8_t*)aux, (const uint8_t*)r, (const uint8_t*)s, on_device != 0};
    *ticket = prove_submit(*p->p, n_proofs, rows, in, proofs_out);
    MB_API_END
}

int mb200_params_bind_circuit(mb200_params* p, const mb200_circuit* c) {
    MB_API_BEGIN
    require_init();
    if (!p || !p->p || !c) fail(MB200_EINVAL, "null argument%s", "");
    Params& P = *p->p;
    if (P.n_inputs != c->n_inputs || P.n_aux != c->n_aux)
        fail(MB200_EINVAL, "circuit and key disagree on the variable counts%s", "");
    if (P.a_len != c->n_inputs + c->a_aux_ones || P.b_len != c->b_input_ones + c->b_aux_ones)
        fail(MB200_EINVAL, "circuit densities do not match the key's query lengths%s", "");
    size_t m = 1;
    while (m < (size_t)c->n_constraints + c->n_inputs) m <<= 1;
    if (m != P.m) fail(MB200_EINVAL, "circuit size does not match the key's domain%s", "");
    r1cs_upload(P.r1cs, *c, P.idx_aux, P.idx_inputs);
    MB_API_END
}

int mb200_circuit_rows(const mb200_circuit* c, size_t n, const uint8_t* inputs, const uint8_t* aux, uint8_t* a_out,
                       uint8_t* b_out, uint8_t* c_out) {
    MB_API_BEGIN
    require_init();
    if (!c || (n && (!inputs || !aux || !a_out || !b_out || !c_out))) fail(MB200_EINVAL, "null argument%s", "");
    if (n == 0) return MB200_OK;
    R1csDev R;
    r1cs_upload(R, *c, 0, c->n_aux);
    const size_t stride = (size_t)c->n_aux + c->n_inputs, rows = (size_t)c->n_constraints + c->n_inputs;
    DevBuf pool(n * stride * 32), abc(n * 3 * rows * 32), flag(4);
    dev_memset(flag.p, 0, 4, g.main);
    uint8_t* pl = pool.as<uint8_t>();
    copy_rows(pl, stride * 32, aux, (size_t)c->n_aux * 32, (size_t)c->n_aux * 32, n, false, g.main);
    copy_rows(pl + (size_t)c->n_aux * 32, stride * 32, inputs, (size_t)c->n_inputs * 32, (size_t)c->n_inputs * 32, n, false,
              g.main);
    check_scalars_dev(pool.as<Fr>(), n * stride, 0, 1, flag.as<uint32_t>(), g.main);
    R1csArgs ra;
    ra.nthreads = n * rows;
    for (int k = 0; k < 3; ++k) {
        ra.rowptr[k] = R.rowptr[k].as<uint32_t>();
        ra.col[k] = R.col[k].as<uint32_t>();
        ra.cidx[k] = R.cidx[k].as<uint32_t>();
    }
    ra.dict = R.dict.as<Fr>();
    ra.ncons = R.ncons;
    ra.rows = (uint32_t)rows;
    ra.pool = pool.as<Fr>();
    ra.pool_stride = stride;
    ra.idx_inputs = c->n_aux;
    ra.abc = abc.as<Fr>();
    ra.order = R.order.as<uint32_t>();
    launch_r1cs_eval(ra, g.main);
    uint32_t bad = 0;
    copy_d2h(&bad, flag.p, 4, g.main);
    uint8_t* outs[3] = {a_out, b_out, c_out};
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k)
            copy_d2h(outs[k] + i * rows * 32, abc.as<uint8_t>() + (i * 3 + k) * rows * 32, rows * 32, g.main);
    stream_sync(g.main);
    if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
    MB_API_END
}

int mb200_verify_batch(const mb200_params* p, size_t n, const uint8_t* proofs_uncompressed, const uint8_t* inputs,
                       uint8_t* ok_out) {
    MB_API_BEGIN
    require_init();
    if (!p || !p->p || (n && (!proofs_uncompressed || !inputs || !ok_out))) fail(MB200_EINVAL, "null argument%s", "");
    if (n == 0) return MB200_OK;
    const Params& P = *p->p;
    // A | B | C in zkcrypto uncompressed form -> three device arrays
    std::vector<uint8_t> ha(n * 96), hb(n * 192), hc(n * 96);
    for (size_t i = 0; i < n; ++i) {
        memcpy(&ha[i * 96], proofs_uncompressed + i * 384, 96);
        memcpy(&hb[i * 192], proofs_uncompressed + i * 384 + 96, 192);
        memcpy(&hc[i * 96], proofs_uncompressed + i * 384 + 288, 96);
    }
    DevBuf ra(n * 96), rb(n * 192), rc(n * 96), pa(n * sizeof(G1Affine)), pb(n * sizeof(G2Affine)),
        pc(n * sizeof(G1Affine)), bad(4), din(n * P.n_inputs * 32), ok(n * 4);
    copy_h2d(ra.p, ha.data(), ha.size(), g.main);
    copy_h2d(rb.p, hb.data(), hb.size(), g.main);
    copy_h2d(rc.p, hc.data(), hc.size(), g.main);
    copy_h2d(din.p, inputs, n * P.n_inputs * 32, g.main);
    dev_memset(bad.p, 0, 4, g.main);
    DecodeArgs d1{n, ra.as<uint8_t>(), pa.p, bad.as<uint32_t>()}, d2{n, rb.as<uint8_t>(), pb.p, bad.as<uint32_t>()},
        d3{n, rc.as<uint8_t>(), pc.p, bad.as<uint32_t>()};
    launch_decode_g1(d1, g.main);
    launch_decode_g2(d2, g.main);
    launch_decode_g1(d3, g.main);
    check_scalars_dev(din.as<Fr>(), n * P.n_inputs, 0, 1, bad.as<uint32_t>(), g.main);
    VerifyArgs va;
    va.nthreads = n;
    va.pa = pa.as<G1Affine>();
    va.pb = pb.as<G2Affine>();
    va.pc = pc.as<G1Affine>();
    va.inputs = din.as<uint32_t>();
    va.input_stride = P.n_inputs;
    va.n_inputs = P.n_inputs;
    va.vk = {P.vk_ic.as<G1Affine>(), P.vk_g2.as<G2Affine>() + 1, P.vk_g2.as<G2Affine>() + 2, P.vk_ab.as<Fp12>()};
    va.ok = ok.as<uint32_t>();
    launch_verify_proofs(va, g.main);
    std::vector<uint32_t> hok(n);
    uint32_t hbad = 0;
    copy_d2h(hok.data(), ok.p, n * 4, g.main);
    copy_d2h(&hbad, bad.p, 4, g.main);
    stream_sync(g.main);
    if (hbad) fail(MB200_EPARSE, "malformed point or non-canonical input%s", "");
    for (size_t i = 0; i < n; ++i) ok_out[i] = hok[i] ? 1 : 0;
    MB_API_END
}

int mb200_verify_proofs(const mb200_params* p, size_t n, const uint8_t* proofs, const uint8_t* inputs, uint8_t* ok_out) {
    MB_API_BEGIN
    require_init();
    if (!p || !p->p || (n && (!proofs || !inputs || !ok_out))) fail(MB200_EINVAL, "null argument%s", "");
    if (n == 0) return MB200_OK;
    const Params& P = *p->p;
    DevBuf raw(n * 192), pa(n * sizeof(G1Affine)), pb(n * sizeof(G2Affine)), pc(n * sizeof(G1Affine)), bad(n * 4),
        flag(4), din(n * P.n_inputs * 32), ok(n * 4);
    copy_h2d(raw.p, proofs, n * 192, g.main);
    copy_h2d(din.p, inputs, n * P.n_inputs * 32, g.main);
    dev_memset(bad.p, 0, n * 4, g.main);
    dev_memset(flag.p, 0, 4, g.main);
    ProofReadArgs ra{n * 3, raw.as<uint8_t>(), pa.as<G1Affine>(), pb.as<G2Affine>(), pc.as<G1Affine>(), bad.as<uint32_t>()};
    launch_proof_read(ra, g.main);
    check_scalars_dev(din.as<Fr>(), n * P.n_inputs, 0, 1, flag.as<uint32_t>(), g.main);
    VerifyArgs va;
    va.nthreads = n;
    va.pa = pa.as<G1Affine>();
    va.pb = pb.as<G2Affine>();
    va.pc = pc.as<G1Affine>();
    va.inputs = din.as<uint32_t>();
    va.input_stride = P.n_inputs;
    va.n_inputs = P.n_inputs;
    va.vk = {P.vk_ic.as<G1Affine>(), P.vk_g2.as<G2Affine>() + 1, P.vk_g2.as<G2Affine>() + 2, P.vk_ab.as<Fp12>()};
    va.ok = ok.as<uint32_t>();
    launch_verify_proofs(va, g.main);
    std::vector<uint32_t> hok(n), hbad(n);
    uint32_t hflag = 0;
    copy_d2h(hok.data(), ok.p, n * 4, g.main);
    copy_d2h(hbad.data(), bad.p, n * 4, g.main);
    copy_d2h(&hflag, flag.p, 4, g.main);
    stream_sync(g.main);
    if (hflag) fail(MB200_ESCALAR, "a public input is not canonical (>= r)%s", "");
    for (size_t i = 0; i < n; ++i) ok_out[i] = (hok[i] && !hbad[i]) ? 1 : 0;
    MB_API_END
}

int mb200_verify_proofs_batch(const mb200_params* p, size_t n, const uint8_t* proofs, const uint8_t* inputs,
                              const uint8_t* z, int* all_ok) {
    MB_API_BEGIN
    require_init();
    if (!p || !p->p || !all_ok || (n && (!proofs || !inputs || !z))) fail(MB200_EINVAL, "null argument%s", "");
    *all_ok = 1;
    if (n == 0) return MB200_OK;
    const Params& P = *p->p;
    DevBuf raw(n * 192), pa(n * sizeof(G1Affine)), pb(n * sizeof(G2Affine)), pc(n * sizeof(G1Affine)), bad(n * 4),
        flag(4), din(n * P.n_inputs * 32), dz(n * 16), f(n * sizeof(Fp12)), zc(n * sizeof(G1XYZZ)),
        zx(n * P.n_inputs * sizeof(Fr)), ok(4);
    copy_h2d(raw.p, proofs, n * 192, g.main);
    copy_h2d(din.p, inputs, n * P.n_inputs * 32, g.main);
    copy_h2d(dz.p, z, n * 16, g.main);
    dev_memset(bad.p, 0, n * 4, g.main);
    dev_memset(flag.p, 0, 4, g.main);
    ProofReadArgs ra{n * 3, raw.as<uint8_t>(), pa.as<G1Affine>(), pb.as<G2Affine>(), pc.as<G1Affine>(), bad.as<uint32_t>()};
    launch_proof_read(ra, g.main);
    check_scalars_dev(din.as<Fr>(), n * P.n_inputs, 0, 1, flag.as<uint32_t>(), g.main);
    BatchMillerArgs ma{n, pa.as<G1Affine>(), pb.as<G2Affine>(), pc.as<G1Affine>(), din.as<uint32_t>(), P.n_inputs,
                       dz.as<uint32_t>(), f.as<Fp12>(), zc.as<G1XYZZ>(), zx.as<Fr>()};
    launch_batch_miller(ma, g.main);
    BatchFinalArgs fa;
    fa.nthreads = 1;
    fa.n = n;
    fa.f = f.as<Fp12>();
    fa.zc = zc.as<G1XYZZ>();
    fa.zx = zx.as<Fr>();
    fa.n_inputs = P.n_inputs;
    fa.alpha = P.vk_g1.as<G1Affine>();
    fa.beta = P.vk_g2.as<G2Affine>();
    fa.vk = {P.vk_ic.as<G1Affine>(), P.vk_g2.as<G2Affine>() + 1, P.vk_g2.as<G2Affine>() + 2, P.vk_ab.as<Fp12>()};
    fa.ok = ok.as<uint32_t>();
    launch_batch_final(fa, g.main);
    std::vector<uint32_t> hbad(n);
    uint32_t hok = 0, hflag = 0;
    copy_d2h(&hok, ok.p, 4, g.main);
    copy_d2h(hbad.data(), bad.p, n * 4, g.main);
    copy_d2h(&hflag, flag.p, 4, g.main);
    stream_sync(g.main);
    if (hflag) fail(MB200_ESCALAR, "a public input is not canonical (>= r)%s", "");
    int good = hok ? 1 : 0;
    for (size_t i = 0; i < n; ++i)
        if (hbad[i]) good = 0;
    *all_ok = good;
    MB_API_END
}

int mb200_prove_batch_witness(const mb200_params* p, size_t n_proofs, const uint8_t* inputs, const uint8_t* aux,
                              const uint8_t* r, const uint8_t* s, uint8_t* proofs_out) {
    MB_API_BEGIN
    if (!p || !p->p) fail(MB200_EINVAL, "null parameters%s", "");
    ProveInputs in{nullptr, nullptr, nullptr, inputs, aux, r, s, false};
    return prove_impl(*p->p, n_proofs, (size_t)p->p->r1cs.ncons + p->p->n_inputs, in, proofs_out);
    MB_API_END
}

int mb200_prove_wait(uint64_t ticket) {
    MB_API_BEGIN
    require_init();
    prove_wait(ticket);
    MB_API_END
}

int mb200_msm_g1(const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t out[96]) {
    MB_API_BEGIN
    require_init();
    if ((n && (!bases || !scalars)) || !out) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf raw(n * 96), pts(n * sizeof(G1Affine)), bad(4), res(sizeof(G1XYZZ)), enc(96);
    copy_h2d(raw.p, bases, n * 96, g.main);
    dev_memset(bad.p, 0, 4, g.main);
    DecodeArgs da{n, raw.as<uint8_t>(), pts.p, bad.as<uint32_t>()};
    launch_decode_g1(da, g.main);
    uint32_t bad_h = 0;
    copy_d2h(&bad_h, bad.p, 4, g.main);
    stream_sync(g.main);
    if (bad_h) fail(MB200_EPARSE, "malformed G1 encoding%s", "");
    msm_on_device_bases<Fp>(pts.as<G1Affine>(), scalars, n, res.as<G1XYZZ>());
    EncodeArgs ea{1, res.p, enc.as<uint8_t>()};
    launch_encode_g1(ea, g.main);
    copy_d2h(out, enc.p, 96, g.main);
    stream_sync(g.main);
    MB_API_END
}

int mb200_msm_g2(const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t out[192]) {
    MB_API_BEGIN
    require_init();
    if ((n && (!bases || !scalars)) || !out) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf raw(n * 192), pts(n * sizeof(G2Affine)), bad(4), res(sizeof(G2XYZZ)), enc(192);
    copy_h2d(raw.p, bases, n * 192, g.main);
    dev_memset(bad.p, 0, 4, g.main);
    DecodeArgs da{n, raw.as<uint8_t>(), pts.p, bad.as<uint32_t>()};
    launch_decode_g2(da, g.main);
    uint32_t bad_h = 0;
    copy_d2h(&bad_h, bad.p, 4, g.main);
    stream_sync(g.main);
    if (bad_h) fail(MB200_EPARSE, "malformed G2 encoding%s", "");
    msm_on_device_bases<Fp2>(pts.as<G2Affine>(), scalars, n, res.as<G2XYZZ>());
    EncodeArgs ea{1, res.p, enc.as<uint8_t>()};
    launch_encode_g2(ea, g.main);
    copy_d2h(out, enc.p, 192, g.main);
    stream_sync(g.main);
    MB_API_END
}

int mb200_g1_bases_upload(const uint8_t* bases, size_t n, void** dev_bases) {
    MB_API_BEGIN
    require_init();
    if (!bases || !dev_bases) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf raw(n * 96), bad(4);
    void* pts = dev_alloc(n * sizeof(G1Affine));
    copy_h2d(raw.p, bases, n * 96, g.main);
    dev_memset(bad.p, 0, 4, g.main);
    DecodeArgs da{n, raw.as<uint8_t>(), pts, bad.as<uint32_t>()};
    launch_decode_g1(da, g.main);
    uint32_t bad_h = 0;
    copy_d2h(&bad_h, bad.p, 4, g.main);
    stream_sync(g.main);
    if (bad_h) {
        dev_free(pts);
        fail(MB200_EPARSE, "malformed G1 encoding%s", "");
    }
    *dev_bases = pts;
    MB_API_END
}

int mb200_dev_free(void* p) {
    MB_API_BEGIN
    dev_free(p);
    MB_API_END
}

int mb200_msm_g1_partial(const void* dev_bases, const uint8_t* scalars, size_t n, uint8_t out_partial[192]) {
    MB_API_BEGIN
    require_init();
    if ((n && (!dev_bases || !scalars)) || !out_partial) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf res(sizeof(G1XYZZ));
    msm_on_device_bases<Fp>((const G1Affine*)dev_bases, scalars, n, res.as<G1XYZZ>());
    copy_d2h(out_partial, res.p, sizeof(G1XYZZ), g.main);
    stream_sync(g.main);
    MB_API_END
}

int mb200_g1_sum_partials(const uint8_t* partials, size_t count, uint8_t out[96]) {
    MB_API_BEGIN
    require_init();
    if ((count && !partials) || !out) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf parts(count * sizeof(G1XYZZ)), enc(96);
    copy_h2d(parts.p, partials, count * sizeof(G1XYZZ), g.main);
    SumPartialsArgs sa{1, parts.as<G1XYZZ>(), count, enc.as<uint8_t>()};
    launch_sum_partials(sa, g.main);
    copy_d2h(out, enc.p, 96, g.main);
    stream_sync(g.main);
    MB_API_END
}

int mb200_ntt(uint8_t* data, unsigned log_n, int inverse, int coset) {
    MB_API_BEGIN
    require_init();
    if (!data) fail(MB200_EINVAL, "null buffer%s", "");
    NttCache& c = ntt_cache(log_n);
    size_t n = (size_t)1 << log_n;
    DevBuf d(n * 32), t0(n * 32), t1(n * 32), o(n * 32), flag(4);
    copy_h2d(d.p, data, n * 32, g.main);
    dev_memset(flag.p, 0, 4, g.main);
    check_scalars_dev(d.as<Fr>(), n, 0, 1, flag.as<uint32_t>(), g.main);
    NttPlan p;
    p.inverse = inverse != 0;
    if (!inverse && coset) p.in_scale = c.gpow.as<Fr>();
    if (inverse) p.out_scale = coset ? c.dom.cos_inv.as<Fr>() : c.dom.minv_tab.as<Fr>();
    ntt_run(c.dom, p, 1, d.as<Fr>(), n, o.as<Fr>(), n, t0.as<Fr>(), t1.as<Fr>(), g.main);
    uint32_t bad = 0;
    copy_d2h(&bad, flag.p, 4, g.main);
    copy_d2h(data, o.p, n * 32, g.main);
    stream_sync(g.main);
    if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
    MB_API_END
}

int mb200_h_coeffs(const uint8_t* a, const uint8_t* b, const uint8_t* c, size_t rows, uint8_t* out) {
    MB_API_BEGIN
    require_init();
    if (!a || !b || !c || !out || rows < 2) fail(MB200_EINVAL, "bad argument (rows must be >= 2)%s", "");
    unsigned log_n = 0;
    while (((size_t)1 << log_n) < rows) log_n++;
    NttCache& cache = ntt_cache(log_n);
    size_t n = (size_t)1 << log_n;
    DevBuf abc(3 * rows * 32), w0(3 * n * 32), w1(3 * n * 32), w2(3 * n * 32), w3(3 * n * 32), h(n * 32), flag(4);
    copy_h2d(abc.as<uint8_t>(), a, rows * 32, g.main);
    copy_h2d(abc.as<uint8_t>() + rows * 32, b, rows * 32, g.main);
    copy_h2d(abc.as<uint8_t>() + 2 * rows * 32, c, rows * 32, g.main);
    dev_memset(flag.p, 0, 4, g.main);
    check_scalars_dev(abc.as<Fr>(), 3 * rows, 0, 1, flag.as<uint32_t>(), g.main);
    h_pipeline(cache.dom, 1, (uint32_t)rows, abc.as<Fr>(), rows, h.as<Fr>(), n, w0.as<Fr>(), w1.as<Fr>(), w2.as<Fr>(),
               w3.as<Fr>(), g.main);
    uint32_t bad = 0;
    copy_d2h(&bad, flag.p, 4, g.main);
    copy_d2h(out, h.p, (n - 1) * 32, g.main);
    stream_sync(g.main);
    if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
    MB_API_END
}

static void fr_mul_dev(const void* a, const void* b, size_t n, void* out) {
    FrMulArgs fa{n, (const Fr*)a, (const Fr*)b, (Fr*)out};
    launch_fr_mul_kernel(fa, g.main);
}
int mb200_fr_mul(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
    MB_API_BEGIN
    require_init();
    if (n && (!a || !b || !out)) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf da(n * 32), db(n * 32), dc(n * 32);
    copy_h2d(da.p, a, n * 32, g.main);
    copy_h2d(db.p, b, n * 32, g.main);
    fr_mul_dev(da.p, db.p, n, dc.p);
    copy_d2h(out, dc.p, n * 32, g.main);
    stream_sync(g.main);
    MB_API_END
}
int mb200_fr_mul_device(const void* a, const void* b, size_t n, void* out) {
    MB_API_BEGIN
    require_init();
    if (n && (!a || !b || !out)) fail(MB200_EINVAL, "null buffer%s", "");
    fr_mul_dev(a, b, n, out);
    stream_sync(g.main);
    MB_API_END
}

int mb200_set_option(const char* name, long value) {
    MB_API_BEGIN
    require_init();
    if (!name) fail(MB200_EINVAL, "null name%s", "");
    if (!strcmp(name, "chunk")) {
        if (value < 1 || value > 4096) fail(MB200_EINVAL, "chunk out of range%s (%ld)", "", value);
        g.chunk = (uint32_t)value;
    } else if (!strcmp(name, "streams")) {
        if (value < 1 || value > 8) fail(MB200_EINVAL, "streams out of range%s (%ld)", "", value);
        set_ctx_count((size_t)value);
    } else if (!strcmp(name, "verify")) {
        g.verify = (int)value;
    } else if (!strcmp(name, "profile")) {
        g_msm_profile.enabled = value != 0;
        g_msm_profile.acc_ms = 0;
        g_msm_profile.acc_launches = 0;
        g_msm_profile.acc_entries_bound = 0;
    } else {
        fail(MB200_EINVAL, "unknown option %s", name);
    }
    MB_API_END
}

int mb200_get_counter(const char* name, double* value) {
    MB_API_BEGIN
    if (!name || !value) fail(MB200_EINVAL, "null argument%s", "");
    if (!strcmp(name, "launches")) *value = (double)g_launches;
    else if (!strcmp(name, "acc_launches")) *value = (double)g_msm_profile.acc_launches;
    else if (!strcmp(name, "acc_us")) *value = g_msm_profile.acc_ms * 1000.0;
    else if (!strcmp(name, "acc_bytes")) *value = (double)g_msm_profile.acc_entries_bound;
    else if (!strcmp(name, "last_batch_us")) *value = g.last_batch_ms * 1000.0;
    else if (!strcmp(name, "verified")) *value = (double)g.verified;
    else if (!strcmp(name, "verify_failed")) *value = (double)g.verify_failed;
    else fail(MB200_EINVAL, "unknown counter %s", name);
    MB_API_END
}

int mb200_selftest(void) {
    std::lock_guard<std::mutex> lk(g.mu);
    try {
        require_init();
        uint32_t bad = 0;
        DevBuf d(4);
        dev_memset(d.p, 0, 4, g.main);
        SelfTestArgs a;
        a.nthreads = 4096;
        a.mismatches = d.as<uint32_t>();
        a.g1 = g1_generator_host();
        a.g2 = g2_generator_host();
        launch_selftest_kernel(a, g.main);
        // pairing: x-chain final exponentiation against the plain power, Frobenius consistency
        DevBuf gens(sizeof(G1Affine) + sizeof(G2Affine));
        G1Affine hg1 = g1_generator_host();
        G2Affine hg2 = g2_generator_host();
        copy_h2d(gens.p, &hg1, sizeof hg1, g.main);
        copy_h2d((char*)gens.p + sizeof(G1Affine), &hg2, sizeof hg2, g.main);
        PairSelfTestArgs pa{1, gens.as<G1Affine>(), (const G2Affine*)((char*)gens.p + sizeof(G1Affine)), d.as<uint32_t>()};
        launch_pair_selftest(pa, g.main);
        copy_d2h(&bad, d.p, 4, g.main);
        stream_sync(g.main);
        return (int)bad;
    } catch (const Exc& e) {
        return e.code;
    }
}

int mb200_bench_fpmul(double* muls_per_second) {
    MB_API_BEGIN
    require_init();
    if (!muls_per_second) fail(MB200_EINVAL, "null argument%s", "");
#ifndef MB200_EMU
    FpMulBenchArgs a;
    a.nthreads = (size_t)148 * 2048 * 4;
    a.iters = 512;
    DevBuf sink(a.nthreads * sizeof(Fp));
    a.sink = sink.as<Fp>();
    launch_fpmul_bench(a, g.main);  // warm-up
    cudaEvent_t e0, e1;
    MB_CUDA(cudaEventCreate(&e0));
    MB_CUDA(cudaEventCreate(&e1));
    MB_CUDA(cudaEventRecord(e0, g.main));
    launch_fpmul_bench(a, g.main);
    MB_CUDA(cudaEventRecord(e1, g.main));
    MB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *muls_per_second = (double)a.nthreads * a.iters * 4 / (ms * 1e-3);
#else
    *muls_per_second = 0;
#endif
    MB_API_END
}

int mb200_bench_latency(int mode, double* ns_per_op) {
    MB_API_BEGIN
    require_init();
    if (!ns_per_op || mode < 0 || mode > 3) fail(MB200_EINVAL, "bad argument%s", "");
#ifndef MB200_EMU
    LatencyArgs a;
    a.nthreads = 32;
    a.iters = 2000;
    a.mode = (uint32_t)mode;
    a.g1 = g1_generator_host();
    DevBuf sink(32 * sizeof(Fp));
    a.sink = sink.as<Fp>();
    launch_latency_kernel(a, g.main);
    cudaEvent_t e0, e1;
    MB_CUDA(cudaEventCreate(&e0));
    MB_CUDA(cudaEventCreate(&e1));
    MB_CUDA(cudaEventRecord(e0, g.main));
    launch_latency_kernel(a, g.main);
    MB_CUDA(cudaEventRecord(e1, g.main));
    MB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ns_per_op = (double)ms * 1e6 / a.iters;
#else
    *ns_per_op = 0;
#endif
    MB_API_END
}

const char* mb200_strerror(int code) {
    switch (code) {
        case MB200_OK: return "ok";
        case MB200_EINVAL: return "invalid argument";
        case MB200_EPARSE: return "malformed parameters";
        case MB200_ECUDA: return "CUDA failure or no device";
        case MB200_ENOMEM: return "out of memory";
        case MB200_ESTATE: return "mb200_init has not been called";
        case MB200_ESCALAR: return "scalar is not canonical";
        case MB200_ESYNTH: return "witness generation failed";
        case MB200_EVERIFY: return "proof does not verify";
        default: return "unknown error";
    }
}
const char* mb200_last_error(void) { return last_error().msg; }

}  // extern "C"
