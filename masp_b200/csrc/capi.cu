// C ABI of libmasp_b200 (include/masp_b200.h).  Host-side orchestration in
// C++ above hand-written sm_100a kernels; no CPU compute path exists here:
// every entry point needs mb200_init to have found a CUDA device.
#include <algorithm>
#include <array>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "host/blake2b.hpp"
#include "host/circuit_obj.hpp"
#include "prover.cuh"
#include "synth.cuh"
#include "misc.cuh"

namespace mb {

std::atomic<unsigned long long> g_launches{0};
MsmProfile g_msm_profile;

// ---------------------------------------------------------------------------
// process state: one Device per GPU opened by mb200_init.  The reference's prover is ONE object
// shared by every thread of the process (masp_proofs/src/prover.rs:27-33, 156-261: `&self`,
// immutable parameters); here that object spans all opened GPUs: keys are replicated to every
// device, a batch is cut into per-device slices, standalone calls run on device slot 0.
// ---------------------------------------------------------------------------
struct NttCache {
    NttDomain dom;
    DevBuf gpow;  // g^i, Montgomery (standalone coset_fft)
};
struct TicketRes;
struct Device {
    int id = 0;      // CUDA ordinal
    int slot = 0;    // index into State::devs
    cudaStream_t main = 0;
    std::vector<ProveCtx> ctxs;
    cudaStream_t vstream[4] = {0, 0, 0, 0};  // the self-check's streams (round robin over chunks)
    size_t next_v = 0;
    size_t next_ctx = 0;  // chunks go round the contexts across batches, so consecutive batches overlap
    std::map<unsigned, NttCache*> ntt;
    MsmScratch msm;
    std::vector<TicketRes*> res_free;
    DevBuf peer_parts;  // slot 0 only: where the other devices' MSM partials land (K6)
    // standalone MSM: scalars arrive in slabs through two buffers, uploaded on `copy` while the
    // previous slab is being sorted and accumulated on `main`.  All of it persists across calls:
    // cudaMalloc / cudaFree per call cost more than a 2^16 MSM itself.
    cudaStream_t copy = 0;
    DevBuf slab[2], msm_flag, msm_enc;
#ifndef MB200_EMU
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
#endif
};
struct State {
    bool inited = false;
    std::vector<std::unique_ptr<Device>> devs;
    uint32_t chunk = 64;
    uint32_t n_streams = 3;
    int verify = 0;  // option "verify": 1 = check every proof before returning it (failure = MB200_EVERIFY),
                     // 2 = check and only count the failures (counter "verify_failed"; for measuring the cost)
    unsigned long long verify_failed = 0, verified = 0;
    size_t next_dev = 0;  // small batches rotate over the devices
    size_t msm_slab = (size_t)1 << 22;  // standalone MSM: scalars per upload / sort / accumulate slab (option "msm_slab")
    std::mutex mu;
    double last_batch_ms = 0;  // device time of the last prove call (CUDA events on the chunk streams, max over devices)
};
static State g;

static void require_init() {
    if (!g.inited) fail(MB200_ESTATE, "mb200_init has not been called%s", "");
}
// Make `d` the calling thread's current device (CUDA's current device is per host thread).
static inline void use(const Device& d) {
#ifndef MB200_EMU
    MB_CUDA(cudaSetDevice(d.id));
#else
    (void)d;
#endif
}
static inline Device& dev0() {
    require_init();
    Device& d = *g.devs[0];
    use(d);
    return d;
}

// Runs fn(device) for every listed device, one host thread per device beyond the first, and
// rethrows the first failure on the calling thread (fail() records its message per thread).
template <class Fn>
static void for_devices(const std::vector<Device*>& ds, Fn fn) {
    if (ds.empty()) return;
    if (ds.size() == 1) {
        use(*ds[0]);
        fn(*ds[0]);
        return;
    }
    struct Outcome {
        int code = 0;
        char msg[256] = {0};
    };
    std::vector<Outcome> res(ds.size());
    auto body = [&](size_t i) {
        try {
            use(*ds[i]);
            fn(*ds[i]);
        } catch (const Exc& e) {
            res[i].code = e.code;
            memcpy(res[i].msg, last_error().msg, sizeof res[i].msg);
        } catch (const std::bad_alloc&) {
            res[i].code = MB200_ENOMEM;
            snprintf(res[i].msg, sizeof res[i].msg, "host allocation failed");
        } catch (...) {
            res[i].code = MB200_ECUDA;
            snprintf(res[i].msg, sizeof res[i].msg, "unexpected failure on device slot %d", (int)i);
        }
    };
    std::vector<std::thread> th;
    size_t started = 1;
    try {
        for (size_t i = 1; i < ds.size(); ++i) {
            th.emplace_back(body, i);
            started = i + 1;
        }
    } catch (...) {  // std::system_error: could not spawn; the remaining devices run on this thread
    }
    body(0);
    for (size_t i = started; i < ds.size(); ++i) body(i);
    for (auto& t : th) t.join();
    use(*g.devs[0]);
    for (auto& r : res)
        if (r.code) {
            last_error().code = r.code;
            memcpy(last_error().msg, r.msg, sizeof r.msg);
            throw Exc{r.code};
        }
}
static std::vector<Device*> all_devices() {
    std::vector<Device*> v;
    for (auto& d : g.devs) v.push_back(d.get());
    return v;
}

static void destroy_ctxs(Device& d) {
#ifndef MB200_EMU
    for (auto& c : d.ctxs)
        if (c.have_stream) {
            cudaStreamSynchronize(c.stream);
            cudaStreamDestroy(c.stream);
            cudaEventDestroy(c.ev_inputs);
            cudaEventDestroy(c.ev_tail);
            for (int i = 0; i < 3; ++i) {
                cudaStreamSynchronize(c.side[i]);
                cudaStreamDestroy(c.side[i]);
                cudaEventDestroy(c.ev_side[i]);
            }
        }
#endif
    d.ctxs.clear();
}
static void set_ctx_count(Device& d, size_t n) {
    destroy_ctxs(d);
    d.ctxs.resize(n);
#ifndef MB200_EMU
    for (auto& c : d.ctxs) {
        MB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        MB_CUDA(cudaEventCreateWithFlags(&c.ev_inputs, cudaEventDisableTiming));
        MB_CUDA(cudaEventCreateWithFlags(&c.ev_tail, cudaEventDisableTiming));
        for (int i = 0; i < 3; ++i) {
            MB_CUDA(cudaStreamCreateWithFlags(&c.side[i], cudaStreamNonBlocking));
            MB_CUDA(cudaEventCreateWithFlags(&c.ev_side[i], cudaEventDisableTiming));
        }
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

// One MSM over device-resident bases on device `d`; scalars from host memory (pinned memory makes
// the uploads asynchronous).  out_dev: one XYZZ on `d`, complete when this returns.
// The scalars are cut into slabs of MSM_SLAB: slab k+1 crosses PCIe while slab k is sorted and
// accumulated into one shared set of buckets; the bucket reduction runs once, after the last slab.
template <class F>
static void msm_on_device_bases(Device& d, const Affine<F>* bases, const uint8_t* scalars_host, size_t n, XYZZ<F>* out_dev) {
    if (n == 0) {
        dev_memset(out_dev, 0, sizeof(XYZZ<F>), d.main);
        stream_sync(d.main);
        return;
    }
    if (n >= (1ull << 31)) fail(MB200_EINVAL, "MSM of %s%ld bases is too large", "", (long)n);
    // up to eight slabs once they are worth their fixed cost (every slab walks all bucket sets: >= 2^21
    // scalars each, measured -- 2^18 made the 2^20 / 2^22 witness-like MSMs 30 % slower): only the first
    // slab's upload is exposed, the rest crosses PCIe under the previous slab's kernels
    const size_t slab = std::min(g.msm_slab, std::max<size_t>((size_t)1 << 21, (n + 7) / 8));
    const size_t nslab = (n + slab - 1) / slab;
    const size_t per = (n + nslab - 1) / nslab;  // equal slabs, each <= msm_slab
    for (int b = 0; b < (nslab > 1 ? 2 : 1); ++b) d.slab[b].ensure(per * 32);
    d.msm_flag.ensure(4);
    dev_memset(d.msm_flag.p, 0, 4, d.main);
    const uint32_t c = standalone_window(n);
    uint32_t n_ones = (uint32_t)std::min<size_t>(256, std::max<size_t>(1, per / 256));
    auto upload = [&](size_t k) {
        const size_t lo = k * per, cnt = std::min(per, n - lo);
        const int b = (int)(k & 1);
#ifndef MB200_EMU
        if (k >= 2) MB_CUDA(cudaStreamWaitEvent(d.copy, d.ev_free[b], 0));   // slab k-2 has been consumed
        else if (k == 0) {
            MB_CUDA(cudaEventRecord(d.ev_free[0], d.main));                  // order after earlier work on main
            MB_CUDA(cudaStreamWaitEvent(d.copy, d.ev_free[0], 0));
        }
#endif
        cudaStream_t cs = d.copy;
        copy_h2d(d.slab[b].p, scalars_host + lo * 32, cnt * 32, cs);
#ifndef MB200_EMU
        MB_CUDA(cudaEventRecord(d.ev_up[b], cs));
#endif
    };
    upload(0);
    MsmClass last;
    for (size_t k = 0; k < nslab; ++k) {
        const size_t lo = k * per, cnt = std::min(per, n - lo);
        const int b = (int)(k & 1);
        if (k + 1 < nslab) upload(k + 1);
#ifndef MB200_EMU
        MB_CUDA(cudaStreamWaitEvent(d.main, d.ev_up[b], 0));
#endif
        ValidateArgs va{cnt, d.slab[b].as<Fr>(), cnt, cnt, d.msm_flag.as<uint32_t>()};
        launch_validate_scalars(va, d.main);
        // every slab sorts and accumulates into the SAME bucket sets (same c, same layout); the bucket
        // reduction and the Horner step over the windows run once, after the last slab
        last = msm_make_class(bases + lo, nullptr, (uint32_t)cnt, c, false, n_ones);
        msm_accumulate_buckets<F>(last, 1, d.slab[b].as<uint32_t>(), cnt, d.msm, d.main, k > 0);
#ifndef MB200_EMU
        MB_CUDA(cudaEventRecord(d.ev_free[b], d.main));
#endif
    }
    msm_reduce_buckets<F>(last, 1, out_dev, d.msm, d.main);
    uint32_t bad = 0;
    copy_d2h(&bad, d.msm_flag.p, 4, d.main);
    stream_sync(d.main);
    if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
}

static NttCache& ntt_cache(Device& d, unsigned log_n) {
    auto it = d.ntt.find(log_n);
    if (it != d.ntt.end()) return *it->second;
    std::unique_ptr<NttCache> c(new NttCache());
    c->dom.build(log_n, d.main);
    size_t n = (size_t)1 << log_n;
    c->gpow.alloc(n * sizeof(Fr));
    PowArgs pa;
    pa.nthreads = n;
    pa.out = c->gpow.as<Fr>();
    pa.base = fr_from_u64_host(7);
    pa.scale = Fr::one();
    launch_fr_powers(pa, d.main);
    NttCache* raw = c.release();
    d.ntt[log_n] = raw;
    return *raw;
}

static void check_scalars_dev(const Fr* base, size_t per_row, size_t row_stride, size_t nrows, uint32_t* flag,
                              cudaStream_t s) {
    ValidateArgs va{per_row * nrows, base, per_row, row_stride, flag};
    launch_validate_scalars(va, s);
}

// Buffers a batch slice needs for its lifetime.  They are recycled across batches: cudaMalloc /
// cudaFree / cudaFreeHost synchronise the whole device, which would serialise batches that are
// meant to overlap (submit k+1 while k is still running, wait k while k+1 runs).
struct TicketRes {
    uint8_t* staged = nullptr;     // pinned: proofs land here
    size_t staged_cap = 0;
    uint32_t* verdicts = nullptr;  // pinned: per-proof result of the self-check (option "verify")
    size_t verdicts_cap = 0;
    DevBuf v_a, v_b, v_c, v_in, v_ok;  // self-check inputs / verdicts of this slice (device)
    DevBuf flag;                   // set by the canonical-scalar checks of this slice
    ~TicketRes() {
        host_free_pinned(staged);
        host_free_pinned(verdicts);
    }
};
static TicketRes* res_acquire(Device& d) {
    if (d.res_free.empty()) return new TicketRes();
    TicketRes* r = d.res_free.back();
    d.res_free.pop_back();
    return r;
}

// The part of a submitted batch that runs on one device: proofs [first, first + count).
struct Slice {
    Device* dev = nullptr;
    size_t first = 0, count = 0;
    TicketRes* res = nullptr;
    uint8_t* staged = nullptr;     // = res->staged
    uint32_t* verdicts = nullptr;  // = res->verdicts when the self-check is on
    uint32_t* flag = nullptr;      // = res->flag
#ifndef MB200_EMU
    cudaEvent_t ev_v[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev0 = nullptr;
    std::vector<cudaEvent_t> ev1;
#endif
};
// One submitted batch: everything needed to finish it later.
struct Ticket {
    std::vector<Slice> slices;
    uint8_t* out = nullptr;  // caller's buffer
    size_t n_proofs = 0;
    int verify = 0;          // option "verify" as it was at submit time
};
static std::map<uint64_t, Ticket*> g_tickets;
static uint64_t g_next_ticket = 1;

static void slice_release(Slice& sl) {
#ifndef MB200_EMU
    if (sl.ev0) cudaEventDestroy(sl.ev0);
    for (auto& e : sl.ev_v)
        if (e) cudaEventDestroy(e);
    for (auto& e : sl.ev1)
        if (e) cudaEventDestroy(e);
    sl.ev0 = nullptr;
    sl.ev1.clear();
    for (auto& e : sl.ev_v) e = nullptr;
#endif
    if (sl.res) sl.dev->res_free.push_back(sl.res);
    sl.res = nullptr;
}
static void ticket_destroy(Ticket* t) {
    if (!t) return;
    for (auto& sl : t->slices) slice_release(sl);
    delete t;
}

// How a batch is cut over the devices: whole chunks, contiguous, as even as the chunk size
// allows; a batch of fewer chunks than devices uses that many devices, starting at a rotating
// slot so that small batches submitted back to back land on different GPUs.
static std::vector<Slice> plan_slices(size_t n_proofs) {
    const size_t nd = g.devs.size(), chunk = g.chunk;
    const size_t units = (n_proofs + chunk - 1) / chunk;
    const size_t used = std::min(nd, units);
    std::vector<Slice> out;
    size_t first = 0;
    for (size_t k = 0; k < used; ++k) {
        size_t u = units / used + (k < units % used ? 1 : 0);
        size_t count = std::min(u * chunk, n_proofs - first);
        Slice sl;
        sl.dev = g.devs[(g.next_dev + k) % nd].get();
        sl.first = first;
        sl.count = count;
        out.push_back(sl);
        first += count;
    }
    g.next_dev = (g.next_dev + used) % nd;
    return out;
}

// Enqueue one slice on its device's chunk contexts; no host synchronisation (with pinned inputs).
static void slice_submit(const Params& P, Slice& sl, int verify, size_t rows, const ProveInputs& in) {
    Device& d = *sl.dev;
    TicketRes* R = sl.res = res_acquire(d);
    if (R->staged_cap < sl.count * 192) {
        host_free_pinned(R->staged);
        R->staged = nullptr;
        R->staged_cap = 0;
        R->staged = (uint8_t*)host_alloc_pinned(sl.count * 192);
        R->staged_cap = sl.count * 192;
    }
    sl.staged = R->staged;
    VerifySink sink;
    if (verify) {
        if (R->verdicts_cap < sl.count * 4) {
            host_free_pinned(R->verdicts);
            R->verdicts = nullptr;
            R->verdicts_cap = 0;
            R->verdicts = (uint32_t*)host_alloc_pinned(sl.count * 4);
            R->verdicts_cap = sl.count * 4;
        }
        sl.verdicts = R->verdicts;
        R->v_a.ensure(sl.count * sizeof(G1Affine));
        R->v_b.ensure(sl.count * sizeof(G2Affine));
        R->v_c.ensure(sl.count * sizeof(G1Affine));
        R->v_in.ensure(sl.count * (size_t)P.n_inputs * 32);
        R->v_ok.ensure(sl.count * 4);
        sink.a = R->v_a.as<G1Affine>();
        sink.b = R->v_b.as<G2Affine>();
        sink.c = R->v_c.as<G1Affine>();
        sink.inputs = R->v_in.as<uint32_t>();
        sink.ok_dev = R->v_ok.as<uint32_t>();
        sink.ok_host = sl.verdicts;
    }
    R->flag.ensure(4);
    sl.flag = R->flag.as<uint32_t>();
    dev_memset(sl.flag, 0, 4, d.main);
#ifndef MB200_EMU
    // every chunk stream starts after `ev0` (and so after the flag reset)
    sl.ev1.assign(d.ctxs.size(), nullptr);
    MB_CUDA(cudaEventCreate(&sl.ev0));
    for (auto& e : sl.ev1) MB_CUDA(cudaEventCreate(&e));
    MB_CUDA(cudaEventRecord(sl.ev0, d.main));
    for (auto& x : d.ctxs) MB_CUDA(cudaStreamWaitEvent(x.stream, sl.ev0, 0));
#endif
    // the slice's view of the caller's buffers; chunk offsets below are relative to the slice
    ProveInputs sin = in;
    const size_t rb = rows * 32;
    if (in.a) {
        sin.a = in.a + sl.first * rb;
        sin.b = in.b + sl.first * rb;
        sin.c = in.c + sl.first * rb;
    }
    sin.inputs = in.inputs + sl.first * (size_t)P.n_inputs * 32;
    sin.aux = in.aux + sl.first * (size_t)P.n_aux * 32;
    sin.r = in.r + sl.first * 32;
    sin.s = in.s + sl.first * 32;
    for (size_t first = 0; first < sl.count; first += g.chunk) {
        uint32_t count = (uint32_t)std::min<size_t>(g.chunk, sl.count - first);
        ProveCtx& x = d.ctxs[d.next_ctx++ % d.ctxs.size()];
#ifndef MB200_EMU
        sink.stream = d.vstream[d.next_v++ % 4];
#endif
        prove_chunk(P, x, sin, first, count, rows, sl.staged, sink);
        // canonical-scalar check on what was just staged (abc and aux..s of the pool)
        check_scalars_dev(x.abc.as<Fr>(), (size_t)count * 3 * rows, 0, 1, sl.flag, x.stream);
        check_scalars_dev(x.pool.as<Fr>() + P.idx_aux, P.idx_one - P.idx_aux, P.pool_stride, count, sl.flag, x.stream);
    }
#ifndef MB200_EMU
    for (size_t i = 0; i < d.ctxs.size(); ++i) MB_CUDA(cudaEventRecord(sl.ev1[i], d.ctxs[i].stream));
    if (verify)
        for (int i = 0; i < 4; ++i) {
            MB_CUDA(cudaEventCreateWithFlags(&sl.ev_v[i], cudaEventDisableTiming));
            MB_CUDA(cudaEventRecord(sl.ev_v[i], d.vstream[i]));
        }
#endif
}

// Enqueue a whole batch: one slice per device, each enqueued by its own host thread.
static uint64_t prove_submit(const std::vector<Params*>& reps, size_t n_proofs, size_t rows, const ProveInputs& in,
                             uint8_t* proofs_out) {
    require_init();
    const Params& P0 = *reps[0];
    if (!in.inputs || !in.aux || !in.r || !in.s || !proofs_out) fail(MB200_EINVAL, "null buffer%s", "");
    if (!in.a && !in.b && !in.c) {  // witness-only: rows come from the bound circuit
        if (!P0.r1cs.bound) fail(MB200_EINVAL, "no circuit bound to these parameters%s", "");
        if (rows != (size_t)P0.r1cs.ncons + P0.n_inputs) fail(MB200_EINVAL, "rows do not match the bound circuit%s", "");
    } else if (!in.a || !in.b || !in.c) {
        fail(MB200_EINVAL, "null buffer%s", "");
    }
    size_t m = 1;
    while (m < rows) m <<= 1;
    if (rows == 0 || m != P0.m)
        fail(MB200_EINVAL, "rows does not match the key's domain%s (key m = %ld)", "", (long)P0.m);
    Ticket* t = new Ticket();
    try {
        t->n_proofs = n_proofs;
        t->out = proofs_out;
        t->verify = g.verify;
        if (in.on_device) {  // the inputs are on the caller's current device: the whole batch runs there
            Slice sl;
            sl.dev = g.devs[0].get();
#ifndef MB200_EMU
            int cur = 0;
            MB_CUDA(cudaGetDevice(&cur));
            for (auto& d : g.devs)
                if (d->id == cur) sl.dev = d.get();
#endif
            sl.first = 0;
            sl.count = n_proofs;
            t->slices.push_back(sl);
        } else {
            t->slices = plan_slices(n_proofs);
        }
        std::vector<Device*> ds;
        for (auto& sl : t->slices) ds.push_back(sl.dev);
        const int verify = t->verify;
        for_devices(ds, [&](Device& d) {
            for (auto& sl : t->slices)
                if (sl.dev == &d) slice_submit(*reps[d.slot], sl, verify, rows, in);
        });
    } catch (...) {
        // nothing of this batch may still be running when its buffers go back to the pool
#ifndef MB200_EMU
        for (auto& sl : t->slices) {
            cudaSetDevice(sl.dev->id);
            for (auto& x : sl.dev->ctxs) cudaStreamSynchronize(x.stream);
            for (auto& v : sl.dev->vstream) cudaStreamSynchronize(v);
        }
        cudaSetDevice(g.devs[0]->id);
#endif
        ticket_destroy(t);
        throw;
    }
    uint64_t id = g_next_ticket++;
    g_tickets[id] = t;
    return id;
}

// Wait for a submitted batch, check its flags, hand the proofs to the caller.
static void prove_wait(uint64_t id) {
    auto it = g_tickets.find(id);
    if (it == g_tickets.end()) fail(MB200_EINVAL, "unknown ticket%s (%ld)", "", (long)id);
    Ticket* t = it->second;
    g_tickets.erase(it);
    try {
        g.last_batch_ms = 0;
        uint32_t bad = 0;
        for (auto& sl : t->slices) {
            Device& d = *sl.dev;
            use(d);
#ifndef MB200_EMU
            for (auto& e : sl.ev1) {
                MB_CUDA(cudaEventSynchronize(e));
                float ms = 0;
                MB_CUDA(cudaEventElapsedTime(&ms, sl.ev0, e));
                if (ms > g.last_batch_ms) g.last_batch_ms = ms;
            }
            for (auto& e : sl.ev_v)
                if (e) MB_CUDA(cudaEventSynchronize(e));
#endif
            uint32_t b = 0;
            copy_d2h(&b, sl.flag, 4, d.main);
            stream_sync(d.main);
            bad |= b;
        }
        use(*g.devs[0]);
        if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
        for (auto& sl : t->slices) memcpy(t->out + sl.first * 192, sl.staged, sl.count * 192);
        for (auto& sl : t->slices)
            if (sl.verdicts)
                for (size_t i = 0; i < sl.count; ++i) {
                    g.verified++;
                    if (!sl.verdicts[i]) {
                        g.verify_failed++;
                        if (t->verify == 1)
                            fail(MB200_EVERIFY, "proof %s%ld does not satisfy the verification equation", "",
                                 (long)(sl.first + i));
                    }
                }
    } catch (...) {
#ifndef MB200_EMU
        // a failed wait must not hand buffers back while later slices are still running
        for (auto& sl : t->slices) {
            cudaSetDevice(sl.dev->id);
            for (auto& e : sl.ev1)
                if (e) cudaEventSynchronize(e);
            for (auto& e : sl.ev_v)
                if (e) cudaEventSynchronize(e);
        }
        cudaSetDevice(g.devs[0]->id);
#endif
        ticket_destroy(t);
        throw;
    }
    ticket_destroy(t);
}

// CSR matrices of a recorded circuit -> device, columns rewritten to scalar-pool
// indices, coefficients interned into a dictionary (Montgomery form; 0 -> +1, 1 -> -1).
static void r1cs_upload(Device& d, R1csDev& R, const mb200_circuit& c, size_t idx_aux, size_t idx_inputs) {
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
        copy_h2d(R.rowptr[k].p, M.rowptr.data(), M.rowptr.size() * 4, d.main);
        copy_h2d(R.col[k].p, col.data(), col.size() * 4, d.main);
        copy_h2d(R.cidx[k].p, cidx.data(), cidx.size() * 4, d.main);
        stream_sync(d.main);  // the staging vectors die at the end of this iteration
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
        copy_h2d(R.order.p, order.data(), order.size() * 4, d.main);
        stream_sync(d.main);
    }
    // mbh::Fr (4 x u64 Montgomery, R = 2^256) has the same memory image as the device Fr (8 x u32)
    R.dict.alloc(dict_vals.size() * 32);
    copy_h2d(R.dict.p, dict_vals.data(), dict_vals.size() * 32, d.main);
    stream_sync(d.main);
    R.ncons = c.n_constraints;
    R.n_inputs = c.n_inputs;
    R.n_aux = c.n_aux;
    R.bound = true;
}

static int prove_impl(const std::vector<Params*>& reps, size_t n_proofs, size_t rows, const ProveInputs& in,
                      uint8_t* proofs_out) {
    require_init();
    if (n_proofs == 0) return MB200_OK;
    prove_wait(prove_submit(reps, n_proofs, rows, in, proofs_out));
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
    std::vector<Params*> rep;  // one replica per opened device, indexed by Device::slot
    const Params& p0() const { return *rep[0]; }
};
static void params_destroy(mb200_params* p) {
    if (!p) return;
    for (size_t i = 0; i < p->rep.size(); ++i)
        if (p->rep[i]) {
#ifndef MB200_EMU
            if (i < g.devs.size()) {
                cudaSetDevice(g.devs[i]->id);
                cudaDeviceSynchronize();
            }
#endif
            delete p->rep[i];
        }
#ifndef MB200_EMU
    if (!g.devs.empty()) cudaSetDevice(g.devs[0]->id);
#endif
    delete p;
}
static const std::vector<Params*>& replicas(const mb200_params* p) {
    if (!p || p->rep.empty() || !p->rep[0]) fail(MB200_EINVAL, "null parameters%s", "");
    if (p->rep.size() != g.devs.size()) fail(MB200_ESTATE, "parameters were loaded under a different mb200_init%s", "");
    return p->rep;
}

extern "C" {

int mb200_init(const int* device_ids, int n_devices) {
    MB_API_BEGIN
    if (g.inited) return MB200_OK;
    std::vector<int> ids;
#ifndef MB200_EMU
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        fail(MB200_ECUDA, "no CUDA device: %s (this library has no CPU path)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device_ids && n_devices > 0) {
        ids.assign(device_ids, device_ids + n_devices);
    } else if (n_devices < 0) {  // every visible device
        for (int i = 0; i < count; ++i) ids.push_back(i);
    } else {
        int dev = 0;
        MB_CUDA(cudaGetDevice(&dev));
        ids.push_back(dev);
    }
    if (ids.size() > 64) fail(MB200_EINVAL, "too many devices%s (%ld)", "", (long)ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {
        if (ids[i] < 0 || ids[i] >= count) fail(MB200_EINVAL, "device id out of range%s (%ld)", "", (long)ids[i]);
        for (size_t j = 0; j < i; ++j)
            if (ids[j] == ids[i]) fail(MB200_EINVAL, "device id listed twice%s (%ld)", "", (long)ids[i]);
        cudaDeviceProp prop;
        MB_CUDA(cudaGetDeviceProperties(&prop, ids[i]));
        if (prop.major < 10) fail(MB200_ECUDA, "device %s is not sm_100 (compute capability %ld.x)", prop.name, (long)prop.major);
    }
#else
    // host emulation (tests only): "devices" are separate bookkeeping over the same host memory
    int n = (device_ids && n_devices > 0) ? n_devices : 1;
    if (n > 64) fail(MB200_EINVAL, "too many devices%s (%ld)", "", (long)n);
    for (int i = 0; i < n; ++i) ids.push_back(i);
#endif
    g.chunk = env_u32("MB200_CHUNK", 64);
    g.n_streams = env_u32("MB200_STREAMS", 3);
    if (g.n_streams < 1 || g.n_streams > 8) g.n_streams = 3;
    try {
        for (size_t i = 0; i < ids.size(); ++i) {
            std::unique_ptr<Device> d(new Device());
            d->id = ids[i];
            d->slot = (int)i;
            g.devs.push_back(std::move(d));
        }
        for (auto& dp : g.devs) {
            Device& d = *dp;
            use(d);
#ifndef MB200_EMU
            MB_CUDA(cudaStreamCreateWithFlags(&d.main, cudaStreamNonBlocking));
            MB_CUDA(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
            for (int b = 0; b < 2; ++b) {
                MB_CUDA(cudaEventCreateWithFlags(&d.ev_up[b], cudaEventDisableTiming));
                MB_CUDA(cudaEventCreateWithFlags(&d.ev_free[b], cudaEventDisableTiming));
            }
            for (auto& v : d.vstream) MB_CUDA(cudaStreamCreateWithFlags(&v, cudaStreamNonBlocking));
#endif
            set_ctx_count(d, g.n_streams);
        }
#ifndef MB200_EMU
        // K6 (partials of a base-split MSM travel GPU -> GPU): let slot 0 be written by its peers
        for (size_t i = 1; i < g.devs.size(); ++i) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g.devs[i]->id, g.devs[0]->id);
            if (can) {
                cudaSetDevice(g.devs[i]->id);
                cudaError_t pe = cudaDeviceEnablePeerAccess(g.devs[0]->id, 0);
                if (pe != cudaSuccess) cudaGetLastError();  // already enabled, or unsupported: cudaMemcpyPeerAsync still works
            }
        }
#endif
        use(*g.devs[0]);
    } catch (...) {
        g.devs.clear();
        throw;
    }
    g.inited = true;
    MB_API_END
}

int mb200_device_count(void) {
    std::lock_guard<std::mutex> lk(g.mu);
    return g.inited ? (int)g.devs.size() : 0;
}

int mb200_shutdown(void) {
    MB_API_BEGIN
    if (!g.inited) return MB200_OK;
    for (auto& kv : g_tickets) ticket_destroy(kv.second);  // batches never waited for
    g_tickets.clear();
    for (auto& dp : g.devs) {
        Device& d = *dp;
#ifndef MB200_EMU
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
#endif
        for (auto& kv : d.ntt) delete kv.second;
        d.ntt.clear();
        destroy_ctxs(d);
        for (auto* r : d.res_free) delete r;
        d.res_free.clear();
        d.msm = MsmScratch();
        d.peer_parts.release();
        for (auto& b : d.slab) b.release();
        d.msm_flag.release();
        d.msm_enc.release();
#ifndef MB200_EMU
        cudaStreamDestroy(d.copy);
        for (int b = 0; b < 2; ++b) {
            cudaEventDestroy(d.ev_up[b]);
            cudaEventDestroy(d.ev_free[b]);
        }
        cudaStreamDestroy(d.main);
        for (auto& v : d.vstream) cudaStreamDestroy(v);
#endif
    }
    g.devs.clear();
    g.inited = false;
    MB_API_END
}

static mb200_params* load_replicated(const uint8_t* bytes, size_t len, const uint8_t* a_aux_density,
                                     const uint8_t* b_input_density, const uint8_t* b_aux_density);

int mb200_params_load(const uint8_t* bytes, size_t len, const uint8_t* a_aux_density, const uint8_t* b_input_density,
                      const uint8_t* b_aux_density, mb200_params** out) {
    MB_API_BEGIN
    require_init();
    if (!out) fail(MB200_EINVAL, "null out pointer%s", "");
    *out = nullptr;
    *out = load_replicated(bytes, len, a_aux_density, b_input_density, b_aux_density);
    MB_API_END
}

// load_parameters / parse_parameters of the reference (masp_proofs/src/lib.rs:278-325, 343-388) for ONE
// file: (1) the size must equal expected_bytes before anything is hashed; (2) Parameters::read(.., false)
// consumes the key; (3) the WHOLE stream, transcript included, is BLAKE2b-512 hashed and the hex digest
// compared.  The reference panics where this returns MB200_EPARAMS / MB200_EIO / MB200_EPARSE.
static void check_stream(const uint8_t* bytes, size_t len, uint64_t expected_bytes, const char* expected_hex,
                         const char* what) {
    if (expected_bytes && len != expected_bytes)
        fail(MB200_EPARAMS, "%s: parameter file size is not correct (%ld bytes)", what, (long)len);
    if (expected_hex && *expected_hex) {
        if (strlen(expected_hex) != 128) fail(MB200_EINVAL, "expected BLAKE2b-512 digest must be 128 hex digits%s", "");
        char got[129];
        mbh::blake2b512_hex(bytes, len, got);
        for (int i = 0; i < 128; ++i) {
            char e = expected_hex[i];
            if (e >= 'A' && e <= 'F') e = (char)(e - 'A' + 'a');
            if (e != got[i]) {
                got[16] = 0;
                fail(MB200_EPARAMS, "parameter file is not correct: BLAKE2b-512 = %s... over %ld bytes", got, (long)len);
            }
        }
    }
}
static mb200_params* load_replicated(const uint8_t* bytes, size_t len, const uint8_t* a_aux_density,
                                     const uint8_t* b_input_density, const uint8_t* b_aux_density) {
    std::unique_ptr<mb200_params, void (*)(mb200_params*)> p(new mb200_params(), params_destroy);
    p->rep.assign(g.devs.size(), nullptr);
    // replicate: every device ingests the bytes and expands its own window tables, in parallel
    for_devices(all_devices(), [&](Device& d) {
        p->rep[d.slot] = params_load(bytes, len, a_aux_density, b_input_density, b_aux_density, d.main);
    });
    return p.release();
}

int mb200_params_load_verified(const uint8_t* bytes, size_t len, uint64_t expected_bytes, const char* expected_blake2b_hex,
                               const uint8_t* a_aux_density, const uint8_t* b_input_density,
                               const uint8_t* b_aux_density, mb200_params** out) {
    MB_API_BEGIN
    require_init();
    if (!out || !bytes) fail(MB200_EINVAL, "null argument%s", "");
    *out = nullptr;
    if (expected_bytes && len != expected_bytes) check_stream(bytes, len, expected_bytes, nullptr, "stream");
    mb200_params* p = load_replicated(bytes, len, a_aux_density, b_input_density, b_aux_density);
    try {
        check_stream(bytes, len, expected_bytes, expected_blake2b_hex, "stream");
    } catch (...) {
        params_destroy(p);
        throw;
    }
    *out = p;
    MB_API_END
}

int mb200_params_load_file(const char* path, uint64_t expected_bytes, const char* expected_blake2b_hex,
                           const uint8_t* a_aux_density, const uint8_t* b_input_density, const uint8_t* b_aux_density,
                           mb200_params** out) {
    MB_API_BEGIN
    require_init();
    if (!out || !path) fail(MB200_EINVAL, "null argument%s", "");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) fail(MB200_EIO, "couldn't load parameters file %s", path);
    std::vector<uint8_t> buf;
    try {
        // verify_file_size first (lib.rs:284-311): nothing is read or hashed when the size is off
        if (fseek(f, 0, SEEK_END) != 0) fail(MB200_EIO, "cannot seek in %s", path);
        long size = ftell(f);
        if (size < 0) fail(MB200_EIO, "cannot size %s", path);
        if (expected_bytes && (uint64_t)size != expected_bytes)
            fail(MB200_EPARAMS, "%s: parameter file size is not correct (%ld bytes)", path, size);
        rewind(f);
        buf.resize((size_t)size);
        if (size && fread(buf.data(), 1, (size_t)size, f) != (size_t)size) fail(MB200_EIO, "short read from %s", path);
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
    mb200_params* p = load_replicated(buf.data(), buf.size(), a_aux_density, b_input_density, b_aux_density);
    try {
        check_stream(buf.data(), buf.size(), expected_bytes, expected_blake2b_hex, path);
    } catch (...) {
        params_destroy(p);
        throw;
    }
    *out = p;
    MB_API_END
}

int mb200_masp_params_spec(int kind, uint64_t* expected_bytes, char blake2b_hex[129], const char** file_name) {
    // masp_proofs/src/lib.rs:61-76
    static const struct {
        const char* name;
        const char* hash;
        uint64_t bytes;
    } spec[3] = {
        {"masp-spend.params", "196e7c717f25e16653431559ce2c8816e750a4490f98696e3c031efca37e25e0647182b7b013660806db11eb2b1e365fb2d6a0f24dbbd9a4a8314fef10a7cba2", 49848572ull},
        {"masp-output.params", "eafc3b1746cccc8b9eed2b69395692c5892f6aca83552a07dceb2dcbaa64dcd0e22434260b3aa3b049b633a08b008988cbe0d31effc77e2bc09bfab690a23724", 16398620ull},
        {"masp-convert.params", "dc4aaf3c3ce056ab448b6c4a7f43c1d68502c2902ea89ab8769b1524a2e8ace9a5369621a73ee1daa52aec826907a19974a37874391cf8f11bbe0b0420de1ab7", 22570940ull},
    };
    if (kind < 0 || kind > 2) return MB200_EINVAL;
    if (expected_bytes) *expected_bytes = spec[kind].bytes;
    if (blake2b_hex) memcpy(blake2b_hex, spec[kind].hash, 129);
    if (file_name) *file_name = spec[kind].name;
    return MB200_OK;
}

int mb200_blake2b512(const uint8_t* bytes, size_t len, uint8_t out[64]) {
    if ((len && !bytes) || !out) return MB200_EINVAL;
    mbh::Blake2b b;
    b.update(bytes, len);
    b.finish(out);
    return MB200_OK;
}

int mb200_params_info(const mb200_params* p, uint64_t info[10]) {
    MB_API_BEGIN
    if (!p || p->rep.empty() || !p->rep[0] || !info) fail(MB200_EINVAL, "null argument%s", "");
    const Params& P = p->p0();
    info[0] = P.n_inputs; info[1] = P.n_aux; info[2] = P.h_len; info[3] = P.a_len; info[4] = P.b_len;
    info[5] = P.m; info[6] = P.consumed; info[7] = P.table_bytes; info[8] = P.k_hl.c; info[9] = P.k_a.c;
    MB_API_END
}

void mb200_params_free(mb200_params* p) {
    std::lock_guard<std::mutex> lk(g.mu);
    params_destroy(p);
}

size_t mb200_params_synth_size(uint32_t n_inputs, uint32_t h_len, uint32_t l_len, uint32_t a_len, uint32_t b_len) {
    return params_synth_size(n_inputs, h_len, l_len, a_len, b_len);
}

int mb200_params_synthesize(uint64_t seed, uint32_t n_inputs, uint32_t h_len, uint32_t l_len, uint32_t a_len,
                            uint32_t b_len, uint8_t* out, size_t out_len) {
    MB_API_BEGIN
    Device& d = dev0();
    size_t need = params_synth_size(n_inputs, h_len, l_len, a_len, b_len);
    if (!out || out_len < need) fail(MB200_EINVAL, "output buffer too small%s (need %ld bytes)", "", (long)need);
    DevBuf buf(need);
    params_synthesize(seed, n_inputs, h_len, l_len, a_len, b_len, buf.as<uint8_t>(), d.main);
    copy_d2h(out, buf.p, need, d.main);
    stream_sync(d.main);
    MB_API_END
}

int mb200_synth_points(uint64_t seed, uint32_t stream, uint64_t start, size_t n, int group, uint8_t* out) {
    MB_API_BEGIN
    Device& d = dev0();
    if (!out || (group != 1 && group != 2)) fail(MB200_EINVAL, "bad argument%s", "");
    size_t bytes = n * (group == 1 ? 96 : 192);
    DevBuf buf(bytes);
    SynthArgs a;
    a.nthreads = n;
    a.key = stream_key(seed, stream);
    a.start = start;
    a.out = buf.as<uint8_t>();
    a.g1 = g1_generator_host();
    a.g2 = g2_generator_host();
    if (group == 1) launch_synth_g1(a, d.main);
    else launch_synth_g2(a, d.main);
    copy_d2h(out, buf.p, bytes, d.main);
    stream_sync(d.main);
    MB_API_END
}

int mb200_prove_batch(const mb200_params* p, size_t n_proofs, size_t rows, const uint8_t* a_evals,
                      const uint8_t* b_evals, const uint8_t* c_evals, const uint8_t* inputs, const uint8_t* aux,
                      const uint8_t* r, const uint8_t* s, uint8_t* proofs_out) {
    MB_API_BEGIN
    require_init();
    ProveInputs in{a_evals, b_evals, c_evals, inputs, aux, r, s, false};
    return prove_impl(replicas(p), n_proofs, rows, in, proofs_out);
    MB_API_END
}

int mb200_prove_batch_device(const mb200_params* p, size_t n_proofs, size_t rows, const void* a_evals,
                             const void* b_evals, const void* c_evals, const void* inputs, const void* aux,
                             const void* r, const void* s, uint8_t* proofs_out) {
    MB_API_BEGIN
    require_init();
    ProveInputs in{(const uint8_t*)a_evals, (const uint8_t*)b_evals, (const uint8_t*)c_evals, (const uint8_t*)inputs,
                   (const uint8_t*)aux, (const uint8_t*)r, (const uint8_t*)s, true};
    return prove_impl(replicas(p), n_proofs, rows, in, proofs_out);
    MB_API_END
}

int mb200_prove_submit(const mb200_params* p, size_t n_proofs, size_t rows, const void* a_evals, const void* b_evals,
                       const void* c_evals, const void* inputs, const void* aux, const void* r, const void* s,
                       int on_device, uint8_t* proofs_out, uint64_t* ticket) {
    MB_API_BEGIN
    require_init();
    if (!ticket || n_proofs == 0) fail(MB200_EINVAL, "bad argument%s", "");
    ProveInputs in{(const uint8_t*)a_evals, (const uint8_t*)b_evals, (const uint8_t*)c_evals, (const uint8_t*)inputs,
                   (const uint8_t*)aux, (const uint8_t*)r, (const uint8_t*)s, on_device != 0};
    *ticket = prove_submit(replicas(p), n_proofs, rows, in, proofs_out);
    MB_API_END
}

int mb200_params_bind_circuit(mb200_params* p, const mb200_circuit* c) {
    MB_API_BEGIN
    require_init();
    if (!c) fail(MB200_EINVAL, "null argument%s", "");
    const std::vector<Params*>& reps = replicas(p);
    const Params& P = *reps[0];
    if (P.n_inputs != c->n_inputs || P.n_aux != c->n_aux)
        fail(MB200_EINVAL, "circuit and key disagree on the variable counts%s", "");
    if (P.a_len != c->n_inputs + c->a_aux_ones || P.b_len != c->b_input_ones + c->b_aux_ones)
        fail(MB200_EINVAL, "circuit densities do not match the key's query lengths%s", "");
    // equal counts are not enough: the key's base <-> variable maps were built from the bitmaps it
    // was loaded with, and a circuit with other positions would give silently wrong proofs
    if (!density_equal(P.a_aux_density, c->a_aux_density, c->n_aux) ||
        !density_equal(P.b_input_density, c->b_input_density, c->n_inputs) ||
        !density_equal(P.b_aux_density, c->b_aux_density, c->n_aux))
        fail(MB200_EINVAL, "the key was loaded with density bitmaps that are not this circuit's%s", "");
    size_t m = 1;
    while (m < (size_t)c->n_constraints + c->n_inputs) m <<= 1;
    if (m != P.m) fail(MB200_EINVAL, "circuit size does not match the key's domain%s", "");
    for_devices(all_devices(), [&](Device& d) {
        Params& Pd = *reps[d.slot];
        r1cs_upload(d, Pd.r1cs, *c, Pd.idx_aux, Pd.idx_inputs);
    });
    MB_API_END
}

int mb200_circuit_rows(const mb200_circuit* c, size_t n, const uint8_t* inputs, const uint8_t* aux, uint8_t* a_out,
                       uint8_t* b_out, uint8_t* c_out) {
    MB_API_BEGIN
    Device& d = dev0();
    if (!c || (n && (!inputs || !aux || !a_out || !b_out || !c_out))) fail(MB200_EINVAL, "null argument%s", "");
    if (n == 0) return MB200_OK;
    R1csDev R;
    r1cs_upload(d, R, *c, 0, c->n_aux);
    const size_t stride = (size_t)c->n_aux + c->n_inputs, rows = (size_t)c->n_constraints + c->n_inputs;
    DevBuf pool(n * stride * 32), abc(n * 3 * rows * 32), flag(4);
    dev_memset(flag.p, 0, 4, d.main);
    uint8_t* pl = pool.as<uint8_t>();
    copy_rows(pl, stride * 32, aux, (size_t)c->n_aux * 32, (size_t)c->n_aux * 32, n, false, d.main);
    copy_rows(pl + (size_t)c->n_aux * 32, stride * 32, inputs, (size_t)c->n_inputs * 32, (size_t)c->n_inputs * 32, n, false,
              d.main);
    check_scalars_dev(pool.as<Fr>(), n * stride, 0, 1, flag.as<uint32_t>(), d.main);
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
    launch_r1cs_eval(ra, d.main);
    uint32_t bad = 0;
    copy_d2h(&bad, flag.p, 4, d.main);
    uint8_t* outs[3] = {a_out, b_out, c_out};
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k)
            copy_d2h(outs[k] + i * rows * 32, abc.as<uint8_t>() + (i * 3 + k) * rows * 32, rows * 32, d.main);
    stream_sync(d.main);
    if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
    MB_API_END
}

int mb200_verify_batch(const mb200_params* p, size_t n, const uint8_t* proofs_uncompressed, const uint8_t* inputs,
                       uint8_t* ok_out) {
    MB_API_BEGIN
    Device& d = dev0();
    if (n && (!proofs_uncompressed || !inputs || !ok_out)) fail(MB200_EINVAL, "null argument%s", "");
    if (n == 0) return MB200_OK;
    const Params& P = *replicas(p)[0];
    // A | B | C in zkcrypto uncompressed form -> three device arrays
    std::vector<uint8_t> ha(n * 96), hb(n * 192), hc(n * 96);
    for (size_t i = 0; i < n; ++i) {
        memcpy(&ha[i * 96], proofs_uncompressed + i * 384, 96);
        memcpy(&hb[i * 192], proofs_uncompressed + i * 384 + 96, 192);
        memcpy(&hc[i * 96], proofs_uncompressed + i * 384 + 288, 96);
    }
    DevBuf ra(n * 96), rb(n * 192), rc(n * 96), pa(n * sizeof(G1Affine)), pb(n * sizeof(G2Affine)),
        pc(n * sizeof(G1Affine)), bad(4), din(n * P.n_inputs * 32), ok(n * 4);
    copy_h2d(ra.p, ha.data(), ha.size(), d.main);
    copy_h2d(rb.p, hb.data(), hb.size(), d.main);
    copy_h2d(rc.p, hc.data(), hc.size(), d.main);
    copy_h2d(din.p, inputs, n * P.n_inputs * 32, d.main);
    dev_memset(bad.p, 0, 4, d.main);
    DecodeArgs d1{n, ra.as<uint8_t>(), pa.p, bad.as<uint32_t>(), 1}, d2{n, rb.as<uint8_t>(), pb.p, bad.as<uint32_t>(), 1},
        d3{n, rc.as<uint8_t>(), pc.p, bad.as<uint32_t>(), 1};
    launch_decode_g1(d1, d.main);
    launch_decode_g2(d2, d.main);
    launch_decode_g1(d3, d.main);
    check_scalars_dev(din.as<Fr>(), n * P.n_inputs, 0, 1, bad.as<uint32_t>(), d.main);
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
    launch_verify_proofs(va, d.main);
    std::vector<uint32_t> hok(n);
    uint32_t hbad = 0;
    copy_d2h(hok.data(), ok.p, n * 4, d.main);
    copy_d2h(&hbad, bad.p, 4, d.main);
    stream_sync(d.main);
    if (hbad) fail(MB200_EPARSE, "malformed point or non-canonical input%s", "");
    for (size_t i = 0; i < n; ++i) ok_out[i] = hok[i] ? 1 : 0;
    MB_API_END
}

int mb200_verify_proofs(const mb200_params* p, size_t n, const uint8_t* proofs, const uint8_t* inputs, uint8_t* ok_out) {
    MB_API_BEGIN
    Device& d = dev0();
    if (n && (!proofs || !inputs || !ok_out)) fail(MB200_EINVAL, "null argument%s", "");
    if (n == 0) return MB200_OK;
    const Params& P = *replicas(p)[0];
    DevBuf raw(n * 192), pa(n * sizeof(G1Affine)), pb(n * sizeof(G2Affine)), pc(n * sizeof(G1Affine)), bad(n * 4),
        flag(4), din(n * P.n_inputs * 32), ok(n * 4);
    copy_h2d(raw.p, proofs, n * 192, d.main);
    copy_h2d(din.p, inputs, n * P.n_inputs * 32, d.main);
    dev_memset(bad.p, 0, n * 4, d.main);
    dev_memset(flag.p, 0, 4, d.main);
    ProofReadArgs ra{n * 3, raw.as<uint8_t>(), pa.as<G1Affine>(), pb.as<G2Affine>(), pc.as<G1Affine>(), bad.as<uint32_t>()};
    launch_proof_read(ra, d.main);
    check_scalars_dev(din.as<Fr>(), n * P.n_inputs, 0, 1, flag.as<uint32_t>(), d.main);
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
    launch_verify_proofs(va, d.main);
    std::vector<uint32_t> hok(n), hbad(n);
    uint32_t hflag = 0;
    copy_d2h(hok.data(), ok.p, n * 4, d.main);
    copy_d2h(hbad.data(), bad.p, n * 4, d.main);
    copy_d2h(&hflag, flag.p, 4, d.main);
    stream_sync(d.main);
    if (hflag) fail(MB200_ESCALAR, "a public input is not canonical (>= r)%s", "");
    for (size_t i = 0; i < n; ++i) ok_out[i] = (hok[i] && !hbad[i]) ? 1 : 0;
    MB_API_END
}

int mb200_verify_proofs_batch(const mb200_params* p, size_t n, const uint8_t* proofs, const uint8_t* inputs,
                              const uint8_t* z, int* all_ok) {
    MB_API_BEGIN
    Device& d = dev0();
    if (!all_ok || (n && (!proofs || !inputs || !z))) fail(MB200_EINVAL, "null argument%s", "");
    *all_ok = 1;
    if (n == 0) return MB200_OK;
    for (size_t i = 0; i < n; ++i) {  // a zero coefficient would drop proof i from the combined equation
        bool zero = true;
        for (int k = 0; k < 16; ++k) zero = zero && z[16 * i + k] == 0;
        if (zero) fail(MB200_EINVAL, "batch coefficient %s%ld is zero", "", (long)i);
    }
    const Params& P = *replicas(p)[0];
    DevBuf raw(n * 192), pa(n * sizeof(G1Affine)), pb(n * sizeof(G2Affine)), pc(n * sizeof(G1Affine)), bad(n * 4),
        flag(4), din(n * P.n_inputs * 32), dz(n * 16), f(n * sizeof(Fp12)), zc(n * sizeof(G1XYZZ)),
        zx(n * P.n_inputs * sizeof(Fr)), ok(4);
    copy_h2d(raw.p, proofs, n * 192, d.main);
    copy_h2d(din.p, inputs, n * P.n_inputs * 32, d.main);
    copy_h2d(dz.p, z, n * 16, d.main);
    dev_memset(bad.p, 0, n * 4, d.main);
    dev_memset(flag.p, 0, 4, d.main);
    ProofReadArgs ra{n * 3, raw.as<uint8_t>(), pa.as<G1Affine>(), pb.as<G2Affine>(), pc.as<G1Affine>(), bad.as<uint32_t>()};
    launch_proof_read(ra, d.main);
    check_scalars_dev(din.as<Fr>(), n * P.n_inputs, 0, 1, flag.as<uint32_t>(), d.main);
    BatchMillerArgs ma{n, pa.as<G1Affine>(), pb.as<G2Affine>(), pc.as<G1Affine>(), din.as<uint32_t>(), P.n_inputs,
                       dz.as<uint32_t>(), f.as<Fp12>(), zc.as<G1XYZZ>(), zx.as<Fr>()};
    launch_batch_miller(ma, d.main);
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
    launch_batch_final(fa, d.main);
    std::vector<uint32_t> hbad(n);
    uint32_t hok = 0, hflag = 0;
    copy_d2h(&hok, ok.p, 4, d.main);
    copy_d2h(hbad.data(), bad.p, n * 4, d.main);
    copy_d2h(&hflag, flag.p, 4, d.main);
    stream_sync(d.main);
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
    require_init();
    const std::vector<Params*>& reps = replicas(p);
    ProveInputs in{nullptr, nullptr, nullptr, inputs, aux, r, s, false};
    return prove_impl(reps, n_proofs, (size_t)reps[0]->r1cs.ncons + reps[0]->n_inputs, in, proofs_out);
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
    Device& d = dev0();
    if ((n && (!bases || !scalars)) || !out) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf raw(n * 96), pts(n * sizeof(G1Affine)), bad(4), res(sizeof(G1XYZZ)), enc(96);
    copy_h2d(raw.p, bases, n * 96, d.main);
    dev_memset(bad.p, 0, 4, d.main);
    DecodeArgs da{n, raw.as<uint8_t>(), pts.p, bad.as<uint32_t>()};
    launch_decode_g1(da, d.main);
    uint32_t bad_h = 0;
    copy_d2h(&bad_h, bad.p, 4, d.main);
    stream_sync(d.main);
    if (bad_h) fail(MB200_EPARSE, "malformed G1 encoding%s", "");
    msm_on_device_bases<Fp>(d, pts.as<G1Affine>(), scalars, n, res.as<G1XYZZ>());
    EncodeArgs ea{1, res.p, enc.as<uint8_t>()};
    launch_encode_g1(ea, d.main);
    copy_d2h(out, enc.p, 96, d.main);
    stream_sync(d.main);
    MB_API_END
}

int mb200_msm_g2(const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t out[192]) {
    MB_API_BEGIN
    Device& d = dev0();
    if ((n && (!bases || !scalars)) || !out) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf raw(n * 192), pts(n * sizeof(G2Affine)), bad(4), res(sizeof(G2XYZZ)), enc(192);
    copy_h2d(raw.p, bases, n * 192, d.main);
    dev_memset(bad.p, 0, 4, d.main);
    DecodeArgs da{n, raw.as<uint8_t>(), pts.p, bad.as<uint32_t>()};
    launch_decode_g2(da, d.main);
    uint32_t bad_h = 0;
    copy_d2h(&bad_h, bad.p, 4, d.main);
    stream_sync(d.main);
    if (bad_h) fail(MB200_EPARSE, "malformed G2 encoding%s", "");
    msm_on_device_bases<Fp2>(d, pts.as<G2Affine>(), scalars, n, res.as<G2XYZZ>());
    EncodeArgs ea{1, res.p, enc.as<uint8_t>()};
    launch_encode_g2(ea, d.main);
    copy_d2h(out, enc.p, 192, d.main);
    stream_sync(d.main);
    MB_API_END
}

int mb200_g1_bases_upload(const uint8_t* bases, size_t n, void** dev_bases) {
    MB_API_BEGIN
    Device& d = dev0();
    if (!bases || !dev_bases) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf raw(n * 96), bad(4);
    void* pts = dev_alloc(n * sizeof(G1Affine));
    copy_h2d(raw.p, bases, n * 96, d.main);
    dev_memset(bad.p, 0, 4, d.main);
    DecodeArgs da{n, raw.as<uint8_t>(), pts, bad.as<uint32_t>()};
    launch_decode_g1(da, d.main);
    uint32_t bad_h = 0;
    copy_d2h(&bad_h, bad.p, 4, d.main);
    stream_sync(d.main);
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
    Device& d = dev0();
    if ((n && (!dev_bases || !scalars)) || !out_partial) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf res(sizeof(G1XYZZ));
    msm_on_device_bases<Fp>(d, (const G1Affine*)dev_bases, scalars, n, res.as<G1XYZZ>());
    copy_d2h(out_partial, res.p, sizeof(G1XYZZ), d.main);
    stream_sync(d.main);
    MB_API_END
}

int mb200_g1_sum_partials(const uint8_t* partials, size_t count, uint8_t out[96]) {
    MB_API_BEGIN
    Device& d = dev0();
    if ((count && !partials) || !out) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf parts(count * sizeof(G1XYZZ)), enc(96);
    copy_h2d(parts.p, partials, count * sizeof(G1XYZZ), d.main);
    SumPartialsArgs sa{1, parts.as<G1XYZZ>(), count, enc.as<uint8_t>()};
    launch_sum_partials(sa, d.main);
    copy_d2h(out, enc.p, 96, d.main);
    stream_sync(d.main);
    MB_API_END
}

int mb200_msm_g1_partial_device(const void* dev_bases, const uint8_t* scalars, size_t n, void* dev_partial) {
    MB_API_BEGIN
    Device& d = dev0();
    if ((n && (!dev_bases || !scalars)) || !dev_partial) fail(MB200_EINVAL, "null buffer%s", "");
    msm_on_device_bases<Fp>(d, (const G1Affine*)dev_bases, scalars, n, (G1XYZZ*)dev_partial);
    MB_API_END
}

int mb200_g1_sum_partials_device(const void* dev_partials, size_t count, uint8_t out[96]) {
    MB_API_BEGIN
    Device& d = dev0();
    if ((count && !dev_partials) || !out) fail(MB200_EINVAL, "null buffer%s", "");
    d.msm_enc.ensure(96);
    SumPartialsArgs sa{1, (const G1XYZZ*)dev_partials, count, d.msm_enc.as<uint8_t>()};
    launch_sum_partials(sa, d.main);
    copy_d2h(out, d.msm_enc.p, 96, d.main);
    stream_sync(d.main);
    MB_API_END
}

}  // extern "C"

// Bases of a single large MSM, range-split over the opened devices (BASELINE config 3 inside one
// process): shard k holds bases [lo_k, hi_k) decoded into Montgomery form on device slot k.
struct mb200_g1_bases {
    size_t n = 0;
    std::vector<size_t> lo;      // nd + 1 boundaries
    std::vector<void*> shard;    // device pointers, one per slot (nullptr for an empty range)
    std::vector<void*> partial;  // one XYZZ per slot, on that slot's device
};
static void g1_bases_destroy(mb200_g1_bases* b) {
    if (!b) return;
    for (size_t k = 0; k < b->shard.size() && k < g.devs.size(); ++k) {
#ifndef MB200_EMU
        cudaSetDevice(g.devs[k]->id);
#endif
        dev_free(b->shard[k]);
        dev_free(b->partial[k]);
    }
#ifndef MB200_EMU
    if (!g.devs.empty()) cudaSetDevice(g.devs[0]->id);
#endif
    delete b;
}

extern "C" {

int mb200_g1_bases_new(const uint8_t* bases, size_t n, mb200_g1_bases** out) {
    MB_API_BEGIN
    require_init();
    if (!out || (n && !bases)) fail(MB200_EINVAL, "null buffer%s", "");
    *out = nullptr;
    const size_t nd = g.devs.size();
    std::unique_ptr<mb200_g1_bases, void (*)(mb200_g1_bases*)> B(new mb200_g1_bases(), g1_bases_destroy);
    B->n = n;
    B->lo.resize(nd + 1);
    for (size_t k = 0; k <= nd; ++k) B->lo[k] = n / nd * k + std::min(k, n % nd);
    B->shard.assign(nd, nullptr);
    B->partial.assign(nd, nullptr);
    for_devices(all_devices(), [&](Device& d) {
        const size_t lo = B->lo[d.slot], cnt = B->lo[d.slot + 1] - lo;
        B->partial[d.slot] = dev_alloc(sizeof(G1XYZZ));
        if (!cnt) return;
        DevBuf raw(cnt * 96), bad(4);
        B->shard[d.slot] = dev_alloc(cnt * sizeof(G1Affine));
        copy_h2d(raw.p, bases + lo * 96, cnt * 96, d.main);
        dev_memset(bad.p, 0, 4, d.main);
        DecodeArgs da{cnt, raw.as<uint8_t>(), B->shard[d.slot], bad.as<uint32_t>()};
        launch_decode_g1(da, d.main);
        uint32_t bad_h = 0;
        copy_d2h(&bad_h, bad.p, 4, d.main);
        stream_sync(d.main);
        if (bad_h) fail(MB200_EPARSE, "malformed G1 encoding%s", "");
    });
    *out = B.release();
    MB_API_END
}

void mb200_g1_bases_free(mb200_g1_bases* b) {
    std::lock_guard<std::mutex> lk(g.mu);
    g1_bases_destroy(b);
}

/* sum_i s_i B_i with the bases split over the devices: every device reduces its range to one
 * XYZZ partial, the partials travel GPU -> GPU into slot 0's memory (cudaMemcpyPeerAsync: NVLink,
 * no host hop) and one thread there adds them and encodes the result -- SURVEY.md's K6. */
int mb200_msm_g1_bases(const mb200_g1_bases* b, const uint8_t* scalars, size_t n, uint8_t out[96]) {
    MB_API_BEGIN
    require_init();
    if (!b || !out || (n && !scalars)) fail(MB200_EINVAL, "null buffer%s", "");
    if (n != b->n || b->shard.size() != g.devs.size()) fail(MB200_EINVAL, "scalar count does not match the bases%s (%ld)", "", (long)b->n);
    const size_t nd = g.devs.size();
    Device& d0 = *g.devs[0];
    use(d0);
    d0.peer_parts.ensure(nd * sizeof(G1XYZZ));
    G1XYZZ* gathered = d0.peer_parts.as<G1XYZZ>();
    for_devices(all_devices(), [&](Device& d) {
        const size_t lo = b->lo[d.slot], cnt = b->lo[d.slot + 1] - lo;
        G1XYZZ* part = (G1XYZZ*)b->partial[d.slot];
        msm_on_device_bases<Fp>(d, (const G1Affine*)b->shard[d.slot], scalars + lo * 32, cnt, part);
#ifndef MB200_EMU
        if (d.slot == 0) MB_CUDA(cudaMemcpyAsync(gathered, part, sizeof(G1XYZZ), cudaMemcpyDeviceToDevice, d.main));
        else MB_CUDA(cudaMemcpyPeerAsync(gathered + d.slot, d0.id, part, d.id, sizeof(G1XYZZ), d.main));
        stream_sync(d.main);
#else
        memcpy(gathered + d.slot, part, sizeof(G1XYZZ));
#endif
    });
    use(d0);
    d0.msm_enc.ensure(96);
    SumPartialsArgs sa{1, gathered, nd, d0.msm_enc.as<uint8_t>()};
    launch_sum_partials(sa, d0.main);
    copy_d2h(out, d0.msm_enc.p, 96, d0.main);
    stream_sync(d0.main);
    MB_API_END
}

int mb200_ntt(uint8_t* data, unsigned log_n, int inverse, int coset) {
    MB_API_BEGIN
    Device& d = dev0();
    if (!data) fail(MB200_EINVAL, "null buffer%s", "");
    NttCache& c = ntt_cache(d, log_n);
    size_t n = (size_t)1 << log_n;
    DevBuf in(n * 32), t0(n * 32), t1(n * 32), o(n * 32), flag(4);
    copy_h2d(in.p, data, n * 32, d.main);
    dev_memset(flag.p, 0, 4, d.main);
    check_scalars_dev(in.as<Fr>(), n, 0, 1, flag.as<uint32_t>(), d.main);
    NttPlan p;
    p.inverse = inverse != 0;
    if (!inverse && coset) p.in_scale = c.gpow.as<Fr>();
    if (inverse) p.out_scale = coset ? c.dom.cos_inv.as<Fr>() : c.dom.minv_tab.as<Fr>();
    ntt_run(c.dom, p, 1, in.as<Fr>(), n, o.as<Fr>(), n, t0.as<Fr>(), t1.as<Fr>(), d.main);
    uint32_t bad = 0;
    copy_d2h(&bad, flag.p, 4, d.main);
    copy_d2h(data, o.p, n * 32, d.main);
    stream_sync(d.main);
    if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
    MB_API_END
}

int mb200_h_coeffs(const uint8_t* a, const uint8_t* b, const uint8_t* c, size_t rows, uint8_t* out) {
    MB_API_BEGIN
    Device& d = dev0();
    if (!a || !b || !c || !out || rows < 2) fail(MB200_EINVAL, "bad argument (rows must be >= 2)%s", "");
    unsigned log_n = 0;
    while (((size_t)1 << log_n) < rows) log_n++;
    NttCache& cache = ntt_cache(d, log_n);
    size_t n = (size_t)1 << log_n;
    DevBuf abc(3 * rows * 32), w0(3 * n * 32), w1(3 * n * 32), w2(3 * n * 32), w3(3 * n * 32), h(n * 32), flag(4);
    copy_h2d(abc.as<uint8_t>(), a, rows * 32, d.main);
    copy_h2d(abc.as<uint8_t>() + rows * 32, b, rows * 32, d.main);
    copy_h2d(abc.as<uint8_t>() + 2 * rows * 32, c, rows * 32, d.main);
    dev_memset(flag.p, 0, 4, d.main);
    check_scalars_dev(abc.as<Fr>(), 3 * rows, 0, 1, flag.as<uint32_t>(), d.main);
    h_pipeline(cache.dom, 1, (uint32_t)rows, abc.as<Fr>(), rows, h.as<Fr>(), n, w0.as<Fr>(), w1.as<Fr>(), w2.as<Fr>(),
               w3.as<Fr>(), d.main);
    uint32_t bad = 0;
    copy_d2h(&bad, flag.p, 4, d.main);
    copy_d2h(out, h.p, (n - 1) * 32, d.main);
    stream_sync(d.main);
    if (bad) fail(MB200_ESCALAR, "a scalar is not canonical (>= r)%s", "");
    MB_API_END
}

static void fr_mul_dev(Device& d, const void* a, const void* b, size_t n, void* out) {
    FrMulArgs fa{n, (const Fr*)a, (const Fr*)b, (Fr*)out};
    launch_fr_mul_kernel(fa, d.main);
}
int mb200_fr_mul(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
    MB_API_BEGIN
    Device& d = dev0();
    if (n && (!a || !b || !out)) fail(MB200_EINVAL, "null buffer%s", "");
    DevBuf da(n * 32), db(n * 32), dc(n * 32);
    copy_h2d(da.p, a, n * 32, d.main);
    copy_h2d(db.p, b, n * 32, d.main);
    fr_mul_dev(d, da.p, db.p, n, dc.p);
    copy_d2h(out, dc.p, n * 32, d.main);
    stream_sync(d.main);
    MB_API_END
}
int mb200_fr_mul_device(const void* a, const void* b, size_t n, void* out) {
    MB_API_BEGIN
    Device& d = dev0();
    if (n && (!a || !b || !out)) fail(MB200_EINVAL, "null buffer%s", "");
    fr_mul_dev(d, a, b, n, out);
    stream_sync(d.main);
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
        g.n_streams = (uint32_t)value;
        for (auto& d : g.devs) {
            use(*d);
            set_ctx_count(*d, (size_t)value);
        }
        use(*g.devs[0]);
    } else if (!strcmp(name, "msm_slab")) {
        if (value < 1 || value > (1l << 26)) fail(MB200_EINVAL, "msm_slab out of range%s (%ld)", "", value);
        g.msm_slab = (size_t)value;
    } else if (!strcmp(name, "verify")) {
        if (value < 0 || value > 2) fail(MB200_EINVAL, "verify out of range%s (%ld)", "", value);
        g.verify = (int)value;
    } else if (!strcmp(name, "profile")) {
        std::lock_guard<std::mutex> pl(g_msm_profile.mu);
        g_msm_profile.enabled = value != 0;
        g_msm_profile.acc_ms = 0;
        g_msm_profile.acc_launches = 0;
        g_msm_profile.acc_entries_bound = 0;
        for (auto& ph : g_msm_profile.phase_ms) ph = 0;
    } else {
        fail(MB200_EINVAL, "unknown option %s", name);
    }
    MB_API_END
}

int mb200_get_counter(const char* name, double* value) {
    MB_API_BEGIN
    if (!name || !value) fail(MB200_EINVAL, "null argument%s", "");
    std::lock_guard<std::mutex> pl(g_msm_profile.mu);
    if (!strcmp(name, "launches")) *value = (double)g_launches.load();
    else if (!strcmp(name, "acc_launches")) *value = (double)g_msm_profile.acc_launches;
    else if (!strcmp(name, "acc_us")) *value = g_msm_profile.acc_ms * 1000.0;
    else if (!strcmp(name, "acc_bytes")) *value = (double)g_msm_profile.acc_entries_bound;
    else if (!strcmp(name, "msm_upload_us")) *value = g_msm_profile.phase_ms[0] * 1000.0;
    else if (!strcmp(name, "msm_sort_us")) *value = g_msm_profile.phase_ms[1] * 1000.0;
    else if (!strcmp(name, "msm_reduce_us")) *value = g_msm_profile.phase_ms[2] * 1000.0;
    else if (!strcmp(name, "last_batch_us")) *value = g.last_batch_ms * 1000.0;
    else if (!strcmp(name, "verified")) *value = (double)g.verified;
    else if (!strcmp(name, "verify_failed")) *value = (double)g.verify_failed;
    else if (!strcmp(name, "devices")) *value = (double)g.devs.size();
    else fail(MB200_EINVAL, "unknown counter %s", name);
    MB_API_END
}

// Every opened device runs the self-test; the result is the total number of mismatches.
int mb200_selftest(void) {
    std::lock_guard<std::mutex> lk(g.mu);
    try {
        require_init();
        std::atomic<unsigned> total{0};
        for_devices(all_devices(), [&](Device& d) {
            uint32_t bad = 0;
            DevBuf dv(4);
            dev_memset(dv.p, 0, 4, d.main);
            SelfTestArgs a;
            a.nthreads = 4096;
            a.mismatches = dv.as<uint32_t>();
            a.g1 = g1_generator_host();
            a.g2 = g2_generator_host();
            launch_selftest_kernel(a, d.main);
            // pairing: x-chain final exponentiation against the plain power, Frobenius consistency
            DevBuf gens(sizeof(G1Affine) + sizeof(G2Affine));
            G1Affine hg1 = g1_generator_host();
            G2Affine hg2 = g2_generator_host();
            copy_h2d(gens.p, &hg1, sizeof hg1, d.main);
            copy_h2d((char*)gens.p + sizeof(G1Affine), &hg2, sizeof hg2, d.main);
            PairSelfTestArgs pa{1, gens.as<G1Affine>(), (const G2Affine*)((char*)gens.p + sizeof(G1Affine)), dv.as<uint32_t>()};
            launch_pair_selftest(pa, d.main);
            copy_d2h(&bad, dv.p, 4, d.main);
            stream_sync(d.main);
            total += bad;
        });
        return (int)total.load();
    } catch (const Exc& e) {
        return e.code;
    } catch (const std::bad_alloc&) {
        return MB200_ENOMEM;
    }
}

int mb200_bench_fpmul(double* muls_per_second) {
    MB_API_BEGIN
    Device& d = dev0();
    if (!muls_per_second) fail(MB200_EINVAL, "null argument%s", "");
#ifndef MB200_EMU
    FpMulBenchArgs a;
    a.nthreads = (size_t)148 * 2048 * 4;
    a.iters = 512;
    DevBuf sink(a.nthreads * sizeof(Fp));
    a.sink = sink.as<Fp>();
    launch_fpmul_bench(a, d.main);  // warm-up
    cudaEvent_t e0, e1;
    MB_CUDA(cudaEventCreate(&e0));
    MB_CUDA(cudaEventCreate(&e1));
    MB_CUDA(cudaEventRecord(e0, d.main));
    launch_fpmul_bench(a, d.main);
    MB_CUDA(cudaEventRecord(e1, d.main));
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
    Device& d = dev0();
    if (!ns_per_op || mode < 0 || mode > 3) fail(MB200_EINVAL, "bad argument%s", "");
#ifndef MB200_EMU
    LatencyArgs a;
    a.nthreads = 32;
    a.iters = 2000;
    a.mode = (uint32_t)mode;
    a.g1 = g1_generator_host();
    DevBuf sink(32 * sizeof(Fp));
    a.sink = sink.as<Fp>();
    launch_latency_kernel(a, d.main);
    cudaEvent_t e0, e1;
    MB_CUDA(cudaEventCreate(&e0));
    MB_CUDA(cudaEventCreate(&e1));
    MB_CUDA(cudaEventRecord(e0, d.main));
    launch_latency_kernel(a, d.main);
    MB_CUDA(cudaEventRecord(e1, d.main));
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
        case MB200_EPARAMS: return "parameter file size or BLAKE2b digest is not the expected one";
        case MB200_EIO: return "cannot read the parameter file";
        default: return "unknown error";
    }
}
const char* mb200_last_error(void) { return last_error().msg; }

}  // extern "C"
