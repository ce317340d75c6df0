// Pippenger bucket MSM over BLS12-381 G1 / G2, batched over independent
// instances (one instance = one proof's query, or one standalone MSM).
//
// Replaces bellperson's `multiexp` for the H, L, A, B1 (G1) and B2 (G2)
// queries (call sites of create_random_proof at
// masp_proofs/src/sapling/prover.rs:116-117, 201-202, 251-252).
//
// Schedule (B200-first, not the reference's window-per-task loop):
//   1. msm_count    one thread per (instance, base): signed c-bit digits of the
//                   scalar -> histogram over buckets.  0 is skipped; 1 goes to
//                   a spread set of "ones" buckets (the reference's 0/1 fast
//                   path) so no bucket is a hot spot.
//   2. scan         exclusive prefix sum of the histogram.
//   3. msm_scatter  same walk, writes (table index, sign) entries bucket-sorted.
//   4. msm_accumulate  one thread per bucket: XYZZ accumulator in registers,
//                   affine table points gathered from HBM/L2, 8M+2S mixed add.
//   5. reduce       sum_b b * S_b by chunked running sums, a few levels.
// With a precomputed table (2^(c w) * B_k for every window w, built once at
// key load and resident in HBM) all windows of an instance share ONE bucket
// set: no per-window reduction and no doublings on the proving path.
#pragma once
#include <mutex>

#include "ec.cuh"

namespace mb {


struct MsmClass {
    const void* table;     // Affine<F>[ (precomp ? nwin : 1) * n_bases ]
    const uint32_t* sel;   // n_bases: index of each base's scalar inside an instance's scalar pool (nullptr: base k <-> scalar k)
    uint32_t n_bases;
    uint32_t c;            // window bits
    uint32_t nwin;         // windows: nwin * c >= 256
    uint32_t nb;           // 2^(c-1): digit magnitudes 1..nb
    uint32_t precomp;      // 1: table holds every window, one bucket set; 0: one bucket set per window
    uint32_t nsets;        // precomp ? 1 : nwin
    uint32_t n_ones;       // spread buckets for scalar == 1 (per set; only set 0 is populated)
    uint32_t set_stride;   // nb + 1 + n_ones
    uint32_t inst_stride;  // nsets * set_stride
};

inline uint32_t msm_nwin(uint32_t c) { return (256 + c - 1) / c; }

inline MsmClass msm_make_class(const void* table, const uint32_t* sel, uint32_t n_bases, uint32_t c, bool precomp,
                               uint32_t n_ones = 0) {
    MsmClass k;
    k.table = table;
    k.sel = sel;
    k.n_bases = n_bases;
    k.c = c;
    k.nwin = msm_nwin(c);
    k.nb = 1u << (c - 1);
    k.precomp = precomp ? 1 : 0;
    k.nsets = precomp ? 1 : k.nwin;
    uint32_t ones = n_ones ? n_ones : n_bases / 256;  // n_ones given: slabs of one MSM share a bucket layout
    if (ones < 1) ones = 1;
    if (ones > 256) ones = 256;
    k.n_ones = ones;
    k.set_stride = k.nb + 1 + k.n_ones;
    k.inst_stride = k.nsets * k.set_stride;
    return k;
}

// ---------------------------------------------------------------------------
// 1 + 3: digit walk.  MODE 0 counts, MODE 1 scatters.
// ---------------------------------------------------------------------------
struct DigitArgs {
    size_t nthreads;  // n_inst * n_bases
    MsmClass k;
    const uint32_t* pool;   // scalars, 8 limbs each, plain little-endian, < r
    size_t pool_stride;     // scalars per instance
    uint32_t* counts;       // [n_inst * inst_stride]           (MODE 0)
    uint32_t* cursor;       // running write position per bucket (MODE 1)
    uint32_t* entries;      // (table index << 1) | negate        (MODE 1)
};

template <int MODE>
MB_HD void digit_emit(const DigitArgs& a, size_t bucket, uint32_t entry) {
    if (MODE == 0) {
        MB_ATOMIC_ADD(&a.counts[bucket], 1u);
    } else {
        uint32_t pos = MB_ATOMIC_ADD(&a.cursor[bucket], 1u);
        a.entries[pos] = entry;
    }
}

template <int MODE>
MB_HD void digit_body(const DigitArgs& a, size_t tid) {
    const MsmClass& k = a.k;
    size_t inst = tid / k.n_bases;
    uint32_t base = (uint32_t)(tid - inst * k.n_bases);
    const uint32_t* sp = a.pool + (inst * a.pool_stride + (k.sel ? k.sel[base] : base)) * 8;
    uint32_t s[8];
    uint32_t any_hi = 0;
    MB_UNROLL
    for (int i = 0; i < 8; ++i) s[i] = sp[i];
    MB_UNROLL
    for (int i = 1; i < 8; ++i) any_hi |= s[i];
    if (any_hi == 0 && s[0] == 0) return;
    size_t b0 = inst * k.inst_stride;
    if (any_hi == 0 && s[0] == 1) {
        digit_emit<MODE>(a, b0 + k.nb + 1 + base % k.n_ones, base << 1);
        return;
    }
    uint32_t carry = 0;
    const uint32_t c = k.c, mask = (1u << c) - 1;
    for (uint32_t w = 0; w < k.nwin; ++w) {
        uint32_t bit = w * c;
        uint32_t limb = bit >> 5, off = bit & 31;
        uint32_t d = 0;
        if (limb < 8) {
            d = s[limb] >> off;
            if (off + c > 32 && limb + 1 < 8) d |= s[limb + 1] << (32 - off);
            d &= mask;
        }
        d += carry;
        uint32_t neg = 0;
        carry = 0;
        if (d > k.nb) {  // use d - 2^c, borrow one from the next window
            d = (1u << c) - d;
            neg = 1;
            carry = 1;
        }
        if (d == 0) continue;
        uint32_t set = k.precomp ? 0 : w;
        uint32_t tw = k.precomp ? w : 0;
        digit_emit<MODE>(a, b0 + (size_t)set * k.set_stride + d, ((tw * k.n_bases + base) << 1) | neg);
    }
}
MB_HD void digit_count_body(const DigitArgs& a, size_t tid) { digit_body<0>(a, tid); }
MB_HD void digit_scatter_body(const DigitArgs& a, size_t tid) { digit_body<1>(a, tid); }
MB_K_MSM_G1(msm_count, DigitArgs, digit_count_body, 256)
MB_K_MSM_G1(msm_scatter, DigitArgs, digit_scatter_body, 256)

// ---------------------------------------------------------------------------
// 2: exclusive scan of `counts` -> offsets and cursor (three small kernels)
// ---------------------------------------------------------------------------
struct ScanArgs {
    size_t nthreads;
    const uint32_t* counts;
    uint32_t* offsets;
    uint32_t* cursor;
    uint32_t* partial;  // one per chunk
    size_t n;           // elements
    uint32_t chunk;     // elements per thread
};
MB_HD void scan_sum_body(const ScanArgs& a, size_t tid) {
    size_t lo = tid * a.chunk, hi = lo + a.chunk;
    if (hi > a.n) hi = a.n;
    uint32_t s = 0;
    for (size_t i = lo; i < hi; ++i) s += a.counts[i];
    a.partial[tid] = s;
}
MB_HD void scan_top_body(const ScanArgs& a, size_t) {
    size_t nchunks = (a.n + a.chunk - 1) / a.chunk;
    uint32_t run = 0;
    for (size_t i = 0; i < nchunks; ++i) {
        uint32_t v = a.partial[i];
        a.partial[i] = run;
        run += v;
    }
}
MB_HD void scan_write_body(const ScanArgs& a, size_t tid) {
    size_t lo = tid * a.chunk, hi = lo + a.chunk;
    if (hi > a.n) hi = a.n;
    uint32_t run = a.partial[tid];
    for (size_t i = lo; i < hi; ++i) {
        uint32_t v = a.counts[i];
        a.offsets[i] = run;
        if (a.cursor) a.cursor[i] = run;
        run += v;
    }
}
MB_K_MSM_G1(scan_sum, ScanArgs, scan_sum_body, 128)
MB_K_MSM_G1(scan_top, ScanArgs, scan_top_body, 32)
MB_K_MSM_G1(scan_write, ScanArgs, scan_write_body, 128)

// ---------------------------------------------------------------------------
// 3b: tasks.  A bucket's entry list is cut into segments of at most SEG_LEN
// entries; one thread accumulates one segment and the (few) partial sums of a
// split bucket are added afterwards.  No single bucket can serialise the
// launch: with merged windows the top, partly filled window concentrates its
// digits on a handful of buckets, and real witnesses repeat small values.
// Tasks are handed to threads in descending length (a counting sort), so the
// 32 lanes of a warp run equal trip counts and the long tasks start first.
// ---------------------------------------------------------------------------
static const uint32_t SEG_LEN = 128;
static const uint32_t ORDER_CAP = SEG_LEN + 1;
struct SegArgs {
    size_t nthreads;          // buckets
    const uint32_t* counts;   // entries per bucket
    const uint32_t* offsets;  // first entry of each bucket
    uint32_t* nseg;           // segments per bucket (>= 1)
    const uint32_t* seg_off;  // exclusive scan of nseg
    uint32_t* task_bucket;    // per task
    uint32_t* task_start;     // per task: first entry
    uint32_t* task_len;       // per task: entries
    uint32_t* ntasks;         // total number of tasks
};
MB_HD void seg_count_body(const SegArgs& a, size_t tid) {
    uint32_t n = a.counts[tid];
    a.nseg[tid] = n == 0 ? 1u : (n + SEG_LEN - 1) / SEG_LEN;
}
MB_HD void seg_fill_body(const SegArgs& a, size_t tid) {
    uint32_t n = a.counts[tid], ns = a.nseg[tid], t0 = a.seg_off[tid], e0 = a.offsets[tid];
    for (uint32_t j = 0; j < ns; ++j) {
        uint32_t done = j * SEG_LEN;
        a.task_bucket[t0 + j] = (uint32_t)tid;
        a.task_start[t0 + j] = e0 + done;
        a.task_len[t0 + j] = n - done < SEG_LEN ? n - done : SEG_LEN;
    }
    if (tid + 1 == a.nthreads) *a.ntasks = t0 + ns;
}
MB_K_MSM_G1(seg_count, SegArgs, seg_count_body, 256)
MB_K_MSM_G1(seg_fill, SegArgs, seg_fill_body, 256)

struct OrderArgs {
    size_t nthreads;  // upper bound on tasks (hist / scatter) or 1 (scan)
    const uint32_t* task_len;
    const uint32_t* ntasks;
    uint32_t* hist;   // ORDER_CAP bins, zeroed
    uint32_t* order;  // task ids, longest first
};
MB_HD void order_hist_body(const OrderArgs& a, size_t tid) {
    if (tid >= *a.ntasks) return;
    MB_ATOMIC_ADD(&a.hist[SEG_LEN - a.task_len[tid]], 1u);
}
MB_HD void order_scan_body(const OrderArgs& a, size_t) {
    uint32_t run = 0;
    for (uint32_t i = 0; i < ORDER_CAP; ++i) {
        uint32_t v = a.hist[i];
        a.hist[i] = run;
        run += v;
    }
}
MB_HD void order_scatter_body(const OrderArgs& a, size_t tid) {
    if (tid >= *a.ntasks) return;
    uint32_t pos = MB_ATOMIC_ADD(&a.hist[SEG_LEN - a.task_len[tid]], 1u);
    a.order[pos] = (uint32_t)tid;
}
MB_K_MSM_G1(order_hist, OrderArgs, order_hist_body, 256)
MB_K_MSM_G1(order_scan, OrderArgs, order_scan_body, 32)
MB_K_MSM_G1(order_scatter, OrderArgs, order_scatter_body, 256)

// ---------------------------------------------------------------------------
// 4: segment accumulation (the hot kernel)
// ---------------------------------------------------------------------------
template <class F>
struct AccArgs {
    size_t nthreads;  // upper bound on tasks
    const Affine<F>* table;
    const uint32_t* entries;
    const uint32_t* task_start;
    const uint32_t* task_len;
    const uint32_t* order;  // thread -> task
    const uint32_t* ntasks;
    XYZZ<F>* partials;      // per task
    // slabs of one standalone MSM share their buckets: a bucket's first segment starts from the sum the
    // earlier slabs left there instead of the identity (nullptr on the proving path and for the first slab)
    const XYZZ<F>* carry;
    const uint32_t* task_bucket;
    const uint32_t* seg_off;
};
template <class F>
MB_HD void acc_body(const AccArgs<F>& a, size_t tid) {
    if (tid >= *a.ntasks) return;
    uint32_t t = a.order[tid];
    uint32_t n = a.task_len[t];
    XYZZ<F> acc = XYZZ<F>::inf();
    if (a.carry) {
        const uint32_t b = a.task_bucket[t];
        if (a.seg_off[b] == t) acc = a.carry[b];
    }
    const uint32_t* e = a.entries + a.task_start[t];
    MB_NOUNROLL
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t ent = e[i];
        Affine<F> q = a.table[ent >> 1];
        xyzz_madd(acc, q, (ent & 1) != 0);
    }
    a.partials[t] = acc;
}
MB_HD void acc_g1_body(const AccArgs<Fp>& a, size_t tid) { acc_body<Fp>(a, tid); }
MB_HD void acc_g2_body(const AccArgs<Fp2>& a, size_t tid) { acc_body<Fp2>(a, tid); }
MB_K_MSM_G1(msm_accumulate_g1, AccArgs<Fp>, acc_g1_body, 128)
// Measured on B200 and rejected (profiles/r02_variants_ab.jsonl, r01_acc_128reg_ab.jsonl; baseline 492-494
// Spend proofs/s): a 128-register cap for a fourth resident block (-1 %); the accumulator in shared
// memory, 4 blocks per SM at 128 registers (414: `no_instructions` stalls x 3.6, the i-cache thrashes with
// 16 warps at 16 places of a 72 KB loop body); one lock-step block per SM with a barrier per
// iteration (480: `no_instructions` gone, 94 k -> 1 k samples, but the multiplier drops from 82 % to
// 78 % busy waiting at the barrier); lock-step with L2 prefetch (476) and with shared accumulators
// (449).  The plain kernel at 166 registers and 12 warps per SM stays.
MB_K_MSM_G2(msm_accumulate_g2, AccArgs<Fp2>, acc_g2_body, 64)
// (its accumulator in shared memory -- 168 registers, 12 warps per SM instead of 255 and 8 --
// measured 457 against 492 proofs/s: the ~800 LDS / STS per addition cost more than the spills.)

// bucket sum = sum of its segments' partial sums.  Almost every bucket has ONE segment; a heavy
// bucket (the partly filled top window of a 255-bit scalar puts n / 8 entries into each of eight
// buckets; real witnesses repeat small values) has thousands, and one thread adding them serially
// was the cliff of the 2^24 MSM: 420 of 650 ms in this step (profiles/r02_launches_msm24.csv).  So the
// partial sums of a bucket are folded as an in-place tree of fan-in COMBINE_F: in round r the thread
// of segment j (j a multiple of F^(r+1)) adds the slots j + i F^r, i < F, into its own.  Groups are
// disjoint, so no second buffer and no synchronisation inside a round; buckets with a single segment
// cost one early exit per round.  After ceil(log_F(max segments)) rounds slot 0 holds the bucket sum.
static const uint32_t COMBINE_F = 8;
template <class F>
struct CombineRoundArgs {
    size_t nthreads;  // upper bound on tasks
    XYZZ<F>* partials;
    const uint32_t* task_bucket;
    const uint32_t* seg_off;
    const uint32_t* nseg;
    const uint32_t* ntasks;
    uint32_t stride;  // COMBINE_F^round
};
template <class F>
MB_HD void combine_round_body(const CombineRoundArgs<F>& a, size_t tid) {
    if (tid >= *a.ntasks) return;
    const uint32_t b = a.task_bucket[tid];
    const uint32_t ns = a.nseg[b];
    if (ns <= a.stride) return;  // nothing left to fold in this bucket
    const uint32_t j = (uint32_t)tid - a.seg_off[b];
    if (j % (a.stride * COMBINE_F) != 0 || j + a.stride >= ns) return;
    XYZZ<F> acc = a.partials[tid];
    MB_NOUNROLL
    for (uint32_t i = 1; i < COMBINE_F; ++i) {
        uint64_t idx = (uint64_t)j + (uint64_t)i * a.stride;
        if (idx >= ns) break;
        xyzz_add_cold(acc, a.partials[tid + (size_t)i * a.stride]);
    }
    a.partials[tid] = acc;
}
MB_HD void combine_round_g1_body(const CombineRoundArgs<Fp>& a, size_t tid) { combine_round_body<Fp>(a, tid); }
MB_HD void combine_round_g2_body(const CombineRoundArgs<Fp2>& a, size_t tid) { combine_round_body<Fp2>(a, tid); }
MB_K_RED_G1(msm_combine_round_g1, CombineRoundArgs<Fp>, combine_round_g1_body, 128)
MB_K_RED_G2(msm_combine_round_g2, CombineRoundArgs<Fp2>, combine_round_g2_body, 64)

template <class F>
struct CombineArgs {
    size_t nthreads;  // buckets
    const XYZZ<F>* partials;
    const uint32_t* seg_off;
    XYZZ<F>* buckets;
};
template <class F>
MB_HD void combine_body(const CombineArgs<F>& a, size_t tid) {
    a.buckets[tid] = a.partials[a.seg_off[tid]];
}
MB_HD void combine_g1_body(const CombineArgs<Fp>& a, size_t tid) { combine_body<Fp>(a, tid); }
MB_HD void combine_g2_body(const CombineArgs<Fp2>& a, size_t tid) { combine_body<Fp2>(a, tid); }
MB_K_RED_G1(msm_combine_g1, CombineArgs<Fp>, combine_g1_body, 128)
MB_K_RED_G2(msm_combine_g2, CombineArgs<Fp2>, combine_g2_body, 64)

// ---------------------------------------------------------------------------
// 5: reduction  W = sum_i i * X_i  (+ plain sum of the ones buckets)
//
// Chunks of T consecutive entries: run_k = sum X_i, acc_k = sum (i - kT) X_i.
//   W(X) = sum_k acc_k + T * W(run)
// Carrying P (a plain-sum array) alongside X,
//   level 0:  X' = run, P' = acc                 (ones chunks: X' = 0, P' = sum)
//   level l:  X' = run, P'_k = sum P_j + T^l * acc_k
// until one entry is left; the answer is P[0].
// ---------------------------------------------------------------------------
template <class F>
struct RedArgs {
    size_t nthreads;      // jobs * n_out
    const XYZZ<F>* X;     // [jobs][stride_in]
    const XYZZ<F>* P;     // level >= 1 only
    XYZZ<F>* Xo;          // [jobs][stride_out]
    XYZZ<F>* Po;
    uint32_t n_weighted;  // level 0: nb + 1; level >= 1: L (all entries)
    uint32_t n_plain;     // level 0: n_ones; level >= 1: 0
    uint32_t stride_in, stride_out;
    uint32_t n_wout;      // ceil(n_weighted / T)
    uint32_t n_out;       // n_wout + ceil(n_plain / T)
    uint32_t T;
    uint32_t shift;       // doublings applied to acc at this level: level * log2(T)
    uint32_t level;
};
template <class F>
MB_HD void red_body(const RedArgs<F>& a, size_t tid) {
    size_t job = tid / a.n_out;
    uint32_t k = (uint32_t)(tid - job * a.n_out);
    const XYZZ<F>* X = a.X + job * a.stride_in;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    if (k < a.n_wout) {
        uint32_t lo = k * a.T, hi = lo + a.T;
        if (hi > a.n_weighted) hi = a.n_weighted;
        MB_NOUNROLL
        for (uint32_t j = hi; j-- > lo + 1;) {
            xyzz_add_cold(run, X[j]);
            xyzz_add_cold(acc, run);
        }
        xyzz_add_cold(run, X[lo]);
        MB_NOUNROLL
        for (uint32_t d = 0; d < a.shift; ++d) acc = xyzz_dbl_cold(acc);
        if (a.level > 0) {
            const XYZZ<F>* P = a.P + job * a.stride_in;
            MB_NOUNROLL
            for (uint32_t j = lo; j < hi; ++j) xyzz_add_cold(acc, P[j]);
        }
    } else {  // level 0 only: a chunk of ones buckets, plain sum
        uint32_t lo = a.n_weighted + (k - a.n_wout) * a.T, hi = lo + a.T;
        if (hi > a.n_weighted + a.n_plain) hi = a.n_weighted + a.n_plain;
        MB_NOUNROLL
        for (uint32_t j = lo; j < hi; ++j) xyzz_add_cold(acc, X[j]);
    }
    a.Xo[job * a.stride_out + k] = run;
    a.Po[job * a.stride_out + k] = acc;
}
MB_HD void red_g1_body(const RedArgs<Fp>& a, size_t tid) { red_body<Fp>(a, tid); }
MB_HD void red_g2_body(const RedArgs<Fp2>& a, size_t tid) { red_body<Fp2>(a, tid); }
MB_K_RED_G1(msm_reduce_g1, RedArgs<Fp>, red_g1_body, 64)
MB_K_RED_G2(msm_reduce_g2, RedArgs<Fp2>, red_g2_body, 32)

// Horner over per-window results (non-precomputed tables only):
// out[inst] = sum_w 2^(c w) * R[inst][w]
template <class F>
struct HornerArgs {
    size_t nthreads;  // instances
    const XYZZ<F>* R;
    uint32_t r_stride;  // entries between consecutive window results
    uint32_t nsets, c;
    XYZZ<F>* out;
};
template <class F>
MB_HD void horner_body(const HornerArgs<F>& a, size_t tid) {
    const XYZZ<F>* R = a.R + tid * (size_t)a.nsets * a.r_stride;
    XYZZ<F> acc = R[(size_t)(a.nsets - 1) * a.r_stride];
    MB_NOUNROLL
    for (uint32_t w = a.nsets - 1; w-- > 0;) {
        MB_NOUNROLL
        for (uint32_t d = 0; d < a.c; ++d) acc = xyzz_dbl_cold(acc);
        xyzz_add_cold(acc, R[(size_t)w * a.r_stride]);
    }
    a.out[tid] = acc;
}
MB_HD void horner_g1_body(const HornerArgs<Fp>& a, size_t tid) { horner_body<Fp>(a, tid); }
MB_HD void horner_g2_body(const HornerArgs<Fp2>& a, size_t tid) { horner_body<Fp2>(a, tid); }
MB_K_RED_G1(msm_horner_g1, HornerArgs<Fp>, horner_g1_body, 32)
MB_K_RED_G2(msm_horner_g2, HornerArgs<Fp2>, horner_g2_body, 32)

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
struct MsmScratch {
    DevBuf counts, offsets, cursor, partial, entries, buckets, lx[2], lp[2], order, ohist;
    DevBuf nseg, seg_off, task_bucket, task_start, task_len, ntasks, partials;
};

struct MsmProfile {  // optional CUDA-event timing of the accumulate kernel (all devices add into it)
    std::mutex mu;
    bool enabled = false;
    double acc_ms = 0;
    unsigned long long acc_launches = 0;
    unsigned long long acc_entries_bound = 0;
    double phase_ms[3] = {0, 0, 0};  // standalone MSM: scalar upload, digit sort, bucket reduction
};
extern MsmProfile g_msm_profile;

static const uint32_t SCAN_CHUNK = 512;
// reduction fan-in (log2): wide at level 0, narrow above it -- swept on B200 (profiles/r01_reduce_fanin_sweep.jsonl)
static const uint32_t RED_LOG_T0 = 3, RED_LOG_T1 = 2;

template <class F>
inline void launch_acc(const AccArgs<F>& a, cudaStream_t s);
template <>
inline void launch_acc<Fp>(const AccArgs<Fp>& a, cudaStream_t s) { launch_msm_accumulate_g1(a, s); }
template <>
inline void launch_acc<Fp2>(const AccArgs<Fp2>& a, cudaStream_t s) { launch_msm_accumulate_g2(a, s); }
template <class F>
inline void launch_combine_round(const CombineRoundArgs<F>& a, cudaStream_t s);
template <>
inline void launch_combine_round<Fp>(const CombineRoundArgs<Fp>& a, cudaStream_t s) { launch_msm_combine_round_g1(a, s); }
template <>
inline void launch_combine_round<Fp2>(const CombineRoundArgs<Fp2>& a, cudaStream_t s) { launch_msm_combine_round_g2(a, s); }
template <class F>
inline void launch_combine(const CombineArgs<F>& a, cudaStream_t s);
template <>
inline void launch_combine<Fp>(const CombineArgs<Fp>& a, cudaStream_t s) { launch_msm_combine_g1(a, s); }
template <>
inline void launch_combine<Fp2>(const CombineArgs<Fp2>& a, cudaStream_t s) { launch_msm_combine_g2(a, s); }
template <class F>
inline void launch_red(const RedArgs<F>& a, cudaStream_t s);
template <>
inline void launch_red<Fp>(const RedArgs<Fp>& a, cudaStream_t s) { launch_msm_reduce_g1(a, s); }
template <>
inline void launch_red<Fp2>(const RedArgs<Fp2>& a, cudaStream_t s) { launch_msm_reduce_g2(a, s); }
template <class F>
inline void launch_horner(const HornerArgs<F>& a, cudaStream_t s);
template <>
inline void launch_horner<Fp>(const HornerArgs<Fp>& a, cudaStream_t s) { launch_msm_horner_g1(a, s); }
template <>
inline void launch_horner<Fp2>(const HornerArgs<Fp2>& a, cudaStream_t s) { launch_msm_horner_g2(a, s); }

// Runs one class over n_inst instances.  out: n_inst XYZZ results (device), complete on s.
// (Batched-affine "pair rounds" in front of the accumulation -- 6 instead of 10 multiplications per
// addition, one shared inversion per launch -- were built and measured in round 1: 434 / 416 / 402
// proofs/s with 1 / 2 / 3 rounds against 484 without, profiles/r01_pair_rounds_v2_sweep.jsonl; removed.)
template <class F>
void msm_reduce_buckets(const MsmClass& k, uint32_t n_inst, XYZZ<F>* out, MsmScratch& w, cudaStream_t s);

// Steps 1-4 and the fold of split buckets: w.buckets receives (add_into: is increased by) the bucket sums.
template <class F>
void msm_accumulate_buckets(const MsmClass& k, uint32_t n_inst, const uint32_t* pool, size_t pool_stride,
                            MsmScratch& w, cudaStream_t s, bool add_into);

template <class F>
void msm_run(const MsmClass& k, uint32_t n_inst, const uint32_t* pool, size_t pool_stride, XYZZ<F>* out,
             MsmScratch& w, cudaStream_t s) {
    if (n_inst == 0) return;
    msm_accumulate_buckets<F>(k, n_inst, pool, pool_stride, w, s, false);
    msm_reduce_buckets<F>(k, n_inst, out, w, s);
}

template <class F>
void msm_accumulate_buckets(const MsmClass& k, uint32_t n_inst, const uint32_t* pool, size_t pool_stride,
                            MsmScratch& w, cudaStream_t s, bool add_into) {
    if (n_inst == 0) return;
    size_t nbuckets = (size_t)n_inst * k.inst_stride;
    size_t max_entries = (size_t)n_inst * k.n_bases * k.nwin;
    if (max_entries >= (1ull << 32) || (size_t)k.nwin * k.n_bases >= (1ull << 31))
        fail(MB200_EINVAL, "MSM too large for 32-bit entry indices%s (%ld entries)", "", (long)max_entries);
    size_t nchunks = (nbuckets + SCAN_CHUNK - 1) / SCAN_CHUNK;
    w.counts.ensure(nbuckets * 4);
    w.offsets.ensure(nbuckets * 4);
    w.cursor.ensure(nbuckets * 4);
    w.partial.ensure(nchunks * 4);
    w.entries.ensure(max_entries * 4);
    w.buckets.ensure(nbuckets * sizeof(XYZZ<F>));
    dev_memset(w.counts.p, 0, nbuckets * 4, s);

    DigitArgs da;
    da.nthreads = (size_t)n_inst * k.n_bases;
    da.k = k;
    da.pool = pool;
    da.pool_stride = pool_stride;
    da.counts = w.counts.as<uint32_t>();
    da.cursor = w.cursor.as<uint32_t>();
    da.entries = w.entries.as<uint32_t>();
    launch_msm_count(da, s);

    ScanArgs sa;
    sa.counts = w.counts.as<uint32_t>();
    sa.offsets = w.offsets.as<uint32_t>();
    sa.cursor = w.cursor.as<uint32_t>();
    sa.partial = w.partial.as<uint32_t>();
    sa.n = nbuckets;
    sa.chunk = SCAN_CHUNK;
    sa.nthreads = nchunks;
    launch_scan_sum(sa, s);
    sa.nthreads = 1;
    launch_scan_top(sa, s);
    sa.nthreads = nchunks;
    launch_scan_write(sa, s);

    launch_msm_scatter(da, s);

    const uint32_t* b_counts = w.counts.as<uint32_t>();
    const uint32_t* b_offsets = w.offsets.as<uint32_t>();

    // tasks: segments of at most SEG_LEN entries, longest first
    size_t max_tasks = nbuckets + max_entries / SEG_LEN + 1;
    if (max_tasks >= (1ull << 32)) fail(MB200_EINVAL, "too many buckets%s (%ld)", "", (long)nbuckets);
    w.nseg.ensure(nbuckets * 4);
    w.seg_off.ensure(nbuckets * 4);
    w.task_bucket.ensure(max_tasks * 4);
    w.task_start.ensure(max_tasks * 4);
    w.task_len.ensure(max_tasks * 4);
    w.order.ensure(max_tasks * 4);
    w.ntasks.ensure(4);
    w.ohist.ensure(ORDER_CAP * 4);
    w.partials.ensure(max_tasks * sizeof(XYZZ<F>));
    SegArgs ga;
    ga.nthreads = nbuckets;
    ga.counts = b_counts;
    ga.offsets = b_offsets;
    ga.nseg = w.nseg.as<uint32_t>();
    ga.seg_off = w.seg_off.as<uint32_t>();
    ga.task_bucket = w.task_bucket.as<uint32_t>();
    ga.task_start = w.task_start.as<uint32_t>();
    ga.task_len = w.task_len.as<uint32_t>();
    ga.ntasks = w.ntasks.as<uint32_t>();
    launch_seg_count(ga, s);
    ScanArgs sg;
    sg.counts = w.nseg.as<uint32_t>();
    sg.offsets = w.seg_off.as<uint32_t>();
    sg.cursor = nullptr;
    sg.partial = w.partial.as<uint32_t>();
    sg.n = nbuckets;
    sg.chunk = SCAN_CHUNK;
    sg.nthreads = nchunks;
    launch_scan_sum(sg, s);
    sg.nthreads = 1;
    launch_scan_top(sg, s);
    sg.nthreads = nchunks;
    launch_scan_write(sg, s);
    launch_seg_fill(ga, s);

    dev_memset(w.ohist.p, 0, ORDER_CAP * 4, s);
    OrderArgs oa;
    oa.task_len = w.task_len.as<uint32_t>();
    oa.ntasks = w.ntasks.as<uint32_t>();
    oa.hist = w.ohist.as<uint32_t>();
    oa.order = w.order.as<uint32_t>();
    oa.nthreads = max_tasks;
    launch_order_hist(oa, s);
    oa.nthreads = 1;
    launch_order_scan(oa, s);
    oa.nthreads = max_tasks;
    launch_order_scatter(oa, s);

    AccArgs<F> aa;
    aa.nthreads = max_tasks;
    aa.table = (const Affine<F>*)k.table;
    aa.entries = w.entries.as<uint32_t>();
    aa.task_start = w.task_start.as<uint32_t>();
    aa.task_len = w.task_len.as<uint32_t>();
    aa.order = w.order.as<uint32_t>();
    aa.ntasks = w.ntasks.as<uint32_t>();
    aa.partials = w.partials.as<XYZZ<F>>();
    aa.carry = add_into ? w.buckets.as<XYZZ<F>>() : nullptr;
    aa.task_bucket = w.task_bucket.as<uint32_t>();
    aa.seg_off = w.seg_off.as<uint32_t>();
#ifndef MB200_EMU
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_msm_profile.enabled) {
        MB_CUDA(cudaEventCreate(&e0));
        MB_CUDA(cudaEventCreate(&e1));
        MB_CUDA(cudaEventRecord(e0, s));
    }
#endif
    launch_acc<F>(aa, s);
#ifndef MB200_EMU
    if (g_msm_profile.enabled) {
        MB_CUDA(cudaEventRecord(e1, s));
        MB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        {
            std::lock_guard<std::mutex> pl(g_msm_profile.mu);
            g_msm_profile.acc_ms += ms;
            g_msm_profile.acc_launches++;
            g_msm_profile.acc_entries_bound += (unsigned long long)n_inst * k.n_bases * (sizeof(Affine<F>) + 32);
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
#endif
    {   // fold the partial sums of split buckets: rounds of fan-in COMBINE_F, enough for the fullest possible bucket
        uint64_t worst = ((uint64_t)k.n_bases * (k.precomp ? k.nwin : 1) + SEG_LEN - 1) / SEG_LEN;
        CombineRoundArgs<F> cr;
        cr.nthreads = max_tasks;
        cr.partials = w.partials.as<XYZZ<F>>();
        cr.task_bucket = w.task_bucket.as<uint32_t>();
        cr.seg_off = w.seg_off.as<uint32_t>();
        cr.nseg = w.nseg.as<uint32_t>();
        cr.ntasks = w.ntasks.as<uint32_t>();
        for (uint64_t stride = 1; stride < worst; stride *= COMBINE_F) {
            cr.stride = (uint32_t)stride;
            launch_combine_round<F>(cr, s);
        }
    }
    CombineArgs<F> ca;
    ca.nthreads = nbuckets;
    ca.partials = w.partials.as<XYZZ<F>>();
    ca.seg_off = w.seg_off.as<uint32_t>();
    ca.buckets = w.buckets.as<XYZZ<F>>();
    launch_combine<F>(ca, s);
}

// Step 5: out[inst] = sum over the bucket sets of sum_b b * S_b (times the window weights).
template <class F>
void msm_reduce_buckets(const MsmClass& k, uint32_t n_inst, XYZZ<F>* out, MsmScratch& w, cudaStream_t s) {
    // reduction levels
    uint32_t jobs = n_inst * k.nsets;
    uint32_t n_w = k.nb + 1, n_p = k.n_ones, stride_in = k.set_stride;
    const XYZZ<F>* X = w.buckets.as<XYZZ<F>>();
    const XYZZ<F>* P = nullptr;
    int flip = 0;
    uint32_t level = 0, shift = 0;
    // fan-in: wide at level 0 (many chunks: throughput-bound), narrow above it (few threads:
    // the serial chain of 2 T additions per level is what the stream waits for)
    for (;;) {
        const uint32_t log_t = level == 0 ? RED_LOG_T0 : RED_LOG_T1, T = 1u << log_t;
        RedArgs<F> ra;
        ra.X = X;
        ra.P = P;
        ra.n_weighted = n_w;
        ra.n_plain = n_p;
        ra.stride_in = stride_in;
        ra.T = T;
        ra.n_wout = (n_w + T - 1) / T;
        ra.n_out = ra.n_wout + (n_p + T - 1) / T;
        ra.stride_out = ra.n_out;
        ra.shift = shift;
        shift += log_t;
        ra.level = level;
        ra.nthreads = (size_t)jobs * ra.n_out;
        w.lx[flip].ensure(ra.nthreads * sizeof(XYZZ<F>));
        w.lp[flip].ensure(ra.nthreads * sizeof(XYZZ<F>));
        ra.Xo = w.lx[flip].as<XYZZ<F>>();
        ra.Po = w.lp[flip].as<XYZZ<F>>();
        launch_red<F>(ra, s);
        X = ra.Xo;
        P = ra.Po;
        n_w = ra.n_out;
        n_p = 0;
        stride_in = ra.n_out;
        flip ^= 1;
        level++;
        if (ra.n_out == 1) break;
    }
    // P now holds one point per (instance, set)
    if (k.nsets == 1) {
        copy_d2d(out, P, (size_t)n_inst * sizeof(XYZZ<F>), s);
    } else {
        HornerArgs<F> ha;
        ha.nthreads = n_inst;
        ha.R = P;
        ha.r_stride = 1;
        ha.nsets = k.nsets;
        ha.c = k.c;
        ha.out = out;
        launch_horner<F>(ha, s);
    }
}

}  // namespace mb
