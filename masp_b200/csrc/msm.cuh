// Pippenger bucket MSM over BLS12-381 G1 / G2, batched over independent
// instances (one instance = one proof's query, or one standalone MSM).
//
// Replaces bellperson's `multiexp` for the H, L, A, B1 (G1) and B2 (G2)
// queries (call sites of create_random_proof at
// masp_proofs/src/sapling/prover.rs:116-117, 201-202, 251-252).
//
// Schedule (B200-first, not the reference's window-per-task loop):
//   1. msm_count    one thread per (instance, base): the scalar is folded to
//                   min(s, r - s) (sign carried by the entries), cut into signed
//                   c-bit digits -> histogram over buckets.  0 is skipped; 1 goes
//                   to a spread set of "ones" buckets (the reference's 0/1 fast
//                   path) so no bucket is a hot spot.
//   2. scan         exclusive prefix sum of the histogram (block-cooperative tiles).
//   3. msm_scatter  same walk, writes (table index, sign) entries bucket-sorted.
//   3b. seg / order buckets are cut into segments of <= 128 entries (tasks), tasks
//                   are counting-sorted longest first.
//   4. msm_accumulate  one thread per task: XYZZ accumulator in registers, affine
//                   table points gathered from HBM/L2, mixed add with 10 products
//                   and 9 Montgomery reductions (ec.cuh madd_y3); a bucket with a
//                   single segment is written directly, split ones are folded by
//                   msm_combine_round (tree of fan-in 8).
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

// Windows needed for a scalar in signed c-bit digits.  The digit walk first folds a scalar above
// (r - 1) / 2 to r - s and flips the sign of all of its entries, so every scalar it decomposes is at
// most (r - 1) / 2 < 2^254: ceil(254 / c) windows cover every bit, and the top digit never borrows
// from a further window as long as its largest value plus the carry from below stays <= 2^(c-1)
// (checked per c; otherwise one more window).  c = 17 -> 15 windows, c = 15 -> 17, where the unfolded
// range 0 .. r - 1 needs 16 and 18: one bucket addition per full-width scalar less.
inline uint32_t msm_nwin(uint32_t c) {
    uint32_t h[8];  // (r - 1) / 2
    for (int i = 0; i < 8; ++i) h[i] = FrCfg::half(i);
    uint32_t nwin = (254 + c - 1) / c;
    uint32_t lo = (nwin - 1) * c;  // first bit of the top window
    uint64_t top = 0;              // largest top digit: ((r - 1) / 2) >> lo
    for (int i = 7; i >= 0; --i) {
        if ((uint32_t)(32 * i + 31) < lo) break;
        if ((uint32_t)(32 * i) >= lo) top |= (uint64_t)h[i] << (32 * i - lo);
        else top |= (uint64_t)(h[i] >> (lo - 32 * i));
    }
    if (top + 1 > (1ull << (c - 1))) ++nwin;
    return nwin;
}

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

// one window: raw digit + carry -> signed digit, borrow for the next window, entry
template <int MODE>
MB_HD void digit_window(const DigitArgs& a, const MsmClass& k, uint32_t d, uint32_t w, uint32_t& carry, uint32_t flip,
                        size_t b0, uint32_t base) {
    d += carry;
    uint32_t neg = 0;
    carry = 0;
    if (d > k.nb) {  // use d - 2^c, borrow one from the next window
        d = (1u << k.c) - d;
        neg = 1;
        carry = 1;
    }
    neg ^= flip;
    if (d == 0) return;
    uint32_t set = k.precomp ? 0 : w;
    uint32_t tw = k.precomp ? w : 0;
    digit_emit<MODE>(a, b0 + (size_t)set * k.set_stride + d, ((tw * k.n_bases + base) << 1) | neg);
}

template <int MODE>
MB_HD void digit_body(const DigitArgs& a, size_t tid) {
    const MsmClass& k = a.k;
    size_t inst = tid / k.n_bases;
    uint32_t base = (uint32_t)(tid - inst * k.n_bases);
    const uint32_t* sp = a.pool + (inst * a.pool_stride + (k.sel ? k.sel[base] : base)) * 8;
    uint32_t s[8];
    uint32_t any_hi = 0;
#ifdef MB200_EMU
    for (int i = 0; i < 8; ++i) s[i] = sp[i];
#else
    {   // pool entries are 32-byte aligned: two 16-byte loads
        const uint4 lo4 = reinterpret_cast<const uint4*>(sp)[0], hi4 = reinterpret_cast<const uint4*>(sp)[1];
        s[0] = lo4.x; s[1] = lo4.y; s[2] = lo4.z; s[3] = lo4.w;
        s[4] = hi4.x; s[5] = hi4.y; s[6] = hi4.z; s[7] = hi4.w;
    }
#endif
    MB_UNROLL
    for (int i = 1; i < 8; ++i) any_hi |= s[i];
    if (any_hi == 0 && s[0] == 0) return;
    size_t b0 = inst * k.inst_stride;
    // s > (r - 1) / 2: walk r - s and negate every entry (msm_nwin counts windows for the folded range)
    uint32_t flip = 0;
    {
        uint32_t gt = 0, decided = 0;
        MB_UNROLL
        for (int i = 7; i >= 0; --i) {
            uint32_t hi = FrCfg::half(i);
            uint32_t g = s[i] > hi ? 1u : 0u, l = s[i] < hi ? 1u : 0u;
            gt |= g & ~decided;
            decided |= g | l;
        }
        flip = gt & 1u;
        if (flip) {
            uint32_t borrow = 0;
            MB_UNROLL
            for (int i = 0; i < 8; ++i) {
                uint32_t m = FrCfg::mod(i);
                uint32_t d0 = m - s[i];
                uint32_t b1 = m < s[i] ? 1u : 0u;
                uint32_t d1 = d0 - borrow;
                uint32_t b2 = d0 < borrow ? 1u : 0u;
                s[i] = d1;
                borrow = b1 | b2;
            }
            any_hi = 0;
            MB_UNROLL
            for (int i = 1; i < 8; ++i) any_hi |= s[i];
        }
    }
    if (any_hi == 0 && s[0] == 1) {
        digit_emit<MODE>(a, b0 + k.nb + 1 + base % k.n_ones, (base << 1) | flip);
        return;
    }
    // windows are peeled off a 64-bit buffer refilled limb by limb: the limbs stay in registers
    // (indexing s[] by a run-time window position had put the scalar into local memory)
    uint32_t carry = 0, w = 0, nbits = 0;
    const uint32_t c = k.c, mask = (1u << c) - 1;
    uint64_t buf = 0;
    MB_UNROLL
    for (int i = 0; i < 8; ++i) {
        buf |= (uint64_t)s[i] << nbits;
        nbits += 32;
        while (nbits >= c && w < k.nwin) {
            digit_window<MODE>(a, k, (uint32_t)buf & mask, w, carry, flip, b0, base);
            buf >>= c;
            nbits -= c;
            ++w;
        }
    }
    for (; w < k.nwin; ++w) {  // the partly filled top window, then windows that only receive a carry
        digit_window<MODE>(a, k, (uint32_t)buf & mask, w, carry, flip, b0, base);
        buf = 0;
    }
}
MB_HD void digit_count_body(const DigitArgs& a, size_t tid) { digit_body<0>(a, tid); }
MB_HD void digit_scatter_body(const DigitArgs& a, size_t tid) { digit_body<1>(a, tid); }
MB_K_MSM_G1(msm_count, DigitArgs, digit_count_body, 256)
MB_K_MSM_G1(msm_scatter, DigitArgs, digit_scatter_body, 256)

// ---------------------------------------------------------------------------
// 2: exclusive scan of `counts` -> offsets and cursor.  Three block-cooperative kernels over
// tiles of SCAN_TILE elements: tile sums (coalesced 16-byte loads, shuffle reduction), an exclusive
// scan of the tile sums by one block, and the tile-local scan that writes offsets (and the scatter
// cursor) back with 16-byte stores.  (Round 1 gave every thread 512 consecutive elements: three
// launches of ~110 / 44 / 40 us for 2.1 M buckets, strided across the warp; this form moves the
// same 25 MB at streaming speed.)
// ---------------------------------------------------------------------------
struct ScanArgs {
    const uint32_t* counts;
    uint32_t* offsets;
    uint32_t* cursor;   // may be nullptr
    uint32_t* partial;  // one per tile
    size_t n;           // elements
};
static const uint32_t SCAN_THREADS = 256, SCAN_PER_THREAD = 8, SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;
inline size_t scan_tiles(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }
void launch_scan(const ScanArgs& a, cudaStream_t s);  // all three steps, in order, on s

#ifdef MB_DEFINE_MSM_G1
#ifdef MB200_EMU
void launch_scan(const ScanArgs& a, cudaStream_t) {
    uint32_t run = 0;
    for (size_t i = 0; i < a.n; ++i) {
        uint32_t v = a.counts[i];
        a.offsets[i] = run;
        if (a.cursor) a.cursor[i] = run;
        run += v;
    }
    ::mb::g_launches += 3;
}
#else
// this thread's SCAN_PER_THREAD consecutive elements (zero beyond n); buffers come from cudaMalloc,
// so element 8 k is 32-byte aligned
__device__ __forceinline__ void scan_load(const uint32_t* src, size_t n, size_t first, uint32_t v[SCAN_PER_THREAD]) {
    if (first + SCAN_PER_THREAD <= n) {
        uint4 x = *reinterpret_cast<const uint4*>(src + first), y = *reinterpret_cast<const uint4*>(src + first + 4);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
        for (uint32_t i = 0; i < SCAN_PER_THREAD; ++i) v[i] = first + i < n ? src[first + i] : 0u;
    }
}
// exclusive prefix of `mine` over the block's threads (thread order); *total: the block's sum
__device__ __forceinline__ uint32_t scan_block_exclusive(uint32_t mine, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (uint32_t w = 0; w < SCAN_THREADS / 32; ++w) {
        uint32_t ws = warp_sums[w];
        before += w < warp ? ws : 0u;
        all += ws;
    }
    __syncthreads();  // warp_sums may be reused by the caller's next tile
    *total = all;
    return before + incl - mine;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const ScanArgs a) {
    uint32_t v[SCAN_PER_THREAD];
    scan_load(a.counts, a.n, (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_PER_THREAD, v);
    uint32_t sum = 0, total;
#pragma unroll
    for (uint32_t i = 0; i < SCAN_PER_THREAD; ++i) sum += v[i];
    scan_block_exclusive(sum, &total);
    if (threadIdx.x == 0) a.partial[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets(const ScanArgs a, size_t ntiles) {
    uint32_t carry = 0;  // one block walks the tile sums, SCAN_TILE at a time
    for (size_t base = 0; base < ntiles; base += SCAN_TILE) {
        uint32_t v[SCAN_PER_THREAD];
        const size_t first = base + (size_t)threadIdx.x * SCAN_PER_THREAD;
        scan_load(a.partial, ntiles, first, v);
        uint32_t sum = 0, total;
#pragma unroll
        for (uint32_t i = 0; i < SCAN_PER_THREAD; ++i) sum += v[i];
        uint32_t run = carry + scan_block_exclusive(sum, &total);
#pragma unroll
        for (uint32_t i = 0; i < SCAN_PER_THREAD; ++i) {
            if (first + i < ntiles) a.partial[first + i] = run;
            run += v[i];
        }
        carry += total;
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_write(const ScanArgs a) {
    uint32_t v[SCAN_PER_THREAD];
    const size_t first = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_PER_THREAD;
    scan_load(a.counts, a.n, first, v);
    uint32_t sum = 0, total;
#pragma unroll
    for (uint32_t i = 0; i < SCAN_PER_THREAD; ++i) sum += v[i];
    uint32_t run = a.partial[blockIdx.x] + scan_block_exclusive(sum, &total);
    uint32_t o[SCAN_PER_THREAD];
#pragma unroll
    for (uint32_t i = 0; i < SCAN_PER_THREAD; ++i) {
        o[i] = run;
        run += v[i];
    }
    if (first + SCAN_PER_THREAD <= a.n) {
        const uint4 x = make_uint4(o[0], o[1], o[2], o[3]), y = make_uint4(o[4], o[5], o[6], o[7]);
        *reinterpret_cast<uint4*>(a.offsets + first) = x;
        *reinterpret_cast<uint4*>(a.offsets + first + 4) = y;
        if (a.cursor) {
            *reinterpret_cast<uint4*>(a.cursor + first) = x;
            *reinterpret_cast<uint4*>(a.cursor + first + 4) = y;
        }
    } else {
#pragma unroll
        for (uint32_t i = 0; i < SCAN_PER_THREAD; ++i)
            if (first + i < a.n) {
                a.offsets[first + i] = o[i];
                if (a.cursor) a.cursor[first + i] = o[i];
            }
    }
}
void launch_scan(const ScanArgs& a, cudaStream_t s) {
    if (!a.n) return;
    const size_t ntiles = scan_tiles(a.n);
    scan_tile_sums<<<(unsigned)ntiles, SCAN_THREADS, 0, s>>>(a);
    scan_tile_offsets<<<1, SCAN_THREADS, 0, s>>>(a, ntiles);
    scan_tile_write<<<(unsigned)ntiles, SCAN_THREADS, 0, s>>>(a);
    MB_CUDA(cudaGetLastError());
    ::mb::g_launches += 3;
}
#endif
#endif

// ---------------------------------------------------------------------------
// 3b: tasks.  A bucket's entry list is cut into segments of at most SEG_LEN
// entries; one thread accumulates one segment and the (few) partial sums of a
// split bucket are added afterwards.  No single bucket can serialise the
// launch: with merged windows the top, partly filled window concentrates its
// digits on a handful of buckets, and real witnesses repeat small values.
// Tasks are handed to threads in descending length (a counting sort), so the
// 32 lanes of a warp run equal trip counts and the long tasks start first.
// ---------------------------------------------------------------------------
static const uint32_t SEG_LEN = 128;  // 256: same rate; 64: -4 % (profiles/r02_ab_ntt_block_and_segments.jsonl)
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
// Almost every task of a proof-sized launch has the same length, so the 32 lanes of a warp hit the
// same bin: the lanes with equal bins elect a leader that adds their count once and hands out ranks
// (one L2 atomic per warp and bin instead of 32 serialised on one address).
MB_HD uint32_t order_bin_add(uint32_t* hist, uint32_t bin) {
#if defined(__CUDA_ARCH__)
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, bin);
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(&hist[bin], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
#else
    return MB_ATOMIC_ADD(&hist[bin], 1u);
#endif
}
MB_HD void order_hist_body(const OrderArgs& a, size_t tid) {
    if (tid >= *a.ntasks) return;
    order_bin_add(a.hist, SEG_LEN - a.task_len[tid]);
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
    uint32_t pos = order_bin_add(a.hist, SEG_LEN - a.task_len[tid]);
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
    XYZZ<F>* partials;      // per task: the partial sums of buckets that were cut into several segments
    XYZZ<F>* buckets;       // per bucket: a bucket with a single segment (almost all) is written here directly
    // ranges of one MSM (slabs of a standalone MSM, base ranges of the H+L query) share their buckets: a
    // bucket's first segment starts from the sum the earlier ranges left in `buckets` instead of the identity
    uint32_t add_into;
    const uint32_t* task_bucket;
    const uint32_t* seg_off;
    const uint32_t* nseg;
};
template <class F>
MB_HD void acc_body(const AccArgs<F>& a, size_t tid) {
    if (tid >= *a.ntasks) return;
    uint32_t t = a.order[tid];
    uint32_t n = a.task_len[t];
    XYZZ<F> acc = XYZZ<F>::inf();
    const uint32_t b = a.task_bucket[t];
    const bool single = a.nseg[b] == 1;
    if (a.add_into && (single || a.seg_off[b] == t)) acc = a.buckets[b];
    const uint32_t* e = a.entries + a.task_start[t];
    MB_NOUNROLL
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t ent = e[i];
        Affine<F> q = a.table[ent >> 1];
        xyzz_madd(acc, q, (ent & 1) != 0);
    }
    if (single) a.buckets[b] = acc;  // no combine step for this bucket
    else a.partials[t] = acc;
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
// (449).  Round 2, profiles/r02_ab_acc_register_caps.jsonl: __maxnreg__ 152 / 144 / 136 with one-warp blocks
// (13 / 14 / 15 resident warps, 24-80 B of spill): 523 / 519 / 526 against 531; one-warp blocks at the
// full register count: 525.  The plain kernel at 164 registers and 12 warps per SM stays.
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
    const uint32_t* nseg;
    XYZZ<F>* buckets;
};
template <class F>
MB_HD void combine_body(const CombineArgs<F>& a, size_t tid) {
    if (a.nseg[tid] > 1) a.buckets[tid] = a.partials[a.seg_off[tid]];  // single segments were written by the accumulate kernel
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
// INL: additions and doublings inlined (ptxas overlaps their independent multiplications: about half the
// latency of the out-of-line bodies per level, at ~10x the code).  The proving path keeps the out-of-line
// form -- its reductions hide under other chunks' kernels and a small i-cache footprint matters more
// there -- the standalone MSM, whose caller waits for exactly this chain, takes the inlined one.
template <class F, bool INL>
MB_HD void red_add(XYZZ<F>& acc, const XYZZ<F>& q) {
    if (INL) xyzz_add(acc, q);
    else xyzz_add_cold(acc, q);
}
template <class F, bool INL = false>
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
            red_add<F, INL>(run, X[j]);
            red_add<F, INL>(acc, run);
        }
        red_add<F, INL>(run, X[lo]);
        MB_NOUNROLL
        for (uint32_t d = 0; d < a.shift; ++d) acc = INL ? xyzz_dbl(acc) : xyzz_dbl_cold(acc);
        if (a.level > 0) {
            const XYZZ<F>* P = a.P + job * a.stride_in;
            MB_NOUNROLL
            for (uint32_t j = lo; j < hi; ++j) red_add<F, INL>(acc, P[j]);
        }
    } else {  // level 0 only: a chunk of ones buckets, plain sum
        uint32_t lo = a.n_weighted + (k - a.n_wout) * a.T, hi = lo + a.T;
        if (hi > a.n_weighted + a.n_plain) hi = a.n_weighted + a.n_plain;
        MB_NOUNROLL
        for (uint32_t j = lo; j < hi; ++j) red_add<F, INL>(acc, X[j]);
    }
    a.Xo[job * a.stride_out + k] = run;
    a.Po[job * a.stride_out + k] = acc;
}
MB_HD void red_g1_body(const RedArgs<Fp>& a, size_t tid) { red_body<Fp>(a, tid); }
MB_HD void red_g2_body(const RedArgs<Fp2>& a, size_t tid) { red_body<Fp2>(a, tid); }
MB_K_RED_G1(msm_reduce_g1, RedArgs<Fp>, red_g1_body, 64)
MB_K_RED_G2(msm_reduce_g2, RedArgs<Fp2>, red_g2_body, 32)
MB_HD void red_inl_g1_body(const RedArgs<Fp>& a, size_t tid) { red_body<Fp, true>(a, tid); }
MB_K_MSM_G1(msm_reduce_inl_g1, RedArgs<Fp>, red_inl_g1_body, 32)

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
    // ~255 dependent doublings on one thread: the whole standalone MSM waits for this chain, so the
    // doubling is inlined here (the kernel lives in the unit with inlined multiplications) and ptxas
    // overlaps the independent multiplications of one doubling (U^2 | X^2, then U V | X V | M^2 | V ZZ,
    // then W Y | W ZZZ | M (S - X3)): three multiplication latencies per doubling instead of nine.
    XYZZ<F> acc = R[(size_t)(a.nsets - 1) * a.r_stride];
    MB_NOUNROLL
    for (uint32_t w = a.nsets - 1; w-- > 0;) {
        MB_NOUNROLL
        for (uint32_t d = 0; d < a.c; ++d) acc = xyzz_dbl(acc);
        xyzz_add_cold(acc, R[(size_t)w * a.r_stride]);
    }
    a.out[tid] = acc;
}
MB_HD void horner_g1_body(const HornerArgs<Fp>& a, size_t tid) { horner_body<Fp>(a, tid); }
MB_HD void horner_g2_body(const HornerArgs<Fp2>& a, size_t tid) { horner_body<Fp2>(a, tid); }
MB_K_MSM_G1(msm_horner_g1, HornerArgs<Fp>, horner_g1_body, 32)
MB_K_MSM_G2(msm_horner_g2, HornerArgs<Fp2>, horner_g2_body, 32)

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
inline void launch_red(const RedArgs<F>& a, cudaStream_t s, bool latency);
template <>
inline void launch_red<Fp>(const RedArgs<Fp>& a, cudaStream_t s, bool latency) {
    if (latency) launch_msm_reduce_inl_g1(a, s);
    else launch_msm_reduce_g1(a, s);
}
template <>
inline void launch_red<Fp2>(const RedArgs<Fp2>& a, cudaStream_t s, bool) { launch_msm_reduce_g2(a, s); }
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
    size_t nchunks = scan_tiles(nbuckets);
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
    launch_scan(sa, s);

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
    launch_scan(sg, s);
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
    aa.buckets = w.buckets.as<XYZZ<F>>();
    aa.add_into = add_into ? 1 : 0;
    aa.task_bucket = w.task_bucket.as<uint32_t>();
    aa.seg_off = w.seg_off.as<uint32_t>();
    aa.nseg = w.nseg.as<uint32_t>();
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
    ca.nseg = w.nseg.as<uint32_t>();
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
        launch_red<F>(ra, s, !k.precomp);  // no tables = a standalone MSM: its caller waits for this chain
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
