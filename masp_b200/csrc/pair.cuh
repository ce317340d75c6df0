// Batched-affine pair rounds in front of the bucket accumulation.
//
// A bucket holding k points needs k - 1 additions whatever the schedule.  The
// register-resident XYZZ accumulator (msm_accumulate) pays 10 field
// multiplications for each.  An affine addition costs 6 -- one of them the
// division -- if the division is shared: here one round adds the points of
// every bucket in PAIRS (entry 2i with entry 2i+1, an odd one is copied), all
// pairs of the launch are independent, and ONE field inversion serves the whole
// launch (Montgomery's trick, three levels deep):
//   pair_a     each thread walks B consecutive output slots forward, multiplies
//              its denominators together, stores the running prefixes (HBM) and
//              its product;
//   binv_*     the thread products are inverted together: fan-in 64 product
//              tree up, one Fermat inversion at the top (a single thread; the
//              only latency-bound step), back-substitution down;
//   pair_b     each thread walks its slots backward, peels one inverse per slot
//              and writes the affine sum.
// In SIMT a warp-shared inversion would not do: a warp executing the ~570
// multiplications of an inversion costs the same issue slots whether one lane
// or all 32 need the result (measured: profiles/r01_pair_rounds_v1_sweep.jsonl),
// so the inversion has to be amortised over the launch, not over the warp.
// A round halves every bucket; after R rounds the remaining points (already
// affine, signs applied) go to the XYZZ accumulator.  With R = 3, 7/8 of all
// additions cost ~6 multiplications instead of 10.
//
// Layout: round r reads bucket b's points at in[off_r[b] .. off_r[b] + cnt_r[b])
// and writes its ceil(cnt/2) results at out[off_(r+1)[b] ..), with the closed form
//     off_(r+1)[b] = ceil((off_r[b] + b) / 2),     cnt_(r+1)[b] = ceil(cnt_r[b] / 2),
// which never overlaps (ceil(x + m/2) - ceil(x) >= floor(m/2)) and needs no scan.
// Intermediate points live in two ping-pong buffers in HBM: this is the part of
// the path that trades the idle HBM bandwidth for multiplier issue slots.
//
// Exact for every input: P + P takes the tangent slope, P + (-P) and infinities
// are resolved without touching the shared inversion (their denominator is 1).
#pragma once
#include "ec.cuh"

namespace mb {

template <class F>
struct PairArgs {
    uint32_t nbuckets;
    const uint32_t* off_in;   // [nbuckets]
    const uint32_t* cnt_in;   // [nbuckets]
    uint32_t* off_out;        // [nbuckets + 1], last = end of the output index space
    uint32_t* cnt_out;        // [nbuckets]
    const Affine<F>* table;   // round 0: points are table[entries[.] >> 1], negated if the low bit is set
    const uint32_t* entries;  // round 0 only (nullptr afterwards)
    const Affine<F>* win;     // rounds >= 1
    Affine<F>* wout;
    size_t nthreads;          // layout kernel: nbuckets + 1; pair_a / pair_b: worker threads T
    uint32_t B;               // output slots per worker thread
    F* gpre;                  // [B][T] running prefix products (slot-major: coalesced across lanes)
    F* gprod;                 // [T] product of each thread's denominators
    const F* ginv;            // [T] their inverses (after binv_*)
    uint32_t* glast;          // [T] bucket of each thread's last slot
};

// per-bucket layout of the next round
template <class F>
MB_HD void pair_layout_body(const PairArgs<F>& a, size_t tid) {
    uint32_t nb = a.nbuckets;
    if (tid < nb) {
        a.off_out[tid] = (uint32_t)(((size_t)a.off_in[tid] + tid + 1) >> 1);
        a.cnt_out[tid] = (a.cnt_in[tid] + 1) >> 1;
    } else if (tid == nb) {
        size_t end_in = nb ? (size_t)a.off_in[nb - 1] + a.cnt_in[nb - 1] : 0;
        a.off_out[nb] = (uint32_t)((end_in + nb + 1) >> 1);
    }
}
MB_HD void pair_layout_g1_body(const PairArgs<Fp>& a, size_t tid) { pair_layout_body<Fp>(a, tid); }
MB_HD void pair_layout_g2_body(const PairArgs<Fp2>& a, size_t tid) { pair_layout_body<Fp2>(a, tid); }
MB_K_MSM_G1(pair_layout_g1, PairArgs<Fp>, pair_layout_g1_body, 256)
MB_K_MSM_G2(pair_layout_g2, PairArgs<Fp2>, pair_layout_g2_body, 256)

template <class F>
MB_HD Affine<F> pair_in(const PairArgs<F>& a, uint32_t b, uint32_t p) {
    size_t idx = (size_t)a.off_in[b] + p;
    if (a.entries) {
        uint32_t e = a.entries[idx];
        Affine<F> q = a.table[e >> 1];
        if (e & 1) q.y = F::neg(q.y);
        return q;
    }
    return a.win[idx];
}

enum : int { PAIR_NONE = 0, PAIR_COPY, PAIR_ADD, PAIR_DBL, PAIR_INF, PAIR_FIRST, PAIR_SECOND };

// what to do with (p1, p2), and the denominator it contributes to the shared inversion
template <class F>
MB_HD int pair_classify(const Affine<F>& p1, const Affine<F>& p2, F& den) {
    bool i1 = p1.is_inf(), i2 = p2.is_inf();
    if (i1 || i2) return i1 ? (i2 ? PAIR_INF : PAIR_SECOND) : PAIR_FIRST;
    den = F::sub(p2.x, p1.x);
    if (!den.is_zero()) return PAIR_ADD;
    if (p1.y.eq(p2.y) && !p1.y.is_zero()) {
        den = F::dbl(p1.y);
        return PAIR_DBL;
    }
    return PAIR_INF;  // P + (-P), or a 2-torsion point doubled
}

// phase A: walk the B slots forward, collect denominators
template <class F>
MB_HD void pair_a_body(const PairArgs<F>& a, size_t t) {
    const uint32_t nb = a.nbuckets, B = a.B;
    const size_t end = a.off_out[nb], T = a.nthreads;
    size_t j0 = t * (size_t)B;
    F prod = F::one();
    uint32_t b = 0;
    if (j0 < end) {
        uint32_t lo = 0, hi = nb;  // largest b with off_out[b] <= j0
        while (hi - lo > 1) {
            uint32_t mid = lo + ((hi - lo) >> 1);
            if ((size_t)a.off_out[mid] <= j0) lo = mid;
            else hi = mid;
        }
        b = lo;
        uint32_t n = end - j0 < (size_t)B ? (uint32_t)(end - j0) : B;
        MB_NOUNROLL
        for (uint32_t s = 0; s < n; ++s) {
            size_t j = j0 + s;
            while (b + 1 < nb && j >= (size_t)a.off_out[b + 1]) ++b;
            uint32_t i = (uint32_t)(j - a.off_out[b]), k = a.cnt_in[b];
            if (2 * i + 1 < k) {
                Affine<F> p1 = pair_in(a, b, 2 * i), p2 = pair_in(a, b, 2 * i + 1);
                F den;
                int ty = pair_classify(p1, p2, den);
                if (ty == PAIR_ADD || ty == PAIR_DBL) {
                    a.gpre[(size_t)s * T + t] = prod;
                    prod = F::mul(prod, den);
                }
            }
        }
    }
    a.gprod[t] = prod;
    a.glast[t] = b;
}

// phase B: walk backward, peel one inverse per slot, add
template <class F>
MB_HD void pair_b_body(const PairArgs<F>& a, size_t t) {
    const uint32_t nb = a.nbuckets, B = a.B;
    const size_t end = a.off_out[nb], T = a.nthreads;
    size_t j0 = t * (size_t)B;
    if (j0 >= end) return;
    uint32_t n = end - j0 < (size_t)B ? (uint32_t)(end - j0) : B;
    uint32_t b = a.glast[t];
    F acc = a.ginv[t];  // inverse of the product of the denominators of slots <= s
    MB_NOUNROLL
    for (uint32_t s = n; s-- > 0;) {
        size_t j = j0 + s;
        while (j < (size_t)a.off_out[b]) --b;
        uint32_t i = (uint32_t)(j - a.off_out[b]), k = a.cnt_in[b];
        Affine<F>* out = a.wout + j;
        if (2 * i + 1 < k) {
            Affine<F> p1 = pair_in(a, b, 2 * i), p2 = pair_in(a, b, 2 * i + 1);
            F den;
            int ty = pair_classify(p1, p2, den);
            if (ty == PAIR_ADD || ty == PAIR_DBL) {
                F inv = F::mul(acc, a.gpre[(size_t)s * T + t]);
                acc = F::mul(acc, den);
                F num;
                if (ty == PAIR_ADD) {
                    num = F::sub(p2.y, p1.y);
                } else {
                    F xx = F::sqr(p1.x);
                    num = F::add(F::dbl(xx), xx);
                }
                F lam = F::mul(num, inv);
                F x3 = F::sub(F::sub(F::sqr(lam), p1.x), p2.x);
                F y3 = F::sub(F::mul(lam, F::sub(p1.x, x3)), p1.y);
                *out = {x3, y3};
            } else if (ty == PAIR_FIRST) {
                *out = p1;
            } else if (ty == PAIR_SECOND) {
                *out = p2;
            } else {
                *out = Affine<F>::inf();
            }
        } else if (2 * i < k) {
            *out = pair_in(a, b, 2 * i);  // odd one out
        }
    }
}
MB_HD void pair_a_g1_body(const PairArgs<Fp>& a, size_t t) { pair_a_body<Fp>(a, t); }
MB_HD void pair_b_g1_body(const PairArgs<Fp>& a, size_t t) { pair_b_body<Fp>(a, t); }
MB_HD void pair_a_g2_body(const PairArgs<Fp2>& a, size_t t) { pair_a_body<Fp2>(a, t); }
MB_HD void pair_b_g2_body(const PairArgs<Fp2>& a, size_t t) { pair_b_body<Fp2>(a, t); }
MB_K_MSM_G1(pair_a_g1, PairArgs<Fp>, pair_a_g1_body, 128)
MB_K_MSM_G1(pair_b_g1, PairArgs<Fp>, pair_b_g1_body, 128)
MB_K_MSM_G2(pair_a_g2, PairArgs<Fp2>, pair_a_g2_body, 64)
MB_K_MSM_G2(pair_b_g2, PairArgs<Fp2>, pair_b_g2_body, 64)

// ---------------------------------------------------------------------------
// batch inversion of a device array (all elements non-zero): product tree of
// fan-in G up, one inversion at the top, back-substitution down
// ---------------------------------------------------------------------------
template <class F>
struct BinvArgs {
    size_t nthreads;   // chunks at this level (top: 1)
    const F* x;        // level input, n elements
    F* pfx;            // exclusive prefix products inside each chunk
    F* up;             // chunk totals = next level's input (up kernel)
    const F* inv_up;   // inverses of the chunk totals (down kernel)
    F* inv;            // inverses of x (down / top kernel)
    size_t n;
    uint32_t G;
};
template <class F>
MB_HD void binv_up_body(const BinvArgs<F>& a, size_t u) {
    size_t lo = u * a.G, hi = lo + a.G < a.n ? lo + a.G : a.n;
    F run = F::one();
    MB_NOUNROLL
    for (size_t i = lo; i < hi; ++i) {
        a.pfx[i] = run;
        run = F::mul(run, a.x[i]);
    }
    a.up[u] = run;
}
template <class F>
MB_HD void binv_down_body(const BinvArgs<F>& a, size_t u) {
    size_t lo = u * a.G, hi = lo + a.G < a.n ? lo + a.G : a.n;
    F it = a.inv_up[u];
    MB_NOUNROLL
    for (size_t i = hi; i-- > lo;) {
        a.inv[i] = F::mul(it, a.pfx[i]);
        it = F::mul(it, a.x[i]);
    }
}
template <class F>
MB_HD void binv_top_body(const BinvArgs<F>& a, size_t) {
    F run = F::one();
    MB_NOUNROLL
    for (size_t i = 0; i < a.n; ++i) {
        a.pfx[i] = run;
        run = F::mul(run, a.x[i]);
    }
    F it = field_inv_cold(run);
    MB_NOUNROLL
    for (size_t i = a.n; i-- > 0;) {
        a.inv[i] = F::mul(it, a.pfx[i]);
        it = F::mul(it, a.x[i]);
    }
}
MB_HD void binv_up_g1_body(const BinvArgs<Fp>& a, size_t u) { binv_up_body<Fp>(a, u); }
MB_HD void binv_down_g1_body(const BinvArgs<Fp>& a, size_t u) { binv_down_body<Fp>(a, u); }
MB_HD void binv_top_g1_body(const BinvArgs<Fp>& a, size_t u) { binv_top_body<Fp>(a, u); }
MB_HD void binv_up_g2_body(const BinvArgs<Fp2>& a, size_t u) { binv_up_body<Fp2>(a, u); }
MB_HD void binv_down_g2_body(const BinvArgs<Fp2>& a, size_t u) { binv_down_body<Fp2>(a, u); }
MB_HD void binv_top_g2_body(const BinvArgs<Fp2>& a, size_t u) { binv_top_body<Fp2>(a, u); }
MB_K_MSM_G1(binv_up_g1, BinvArgs<Fp>, binv_up_g1_body, 64)
MB_K_MSM_G1(binv_down_g1, BinvArgs<Fp>, binv_down_g1_body, 64)
MB_K_MSM_G1(binv_top_g1, BinvArgs<Fp>, binv_top_g1_body, 32)
MB_K_MSM_G2(binv_up_g2, BinvArgs<Fp2>, binv_up_g2_body, 64)
MB_K_MSM_G2(binv_down_g2, BinvArgs<Fp2>, binv_down_g2_body, 64)
MB_K_MSM_G2(binv_top_g2, BinvArgs<Fp2>, binv_top_g2_body, 32)

template <class F> struct PairLaunch;
template <> struct PairLaunch<Fp> {
    static void layout(const PairArgs<Fp>& a, cudaStream_t s) { launch_pair_layout_g1(a, s); }
    static void pa(const PairArgs<Fp>& a, cudaStream_t s) { launch_pair_a_g1(a, s); }
    static void pb(const PairArgs<Fp>& a, cudaStream_t s) { launch_pair_b_g1(a, s); }
    static void up(const BinvArgs<Fp>& a, cudaStream_t s) { launch_binv_up_g1(a, s); }
    static void down(const BinvArgs<Fp>& a, cudaStream_t s) { launch_binv_down_g1(a, s); }
    static void top(const BinvArgs<Fp>& a, cudaStream_t s) { launch_binv_top_g1(a, s); }
};
template <> struct PairLaunch<Fp2> {
    static void layout(const PairArgs<Fp2>& a, cudaStream_t s) { launch_pair_layout_g2(a, s); }
    static void pa(const PairArgs<Fp2>& a, cudaStream_t s) { launch_pair_a_g2(a, s); }
    static void pb(const PairArgs<Fp2>& a, cudaStream_t s) { launch_pair_b_g2(a, s); }
    static void up(const BinvArgs<Fp2>& a, cudaStream_t s) { launch_binv_up_g2(a, s); }
    static void down(const BinvArgs<Fp2>& a, cudaStream_t s) { launch_binv_down_g2(a, s); }
    static void top(const BinvArgs<Fp2>& a, cudaStream_t s) { launch_binv_top_g2(a, s); }
};

static const uint32_t BINV_G = 64;
// elements of scratch (in F) that batch_inverse_device needs for n inputs: pfx + inv per level, totals
inline size_t binv_scratch_elems(size_t n) {
    size_t tot = 0;
    for (;;) {
        tot += 2 * n;  // pfx, inv of this level
        if (n <= BINV_G) break;
        n = (n + BINV_G - 1) / BINV_G;
        tot += n;      // the next level's x
    }
    return tot;
}
// inv_out[i] = 1 / x[i], i < n (device arrays; x is left intact); returns nothing, all on stream s
template <class F>
inline void batch_inverse_device(const F* x, size_t n, F* inv_out, F* scratch, cudaStream_t s) {
    struct Level { const F* x; F* pfx; F* inv; size_t n; };
    Level lv[8];
    int L = 0;
    F* p = scratch;
    const F* cur = x;
    size_t m = n;
    for (;;) {
        lv[L].x = cur;
        lv[L].n = m;
        lv[L].pfx = p;
        p += m;
        lv[L].inv = L == 0 ? inv_out : p;
        if (L != 0) p += m;
        if (m <= BINV_G) break;
        size_t nx = (m + BINV_G - 1) / BINV_G;
        BinvArgs<F> a;
        a.nthreads = nx;
        a.x = cur;
        a.pfx = lv[L].pfx;
        a.up = p;
        a.inv_up = nullptr;
        a.inv = nullptr;
        a.n = m;
        a.G = BINV_G;
        PairLaunch<F>::up(a, s);
        cur = p;
        p += nx;
        m = nx;
        ++L;
    }
    {
        BinvArgs<F> a;
        a.nthreads = 1;
        a.x = lv[L].x;
        a.pfx = lv[L].pfx;
        a.up = nullptr;
        a.inv_up = nullptr;
        a.inv = lv[L].inv;
        a.n = lv[L].n;
        a.G = BINV_G;
        PairLaunch<F>::top(a, s);
    }
    for (int l = L - 1; l >= 0; --l) {
        BinvArgs<F> a;
        a.nthreads = lv[l + 1].n;
        a.x = lv[l].x;
        a.pfx = lv[l].pfx;
        a.up = nullptr;
        a.inv_up = lv[l + 1].inv;
        a.inv = lv[l].inv;
        a.n = lv[l].n;
        a.G = BINV_G;
        PairLaunch<F>::down(a, s);
    }
}

}  // namespace mb
