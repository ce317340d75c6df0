// Radix-2 NTT over the BLS12-381 scalar field on bellman's evaluation domain,
// and the fused H-polynomial pipeline.
//
// Replaces EvaluationDomain::{ifft, coset_fft, mul_assign, sub_assign,
// divide_by_z_on_coset, icoset_fft} (SURVEY.md §8 a-3, Appendix A "H").  The
// domain is bellman's: omega = ROOT_OF_UNITY^(2^(32 - log m)),
// ROOT_OF_UNITY = 7^((r-1)/2^32), coset generator 7.
//
// Formulation: Stockham autosort, natural order in and out, no bit-reversal pass.  For the
// sizes the circuits use (2^12 ... 2^18) a transform is TWO shared-memory kernels
// (ntt_smem.cuh: 2 + 2 sweeps over HBM, the first kernel's contiguous runs stored with TMA bulk
// copies); other sizes run one radix-8 pass per launch (tail pass of radix 4 or 2), one thread
// per radix-point butterfly held in registers.  Measured on B200 (profiles/r02_variants_ab.jsonl):
// 492 -> 517 Spend proofs/s for the shared-memory form, 522 with the six-transform H pipeline
// below.  Data stay in PLAIN form
// end to end: with the twiddles stored in Montgomery form, montmul(x, w*R)
// = x*w, so no conversion pass is needed on either side.  The scalings are
// fused into the first / last pass of a transform:
//   ifft  1/m and coset g^i      -> in_scale of the following coset fft
//   a*b                          -> "pointwise" load of the final inverse
//   1/m and g^-i, - c(X), 1/Z(g) -> out_scale / sub / post_k of the final inverse's last store
#pragma once
#include "field.cuh"

namespace mb {


struct NttArgs {
    size_t nthreads;      // batch * (n >> K)
    const Fr* src;        // batch items, src_stride apart
    Fr* dst;
    size_t src_stride, dst_stride;
    uint32_t src_len;     // elements present per item; reads beyond are zero (first pass)
    uint32_t log_n;
    uint32_t ns;          // product of the radices of the previous passes
    uint32_t inverse;
    const Fr* tw;         // omega^i, Montgomery form, n entries
    const Fr* in_scale;   // optional, Montgomery form, indexed by source position
    const Fr* out_scale;  // optional, Montgomery form, indexed by destination position
    // pointwise load: x = a * b (plain product); b laid out as src
    const Fr* srcb;
    // sub: after out_scale, val = (val - sub[o] * k3) * k2 on the last store
    const Fr* sub;
    size_t sub_stride;
    Fr k2, k3;
};

template <int K>
MB_HD void ntt_body(const NttArgs& a, size_t tid) {
    constexpr int R = 1 << K;
    const uint32_t n = 1u << a.log_n;
    const uint32_t per = n >> K;
    size_t item = tid / per;
    uint32_t j = (uint32_t)(tid - item * per);
    uint32_t kk = j & (a.ns - 1);
    const Fr* src = a.src + item * a.src_stride;
    Fr v[R];
    MB_UNROLL
    for (int r = 0; r < R; ++r) {
        uint32_t idx = j + (uint32_t)r * per;
        if (idx < a.src_len) {
            v[r] = src[idx];
            if (a.srcb) v[r] = Fr::mul(Fr::mul(v[r], a.srcb[item * a.src_stride + idx]), Fr::r2());
            if (a.in_scale) v[r] = Fr::mul(v[r], a.in_scale[idx]);
        } else {
            v[r] = Fr::zero();
        }
    }
    if (kk != 0) {
        uint32_t step = kk * (n / (a.ns * R));
        MB_UNROLL
        for (int r = 1; r < R; ++r) {
            uint32_t idx = step * (uint32_t)r;
            if (a.inverse) idx = (n - idx) & (n - 1);
            v[r] = Fr::mul(v[r], a.tw[idx]);
        }
    }
    // radix-R DFT in registers: K decimation-in-frequency stages
    Fr w[R / 2 > 1 ? R / 2 : 1];
    MB_UNROLL
    for (int e = 1; e < R / 2; ++e) {
        uint32_t idx = (uint32_t)e * (n >> K);
        if (a.inverse) idx = n - idx;
        w[e] = a.tw[idx];
    }
    MB_UNROLL
    for (int span = R / 2; span >= 1; span >>= 1) {
        MB_UNROLL
        for (int start = 0; start < R; start += 2 * span) {
            MB_UNROLL
            for (int q = 0; q < span; ++q) {
                Fr x = v[start + q], y = v[start + q + span];
                v[start + q] = Fr::add(x, y);
                Fr d = Fr::sub(x, y);
                int e = q * ((R / 2) / span);
                v[start + q + span] = e == 0 ? d : Fr::mul(d, w[e]);
            }
        }
    }
    Fr* dst = a.dst + item * a.dst_stride;
    uint32_t j0 = (j / a.ns) * a.ns * R + kk;
    MB_UNROLL
    for (int i = 0; i < R; ++i) {
        int q = 0;  // bit reversal of i over K bits
        MB_UNROLL
        for (int b = 0; b < K; ++b) q |= ((i >> b) & 1) << (K - 1 - b);
        uint32_t o = j0 + (uint32_t)q * a.ns;
        Fr val = v[i];
        if (a.out_scale) val = Fr::mul(val, a.out_scale[o]);
        if (a.sub) val = Fr::mul(Fr::sub(val, Fr::mul(a.sub[item * a.sub_stride + o], a.k3)), a.k2);
        dst[o] = val;
    }
}
MB_HD void ntt1_body(const NttArgs& a, size_t tid) { ntt_body<1>(a, tid); }
MB_HD void ntt2_body(const NttArgs& a, size_t tid) { ntt_body<2>(a, tid); }
MB_HD void ntt3_body(const NttArgs& a, size_t tid) { ntt_body<3>(a, tid); }
MB_K_NTT(ntt_pass_r2, NttArgs, ntt1_body, 128)
MB_K_NTT(ntt_pass_r4, NttArgs, ntt2_body, 128)
// The radix-8 pass is latency-bound on its own (152 registers, 12 warps per SM, 45 % of the
// multiplier, 6 x the algorithmic HBM traffic: profiles/r01_ncu_ntt_pass_r8.md), which is why the
// circuits' sizes go through ntt_smem.cuh; this kernel serves the sizes outside 2^12 ... 2^18.
MB_K_NTT(ntt_pass_r8, NttArgs, ntt3_body, 128)

// ---------------------------------------------------------------------------
// domain tables
// ---------------------------------------------------------------------------
struct PowArgs {
    size_t nthreads;
    Fr* out;
    Fr base;   // Montgomery
    Fr scale;  // Montgomery
};
// out[i] = scale * base^i
MB_HD void pow_body(const PowArgs& a, size_t tid) {
    Fr acc = a.scale, b = a.base;
    size_t e = tid;
    while (e) {
        if (e & 1) acc = Fr::mul(acc, b);
        b = Fr::sqr(b);
        e >>= 1;
    }
    a.out[tid] = acc;
}
MB_K_NTT(fr_powers, PowArgs, pow_body, 128)

struct FrMulArgs {
    size_t nthreads;
    const Fr* a;
    const Fr* b;
    Fr* out;
};
// plain in, plain out
MB_HD void frmul_body(const FrMulArgs& a, size_t tid) { a.out[tid] = Fr::mul(Fr::mul(a.a[tid], a.b[tid]), Fr::r2()); }
MB_K_NTT(fr_mul_kernel, FrMulArgs, frmul_body, 128)

inline Fr fr_from_u64_host(uint64_t x) {
    Fr t = Fr::zero();
    t.v[0] = (uint32_t)x;
    t.v[1] = (uint32_t)(x >> 32);
    return Fr::from_std(t);
}
inline Fr fr_pow_host(Fr b, uint64_t e) {
    Fr acc = Fr::one();
    while (e) {
        if (e & 1) acc = Fr::mul(acc, b);
        b = Fr::sqr(b);
        e >>= 1;
    }
    return acc;
}

struct NttDomain {
    uint32_t log_n = 0;
    DevBuf tw, cos_fwd, cos_inv, minv_tab;
    Fr k2;         // 1 / (g^m - 1), Montgomery
    Fr minv;       // Montgomery

    void build(uint32_t logn, cudaStream_t s) {
        if (logn == 0 || logn > 28) fail(MB200_EINVAL, "NTT size 2^%s%ld out of range", "", (long)logn);
        log_n = logn;
        size_t n = (size_t)1 << logn;
        // ROOT_OF_UNITY = 7^((r-1)/2^32), Montgomery form
        Fr omega;
        const uint32_t root[8] = {0x5f0e466au, 0xb9b58d8cu, 0x1819d7ecu, 0x5b1b4c80u,
                                  0x52a31e64u, 0x0af53ae3u, 0x19e9b27bu, 0x5bf3addau};
        for (int i = 0; i < 8; ++i) omega.v[i] = root[i];
        for (uint32_t i = logn; i < 32; ++i) omega = Fr::sqr(omega);
        Fr g = fr_from_u64_host(7);
        Fr ginv = Fr::inv(g);
        minv = Fr::inv(fr_from_u64_host(n));
        Fr zinv = Fr::inv(Fr::sub(fr_pow_host(g, n), Fr::one()));  // 1 / (g^m - 1), Montgomery
        k2 = zinv;
        tw.alloc(n * sizeof(Fr));
        cos_fwd.alloc(n * sizeof(Fr));
        cos_inv.alloc(n * sizeof(Fr));
        minv_tab.alloc(n * sizeof(Fr));
        PowArgs pa;
        pa.nthreads = n;
        pa.out = tw.as<Fr>(); pa.base = omega; pa.scale = Fr::one();
        launch_fr_powers(pa, s);
        pa.out = cos_fwd.as<Fr>(); pa.base = g; pa.scale = minv;
        launch_fr_powers(pa, s);
        pa.out = cos_inv.as<Fr>(); pa.base = ginv; pa.scale = minv;
        launch_fr_powers(pa, s);
        pa.out = minv_tab.as<Fr>(); pa.base = Fr::one(); pa.scale = minv;
        launch_fr_powers(pa, s);
    }
};

struct NttPlan {
    const Fr* in_scale = nullptr;
    const Fr* out_scale = nullptr;
    const Fr* srcb = nullptr;  // plain product with srcb on load
    uint32_t src_len = 0;      // 0 = n
    bool inverse = false;
    // after out_scale, val = (val - sub[o] * sub_k) * post_k on the last store
    const Fr* sub = nullptr;
    size_t sub_stride = 0;
    Fr sub_k{}, post_k{};
};

}  // namespace mb
#include "ntt_smem.cuh"
namespace mb {

// One transform of `batch` items: src -> dst through the two scratch buffers
// (each batch * n elements, stride n).  src / dst strides are free.
inline void ntt_run(const NttDomain& d, const NttPlan& p, uint32_t batch, const Fr* src, size_t src_stride, Fr* dst,
                    size_t dst_stride, Fr* tmp0, Fr* tmp1, cudaStream_t s) {
    uint32_t n = 1u << d.log_n;
    uint32_t left = d.log_n, ns = 1;
    if (ntt_smem_supported(d.log_n)) {
        // two shared-memory kernels (ntt_smem.cuh) instead of the pass loop below
        NttFusedArgs f;
        f.log_n = d.log_n;
        f.inverse = p.inverse ? 1 : 0;
        f.tw = d.tw.as<Fr>();
        f.k2 = d.k2;
        f.sub = nullptr;
        f.sub_stride = 0;
        f.k3 = d.k2;
        f.group = 0;
        f.L = NTT_SMEM_L1;
        f.logC = NTT_SMEM_LOG_ELEMS - f.L;
        f.nblocks = (size_t)batch * ((n >> f.L) >> f.logC);
        f.src = src;
        f.src_stride = src_stride;
        f.dst = tmp0;
        f.dst_stride = n;
        f.src_len = p.src_len ? p.src_len : n;
        f.in_scale = p.in_scale;
        f.srcb = p.srcb;
        f.out_scale = nullptr;
        launch_ntt_fused(f, s);
        f.group = 1;
        f.L = d.log_n - NTT_SMEM_L1;
        f.logC = NTT_SMEM_LOG_ELEMS - f.L;
        f.nblocks = (size_t)batch * ((n >> f.L) >> f.logC);
        f.src = tmp0;
        f.src_stride = n;
        f.dst = dst;
        f.dst_stride = dst_stride;
        f.src_len = n;
        f.in_scale = nullptr;
        f.srcb = nullptr;
        f.out_scale = p.out_scale;
        if (p.sub) {
            f.sub = p.sub;
            f.sub_stride = p.sub_stride;
            f.k3 = p.sub_k;
            f.k2 = p.post_k;
        }
        launch_ntt_fused(f, s);
        return;
    }
    const Fr* cur = src;
    size_t cur_stride = src_stride;
    int flip = 0;
    bool first = true;
    while (left) {
        uint32_t K = left >= 3 ? 3 : left;
        bool last = left == K;
        NttArgs a;
        a.nthreads = (size_t)batch * (n >> K);
        a.src = cur;
        a.src_stride = cur_stride;
        a.dst = last ? dst : (flip ? tmp1 : tmp0);
        a.dst_stride = last ? dst_stride : n;
        a.src_len = first && p.src_len ? p.src_len : n;
        a.log_n = d.log_n;
        a.ns = ns;
        a.inverse = p.inverse ? 1 : 0;
        a.tw = d.tw.as<Fr>();
        a.in_scale = first ? p.in_scale : nullptr;
        a.out_scale = last ? p.out_scale : nullptr;
        a.srcb = first ? p.srcb : nullptr;
        a.k2 = p.post_k;
        a.sub = last ? p.sub : nullptr;
        a.sub_stride = p.sub_stride;
        a.k3 = p.sub_k;
        if (K == 3) launch_ntt_pass_r8(a, s);
        else if (K == 2) launch_ntt_pass_r4(a, s);
        else launch_ntt_pass_r2(a, s);
        cur = a.dst;
        cur_stride = a.dst_stride;
        flip ^= 1;
        ns <<= K;
        left -= K;
        first = false;
    }
}

// H pipeline for `batch` proofs.  abc: [batch][3][rows] plain scalars (a, b, c
// evaluation vectors, rows <= m each, polynomials `poly_stride` apart).  Writes
// the m coefficients of each proof's quotient to hout + proof * hout_stride (the
// caller ignores coefficient m-1, as bellman truncates it).
// work0..work3: scratch of batch * 3 * m elements each.
//
// SIX transforms where bellman runs seven (3 ifft, 3 coset_fft, 1 icoset_fft): the inverse coset
// transform is linear, so
//   icoset[(a b - c)(g w^i) / (g^m - 1)] = (icoset[(a b)(g w^i)] - c(X)) / (g^m - 1)
// coefficient by coefficient, and c(X) is already there after step 1: the coset transform of c
// is never needed.  Same field elements as the seven-transform form for ANY rows (satisfied or
// not), hence the same proof bytes (tests/test_emu.py, tests/test_gpu_parity.py::test_h_*).
inline void h_pipeline(const NttDomain& d, uint32_t batch, uint32_t rows, const Fr* abc, size_t poly_stride,
                       Fr* hout, size_t hout_stride, Fr* work0, Fr* work1, Fr* work2, Fr* work3, cudaStream_t s) {
    uint32_t n = 1u << d.log_n;
    // 1. three inverse transforms per proof (unscaled), zero-padded from `rows`
    NttPlan p1;
    p1.inverse = true;
    p1.src_len = rows;
    ntt_run(d, p1, batch * 3, abc, poly_stride, work2, n, work0, work1, s);
    // 2. coset forward transforms of a and b; 1/m and g^i folded into the load
    NttPlan q2;
    q2.in_scale = d.cos_fwd.as<Fr>();
    for (uint32_t poly = 0; poly < 2; ++poly)  // polynomials of a proof sit n apart
        ntt_run(d, q2, batch, work2 + (size_t)poly * n, 3 * (size_t)n, work3 + (size_t)poly * n, 3 * (size_t)n, work0,
                work1, s);
    // 3. a * b on load, inverse transform; g^-i / m, - c(X) and 1 / (g^m - 1) on the last store
    NttPlan q3;
    q3.inverse = true;
    q3.srcb = work3 + n;
    q3.out_scale = d.cos_inv.as<Fr>();
    q3.sub = work2 + 2 * (size_t)n;  // raw inverse transform of c: still lacks its 1/m
    q3.sub_stride = 3 * (size_t)n;
    q3.sub_k = d.minv;
    q3.post_k = d.k2;
    ntt_run(d, q3, batch, work3, 3 * (size_t)n, hout, hout_stride, work0, work1, s);
}

}  // namespace mb
