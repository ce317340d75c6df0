// Micro-benchmark: can FP64-pipe field arithmetic (csrc/fpd.cuh) run beside the
// integer Montgomery path (csrc/field.cuh) and add to its throughput?
//   pipe_mix [chains_log2=20] [adds_per_chain=64]
// Prints one JSON line per configuration; every FP64/mixed result is compared
// word for word with the integer path's.
#include <cuda_runtime.h>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include "../../masp_b200/csrc/ec.cuh"
#include "fpd.cuh"
using namespace mb;
unsigned long long mb::g_launches;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

struct G1D { FpD x, y, zz, zzz; };

__device__ __forceinline__ bool madd_d(G1D& acc, const FpD& qx, const FpD& qy, double sgn) {
    FpD U2 = FpD::mul(qx, acc.zz);
    FpD S2 = FpD::mul(qy, acc.zzz);
    FpD P = FpD::sub(U2, acc.x);
    FpD R = FpD::ssub(sgn, S2, acc.y);
    double lo = d_add(P.v[0], -FpD::rnd24(P.v[0]));
    if (lo == 0.0 || lo == 21845.0 || lo == -21845.0) return false;
    FpD PP = FpD::sqr(P);
    FpD PPP = FpD::mul(P, PP);
    FpD Q = FpD::mul(acc.x, PP);
    FpD X3 = FpD::sub(FpD::sub(FpD::sqr(R), PPP), FpD::add(Q, Q));
    X3.reduce();
    FpD Y3 = FpD::sub(FpD::mul(R, FpD::sub(Q, X3)), FpD::mul(acc.y, PPP));
    Y3.reduce();
    acc.zz = FpD::mul(acc.zz, PP);
    acc.zzz = FpD::mul(acc.zzz, PPP);
    acc.x = X3;
    acc.y = Y3;
    return true;
}

__device__ __noinline__ void chain_int(const G1Affine* tab, uint32_t ntab, uint32_t chain, uint32_t K, G1XYZZ* out) {
    G1XYZZ acc = G1XYZZ::inf();
    uint32_t idx = chain * 2654435761u;
    #pragma unroll 1
    for (uint32_t i = 0; i < K; ++i) {
        idx = idx * 1664525u + 1013904223u;
        G1Affine q = tab[(idx >> 8) % ntab];
        xyzz_madd(acc, q, (idx >> 7) & 1);
    }
    out[chain] = acc;
}

__device__ __noinline__ void chain_f64(const G1Affine* tab, uint32_t ntab, uint32_t chain, uint32_t K, G1XYZZ* out) {
    G1D acc;
    bool inf = true;
    uint32_t idx = chain * 2654435761u;
    #pragma unroll 1
    for (uint32_t i = 0; i < K; ++i) {
        idx = idx * 1664525u + 1013904223u;
        G1Affine q = tab[(idx >> 8) % ntab];
        bool neg = (idx >> 7) & 1;
        if (q.is_inf()) continue;
        FpD qx = FpD::from_fp(q.x), qy = FpD::from_fp(q.y);
        if (inf) {
            qx.carry();
            qy.carry();
            acc.x = qx;
            if (neg) qy = FpD::sub(FpD::zero(), qy);
            acc.y = qy;
            acc.zz = FpD::one();
            acc.zzz = acc.zz;
            inf = false;
            continue;
        }
        if (!madd_d(acc, qx, qy, neg ? -1.0 : 1.0)) {
            // exceptional (or a 2^-22 false alarm): exact integer formulas
            G1XYZZ a = {FpD::to_fp(acc.x), FpD::to_fp(acc.y), FpD::to_fp(acc.zz), FpD::to_fp(acc.zzz)};
            xyzz_madd_cold(a, q, neg);
            if (a.is_inf()) { inf = true; continue; }
            acc.x = FpD::from_fp(a.x); acc.x.carry();
            acc.y = FpD::from_fp(a.y); acc.y.carry();
            acc.zz = FpD::from_fp(a.zz); acc.zz.carry();
            acc.zzz = FpD::from_fp(a.zzz); acc.zzz.carry();
        }
    }
    G1XYZZ r = G1XYZZ::inf();
    if (!inf) r = {FpD::to_fp(acc.x), FpD::to_fp(acc.y), FpD::to_fp(acc.zz), FpD::to_fp(acc.zzz)};
    out[chain] = r;
}

// warp roles: warp w of a block runs the FP64 path iff (w % PERIOD) < NFP
template <int BLOCK, int PERIOD, int NFP>
__global__ void __launch_bounds__(BLOCK, 1) mix_kernel(const G1Affine* tab, uint32_t ntab, uint32_t nchains, uint32_t K,
                                                       G1XYZZ* out, unsigned* counter, unsigned* done_by_role) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool fp = (int)(warp % PERIOD) < NFP;
    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        uint32_t chain = t * 32 + lane;
        if (t * 32 >= nchains) break;
        if (lane == 0) atomicAdd(done_by_role + (fp ? 1 : 0), 1u);
        if (chain < nchains) {
            if (NFP > 0 && fp) chain_f64(tab, ntab, chain, K, out);
            else if (NFP < PERIOD) chain_int(tab, ntab, chain, K, out);
        }
    }
}

// raw multiplier chains
__global__ void __launch_bounds__(256) mul_int_kernel(const Fp* in, Fp* out, int K) {
    size_t t = blockIdx.x * 256 + threadIdx.x;
    Fp a = in[t], b = in[t + 1];
    #pragma unroll 1
    for (int i = 0; i < K; ++i) { a = Fp::mul(a, b); b = Fp::mul(b, a); }
    out[t] = Fp::add(a, b);
}
__global__ void __launch_bounds__(256) mul_f64_kernel(const Fp* in, Fp* out, int K) {
    size_t t = blockIdx.x * 256 + threadIdx.x;
    FpD a = FpD::from_fp(in[t]), b = FpD::from_fp(in[t + 1]);
    a.carry(); b.carry();
    #pragma unroll 1
    for (int i = 0; i < K; ++i) { a = FpD::mul(a, b); b = FpD::mul(b, a); }
    out[t] = Fp::add(FpD::to_fp(a), FpD::to_fp(b));
}
// raw pipe rates: 8 independent chains per thread
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int K) {
    double x[8];
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x + j;
    double m = 1.0000001, c = 0.5;
    #pragma unroll 1
    for (int i = 0; i < K; ++i) {
        #pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __fma_rn(x[j], m, c);
    }
    double s = 0;
    for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) imadw_kernel(unsigned long long* out, int K, uint32_t m) {
    unsigned long long x[8];
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x + j;
    #pragma unroll 1
    for (int i = 0; i < K; ++i) {
        #pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (unsigned long long)(uint32_t)x[j] * m + x[j];
    }
    unsigned long long s = 0;
    for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
// both raw chains in the same thread / in alternate warps
__global__ void __launch_bounds__(256) both_kernel(double* out, int K, uint32_t m, int mode) {
    double x[8];
    unsigned long long y[8];
    for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x + j; y[j] = threadIdx.x + j; }
    double dm = 1.0000001, c = 0.5;
    bool do_f = mode == 0 || ((threadIdx.x >> 5) & 1), do_i = mode == 0 || !((threadIdx.x >> 5) & 1);
    #pragma unroll 1
    for (int i = 0; i < K; ++i) {
        if (do_f) {
            #pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __fma_rn(x[j], dm, c);
        }
        if (do_i) {
            #pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = (unsigned long long)(uint32_t)y[j] * m + y[j];
        }
    }
    double s = 0;
    for (int j = 0; j < 8; ++j) s += x[j] + (double)y[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

static float timed(void (*fn)(void*), void* ctx, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    fn(ctx);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) fn(ctx);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

struct MixCtx {
    const G1Affine* tab; uint32_t ntab, nchains, K; G1XYZZ* out; unsigned* counter; unsigned* roles; int nsm; int which;
};
template <int BLOCK, int PERIOD, int NFP>
static void run_mix(void* p) {
    MixCtx* c = (MixCtx*)p;
    CK(cudaMemsetAsync(c->counter, 0, 4));
    CK(cudaMemsetAsync(c->roles, 0, 8));
    mix_kernel<BLOCK, PERIOD, NFP><<<c->nsm, BLOCK>>>(c->tab, c->ntab, c->nchains, c->K, c->out, c->counter, c->roles);
}

int main(int argc, char** argv) {
    int lg = argc > 1 ? atoi(argv[1]) : 19;
    uint32_t K = argc > 2 ? atoi(argv[2]) : 64;
    uint32_t nchains = 1u << lg;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int nsm = prop.multiProcessorCount;
    double clk = prop.clockRate * 1e3;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_hz\": %.0f}\n", prop.name, nsm, clk);

    // table of pseudo-random "points" (arbitrary field elements; the formulas do not care)
    uint32_t ntab = 1u << 16;
    std::vector<G1Affine> htab(ntab);
    uint64_t s = 0x9e3779b97f4a7c15ull;
    for (auto& q : htab) {
        for (int c = 0; c < 2; ++c) {
            Fp& f = c ? q.y : q.x;
            for (int i = 0; i < 12; ++i) { s = s * 6364136223846793005ull + 1442695040888963407ull; f.v[i] = (uint32_t)(s >> 32); }
            f.v[11] &= 0x0fffffffu;  // < p
        }
    }
    htab[5] = G1Affine::inf();
    htab[7] = htab[6];  // repeated point
    G1Affine* tab; CK(cudaMalloc(&tab, ntab * sizeof(G1Affine)));
    CK(cudaMemcpy(tab, htab.data(), ntab * sizeof(G1Affine), cudaMemcpyHostToDevice));
    G1XYZZ *out_ref, *out; CK(cudaMalloc(&out_ref, nchains * sizeof(G1XYZZ))); CK(cudaMalloc(&out, nchains * sizeof(G1XYZZ)));
    unsigned *counter, *roles; CK(cudaMalloc(&counter, 4)); CK(cudaMalloc(&roles, 8));

    // ---- raw pipes ----
    {
        double* d; CK(cudaMalloc(&d, (size_t)nsm * 8 * 256 * 8));
        int Kp = 4096;
        struct C { double* d; int K; int nsm; int mode; } c = {d, Kp, nsm, 0};
        float ms = timed([](void* p) { C* c = (C*)p; dfma_kernel<<<c->nsm * 8, 256>>>(c->d, c->K); }, &c, 5);
        double ops = (double)nsm * 8 * 256 * 8 * Kp;
        printf("{\"raw\": \"dfma\", \"ms\": %.3f, \"per_sm_per_clk\": %.2f}\n", ms, ops / (ms * 1e-3) / nsm / clk);
        ms = timed([](void* p) { C* c = (C*)p; imadw_kernel<<<c->nsm * 8, 256>>>((unsigned long long*)c->d, c->K, 0x9e3779b9u); }, &c, 5);
        printf("{\"raw\": \"imad.wide\", \"ms\": %.3f, \"per_sm_per_clk\": %.2f}\n", ms, ops / (ms * 1e-3) / nsm / clk);
        ms = timed([](void* p) { C* c = (C*)p; both_kernel<<<c->nsm * 8, 256>>>(c->d, c->K, 0x9e3779b9u, 0); }, &c, 5);
        printf("{\"raw\": \"dfma+imad.wide same thread\", \"ms\": %.3f, \"each_per_sm_per_clk\": %.2f}\n", ms, ops / (ms * 1e-3) / nsm / clk);
        ms = timed([](void* p) { C* c = (C*)p; both_kernel<<<c->nsm * 8, 256>>>(c->d, c->K, 0x9e3779b9u, 1); }, &c, 5);
        printf("{\"raw\": \"dfma / imad.wide alternate warps\", \"ms\": %.3f, \"each_per_sm_per_clk\": %.2f}\n", ms, ops / 2 / (ms * 1e-3) / nsm / clk);
        CK(cudaFree(d));
    }
    // ---- multiplier chains ----
    {
        size_t nt = (size_t)nsm * 6 * 256;
        Fp *in, *o1, *o2; CK(cudaMalloc(&in, (nt + 1) * sizeof(Fp))); CK(cudaMalloc(&o1, nt * sizeof(Fp))); CK(cudaMalloc(&o2, nt * sizeof(Fp)));
        CK(cudaMemcpy(in, htab.data(), (nt + 1) * sizeof(Fp) < ntab * sizeof(G1Affine) ? (nt + 1) * sizeof(Fp) : ntab * sizeof(G1Affine), cudaMemcpyHostToDevice));
        int Km = 256;
        struct C { Fp* in; Fp* o; int K; int nsm; } c1 = {in, o1, Km, nsm}, c2 = {in, o2, Km, nsm};
        float ms1 = timed([](void* p) { C* c = (C*)p; mul_int_kernel<<<c->nsm * 6, 256>>>(c->in, c->o, c->K); }, &c1, 3);
        float ms2 = timed([](void* p) { C* c = (C*)p; mul_f64_kernel<<<c->nsm * 6, 256>>>(c->in, c->o, c->K); }, &c2, 3);
        std::vector<Fp> h1(nt), h2(nt);
        CK(cudaMemcpy(h1.data(), o1, nt * sizeof(Fp), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h2.data(), o2, nt * sizeof(Fp), cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < nt; ++i) if (memcmp(&h1[i], &h2[i], sizeof(Fp))) bad++;
        double muls = (double)nt * Km * 2;
        printf("{\"mul_chain\": \"int\", \"ms\": %.3f, \"fp_mul_per_s\": %.4g, \"clk_per_mul_per_sm\": %.2f}\n", ms1, muls / (ms1 * 1e-3), clk * nsm * ms1 * 1e-3 / muls);
        printf("{\"mul_chain\": \"f64\", \"ms\": %.3f, \"fp_mul_per_s\": %.4g, \"clk_per_mul_per_sm\": %.2f, \"mismatch\": %zu}\n", ms2, muls / (ms2 * 1e-3), clk * nsm * ms2 * 1e-3 / muls, bad);
    }
    // ---- point-addition chains with warp roles ----
    MixCtx c = {tab, ntab, nchains, K, out_ref, counter, roles, nsm, 0};
    std::vector<G1XYZZ> href(nchains), h(nchains);
    double adds = (double)nchains * K;
    float ms = timed(run_mix<384, 1, 0>, &c, 2);
    CK(cudaMemcpy(href.data(), out_ref, nchains * sizeof(G1XYZZ), cudaMemcpyDeviceToHost));
    printf("{\"mix\": \"int only, 12 warps/SM\", \"ms\": %.2f, \"adds_per_s\": %.4g}\n", ms, adds / (ms * 1e-3));
    c.out = out;
#define RUN(B, P, NF, label)                                                                        \
    {                                                                                               \
        CK(cudaMemset(out, 0xff, nchains * sizeof(G1XYZZ)));                                        \
        float m = timed(run_mix<B, P, NF>, &c, 2);                                                  \
        CK(cudaMemcpy(h.data(), out, nchains * sizeof(G1XYZZ), cudaMemcpyDeviceToHost));            \
        unsigned r[2];                                                                              \
        CK(cudaMemcpy(r, roles, 8, cudaMemcpyDeviceToHost));                                        \
        size_t bad = 0;                                                                             \
        for (uint32_t i = 0; i < nchains; ++i) if (memcmp(&h[i], &href[i], sizeof(G1XYZZ))) bad++;  \
        printf("{\"mix\": \"%s\", \"ms\": %.2f, \"adds_per_s\": %.4g, \"vs_int\": %.3f, \"fp64_share\": %.3f, \"mismatch\": %zu}\n", \
               label, m, adds / (m * 1e-3), ms / m, (double)r[1] / (r[0] + r[1]), bad);             \
        fflush(stdout);                                                                             \
    }
    RUN(256, 1, 0, "int only, 8 warps/SM")
    RUN(128, 1, 1, "f64 only, 4 warps/SM")
    RUN(256, 1, 1, "f64 only, 8 warps/SM")
    RUN(384, 1, 1, "f64 only, 12 warps/SM")
    RUN(384, 3, 1, "8 int + 4 f64 warps/SM")
    RUN(384, 2, 1, "6 int + 6 f64 warps/SM")
    RUN(384, 3, 2, "4 int + 8 f64 warps/SM")
    RUN(256, 2, 1, "4 int + 4 f64 warps/SM")
    RUN(512, 4, 1, "12 int + 4 f64 warps/SM")
    RUN(512, 2, 1, "8 int + 8 f64 warps/SM")
    return 0;
}
