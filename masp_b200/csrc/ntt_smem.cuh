// Shared-memory formulation of the Stockham NTT: two kernels instead of six passes.
//
// The default for 2^12 <= n <= 2^18 (every MASP circuit): measured on B200 at 492 -> 517 Spend
// proofs/s against the pass-per-launch kernels (profiles/r02_variants_ab.jsonl), bit-identical
// to them (tests/test_emu.py runs both forms on the same data).
//
// The first three radix-8 passes of the autosort transform of size n are, for every residue
// g mod n/512, one 512-point transform over the elements g + s * n/512, whose result lands in
// the contiguous run [512 g, 512 g + 512).  The remaining passes are, for every g < 512, one
// transform of size n/512 over the elements g + 512 u with twiddles that depend on g, landing
// in place.  So:
//
//   kernel 1 (group 0)  gathers C = 4 such 512-point problems (128-byte segments), runs three
//                       radix-8 passes in shared memory and writes four contiguous 16 KB runs,
//                       each as one TMA bulk copy (cp.async.bulk.global.shared::cta; SASS UBLKCP.G.S),
//                       the one contiguous tile on this path;
//   kernel 2 (group 1)  gathers C = 2048 / (n/512) problems of size n/512 (32 C-byte segments),
//                       runs the remaining passes (radix 8, 8 and a tail of 4 or 2) in shared
//                       memory and scatters with the same pattern.
//
// Global traffic per transform drops from 6 reads + 6 writes of the polynomial to 2 + 2.  A
// block is 256 threads on 72 KB of shared memory (2048 scalars plus one pad scalar per array); every pass is "each thread
// reads its 8 values, barrier, each thread writes its 8 results, barrier", so one buffer serves
// in place.  The fused scalings of ntt.cuh keep their places: zero padding / coset scale /
// (a b - c) / Z on the loads of kernel 1, the output scale on the stores of kernel 2.
//
// Host emulation (MB200_EMU): the same source with the thread loop written out
// (MB_FOR_THREADS) and the register file of a block as an array, barriers being the loop ends.
#pragma once

namespace mb {

struct NttFusedArgs {
    size_t nblocks;       // batch * (n >> L) / C
    const Fr* src;
    Fr* dst;
    size_t src_stride, dst_stride;
    uint32_t src_len, log_n, L, logC, group, inverse;
    const Fr* tw;
    const Fr* in_scale;
    const Fr* out_scale;
    const Fr* srcb;       // plain product a * b on load
    const Fr* sub;        // optional, group 1 store: val = (val - sub[idx] * k3) * k2
    size_t sub_stride;
    Fr k2, k3;
};

// 2048 scalars per block, 2 blocks per SM.  1024 (128 threads, 4 blocks per SM at the same 128 registers)
// measured 554 against 559 Spend proofs/s (profiles/r02_ab_ntt_block_and_segments.jsonl).
static const uint32_t NTT_SMEM_LOG_ELEMS = 11, NTT_SMEM_ELEMS = 1u << NTT_SMEM_LOG_ELEMS,
                      NTT_SMEM_THREADS = NTT_SMEM_ELEMS / 8, NTT_SMEM_L1 = 9;
// arrays sit Ln + 1 elements apart: consecutive arrays start 8 banks apart, so the gather / scatter phases
// (consecutive threads -> consecutive arrays) do not pile onto one bank group; at most 2048 / 8 arrays
static const uint32_t NTT_SMEM_ALLOC = NTT_SMEM_ELEMS + NTT_SMEM_ELEMS / 8;

#ifdef MB200_EMU
#define MB_FOR_THREADS(t) for (uint32_t t = 0; t < NTT_SMEM_THREADS; ++t)
#define MB_TSLOT(t) (t)
#define MB_TSLOTS NTT_SMEM_THREADS
#define MB_BLOCK_SYNC()
#define MB_BLOCK_FN inline
#else
#define MB_FOR_THREADS(t) for (uint32_t t = threadIdx.x, _mb_once = 1; _mb_once; _mb_once = 0)
#define MB_TSLOT(t) 0
#define MB_TSLOTS 1
#define MB_BLOCK_SYNC() __syncthreads()
#define MB_BLOCK_FN __device__ __forceinline__
#endif

// one radix-2^K work item of a local pass on values already in registers: twiddles, butterfly.
// v[i] come in as the elements j + i * per; they leave in DFT order q = bitrev(i).
template <int K>
MB_HD void ntt_fused_item(Fr* v, const NttFusedArgs& a, uint32_t kk_glob, uint32_t ns_glob) {
    constexpr int R = 1 << K;
    const uint32_t n = 1u << a.log_n;
    if (kk_glob != 0) {
        uint32_t step = kk_glob * (n / (ns_glob * R));
        MB_UNROLL
        for (int r = 1; r < R; ++r) {
            uint32_t idx = (step * (uint32_t)r) & (n - 1);
            if (a.inverse) idx = (n - idx) & (n - 1);
            v[r] = Fr::mul(v[r], a.tw[idx]);
        }
    }
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
}

// Element i of an array lives at slot i ^ ((i >> 3) & 7): a permutation inside every aligned group of eight.
// The first radix-8 pass writes 8 j + q for consecutive j (256 bytes apart: one bank group for the whole
// quarter-warp); swizzled, those eight lanes land on eight different slots.  Reads of consecutive elements
// and the strided writes of the later passes stay conflict-free under it.  `nat` = natural order (the layout
// the TMA bulk store needs for the last pass of kernel 1).
MB_HD uint32_t ntt_slot(uint32_t i, bool nat) { return nat ? i : (i ^ ((i >> 3) & 7u)); }

// one local pass over the block's C arrays of 2^L elements; every thread owns 8 values
template <int K>
MB_BLOCK_FN void ntt_fused_pass(const NttFusedArgs& a, Fr* sm, Fr (*regs)[8], uint32_t g0, uint32_t ns_loc,
                                bool write_nat) {
    constexpr int R = 1 << K;
    constexpr int ITEMS = 8 / R;               // work items per thread
    const uint32_t Ln = 1u << a.L, Lp = Ln + 1, per = Ln >> K, S = 1u << (a.log_n - a.L);
    MB_FOR_THREADS(t) {
        Fr* v = regs[MB_TSLOT(t)];
        MB_UNROLL
        for (int it = 0; it < ITEMS; ++it) {
            uint32_t w = t + (uint32_t)it * NTT_SMEM_THREADS;
            uint32_t c = w / per, j = w - c * per;
            MB_UNROLL
            for (int r = 0; r < R; ++r) v[it * R + r] = sm[c * Lp + ntt_slot(j + (uint32_t)r * per, false)];
        }
    }
    MB_BLOCK_SYNC();
    MB_FOR_THREADS(t) {
        Fr* v = regs[MB_TSLOT(t)];
        MB_UNROLL
        for (int it = 0; it < ITEMS; ++it) {
            uint32_t w = t + (uint32_t)it * NTT_SMEM_THREADS;
            uint32_t c = w / per, j = w - c * per;
            uint32_t kk = j & (ns_loc - 1);
            uint32_t kk_glob = a.group == 0 ? kk : (g0 + c) + S * kk;
            uint32_t ns_glob = a.group == 0 ? ns_loc : S * ns_loc;
            ntt_fused_item<K>(v + it * R, a, kk_glob, ns_glob);
            uint32_t j0 = (j / ns_loc) * ns_loc * R + kk;
            MB_UNROLL
            for (int i = 0; i < R; ++i) {
                int q = 0;
                MB_UNROLL
                for (int b = 0; b < K; ++b) q |= ((i >> b) & 1) << (K - 1 - b);
                sm[c * Lp + ntt_slot(j0 + (uint32_t)q * ns_loc, write_nat)] = v[it * R + i];
            }
        }
    }
    MB_BLOCK_SYNC();
}

MB_BLOCK_FN void ntt_fused_block(const NttFusedArgs& a, size_t blk, Fr* sm, Fr (*regs)[8]) {
    const uint32_t n = 1u << a.log_n, Ln = 1u << a.L, Lp = Ln + 1, C = 1u << a.logC, S = n >> a.L;
    const uint32_t blocks_per_item = S >> a.logC;
    const size_t item = blk / blocks_per_item;
    const uint32_t g0 = (uint32_t)(blk - item * blocks_per_item) << a.logC;
    const Fr* src = a.src + item * a.src_stride;
    Fr* dst = a.dst + item * a.dst_stride;
    // gather: element s of array c is src[g0 + c + s * S]; consecutive threads take consecutive c
    MB_FOR_THREADS(t) {
        MB_UNROLL
        for (int i = 0; i < 8; ++i) {
            uint32_t e = t + (uint32_t)i * NTT_SMEM_THREADS;
            uint32_t c = e & (C - 1), s = e >> a.logC;
            uint32_t idx = g0 + c + s * S;
            Fr x = Fr::zero();
            if (idx < a.src_len) {
                x = src[idx];
                if (a.srcb) x = Fr::mul(Fr::mul(x, a.srcb[item * a.src_stride + idx]), Fr::r2());
                if (a.in_scale) x = Fr::mul(x, a.in_scale[idx]);
            }
            sm[c * Lp + ntt_slot(s, false)] = x;
        }
    }
    MB_BLOCK_SYNC();
    // the last pass of kernel 1 leaves natural order: the TMA bulk store copies the arrays as they lie
    const bool nat_out = a.group == 0;
    uint32_t left = a.L, ns = 1;
    while (left >= 3) {
        ntt_fused_pass<3>(a, sm, regs, g0, ns, nat_out && left == 3);
        ns <<= 3;
        left -= 3;
    }
    if (left == 2) ntt_fused_pass<2>(a, sm, regs, g0, ns, nat_out);
    if (left == 1) ntt_fused_pass<1>(a, sm, regs, g0, ns, nat_out);
    if (a.group == 0) {
        // array c is the contiguous run dst[(g0 + c) * 2^L ...]
#ifndef MB200_EMU
        // TMA: one bulk copy per array, all issued by one thread; the generic-proxy writes of the
        // last pass are made visible to the async proxy first.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            MB_NOUNROLL
            for (uint32_t c = 0; c < C; ++c) {
                uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm + c * Lp);
                Fr* gptr = dst + (size_t)(g0 + c) * Ln;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gptr),
                             "r"(saddr), "r"(Ln * (uint32_t)sizeof(Fr))
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
#else
        MB_FOR_THREADS(t) {
            MB_UNROLL
            for (int i = 0; i < 8; ++i) {
                uint32_t e = t + (uint32_t)i * NTT_SMEM_THREADS;
                uint32_t c = e >> a.L, k = e & (Ln - 1);
                dst[(size_t)(g0 + c) * Ln + k] = sm[c * Lp + ntt_slot(k, nat_out)];
            }
        }
#endif
    } else {
        MB_FOR_THREADS(t) {
            MB_UNROLL
            for (int i = 0; i < 8; ++i) {
                uint32_t e = t + (uint32_t)i * NTT_SMEM_THREADS;
                uint32_t c = e & (C - 1), o = e >> a.logC;
                uint32_t idx = g0 + c + o * S;
                Fr val = sm[c * Lp + ntt_slot(o, false)];
                if (a.out_scale) val = Fr::mul(val, a.out_scale[idx]);
                if (a.sub) val = Fr::mul(Fr::sub(val, Fr::mul(a.sub[item * a.sub_stride + idx], a.k3)), a.k2);
                dst[idx] = val;
            }
        }
    }
}

#ifdef MB200_EMU
#ifdef MB_DEFINE_NTT
void launch_ntt_fused(const NttFusedArgs& a, cudaStream_t) {
    Fr* sm = new Fr[NTT_SMEM_ALLOC];
    Fr(*regs)[8] = new Fr[NTT_SMEM_THREADS][8];
    for (size_t b = 0; b < a.nblocks; ++b) ntt_fused_block(a, b, sm, regs);
    delete[] sm;
    delete[] regs;
    ::mb::g_launches++;
}
#else
void launch_ntt_fused(const NttFusedArgs& a, cudaStream_t);
#endif
#else
#ifdef MB_DEFINE_NTT
__global__ void __launch_bounds__(NTT_SMEM_THREADS, 512 / NTT_SMEM_THREADS) ntt_fused(const NttFusedArgs a) {
    extern __shared__ __align__(128) unsigned char ntt_fused_smem[];
    Fr regs[1][8];
    ntt_fused_block(a, blockIdx.x, reinterpret_cast<Fr*>(ntt_fused_smem), regs);
}
void launch_ntt_fused(const NttFusedArgs& a, cudaStream_t s) {
    if (!a.nblocks) return;
    static std::atomic<unsigned long long> attr_done{0};
    once_per_device(attr_done, [] {
        MB_CUDA(cudaFuncSetAttribute(ntt_fused, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(NTT_SMEM_ALLOC * sizeof(Fr))));
    });
    ntt_fused<<<(unsigned)a.nblocks, NTT_SMEM_THREADS, NTT_SMEM_ALLOC * sizeof(Fr), s>>>(a);
    MB_CUDA(cudaGetLastError());
    ::mb::g_launches++;
}
#else
void launch_ntt_fused(const NttFusedArgs& a, cudaStream_t s);
#endif
#endif

inline bool ntt_smem_supported(uint32_t log_n) { return log_n >= NTT_SMEM_L1 + 3 && log_n <= 2 * NTT_SMEM_L1; }

}  // namespace mb
