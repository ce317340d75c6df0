// Montgomery arithmetic over the BLS12-381 base field Fp (12 x 32-bit limbs)
// and scalar field Fr (8 x 32-bit limbs), kept in registers.
//
// Replaces, for the proving path, the field layer the reference takes from
// nam-blstrs / nam-blst (reference Cargo.lock:1385-1411).
//
// Multiplication is operand-scanning Montgomery with the partial products
// split over two accumulators by limb parity ("even"/"odd" columns), so that
// every 32x32->64 product (mad.lo.cc / madc.hi.cc pair on the IMAD pipe) joins
// one unbroken carry chain; the two accumulators are one limb apart and swap
// roles on every 32-bit shift of the reduction.
#pragma once
#include "rt.cuh"

namespace mb {

// ---------------------------------------------------------------------------
// carry-chain primitives: PTX on the device, exact emulation on the host
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
MB_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MB_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MB_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MB_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MB_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MB_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MB_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MB_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MB_D void mul_wide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
    uint64_t r;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    lo = (uint32_t)r;
    hi = (uint32_t)(r >> 32);
}
MB_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MB_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MB_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MB_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MB_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// The host build models CC.CF with one thread-local flag; each helper has the
// exact semantics of the PTX instruction it stands for.
inline uint32_t& cf() { static thread_local uint32_t f = 0; return f; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cf(); cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cf(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; cf() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cf(); cf() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cf(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline void mul_wide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) { uint64_t r = (uint64_t)a * b; lo = (uint32_t)r; hi = (uint32_t)(r >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + cf(); cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c; cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + cf(); cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)((((uint64_t)a * b) >> 32) + c + cf()); }
#endif

// ---------------------------------------------------------------------------
// field parameters (little-endian 32-bit limbs)
// ---------------------------------------------------------------------------
struct FpCfg {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;  // -p^{-1} mod 2^32
    MB_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t t[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return t[i];
    }
    MB_HD static constexpr uint32_t r1(int i) {  // 2^384 mod p
        constexpr uint32_t t[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return t[i];
    }
    MB_HD static constexpr uint32_t r2(int i) {  // 2^768 mod p
        constexpr uint32_t t[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return t[i];
    }
    MB_HD static constexpr uint32_t half(int i) {  // (p - 1) / 2
        constexpr uint32_t t[12] = {0xffffd555u, 0xdcff7fffu, 0x58a9ffffu, 0x0f55ffffu, 0x7b587b12u, 0xb3986950u,
                                    0x79c2895fu, 0xb23ba5c2u, 0x21a5d66bu, 0x258dd3dbu, 0x1cbff34du, 0x0d0088f5u};
        return t[i];
    }
};
struct FrCfg {
    static constexpr int N = 8;
    static constexpr uint32_t INV = 0xffffffffu;
    MB_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t t[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return t[i];
    }
    MB_HD static constexpr uint32_t r1(int i) {  // 2^256 mod r
        constexpr uint32_t t[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return t[i];
    }
    MB_HD static constexpr uint32_t r2(int i) {  // 2^512 mod r
        constexpr uint32_t t[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return t[i];
    }
    MB_HD static constexpr uint32_t half(int i) {
        constexpr uint32_t t[8] = {0x80000000u, 0x7fffffffu, 0x7fff2dffu, 0xa9ded201u,
                                   0x04d0ec02u, 0x199cec04u, 0x94cebea4u, 0x39f6d3a9u};
        return t[i];
    }
};

// ---------------------------------------------------------------------------
// Mont<Cfg>: an element in Montgomery form, fully reduced (< modulus)
// ---------------------------------------------------------------------------
#if defined(__CUDACC__)
// -r^-1 mod 2^32 is 2^32 - 1.  Given as a literal, ptxas turns m = t0 * INV into a negation, learns that
// t0 + m * r0 vanishes, and in simplifying that loses the lo / hi pairing of the whole reduction row: every
// m * r_i of an Fr multiplication became IMAD.X + IMAD.HI.X instead of one IMAD.WIDE.X (48 extra
// instructions on the multiplier pipe per multiplication, a quarter of the NTT kernels' IMAD count).
// Read from constant memory the factor is opaque and the row stays regular.
static __constant__ uint32_t mb_inv_minus_one = 0xffffffffu;
#endif

template <class C>
struct Mont {
    static constexpr int N = C::N;
    uint32_t v[N];

    MB_HD static Mont zero() {
        Mont r;
        MB_UNROLL
        for (int i = 0; i < N; ++i) r.v[i] = 0;
        return r;
    }
    MB_HD static Mont one() {
        Mont r;
        MB_UNROLL
        for (int i = 0; i < N; ++i) r.v[i] = C::r1(i);
        return r;
    }
    MB_HD static Mont r2() {
        Mont r;
        MB_UNROLL
        for (int i = 0; i < N; ++i) r.v[i] = C::r2(i);
        return r;
    }
    MB_HD bool is_zero() const {
        uint32_t o = 0;
        MB_UNROLL
        for (int i = 0; i < N; ++i) o |= v[i];
        return o == 0;
    }
    MB_HD bool eq(const Mont& b) const {
        uint32_t o = 0;
        MB_UNROLL
        for (int i = 0; i < N; ++i) o |= v[i] ^ b.v[i];
        return o == 0;
    }

    // r = a + b mod p
    MB_HD static Mont add(const Mont& a, const Mont& b) {
        Mont s, t;
        s.v[0] = add_cc(a.v[0], b.v[0]);
        MB_UNROLL
        for (int i = 1; i < N - 1; ++i) s.v[i] = addc_cc(a.v[i], b.v[i]);
        s.v[N - 1] = addc(a.v[N - 1], b.v[N - 1]);  // 2p < 2^(32N): no carry out
        t.v[0] = sub_cc(s.v[0], C::mod(0));
        MB_UNROLL
        for (int i = 1; i < N; ++i) t.v[i] = subc_cc(s.v[i], C::mod(i));
        uint32_t borrow = subc(0, 0);  // all ones iff s < p
        MB_UNROLL
        for (int i = 0; i < N; ++i) s.v[i] = borrow ? s.v[i] : t.v[i];
        return s;
    }
    // r = a - b mod p
    MB_HD static Mont sub(const Mont& a, const Mont& b) {
        Mont d;
        d.v[0] = sub_cc(a.v[0], b.v[0]);
        MB_UNROLL
        for (int i = 1; i < N; ++i) d.v[i] = subc_cc(a.v[i], b.v[i]);
        uint32_t m = subc(0, 0);  // all ones iff a < b
        d.v[0] = add_cc(d.v[0], C::mod(0) & m);
        MB_UNROLL
        for (int i = 1; i < N - 1; ++i) d.v[i] = addc_cc(d.v[i], C::mod(i) & m);
        d.v[N - 1] = addc(d.v[N - 1], C::mod(N - 1) & m);
        return d;
    }
    MB_HD static Mont neg(const Mont& a) {
        Mont d;
        d.v[0] = sub_cc(C::mod(0), a.v[0]);
        MB_UNROLL
        for (int i = 1; i < N - 1; ++i) d.v[i] = subc_cc(C::mod(i), a.v[i]);
        d.v[N - 1] = subc(C::mod(N - 1), a.v[N - 1]);
        uint32_t nz = 0;
        MB_UNROLL
        for (int i = 0; i < N; ++i) nz |= a.v[i];
        MB_UNROLL
        for (int i = 0; i < N; ++i) d.v[i] = nz ? d.v[i] : 0u;
        return d;
    }
    MB_HD static Mont dbl(const Mont& a) { return add(a, a); }

    // one row of the parity-split accumulation: ev += x[even] * y, od += x[odd] * y
    // (od first: its chain has no carry out; the ev chain's carry lands in od[N-1])
    MB_HD static void row_acc(uint32_t* ev, uint32_t* od, const uint32_t* x, uint32_t y) {
        od[0] = mad_lo_cc(x[1], y, od[0]);
        od[1] = madc_hi_cc(x[1], y, od[1]);
        MB_UNROLL
        for (int j = 2; j < N; j += 2) {
            od[j] = madc_lo_cc(x[j + 1], y, od[j]);
            od[j + 1] = madc_hi_cc(x[j + 1], y, od[j + 1]);
        }
        ev[0] = mad_lo_cc(x[0], y, ev[0]);
        ev[1] = madc_hi_cc(x[0], y, ev[1]);
        MB_UNROLL
        for (int j = 2; j < N; j += 2) {
            ev[j] = madc_lo_cc(x[j], y, ev[j]);
            ev[j + 1] = madc_hi_cc(x[j], y, ev[j + 1]);
        }
        od[N - 1] = addc(od[N - 1], 0);
    }
    // Montgomery step on the current row: make ev[0] vanish
    MB_HD static void row_redc(uint32_t* ev, uint32_t* od) {
#if defined(__CUDA_ARCH__)
        uint32_t m = mul_lo(ev[0], C::INV == 0xffffffffu ? mb_inv_minus_one : C::INV);
#else
        uint32_t m = mul_lo(ev[0], C::INV);
#endif
        uint32_t p[N];
        MB_UNROLL
        for (int i = 0; i < N; ++i) p[i] = C::mod(i);
        row_acc(ev, od, p, m);
    }
    // Shift the pair right by one limb and add x * y: `ev` (the old odd
    // accumulator) becomes the even one; `od` (the old even one, whose limb 0
    // is now zero) slides down two limbs and becomes the odd one.
    MB_HD static void row_shift_acc(uint32_t* ev, uint32_t* od, const uint32_t* x, uint32_t y) {
        ev[0] = add_cc(ev[0], od[1]);
        MB_UNROLL
        for (int j = 0; j < N - 2; j += 2) {
            od[j] = madc_lo_cc(x[j + 1], y, od[j + 2]);
            od[j + 1] = madc_hi_cc(x[j + 1], y, od[j + 3]);
        }
        od[N - 2] = madc_lo_cc(x[N - 1], y, 0);
        od[N - 1] = madc_hi(x[N - 1], y, 0);
        ev[0] = mad_lo_cc(x[0], y, ev[0]);
        ev[1] = madc_hi_cc(x[0], y, ev[1]);
        MB_UNROLL
        for (int j = 2; j < N; j += 2) {
            ev[j] = madc_lo_cc(x[j], y, ev[j]);
            ev[j + 1] = madc_hi_cc(x[j], y, ev[j + 1]);
        }
        od[N - 1] = addc(od[N - 1], 0);
    }

    // r = a * b * 2^(-32N) mod p.  Units that hold only latency-bound code
    // (MB_COLD_MUL) call one out-of-line body: their kernels are a few hundred
    // instructions instead of tens of thousands and stay in the instruction cache.
#ifdef MB_COLD_MUL
    MB_COLD Mont mul(const Mont& a, const Mont& b) { return mul_inline(a, b); }
#else
    MB_HD static Mont mul(const Mont& a, const Mont& b) { return mul_inline(a, b); }
#endif
    MB_HD static Mont mul_inline(const Mont& a, const Mont& b) {
        uint32_t ev[N], od[N];
        // first row: plain products
        MB_UNROLL
        for (int j = 0; j < N; j += 2) {
            mul_wide(a.v[j], b.v[0], ev[j], ev[j + 1]);
            mul_wide(a.v[j + 1], b.v[0], od[j], od[j + 1]);
        }
        row_redc(ev, od);
        MB_UNROLL
        for (int i = 1; i < N; i += 2) {
            row_shift_acc(od, ev, a.v, b.v[i]);
            row_redc(od, ev);
            if (i + 1 < N) {
                row_shift_acc(ev, od, a.v, b.v[i + 1]);
                row_redc(ev, od);
            }
        }
        // N is even: `od` ended as the even accumulator with od[0] == 0.
        // result = (od >> 32) + ev
        Mont s, t;
        s.v[0] = add_cc(ev[0], od[1]);
        MB_UNROLL
        for (int i = 1; i < N - 1; ++i) s.v[i] = addc_cc(ev[i], od[i + 1]);
        s.v[N - 1] = addc(ev[N - 1], 0);
        // s < 2p: one conditional subtraction
        t.v[0] = sub_cc(s.v[0], C::mod(0));
        MB_UNROLL
        for (int i = 1; i < N; ++i) t.v[i] = subc_cc(s.v[i], C::mod(i));
        uint32_t borrow = subc(0, 0);
        MB_UNROLL
        for (int i = 0; i < N; ++i) s.v[i] = borrow ? s.v[i] : t.v[i];
        return s;
    }
    // r = (a * b + c * d) * 2^(-32N) mod p: both products share one accumulator pair and ONE Montgomery
    // reduction (N rows of N wide multiplies saved against two multiplications and an addition).
    // Bounds: the running value stays below 3p (2^32 + 1), which the pair's N + 1 limbs hold when
    // 3p < 2^(32N): true for Fp (p = 0.10 * 2^384), NOT for Fr (r = 0.45 * 2^256), hence SOP2_OK; the
    // result is below (2 p^2 + 2^(32N) p) / 2^(32N) < 1.21 p, so one conditional subtraction brings it under p.
    static constexpr bool SOP2_OK = C::mod(N - 1) < 0x55555555u;  // 3p < 2^(32N)
    static constexpr bool SOP4_OK = C::mod(N - 1) < 0x33333333u;  // 5p < 2^(32N)
    MB_HD static Mont sop2_inline(const Mont& a, const Mont& b, const Mont& c, const Mont& d) {
        static_assert(SOP2_OK, "modulus too close to 2^(32N) for a fused two-product reduction");
        uint32_t ev[N], od[N];
        MB_UNROLL
        for (int j = 0; j < N; j += 2) {
            mul_wide(a.v[j], b.v[0], ev[j], ev[j + 1]);
            mul_wide(a.v[j + 1], b.v[0], od[j], od[j + 1]);
        }
        row_acc(ev, od, c.v, d.v[0]);
        row_redc(ev, od);
        MB_UNROLL
        for (int i = 1; i < N; i += 2) {
            row_shift_acc(od, ev, a.v, b.v[i]);
            row_acc(od, ev, c.v, d.v[i]);
            row_redc(od, ev);
            if (i + 1 < N) {
                row_shift_acc(ev, od, a.v, b.v[i + 1]);
                row_acc(ev, od, c.v, d.v[i + 1]);
                row_redc(ev, od);
            }
        }
        Mont s, t;
        s.v[0] = add_cc(ev[0], od[1]);
        MB_UNROLL
        for (int i = 1; i < N - 1; ++i) s.v[i] = addc_cc(ev[i], od[i + 1]);
        s.v[N - 1] = addc(ev[N - 1], 0);
        t.v[0] = sub_cc(s.v[0], C::mod(0));
        MB_UNROLL
        for (int i = 1; i < N; ++i) t.v[i] = subc_cc(s.v[i], C::mod(i));
        uint32_t borrow = subc(0, 0);
        MB_UNROLL
        for (int i = 0; i < N; ++i) s.v[i] = borrow ? s.v[i] : t.v[i];
        return s;
    }
    // the same with four products (Fp2: both components of R T - Y PPP): the running value stays below
    // 5p (2^32 + 1) < 0.51 * 2^(32 (N + 1)), the result below (4 p^2 + 2^(32N) p) / 2^(32N) < 1.41 p
    MB_HD static Mont sop4_inline(const Mont& a0, const Mont& b0, const Mont& a1, const Mont& b1, const Mont& a2,
                                  const Mont& b2, const Mont& a3, const Mont& b3) {
        static_assert(SOP4_OK, "modulus too close to 2^(32N) for a fused four-product reduction");
        uint32_t ev[N], od[N];
        MB_UNROLL
        for (int j = 0; j < N; j += 2) {
            mul_wide(a0.v[j], b0.v[0], ev[j], ev[j + 1]);
            mul_wide(a0.v[j + 1], b0.v[0], od[j], od[j + 1]);
        }
        row_acc(ev, od, a1.v, b1.v[0]);
        row_acc(ev, od, a2.v, b2.v[0]);
        row_acc(ev, od, a3.v, b3.v[0]);
        row_redc(ev, od);
        MB_UNROLL
        for (int i = 1; i < N; i += 2) {
            row_shift_acc(od, ev, a0.v, b0.v[i]);
            row_acc(od, ev, a1.v, b1.v[i]);
            row_acc(od, ev, a2.v, b2.v[i]);
            row_acc(od, ev, a3.v, b3.v[i]);
            row_redc(od, ev);
            if (i + 1 < N) {
                row_shift_acc(ev, od, a0.v, b0.v[i + 1]);
                row_acc(ev, od, a1.v, b1.v[i + 1]);
                row_acc(ev, od, a2.v, b2.v[i + 1]);
                row_acc(ev, od, a3.v, b3.v[i + 1]);
                row_redc(ev, od);
            }
        }
        Mont s, t;
        s.v[0] = add_cc(ev[0], od[1]);
        MB_UNROLL
        for (int i = 1; i < N - 1; ++i) s.v[i] = addc_cc(ev[i], od[i + 1]);
        s.v[N - 1] = addc(ev[N - 1], 0);
        t.v[0] = sub_cc(s.v[0], C::mod(0));
        MB_UNROLL
        for (int i = 1; i < N; ++i) t.v[i] = subc_cc(s.v[i], C::mod(i));
        uint32_t borrow = subc(0, 0);
        MB_UNROLL
        for (int i = 0; i < N; ++i) s.v[i] = borrow ? s.v[i] : t.v[i];
        return s;
    }
    // (A dedicated squaring -- cross products once, product-scanning reduction, 222 instead of 288
    // wide multiplies -- is exact but measured 2 % SLOWER inside the accumulation kernel: its
    // three-word column accumulator serialises what the row-wise form leaves independent.
    // Kept with its test under scripts/microbench/sqr_ps.cuh.)
    // Row-wise triangular squaring (MB_TRI_SQR units): a^2 = sum_i a_i B^i * (a_i B^i + 2 sum_{j>i} a_j B^j),
    // B = 2^32.  Row i multiplies a_i into the vector (a_i, 2 U_i), U_i = the limbs above i, whose doubled
    // limbs are e_{i+1} = a_{i+1} << 1 and d_j = (a_j << 1) | (a_{j-1} >> 31) above that (a < 2^(32N - 1) for
    // both moduli, so nothing is shifted out).  Every contribution to limb k comes from a row <= k, so the
    // interleaved Montgomery steps see the same low limbs as in mul(a, a) and the result is bit-identical; row i
    // simply has no products below position i: 78 instead of 144 product multiplies for N = 12, on the same
    // two accumulators (the skipped pairs of the odd chain become carry-propagating moves on the ALU pipe).
    // Row k brings its doubled cross terms in ahead of the rows that would add them in mul(a, a), so the running
    // value is bounded by 2a + p < 3p instead of 2p: the same head-room condition as sop2 (Fp yes, Fr no).
    MB_HD static void row_shift_acc_tri(uint32_t* ev, uint32_t* od, const uint32_t* x, uint32_t y, int j0) {
        ev[0] = add_cc(ev[0], od[1]);
        MB_UNROLL
        for (int j = 0; j < N - 2; j += 2) {
            if (j + 1 >= j0) {
                od[j] = madc_lo_cc(x[j + 1], y, od[j + 2]);
                od[j + 1] = madc_hi_cc(x[j + 1], y, od[j + 3]);
            } else {
                od[j] = addc_cc(od[j + 2], 0);
                od[j + 1] = addc_cc(od[j + 3], 0);
            }
        }
        od[N - 2] = madc_lo_cc(x[N - 1], y, 0);
        od[N - 1] = madc_hi(x[N - 1], y, 0);
        bool first = true;
        MB_UNROLL
        for (int j = 0; j < N; j += 2) {
            if (j >= j0) {
                ev[j] = first ? mad_lo_cc(x[j], y, ev[j]) : madc_lo_cc(x[j], y, ev[j]);
                ev[j + 1] = madc_hi_cc(x[j], y, ev[j + 1]);
                first = false;
            }
        }
        if (!first) od[N - 1] = addc(od[N - 1], 0);  // the even chain's carry (no even term at all for j0 = N - 1)
    }
    MB_HD static Mont sqr_inline(const Mont& a) {
        static_assert(SOP2_OK, "modulus too close to 2^(32N) for the triangular squaring");
        uint32_t d[N], ev[N], od[N], x[N];
        d[0] = a.v[0] << 1;
        MB_UNROLL
        for (int j = 1; j < N; ++j) d[j] = (a.v[j] << 1) | (a.v[j - 1] >> 31);
        // row 0: a_0 * (a_0, e_1, d_2, ...)
        MB_UNROLL
        for (int j = 0; j < N; ++j) x[j] = j == 0 ? a.v[0] : (j == 1 ? (a.v[1] << 1) : d[j]);
        MB_UNROLL
        for (int j = 0; j < N; j += 2) {
            mul_wide(x[j], a.v[0], ev[j], ev[j + 1]);
            mul_wide(x[j + 1], a.v[0], od[j], od[j + 1]);
        }
        row_redc(ev, od);
        MB_UNROLL
        for (int i = 1; i < N; i += 2) {
            MB_UNROLL
            for (int j = 0; j < N; ++j) x[j] = j == i ? a.v[j] : (j == i + 1 ? (a.v[j] << 1) : d[j]);
            row_shift_acc_tri(od, ev, x, a.v[i], i);
            row_redc(od, ev);
            if (i + 1 < N) {
                MB_UNROLL
                for (int j = 0; j < N; ++j) x[j] = j == i + 1 ? a.v[j] : (j == i + 2 ? (a.v[j] << 1) : d[j]);
                row_shift_acc_tri(ev, od, x, a.v[i + 1], i + 1);
                row_redc(ev, od);
            }
        }
        Mont s, t;
        s.v[0] = add_cc(ev[0], od[1]);
        MB_UNROLL
        for (int i = 1; i < N - 1; ++i) s.v[i] = addc_cc(ev[i], od[i + 1]);
        s.v[N - 1] = addc(ev[N - 1], 0);
        t.v[0] = sub_cc(s.v[0], C::mod(0));
        MB_UNROLL
        for (int i = 1; i < N; ++i) t.v[i] = subc_cc(s.v[i], C::mod(i));
        uint32_t borrow = subc(0, 0);
        MB_UNROLL
        for (int i = 0; i < N; ++i) s.v[i] = borrow ? s.v[i] : t.v[i];
        return s;
    }
#if defined(MB_TRI_SQR) && !defined(MB_COLD_MUL)
    MB_HD static Mont sqr(const Mont& a) {
        if constexpr (SOP2_OK) return sqr_inline(a);
        else return mul(a, a);
    }
#else
    MB_HD static Mont sqr(const Mont& a) { return mul(a, a); }
#endif

    // reference multiplication with 64-bit temporaries (self-test only)
    MB_HD static Mont mul_portable(const Mont& a, const Mont& b) {
        uint32_t t[N + 2];
        for (int i = 0; i < N + 2; ++i) t[i] = 0;
        for (int i = 0; i < N; ++i) {
            uint64_t c = 0;
            for (int j = 0; j < N; ++j) {
                c += (uint64_t)a.v[j] * b.v[i] + t[j];
                t[j] = (uint32_t)c;
                c >>= 32;
            }
            c += t[N];
            t[N] = (uint32_t)c;
            t[N + 1] = (uint32_t)(c >> 32);
            uint32_t m = t[0] * C::INV;
            c = (uint64_t)m * C::mod(0) + t[0];
            c >>= 32;
            for (int j = 1; j < N; ++j) {
                c += (uint64_t)m * C::mod(j) + t[j];
                t[j - 1] = (uint32_t)c;
                c >>= 32;
            }
            c += t[N];
            t[N - 1] = (uint32_t)c;
            t[N] = t[N + 1] + (uint32_t)(c >> 32);
        }
        // conditional subtract
        uint32_t d[N];
        uint64_t br = 0;
        for (int i = 0; i < N; ++i) {
            uint64_t x = (uint64_t)t[i] - C::mod(i) - br;
            d[i] = (uint32_t)x;
            br = (x >> 63) & 1;
        }
        Mont r;
        bool ge = t[N] != 0 || br == 0;
        for (int i = 0; i < N; ++i) r.v[i] = ge ? d[i] : t[i];
        return r;
    }

    // plain integer (already < modulus) <-> Montgomery form
    MB_HD static Mont from_std(const Mont& a) { return mul(a, r2()); }
    MB_HD static Mont to_std(const Mont& a) {
        Mont o = zero();
        o.v[0] = 1;
        return mul(a, o);
    }
    // a >= modulus ? (plain integers)
    MB_HD static bool std_ge_mod(const Mont& a) {
        sub_cc(a.v[0], C::mod(0));
        MB_UNROLL
        for (int i = 1; i < N; ++i) subc_cc(a.v[i], C::mod(i));
        return subc(0, 0) == 0;
    }
    // plain integer a > (p-1)/2 ?
    MB_HD static bool std_gt_half(const Mont& a) {
        sub_cc(C::half(0), a.v[0]);
        MB_UNROLL
        for (int i = 1; i < N; ++i) subc_cc(C::half(i), a.v[i]);
        return subc(0, 0) != 0;
    }

    // a^-1 by the binary extended Euclidean algorithm on the plain integers (u, v) = (a R mod p, p) with
    // cofactors (x1, x2): ~2 * 32 N rounds of a shift or a subtraction on N limbs, against the Fermat power's
    // ~1.5 * 32 N dependent multiplications of N^2 wide multiplies each -- this chain is what a caller of
    // `proof_finish` / `sum_partials` (one thread per point) waits for.  The loop yields (a R)^-1 as a plain
    // integer; one multiplication by R^3 turns it into a^-1 R.  inv(0) = 0, like the power.
    MB_HD static Mont inv(const Mont& a) {
        if (a.is_zero()) return zero();
        Mont u = a, v, x1 = zero(), x2 = zero();
        MB_UNROLL
        for (int i = 0; i < N; ++i) v.v[i] = C::mod(i);
        x1.v[0] = 1;
        MB_NOUNROLL
        for (;;) {
            MB_NOUNROLL
            while ((u.v[0] & 1u) == 0) {
                shr1(u, 0);
                halve(x1);
            }
            if (is_plain_one(u)) break;
            MB_NOUNROLL
            while ((v.v[0] & 1u) == 0) {
                shr1(v, 0);
                halve(x2);
            }
            if (is_plain_one(v)) break;
            if (plain_geq(u, v)) {
                plain_sub(u, v);
                x1 = sub(x1, x2);
            } else {
                plain_sub(v, u);
                x2 = sub(x2, x1);
            }
        }
        const Mont r3 = mul(r2(), r2());
        return mul(is_plain_one(u) ? x1 : x2, r3);
    }
    MB_HD static bool is_plain_one(const Mont& a) {
        uint32_t o = a.v[0] ^ 1u;
        MB_UNROLL
        for (int i = 1; i < N; ++i) o |= a.v[i];
        return o == 0;
    }
    MB_HD static void shr1(Mont& a, uint32_t top_bit) {  // a = (top_bit : a) >> 1
        MB_UNROLL
        for (int i = 0; i < N - 1; ++i) a.v[i] = (a.v[i] >> 1) | (a.v[i + 1] << 31);
        a.v[N - 1] = (a.v[N - 1] >> 1) | (top_bit << 31);
    }
    MB_HD static void halve(Mont& x) {  // x / 2 mod p for x < p (x + p < 2^(32N) for both moduli: no carry out)
        if (x.v[0] & 1u) {
            x.v[0] = add_cc(x.v[0], C::mod(0));
            MB_UNROLL
            for (int i = 1; i < N - 1; ++i) x.v[i] = addc_cc(x.v[i], C::mod(i));
            x.v[N - 1] = addc(x.v[N - 1], C::mod(N - 1));
        }
        shr1(x, 0);
    }
    MB_HD static bool plain_geq(const Mont& a, const Mont& b) {
        sub_cc(a.v[0], b.v[0]);
        MB_UNROLL
        for (int i = 1; i < N; ++i) subc_cc(a.v[i], b.v[i]);
        return subc(0, 0) == 0;
    }
    MB_HD static void plain_sub(Mont& a, const Mont& b) {  // a -= b, a >= b
        a.v[0] = sub_cc(a.v[0], b.v[0]);
        MB_UNROLL
        for (int i = 1; i < N - 1; ++i) a.v[i] = subc_cc(a.v[i], b.v[i]);
        a.v[N - 1] = subc(a.v[N - 1], b.v[N - 1]);
    }
    // a^(p-2) by square-and-multiply over the bits of p - 2 (the self-test's reference for inv)
    MB_HD static Mont inv_fermat(const Mont& a) {
        uint32_t e[N];
        uint32_t borrow = 2;
        for (int i = 0; i < N; ++i) {
            uint32_t m = C::mod(i);
            e[i] = m - borrow;
            borrow = m < borrow ? 1u : 0u;
        }
        // a chain of ~1.5 * 32 N dependent multiplications: the inlined body has two thirds of the
        // out-of-line latency (profiles/r01_latency_microbench.txt) and this loop is its only copy per unit
        Mont r = one();
        MB_NOUNROLL
        for (int i = 32 * N - 1; i >= 0; --i) {
            r = mul_inline(r, r);
            if ((e[i >> 5] >> (i & 31)) & 1) r = mul_inline(r, a);
        }
        return r;
    }
};

typedef Mont<FpCfg> Fp;
typedef Mont<FrCfg> Fr;

// ---------------------------------------------------------------------------
// Fp2 = Fp[u] / (u^2 + 1)
// ---------------------------------------------------------------------------
struct Fp2 {
    Fp c0, c1;
    MB_HD static Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
    MB_HD static Fp2 one() { return {Fp::one(), Fp::zero()}; }
    MB_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    MB_HD bool eq(const Fp2& b) const { return c0.eq(b.c0) && c1.eq(b.c1); }
    MB_HD static Fp2 add(const Fp2& a, const Fp2& b) { return {Fp::add(a.c0, b.c0), Fp::add(a.c1, b.c1)}; }
    MB_HD static Fp2 sub(const Fp2& a, const Fp2& b) { return {Fp::sub(a.c0, b.c0), Fp::sub(a.c1, b.c1)}; }
    MB_HD static Fp2 neg(const Fp2& a) { return {Fp::neg(a.c0), Fp::neg(a.c1)}; }
    MB_HD static Fp2 dbl(const Fp2& a) { return {Fp::dbl(a.c0), Fp::dbl(a.c1)}; }
    MB_HD static Fp2 mul(const Fp2& a, const Fp2& b) {
#if defined(MB_FP2_SOP) && !defined(MB_COLD_MUL)
        // same 864 wide multiplies as Karatsuba's three multiplications, as two fused two-product reductions
        Fp na1 = Fp::neg(a.c1);
        return {Fp::sop2_inline(a.c0, b.c0, na1, b.c1), Fp::sop2_inline(a.c0, b.c1, a.c1, b.c0)};
#endif
        Fp t0 = Fp::mul(a.c0, b.c0);
        Fp t1 = Fp::mul(a.c1, b.c1);
        Fp t2 = Fp::mul(Fp::add(a.c0, a.c1), Fp::add(b.c0, b.c1));
        return {Fp::sub(t0, t1), Fp::sub(Fp::sub(t2, t0), t1)};
    }
    MB_HD static Fp2 sqr(const Fp2& a) {
        Fp t = Fp::mul(a.c0, a.c1);
        return {Fp::mul(Fp::add(a.c0, a.c1), Fp::sub(a.c0, a.c1)), Fp::dbl(t)};
    }
    MB_HD static Fp2 inv(const Fp2& a) {
        Fp t = Fp::inv(Fp::add(Fp::sqr(a.c0), Fp::sqr(a.c1)));
        return {Fp::mul(a.c0, t), Fp::neg(Fp::mul(a.c1, t))};
    }
};

}  // namespace mb
