#define MB200_EMU
#include "/root/repo/masp_b200/csrc/field.cuh"
#include <random>
using namespace mb;
unsigned long long mb::g_launches;
int main() {
    std::mt19937_64 rng(7);
    int bad = 0;
    for (int it = 0; it < 200000; ++it) {
        Fp a;
        for (;;) { for (int i = 0; i < 12; ++i) a.v[i] = (uint32_t)rng(); a.v[11] &= 0x1fffffff; if (!Fp::std_ge_mod(a)) break; }
        if (it == 0) a = Fp::zero();
        if (it == 1) { for (int i = 0; i < 12; ++i) a.v[i] = FpCfg::mod(i); a.v[0] -= 1; }
        if (it == 2) { for (int i = 0; i < 12; ++i) a.v[i] = 0xffffffffu; a.v[11] = 0x1a0111e9u; }
        if (it == 3) a = Fp::one();
        if (!Fp::sqr_inline(a).eq(Fp::mul_portable(a, a))) { if (bad++ < 5) printf("Fp mismatch %d\n", it); }
        Fr b;
        for (;;) { for (int i = 0; i < 8; ++i) b.v[i] = (uint32_t)rng(); b.v[7] &= 0x7fffffff; if (!Fr::std_ge_mod(b)) break; }
        if (it == 1) { for (int i = 0; i < 8; ++i) b.v[i] = FrCfg::mod(i); b.v[0] -= 1; }
        if (!Fr::sqr_inline(b).eq(Fr::mul_portable(b, b))) { if (bad++ < 5) printf("Fr mismatch %d\n", it); }
    }
    printf("bad=%d\n", bad);
    return bad != 0;
}
