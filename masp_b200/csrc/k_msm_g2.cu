// Kernel definitions of group MSM_G2 (see rt.cuh: one translation unit per group).
// Fp2 products as fused sums of Fp products sharing one Montgomery reduction (field.cuh sop2 / sop4): same
// multiply count as Karatsuba for a plain Fp2 product, two reductions fewer in the mixed addition's Y3, and
// 124 instead of 628 bytes of spill code at the 255-register cap.  B200, Spend proofs/s: Karatsuba 550.0,
// fused products 555.8, + fused Y3 557.1 (profiles/r02_ab_g2_fused.jsonl).
#define MB_FP2_SOP
#define MB_DEFINE_MSM_G2
#include "msm.cuh"
