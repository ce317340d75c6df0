// Kernel definitions of group MSM_G1 (see rt.cuh: one translation unit per group).
// the two squarings of the G1 mixed addition (P^2, R^2) as triangular row-wise squarings: 78 instead of 144
// product multiplies each, same accumulator pair, bit-identical to mul(a, a) (field.cuh sqr_inline)
#define MB_TRI_SQR
#define MB_DEFINE_MSM_G1
#include "msm.cuh"
