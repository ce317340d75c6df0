// Kernel definitions of group MSM_G2 (see rt.cuh: one translation unit per group).
#define MB_DEFINE_MSM_G2
#include "msm.cuh"
