// Kernel definitions of group MSM_G1 (see rt.cuh: one translation unit per group).
#define MB_DEFINE_MSM_G1
#include "msm.cuh"
