// Kernel definitions of group RED_G2: latency-bound bucket reduction (out-of-line multiplies).
#define MB_COLD_MUL
#define MB_DEFINE_RED_G2
#include "msm.cuh"
