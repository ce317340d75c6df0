// Kernel definitions of group NTT (see rt.cuh: one translation unit per group).
#define MB_DEFINE_NTT
#include "ntt.cuh"
#include "r1cs.cuh"
