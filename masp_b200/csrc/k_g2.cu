// Kernel definitions of group G2 (see rt.cuh: one translation unit per group).
#define MB_COLD_MUL
#define MB_DEFINE_G2
#include "prover.cuh"
#include "synth.cuh"
