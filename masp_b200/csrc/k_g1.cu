// Kernel definitions of group G1 (see rt.cuh: one translation unit per group).
#define MB_COLD_MUL
#define MB_DEFINE_G1
#include "prover.cuh"
#include "synth.cuh"
