// Kernel definitions of group MISC (see rt.cuh: one translation unit per group).
#define MB_COLD_MUL
#define MB_DEFINE_MISC
#include "misc.cuh"
