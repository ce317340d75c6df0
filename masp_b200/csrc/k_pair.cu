// Kernel definitions of the pairing group (verification self-check); see rt.cuh.
#define MB_COLD_MUL
#define MB_DEFINE_PAIR
#include "pairing.cuh"
