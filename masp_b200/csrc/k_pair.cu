// Kernel definitions of the pairing group (verification self-check); see rt.cuh.
// The tower functions (f2_mul, f6_mul, f12_mul, ...) are out-of-line bodies already; the
// field multiplication stays inlined inside them: this code is a chain of dependent
// multiplications and the inlined form has two thirds of the latency (profiles/r01_latency_microbench.txt).
#define MB_DEFINE_PAIR
#include "pairing.cuh"
