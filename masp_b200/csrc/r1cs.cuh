// Row evaluations of a rank-1 constraint system on the device.
//
// bellman's ProvingAssignment evaluates every constraint's three linear
// combinations on the CPU while the circuit is synthesised
// (`enforce` -> `eval`, called from Circuit::synthesize at
// masp_proofs/src/circuit/sapling.rs:139-417, 419-596, convert.rs:29-128;
// SURVEY.md §8 a-2).  The matrices are the same for every proof of a circuit,
// so here they are recorded once (csrc/host/gadgets.hpp), kept in HBM as
// CSR, and a = A z, b = B z, c = C z become a sparse matrix-vector product
// over the proof's scalar pool: the host ships only the witness z.
//
// Coefficients are indices into a small dictionary held in Montgomery form
// (z stays a plain integer: montmul(z, k R) = z k); entries 0 and 1 are +1 and
// -1, which cover most of the non-zeros and need no multiplication.
#pragma once
#include "field.cuh"

namespace mb {

struct R1csArgs {
    size_t nthreads;  // proofs * rows
    const uint32_t* rowptr[3];
    const uint32_t* col[3];   // pool index of the variable
    const uint32_t* cidx[3];  // dictionary index of the coefficient
    const Fr* dict;
    uint32_t ncons, rows;     // rows = ncons + n_inputs
    const Fr* pool;
    size_t pool_stride, idx_inputs;
    Fr* abc;                  // [proof][3][rows]
    const uint32_t* order;    // constraint rows by descending non-zero count: the 32 lanes of a warp
                              // get rows of similar length (a few rows have ~255 terms, most have 1-3)
};
MB_HD void r1cs_eval_body(const R1csArgs& a, size_t tid) {
    size_t proof = tid / a.rows;
    uint32_t row = (uint32_t)(tid - proof * a.rows);
    if (row < a.ncons) row = a.order[row];
    const Fr* z = a.pool + proof * a.pool_stride;
    Fr* out = a.abc + proof * 3 * (size_t)a.rows + row;
    if (row >= a.ncons) {  // bellman appends one row per input: a = input, b = c = 0
        out[0] = z[a.idx_inputs + (row - a.ncons)];
        out[a.rows] = Fr::zero();
        out[2 * (size_t)a.rows] = Fr::zero();
        return;
    }
    MB_NOUNROLL
    for (int k = 0; k < 3; ++k) {
        Fr acc = Fr::zero();
        uint32_t e0 = a.rowptr[k][row], e1 = a.rowptr[k][row + 1];
        MB_NOUNROLL
        for (uint32_t e = e0; e < e1; ++e) {
            Fr v = z[a.col[k][e]];
            uint32_t ci = a.cidx[k][e];
            if (ci == 0) acc = Fr::add(acc, v);
            else if (ci == 1) acc = Fr::sub(acc, v);
            else acc = Fr::add(acc, Fr::mul(v, a.dict[ci]));
        }
        out[(size_t)k * a.rows] = acc;
    }
}
MB_K_NTT(r1cs_eval, R1csArgs, r1cs_eval_body, 128)

struct R1csDev {
    DevBuf rowptr[3], col[3], cidx[3], dict, order;
    uint32_t ncons = 0, n_inputs = 0, n_aux = 0;
    bool bound = false;
};

}  // namespace mb
