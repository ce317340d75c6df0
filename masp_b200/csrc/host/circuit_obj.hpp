// The object behind `mb200_circuit`: one recorded MASP circuit (its three
// sparse matrices, counts, density bitmaps) plus the witness layout.
#pragma once
#include <string>
#include <vector>

#include "gadgets.hpp"

struct mb200_circuit {
    int kind = 0;          // MB200_CIRCUIT_*
    uint32_t depth = 0;    // Merkle depth (Spend, Convert)
    uint32_t n_inputs = 0, n_aux = 0, n_constraints = 0;
    size_t witness_bytes = 0;
    mbh::Matrix A, B, C;
    std::vector<uint8_t> a_aux_density, b_input_density, b_aux_density;  // LSB-first bitmaps
    uint32_t a_aux_ones = 0, b_input_ones = 0, b_aux_ones = 0;
    std::string hash_hex;  // TestConstraintSystem::hash of the recorded system
};
