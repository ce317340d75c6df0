/* masp_b200 -- C ABI of the Blackwell-native Groth16 proving path for the MASP
 * Spend / Output / Convert circuits.
 *
 * The reference (namada-net/masp) exposes this path as a Rust trait, not an
 * FFI: a grep for `extern "C"` in /root/reference finds
 * nothing, so the boundary below is the set of entry points a Rust shim
 * (`impl TxProver for B200TxProver`) binds.  Each entry cites
 * the reference interface it stands behind.
 *
 * Conventions
 *   - every function returns 0 on success or a negative MB200_E* code; nothing
 *     unwinds or aborts across the boundary (the reference panics instead:
 *     masp_proofs/src/sapling/prover.rs:117,202,252; src/lib.rs:290-293);
 *   - all buffers are caller-owned; "host" buffers are ordinary host memory
 *     (pinned memory makes the copies asynchronous), "device" buffers are
 *     device pointers on the current device;
 *   - scalars are 32-byte little-endian canonical integers < r
 *     (= PrimeField::to_repr of bls12_381::Scalar);
 *   - points use bellman's wire encodings (zkcrypto uncompressed / compressed).
 *   - one process drives every GPU it passes to mb200_init -- the reference's prover is one
 *     `&self` object shared by all threads (masp_proofs/src/prover.rs:27-33, 156-261): keys are
 *     replicated to each device, a batch is cut into per-device slices (one host thread
 *     enqueues each), standalone calls run on device_ids[0].  One rank per GPU under
 *     torch.distributed (each rank passing its own device) works the same way.  Calls are
 *     serialised internally.
 */
#ifndef MASP_B200_H
#define MASP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB200_OK 0
#define MB200_EINVAL (-1)  /* bad argument / shape mismatch */
#define MB200_EPARSE (-2)  /* malformed Parameters bytes */
#define MB200_ECUDA (-3)   /* CUDA runtime failure (including: no device) */
#define MB200_ENOMEM (-4)
#define MB200_ESTATE (-5)  /* mb200_init not called */
#define MB200_ESCALAR (-6) /* a scalar is not canonical (>= r) */
#define MB200_EVERIFY (-8) /* a proof failed the self-check (the reference returns Err(()) at sapling/prover.rs:148, :266) */
#define MB200_EPARAMS (-9) /* parameter file: wrong size or BLAKE2b-512 digest (the reference panics: lib.rs:290-293, 359-388) */
#define MB200_EIO (-10)    /* parameter file cannot be opened / read (the reference panics: lib.rs:316-318) */
#define MB200_ESYNTH (-7)  /* witness generation failed (bellman SynthesisError: division by zero / unsatisfiable) */

#define MB200_PROOF_BYTES 192 /* GROTH_PROOF_SIZE, masp_primitives/src/transaction/components.rs:14-15 */

typedef struct mb200_params mb200_params;

/* Opens the listed devices (NULL / 0 = the current device; NULL / -1 = every visible device)
 * and creates their streams and chunk contexts.  device_ids[0] is the primary device (standalone
 * MSM / NTT / verification calls, the destination of split-MSM partials).  Fails with MB200_ECUDA
 * when there is no usable GPU: there is no CPU fallback. */
int mb200_init(const int* device_ids, int n_devices);
int mb200_shutdown(void);
/* devices opened by mb200_init (0 before it) */
int mb200_device_count(void);

/* groth16::Parameters::<Bls12>::read(reader, false) as called at
 * masp_proofs/src/lib.rs:336-341, plus the three density bitmaps bellman's
 * ProvingAssignment records during synthesis (a_aux_density, b_input_density,
 * b_aux_density; LSB-first bit i = variable i; NULL = all dense).  Bytes after
 * the Parameters encoding (the MPC transcript of the real .params files,
 * lib.rs:343-388) are ignored; mb200_params_info reports how many were
 * consumed.  The key is converted to Montgomery limbs and expanded into
 * per-window tables resident in HBM; it is immutable and may be shared. */
int mb200_params_load(const uint8_t* bytes, size_t len, const uint8_t* a_aux_density,
                      const uint8_t* b_input_density, const uint8_t* b_aux_density, mb200_params** out);
/* load_parameters / parse_parameters for one file (masp_proofs/src/lib.rs:278-325, 343-388): the size is
 * checked against expected_bytes BEFORE anything is read (verify_file_size), Parameters::read(.., false)
 * consumes the key, and the whole stream -- the MPC transcript behind the key included -- is BLAKE2b-512
 * hashed and its hex digest compared with expected_blake2b_hex (verify_hash).  expected_bytes = 0 or an
 * empty / NULL digest skips that check.  mb200_masp_params_spec returns the reference's constants
 * (lib.rs:61-76) for kind = MB200_CIRCUIT_{SPEND,OUTPUT,CONVERT}.  Failures: MB200_EIO, MB200_EPARAMS,
 * MB200_EPARSE where the reference panics. */
int mb200_params_load_file(const char* path, uint64_t expected_bytes, const char* expected_blake2b_hex,
                           const uint8_t* a_aux_density, const uint8_t* b_input_density, const uint8_t* b_aux_density,
                           mb200_params** out);
/* the same for bytes already in memory (LocalTxProver::from_bytes -> parse_parameters, prover.rs:120-136) */
int mb200_params_load_verified(const uint8_t* bytes, size_t len, uint64_t expected_bytes, const char* expected_blake2b_hex,
                               const uint8_t* a_aux_density, const uint8_t* b_input_density,
                               const uint8_t* b_aux_density, mb200_params** out);
int mb200_masp_params_spec(int kind, uint64_t* expected_bytes, char blake2b_hex[129], const char** file_name);
/* BLAKE2b-512 of a byte string (host only; what `b2sum` prints) */
int mb200_blake2b512(const uint8_t* bytes, size_t len, uint8_t out[64]);
/* info[0..9] = n_inputs, n_aux, h_len, a_len, b_len, m, bytes consumed,
 * table bytes in HBM, window bits of the H+L table, window bits of the A table */
int mb200_params_info(const mb200_params* p, uint64_t info[10]);
void mb200_params_free(mb200_params* p);

/* bellman generate_random_parameters as the reference benches use it
 * (masp_proofs/benches/sapling.rs:24-36): a key of the given query lengths
 * whose points are PRNG scalars times the generators (masp_b200/synthetic.py
 * documents the derivation).  Writes mb200_params_synth_size(...) bytes of
 * Parameters encoding to host memory. */
size_t mb200_params_synth_size(uint32_t n_inputs, uint32_t h_len, uint32_t l_len, uint32_t a_len, uint32_t b_len);
int mb200_params_synthesize(uint64_t seed, uint32_t n_inputs, uint32_t h_len, uint32_t l_len, uint32_t a_len,
                            uint32_t b_len, uint8_t* out, size_t out_len);
/* n points (stream, start..start+n) * generator, uncompressed; group 1 = G1 (96 B), 2 = G2 (192 B) */
int mb200_synth_points(uint64_t seed, uint32_t stream, uint64_t start, size_t n, int group, uint8_t* out);

/* bellman create_proof(circuit, params, r, s) after synthesis, for a batch
 * (bellperson create_proof_batch), behind masp_proofs/src/sapling/prover.rs:
 * 116-117 / 201-202 / 251-252 and Proof::write at masp_proofs/src/prover.rs:190-193.
 *   rows            n_constraints + n_inputs (bellman appends one row per input)
 *   a/b/c_evals     n_proofs x rows scalars: the per-row evaluations <A_i,z>, <B_i,z>, <C_i,z>
 *   inputs          n_proofs x n_inputs scalars, inputs[0] = 1
 *   aux             n_proofs x n_aux scalars
 *   r, s            n_proofs scalars each; never drawn inside
 *   proofs_out      n_proofs x 192 bytes, host memory */
int mb200_prove_batch(const mb200_params* p, size_t n_proofs, size_t rows, const uint8_t* a_evals,
                      const uint8_t* b_evals, const uint8_t* c_evals, const uint8_t* inputs, const uint8_t* aux,
                      const uint8_t* r, const uint8_t* s, uint8_t* proofs_out);
/* Same, with every input already resident in device memory. */
int mb200_prove_batch_device(const mb200_params* p, size_t n_proofs, size_t rows, const void* a_evals,
                             const void* b_evals, const void* c_evals, const void* inputs, const void* aux,
                             const void* r, const void* s, uint8_t* proofs_out);

/* Streaming form of the two calls above: submit enqueues the whole batch and
 * returns a ticket without waiting; wait blocks until that batch is done and
 * its proofs are in proofs_out.  Batches submitted back to back overlap on the
 * device (the latency-bound tail of one runs under the head of the next).
 * Input and output buffers must stay valid until wait returns; with host
 * inputs (on_device = 0) pinned memory is needed for the copies to be
 * asynchronous.  Tickets must be waited on in any order, each exactly once. */
int mb200_prove_submit(const mb200_params* p, size_t n_proofs, size_t rows, const void* a_evals, const void* b_evals,
                       const void* c_evals, const void* inputs, const void* aux, const void* r, const void* s,
                       int on_device, uint8_t* proofs_out, uint64_t* ticket);
int mb200_prove_wait(uint64_t ticket);

/* ---- circuits: the step in front of create_proof ----------
 * The reference hands bellman a `Circuit` (masp_proofs/src/circuit/sapling.rs:
 * Spend :139-417, Output :419-596; circuit/convert.rs: Convert :29-128) and
 * bellman's ProvingAssignment runs its `synthesize` on the CPU for every proof,
 * evaluating all constraints as it goes.  Here a circuit is recorded ONCE into
 * an mb200_circuit (three CSR matrices, density bitmaps, structural hash);
 * per proof the host only computes the witness (inputs + aux) and the GPU
 * evaluates the rows (r1cs_eval kernel).
 *
 * A witness is a flat array of 32-byte little-endian fields; Jubjub points are
 * affine (u, v), Jubjub scalars plain integers, value a u64 in the low 8 bytes,
 * a Merkle path `depth` pairs (sibling scalar, is_right 0/1), leaf first:
 *   Spend   (circuit::sapling::Spend):  ak.u ak.v nsk g_d.u g_d.v asset_generator.u .v
 *                                       value rcv rcm ar anchor path...      (12 + 2 depth fields)
 *   Output  (circuit::sapling::Output): asset_identifier[32 bytes] asset_generator.u .v
 *                                       value rcv g_d.u g_d.v pk_d.u pk_d.v rcm esk    (11 fields)
 *   Convert (circuit::convert::Convert): asset_generator.u .v value rcv anchor path... (5 + 2 depth)
 */
#define MB200_CIRCUIT_SPEND 0
#define MB200_CIRCUIT_OUTPUT 1
#define MB200_CIRCUIT_CONVERT 2
typedef struct mb200_circuit mb200_circuit;
/* merkle_depth: the reference uses 32 (masp_primitives/src/sapling.rs SAPLING_COMMITMENT_TREE_DEPTH);
 * ignored for Output.  Host only: works without a device. */
int mb200_circuit_new(int kind, uint32_t merkle_depth, mb200_circuit** out);
void mb200_circuit_free(mb200_circuit* c);
/* info[0..9] = n_inputs (incl. ONE), n_aux, n_constraints, nnz(A), nnz(B), nnz(C), witness bytes,
 * ones in a_aux_density, b_input_density, b_aux_density */
int mb200_circuit_info(const mb200_circuit* c, uint64_t info[10]);
/* bellman TestConstraintSystem::hash of the recorded system, 64 hex digits + NUL: the value the
 * reference pins at circuit/sapling.rs:730-741, 1024-1045 and circuit/convert.rs:218-224 */
int mb200_circuit_hash(const mb200_circuit* c, char out_hex[65]);
/* LSB-first bitmaps, (n_aux+7)/8, (n_inputs+7)/8, (n_aux+7)/8 bytes: what mb200_params_load takes */
int mb200_circuit_densities(const mb200_circuit* c, uint8_t* a_aux, uint8_t* b_input, uint8_t* b_aux);
/* CSR export of matrix 0/1/2 = A/B/C (any pointer may be NULL): rowptr[n_constraints+1], col[nnz]
 * (input index, or 0x80000000 | aux index), coef[nnz x 32] canonical little-endian */
int mb200_circuit_matrix(const mb200_circuit* c, int which, uint32_t* rowptr, uint32_t* col, uint8_t* coef);
/* Circuit::synthesize, witness part only: n witnesses -> n x n_inputs and n x n_aux scalars.
 * n_threads <= 0: all host threads.  MB200_ESCALAR: a field is out of range; MB200_ESYNTH: the
 * reference's synthesis would have returned an error for this witness. */
int mb200_circuit_synthesize(const mb200_circuit* c, size_t n, const uint8_t* witnesses, uint8_t* inputs_out,
                             uint8_t* aux_out, int n_threads);
/* 1 when witnesses are generated eight at a time on AVX-512 IFMA lanes (csrc/circuits_simd.cpp), 0 when the
 * CPU lacks IFMA (or MB200_WITNESS_SCALAR=1) and the one-at-a-time generator runs.  Same bytes either way. */
int mb200_circuit_simd(void);
/* The Pedersen hash exactly as the circuits compute it (circuit/pedersen_hash.rs:19-103 in witness
 * form = masp_primitives::sapling::pedersen_hash): bits[0..6) are the personalization, the rest the
 * message, one byte per bit; the result is the affine (u, v) of the hash point.  The reference's
 * golden vectors for this function (masp_primitives/src/test_vectors/pedersen_hash_vectors.rs)
 * pin the product's window tables, Montgomery chains and Edwards maps. */
int mb200_pedersen_hash(const uint8_t* bits, size_t n_bits, uint8_t u_out[32], uint8_t v_out[32]);
/* The Merkle root a Spend / Convert witness leads to (the `cur` the circuit compares with the
 * anchor, circuit/sapling.rs:343-372, convert.rs:96-124); the witness's own anchor field is ignored.
 * For callers that hold a path but not the tree (tests, benches). */
int mb200_circuit_root(const mb200_circuit* c, const uint8_t* witness, uint8_t root_out[32]);
/* The per-row evaluations a = A z, b = B z, c = C z of n witnesses, computed on the device:
 * n x (n_constraints + n_inputs) scalars each, the trailing n_inputs rows being bellman's
 * input rows (a = input_i, b = c = 0).  What ProvingAssignment holds after synthesize. */
int mb200_circuit_rows(const mb200_circuit* c, size_t n, const uint8_t* inputs, const uint8_t* aux, uint8_t* a_out,
                       uint8_t* b_out, uint8_t* c_out);
/* Attach the circuit's matrices to a loaded key (uploads them once); the key's n_inputs / n_aux /
 * query lengths must match the circuit's counts and densities. */
int mb200_params_bind_circuit(mb200_params* p, const mb200_circuit* c);
/* create_proof for a batch given only the witnesses; needs a bound circuit.  inputs / aux as
 * produced by mb200_circuit_synthesize (host memory). */
int mb200_prove_batch_witness(const mb200_params* p, size_t n_proofs, const uint8_t* inputs, const uint8_t* aux,
                              const uint8_t* r, const uint8_t* s, uint8_t* proofs_out);

/* ---- verification: groth16::verify_proof as the reference calls it right
 * after proving (masp_proofs/src/sapling/prover.rs:148, :266), on the device, one thread per
 * proof.  Two forms:
 *   - mb200_set_option("verify", 1): every prove call checks its own proofs with the key's
 *     verifying key and the witnesses' public inputs before returning them; a failure makes the
 *     call return MB200_EVERIFY (the reference's Err(())).  Value 2 runs the same check but only
 *     counts (counters "verified", "verify_failed"): for measuring its cost on synthetic keys.
 *   - mb200_verify_batch: n proofs given as uncompressed points A (96) | B (192) | C (96) and
 *     n x n_inputs public-input scalars (inputs[0] = 1); ok_out[i] = 1 iff
 *     e(A,B) = e(alpha,beta) e(sum x_i IC_i, gamma) e(C, delta). */
int mb200_verify_batch(const mb200_params* p, size_t n, const uint8_t* proofs_uncompressed, const uint8_t* inputs,
                       uint8_t* ok_out);
/* The same for proofs as they travel: n x 192 bytes (Proof::write: A, B, C compressed).  Reading
 * them is bellman's Proof::read -- each point must decompress to a curve point of the prime-order
 * subgroup -- and a proof that does not read is simply not accepted (ok_out[i] = 0), as at the
 * verifier's call sites masp_proofs/src/sapling/verifier/single.rs:60, 77, 93. */
int mb200_verify_proofs(const mb200_params* p, size_t n, const uint8_t* proofs, const uint8_t* inputs, uint8_t* ok_out);
/* The randomised batch check of bellman's groth16::batch::Verifier (the reference's BatchValidator,
 * masp_proofs/src/sapling/verifier/batch.rs:24-31, 85-160): one verdict for the whole batch, n + 3
 * Miller loops and a single final exponentiation.  z: n x 16 bytes of caller-drawn randomness (the
 * 128-bit coefficients; drawn by the caller like r and s, so a run is reproducible).  *all_ok = 1
 * iff every proof reads and the combined equation holds; an invalid proof passes only with
 * probability about 2^-128 over z. */
int mb200_verify_proofs_batch(const mb200_params* p, size_t n, const uint8_t* proofs, const uint8_t* inputs,
                              const uint8_t* z, int* all_ok);

/* multiexp over arbitrary bases (ec-gpu-gen `multiexp`, full density);
 * result uncompressed.  No table is precomputed on this path. */
int mb200_msm_g1(const uint8_t* bases_uncompressed, const uint8_t* scalars, size_t n, uint8_t out_uncompressed[96]);
int mb200_msm_g2(const uint8_t* bases_uncompressed, const uint8_t* scalars, size_t n, uint8_t out_uncompressed[192]);
/* Split-MSM support: upload bases once, run on device-resident bases, get the
 * projective partial (XYZZ, 4 x 48 bytes of Montgomery limbs), add partials. */
int mb200_g1_bases_upload(const uint8_t* bases_uncompressed, size_t n, void** dev_bases);
int mb200_dev_free(void* dev_ptr);
int mb200_msm_g1_partial(const void* dev_bases, const uint8_t* scalars, size_t n, uint8_t out_partial[192]);
int mb200_g1_sum_partials(const uint8_t* partials, size_t count, uint8_t out_uncompressed[96]);
/* The same two steps without a host hop, for one-rank-per-GPU callers (BASELINE config 3 under
 * torch.distributed): the partial is written to device memory (dev_partial: 192 bytes, e.g. a
 * row of the tensor handed to ncclAllGather), and the gathered partials are added and encoded
 * on the device (SURVEY.md kernel K6, partial_point_allgather_add). */
int mb200_msm_g1_partial_device(const void* dev_bases, const uint8_t* scalars, size_t n, void* dev_partial);
int mb200_g1_sum_partials_device(const void* dev_partials, size_t count, uint8_t out_uncompressed[96]);
/* The single-process form: bases range-split over every opened device at upload; one call
 * reduces each range on its GPU, moves the 192-byte partials GPU -> GPU (NVLink peer copies) to
 * the primary device and adds them there. */
typedef struct mb200_g1_bases mb200_g1_bases;
int mb200_g1_bases_new(const uint8_t* bases_uncompressed, size_t n, mb200_g1_bases** out);
void mb200_g1_bases_free(mb200_g1_bases* b);
int mb200_msm_g1_bases(const mb200_g1_bases* b, const uint8_t* scalars, size_t n, uint8_t out_uncompressed[96]);

/* EvaluationDomain::{fft, ifft, coset_fft, icoset_fft} in place on host data
 * (2^log_n scalars). */
int mb200_ntt(uint8_t* data, unsigned log_n, int inverse, int coset);
/* The H-polynomial coefficients (bellman prover: 3 ifft, 3 coset_fft, a*b-c,
 * divide_by_z_on_coset, icoset_fft, drop the last): out = (m-1) scalars. */
int mb200_h_coeffs(const uint8_t* a_evals, const uint8_t* b_evals, const uint8_t* c_evals, size_t rows,
                   uint8_t* out);
/* elementwise product of two scalar vectors (host or device pointers) */
int mb200_fr_mul(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);
int mb200_fr_mul_device(const void* a, const void* b, size_t n, void* out);

/* knobs: "chunk" proofs per in-flight chunk; "streams" in-flight chunks per device; "verify" 0/1/2 self-check;
 * "profile" 0/1 event timing of the accumulation kernels (synchronises: measurement only);
 * "msm_slab" scalars per slab of a standalone MSM (default 2^22: slab k+1 uploads while slab k is reduced) */
int mb200_set_option(const char* name, long value);
/* counters: "launches" (kernels launched so far), "acc_launches",
 * "acc_us" (device time of the bucket-accumulation kernels, microseconds;
 * "acc_bytes" their algorithmic bytes, 128 N per G1 and 224 N per G2 instance;
 * "last_batch_us" device time of the last prove call (max over devices);
 * "devices"; "verified", "verify_failed";
 * acc_* only measured while option "profile" = 1) */
int mb200_get_counter(const char* name, double* value);
/* device self-test of the register-level field / curve arithmetic against
 * straightforward 64-bit code; returns the number of mismatches (0 = pass) */
int mb200_selftest(void);
/* throughput of Fp Montgomery multiplications on the whole chip, per second */
int mb200_bench_fpmul(double* muls_per_second);
/* single-warp latency, nanoseconds per dependent operation: mode 0 inlined Fp
 * multiply, 1 out-of-line Fp multiply, 2 XYZZ doubling, 3 XYZZ addition */
int mb200_bench_latency(int mode, double* ns_per_op);

const char* mb200_strerror(int code);
const char* mb200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
