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
 *   - one process drives one GPU (one rank per GPU under torch.distributed);
 *     calls are serialised internally.
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

#define MB200_PROOF_BYTES 192 /* GROTH_PROOF_SIZE, masp_primitives/src/transaction/components.rs:14-15 */

typedef struct mb200_params mb200_params;

/* Binds the library to one device (device_ids[0]; NULL / 0 = current device)
 * and creates its streams.  Fails with MB200_ECUDA when there is no usable
 * GPU: there is no CPU fallback. */
int mb200_init(const int* device_ids, int n_devices);
int mb200_shutdown(void);

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

/* knobs: "chunk" proofs per in-flight chunk; "streams" in-flight chunks */
int mb200_set_option(const char* name, long value);
/* counters: "launches" (kernels launched so far), "acc_launches",
 * "acc_us" (device time of the bucket-accumulation kernels, microseconds;
 * "acc_bytes" their algorithmic bytes, 128 N per G1 and 224 N per G2 instance;
 * "last_batch_us" device time of the last prove call;
 * only measured while option "profile" = 1) */
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
