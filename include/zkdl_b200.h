/* zkdl_b200.h — C ABI of the B200-native zkDL FC-layer prover (libzkdl_b200.so).
 *
 * The reference (SafeAILab/zkDL) has no FFI layer: its boundary is the C++ header API + ./demo.
 * These entry points are what the reference-named C++ shim classes in zkdl_b200/host/ (FrTensor, G1TensorJacobian,
 * Commitment, zkFC, zkReLU) and the ctypes harness bind; each cites the reference interface it replaces.
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers unless the name ends in _host.  Plain pointers and sizes only.
 *   - PODs are the reference's: Fr = 8 x u32 LE limbs (Montgomery R=2^256 unless stated), Fq = 12 x u32,
 *     G1 affine = {x,y}, G1 Jacobian = {x,y,z}, infinity <=> z == 0  (bls12-381.cuh:120,222,421-430).
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.  Results are valid after the
 *     stream is synchronised (the C++ shims synchronise to keep the reference's blocking semantics).
 *   - Every function returns 0 on success, ZKDL_ERR_* otherwise; zkdl_last_error() gives the message.
 *     ZKDL_ERR_DIM is raised exactly where the reference throws "Incompatible dimensions".
 *   - There is no CPU fallback: without a CUDA device every compute call fails with ZKDL_ERR_CUDA.
 */
#ifndef ZKDL_B200_H
#define ZKDL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t val[8]; } zkdl_fr_t;                 /* blstrs__scalar__Scalar, bls12-381.cuh:120 */
typedef struct { uint32_t val[12]; } zkdl_fq_t;                /* blstrs__fp__Fp, bls12-381.cuh:222 */
typedef struct { zkdl_fq_t x, y; } zkdl_g1_affine_t;           /* bls12-381.cuh:421-424 */
typedef struct { zkdl_fq_t x, y, z; } zkdl_g1_jacobian_t;      /* bls12-381.cuh:426-430 */

enum { ZKDL_OK = 0, ZKDL_ERR_DIM = 1, ZKDL_ERR_CUDA = 2, ZKDL_ERR_ARG = 3,
       ZKDL_ERR_NCCL = 4 /* reserved: the library itself issues no collective (the multi-GPU exchanges of zkdl_b200/parallel.py go
                            through torch.distributed / NCCL); kept so that a future in-library collective does not renumber */ };

const char* zkdl_last_error(void);
int zkdl_version(void);
/* Pre-sizes the scratch arenas of `stream` (and of the calling thread's side streams) so that the first proof does not
 * pay for cudaMalloc.  Optional: arenas also grow on demand and are warm after the first call. */
int zkdl_scratch_reserve(size_t bytes, void* stream);
/* Releases every scratch arena of the current device that holds no live allocation (arenas are otherwise kept for the life
 * of the process).  Synchronises the device.  Call between workloads of very different sizes, never from a timed region. */
int zkdl_scratch_release_all(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t zkdl_launch_count(void);
/* Per-kernel profiler for bench.py's roofline entries (no reference counterpart; the reference times with Timer around
 * whole operators, timer.cpp).  zkdl_prof_enable(1) clears the records and makes every hot-kernel launch record CUDA events
 * on its own stream; zkdl_prof_dump synchronises the device and writes one text line per kernel
 * "name launches total_ms algorithmic_bytes fr_products fq_products" into buf (returns the size needed). */
int zkdl_prof_enable(int on);
size_t zkdl_prof_dump(char* buf, size_t cap);

/* ------------------------------------------------------------------ Fr tensors (fr-tensor.cu) */
enum { ZKDL_OP_ADD = 0, ZKDL_OP_SUB = 1, ZKDL_OP_MUL = 2, ZKDL_OP_NEG = 3, ZKDL_OP_MONT = 4, ZKDL_OP_UNMONT = 5 };
/* Fr_elementwise_{add,sub,neg,mont,unmont,mont_mul} (fr-tensor.cu:20-80); b ignored for unary ops; out may alias a */
int zkdl_fr_elementwise(int op, const zkdl_fr_t* a, const zkdl_fr_t* b, zkdl_fr_t* out, size_t n, void* stream);
/* Fr_broadcast_{add,sub,mont_mul} (fr-tensor.cu:28-33,51-56,82-87); x by value from the host */
int zkdl_fr_broadcast(int op, const zkdl_fr_t* a, const zkdl_fr_t* x_host, zkdl_fr_t* out, size_t n, void* stream);
/* FrTensor::sum (fr-tensor.cu:240-292) -> out[0] */
int zkdl_fr_sum(const zkdl_fr_t* a, size_t n, zkdl_fr_t* out, void* stream);
/* Fr_me_step (fr-tensor.cu:399-409): out has (in_size+1)/2 entries */
int zkdl_fr_fold(const zkdl_fr_t* in, zkdl_fr_t* out, const zkdl_fr_t* x_host, size_t in_size, void* stream);
/* Fr_partial_me_step (fr-tensor.cu:420-432): out has window*ceil(in_size/(2*window)) entries */
int zkdl_fr_partial_fold(const zkdl_fr_t* in, zkdl_fr_t* out, const zkdl_fr_t* x_host, size_t in_size, size_t window, void* stream);
/* FrTensor::operator()(u) / Fr_me (fr-tensor.cu:295-300,411-418): out[0] = T(u); requires 2^(k-1) < n <= 2^k */
int zkdl_fr_me(const zkdl_fr_t* a, size_t n, const zkdl_fr_t* u_host, size_t k, zkdl_fr_t* out, void* stream);
/* FrTensor::partial_me / Fr_partial_me (fr-tensor.cu:370-374,434-443).  out must hold zkdl_partial_me_size() entries */
size_t zkdl_partial_me_size(size_t n, size_t k, size_t window);
int zkdl_fr_partial_me(const zkdl_fr_t* a, size_t n, const zkdl_fr_t* u_host, size_t k, size_t window, zkdl_fr_t* out, void* stream);

/* FrTensor::random / FrTensor::random_int (fr-tensor.cu:302-368): the reference's generator (curand XORWOW, one state per
 * element, curand_init(seed, index, 0)); same seed -> bit-identical tables.  random: 8 draws per element, top limb
 * % 0x73eda753 (value < p, read as Montgomery or plain by the caller); random_int: (draw & (2^num_bits - 1)) - 2^(num_bits-1). */
int zkdl_fr_random(zkdl_fr_t* out, size_t n, uint64_t seed, void* stream);
int zkdl_fr_random_int(zkdl_fr_t* out, uint32_t num_bits, size_t n, uint64_t seed, void* stream);

/* ------------------------------------------------------------------ sumchecks (proof.cu) */
/* inner_product_sumcheck (proof.cu:72-108): proof gets 3k+2 Fr */
int zkdl_ip_sumcheck(const zkdl_fr_t* a, const zkdl_fr_t* b, size_t n, const zkdl_fr_t* u_host, size_t k, zkdl_fr_t* proof, void* stream);
/* hadamard_product_sumcheck (proof.cu:110-150): proof gets 3k+2 Fr */
int zkdl_hp_sumcheck(const zkdl_fr_t* a, const zkdl_fr_t* b, size_t n, const zkdl_fr_t* u_host, const zkdl_fr_t* v_host, size_t k, zkdl_fr_t* proof, void* stream);
/* binary_sumcheck (proof.cu:152-200): proof gets 3k+1 Fr */
int zkdl_bin_sumcheck(const zkdl_fr_t* a, size_t n, const zkdl_fr_t* u_host, const zkdl_fr_t* v_host, size_t k, zkdl_fr_t* proof, void* stream);

/* The same three sumchecks with a Fiat-Shamir transcript on the device (no reference counterpart: the reference draws every
 * challenge from std::random_device up front, proof.cu:3-11, and its proofs bind to nothing).
 * state_in_host: 32 transcript bytes.  Round j: S <- SHA-256(S || c0 || c1 || c2) over the 32-byte little-endian limb images of
 * the round's proof elements; the fold challenge x_j = the digest's 8 little-endian u32 limbs, top limb % 0x73eda753 (the
 * random_vec recipe).  proof: same layout and field elements as zkdl_{ip,hp,bin}_sumcheck would give for those challenges
 * (for kind IP the challenges play the role of u, for HP / BIN of v; u_host is the eq point, fixed before round 0).
 * challenges[k] and state_out[8] (big-endian digest words) are DEVICE outputs. */
enum { ZKDL_FS_IP = 0, ZKDL_FS_HP = 1, ZKDL_FS_BIN = 2 };
int zkdl_sumcheck_fs(int kind, const zkdl_fr_t* a, const zkdl_fr_t* b, size_t n, const zkdl_fr_t* u_host, size_t k, const uint8_t* state_in_host,
                     zkdl_fr_t* proof, zkdl_fr_t* challenges, uint32_t* state_out, void* stream);

/* ------------------------------------------------------------------ zkFC / zkReLU forward (zkfc.cu, zkrelu.cu) */
/* float_to_Fr_kernel (zkfc.cu:63-88): round(x*2^16) -> signed Fr, NOT Montgomery; zero-pads to rows_out x cols_out */
int zkdl_float_to_fr(const float* fs, zkdl_fr_t* out, uint32_t rows_in, uint32_t rows_out, uint32_t cols_in, uint32_t cols_out, void* stream);
/* matrixMultiplyOptimized (zkfc.cu:6-47): C[rowsA x colsB] = A[rowsA x colsA] * B[colsA x colsB] over Fr (Montgomery) */
int zkdl_fr_matmul(const zkdl_fr_t* A, const zkdl_fr_t* B, zkdl_fr_t* C, size_t rowsA, size_t colsA, size_t colsB, void* stream);
/* A zkFC multiplies by the same weight matrix on every call (zkfc.cu:117-126): zkdl_mm_weights keeps its quantised
 * integer copy (int32 + byte planes for the int8 tensor cores) so that zkdl_fr_matmul_prepared does not re-derive it.
 * Results are identical to zkdl_fr_matmul's; W must be the table the copy was made from and must not change meanwhile. */
typedef struct zkdl_mm_weights zkdl_mm_weights;
int zkdl_mm_weights_create(const zkdl_fr_t* W, size_t rows, size_t cols, zkdl_mm_weights** out, void* stream);
int zkdl_mm_weights_destroy(zkdl_mm_weights* w);
int zkdl_fr_matmul_prepared(const zkdl_fr_t* A, const zkdl_fr_t* W, const zkdl_mm_weights* prep, zkdl_fr_t* C, size_t rowsA, void* stream);
/* One hidden layer of the forward pass, zkFC::operator() followed by zkReLU::operator() (demo.cu:30-34; zkfc.cu:117-126,
 * zkrelu.cu:44-52): Z = A W, then act / sign / packed decomposition of Z as zkdl_relu_packed writes them.  On the tensor-core
 * route the activation is applied to the exact integer accumulators in the product's epilogue; results are identical to
 * zkdl_fr_matmul_prepared + zkdl_relu_packed. */
int zkdl_fr_matmul_prepared_relu(const zkdl_fr_t* A, const zkdl_fr_t* W, const zkdl_mm_weights* prep, zkdl_fr_t* Z, size_t rowsA,
                                 zkdl_fr_t* act, zkdl_fr_t* sign, uint32_t* mag_packed, uint16_t* rem_packed, uint32_t* out_of_range, void* stream);
/* relu_kernel (zkrelu.cu:11-41): Z[n], sign[n], mag_bin[32n], rem_bin[16n].  Inputs outside +-2^47 (undefined in the
 * reference) give sign = 0, mag = 0 and are counted in *out_of_range (device u32, may be NULL). */
int zkdl_relu(const zkdl_fr_t* X, zkdl_fr_t* Z, zkdl_fr_t* sign, zkdl_fr_t* mag_bin, zkdl_fr_t* rem_bin, size_t n, uint32_t* out_of_range, void* stream);

/* Same decomposition with the auxiliary input kept bit-packed (48 bits instead of 1536 B per activation):
 * mag_packed[i] = the 32 mag_bin cells of element i (bit k = cell 32i+k), rem_packed[i] = its 16 rem_bin cells.
 * zkdl_relu_expand materialises the reference's 0/1 Fr tables from the packed words (either output may be NULL). */
int zkdl_relu_packed(const zkdl_fr_t* X, zkdl_fr_t* Z, zkdl_fr_t* sign, uint32_t* mag_packed, uint16_t* rem_packed, size_t n, uint32_t* out_of_range, void* stream);
int zkdl_relu_expand(const uint32_t* mag_packed, const uint16_t* rem_packed, zkdl_fr_t* mag_bin, zkdl_fr_t* rem_bin, size_t n, void* stream);

/* ------------------------------------------------------------------ G1 tensors (g1-tensor.cu) */
enum { ZKDL_G1_ADD = 0, ZKDL_G1_SUB = 1, ZKDL_G1_NEG = 2, ZKDL_G1_MADD = 3, ZKDL_G1_MSUB = 4 };
/* G1_jacobian_elementwise_{add,sub,madd,msub,minus} and the broadcast forms (g1-tensor.cu:169-302):
 * nb == n elementwise, nb == 1 broadcast.  For MADD/MSUB b points at affine points. */
int zkdl_g1_elementwise(int op, const zkdl_g1_jacobian_t* a, const void* b, size_t nb, zkdl_g1_jacobian_t* out, size_t n, void* stream);
/* G1_affine_to_jacobian (g1-tensor.cu:142-147) */
int zkdl_g1_affine_to_jacobian(const zkdl_g1_affine_t* a, zkdl_g1_jacobian_t* out, size_t n, void* stream);
/* G1_jacobian_elementwise_mul(_broadcast) (g1-tensor.cu:422-461): out[i] = [x[i]] P[i mod np], x = raw limbs as integer */
int zkdl_g1_mul(const zkdl_g1_jacobian_t* P, size_t np, const zkdl_fr_t* x, size_t n, zkdl_g1_jacobian_t* out, void* stream);
/* G1TensorJacobian::sum (g1-tensor.cu:368-420) -> out[0] */
int zkdl_g1_sum(const zkdl_g1_jacobian_t* a, size_t n, zkdl_g1_jacobian_t* out, void* stream);
/* G1TensorJacobian::operator()(u) / G1_me (g1-tensor.cu:463-491) -> out[0]; requires 2^(k-1) < n <= 2^k */
int zkdl_g1_me(const zkdl_g1_jacobian_t* a, size_t n, const zkdl_fr_t* u_host, size_t k, zkdl_g1_jacobian_t* out, void* stream);
/* normalise to affine-equivalent Jacobian (z = 1, or z = 0 for infinity); not in the reference (it has no inversion) */
int zkdl_g1_normalize(const zkdl_g1_jacobian_t* a, zkdl_g1_jacobian_t* out, size_t n, void* stream);

/* ------------------------------------------------------------------ fixed-base MSM engine (Commitment, commitment.cu) */
typedef struct zkdl_g1_table zkdl_g1_table;     /* opaque: affine window tables 2^(c*w) * G[i] resident in HBM */
/* Precompute tables for n generators (the reference's Commitment is a fixed generator set, demo.cu:81-82).
 * window_bits = 0 picks a default.  full = 1: all windows (no doublings on the proving path);
 * full = 0: bases only (plain Pippenger, used for the large-N MSM sweep). */
int zkdl_g1_table_create(const zkdl_g1_jacobian_t* points, size_t n, int window_bits, int full, zkdl_g1_table** out, void* stream);
int zkdl_g1_table_destroy(zkdl_g1_table* t);
size_t zkdl_g1_table_size(const zkdl_g1_table* t);
size_t zkdl_g1_table_bytes(const zkdl_g1_table* t);
/* Batched Pippenger MSM over shared bases: out[r] = sum_c [s[r*n + c]] G[c], r < m.
 * scalars_mont = 1: scalars are Montgomery Fr and are un-Montgomery'd first (Commitment::commit semantics,
 * commitment.cu:33); 0: the raw limbs are the integer scalar (G1Jacobian_mul semantics, g1-tensor.cu:422-430). */
int zkdl_msm(const zkdl_g1_table* t, const zkdl_fr_t* scalars, size_t m, int scalars_mont, zkdl_g1_jacobian_t* out, void* stream);
/* Commitment::commit (commitment.cu:29-41), intended semantics com[r] = (G * unmont(t[r,:])).sum()   */
int zkdl_commit(const zkdl_g1_table* gens, const zkdl_fr_t* t, size_t nt, zkdl_g1_jacobian_t* com, void* stream);
/* Commitment::me_open (commitment.cu:43-81): proof gets 3k+1 points, ret[0] the final scalar; requires n == 2^k == |gens| */
int zkdl_me_open(const zkdl_g1_table* gens, const zkdl_fr_t* t, size_t n, const zkdl_fr_t* u_host, size_t k,
                 zkdl_g1_jacobian_t* proof, zkdl_fr_t* ret, void* stream);
/* Commitment::open (commitment.cu:83-92): com_eval[0] = com(u_hi) (computed through com_table), then me_open on
 * t.partial_me(u_hi, |gens|).  u has ku entries; ceilLog2(ncom) of them (the last) are u_hi. */
int zkdl_open(const zkdl_g1_table* gens, const zkdl_g1_table* com_table, const zkdl_fr_t* t, size_t nt,
              const zkdl_fr_t* u_host, size_t ku, zkdl_g1_jacobian_t* com_eval, zkdl_g1_jacobian_t* proof, zkdl_fr_t* ret, void* stream);

/* ------------------------------------------------------------------ whole-layer provers (zkfc.cu:128-145, zkrelu.cu:79-100) */
/* zkFC::prove with injected challenges.  X[B*I], W[I*O], Z[B*O] Montgomery; B, I, O powers of two (padded).
 * proof_fr receives [ip sumcheck 3*log I + 2][Z(u) 1][open ret 1]; proof_g1 receives [com(u_hi) 1][me_open 3*log|G| + 1]
 * (SURVEY App. A.12).  Sizes: zkdl_zkfc_proof_sizes(). */
void zkdl_zkfc_proof_sizes(size_t B, size_t I, size_t O, size_t ngens, size_t* n_fr, size_t* n_g1);
int zkdl_zkfc_prove(const zkdl_fr_t* X, const zkdl_fr_t* W, const zkdl_fr_t* Z, size_t B, size_t I, size_t O,
                    const zkdl_g1_table* gens, const zkdl_g1_table* com_table,
                    const zkdl_fr_t* u_bs_host, const zkdl_fr_t* u_in_host, const zkdl_fr_t* u_out_host,
                    zkdl_fr_t* proof_fr, zkdl_g1_jacobian_t* proof_g1, void* stream);
/* zkReLU::prove with injected challenges.  n = |X| (power of two), L = log2 n.
 * proof_fr receives [bin(mag) 3(L+5)+1][mag.partial_me 32][bin(rem) 3(L+4)+1][rem.partial_me 16][hadamard 3L+2]. */
size_t zkdl_zkrelu_proof_size(size_t n);
int zkdl_zkrelu_prove(const zkdl_fr_t* X, const zkdl_fr_t* sign, const zkdl_fr_t* mag_bin, const zkdl_fr_t* rem_bin, size_t n,
                      const zkdl_fr_t* u_z_host, const zkdl_fr_t* v_z_host, const zkdl_fr_t* u_r_host, const zkdl_fr_t* v_r_host,
                      const zkdl_fr_t* u_rec_host, const zkdl_fr_t* u_hp_host, const zkdl_fr_t* v_hp_host,
                      zkdl_fr_t* proof_fr, void* stream);

/* zkReLU::prove on the packed auxiliary input: identical proof elements (the cells are exactly Scalar_ONE/ZERO,
 * zkrelu.cu:34-38), first three binary-sumcheck rounds and both partial_me calls read 48 bits per activation. */
int zkdl_zkrelu_prove_packed(const zkdl_fr_t* X, const zkdl_fr_t* sign, const uint32_t* mag_packed, const uint16_t* rem_packed, size_t n,
                             const zkdl_fr_t* u_z_host, const zkdl_fr_t* v_z_host, const zkdl_fr_t* u_r_host, const zkdl_fr_t* v_r_host,
                             const zkdl_fr_t* u_rec_host, const zkdl_fr_t* u_hp_host, const zkdl_fr_t* v_hp_host,
                             zkdl_fr_t* proof_fr, void* stream);

/* The same provers restricted to some of their independent parts (sub-layer partition over GPUs).  Only the
 * selected segments of the proof buffers are written; a rank that owns a part passes the full-size buffers and ships
 * its segments.  zkFC: ZKDL_FC_SUMCHECK = [ip sumcheck][Z(u)], ZKDL_FC_OPENING = [open ret] + all of proof_g1.
 * zkReLU: ZKDL_RELU_MAG = [bin(mag)][mag.partial_me], ZKDL_RELU_REM = [bin(rem)][rem.partial_me], ZKDL_RELU_HP = [hadamard].
 * W_int (may be NULL): the zkdl_mm_weights copy of W; when W is a table of 32-bit integers (quantised weights) the two
 * weights.partial_me passes (zkfc.cu:139, commitment.cu:88) fold the integers against eq tables: same field elements. */
#define ZKDL_FC_SUMCHECK 1u
#define ZKDL_FC_OPENING 2u
#define ZKDL_RELU_MAG 1u
#define ZKDL_RELU_REM 2u
#define ZKDL_RELU_HP 4u
int zkdl_zkfc_prove_parts(const zkdl_fr_t* X, const zkdl_fr_t* W, const zkdl_mm_weights* W_int, const zkdl_fr_t* Z, size_t B, size_t I, size_t O,
                          const zkdl_g1_table* gens, const zkdl_g1_table* com_table,
                          const zkdl_fr_t* u_bs_host, const zkdl_fr_t* u_in_host, const zkdl_fr_t* u_out_host,
                          zkdl_fr_t* proof_fr, zkdl_g1_jacobian_t* proof_g1, unsigned parts, void* stream);
int zkdl_zkrelu_prove_packed_parts(const zkdl_fr_t* X, const zkdl_fr_t* sign, const uint32_t* mag_packed, const uint16_t* rem_packed, size_t n,
                                   const zkdl_fr_t* u_z_host, const zkdl_fr_t* v_z_host, const zkdl_fr_t* u_r_host, const zkdl_fr_t* v_r_host,
                                   const zkdl_fr_t* u_rec_host, const zkdl_fr_t* u_hp_host, const zkdl_fr_t* v_hp_host,
                                   zkdl_fr_t* proof_fr, unsigned parts, void* stream);

/* ------------------------------------------------------------------ host helpers (proof.cu:3-31) */
/* random_vec with an injected seed: std::mt19937(seed), 8 draws per element, last % 1944954707 */
void zkdl_random_vec_host(uint32_t seed, size_t len, zkdl_fr_t* out_host);
uint32_t zkdl_ceil_log2(uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
