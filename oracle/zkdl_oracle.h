/* zkdl_oracle.h — CPU restatement of the zkDL reference's FC-layer proof path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under zkdl_b200/ (the product) may include, link or call this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, as the checker.
 *
 * Parity pinning: the reference ships NO tests, golden vectors or CPU path, so this
 * restatement is pinned by (i) fixtures produced by the reference's own CUDA code on a B200
 * (oracle/_ref/ref_harness, see oracle/ref_harness.cu and tests/golden/README.md), (ii) a Python big-int
 * model (tests/test_oracle_bigint.py), (iii) the protocol identities of SURVEY.md §4.
 *
 * Every function cites the reference file:line it restates.  Layouts are the reference's PODs:
 *   Fr  = 8 x u32 little-endian limbs  (bls12-381.cuh:120)  == 4 x u64 on x86
 *   Fq  = 12 x u32                      (bls12-381.cuh:222)  == 6 x u64
 *   G1 affine = {x,y} (96 B), Jacobian = {x,y,z} (144 B), infinity <=> z == 0 (bls12-381.cuh:419-430)
 */
#ifndef ZKDL_ORACLE_H
#define ZKDL_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } ofr_t;
typedef struct { uint64_t l[6]; } ofq_t;
typedef struct { ofq_t x, y; } og1a_t;
typedef struct { ofq_t x, y, z; } og1j_t;

/* ---- Fr (bls12-381.cu:213-608) ---- */
void orc_fr_add(const ofr_t* a, const ofr_t* b, ofr_t* out, size_t n);
void orc_fr_sub(const ofr_t* a, const ofr_t* b, ofr_t* out, size_t n);
void orc_fr_mul(const ofr_t* a, const ofr_t* b, ofr_t* out, size_t n);   /* Montgomery product */
void orc_fr_mont(const ofr_t* a, ofr_t* out, size_t n);
void orc_fr_unmont(const ofr_t* a, ofr_t* out, size_t n);
void orc_fr_neg(const ofr_t* a, ofr_t* out, size_t n);
void orc_fr_bcast(const ofr_t* a, const ofr_t* x, int op, ofr_t* out, size_t n); /* op 0 add,1 sub,2 mul */
void orc_fr_sum(const ofr_t* a, size_t n, ofr_t* out);                    /* fr-tensor.cu:240-292 */

/* ---- multilinear folds (fr-tensor.cu:295-300,370-443) ---- */
void orc_fr_me(const ofr_t* a, size_t n, const ofr_t* u, size_t k, ofr_t* out);
/* returns the output length; out must hold n elements */
size_t orc_fr_partial_me(const ofr_t* a, size_t n, const ofr_t* u, size_t k, size_t window, ofr_t* out);

/* ---- sumchecks (proof.cu:55-200); proof lengths 3k+2, 3k+2, 3k+1 ---- */
void orc_ip_sumcheck(const ofr_t* a, const ofr_t* b, size_t n, const ofr_t* u, size_t k, ofr_t* proof);
void orc_hp_sumcheck(const ofr_t* a, const ofr_t* b, size_t n, const ofr_t* u, const ofr_t* v, size_t k, ofr_t* proof);
void orc_bin_sumcheck(const ofr_t* a, size_t n, const ofr_t* u, const ofr_t* v, size_t k, ofr_t* proof);

/* ---- host helpers (proof.cu:3-31) ---- */
void orc_random_vec(uint32_t seed, size_t len, ofr_t* out);               /* mt19937(seed) recipe */
uint32_t orc_ceil_log2(uint32_t n);

/* ---- zkFC / zkReLU forward (zkfc.cu:6-126, zkrelu.cu:11-62) ---- */
void orc_float_to_fr(const float* fs, ofr_t* frs, uint32_t fs_rows, uint32_t frs_rows, uint32_t fs_cols, uint32_t frs_cols);
void orc_fr_matmul(const ofr_t* A, const ofr_t* B, ofr_t* C, size_t rowsA, size_t colsA, size_t colsB);
/* returns number of inputs outside the +-2^47 range (reference behaviour undefined there, App. B9) */
size_t orc_relu(const ofr_t* X, ofr_t* Z, ofr_t* sign, ofr_t* mag_bin, ofr_t* rem_bin, size_t n);

/* ---- Fq / G1 (bls12-381.cu:612-1015,1331-1435; g1-tensor.cu:368-491) ---- */
void orc_fq_mul(const ofq_t* a, const ofq_t* b, ofq_t* out, size_t n);
void orc_g1_double(const og1j_t* a, og1j_t* out, size_t n);
void orc_g1_add(const og1j_t* a, const og1j_t* b, og1j_t* out, size_t n);
void orc_g1_add_mixed(const og1j_t* a, const og1a_t* b, og1j_t* out, size_t n);
void orc_g1_neg(const og1j_t* a, og1j_t* out, size_t n);
/* out[i] = [x[i]] P[i mod np] by the literal 256-step LSB-first ladder on the raw limbs of x (g1-tensor.cu:422-445) */
void orc_g1_mul(const og1j_t* P, size_t np, const ofr_t* x, size_t n, og1j_t* out);
/* same group element as orc_g1_mul (any representative), windowed: for large oracle inputs */
void orc_g1_mul_fast(const og1j_t* P, size_t np, const ofr_t* x, size_t n, og1j_t* out);
void orc_g1_sum(const og1j_t* a, size_t n, og1j_t* out);                  /* g1-tensor.cu:368-420, same tree order */
void orc_g1_me(const og1j_t* a, size_t n, const ofr_t* u, size_t k, og1j_t* out); /* g1-tensor.cu:463-491 */
void orc_g1_to_affine(const og1j_t* a, og1a_t* out, uint8_t* is_inf, size_t n); /* (X/Z^2, Y/Z^3), Montgomery form */
int  orc_g1_eq(const og1j_t* a, const og1j_t* b);                         /* same point (projective compare) */
int  orc_g1_on_curve(const og1j_t* a);
void orc_g1_generator(og1j_t* out);                                       /* g1-tensor.cuh:28-63 */

/* ---- Commitment (commitment.cu:29-92) ---- */
/* intended semantics: com[r] = (G * unmont(t[r,:])).sum()  (SURVEY §0 fact 5, §8c); t in Montgomery form */
void orc_commit(const og1j_t* G, size_t ng, const ofr_t* t, size_t nt, og1j_t* com, int fast);
/* as written, incl. the sum_axis_n_optimized row-mixing bug; only rows 64*b are defined; needs m%64==0, ng%64==0 */
void orc_commit_as_written(const og1j_t* G, size_t ng, const ofr_t* t, size_t nt, og1j_t* com);
/* me_open: proof gets 3k+1 points; returns final scalar in *ret */
void orc_me_open(const ofr_t* t, const og1j_t* G, size_t n, const ofr_t* u, size_t k, og1j_t* proof, ofr_t* ret, int fast);

/* ---- CPU baseline (BASELINE.md "Baseline B"): Pippenger MSM, OpenMP.  Same group element as orc_commit(m=1). ---- */
void orc_msm_pippenger(const og1a_t* bases, const ofr_t* scalars_plain, size_t n, og1j_t* out, int threads);
int  orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
