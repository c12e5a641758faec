// prove.cu — whole-layer provers: zkFC::prove (/root/reference/zkfc.cu:128-145) and zkReLU::prove
// (/root/reference/zkrelu.cu:79-100) with injected challenges, composed from the kernels in fr_kernels.cu / msm.cu.
// All work is enqueued on one stream with stream-ordered scratch memory: no host synchronisation, no cudaMalloc.
#include "common.cuh"
#include "g1.cuh"
#include "../../include/zkdl_b200.h"

struct zkdl_g1_table;
namespace zk {
int open_run(const zkdl_g1_table* gens, const zkdl_g1_table* com_table, const Fr* t, size_t nt, const zkdl_mm_weights* t_int,
             const zkdl_fr_t* u_host, size_t ku, G1Jac* com_eval, G1Jac* proof, Fr* ret, cudaStream_t st);
bool mmw_usable(const zkdl_mm_weights* p, size_t n);
int wfold_cols(const zkdl_mm_weights* p, const zkdl_fr_t* u_host, size_t k, Fr* out, cudaStream_t st);
}
using namespace zk;

static size_t ilog2(size_t v) { size_t l = 0; while (((size_t)1 << l) < v) ++l; return l; }

extern "C" {

void zkdl_zkfc_proof_sizes(size_t B, size_t I, size_t O, size_t ngens, size_t* n_fr, size_t* n_g1) {
  (void)B; (void)O;
  if (n_fr) *n_fr = 3 * ilog2(I) + 2 + 1 + 1;
  if (n_g1) *n_g1 = 1 + 3 * ilog2(ngens) + 1;
}

int zkdl_zkfc_prove(const zkdl_fr_t* X, const zkdl_fr_t* W, const zkdl_fr_t* Z, size_t B, size_t I, size_t O,
                    const zkdl_g1_table* gens, const zkdl_g1_table* com_table,
                    const zkdl_fr_t* u_bs_host, const zkdl_fr_t* u_in_host, const zkdl_fr_t* u_out_host,
                    zkdl_fr_t* proof_fr, zkdl_g1_jacobian_t* proof_g1, void* stream) {
  return zkdl_zkfc_prove_parts(X, W, nullptr, Z, B, I, O, gens, com_table, u_bs_host, u_in_host, u_out_host, proof_fr, proof_g1,
                               ZKDL_FC_SUMCHECK | ZKDL_FC_OPENING, stream);
}

int zkdl_zkfc_prove_parts(const zkdl_fr_t* X, const zkdl_fr_t* W, const zkdl_mm_weights* W_int, const zkdl_fr_t* Z, size_t B, size_t I, size_t O,
                          const zkdl_g1_table* gens, const zkdl_g1_table* com_table,
                          const zkdl_fr_t* u_bs_host, const zkdl_fr_t* u_in_host, const zkdl_fr_t* u_out_host,
                          zkdl_fr_t* proof_fr, zkdl_g1_jacobian_t* proof_g1, unsigned parts, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ZK_REQUIRE(X && W && Z && gens && com_table && proof_fr && proof_g1, ZK_ERR_ARG, "null argument");
  ZK_REQUIRE(parts && !(parts & ~(ZKDL_FC_SUMCHECK | ZKDL_FC_OPENING)), ZK_ERR_ARG, "bad parts mask");
  size_t kb = ilog2(B), ki = ilog2(I), ko = ilog2(O);
  ZK_REQUIRE(((size_t)1 << kb) == B && ((size_t)1 << ki) == I && ((size_t)1 << ko) == O, ZK_ERR_DIM, "Incompatible dimensions 1");
  int rc;
  // The matmul sumcheck and the commitment opening are independent: the sumcheck part runs on a side stream.
  SideStream& ss = side_stream(1, st);
  ForkScope fs(ss, st);
  if ((rc = fs.fork())) return rc;
  void* sst = reinterpret_cast<void*>(ss.stream);
  size_t nip = 3 * ki + 2;
  if (parts & ZKDL_FC_SUMCHECK) {
    Scratch Xr, Wr;
    if ((rc = Xr.alloc(sizeof(Fr) * I, ss.stream))) return rc;
    if ((rc = Wr.alloc(sizeof(Fr) * I, ss.stream))) return rc;
    // X.partial_me(u_bs, inputSize), weights.partial_me(u_out_dim, 1)   (zkfc.cu:139)
    if ((rc = zkdl_fr_partial_me(X, B * I, u_bs_host, kb, I, Xr.as<zkdl_fr_t>(), sst))) return rc;
    if (mmw_usable(W_int, I * O)) rc = wfold_cols(W_int, u_out_host, ko, Wr.as<Fr>(), ss.stream);       // same field elements from the integers
    else rc = zkdl_fr_partial_me(W, I * O, u_out_host, ko, 1, Wr.as<zkdl_fr_t>(), sst);
    if (rc) return rc;
    if ((rc = zkdl_ip_sumcheck(Xr.as<zkdl_fr_t>(), Wr.as<zkdl_fr_t>(), I, u_in_host, ki, proof_fr, sst))) return rc;
    // Z(u_out || u_bs)   (zkfc.cu:141-143)
    zkdl_fr_t uz[64];
    ZK_REQUIRE(ko + kb <= 64, ZK_ERR_DIM, "Incompatible dimensions");
    for (size_t i = 0; i < ko; ++i) uz[i] = u_out_host[i];
    for (size_t i = 0; i < kb; ++i) uz[ko + i] = u_bs_host[i];
    if ((rc = zkdl_fr_me(Z, B * O, uz, ko + kb, proof_fr + nip, sst))) return rc;
  }
  // generators.open(weights, com, u_out || u_in)   (zkfc.cu:144)
  if (parts & ZKDL_FC_OPENING) {
    zkdl_fr_t uo[64];
    ZK_REQUIRE(ko + ki <= 64, ZK_ERR_DIM, "Incompatible dimensions");
    for (size_t i = 0; i < ko; ++i) uo[i] = u_out_host[i];
    for (size_t i = 0; i < ki; ++i) uo[ko + i] = u_in_host[i];
    rc = open_run(gens, com_table, reinterpret_cast<const Fr*>(W), I * O, W_int, uo, ko + ki, reinterpret_cast<G1Jac*>(proof_g1),
                  reinterpret_cast<G1Jac*>(proof_g1 + 1), reinterpret_cast<Fr*>(proof_fr + nip + 1), st);
    if (rc) return rc;
  }
  return fs.join();
}

size_t zkdl_zkrelu_proof_size(size_t n) {
  size_t L = ilog2(n);
  return (3 * (L + 5) + 1) + 32 + (3 * (L + 4) + 1) + 16 + (3 * L + 2);
}

int zkdl_zkrelu_prove(const zkdl_fr_t* X, const zkdl_fr_t* sign, const zkdl_fr_t* mag_bin, const zkdl_fr_t* rem_bin, size_t n,
                      const zkdl_fr_t* u_z_host, const zkdl_fr_t* v_z_host, const zkdl_fr_t* u_r_host, const zkdl_fr_t* v_r_host,
                      const zkdl_fr_t* u_rec_host, const zkdl_fr_t* u_hp_host, const zkdl_fr_t* v_hp_host,
                      zkdl_fr_t* proof_fr, void* stream) {
  ZK_REQUIRE(X && sign && mag_bin && rem_bin && proof_fr, ZK_ERR_ARG, "null argument");
  size_t L = ilog2(n);
  ZK_REQUIRE(((size_t)1 << L) == n && L >= 1, ZK_ERR_DIM, "Incompatible dimensions");
  int rc;
  zkdl_fr_t* p = proof_fr;
  if ((rc = zkdl_bin_sumcheck(mag_bin, 32 * n, u_z_host, v_z_host, L + 5, p, stream))) return rc;      // zkrelu.cu:91
  p += 3 * (L + 5) + 1;
  if ((rc = zkdl_fr_partial_me(mag_bin, 32 * n, u_rec_host, L, 32, p, stream))) return rc;              // zkrelu.cu:92
  p += 32;
  if ((rc = zkdl_bin_sumcheck(rem_bin, 16 * n, u_r_host, v_r_host, L + 4, p, stream))) return rc;      // zkrelu.cu:93
  p += 3 * (L + 4) + 1;
  if ((rc = zkdl_fr_partial_me(rem_bin, 16 * n, u_rec_host, L, 16, p, stream))) return rc;              // zkrelu.cu:94
  p += 16;
  if ((rc = zkdl_hp_sumcheck(X, sign, n, u_hp_host, v_hp_host, L, p, stream))) return rc;              // zkrelu.cu:99
  return ZK_OK;
}

}  // extern "C"
