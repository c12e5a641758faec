// zkdl.hpp — host-side C++ mirror of the reference's public API over the C ABI (include/zkdl_b200.h).
//
// Same names, argument meaning and error behaviour as the reference headers, so demo.cu-style host code compiles
// against it unchanged:  FrTensor (/root/reference/fr-tensor.cuh:62-133), G1TensorAffine / G1TensorJacobian
// (g1-tensor.cuh:65-174), Commitment (commitment.cuh:8-23), the proof.cuh free functions (proof.cuh:13-34),
// zkFC (zkfc.cuh:16-33), zkReLU (zkrelu.cuh:10-20), Timer (timer.hpp).  No kernels live here: every operation is one
// or a few calls into libzkdl_b200.so followed by a stream synchronisation, which preserves the reference's blocking
// semantics ("every op is synchronous").  Shape misuse throws
// std::runtime_error("Incompatible dimensions") exactly where the reference does; unlike the reference, CUDA errors
// are not ignored: they throw std::runtime_error with the library's message.
//
// Additions over the reference (it discards every proof): zkFC::last_proof_fr()/last_proof_g1() and
// zkReLU::last_proof() expose the proof elements of the most recent prove(); set_challenge_seed() makes random_vec
// reproducible (the reference seeds std::mt19937 from std::random_device on every call, proof.cu:5-6).
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <climits>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <memory>
#include <random>
#include <stdexcept>
#include <utility>
#include <vector>
#include "../../include/zkdl_b200.h"

typedef unsigned int uint;
typedef zkdl_fr_t Fr_t;
typedef zkdl_fq_t Fp_t;
typedef zkdl_g1_affine_t G1Affine_t;
typedef zkdl_g1_jacobian_t G1Jacobian_t;

namespace zkdl_host {
inline void check(int rc) {
  if (rc == ZKDL_OK) return;
  if (rc == ZKDL_ERR_DIM) throw std::runtime_error("Incompatible dimensions");
  throw std::runtime_error(std::string("zkdl_b200: ") + zkdl_last_error());
}
inline void cuda_check(cudaError_t e) { if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e)); }
// Every shim operation runs on the calling thread's current stream (default: the legacy default stream, exactly like
// the reference) and synchronises it before returning.  set_thread_stream() lets independent proofs run concurrently
// from several host threads (demo.cpp proves the layers that way); memory then comes from the stream-ordered pool.
cudaStream_t cur_stream();
void set_thread_stream(cudaStream_t s);
inline void* st() { return reinterpret_cast<void*>(cur_stream()); }
inline void sync() { cuda_check(cudaStreamSynchronize(cur_stream())); }
void* dev_alloc_bytes(size_t bytes);
void dev_free(void* p);
template <class T> T* dev_alloc(size_t n) { return static_cast<T*>(dev_alloc_bytes(sizeof(T) * (n ? n : 1))); }
void copy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);        // on cur_stream(), synchronised
uint32_t next_challenge_seed();        // random_device, or a counter when set_challenge_seed() was called
// zkFC::prove / zkReLU::prove return nothing observable in the reference (the proof vectors are dropped, zkfc.cu:139-144).
// With async proves ON they only ENQUEUE the proof on the calling thread's current stream (no allocation, no host
// synchronisation after the object's first proof); last_proof_*() synchronises that stream and downloads on demand.  A host
// loop can thus spread independent layer proofs over several streams from ONE thread (demo.cpp).  Default: OFF (a prove()
// call returns with the proof finished, like every other shim call).
void set_async_prove(bool on);
bool async_prove();
}  // namespace zkdl_host

void set_challenge_seed(uint32_t seed);   // 0 restores std::random_device
void set_thread_challenge_seed(uint32_t seed);   // per-thread seed stream (deterministic proofs from concurrent threads); 0 = use the global one

// ---- constants (g1-tensor.cuh:28-63, bls12-381.cu:3-11)
extern const Fp_t G1_generator_x_mont, G1_generator_y_mont, G1_ONE;
extern const G1Affine_t G1Affine_generator;
extern const G1Jacobian_t G1Jacobian_generator;
extern const Fr_t Fr_ONE_mont, Fr_ZERO;

std::ostream& operator<<(std::ostream& os, const Fr_t& x);
std::ostream& operator<<(std::ostream& os, const Fp_t& x);
std::ostream& operator<<(std::ostream& os, const G1Affine_t& g);
std::ostream& operator<<(std::ostream& os, const G1Jacobian_t& g);

class G1TensorAffine; class G1TensorJacobian; class Commitment; class zkFC; class zkReLU;

// ------------------------------------------------------------------------------------------------ FrTensor
class FrTensor {
 public:
  Fr_t* gpu_data;                    // (private in the reference; friends there, public here)
  const uint size;
  FrTensor(uint size);
  FrTensor(uint size, const Fr_t* cpu_data);
  FrTensor(const FrTensor& t);
  ~FrTensor();
  Fr_t operator()(uint idx) const;
  FrTensor operator+(const FrTensor& t) const;
  FrTensor operator+(const Fr_t& x) const;
  FrTensor& operator+=(const FrTensor& t);
  FrTensor& operator+=(const Fr_t& x);
  FrTensor operator-() const;
  FrTensor operator-(const FrTensor& t) const;
  FrTensor operator-(const Fr_t& x) const;
  FrTensor& operator-=(const FrTensor& t);
  FrTensor& operator-=(const Fr_t& x);
  FrTensor& mont();
  FrTensor& unmont();
  FrTensor operator*(const FrTensor& t) const;
  FrTensor operator*(const Fr_t& x) const;
  FrTensor& operator*=(const FrTensor& t);
  FrTensor& operator*=(const Fr_t& x);
  Fr_t sum() const;
  Fr_t operator()(const std::vector<Fr_t>& u) const;
  std::pair<FrTensor, FrTensor> split(uint window_size) const;
  FrTensor partial_me(std::vector<Fr_t> u, uint window_size) const;
  static FrTensor random_int(uint size, uint num_bits);
  static FrTensor random(uint size);
};
std::ostream& operator<<(std::ostream& os, const FrTensor& A);
Fr_t Fr_me(const FrTensor& t, std::vector<Fr_t>::const_iterator begin, std::vector<Fr_t>::const_iterator end);
FrTensor Fr_partial_me(const FrTensor& t, std::vector<Fr_t>::const_iterator begin, std::vector<Fr_t>::const_iterator end, uint window_size);

// ------------------------------------------------------------------------------------------------ G1 tensors
class G1Tensor { public: const uint size; G1Tensor(uint size) : size(size) {} };

class G1TensorAffine : public G1Tensor {
 public:
  G1Affine_t* gpu_data;
  G1TensorAffine(const G1TensorAffine&);
  G1TensorAffine(uint size);
  G1TensorAffine(uint size, const G1Affine_t&);
  G1TensorAffine(uint size, const G1Affine_t* cpu_data);
  ~G1TensorAffine();
  G1Affine_t operator()(uint idx) const;
  G1TensorAffine operator-() const;
};

class G1TensorJacobian : public G1Tensor {
 public:
  G1Jacobian_t* gpu_data;
  G1TensorJacobian(const G1TensorJacobian&);
  G1TensorJacobian(uint size);
  G1TensorJacobian(uint size, const G1Jacobian_t&);
  G1TensorJacobian(uint size, const G1Jacobian_t* cpu_data);
  G1TensorJacobian(const G1TensorAffine& affine_tensor);
  ~G1TensorJacobian();
  G1Jacobian_t operator()(uint) const;
  G1TensorJacobian operator-() const;
  G1TensorJacobian operator+(const G1TensorJacobian&) const;
  G1TensorJacobian operator+(const G1TensorAffine&) const;
  G1TensorJacobian operator+(const G1Jacobian_t&) const;
  G1TensorJacobian operator+(const G1Affine_t&) const;
  G1TensorJacobian& operator+=(const G1TensorJacobian&);
  G1TensorJacobian& operator+=(const G1TensorAffine&);
  G1TensorJacobian& operator+=(const G1Jacobian_t&);
  G1TensorJacobian& operator+=(const G1Affine_t&);
  G1TensorJacobian operator-(const G1TensorJacobian&) const;
  G1TensorJacobian operator-(const G1TensorAffine&) const;
  G1TensorJacobian operator-(const G1Jacobian_t&) const;
  G1TensorJacobian operator-(const G1Affine_t&) const;
  G1TensorJacobian& operator-=(const G1TensorJacobian&);
  G1TensorJacobian& operator-=(const G1TensorAffine&);
  G1TensorJacobian& operator-=(const G1Jacobian_t&);
  G1TensorJacobian& operator-=(const G1Affine_t&);
  G1Jacobian_t sum() const;
  G1TensorJacobian operator*(const FrTensor&) const;
  G1TensorJacobian& operator*=(const FrTensor&);
  G1Jacobian_t operator()(const std::vector<Fr_t>& u) const;
  // fixed-base window tables of these points: built on first use, shared by copies of the tensor (same points), and
  // dropped by an object whenever ITS points change
  const zkdl_g1_table* table() const;
  void invalidate_table() const;
 private:
  struct TableHolder { zkdl_g1_table* t = nullptr; ~TableHolder(); };
  mutable std::shared_ptr<TableHolder> table_;
  G1TensorJacobian binary(int op, const void* b, size_t nb) const;
  G1TensorJacobian& binary_inplace(int op, const void* b, size_t nb);
};
G1Jacobian_t G1_me(const G1TensorJacobian& t, std::vector<Fr_t>::const_iterator begin, std::vector<Fr_t>::const_iterator end);

// ------------------------------------------------------------------------------------------------ Commitment
class Commitment : public G1TensorJacobian {
 public:
  using G1TensorJacobian::G1TensorJacobian;
  using G1TensorJacobian::operator+;
  using G1TensorJacobian::operator-;
  using G1TensorJacobian::operator*;
  using G1TensorJacobian::operator*=;
  G1TensorJacobian commit(const FrTensor& t) const;
  Fr_t open(const FrTensor& t, const G1TensorJacobian& c, const std::vector<Fr_t>& u) const;
  static Fr_t me_open(const FrTensor& t, const Commitment& generators, std::vector<Fr_t>::const_iterator begin,
                      std::vector<Fr_t>::const_iterator end, std::vector<G1Jacobian_t>& proof);
  // open() with the proof elements returned instead of discarded: proof = [com(u_hi)] ++ me_open proof
  Fr_t open_with_proof(const FrTensor& t, const G1TensorJacobian& c, const std::vector<Fr_t>& u, std::vector<G1Jacobian_t>& proof) const;
};

// ------------------------------------------------------------------------------------------------ proof.cuh
std::vector<Fr_t> random_vec(uint len);
uint ceilLog2(uint num);
template <typename T> std::vector<T> concatenate(const std::vector<std::vector<T>>& vecs) {
  std::vector<T> r; for (const auto& v : vecs) r.insert(r.end(), v.begin(), v.end()); return r;
}
void Fr_ip_sc(const FrTensor& a, const FrTensor& b, std::vector<Fr_t>::const_iterator begin, std::vector<Fr_t>::const_iterator end, std::vector<Fr_t>& proof);
std::vector<Fr_t> inner_product_sumcheck(const FrTensor& a, const FrTensor& b, std::vector<Fr_t> u);
void Fr_hp_sc(const FrTensor& a, const FrTensor& b, std::vector<Fr_t>::const_iterator u_begin, std::vector<Fr_t>::const_iterator u_end,
              std::vector<Fr_t>::const_iterator v_begin, std::vector<Fr_t>::const_iterator v_end, std::vector<Fr_t>& proof);
std::vector<Fr_t> hadamard_product_sumcheck(const FrTensor& a, const FrTensor& b, std::vector<Fr_t> u, std::vector<Fr_t> v);
void Fr_bin_sc(const FrTensor& a, std::vector<Fr_t>::const_iterator u_begin, std::vector<Fr_t>::const_iterator u_end,
               std::vector<Fr_t>::const_iterator v_begin, std::vector<Fr_t>::const_iterator v_end, std::vector<Fr_t>& proof);
std::vector<Fr_t> binary_sumcheck(const FrTensor& a, std::vector<Fr_t> u, std::vector<Fr_t> v);

// ------------------------------------------------------------------------------------------------ zkFC / zkReLU
class zkFC {
 private:
  FrTensor weights;
  G1TensorJacobian com;
  mutable std::vector<Fr_t> proof_fr_;
  mutable std::vector<G1Jacobian_t> proof_g1_;
  struct ProofBuf { Fr_t* fr = nullptr; G1Jacobian_t* g1 = nullptr; size_t nfr = 0, ng1 = 0; cudaStream_t stream = 0; bool pending = false; ~ProofBuf(); };
  mutable std::shared_ptr<ProofBuf> pbuf_;       // device proof buffers, allocated by the first prove() and reused
  void fetch_proof() const;
  struct MMHolder { zkdl_mm_weights* w = nullptr; ~MMHolder(); };
  mutable std::shared_ptr<MMHolder> mm_;         // quantised integer copy of `weights` for operator(), shared by copies
 public:
  const uint inputSize;
  const uint outputSize;
  zkFC(uint input_size, uint output_size, const FrTensor& t, const Commitment& generators);
  FrTensor operator()(const FrTensor& X) const;
  void prove(const FrTensor& X, const FrTensor& Z, Commitment& generators) const;
  static zkFC from_float_gpu_ptr(uint input_size, uint output_size, float* float_gpu_ptr, const Commitment& generators);
  static FrTensor load_float_gpu_input(uint batch_size, uint input_dim, float* input_ptr);
  const std::vector<Fr_t>& last_proof_fr() const { fetch_proof(); return proof_fr_; }
  const std::vector<G1Jacobian_t>& last_proof_g1() const { fetch_proof(); return proof_g1_; }
  const G1TensorJacobian& commitment() const { return com; }
};

class zkReLU {
 protected:
  FrTensor* sign_ptr = nullptr;
  FrTensor* mag_bin_ptr = nullptr;
  FrTensor* rem_bin_ptr = nullptr;
  void reset_ptrs(uint size);
  std::vector<Fr_t> proof_;
  Fr_t* dev_proof_ = nullptr; size_t dev_proof_n_ = 0; cudaStream_t proof_stream_ = 0; bool proof_pending_ = false;
  void fetch_proof();
  uint32_t* mag_packed_ = nullptr;               // the same auxiliary input, 48 bits per activation (device)
  uint16_t* rem_packed_ = nullptr;
  uint n_ = 0;
 public:
  // true (default): operator() also materialises the reference's 0/1 Fr tables (sign/mag_bin/rem_bin pointers valid, as in
  // the reference).  false: only the packed words are kept (1.5 KB -> 6 B per activation); prove() never needs the tables.
  static bool materialize_tables;
  zkReLU() {}
  zkReLU(const zkReLU&) {}                       // aux tensors are per-instance state (std::vector<zkReLU> relus(n), demo.cu:103)
  FrTensor operator()(const FrTensor& X);
  void prove(const FrTensor& X, const FrTensor& Z);
  ~zkReLU();
  const std::vector<Fr_t>& last_proof() { fetch_proof(); return proof_; }
};

// ------------------------------------------------------------------------------------------------ Timer (timer.hpp)
class Timer {
  std::chrono::high_resolution_clock::time_point startTime;
  std::chrono::duration<double> totalTime{0};
  bool isRunning = false;
 public:
  void start() { if (!isRunning) { startTime = std::chrono::high_resolution_clock::now(); isRunning = true; } }
  void stop() { if (isRunning) { totalTime += std::chrono::high_resolution_clock::now() - startTime; isRunning = false; } }
  void reset() { totalTime = std::chrono::duration<double>(0); isRunning = false; }
  double getTotalTime() const { return totalTime.count(); }
};
