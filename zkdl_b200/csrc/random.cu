// random.cu — FrTensor::random / FrTensor::random_int with the reference's generator: curand XORWOW, one state per element
// initialised with curand_init(seed, element index, 0) (/root/reference/fr-tensor.cu:302-347).  Given the same 64-bit seed the
// tables are bit-identical to the reference's, so a reference run and a drop-in run can share generators
// (demo.cu:81-82 multiplies the G1 generator by FrTensor::random) without injecting them.  The reference draws the seed from
// std::random_device; here it is an argument.  (Its `tid > n` guard writes one element past the end; not mirrored.)
#include <curand_kernel.h>
#include "common.cuh"
#include "../../include/zkdl_b200.h"

namespace zk {

__global__ void __launch_bounds__(256) k_fr_random(Fr* __restrict__ out, size_t n, unsigned long long seed) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (tid >= n) return;
  curandState state;
  curand_init(seed, tid, 0, &state);
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = curand(&state);
  r.v[7] %= 1944954707u;                                    // top limb of the modulus: value < p (fr-tensor.cu:346)
  out[tid] = r;
}
__global__ void __launch_bounds__(256) k_fr_random_int(Fr* __restrict__ out, uint32_t num_bits, size_t n, unsigned long long seed) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (tid >= n) return;
  curandState state;
  curand_init(seed, tid, 0, &state);
  Fr v = Fr::zero(), half = Fr::zero();
  v.v[0] = curand(&state) & ((1u << num_bits) - 1u);
  half.v[0] = 1u << (num_bits - 1);
  out[tid] = sub(v, half);                                  // fr-tensor.cu:311-312
}

}  // namespace zk

using namespace zk;

extern "C" {
int zkdl_fr_random(zkdl_fr_t* out, size_t n, uint64_t seed, void* stream) {
  if (n == 0) return ZK_OK;
  ZK_REQUIRE(out, ZK_ERR_ARG, "null argument");
  ZK_LAUNCH(k_fr_random<<<div_up(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<Fr*>(out), n, seed));
  return ZK_OK;
}
int zkdl_fr_random_int(zkdl_fr_t* out, uint32_t num_bits, size_t n, uint64_t seed, void* stream) {
  if (n == 0) return ZK_OK;
  ZK_REQUIRE(out && num_bits >= 1 && num_bits <= 31, ZK_ERR_ARG, "bad argument");
  ZK_LAUNCH(k_fr_random_int<<<div_up(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<Fr*>(out), num_bits, n, seed));
  return ZK_OK;
}
}
