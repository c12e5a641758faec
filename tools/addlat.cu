// addlat.cu — latency of a dependent chain of XYZZ additions for a lone warp (the regime of the bucket-reduction tails),
// with the Fq product outlined (ZK_FQ_OUTLINE_MUL=1, the library default) or inlined (=0).  Build both, run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include "../zkdl_b200/csrc/g1.cuh"
using namespace zk;
__device__ __noinline__ G1XYZZ add_ni(G1XYZZ a, G1XYZZ b) { return xyzz_add(a, b); }
__global__ void k_chain(const G1XYZZ* in, G1XYZZ* out, int n) {
  G1XYZZ acc = in[threadIdx.x], b = in[threadIdx.x + blockDim.x];
  for (int i = 0; i < n; ++i) acc = add_ni(acc, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_mulchain(const Fq* in, Fq* out, int n) {
  Fq a = in[threadIdx.x], b = in[threadIdx.x + blockDim.x];
  for (int i = 0; i < n; ++i) a = mul(a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void k_init(G1XYZZ* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  G1XYZZ r; for (int l = 0; l < 12; ++l) { r.x.v[l] = 0x1234567u * (i + 1) + l; r.y.v[l] = 0x7654321u * (i + 3) + l; r.zz.v[l] = 0x1111111u * (i + 5) + l; r.zzz.v[l] = 0x2222222u * (i + 7) + l; }
  r.x.v[11] &= 0x0fffffff; r.y.v[11] &= 0x0fffffff; r.zz.v[11] &= 0x0fffffff; r.zzz.v[11] &= 0x0fffffff;
  p[i] = r;
}
int main() {
  G1XYZZ *in, *out; cudaMalloc(&in, sizeof(G1XYZZ) * 4096); cudaMalloc(&out, sizeof(G1XYZZ) * 148 * 1024);
  k_init<<<16, 256>>>(in, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int n = 64;
  for (int threads : {32, 128, 256, 512}) for (int blocks : {1, 148}) {
    k_chain<<<blocks, threads>>>(in, out, n); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_chain<<<blocks, threads>>>(in, out, n); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"outline\":%d,\"what\":\"xyzz_add chain\",\"threads\":%d,\"blocks\":%d,\"us_per_add\":%.3f}\n", ZK_FQ_OUTLINE_MUL, threads, blocks, ms * 1e3 / n);
  }
  for (int threads : {32, 128, 512}) {
    k_mulchain<<<1, threads>>>((Fq*)in, (Fq*)out, 256); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_mulchain<<<1, threads>>>((Fq*)in, (Fq*)out, 256); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"outline\":%d,\"what\":\"Fq mul chain\",\"threads\":%d,\"us_per_mul\":%.3f}\n", ZK_FQ_OUTLINE_MUL, threads, ms * 1e3 / 256);
  }
  printf("{\"cuda_status\":%d}\n", (int)cudaGetLastError());
  return 0;
}
