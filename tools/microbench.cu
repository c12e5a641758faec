// microbench.cu — on-box pins for the IMAD roofline: raw IMAD issue rate, Fr / Fq Montgomery
// product throughput (inlined and outlined) and XYZZ mixed-add throughput, at several occupancies.
#include <cstdio>
#include <cuda_runtime.h>
#include "../zkdl_b200/csrc/g1.cuh"
using namespace zk;

__global__ void k_imad(uint32_t* out, int iters) {
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, m = blockIdx.x | 1;
  for (int i = 0; i < iters; ++i) {
    a0 = a0 * m + a1; a1 = a1 * m + a2; a2 = a2 * m + a3; a3 = a3 * m + a4; a4 = a4 * m + a5; a5 = a5 * m + a6; a6 = a6 * m + a7; a7 = a7 * m + a0;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
__global__ void k_imad_wide(uint64_t* out, int iters) {
  uint64_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7; uint32_t m = blockIdx.x | 1;
  for (int i = 0; i < iters; ++i) {
    a0 = (uint64_t)(uint32_t)a0 * m + a1; a1 = (uint64_t)(uint32_t)a1 * m + a2; a2 = (uint64_t)(uint32_t)a2 * m + a3; a3 = (uint64_t)(uint32_t)a3 * m + a4;
    a4 = (uint64_t)(uint32_t)a4 * m + a5; a5 = (uint64_t)(uint32_t)a5 * m + a6; a6 = (uint64_t)(uint32_t)a6 * m + a7; a7 = (uint64_t)(uint32_t)a7 * m + a0;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
template <class F, bool OUTLINE>
__global__ void k_mul(F* out, int iters) {
  F x = F::one(), y = F::r2();
  x.v[0] ^= threadIdx.x; y.v[1] ^= blockIdx.x;
  for (int i = 0; i < iters; ++i) {
    if (OUTLINE) { x = mul_outlined(x, y); y = mul_outlined(y, x); } else { x = mul_impl(x, y); y = mul_impl(y, x); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = add(x, y);
}
template <class F>
__global__ void k_mul_ilp2(F* out, int iters) {   // two independent chains per thread
  F x = F::one(), y = F::r2(), z = F::r2(), w = F::one();
  x.v[0] ^= threadIdx.x; y.v[1] ^= blockIdx.x; z.v[2] ^= threadIdx.x;
  for (int i = 0; i < iters; ++i) { x = mul_outlined(x, y); z = mul_outlined(z, w); y = mul_outlined(y, x); w = mul_outlined(w, z); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = add(add(x, y), add(z, w));
}
__global__ void k_madd(const G1Affine* tab, int ntab, G1XYZZ* out, int iters) {
  G1XYZZ acc = xyzz_inf();
  unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < iters; ++i) { idx = idx * 1664525u + 1013904223u; xyzz_madd(acc, tab[idx % ntab], (idx >> 31) != 0); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_mktab(G1Affine* tab, int n) {   // multiples of the generator, affine via repeated madd + normalise is overkill: use x,y of k*G in XYZZ->affine-free way
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  // not on-curve points are fine for throughput measurement of the common path (no special-case hits expected)
  G1Affine p; p.x = Fq::one(); p.y = Fq::r2(); p.x.v[0] ^= (uint32_t)i * 2654435761u; p.y.v[1] ^= (uint32_t)i * 40503u;
  tab[i] = p;
}
template <class K, class... A>
static float timeit(K k, int blocks, int threads, A... args) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(args...); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<<<blocks, threads>>>(args...); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  void* buf; cudaMalloc(&buf, 512u << 20);
  int sms = 148;
  for (int wps : {4, 8, 16, 32}) {            // warps per SM
    int blocks = sms * wps / 4, threads = 128; int iters = 20000;
    float ms = timeit(k_imad, blocks, threads, (uint32_t*)buf, iters);
    printf("{\"bench\":\"imad32\",\"warps_per_sm\":%d,\"T_imad_per_s\":%.3f}\n", wps, 8.0 * iters * blocks * threads / ms / 1e9);
    ms = timeit(k_imad_wide, blocks, threads, (uint64_t*)buf, iters);
    printf("{\"bench\":\"imad_wide\",\"warps_per_sm\":%d,\"T_imad_per_s\":%.3f}\n", wps, 8.0 * iters * blocks * threads / ms / 1e9);
  }
  for (int wps : {4, 8, 12, 16, 24, 32}) {
    int blocks = sms * wps / 4, threads = 128; int iters = 500;
    float ms = timeit(k_mul<Fq, false>, blocks, threads, (Fq*)buf, iters);
    printf("{\"bench\":\"fq_mul_inline\",\"warps_per_sm\":%d,\"G_mul_per_s\":%.3f,\"T_imad_equiv\":%.3f}\n", wps, 2.0 * iters * blocks * threads / ms / 1e6, 2.0 * iters * blocks * threads / ms / 1e6 * 300 / 1e3);
    ms = timeit(k_mul<Fq, true>, blocks, threads, (Fq*)buf, iters);
    printf("{\"bench\":\"fq_mul_outlined\",\"warps_per_sm\":%d,\"G_mul_per_s\":%.3f,\"T_imad_equiv\":%.3f}\n", wps, 2.0 * iters * blocks * threads / ms / 1e6, 2.0 * iters * blocks * threads / ms / 1e6 * 300 / 1e3);
    ms = timeit(k_mul_ilp2<Fq>, blocks, threads, (Fq*)buf, iters);
    printf("{\"bench\":\"fq_mul_outlined_ilp2\",\"warps_per_sm\":%d,\"G_mul_per_s\":%.3f,\"T_imad_equiv\":%.3f}\n", wps, 4.0 * iters * blocks * threads / ms / 1e6, 4.0 * iters * blocks * threads / ms / 1e6 * 300 / 1e3);
    ms = timeit(k_mul<Fr, false>, blocks, threads, (Fr*)buf, iters);
    printf("{\"bench\":\"fr_mul_inline\",\"warps_per_sm\":%d,\"G_mul_per_s\":%.3f,\"T_imad_equiv\":%.3f}\n", wps, 2.0 * iters * blocks * threads / ms / 1e6, 2.0 * iters * blocks * threads / ms / 1e6 * 136 / 1e3);
  }
  G1Affine* tab = (G1Affine*)buf; int ntab = 1 << 16;
  k_mktab<<<ntab / 128, 128>>>(tab, ntab);
  G1XYZZ* out = (G1XYZZ*)((char*)buf + (64u << 20));
  for (int wps : {4, 8, 12}) {
    int blocks = sms * wps / 4, threads = 128; int iters = 200;
    float ms = timeit(k_madd, blocks, threads, tab, ntab, out, iters);
    printf("{\"bench\":\"xyzz_madd\",\"warps_per_sm\":%d,\"M_madd_per_s\":%.3f,\"us_per_madd_per_thread\":%.3f}\n", wps, 1.0 * iters * blocks * threads / ms / 1e3, ms * 1e3 / iters);
  }
  printf("{\"cuda_status\":%d}\n", (int)cudaGetLastError());
  return 0;
}
