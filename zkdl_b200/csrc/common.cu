// common.cu — process-wide state of libzkdl_b200: last error, launch counter, scratch pool, host helpers.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include "common.cuh"
#include "../../include/zkdl_b200.h"

namespace zk {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

static std::once_flag g_pool_once;
static void pool_init() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return;
  uint64_t thr = UINT64_MAX;                       // keep freed blocks in the pool: no cudaMalloc after warm-up
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
}
int scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
  std::call_once(g_pool_once, pool_init);
  ZK_CUDA(cudaMallocAsync(p, bytes, s));
  return ZK_OK;
}
int scratch_free(void* p, cudaStream_t s) {
  ZK_CUDA(cudaFreeAsync(p, s));
  return ZK_OK;
}
int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0; cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace zk

extern "C" {
const char* zkdl_last_error(void) { return zk::g_err; }
int zkdl_version(void) { return 100; }
uint64_t zkdl_launch_count(void) { return zk::g_launches.load(); }

uint32_t zkdl_ceil_log2(uint32_t num) {            // proof.cu:13-31
  if (num == 0) return 0;
  num--;
  uint32_t r = 0;
  while (num > 0) { num >>= 1; r++; }
  return r;
}

// random_vec (proof.cu:3-11) with an injected seed.  std::mt19937 + uniform_int_distribution<unsigned>(0,UINT_MAX)
// yields the raw tempered 32-bit draws; implemented here without <random> so the stream is pinned by this file.
void zkdl_random_vec_host(uint32_t seed, size_t len, zkdl_fr_t* out) {
  uint32_t mt[624]; int idx = 624;
  mt[0] = seed;
  for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  auto next = [&]() -> uint32_t {
    if (idx >= 624) {
      for (int i = 0; i < 624; ++i) {
        uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  };
  for (size_t i = 0; i < len; ++i) {
    for (int j = 0; j < 8; ++j) out[i].val[j] = next();
    out[i].val[7] %= 1944954707u;
  }
}
}
