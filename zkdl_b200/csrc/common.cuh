// common.cuh — error handling, stream-ordered scratch memory and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "field.cuh"

namespace zk {

// ---- status codes of the C ABI (include/zkdl_b200.h)
enum : int { ZK_OK = 0, ZK_ERR_DIM = 1, ZK_ERR_CUDA = 2, ZK_ERR_ARG = 3, ZK_ERR_NCCL = 4 };

void set_last_error(const char* fmt, ...);

#define ZK_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      zk::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));       \
      return zk::ZK_ERR_CUDA;                                                                          \
    }                                                                                                  \
  } while (0)
#define ZK_CHECK_LAUNCH() ZK_CUDA(cudaGetLastError())
#define ZK_REQUIRE(cond, code, msg)                                                                    \
  do {                                                                                                 \
    if (!(cond)) { zk::set_last_error("%s:%d: %s", __FILE__, __LINE__, msg); return (code); }          \
  } while (0)

// Scratch memory from per-stream stack arenas (common.cu): after warm-up the proving path performs no cudaMalloc /
// cudaFree and no host sync for temporaries (the reference does one of each per operator, fr-tensor.cu:92-113).
int scratch_alloc(void** p, size_t bytes, cudaStream_t s);
int scratch_free(void* p, cudaStream_t s);

struct Scratch {                       // RAII helper used inside the C-ABI functions
  void* p = nullptr; cudaStream_t s = 0;
  ~Scratch() { if (p) scratch_free(p, s); }
  int alloc(size_t bytes, cudaStream_t st) { s = st; return scratch_alloc(&p, bytes ? bytes : 16, st); }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Fork/join of independent sub-proofs onto side streams (a pool of 4 per main stream, so that proofs issued on different
// main streams - from one host thread or several - never share a side stream): fork() makes side stream `idx` wait for
// the work enqueued so far on `main`; join() makes `main` wait for everything enqueued on the side stream since.
struct SideStream {
  cudaStream_t stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int fork(cudaStream_t main);
  int join(cudaStream_t main);
};
SideStream& side_stream(int idx, cudaStream_t main);   // idx < 4; one pool per (device, main stream)
// fork() now, join() on every way out of the scope: an early error return must not leave the side stream still reading
// the caller's buffers while the caller frees them
struct ForkScope {
  SideStream& ss; cudaStream_t main; bool open = false;
  ForkScope(SideStream& s, cudaStream_t m) : ss(s), main(m) {}
  int fork() { int rc = ss.fork(main); open = rc == 0; return rc; }
  int join() { if (!open) return 0; open = false; return ss.join(main); }
  ~ForkScope() { if (open) ss.join(main); }
};

// ---- launch accounting and the optional per-kernel profiler (zkdl_prof_enable / zkdl_prof_dump, common.cu): with the
// profiler on, ZK_LAUNCH_P brackets the launch with CUDA events on ITS stream and files the elapsed time under the kernel's
// name together with the launch's algorithmic work (bytes by SURVEY.md §8d's model, Fr / Fq Montgomery products).
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_prof_on;
int prof_pre(const char* what, cudaStream_t st, double bytes, double fr_mul, double fq_mul);
void prof_post(int idx, cudaStream_t st);
void prof_set_fq_mul(int idx, double fq_mul);
#define ZK_LAUNCH(...)            \
  do {                            \
    __VA_ARGS__;                  \
    zk::g_launches.fetch_add(1);  \
    ZK_CHECK_LAUNCH();            \
  } while (0)
#define ZK_LAUNCH_P(st_, bytes_, fr_, fq_, ...)                                                                          \
  do {                                                                                                                   \
    int pi__ = zk::g_prof_on.load(std::memory_order_relaxed) ? zk::prof_pre(#__VA_ARGS__, st_, bytes_, fr_, fq_) : -1;   \
    __VA_ARGS__;                                                                                                         \
    zk::g_launches.fetch_add(1);                                                                                         \
    if (pi__ >= 0) zk::prof_post(pi__, st_);                                                                             \
    ZK_CHECK_LAUNCH();                                                                                                   \
  } while (0)

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
int num_sms();

#if defined(__CUDACC__)
// ---- block-wide modular sum of `cnt` Fr values per thread (cnt <= 3); result valid in thread 0.
template <int CNT>
__device__ __forceinline__ void block_reduce_fr(Fr* vals, Fr* smem /* CNT * 32 */) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int c = 0; c < CNT; ++c) {
      Fr o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = __shfl_down_sync(0xffffffffu, vals[c].v[i], off);
      vals[c] = add(vals[c], o);
    }
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < CNT; ++c) smem[c * 32 + warp] = vals[c];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < CNT; ++c) vals[c] = (lane < nwarps) ? smem[c * 32 + lane] : Fr::zero();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int c = 0; c < CNT; ++c) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = __shfl_down_sync(0xffffffffu, vals[c].v[i], off);
        vals[c] = add(vals[c], o);
      }
    }
  }
}
#endif

}  // namespace zk
