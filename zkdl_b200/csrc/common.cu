// common.cu — process-wide state of libzkdl_b200: last error, launch counter, scratch pool, host helpers.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <map>
#include <string>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "../../include/zkdl_b200.h"

namespace zk {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_prof_on{0};

namespace {
struct ProfRec { const char* what; cudaEvent_t e0, e1; double bytes, fr_mul, fq_mul; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
}  // namespace
int prof_pre(const char* what, cudaStream_t st, double bytes, double fr_mul, double fq_mul) {
  ProfRec r{what, nullptr, nullptr, bytes, fr_mul, fq_mul};
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
  cudaEventRecord(r.e0, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}
void prof_post(int idx, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx >= 0 && idx < (int)g_prof.size()) cudaEventRecord(g_prof[idx].e1, st);
}
void prof_set_fq_mul(int idx, double fq_mul) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx >= 0 && idx < (int)g_prof.size()) g_prof[idx].fq_mul = fq_mul;
}

void set_last_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

// ---- stream-keyed stack arenas.  All scratch memory of one C-ABI call is allocated and released in LIFO order (RAII
// `Scratch` locals) and is only ever touched by work enqueued on that call's stream, so a per-stream bump allocator is
// exact: memory popped by one call is reused by the next call on the same stream, which the stream orders after it.
// After warm-up there is no cudaMalloc / cudaFree / cudaMallocAsync on the proving path at all.  (cudaMallocAsync was
// used first; with 4-15 concurrent streams its cross-stream reuse logic added 10-50 us per allocation.)
namespace {
struct Block { char* base; size_t cap, top; };
struct Alloc { size_t block, off; bool live; };
struct Arena { std::vector<Block> blocks; std::vector<Alloc> stack; };
std::mutex g_arena_mu;
std::map<std::pair<int, cudaStream_t>, Arena> g_arenas;
constexpr size_t ALIGN = 256;
}  // namespace

int scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_arena_mu);
  int dev = 0; ZK_CUDA(cudaGetDevice(&dev));
  Arena& a = g_arenas[{dev, s}];
  bytes = (bytes + ALIGN - 1) / ALIGN * ALIGN;
  if (a.blocks.empty() || a.blocks.back().top + bytes > a.blocks.back().cap) {
    size_t cap = a.blocks.empty() ? (size_t)64 << 20 : a.blocks.back().cap * 2;
    if (cap < bytes) cap = bytes;
    char* base = nullptr; ZK_CUDA(cudaMalloc((void**)&base, cap));
    a.blocks.push_back({base, cap, 0});
  }
  Block& b = a.blocks.back();
  *p = b.base + b.top;
  a.stack.push_back({a.blocks.size() - 1, b.top, true});
  b.top += bytes;
  return ZK_OK;
}
int scratch_free(void* p, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_arena_mu);
  int dev = 0; ZK_CUDA(cudaGetDevice(&dev));
  Arena& a = g_arenas[{dev, s}];
  for (size_t i = a.stack.size(); i-- > 0;) {
    Alloc& al = a.stack[i];
    if (al.live && a.blocks[al.block].base + al.off == (char*)p) { al.live = false; break; }
  }
  while (!a.stack.empty() && !a.stack.back().live) {       // pop every released allocation on top of the stack
    a.blocks[a.stack.back().block].top = a.stack.back().off;
    a.stack.pop_back();
  }
  if (a.stack.empty() && a.blocks.size() > 1) {           // end of a warm-up call: consolidate its blocks into one
    size_t total = 0;
    cudaStreamSynchronize(s);                             // the call's kernels may still be using the blocks
    for (auto& b : a.blocks) { total += b.cap; cudaFree(b.base); }
    a.blocks.clear();
    char* base = nullptr; ZK_CUDA(cudaMalloc((void**)&base, total));
    a.blocks.push_back({base, total, 0});
  }
  return ZK_OK;
}
static int reserve_one(cudaStream_t s, size_t bytes) {
  void* p = nullptr;
  int rc = scratch_alloc(&p, bytes, s);          // grows the arena to one block of at least `bytes` ...
  if (rc) return rc;
  return scratch_free(p, s);                     // ... and releases it (consolidating if it had several blocks)
}
SideStream& side_stream(int idx, cudaStream_t main) {
  struct Pool { SideStream s[4]; };
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, Pool> pools;       // node-based: references stay valid
  int dev = 0; cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  return pools[{dev, main}].s[idx & 3];
}
int SideStream::fork(cudaStream_t main) {
  if (!stream) {
    ZK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    ZK_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    ZK_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  ZK_CUDA(cudaEventRecord(ev_fork, main));
  ZK_CUDA(cudaStreamWaitEvent(stream, ev_fork, 0));
  return ZK_OK;
}
int SideStream::join(cudaStream_t main) {
  ZK_CUDA(cudaEventRecord(ev_join, stream));
  ZK_CUDA(cudaStreamWaitEvent(main, ev_join, 0));
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
int zkdl_scratch_reserve(size_t bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = zk::reserve_one(st, bytes);
  for (int i = 0; i < 3 && !rc; ++i) {           // the side streams this host thread forks sub-proofs onto
    zk::SideStream& ss = zk::side_stream(i, st);
    if ((rc = ss.fork(st))) break;
    rc = zk::reserve_one(ss.stream, bytes / 2);
    if (!rc) rc = ss.join(st);
  }
  return rc;
}
int zkdl_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(zk::g_prof_mu);
  for (auto& r : zk::g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  zk::g_prof.clear();
  zk::g_prof_on.store(on ? 1 : 0);
  return ZKDL_OK;
}
size_t zkdl_prof_dump(char* buf, size_t cap) {
  // one line per kernel: name launches total_ms bytes fr_mul fq_mul   (the device must be idle: events are synchronised)
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(zk::g_prof_mu);
  struct Agg { double n = 0, ms = 0, bytes = 0, fr = 0, fq = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : zk::g_prof) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    std::string name(r.what);
    size_t cut = name.find("<<<"); if (cut != std::string::npos) name.resize(cut);
    std::string clean; for (char c : name) if (c != ' ') clean += c;
    Agg& a = agg[clean]; a.n += 1; a.ms += ms; a.bytes += r.bytes; a.fr += r.fr_mul; a.fq += r.fq_mul;
  }
  std::string out;
  for (auto& kv : agg) {
    char line[512];
    snprintf(line, sizeof(line), "%s %.0f %.6f %.0f %.0f %.0f\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes, kv.second.fr, kv.second.fq);
    out += line;
  }
  if (buf && cap) { size_t n = out.size() < cap - 1 ? out.size() : cap - 1; memcpy(buf, out.data(), n); buf[n] = 0; }
  return out.size() + 1;
}
int zkdl_scratch_release_all(void) {
  // Frees every scratch arena that holds no live allocation (the arenas are process-global and otherwise kept for the life
  // of the process).  Synchronises the device first: a released block may still be in use by enqueued work.
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { zk::set_last_error("cudaDeviceSynchronize: %s", cudaGetErrorString(e)); return ZKDL_ERR_CUDA; }
  std::lock_guard<std::mutex> lk(zk::g_arena_mu);
  int dev = 0; cudaGetDevice(&dev);
  for (auto it = zk::g_arenas.begin(); it != zk::g_arenas.end();) {
    if (it->first.first == dev && it->second.stack.empty()) {
      for (auto& b : it->second.blocks) cudaFree(b.base);
      it = zk::g_arenas.erase(it);
    } else ++it;
  }
  return ZKDL_OK;
}
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
