// msm.cu — G1 tensors and the fixed-base batched Pippenger MSM behind Commitment::commit / open / me_open.
//
// Replaces /root/reference/g1-tensor.cu (elementwise group ops, G1Jacobian_sum_reduction, the 256-step
// G1Jacobian_mul ladder, G1_me_step) and /root/reference/commitment.cu (commit via |t| independent ladders +
// sum_axis_n_optimized; me_open_step with 5 sequential ladders per thread).
//
// Design.  A Commitment is a fixed generator set, so zkdl_g1_table_create precomputes affine window
// tables 2^(4w) G[i] once; every commitment, opening round and commitment-vector evaluation is then a batched
// Pippenger MSM over those tables with NO doublings on the proving path:
//   count  : signed-digit recode every scalar (digit width c = 4t; c = 8 with tables: 128 buckets per row), histogram (row, bucket) keys
//   scan   : exclusive prefix sums (bucket offsets, partial slots) + chunk plan; ONE single-CTA launch on the few-keys path
//   scatter: counting sort of (window-table index, sign) entries by key
//   accum  : the sorted list cut into equal chunks, one per thread: XYZZ += affine mixed additions (8M+2S), one partial per
//            (chunk, bucket); combine: G lanes per bucket add the partials up
//   reduce : sum_k k*B_k per (row, window-group).  Few rows: k_msm_reduce_coop, four warps share every point addition
//            (one Fq product each per dependency level); many rows: k_msm_reduce_scan (shuffle suffix scan + tree);
//            wide windows (plain Pippenger): k_msm_reduce (running sums + shared-memory scan)
//   final  : Horner over window groups (plain mode only; cooperative doublings for few rows) and XYZZ -> Jacobian
// me_open's log|G| dependent folding rounds are re-expressed as 3*log|G|+1 independent MSMs over the ORIGINAL
// generators (k_open_scalars computes the per-round scalars), so a whole opening is ONE batched MSM.
#include <atomic>
#include <stdlib.h>
#include "common.cuh"
#include "g1_device.cuh"
#include "../../include/zkdl_b200.h"

namespace zk {

int build_eq_table(const Fr* q_dev, const zkdl_fr_t* q_host, int t, int rev, Fr* E, cudaStream_t st);
int fr_partial_me_dev(const Fr* a, size_t n, const zkdl_fr_t* u_host, size_t k, size_t w, Fr* out, cudaStream_t st);

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static constexpr int G1_THREADS = 128;
static constexpr int TABLE_C = 4;                 // table window granularity (bits)
static constexpr int TABLE_W = 64;                // 64 * 4 = 256 bits

static inline unsigned g1_grid(size_t items, int threads) {
  size_t blocks = (items + threads - 1) / threads;
  size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks ? blocks : 1);
}

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void __launch_bounds__(G1_THREADS) k_g1_elementwise(int op, const G1Jac* __restrict__ a, const void* __restrict__ b, size_t nb,
                                                               G1Jac* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    G1Jac pa = a[i];
    if (op == ZKDL_G1_NEG) { pa.y = neg(pa.y); out[i] = pa; continue; }       // G1Jacobian_minus (g1-tensor.cu:17-19)
    G1XYZZ acc = xyzz_from_jac(pa);
    size_t bi = nb == 1 ? 0 : i;
    if (op == ZKDL_G1_MADD || op == ZKDL_G1_MSUB) {
      G1Affine pb = reinterpret_cast<const G1Affine*>(b)[bi];
      xyzz_madd(acc, pb, op == ZKDL_G1_MSUB);
    } else {
      G1XYZZ o = xyzz_from_jac(reinterpret_cast<const G1Jac*>(b)[bi]);
      if (op == ZKDL_G1_SUB) o.y = neg(o.y);
      acc = xyzz_add(acc, o);
    }
    out[i] = xyzz_to_jac(acc);
  }
}
__global__ void __launch_bounds__(G1_THREADS) k_g1_affine_to_jac(const G1Affine* __restrict__ a, G1Jac* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    G1Jac r; r.x = a[i].x; r.y = a[i].y; r.z = Fq::one(); out[i] = r;         // g1-tensor.cu:142-147
  }
}
// out[i] = [x[i]] P[i mod np]; LSB-first double-and-add over the raw limbs (g1-tensor.cu:422-445), XYZZ arithmetic
__global__ void __launch_bounds__(G1_THREADS) k_g1_mul(const G1Jac* __restrict__ P, size_t np, const Fr* __restrict__ x, size_t n, G1Jac* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    G1XYZZ a = xyzz_from_jac(P[i % np]);
    Fr s = x[i];
    int top = -1;
    for (int l = 7; l >= 0; --l) if (s.v[l]) { top = l * 32 + 31 - __clz(s.v[l]); break; }
    G1XYZZ acc = xyzz_inf();
    for (int bit = 0; bit <= top; ++bit) {
      if ((s.v[bit >> 5] >> (bit & 31)) & 1u) acc = xyzz_add(acc, a);
      if (bit < top) a = xyzz_dbl(a);
    }
    out[i] = xyzz_to_jac(acc);
  }
}

// block-wide XYZZ tree sum through shared memory; result in thread 0
__device__ __forceinline__ G1XYZZ block_sum_xyzz(G1XYZZ v, G1XYZZ* sm) {
  sm[threadIdx.x] = v;
  __syncthreads();
  for (unsigned s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] = xyzz_add(sm[threadIdx.x], sm[threadIdx.x + s]);
    __syncthreads();
  }
  return sm[0];
}
extern __shared__ __align__(16) unsigned char g1_dyn_smem[];

__global__ void __launch_bounds__(G1_THREADS) k_g1_sum_partial(const G1Jac* __restrict__ a, size_t n, G1XYZZ* __restrict__ parts) {
  G1XYZZ* sm = reinterpret_cast<G1XYZZ*>(g1_dyn_smem);
  G1XYZZ acc = xyzz_inf();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc = xyzz_add(acc, xyzz_from_jac(a[i]));
  G1XYZZ r = block_sum_xyzz(acc, sm);
  if (threadIdx.x == 0) parts[blockIdx.x] = r;
}
__global__ void __launch_bounds__(G1_THREADS) k_g1_sum_final(const G1XYZZ* __restrict__ parts, unsigned nparts, G1Jac* __restrict__ out) {
  G1XYZZ* sm = reinterpret_cast<G1XYZZ*>(g1_dyn_smem);
  G1XYZZ acc = xyzz_inf();
  for (unsigned i = threadIdx.x; i < nparts; i += blockDim.x) acc = xyzz_add(acc, parts[i]);
  G1XYZZ r = block_sum_xyzz(acc, sm);
  if (threadIdx.x == 0) out[0] = xyzz_to_jac(r);
}

// ------------------------------------------------------------------------------------------------ tables
// tmp[w * n + i] = 2^(TABLE_C * w) * P[i] in XYZZ
__global__ void __launch_bounds__(G1_THREADS) k_table_expand(const G1Jac* __restrict__ pts, size_t n, int windows, G1XYZZ* __restrict__ tmp) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1XYZZ p = xyzz_from_jac(pts[i]);
  for (int w = 0; w < windows; ++w) {
    tmp[(size_t)w * n + i] = p;
    if (w + 1 < windows)
      for (int d = 0; d < TABLE_C; ++d) p = xyzz_dbl(p);
  }
}
// XYZZ -> affine with one inversion per CH points (Montgomery's trick); infinity -> (0,0)
static constexpr int INV_CH = 8;
__global__ void __launch_bounds__(G1_THREADS) k_batch_affine(const G1XYZZ* __restrict__ in, G1Affine* __restrict__ out, size_t total) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t base = t * INV_CH;
  if (base >= total) return;
  Fq prod[INV_CH];
  Fq run = Fq::one();
#pragma unroll
  for (int s = 0; s < INV_CH; ++s) {
    if (base + s < total) { Fq z = in[base + s].zzz; if (!z.is_zero()) run = mul(run, z); }
    prod[s] = run;
  }
  Fq inv = fq_inv(run);
#pragma unroll
  for (int s = INV_CH - 1; s >= 0; --s) {
    if (base + s >= total) continue;
    G1XYZZ p = in[base + s];
    if (p.zzz.is_zero()) { G1Affine z; z.x = Fq::zero(); z.y = Fq::zero(); out[base + s] = z; continue; }
    Fq before = s ? prod[s - 1] : Fq::one();
    Fq izzz = mul(inv, before);
    inv = mul(inv, p.zzz);
    out[base + s] = xyzz_to_affine_with_inv(p, izzz);
  }
}
__global__ void __launch_bounds__(G1_THREADS) k_g1_normalize(const G1Jac* __restrict__ in, G1Jac* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    G1Jac p = in[i];
    if (is_inf(p)) { out[i] = jac_inf(); continue; }
    Fq zi = fq_inv(p.z), zi2 = sqr(zi);
    G1Jac r; r.x = mul(p.x, zi2); r.y = mul(p.y, mul(zi2, zi)); r.z = Fq::one(); out[i] = r;
  }
}

// ------------------------------------------------------------------------------------------------ MSM pipeline
struct MsmCfg {
  size_t n, m;          // bases per row, rows
  int c, W, K;          // digit width, windows, buckets per group = 2^(c-1)
  int NG;               // window groups per row: 1 with full tables, W otherwise
  int full, tstep;      // full tables; table windows per digit window (c / TABLE_C)
  int mont;             // scalars are Montgomery
  int nfull, ctop;      // windows [0, nfull) are c bits wide, the remaining top ones ctop bits (full tables only: every
                        // window feeds the same bucket set, so narrow top windows keep the 255-bit scalar's last few
                        // bits from piling 1/W of all entries into three buckets)
};

template <bool SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(const Fr* __restrict__ scalars, MsmCfg cfg, uint32_t* __restrict__ counters, uint32_t* __restrict__ entries) {
  const size_t total = cfg.m * cfg.n;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t row = idx / cfg.n, i = idx - row * cfg.n;
    Fr mag; bool negative;
    scalar_prepare(scalars[idx], cfg.mont != 0, mag, negative);
    if (mag.is_zero()) continue;
    uint32_t carry = 0;
    for (int w = 0; w < cfg.W; ++w) {
      const int bit = w < cfg.nfull ? w * cfg.c : cfg.nfull * cfg.c + (w - cfg.nfull) * cfg.ctop;
      int32_t d = next_digit_at(mag, bit, w < cfg.nfull ? cfg.c : cfg.ctop, carry);
      if (d == 0) continue;
      uint32_t ad = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      size_t key = (row * cfg.NG + (cfg.full ? 0 : w)) * (size_t)cfg.K + (ad - 1);
      if (!SCATTER) {
        atomicAdd(&counters[key], 1u);
      } else {
        uint32_t pos = atomicAdd(&counters[key], 1u);
        uint32_t base = cfg.full ? (uint32_t)((size_t)(bit / TABLE_C) * cfg.n + i) : (uint32_t)i;
        entries[pos] = base | ((negative != (d < 0)) ? 0x80000000u : 0u);
      }
    }
  }
}

// largest bit length of the normalised scalars |s| (plain-mode window planning)
__global__ void __launch_bounds__(256) k_msm_max_bits(const Fr* __restrict__ scalars, size_t total, int mont, uint32_t* __restrict__ out) {
  uint32_t best = 0;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    Fr mag; bool negative;
    scalar_prepare(scalars[idx], mont != 0, mag, negative);
    for (int l = 7; l >= 0; --l) if (mag.v[l]) { uint32_t b = l * 32 + 32 - __clz(mag.v[l]); if (b > best) best = b; break; }
  }
  for (int o = 16; o > 0; o >>= 1) { uint32_t y = __shfl_down_sync(0xffffffffu, best, o); if (y > best) best = y; }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(out, best);
}

// exclusive scan, three phases, tiles of 1024 * 4
static constexpr int SCAN_T = 1024, SCAN_PER = 4, SCAN_TILE = SCAN_T * SCAN_PER;
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[32];
  unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += y; }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= (unsigned)o) s += y; }
    wsum[lane] = s;
  }
  __syncthreads();
  uint32_t warp_off = warp ? wsum[warp - 1] : 0;
  if (total) *total = wsum[31];
  __syncthreads();
  return warp_off + x - v;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, uint32_t* __restrict__ tile_sums) {
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_PER;
  uint32_t v[SCAN_PER], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER; ++k) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
  uint32_t total;
  uint32_t off = block_exclusive_scan(s, &total);
#pragma unroll
  for (int k = 0; k < SCAN_PER; ++k) { if (base + k < n) out[base + k] = off; off += v[k]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_sums(uint32_t* __restrict__ tile_sums, unsigned ntiles, uint32_t* __restrict__ grand_total) {
  // single block; ntiles may exceed blockDim: serial over chunks
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (unsigned b = 0; b < ntiles; b += SCAN_T) {
    unsigned i = b + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0, total;
    uint32_t off = block_exclusive_scan(v, &total);
    uint32_t c = carry_s;
    if (i < ntiles) tile_sums[i] = off + c;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && grand_total) *grand_total = carry_s;
}
// offsets[i] += tile_sums[tile(i)]; also seeds the scatter cursors and writes offsets[n] = grand total
__global__ void __launch_bounds__(256) k_scan_apply(uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursors, size_t n, const uint32_t* __restrict__ tile_sums,
                                                    const uint32_t* __restrict__ grand_total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i <= n; i += (size_t)gridDim.x * blockDim.x) {
    if (i == n) { offsets[n] = *grand_total; continue; }
    uint32_t o = offsets[i] + tile_sums[i / SCAN_TILE];
    offsets[i] = o; cursors[i] = o;
  }
}

// Few-keys path (openings, commitment-vector evaluations: at most SMALL_KEYS (row, bucket) keys): both exclusive scans
// (bucket offsets, partial slots), the chunk plan and the per-key chunk counts in ONE single-CTA launch instead of eight.
static constexpr uint32_t SMALL_KEYS = 16384;
__global__ void __launch_bounds__(1024) k_msm_scan_small(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursors,
                                                         uint32_t nkeys, uint32_t target, uint32_t* __restrict__ plan, uint32_t* __restrict__ pslot) {
  const uint32_t per = (nkeys + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = threadIdx.x * per, hi = lo + per < nkeys ? lo + per : nkeys;
  uint32_t s = 0, total;
  for (uint32_t k = lo; k < hi; ++k) s += counts[k];
  uint32_t off = block_exclusive_scan(s, &total);
  for (uint32_t k = lo; k < hi; ++k) { offsets[k] = off; cursors[k] = off; off += counts[k]; }
  uint32_t CH = (total + target - 1) / target; if (CH < 4) CH = 4;
  if (threadIdx.x == 0) { offsets[nkeys] = total; plan[0] = CH; plan[1] = (total + CH - 1) / CH; }
  __syncthreads();                                          // offsets[] of the neighbouring thread ranges
  s = 0;
  for (uint32_t k = lo; k < hi; ++k) { uint32_t b = offsets[k], e = offsets[k + 1]; s += b == e ? 0u : ((e - 1) / CH - b / CH + 1u); }
  uint32_t ptotal;
  off = block_exclusive_scan(s, &ptotal);
  for (uint32_t k = lo; k < hi; ++k) {
    uint32_t b = offsets[k], e = offsets[k + 1];
    pslot[k] = off; off += b == e ? 0u : ((e - 1) / CH - b / CH + 1u);
  }
  if (threadIdx.x == 0) pslot[nkeys] = ptotal;
}

// ---- skew-robust bucket accumulation.  The sorted entry list is cut into chunks of CH consecutive entries, one per
// thread, regardless of bucket boundaries, so every thread performs exactly CH mixed additions whatever the digit
// distribution is (a 254-bit scalar leaves the top window with 2 bits of entropy: three buckets receive a third of
// all points each).  A chunk that spans several buckets emits one partial per bucket; partial slot of (key, chunk t)
// = P[key] + t - floor(offset[key] / CH), P = exclusive scan of the per-key chunk counts.
// plan[0] = CH (entries per chunk), plan[1] = number of chunks; decided on the device from the real entry count so that
// the host never synchronises: CH = max(4, ceil(total / target)) => at most target + 1 chunks.
__global__ void k_msm_plan(const uint32_t* __restrict__ offsets, size_t nkeys, uint32_t target, uint32_t* __restrict__ plan) {
  uint32_t total = offsets[nkeys];
  uint32_t CH = (total + target - 1) / target; if (CH < 4) CH = 4;
  plan[0] = CH; plan[1] = (total + CH - 1) / CH;
}
__global__ void __launch_bounds__(256) k_msm_chunk_counts(const uint32_t* __restrict__ offsets, size_t nkeys, const uint32_t* __restrict__ plan, uint32_t* __restrict__ counts) {
  const uint32_t CH = plan[0];
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nkeys; k += (size_t)gridDim.x * blockDim.x) {
    uint32_t b = offsets[k], e = offsets[k + 1];
    counts[k] = b == e ? 0u : ((e - 1) / CH - b / CH + 1u);
  }
}
__device__ __forceinline__ G1XYZZ shfl_xyzz(const G1XYZZ& p, int src_lane_delta_down) {
  G1XYZZ r;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    r.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], src_lane_delta_down);
    r.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], src_lane_delta_down);
    r.zz.v[i] = __shfl_down_sync(0xffffffffu, p.zz.v[i], src_lane_delta_down);
    r.zzz.v[i] = __shfl_down_sync(0xffffffffu, p.zzz.v[i], src_lane_delta_down);
  }
  return r;
}
__global__ void __launch_bounds__(G1_THREADS, 3) k_msm_accumulate(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets,
                                                                  const uint32_t* __restrict__ pslot, const G1Affine* __restrict__ table,
                                                                  G1XYZZ* __restrict__ partials, size_t nkeys, const uint32_t* __restrict__ plan) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t CH = plan[0];
  if (t >= plan[1]) return;
  const uint32_t total = offsets[nkeys];
  uint32_t a = t * CH, b = a + CH < total ? a + CH : total;
  // key of entry a: largest k with offsets[k] <= a
  size_t lo = 0, hi = nkeys;
  while (hi - lo > 1) { size_t mid = (lo + hi) >> 1; if (offsets[mid] <= a) lo = mid; else hi = mid; }
  size_t key = lo;
  uint32_t key_end = offsets[key + 1];
  G1XYZZ acc = xyzz_inf();
  for (uint32_t e = a; e < b; ++e) {
    if (e >= key_end) {                                   // bucket boundary inside the chunk
      partials[pslot[key] + (t - offsets[key] / CH)] = acc;
      acc = xyzz_inf();
      do { ++key; key_end = offsets[key + 1]; } while (e >= key_end);
    }
    uint32_t ent = entries[e];
    G1Affine base = table[ent & 0x7fffffffu];
    xyzz_madd(acc, base, (ent >> 31) != 0);
  }
  partials[pslot[key] + (t - offsets[key] / CH)] = acc;
}
// bucket[key] = sum of its partials, G lanes per key (strided partial sums + a shuffle tree inside the G-lane segment):
// G = 1 when there are many keys, G = 8 on the few-keys path, where every bucket holds dozens of partials.  Buckets with
// more than COMBINE_LIGHT * G partials (the skewed ones) are appended to a list and summed by one warp each in
// k_msm_combine_heavy (strided loads + shuffle tree).
static constexpr uint32_t COMBINE_LIGHT = 6;
template <int G>
__global__ void __launch_bounds__(G1_THREADS) k_msm_combine(const G1XYZZ* __restrict__ partials, const uint32_t* __restrict__ pslot, size_t nkeys,
                                                            G1XYZZ* __restrict__ buckets, uint32_t* __restrict__ heavy_list, uint32_t* __restrict__ heavy_count) {
  const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t key = gid / G; const uint32_t g = (uint32_t)(gid % G);
  const bool valid = key < nkeys;
  if (G == 1 && !valid) return;
  uint32_t base = 0, np = 0;
  if (valid) { base = pslot[key]; np = pslot[key + 1] - base; }
  const bool heavy = np > COMBINE_LIGHT * G;
  G1XYZZ acc = xyzz_inf();
  if (!heavy) for (uint32_t i = g; i < np; i += G) acc = xyzz_add(acc, partials[base + i]);
  if (G > 1) {
#pragma unroll 1
    for (int off = G / 2; off > 0; off >>= 1) {
      G1XYZZ o = shfl_xyzz(acc, off);
      if (g < (uint32_t)off) acc = xyzz_add(acc, o);
    }
  }
  if (valid && g == 0) {
    if (heavy) heavy_list[atomicAdd(heavy_count, 1u)] = (uint32_t)key; else buckets[key] = acc;
  }
}
__global__ void __launch_bounds__(G1_THREADS) k_msm_combine_heavy(const G1XYZZ* __restrict__ partials, const uint32_t* __restrict__ pslot,
                                                                  G1XYZZ* __restrict__ buckets, const uint32_t* __restrict__ heavy_list,
                                                                  const uint32_t* __restrict__ heavy_count) {
  size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= *heavy_count) return;                       // whole warp exits together
  uint32_t key = heavy_list[warp];
  uint32_t base = pslot[key], cnt = pslot[key + 1] - base;
  G1XYZZ tmp = xyzz_inf();
  for (uint32_t i = lane; i < cnt; i += 32) tmp = xyzz_add(tmp, partials[base + i]);
  for (int off = 16; off > 0; off >>= 1) {
    G1XYZZ o = shfl_xyzz(tmp, off);
    if (lane < off) tmp = xyzz_add(tmp, o);
  }
  if (lane == 0) buckets[key] = tmp;
}

// `split` CTAs per window group: CTA j reduces buckets [j*K/split, (j+1)*K/split) to sum_b (b + 1) * B_b (global bucket
// numbers); k_msm_final adds the pieces.  Thread t owns L consecutive buckets: run_t = their sum, tot_t = their sum
// with local weights 1..L (running sums, 2L additions).  The thread offsets then need sum_t t * run_t, which is the
// sum over s >= 1 of the suffix sums E_s = sum_{t >= s} run_t: one Hillis-Steele suffix scan (log T additions per
// thread) replaces a 12-bit double-and-add per thread, which had been 2/3 of this kernel's field multiplications.
__global__ void k_msm_reduce(const G1XYZZ* __restrict__ buckets, int K, int split, G1XYZZ* __restrict__ group_out) {
  G1XYZZ* sm = reinterpret_cast<G1XYZZ*>(g1_dyn_smem);
  const int group = blockIdx.x / split, piece = blockIdx.x % split;
  const int Kp = K / split;                                   // buckets in this piece
  const G1XYZZ* B = buckets + (size_t)group * K;
  const int T = blockDim.x, t = threadIdx.x;
  const int L = Kp >= T ? Kp / T : 1;                         // K, split and T are powers of two
  const int Tp = Kp / L;                                      // threads that own buckets (<= T)
  G1XYZZ run = xyzz_inf(), tot = xyzz_inf();
  if (t < Tp) {
    const int lo = piece * Kp + t * L;
    for (int b = lo + L - 1; b >= lo; --b) { run = xyzz_add(run, B[b]); tot = xyzz_add(tot, run); }
  }
  G1XYZZ E = run;                                             // -> inclusive suffix sum of run over the threads
  sm[t] = E;
  __syncthreads();
  for (int d = 1; d < Tp; d <<= 1) {
    const bool act = t + d < Tp;
    G1XYZZ o;
    if (act) o = sm[t + d];
    __syncthreads();
    if (act) { E = xyzz_add(E, o); sm[t] = E; }
    __syncthreads();
  }
  G1XYZZ val = tot;
  if (t > 0 && t < Tp) {
    for (int l = L; l > 1; l >>= 1) E = xyzz_dbl(E);          // L * E_t
    val = xyzz_add(tot, E);
  } else if (t == 0 && piece) {
    val = xyzz_add(tot, xyzz_mul_small(E, (uint32_t)(piece * Kp)));   // this piece's offset times its total
  }
  G1XYZZ r = block_sum_xyzz(val, sm);
  if (t == 0) group_out[blockIdx.x] = r;
}
// Window groups with few buckets (K <= 1024; the fixed-base c = 8 path has K = 128): one CTA of T <= 128 threads per group,
// L = K / T consecutive buckets per thread, everything else in registers: running sums (2L additions), inclusive suffix
// scan of the thread sums by warp shuffles (5 steps) + one cross-warp step, then a shuffle tree.  With L = 1 the result
// is simply the sum of all suffix sums: 16 dependent additions for 128 buckets (the previous shared-memory version
// needed 36 for 2048 buckets at 8 per thread, on a CTA that saturated one SM).  T = 32 for many rows (throughput:
// 21 warp-additions per row), T = 128 for few (latency).  FINAL (one group per row): writes the Jacobian result itself.
template <bool FINAL>
__global__ void __launch_bounds__(128) k_msm_reduce_scan(const G1XYZZ* __restrict__ buckets, int K, G1XYZZ* __restrict__ group_out, G1Jac* __restrict__ out) {
  __shared__ G1XYZZ wt[4];
  const int T = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = T >> 5;
  const int L = K >= T ? K / T : 1;                           // K and T are powers of two
  const int Tp = K / L;                                       // threads that own buckets (<= T)
  const G1XYZZ* B = buckets + (size_t)blockIdx.x * K;
  G1XYZZ E = xyzz_inf(), tot = xyzz_inf();                    // E: sum of the thread's buckets -> inclusive suffix sum over threads
  if (t < Tp) {
    if (L == 1) E = B[t];
    else for (int b = t * L + L - 1; b >= t * L; --b) { E = xyzz_add(E, B[b]); tot = xyzz_add(tot, E); }
  }
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    G1XYZZ o = shfl_xyzz(E, d);
    if (lane + d < 32) E = xyzz_add(E, o);
  }
  if (nw > 1) {
    if (lane == 0) wt[warp] = E;
    __syncthreads();
    G1XYZZ S = xyzz_inf();
#pragma unroll 1
    for (int w = nw - 1; w > warp; --w) S = xyzz_add(S, wt[w]);
    E = xyzz_add(E, S);
    __syncthreads();
  }
  // sum_b (b + 1) B_b = sum_t tot_t + L * sum_{t >= 1} E_t;  for L = 1 (tot_t = B_t) that is sum_{t >= 0} E_t
  G1XYZZ val = E;
  if (L > 1) {
    val = tot;
    if (t > 0 && t < Tp) {
      for (int l = L; l > 1; l >>= 1) E = xyzz_dbl(E);
      val = xyzz_add(tot, E);
    }
  }
#pragma unroll 1
  for (int off = 16; off > 0; off >>= 1) {
    G1XYZZ o = shfl_xyzz(val, off);
    if (lane < off) val = xyzz_add(val, o);
  }
  if (nw > 1) {
    if (lane == 0) wt[warp] = val;
    __syncthreads();
    if (t == 0)
      for (int w = 1; w < nw; ++w) val = xyzz_add(val, wt[w]);
  }
  if (t == 0) { if (FINAL) out[blockIdx.x] = xyzz_to_jac(val); else group_out[blockIdx.x] = val; }
}
// ---- cooperative bucket reduction for the few-rows case (one opening alone on a GPU).  A dependent XYZZ addition costs a lone
// warp ~15 us (14 Fq products whose carry chains serialise; tools/addlat.cu), and k_msm_reduce_scan is 16 of them in a row.
// Here the FOUR warps of a CTA share every addition of 32 point slots: the products of one dependency level run on four
// different warp schedulers at once (u1 | u2 | s1 | s2, then zz1 zz2 | zzz1 zzz2 | p^2 | r^2, then p^3 | u1 p^2 | zz p^2,
// then s1 p^3 | r (q - x3) | zzz p^3): 4 product latencies per addition instead of 14.  Operands live in shared memory as
// structure-of-arrays (field, limb, slot): conflict-free, and a slot may read another slot's point (the scan).  All reads of a
// call precede all its writes, so `out` may alias an operand.  Infinity operands copy the other one; equal abscissae (P + P,
// P - P: structured inputs only) fall back to the complete single-thread formula for that slot.
struct CoopArr { uint32_t w[4][12][32]; };                       // 32 XYZZ points: x, y, zz, zzz
struct CoopTmp { uint32_t t[11][12][32]; uint32_t slow[32]; };
enum { CT_U1, CT_U2, CT_S1, CT_S2, CT_ZZ12, CT_ZZZ12, CT_P, CT_PP, CT_R, CT_PPP_Q_BASE };   // RR, PPP, Q, Y3B share the tail slots
__device__ __forceinline__ Fq cld(const uint32_t (*f)[32], int slot) { Fq r;
#pragma unroll
  for (int i = 0; i < 12; ++i) r.v[i] = f[i][slot];
  return r; }
__device__ __forceinline__ void cst(uint32_t (*f)[32], int slot, const Fq& v) {
#pragma unroll
  for (int i = 0; i < 12; ++i) f[i][slot] = v.v[i]; }
__device__ __forceinline__ G1XYZZ cld_pt(const CoopArr& A, int slot) { G1XYZZ p; p.x = cld(A.w[0], slot); p.y = cld(A.w[1], slot); p.zz = cld(A.w[2], slot); p.zzz = cld(A.w[3], slot); return p; }
__device__ __forceinline__ void cst_pt(CoopArr& A, int slot, const G1XYZZ& p) { cst(A.w[0], slot, p.x); cst(A.w[1], slot, p.y); cst(A.w[2], slot, p.zz); cst(A.w[3], slot, p.zzz); }

// out[lane] = A[ia] + B[ib] for the lanes with `active`; every thread of the 128-thread CTA must call it
__device__ __noinline__ void coop_add(CoopArr& out, const CoopArr& A, int ia, const CoopArr& B, int ib, bool active, CoopTmp& T, int warp, int lane) {
  uint32_t (*U1)[32] = T.t[0], (*U2)[32] = T.t[1], (*S1)[32] = T.t[2], (*S2)[32] = T.t[3], (*ZZ12)[32] = T.t[4], (*ZZZ12)[32] = T.t[5];
  uint32_t (*P)[32] = T.t[6], (*PP)[32] = T.t[7], (*R)[32] = T.t[8], (*PPP)[32] = T.t[9], (*Q)[32] = T.t[10];
  uint32_t (*RR)[32] = T.t[1], (*Y3B)[32] = T.t[3];                 // U2 / S2 are dead once p and r exist
  bool a_inf = false, b_inf = false;
  if (active) {
    uint32_t za = 0, zb = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) { za |= A.w[2][i][ia]; zb |= B.w[2][i][ib]; }
    a_inf = za == 0; b_inf = zb == 0;
  }
  const bool live = active && !a_inf && !b_inf;
  if (live) {
    if (warp == 0) cst(U1, lane, mul(cld(A.w[0], ia), cld(B.w[2], ib)));
    else if (warp == 1) cst(U2, lane, mul(cld(B.w[0], ib), cld(A.w[2], ia)));
    else if (warp == 2) cst(S1, lane, mul(cld(A.w[1], ia), cld(B.w[3], ib)));
    else cst(S2, lane, mul(cld(B.w[1], ib), cld(A.w[3], ia)));
  }
  __syncthreads();
  Fq rr_keep = Fq::zero();
  if (live) {
    if (warp == 0) cst(ZZ12, lane, mul(cld(A.w[2], ia), cld(B.w[2], ib)));
    else if (warp == 1) cst(ZZZ12, lane, mul(cld(A.w[3], ia), cld(B.w[3], ib)));
    else if (warp == 2) { Fq p = sub(cld(U2, lane), cld(U1, lane)); cst(P, lane, p); cst(PP, lane, sqr(p)); T.slow[lane] = p.is_zero() ? 1u : 0u; }
    else { Fq r = sub(cld(S2, lane), cld(S1, lane)); cst(R, lane, r); rr_keep = sqr(r); }
  } else if (warp == 2) T.slow[lane] = 0u;
  __syncthreads();
  if (live && warp == 3) cst(RR, lane, rr_keep);                     // U2 is dead now (p was formed before the barrier)
  const bool slow = live && T.slow[lane] != 0u;
  Fq keep = Fq::zero();
  if (live && !slow) {
    if (warp == 0) cst(PPP, lane, mul(cld(P, lane), cld(PP, lane)));
    else if (warp == 1) cst(Q, lane, mul(cld(U1, lane), cld(PP, lane)));
    else if (warp == 2) keep = mul(cld(ZZ12, lane), cld(PP, lane));   // zz3
  }
  __syncthreads();
  G1XYZZ res;                                                          // slow path only (warp 0)
  if (live && !slow) {
    if (warp == 0 || warp == 1) {
      Fq q = cld(Q, lane), ppp = cld(PPP, lane);
      Fq x3 = sub(sub(sub(cld(RR, lane), ppp), q), q);
      if (warp == 0) { keep = x3; cst(Y3B, lane, mul(cld(S1, lane), ppp)); }      // S2 is dead (r was formed two barriers ago)
      else keep = mul(cld(R, lane), sub(q, x3));                                  // y3a
    } else if (warp == 3) keep = mul(cld(ZZZ12, lane), cld(PPP, lane));           // zzz3
  } else if (active && a_inf) keep = cld(B.w[warp], ib);                          // inf + b = b (all reads before any write)
  else if (active && b_inf) keep = cld(A.w[warp], ia);
  else if (slow && warp == 0) res = xyzz_add(cld_pt(A, ia), cld_pt(B, ib));       // P + P or P - P: the complete formula
  __syncthreads();
  if (active) {
    if (slow) { if (warp == 0) cst_pt(out, lane, res); }
    else {
      if (live && warp == 1) keep = sub(keep, cld(Y3B, lane));
      cst(out.w[warp], lane, keep);
    }
  }
  __syncthreads();
}

// out[lane] = 2 * A[lane] for the lanes with `active` (dbl-2008-s-1, a = 0): three product levels instead of nine products in a
// row: (2y)^2 | x^2, then u v | x v | m^2 | v zz, then m (s - x3) | w y | w zzz.  Same calling convention as coop_add.
__device__ __noinline__ void coop_dbl(CoopArr& out, const CoopArr& A, bool active, CoopTmp& T, int warp, int lane) {
  uint32_t (*V)[32] = T.t[0], (*XX)[32] = T.t[1], (*W)[32] = T.t[2], (*S)[32] = T.t[3], (*MM)[32] = T.t[4], (*WY)[32] = T.t[5];
  bool inf = true;
  if (active) {
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) z |= A.w[2][i][lane];
    inf = z == 0;
  }
  const bool live = active && !inf;
  if (live) {
    if (warp == 0) cst(V, lane, sqr(dbl(cld(A.w[1], lane))));
    else if (warp == 1) cst(XX, lane, sqr(cld(A.w[0], lane)));
  }
  __syncthreads();
  Fq keep = Fq::zero(), m = Fq::zero();
  if (live) {
    if (warp == 0) cst(W, lane, mul(dbl(cld(A.w[1], lane)), cld(V, lane)));
    else if (warp == 1) cst(S, lane, mul(cld(A.w[0], lane), cld(V, lane)));
    else if (warp == 2) { Fq xx = cld(XX, lane); m = add(dbl(xx), xx); cst(MM, lane, sqr(m)); }
    else keep = mul(cld(V, lane), cld(A.w[2], lane));                 // zz3
  }
  __syncthreads();
  if (live) {
    if (warp == 0 || warp == 2) {
      Fq s_ = cld(S, lane);
      Fq x3 = sub(sub(cld(MM, lane), s_), s_);
      if (warp == 0) keep = x3;
      else keep = mul(m, sub(s_, x3));                                // y3a (warp 2 still holds m)
    } else if (warp == 1) cst(WY, lane, mul(cld(W, lane), cld(A.w[1], lane)));
    else { Fq zz3 = keep; (void)zz3; }
  }
  Fq zzz3 = Fq::zero();
  if (live && warp == 3) zzz3 = mul(cld(W, lane), cld(A.w[3], lane));
  __syncthreads();
  if (live) {                                                         // infinity stays infinity (nothing to write when out == A)
    if (warp == 0) cst(out.w[0], lane, keep);
    else if (warp == 2) cst(out.w[1], lane, sub(keep, cld(WY, lane)));
    else if (warp == 3) { cst(out.w[2], lane, keep); cst(out.w[3], lane, zzz3); }
  } else if (active && &out != &A) {
    cst(out.w[warp], lane, cld(A.w[warp], lane));
  }
  __syncthreads();
}

// Horner over the window groups of a plain (no precomputed windows) Pippenger MSM, one CTA of 128 threads per row:
// acc = 2^c acc + G_w from the top window down.  The 255 dependent doublings are the latency floor of a one-off MSM
// (a lone thread: ~9.6 us each = 2.5 ms, 70 % of a 2^16-point MSM); the four warps share each doubling / addition.
__global__ void __launch_bounds__(128) k_msm_final_coop(const G1XYZZ* __restrict__ groups, int NG, int split, int c, G1Jac* __restrict__ out) {
  __shared__ CoopArr ACC, OP;
  __shared__ CoopTmp T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const G1XYZZ* g = groups + (size_t)blockIdx.x * NG * split;
  auto field = [&](const G1XYZZ& p) -> Fq { return warp == 0 ? p.x : (warp == 1 ? p.y : (warp == 2 ? p.zz : p.zzz)); };
  { G1XYZZ inf = xyzz_inf(); cst(ACC.w[warp], lane, field(inf)); }
  __syncthreads();
  const bool me = lane == 0;
#pragma unroll 1
  for (int w = NG - 1; w >= 0; --w) {
    if (w != NG - 1)
#pragma unroll 1
      for (int d = 0; d < c; ++d) coop_dbl(ACC, ACC, me, T, warp, lane);
#pragma unroll 1
    for (int j = 0; j < split; ++j) {
      if (me) cst(OP.w[warp], 0, field(g[(size_t)w * split + j]));
      __syncthreads();
      coop_add(ACC, ACC, lane, OP, lane, me, T, warp, lane);
    }
  }
  if (threadIdx.x == 0) out[blockIdx.x] = xyzz_to_jac(cld_pt(ACC, 0));
}

// One CTA of 128 threads per row of K = 128 buckets: slot t owns buckets 4t..4t+3.
//   A  running sums: run_t = sum of the slot's buckets, tot_t = their sum with weights 1..4        (6 additions)
//   B  E_t = inclusive suffix sum of run over the slots, in place                                    (5)
//   C  sum_b (b+1) B_b = sum_t tot_t + 4 sum_{t>=1} E_t: both sums by one tree (first step apart, then packed 16 + 16)   (6)
//   D  thread 0: T + [4] F, to Jacobian
template <bool FINAL>
__global__ void __launch_bounds__(128) k_msm_reduce_coop(const G1XYZZ* __restrict__ buckets, G1XYZZ* __restrict__ group_out, G1Jac* __restrict__ out) {
  __shared__ CoopArr RUN, TOT, OPB;
  __shared__ CoopTmp T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const G1XYZZ* B = buckets + (size_t)blockIdx.x * 128;
  auto load_field = [&](const G1XYZZ& p) -> Fq { return warp == 0 ? p.x : (warp == 1 ? p.y : (warp == 2 ? p.zz : p.zzz)); };
  {
    Fq v = load_field(B[4 * lane + 3]);
    cst(RUN.w[warp], lane, v); cst(TOT.w[warp], lane, v);
  }
  __syncthreads();
#pragma unroll 1
  for (int b = 2; b >= 0; --b) {
    cst(OPB.w[warp], lane, load_field(B[4 * lane + b]));
    __syncthreads();
    coop_add(RUN, RUN, lane, OPB, lane, true, T, warp, lane);
    coop_add(TOT, TOT, lane, RUN, lane, true, T, warp, lane);
  }
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) coop_add(RUN, RUN, lane, RUN, (lane + d) & 31, lane + d < 32, T, warp, lane);   // RUN -> E
  // tree, first step per sum; slot 0 of E is excluded from F
  if (lane == 0) { Fq z = Fq::zero(); if (warp >= 2) cst(RUN.w[warp], 0, z); }                                   // E_0 := infinity
  __syncthreads();
  coop_add(TOT, TOT, lane, TOT, (lane + 16) & 31, lane < 16, T, warp, lane);
  coop_add(RUN, RUN, lane, RUN, (lane + 16) & 31, lane < 16, T, warp, lane);
  if (lane >= 16) cst(TOT.w[warp], lane, cld(RUN.w[warp], lane - 16));                                           // pack: TOT[16..31] = partial F
  __syncthreads();
#pragma unroll 1
  for (int off = 8; off > 0; off >>= 1) coop_add(TOT, TOT, lane, TOT, (lane + off) & 31, (lane & 15) < off, T, warp, lane);
  if (threadIdx.x == 0) {
    G1XYZZ Tt = cld_pt(TOT, 0), F = cld_pt(TOT, 16);
    G1XYZZ r = xyzz_add(Tt, xyzz_dbl(xyzz_dbl(F)));
    if (FINAL) out[blockIdx.x] = xyzz_to_jac(r); else group_out[blockIdx.x] = r;
  }
}

__global__ void __launch_bounds__(64) k_msm_final(const G1XYZZ* __restrict__ groups, size_t m, int NG, int split, int c, G1Jac* __restrict__ out) {
  size_t row = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (row >= m) return;
  G1XYZZ acc = xyzz_inf();
  for (int w = NG - 1; w >= 0; --w) {
    if (w != NG - 1)
      for (int d = 0; d < c; ++d) acc = xyzz_dbl(acc);
    const G1XYZZ* g = groups + (row * NG + w) * (size_t)split;
    for (int j = 0; j < split; ++j) acc = xyzz_add(acc, g[j]);
  }
  out[row] = xyzz_to_jac(acc);
}

// ------------------------------------------------------------------------------------------------ me_open scalars
// k_open_fold (single CTA): the log n dependent scalar folds and generator-weight doublings of an opening, every level
// kept: sLv[off_j + i] = s^(j)[i] (off_j = 2n - 2n/2^j), wLv[2^j - 1 + b] = w^(j)[b].  ret = the final folded scalar.
// k_open_rows (whole grid): rows[(3j + {0,1,2}) * n + I] = scalars of T, T0, T1 of round j over the ORIGINAL generators,
// rows[3k * n + I] = weights of the final folded generator.  See DESIGN.md §4 "An opening is one batched MSM".
__global__ void __launch_bounds__(1024) k_open_fold(const Fr* __restrict__ s0, const Fr* __restrict__ u, int k, size_t n, Fr* sLv, Fr* wLv, Fr* __restrict__ ret) {
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) sLv[i] = s0[i];
  if (threadIdx.x == 0) wLv[0] = Fr::one();
  __syncthreads();
  size_t soff = 0;
  for (int j = 0; j < k; ++j) {
    size_t B = (size_t)1 << j, nj = n >> j;
    const Fr* s = sLv + soff; Fr* sn = sLv + soff + nj;
    const Fr* w = wLv + (B - 1); Fr* wn = wLv + (2 * B - 1);
    Fr uj = u[j];
    for (size_t g = threadIdx.x; g < nj / 2; g += blockDim.x) sn[g] = add(s[2 * g], mul(uj, sub(s[2 * g + 1], s[2 * g])));   // commitment.cu:55
    for (size_t b = threadIdx.x; b < B; b += blockDim.x) {    // G' = [u'] G0 + [1-u'] G1  (commitment.cu:56)
      Fr wb = w[b], hi = mul(wb, uj);
      wn[b] = hi; wn[B + b] = sub(wb, hi);
    }
    __syncthreads();
    soff += nj;
  }
  if (threadIdx.x == 0) ret[0] = sLv[soff];
}
__global__ void __launch_bounds__(256) k_open_rows(const Fr* __restrict__ sLv, const Fr* __restrict__ wLv, int k, size_t n, Fr* __restrict__ rows) {
  const size_t total = (size_t)(k + 1) * n;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int j = (int)(idx / n); size_t I = idx - (size_t)j * n;
    size_t B = (size_t)1 << j;
    if (j == k) { rows[(size_t)(3 * k) * n + I] = from_mont(wLv[B - 1 + I]); continue; }
    const Fr* s = sLv + (2 * n - ((2 * n) >> j));
    size_t i = I >> j, b = I & (B - 1);
    Fr wb = wLv[B - 1 + b];
    Fr* T = rows + (size_t)(3 * j) * n; Fr* T0 = T + n; Fr* T1 = T0 + n;
    T[I] = mul(s[i], wb);                         // raw Montgomery limbs of s times Montgomery weight = plain integer
    if (i & 1) { T0[I] = mul(s[i - 1], wb); T1[I] = Fr::zero(); }
    else { T1[I] = mul(s[i + 1], wb); T0[I] = Fr::zero(); }
  }
}

}  // namespace zk

using namespace zk;

struct zkdl_g1_table {
  size_t n; int full; int windows;
  G1Affine* pts; size_t bytes;
};

namespace zk {

static int pick_c(const zkdl_g1_table* t, size_t m) {
  if (const char* e = getenv("ZKDL_MSM_C")) { int c = atoi(e); if (c >= 4 && c <= 16) return c; }   // tuning knob
  if (t->full) return 8;      // 128 buckets per row: the bucket reduction is 16 dependent additions (12 -> 2048 buckets was the
                              // longest kernel of an opening); ZKDL_MSM_C=12 restores the wide windows for A/B runs
  int lg = 0; while (((size_t)1 << (lg + 1)) <= t->n) ++lg;
  int c = lg - 5; if (c < 4) c = 4; if (c > 16) c = 16;
  return c;
}

int msm_run(const zkdl_g1_table* t, const Fr* scalars, size_t m, int mont, int c_override, G1Jac* out, cudaStream_t st) {
  if (m == 0) return ZK_OK;
  MsmCfg cfg;
  cfg.n = t->n; cfg.m = m; cfg.full = t->full; cfg.mont = mont;
  cfg.c = c_override ? c_override : pick_c(t, m);
  int maxbits = 254;
  if (!cfg.full && !c_override) {
    // Plain Pippenger (setup-time commitments and the MSM sweep; never on the proving path): one 4-byte read-back of the
    // largest scalar magnitude lets the windows be planned so that the TOP window is as wide as possible.  A narrow top
    // window (e.g. 3 bits of a 16-bit scalar under c = 13) funnels all points into a handful of buckets.
    Scratch mb; int rc0;
    if ((rc0 = mb.alloc(sizeof(uint32_t), st))) return rc0;
    ZK_CUDA(cudaMemsetAsync(mb.p, 0, sizeof(uint32_t), st));
    ZK_LAUNCH(k_msm_max_bits<<<g1_grid(m * cfg.n, 256), 256, 0, st>>>(scalars, m * cfg.n, mont, mb.as<uint32_t>()));
    uint32_t h = 0;
    ZK_CUDA(cudaMemcpyAsync(&h, mb.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    maxbits = h ? (int)h : 1;
    int c0 = cfg.c, best_c = c0, best_top = -1;
    for (int c = (c0 > 6 ? c0 - 2 : 4); c <= (c0 + 2 < 16 ? c0 + 2 : 16); ++c) {
      int W = (maxbits + 1 + c - 1) / c;                     // one spare bit for the signed-digit carry
      int top = maxbits + 1 - (W - 1) * c;                   // bits that reach the top window
      if (W == 1) top = c;                                   // a single window has no skewed top
      if (top > 4) top = 4;                                  // 4+ bits in the top window is balanced enough: then stay close to c0
      if (top > best_top || (top == best_top && abs(c - c0) < abs(best_c - c0))) { best_top = top; best_c = c; }
    }
    cfg.c = best_c;
    if (maxbits + 1 <= 20 && ((size_t)1 << (maxbits + 1)) <= 8 * cfg.n) cfg.c = maxbits + 1 < 4 ? 4 : maxbits + 1;   // one window covers the scalar
  }
  if (cfg.full) { cfg.c = (cfg.c / TABLE_C) * TABLE_C; if (cfg.c < TABLE_C) cfg.c = TABLE_C; if (cfg.c > 16) cfg.c = 16; }
  cfg.tstep = cfg.c / TABLE_C;
  cfg.W = cfg.full ? (255 + cfg.c - 1) / cfg.c : (maxbits + 1 + cfg.c - 1) / cfg.c;
  cfg.nfull = cfg.W; cfg.ctop = cfg.c;
  if (cfg.full && cfg.c == 12) { cfg.nfull = 20; cfg.ctop = 8; cfg.W = 22; }      // 20 x 12 + 2 x 8 = 256 bits: 7 bits reach the top window
  cfg.K = 1 << (cfg.c - 1);
  cfg.NG = cfg.full ? 1 : cfg.W;
  size_t nkeys = m * (size_t)cfg.NG * cfg.K;
  size_t max_entries = m * cfg.n * (size_t)cfg.W;
  ZK_REQUIRE(max_entries < 0xffffffffull && nkeys < 0x7fffffffull, ZK_ERR_ARG, "MSM too large for 32-bit entry indices");
  ZK_REQUIRE(!cfg.full || (size_t)cfg.W * cfg.tstep * cfg.n < 0x80000000ull, ZK_ERR_ARG, "table too large");
  Scratch counts, offsets, cursors, tiles, total, entries, buckets, groups, pcounts, pslot, ptiles, ptotal, partials; int rc;
  unsigned ntiles = div_up(nkeys, SCAN_TILE);
  if ((rc = counts.alloc(sizeof(uint32_t) * nkeys, st))) return rc;
  if ((rc = offsets.alloc(sizeof(uint32_t) * (nkeys + 1), st))) return rc;
  if ((rc = cursors.alloc(sizeof(uint32_t) * nkeys, st))) return rc;
  if ((rc = tiles.alloc(sizeof(uint32_t) * ntiles, st))) return rc;
  if ((rc = total.alloc(sizeof(uint32_t), st))) return rc;
  if ((rc = entries.alloc(sizeof(uint32_t) * max_entries, st))) return rc;
  if ((rc = buckets.alloc(sizeof(G1XYZZ) * nkeys, st))) return rc;

  ZK_CUDA(cudaMemsetAsync(counts.p, 0, sizeof(uint32_t) * nkeys, st));
  unsigned dgrid = g1_grid(m * cfg.n, 256);
  const bool few_keys = nkeys <= SMALL_KEYS;
  // chunked accumulation: about one resident wave of threads (measured: 0.5 / 1 / 2 waves within 1 %; fewer chunks = fewer
  // partials to combine), each with the same number of mixed additions
  static const double waves = getenv("ZKDL_MSM_WAVES") ? atof(getenv("ZKDL_MSM_WAVES")) : 1.0;   // tuning knob
  uint32_t target = (uint32_t)(num_sms() * 384 * waves);
  const int G = nkeys <= 8192 ? 8 : 1;                          // combine lanes per bucket
  Scratch plan, heavy;
  if ((rc = plan.alloc(sizeof(uint32_t) * 2, st))) return rc;
  size_t max_heavy = (nkeys + (size_t)target + 2) / (COMBINE_LIGHT * G + 1) + 1;
  if ((rc = heavy.alloc(sizeof(uint32_t) * (max_heavy + 1), st))) return rc;
  ZK_CUDA(cudaMemsetAsync(heavy.p, 0, sizeof(uint32_t), st));
  if ((rc = pcounts.alloc(sizeof(uint32_t) * nkeys, st))) return rc;
  if ((rc = pslot.alloc(sizeof(uint32_t) * (nkeys + 1), st))) return rc;
  if ((rc = ptiles.alloc(sizeof(uint32_t) * ntiles, st))) return rc;
  if ((rc = ptotal.alloc(sizeof(uint32_t), st))) return rc;
  if ((rc = partials.alloc(sizeof(G1XYZZ) * (nkeys + (size_t)target + 2), st))) return rc;
  ZK_LAUNCH(k_msm_digits<false><<<dgrid, 256, 0, st>>>(scalars, cfg, counts.as<uint32_t>(), nullptr));
  if (few_keys) {
    ZK_LAUNCH(k_msm_scan_small<<<1, 1024, 0, st>>>(counts.as<uint32_t>(), offsets.as<uint32_t>(), cursors.as<uint32_t>(), (uint32_t)nkeys, target,
                                                    plan.as<uint32_t>(), pslot.as<uint32_t>()));
    ZK_LAUNCH(k_msm_digits<true><<<dgrid, 256, 0, st>>>(scalars, cfg, cursors.as<uint32_t>(), entries.as<uint32_t>()));
  } else {
    ZK_LAUNCH(k_scan_tiles<<<ntiles, SCAN_T, 0, st>>>(counts.as<uint32_t>(), offsets.as<uint32_t>(), nkeys, tiles.as<uint32_t>()));
    ZK_LAUNCH(k_scan_sums<<<1, SCAN_T, 0, st>>>(tiles.as<uint32_t>(), ntiles, total.as<uint32_t>()));
    ZK_LAUNCH(k_scan_apply<<<g1_grid(nkeys + 1, 256), 256, 0, st>>>(offsets.as<uint32_t>(), cursors.as<uint32_t>(), nkeys, tiles.as<uint32_t>(), total.as<uint32_t>()));
    ZK_LAUNCH(k_msm_digits<true><<<dgrid, 256, 0, st>>>(scalars, cfg, cursors.as<uint32_t>(), entries.as<uint32_t>()));
    ZK_LAUNCH(k_msm_plan<<<1, 1, 0, st>>>(offsets.as<uint32_t>(), nkeys, target, plan.as<uint32_t>()));
    ZK_LAUNCH(k_msm_chunk_counts<<<g1_grid(nkeys, 256), 256, 0, st>>>(offsets.as<uint32_t>(), nkeys, plan.as<uint32_t>(), pcounts.as<uint32_t>()));
    ZK_LAUNCH(k_scan_tiles<<<ntiles, SCAN_T, 0, st>>>(pcounts.as<uint32_t>(), pslot.as<uint32_t>(), nkeys, ptiles.as<uint32_t>()));
    ZK_LAUNCH(k_scan_sums<<<1, SCAN_T, 0, st>>>(ptiles.as<uint32_t>(), ntiles, ptotal.as<uint32_t>()));
    ZK_LAUNCH(k_scan_apply<<<g1_grid(nkeys + 1, 256), 256, 0, st>>>(pslot.as<uint32_t>(), pcounts.as<uint32_t>(), nkeys, ptiles.as<uint32_t>(), ptotal.as<uint32_t>()));
  }
  if (g_prof_on.load(std::memory_order_relaxed)) {          // profiling pass only: the real entry count (one synchronous 4-byte read)
    uint32_t nent = 0;
    ZK_CUDA(cudaMemcpyAsync(&nent, offsets.as<uint32_t>() + nkeys, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    // one mixed addition (madd-2008-s) = 8M + 2S = 10 Fq products per sorted entry
    ZK_LAUNCH_P(st, 0.0, 0.0, 10.0 * nent, k_msm_accumulate<<<div_up((size_t)target + 1, G1_THREADS), G1_THREADS, 0, st>>>(
        entries.as<uint32_t>(), offsets.as<uint32_t>(), pslot.as<uint32_t>(), t->pts, partials.as<G1XYZZ>(), nkeys, plan.as<uint32_t>()));
  } else
  ZK_LAUNCH(k_msm_accumulate<<<div_up((size_t)target + 1, G1_THREADS), G1_THREADS, 0, st>>>(entries.as<uint32_t>(), offsets.as<uint32_t>(), pslot.as<uint32_t>(),
                                                                                              t->pts, partials.as<G1XYZZ>(), nkeys, plan.as<uint32_t>()));
  uint32_t* hcount = heavy.as<uint32_t>(); uint32_t* hlist = hcount + 1;
  if (G == 8) ZK_LAUNCH_P(st, 0.0, 0.0, 0.0, k_msm_combine<8><<<div_up(nkeys * 8, G1_THREADS), G1_THREADS, 0, st>>>(partials.as<G1XYZZ>(), pslot.as<uint32_t>(), nkeys, buckets.as<G1XYZZ>(), hlist, hcount));
  else ZK_LAUNCH_P(st, 0.0, 0.0, 0.0, k_msm_combine<1><<<div_up(nkeys, G1_THREADS), G1_THREADS, 0, st>>>(partials.as<G1XYZZ>(), pslot.as<uint32_t>(), nkeys, buckets.as<G1XYZZ>(), hlist, hcount));
  ZK_LAUNCH_P(st, 0.0, 0.0, 0.0, k_msm_combine_heavy<<<div_up(max_heavy * 32, G1_THREADS), G1_THREADS, 0, st>>>(partials.as<G1XYZZ>(), pslot.as<uint32_t>(), buckets.as<G1XYZZ>(), hlist, hcount));
  const size_t ngroups = m * (size_t)cfg.NG;
  if (cfg.K <= 1024) {
    // few buckets per group: register/shuffle reduction, one CTA per group; 32 threads when rows are plentiful, 128 otherwise
    int T = ngroups >= (size_t)num_sms() * 2 ? 32 : 128;
    if (T > cfg.K) T = cfg.K < 32 ? 32 : cfg.K;
    static const bool no_coop = getenv("ZKDL_MSM_NO_COOP") != nullptr;      // A/B knob
    if (cfg.K == 128 && T == 128 && cfg.NG == 1 && !no_coop) {               // few rows: latency matters, four warps share every addition
      ZK_LAUNCH_P(st, 0.0, 0.0, 0.0, k_msm_reduce_coop<true><<<(unsigned)ngroups, 128, 0, st>>>(buckets.as<G1XYZZ>(), nullptr, out));
      return ZK_OK;
    }
    if (cfg.NG == 1) {
      ZK_LAUNCH_P(st, 0.0, 0.0, 0.0, k_msm_reduce_scan<true><<<(unsigned)ngroups, T, 0, st>>>(buckets.as<G1XYZZ>(), cfg.K, nullptr, out));
      return ZK_OK;
    }
    if ((rc = groups.alloc(sizeof(G1XYZZ) * ngroups, st))) return rc;
    ZK_LAUNCH_P(st, 0.0, 0.0, 0.0, k_msm_reduce_scan<false><<<(unsigned)ngroups, T, 0, st>>>(buckets.as<G1XYZZ>(), cfg.K, groups.as<G1XYZZ>(), nullptr));
    static const bool no_coop_f = getenv("ZKDL_MSM_NO_COOP") != nullptr;
    if (m <= 64 && !no_coop_f) ZK_LAUNCH(k_msm_final_coop<<<(unsigned)m, 128, 0, st>>>(groups.as<G1XYZZ>(), cfg.NG, 1, cfg.c, out));
    else ZK_LAUNCH(k_msm_final<<<div_up(m, 64), 64, 0, st>>>(groups.as<G1XYZZ>(), m, cfg.NG, 1, cfg.c, out));
    return ZK_OK;
  }
  // many buckets per group (plain Pippenger with wide windows): up to 256-thread CTAs, RL (8) buckets per thread, more only
  // when 16 CTAs per group are not enough
  static const int RL = getenv("ZKDL_MSM_RED_L") ? atoi(getenv("ZKDL_MSM_RED_L")) : 8;      // tuning knobs (powers of two)
  static const int RT = getenv("ZKDL_MSM_RED_T") ? atoi(getenv("ZKDL_MSM_RED_T")) : 256;
  int T = cfg.K / RL; if (T < 32) T = 32; if (T > RT) T = RT; if (T > cfg.K) T = cfg.K < 32 ? 32 : cfg.K;
  int split = cfg.K / (T * RL); if (split < 1) split = 1; if (split > 16) split = 16;
  if ((rc = groups.alloc(sizeof(G1XYZZ) * m * cfg.NG * split, st))) return rc;
  ZK_LAUNCH_P(st, 0.0, 0.0, 0.0, k_msm_reduce<<<(unsigned)(m * cfg.NG * split), T, sizeof(G1XYZZ) * T, st>>>(buckets.as<G1XYZZ>(), cfg.K, split, groups.as<G1XYZZ>()));
  static const bool no_coop_final = getenv("ZKDL_MSM_NO_COOP") != nullptr;
  if (m <= 64 && !no_coop_final) ZK_LAUNCH(k_msm_final_coop<<<(unsigned)m, 128, 0, st>>>(groups.as<G1XYZZ>(), cfg.NG, split, cfg.c, out));   // few rows: latency
  else ZK_LAUNCH(k_msm_final<<<div_up(m, 64), 64, 0, st>>>(groups.as<G1XYZZ>(), m, cfg.NG, split, cfg.c, out));
  return ZK_OK;
}

static int me_open_run(const zkdl_g1_table* gens, const Fr* t, size_t n, const zkdl_fr_t* u_host, size_t k, G1Jac* proof, Fr* ret, cudaStream_t st) {
  ZK_REQUIRE(n == gens->n, ZK_ERR_DIM, "Incompatible dimensions");                       // commitment.cu:64
  ZK_REQUIRE(k < 31 && n == ((size_t)1 << k), ZK_ERR_DIM, "Incompatible dimensions");    // even halving at every round (commitment.cu:46)
  Scratch ud, rows, sLv, wLv; int rc;
  if ((rc = ud.alloc(sizeof(Fr) * (k ? k : 1), st))) return rc;
  if (k) ZK_CUDA(cudaMemcpyAsync(ud.p, u_host, sizeof(Fr) * k, cudaMemcpyHostToDevice, st));
  if ((rc = rows.alloc(sizeof(Fr) * (3 * k + 1) * n, st))) return rc;
  if ((rc = sLv.alloc(sizeof(Fr) * 2 * n, st))) return rc;
  if ((rc = wLv.alloc(sizeof(Fr) * 2 * n, st))) return rc;
  ZK_LAUNCH(k_open_fold<<<1, 1024, 0, st>>>(t, ud.as<Fr>(), (int)k, n, sLv.as<Fr>(), wLv.as<Fr>(), ret));
  ZK_LAUNCH(k_open_rows<<<g1_grid((k + 1) * n, 256), 256, 0, st>>>(sLv.as<Fr>(), wLv.as<Fr>(), (int)k, n, rows.as<Fr>()));
  return msm_run(gens, rows.as<Fr>(), 3 * k + 1, 0, 0, proof, st);
}

bool mmw_usable(const zkdl_mm_weights* p, size_t n);
int wfold_rows(const zkdl_mm_weights* p, size_t window, const zkdl_fr_t* u_host, size_t k, Fr* out, cudaStream_t st);

int open_run(const zkdl_g1_table* gens, const zkdl_g1_table* com_table, const Fr* t, size_t nt, const zkdl_mm_weights* t_int,
             const zkdl_fr_t* u_host, size_t ku, G1Jac* com_eval, G1Jac* proof, Fr* ret, cudaStream_t st) {
  size_t ncom = com_table->n;
  size_t khi = zkdl_ceil_log2((uint32_t)ncom);
  ZK_REQUIRE(ku >= khi, ZK_ERR_DIM, "Incompatible dimensions");
  size_t klo = ku - khi;
  const zkdl_fr_t* u_hi = u_host + klo;
  ZK_REQUIRE(klo < 31 && gens->n == ((size_t)1 << klo), ZK_ERR_DIM, "Incompatible dimensions");   // commitment.cu:89
  // com(u_hi): G1_me == MSM of com against eq(u_hi, .) (g1-tensor.cu:463-491); guard of G1TensorJacobian::operator()(u)
  if (khi > 0) ZK_REQUIRE(!(ncom <= ((size_t)1 << (khi - 1)) || ncom > ((size_t)1 << khi)), ZK_ERR_DIM, "Incompatible dimensions");
  // com(u_hi) and the opening proper are independent: fork the commitment-vector evaluation onto a side stream so its
  // latency-bound bucket reduction overlaps the opening MSM (joined before returning).
  SideStream& ss = side_stream(0, st);
  Scratch uq, E, tf; int rc;
  ForkScope fs(ss, st);
  if ((rc = fs.fork())) return rc;
  cudaStream_t side = ss.stream;
  {
    Scratch uqs, Es;
    if ((rc = uqs.alloc(sizeof(Fr) * (khi ? khi : 1), side))) return rc;
    if (khi) ZK_CUDA(cudaMemcpyAsync(uqs.p, u_hi, sizeof(Fr) * khi, cudaMemcpyHostToDevice, side));
    if ((rc = Es.alloc(sizeof(Fr) * ((size_t)1 << khi), side))) return rc;
    if ((rc = build_eq_table(uqs.as<Fr>(), u_hi, (int)khi, 0, Es.as<Fr>(), side))) return rc;
    if ((rc = msm_run(com_table, Es.as<Fr>(), 1, 1, 0, com_eval, side))) return rc;
  }
  // t.partial_me(u_hi, |gens|)
  size_t w = gens->n;
  if (khi > 0) ZK_REQUIRE(nt > w * ((size_t)1 << (khi - 1)), ZK_ERR_DIM, "Incompatible dimensions");   // fr-tensor.cu:372
  size_t tf_n = zkdl_partial_me_size(nt, khi, w);
  ZK_REQUIRE(tf_n == w, ZK_ERR_DIM, "Incompatible dimensions");                                         // commitment.cu:64
  if ((rc = tf.alloc(sizeof(Fr) * tf_n, st))) return rc;
  if (mmw_usable(t_int, nt)) rc = wfold_rows(t_int, w, u_hi, khi, tf.as<Fr>(), st);     // quantised weights: fold the integers
  else rc = fr_partial_me_dev(t, nt, u_hi, khi, w, tf.as<Fr>(), st);
  if (rc) return rc;
  rc = me_open_run(gens, tf.as<Fr>(), tf_n, u_host, klo, proof, ret, st);
  int rj = fs.join();
  return rc ? rc : rj;
}

}  // namespace zk

extern "C" {

int zkdl_g1_elementwise(int op, const zkdl_g1_jacobian_t* a, const void* b, size_t nb, zkdl_g1_jacobian_t* out, size_t n, void* stream) {
  if (n == 0) return ZK_OK;
  ZK_REQUIRE(op >= ZKDL_G1_ADD && op <= ZKDL_G1_MSUB, ZK_ERR_ARG, "bad op");
  if (op != ZKDL_G1_NEG) ZK_REQUIRE(b && (nb == 1 || nb == n), ZK_ERR_DIM, "Incompatible dimensions");
  ZK_LAUNCH(k_g1_elementwise<<<g1_grid(n, G1_THREADS), G1_THREADS, 0, S(stream)>>>(op, reinterpret_cast<const G1Jac*>(a), b, nb, reinterpret_cast<G1Jac*>(out), n));
  return ZK_OK;
}
int zkdl_g1_affine_to_jacobian(const zkdl_g1_affine_t* a, zkdl_g1_jacobian_t* out, size_t n, void* stream) {
  if (n == 0) return ZK_OK;
  ZK_LAUNCH(k_g1_affine_to_jac<<<g1_grid(n, G1_THREADS), G1_THREADS, 0, S(stream)>>>(reinterpret_cast<const G1Affine*>(a), reinterpret_cast<G1Jac*>(out), n));
  return ZK_OK;
}
int zkdl_g1_mul(const zkdl_g1_jacobian_t* P, size_t np, const zkdl_fr_t* x, size_t n, zkdl_g1_jacobian_t* out, void* stream) {
  if (n == 0) return ZK_OK;
  ZK_REQUIRE(np > 0 && n % np == 0, ZK_ERR_DIM, "Incompatible dimensions");                // g1-tensor.cu:448
  ZK_LAUNCH(k_g1_mul<<<g1_grid(n, G1_THREADS), G1_THREADS, 0, S(stream)>>>(reinterpret_cast<const G1Jac*>(P), np, reinterpret_cast<const Fr*>(x), n, reinterpret_cast<G1Jac*>(out)));
  return ZK_OK;
}
int zkdl_g1_sum(const zkdl_g1_jacobian_t* a, size_t n, zkdl_g1_jacobian_t* out, void* stream) {
  cudaStream_t st = S(stream);
  unsigned grid = g1_grid(n, G1_THREADS); if (grid > 1024) grid = 1024;
  Scratch parts; int rc = parts.alloc(sizeof(G1XYZZ) * grid, st); if (rc) return rc;
  ZK_LAUNCH(k_g1_sum_partial<<<grid, G1_THREADS, sizeof(G1XYZZ) * G1_THREADS, st>>>(reinterpret_cast<const G1Jac*>(a), n, parts.as<G1XYZZ>()));
  ZK_LAUNCH(k_g1_sum_final<<<1, G1_THREADS, sizeof(G1XYZZ) * G1_THREADS, st>>>(parts.as<G1XYZZ>(), grid, reinterpret_cast<G1Jac*>(out)));
  return ZK_OK;
}
int zkdl_g1_normalize(const zkdl_g1_jacobian_t* a, zkdl_g1_jacobian_t* out, size_t n, void* stream) {
  if (n == 0) return ZK_OK;
  ZK_LAUNCH(k_g1_normalize<<<g1_grid(n, G1_THREADS), G1_THREADS, 0, S(stream)>>>(reinterpret_cast<const G1Jac*>(a), reinterpret_cast<G1Jac*>(out), n));
  return ZK_OK;
}

int zkdl_g1_table_create(const zkdl_g1_jacobian_t* points, size_t n, int window_bits, int full, zkdl_g1_table** out, void* stream) {
  cudaStream_t st = S(stream);
  ZK_REQUIRE(points && out && n > 0, ZK_ERR_ARG, "bad table arguments");
  ZK_REQUIRE(window_bits == 0 || window_bits == TABLE_C, ZK_ERR_ARG, "only 4-bit table windows are supported");
  int windows = full ? TABLE_W : 1;
  size_t total = n * (size_t)windows;
  zkdl_g1_table* t = new zkdl_g1_table();
  t->n = n; t->full = full; t->windows = windows; t->bytes = total * sizeof(G1Affine); t->pts = nullptr;
  cudaError_t e = cudaMalloc(&t->pts, t->bytes);
  if (e != cudaSuccess) { delete t; set_last_error("cudaMalloc table: %s", cudaGetErrorString(e)); return ZK_ERR_CUDA; }
  Scratch tmp; int rc = tmp.alloc(sizeof(G1XYZZ) * total, st);
  if (rc) { cudaFree(t->pts); delete t; return rc; }
  k_table_expand<<<div_up(n, G1_THREADS), G1_THREADS, 0, st>>>(reinterpret_cast<const G1Jac*>(points), n, windows, tmp.as<G1XYZZ>());
  k_batch_affine<<<div_up(div_up(total, INV_CH), G1_THREADS), G1_THREADS, 0, st>>>(tmp.as<G1XYZZ>(), t->pts, total);
  zk::g_launches.fetch_add(2);
  e = cudaGetLastError();
  if (e != cudaSuccess) { cudaFree(t->pts); delete t; set_last_error("table kernels: %s", cudaGetErrorString(e)); return ZK_ERR_CUDA; }
  *out = t;
  return ZK_OK;
}
int zkdl_g1_table_destroy(zkdl_g1_table* t) {
  if (!t) return ZK_OK;
  cudaFree(t->pts);
  delete t;
  return ZK_OK;
}
size_t zkdl_g1_table_size(const zkdl_g1_table* t) { return t ? t->n : 0; }
size_t zkdl_g1_table_bytes(const zkdl_g1_table* t) { return t ? t->bytes : 0; }

int zkdl_msm(const zkdl_g1_table* t, const zkdl_fr_t* scalars, size_t m, int scalars_mont, zkdl_g1_jacobian_t* out, void* stream) {
  ZK_REQUIRE(t && scalars && out, ZK_ERR_ARG, "null argument");
  return msm_run(t, reinterpret_cast<const Fr*>(scalars), m, scalars_mont, 0, reinterpret_cast<G1Jac*>(out), S(stream));
}
int zkdl_commit(const zkdl_g1_table* gens, const zkdl_fr_t* t, size_t nt, zkdl_g1_jacobian_t* com, void* stream) {
  ZK_REQUIRE(gens && t && com, ZK_ERR_ARG, "null argument");
  ZK_REQUIRE(nt % gens->n == 0, ZK_ERR_DIM, "Incompatible dimensions");                    // commitment.cu:31
  return msm_run(gens, reinterpret_cast<const Fr*>(t), nt / gens->n, 1, 0, reinterpret_cast<G1Jac*>(com), S(stream));
}
int zkdl_me_open(const zkdl_g1_table* gens, const zkdl_fr_t* t, size_t n, const zkdl_fr_t* u_host, size_t k,
                 zkdl_g1_jacobian_t* proof, zkdl_fr_t* ret, void* stream) {
  ZK_REQUIRE(gens && t && proof && ret, ZK_ERR_ARG, "null argument");
  return me_open_run(gens, reinterpret_cast<const Fr*>(t), n, u_host, k, reinterpret_cast<G1Jac*>(proof), reinterpret_cast<Fr*>(ret), S(stream));
}
int zkdl_open(const zkdl_g1_table* gens, const zkdl_g1_table* com_table, const zkdl_fr_t* t, size_t nt, const zkdl_fr_t* u_host, size_t ku,
              zkdl_g1_jacobian_t* com_eval, zkdl_g1_jacobian_t* proof, zkdl_fr_t* ret, void* stream) {
  ZK_REQUIRE(gens && com_table && t && com_eval && proof && ret, ZK_ERR_ARG, "null argument");
  return open_run(gens, com_table, reinterpret_cast<const Fr*>(t), nt, nullptr, u_host, ku, reinterpret_cast<G1Jac*>(com_eval),
                  reinterpret_cast<G1Jac*>(proof), reinterpret_cast<Fr*>(ret), S(stream));
}
int zkdl_g1_me(const zkdl_g1_jacobian_t* a, size_t n, const zkdl_fr_t* u_host, size_t k, zkdl_g1_jacobian_t* out, void* stream) {
  cudaStream_t st = S(stream);
  ZK_REQUIRE(k < 31, ZK_ERR_DIM, "Incompatible dimensions");
  if (k == 0) { ZK_REQUIRE(n == 1, ZK_ERR_DIM, "Incompatible dimensions"); ZK_CUDA(cudaMemcpyAsync(out, a, sizeof(G1Jac), cudaMemcpyDeviceToDevice, st)); return ZK_OK; }
  ZK_REQUIRE(!(n <= ((size_t)1 << (k - 1)) || n > ((size_t)1 << k)), ZK_ERR_DIM, "Incompatible dimensions");   // g1-tensor.cu:488
  // a one-shot (plain Pippenger) base table in stream-ordered scratch memory: no cudaMalloc, no host synchronisation
  Scratch tmp, pts, uq, E; int rc;
  if ((rc = tmp.alloc(sizeof(G1XYZZ) * n, st))) return rc;
  if ((rc = pts.alloc(sizeof(G1Affine) * n, st))) return rc;
  ZK_LAUNCH(k_table_expand<<<div_up(n, G1_THREADS), G1_THREADS, 0, st>>>(reinterpret_cast<const G1Jac*>(a), n, 1, tmp.as<G1XYZZ>()));
  ZK_LAUNCH(k_batch_affine<<<div_up(div_up(n, INV_CH), G1_THREADS), G1_THREADS, 0, st>>>(tmp.as<G1XYZZ>(), pts.as<G1Affine>(), n));
  zkdl_g1_table tab; tab.n = n; tab.full = 0; tab.windows = 1; tab.pts = pts.as<G1Affine>(); tab.bytes = sizeof(G1Affine) * n;
  if ((rc = uq.alloc(sizeof(Fr) * k, st))) return rc;
  ZK_CUDA(cudaMemcpyAsync(uq.p, u_host, sizeof(Fr) * k, cudaMemcpyHostToDevice, st));
  if ((rc = E.alloc(sizeof(Fr) * ((size_t)1 << k), st))) return rc;
  if ((rc = build_eq_table(uq.as<Fr>(), u_host, (int)k, 0, E.as<Fr>(), st))) return rc;
  return msm_run(&tab, E.as<Fr>(), 1, 1, 0, reinterpret_cast<G1Jac*>(out), st);
}

}  // extern "C"
