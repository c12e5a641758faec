// matmul.cu — the quantised forward product zkFC::operator() (/root/reference/zkfc.cu:6-47,117-126):
// C[rowsA x colsB] = A[rowsA x colsA] * W[colsA x colsB] over Fr, bit-identical to the reference's Montgomery dot products.
//
// Three kernels, routed on the device (no host synchronisation) by what the operands contain:
//   k_tc_matmul   operands are small signed integers (|a| < 2^23, |w| < 2^15: the demo's 2^16-scaled activations and
//                 weights): byte planes on the int8 tensor cores (mma.sync m16n8k32, s32 accumulators per byte shift),
//                 exact because every partial sum stays below 2^31; needs the prepared weight copy (zkdl_mm_weights)
//   k_i32_matmul  operands fit int32: SIMT integer dot products, 64-bit accumulators when max|A| max|W| K < 2^62, else 128-bit
//   k_fr_matmul   anything else: Montgomery products in Fr
// The exact integer dot product reduced mod p is the same field element as the Fr dot product, so all three agree.
// The forward pass is outside the reference's timed region (demo.cu:124-138) and inside bench.py's e2e leg.
#include <atomic>
#include <stdlib.h>
#include "common.cuh"
#include "fr_device.cuh"
#include "../../include/zkdl_b200.h"

struct zkdl_mm_weights {
  size_t rows, cols;            // K x N
  int32_t* w32;                 // [K][N]
  uint8_t* planes;              // [2][N][K] (W transposed, K contiguous) or nullptr when the shape does not tile
  uint32_t* info;               // device: {not-small flag, 0, max|w|, 0}
  bool small;                   // host copy of !flag: every entry is a 32-bit signed integer
};

namespace zk {
bool umma_matmul_shape_ok(size_t M, size_t K, size_t N);
int umma_matmul_launch(const uint8_t* Ap, const uint8_t* Wp, Fr* C, size_t M, size_t K, size_t N, const uint32_t* info, const ReluOut& ro, cudaStream_t st);

static constexpr int THREADS = 256;
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline const Fr* F(const zkdl_fr_t* p) { return reinterpret_cast<const Fr*>(p); }
static inline Fr* F(zkdl_fr_t* p) { return reinterpret_cast<Fr*>(p); }
static inline unsigned stream_grid(size_t work_items, int threads) {
  size_t blocks = (work_items + threads - 1) / threads;
  size_t cap = (size_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks ? blocks : 1);
}

static constexpr int MM_TILE = 16;
// C = A * B over Fr; 16x16 output tile per CTA, K-tiles staged through shared memory.
__global__ void __launch_bounds__(MM_TILE * MM_TILE) k_fr_matmul(const Fr* __restrict__ A, const Fr* __restrict__ B, Fr* __restrict__ C,
                                                                 size_t rowsA, size_t colsA, size_t colsB, const uint32_t* __restrict__ only_if) {
  if (only_if && *only_if == 0) return;                      // the small-integer fast path already produced C
  __shared__ Fr As[MM_TILE][MM_TILE];
  __shared__ Fr Bs[MM_TILE][MM_TILE];
  const int tx = threadIdx.x % MM_TILE, ty = threadIdx.x / MM_TILE;
  const size_t row = (size_t)blockIdx.y * MM_TILE + ty, col = (size_t)blockIdx.x * MM_TILE + tx;
  Fr sum = Fr::zero();
  for (size_t k0 = 0; k0 < colsA; k0 += MM_TILE) {
    As[ty][tx] = (row < rowsA && k0 + tx < colsA) ? A[row * colsA + k0 + tx] : Fr::zero();
    Bs[ty][tx] = (k0 + ty < colsA && col < colsB) ? B[(k0 + ty) * colsB + col] : Fr::zero();
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < MM_TILE; ++k) sum = add(sum, mul(As[ty][k], Bs[k][tx]));
    __syncthreads();
  }
  if (row < rowsA && col < colsB) C[row * colsB + col] = sum;
}

// ---- small-integer fast path of the forward matmul.  Quantised activations and weights are Montgomery forms of small
// signed integers (|v| < 2^31: inputs at scale 2^16, rescaled activations are u32 magnitudes, zkrelu.cu:29).  The exact
// integer dot product (128-bit accumulator) reduced mod p is the same field element as the Fr dot product, so the result
// is bit-identical; ~5 integer instructions per multiply-add instead of a 136-IMAD Montgomery product.  Any operand
// outside the range raises `flag`, and the generic Fr kernel (launched right after, exiting early otherwise) recomputes.
// flag[0] |= 1 if an element is not a 32-bit signed integer; *maxmag = max |element| (decides the accumulator width)
// PLANES > 0: also the low PLANES bytes of the two's-complement value as byte planes (planes[p * n + i]), the operand
// layout of the int8 tensor-core kernel below.
template <int PLANES>
__global__ void __launch_bounds__(THREADS) k_fr_to_i32(const Fr* __restrict__ in, int32_t* __restrict__ out, size_t n, uint32_t* __restrict__ flag,
                                                       uint32_t* __restrict__ maxmag, uint8_t* __restrict__ planes) {
  uint32_t mymax = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr x = from_mont(in[i]);
    bool hi0 = (x.v[1] | x.v[2] | x.v[3] | x.v[4] | x.v[5] | x.v[6] | x.v[7]) == 0;
    int32_t r = 0;
    if (hi0 && x.v[0] < 0x80000000u) r = (int32_t)x.v[0];
    else {
      Fr m = sub(Fr::zero(), x);                              // p - x
      bool mhi0 = (m.v[1] | m.v[2] | m.v[3] | m.v[4] | m.v[5] | m.v[6] | m.v[7]) == 0;
      if (mhi0 && m.v[0] <= 0x80000000u) r = (int32_t)(0u - m.v[0]);
      else atomicOr(flag, 1u);
    }
    out[i] = r;
#pragma unroll
    for (int p = 0; p < PLANES; ++p) planes[(size_t)p * n + i] = (uint8_t)((uint32_t)r >> (8 * p));
    uint32_t mag = r < 0 ? 0u - (uint32_t)r : (uint32_t)r;
    if (mag > mymax) mymax = mag;
  }
  mymax = __reduce_max_sync(0xffffffffu, mymax);
  if ((threadIdx.x & 31) == 0 && mymax) atomicMax(maxmag, mymax);
}
static constexpr int IM_T = 64, IM_K = 32;                    // 64x64 outputs per CTA, 4x4 per thread, K tiles of 32
// Exact integer product tile.  ACC = int64_t when max|A| * max|W| * colsA < 2^62 (the quantised demo: 2^19 * 2^12 * 2^11),
// else __int128 (always exact for 32-bit operands).
template <typename ACC>
__device__ __forceinline__ void i32_matmul_tile(const int32_t* __restrict__ A, const int32_t* __restrict__ W, Fr* __restrict__ C,
                                                size_t rowsA, size_t colsA, size_t colsB,
                                                int32_t (*As)[IM_T + 4], int32_t (*Ws)[IM_T]) {
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const size_t row0 = (size_t)blockIdx.y * IM_T, col0 = (size_t)blockIdx.x * IM_T;
  ACC acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  for (size_t k0 = 0; k0 < colsA; k0 += IM_K) {
    for (int e = threadIdx.x; e < IM_T * IM_K; e += 256) {
      int r = e / IM_K, k = e % IM_K;                          // A: consecutive threads read consecutive k
      size_t gr = row0 + r, gk = k0 + k;
      As[k][r] = (gr < rowsA && gk < colsA) ? A[gr * colsA + gk] : 0;
      int kk = e / IM_T, c = e % IM_T;                         // W: consecutive threads read consecutive columns
      size_t gk2 = k0 + kk, gc = col0 + c;
      Ws[kk][c] = (gk2 < colsA && gc < colsB) ? W[gk2 * colsB + gc] : 0;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < IM_K; ++k) {
      const int4 av = *reinterpret_cast<const int4*>(&As[k][ty * 4]);
      const int4 wv = *reinterpret_cast<const int4*>(&Ws[k][tx * 4]);
      const int32_t a[4] = {av.x, av.y, av.z, av.w}, w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += (ACC)((int64_t)a[i] * (int64_t)w[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      size_t gr = row0 + ty * 4 + i, gc = col0 + tx * 4 + j;
      if (gr >= rowsA || gc >= colsB) continue;
      __int128 v = acc[i][j];
      bool negative = v < 0;
      unsigned __int128 m = negative ? (unsigned __int128)(-v) : (unsigned __int128)v;
      Fr r = Fr::zero();
      r.v[0] = (uint32_t)m; r.v[1] = (uint32_t)(m >> 32); r.v[2] = (uint32_t)(m >> 64); r.v[3] = (uint32_t)(m >> 96);
      r = to_mont(r);
      C[gr * colsB + gc] = negative ? neg(r) : r;
    }
}
// info[0] = not-all-small flag, info[1] = max|A|, info[2] = max|W|, info[3] = force the 128-bit accumulator (tuning knob),
// info[4] = the tensor-core kernel handles this product (set by k_mm_route)
__global__ void __launch_bounds__(256) k_i32_matmul(const int32_t* __restrict__ A, const int32_t* __restrict__ W, Fr* __restrict__ C,
                                                    size_t rowsA, size_t colsA, size_t colsB, const uint32_t* __restrict__ info) {
  __shared__ __align__(16) int32_t As[IM_K][IM_T + 4];        // [k][row], padded: 16-byte aligned rows, 4-way store conflicts at most
  __shared__ __align__(16) int32_t Ws[IM_K][IM_T];            // [k][col]
  if (info[0] || info[4]) return;                              // the generic Fr kernel / the tensor-core kernel takes over
  const unsigned long long bound = (unsigned long long)info[1] * (unsigned long long)info[2];
  if (!info[3] && bound <= (1ull << 62) / (colsA ? colsA : 1)) i32_matmul_tile<long long>(A, W, C, rowsA, colsA, colsB, As, Ws);
  else i32_matmul_tile<__int128>(A, W, C, rowsA, colsA, colsB, As, Ws);
}


// ---- int8 tensor-core path.  a = a2 * 2^16 + a1 * 2^8 + a0 (a2 signed, a1, a0 unsigned bytes; |a| < 2^23),
// w = w1 * 2^8 + w0 (w1 signed, w0 unsigned; |w| < 2^15).  The six byte-plane products are accumulated per shift
// s = i + j in separate s32 accumulators: at most 2 products x 255 x 255 x K < 2^31 for K <= 16384.
static constexpr int TC_A_PLANES = 3, TC_W_PLANES = 2, TC_SHIFTS = 4;
static constexpr int TC_M = 64, TC_N = 64, TC_K = 64;          // CTA tile; 8 warps: 4 row blocks of 16 x 2 column blocks of 32
static constexpr int TC_STRIDE = TC_K + 16;                    // smem row stride (bytes): 20 words, conflict-free fragment loads
static constexpr uint32_t TC_A_LIMIT = 1u << 23, TC_W_LIMIT = 1u << 15;
static constexpr size_t TC_MAX_K = 16384;

template <bool SA, bool SB>
__device__ __forceinline__ void mma_i8(int32_t (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if (SA && SB)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else if (SA)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else if (SB)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// info[4] = 1 iff the tensor-core kernel takes this product (operands small enough); evaluated once, on the device
__global__ void k_mm_route(uint32_t* __restrict__ info, int have_planes) {
  info[4] = (have_planes && !info[0] && !info[3] && info[1] < TC_A_LIMIT && info[2] < TC_W_LIMIT) ? 1u : 0u;
}

// Ap: [3][M][K] byte planes of A, Wp: [2][N][K] byte planes of W^T.  M % 64 == 0, N % 64 == 0, K % 64 == 0.
__global__ void __launch_bounds__(256) k_tc_matmul(const uint8_t* __restrict__ Ap, const uint8_t* __restrict__ Wp, Fr* __restrict__ C,
                                                   size_t M, size_t K, size_t N, const uint32_t* __restrict__ info) {
  if (!info[4]) return;
  __shared__ __align__(16) uint8_t As[TC_A_PLANES][TC_M][TC_STRIDE];
  __shared__ __align__(16) uint8_t Ws[TC_W_PLANES][TC_N][TC_STRIDE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;
  const size_t row0 = (size_t)blockIdx.y * TC_M, col0 = (size_t)blockIdx.x * TC_N;
  int32_t acc[TC_SHIFTS][4][4];
#pragma unroll
  for (int s = 0; s < TC_SHIFTS; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[s][j][e] = 0;
  for (size_t k0 = 0; k0 < K; k0 += TC_K) {
    // 16-byte chunks: A 3 x 64 rows x 4, W 2 x 64 rows x 4
    for (int c = threadIdx.x; c < (TC_A_PLANES + TC_W_PLANES) * 64 * 4; c += 256) {
      const int plane = c / 256, r = (c % 256) / 4, q = c % 4;
      if (plane < TC_A_PLANES) {
        const int4 v = *reinterpret_cast<const int4*>(Ap + ((size_t)plane * M + row0 + r) * K + k0 + q * 16);
        *reinterpret_cast<int4*>(&As[plane][r][q * 16]) = v;
      } else {
        const int pw = plane - TC_A_PLANES;
        const int4 v = *reinterpret_cast<const int4*>(Wp + ((size_t)pw * N + col0 + r) * K + k0 + q * 16);
        *reinterpret_cast<int4*>(&Ws[pw][r][q * 16]) = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TC_K; kk += 32) {
      uint32_t a[TC_A_PLANES][4], b[TC_W_PLANES][4][2];
#pragma unroll
      for (int p = 0; p < TC_A_PLANES; ++p) {
        a[p][0] = *reinterpret_cast<const uint32_t*>(&As[p][wm + g][kk + t * 4]);
        a[p][1] = *reinterpret_cast<const uint32_t*>(&As[p][wm + g + 8][kk + t * 4]);
        a[p][2] = *reinterpret_cast<const uint32_t*>(&As[p][wm + g][kk + 16 + t * 4]);
        a[p][3] = *reinterpret_cast<const uint32_t*>(&As[p][wm + g + 8][kk + 16 + t * 4]);
      }
#pragma unroll
      for (int p = 0; p < TC_W_PLANES; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          b[p][j][0] = *reinterpret_cast<const uint32_t*>(&Ws[p][wn + j * 8 + g][kk + t * 4]);
          b[p][j][1] = *reinterpret_cast<const uint32_t*>(&Ws[p][wn + j * 8 + g][kk + 16 + t * 4]);
        }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mma_i8<false, false>(acc[0][j], a[0], b[0][j]);
        mma_i8<false, false>(acc[1][j], a[1], b[0][j]);
        mma_i8<false, true>(acc[1][j], a[0], b[1][j]);
        mma_i8<true, false>(acc[2][j], a[2], b[0][j]);
        mma_i8<false, true>(acc[2][j], a[1], b[1][j]);
        mma_i8<true, true>(acc[3][j], a[2], b[1][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const size_t gr = row0 + wm + g + (e >> 1) * 8, gc = col0 + wn + j * 8 + t * 2 + (e & 1);
      long long v = 0;
#pragma unroll
      for (int s = TC_SHIFTS - 1; s >= 0; --s) v = v * 256 + (long long)acc[s][j][e];     // |v| < 2^29 * 2^24 * 1.01
      const bool negative = v < 0;
      const unsigned long long m = negative ? (unsigned long long)(-v) : (unsigned long long)v;
      Fr r = Fr::zero();
      r.v[0] = (uint32_t)m; r.v[1] = (uint32_t)(m >> 32);
      r = to_mont(r);
      C[gr * N + gc] = negative ? neg(r) : r;
    }
}

// W^T byte planes from the int32 copy: planes[(p * N + n) * K + k] = byte p of W[k][n]
__global__ void __launch_bounds__(THREADS) k_w_planes(const int32_t* __restrict__ w32, uint8_t* __restrict__ planes, size_t K, size_t N) {
  const size_t total = K * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t k = i / N, n = i - k * N;
    const uint32_t v = (uint32_t)w32[i];
#pragma unroll
    for (int p = 0; p < TC_W_PLANES; ++p) planes[((size_t)p * N + n) * K + k] = (uint8_t)(v >> (8 * p));
  }
}

// ---- folds of the quantised weight table against an eq table (zkFC::prove's two weights.partial_me calls, zkfc.cu:139
// and commitment.cu:88): out = sum_i w_i * E_i with the integers from zkdl_mm_weights, bit-identical to the chain of
// Fr_partial_me_step folds (a multilinear evaluation is the eq-weighted sum) at ~1/10 of its multiplications.
int build_eq_table(const Fr* q_dev, const zkdl_fr_t* q_host, int t, int rev, Fr* E, cudaStream_t st);

// Signed weights without a sign branch: w + 2^31 is an unsigned 32-bit factor, so
//   sum_i w_i E_i = sum_i (w_i + 2^31) E_i - 2^31 sum_i E_i,
// one accumulator and one correction per output.  Over a complete eq table sum_i E_i = 1 (Montgomery one), so the correction
// is the constant mont(2^31); a zero-padded (partial) row range sums the E it covers.  (The first version kept separate
// positive / negative accumulators: the divergent pair cost ~100 instructions per multiply-add.)
__device__ __forceinline__ uint32_t w_offset(int32_t w) { return (uint32_t)w ^ 0x80000000u; }
__device__ __forceinline__ Fr offset_correction(const Fr& esum) {        // 2^31 * esum mod p
  ISum z; isum_zero(z); isum_mac(z, 0x80000000u, esum); return isum_reduce(z);
}
__device__ __forceinline__ void isum_add_fr(ISum& a, const Fr& e) {
  uint64_t c = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { uint64_t t = (uint64_t)a.v[j] + e.v[j] + c; a.v[j] = (uint32_t)t; c = t >> 32; }
  uint64_t t = (uint64_t)a.v[8] + c; a.v[8] = (uint32_t)t; a.v[9] += (uint32_t)(t >> 32);
}
__device__ __forceinline__ ISum isum_shfl_down(const ISum& a, int d) {
  ISum r;
#pragma unroll
  for (int j = 0; j < 10; ++j) r.v[j] = __shfl_down_sync(0xffffffffu, a.v[j], d);
  return r;
}
// window 1: out[r] = sum_c w[r][c] * E[c], c over the WHOLE eq table (cols = 2^k); one warp per row
__global__ void __launch_bounds__(256) k_wfold_cols(const int32_t* __restrict__ w, const Fr* __restrict__ E, size_t rows, size_t cols, Fr* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const Fr corr = offset_correction(Fr::one());
  for (size_t r = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += (size_t)gridDim.x * (blockDim.x >> 5)) {
    ISum acc; isum_zero(acc);
    const int32_t* row = w + r * cols;
    for (size_t c = lane; c < cols; c += 32) isum_mac(acc, w_offset(row[c]), E[c]);
    for (int d = 16; d > 0; d >>= 1) { ISum o = isum_shfl_down(acc, d); isum_add(acc, o); }
    if (lane == 0) out[r] = sub(isum_reduce(acc), corr);
  }
}
// window W: partial[s * W + c] = sum over the rows of slice s of (w[r][c] + 2^31) * E[r]
__global__ void __launch_bounds__(128) k_wfold_rows(const int32_t* __restrict__ w, const Fr* __restrict__ E, size_t rows, size_t W, int slices,
                                                    ISum* __restrict__ partial) {
  const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (c >= W) return;
  const size_t per = (rows + slices - 1) / slices, r0 = s * per, r1 = r0 + per < rows ? r0 + per : rows;
  ISum acc; isum_zero(acc);
  for (size_t r = r0; r < r1; ++r) isum_mac(acc, w_offset(w[r * W + c]), E[r]);
  partial[(size_t)s * W + c] = acc;
}
// full: the rows cover the whole eq table (sum E = 1); otherwise the covered E are summed here (cold path)
__global__ void __launch_bounds__(128) k_wfold_rows_final(const ISum* __restrict__ partial, const Fr* __restrict__ E, size_t rows, size_t W, int slices,
                                                          int full, Fr* __restrict__ out) {
  const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (c >= W) return;
  ISum acc = partial[c];
  for (int s = 1; s < slices; ++s) isum_add(acc, partial[(size_t)s * W + c]);
  Fr esum = Fr::one();
  if (!full) {
    ISum es; isum_zero(es);
    for (size_t r = 0; r < rows; ++r) isum_add_fr(es, E[r]);
    esum = isum_reduce(es);
  }
  out[c] = sub(isum_reduce(acc), offset_correction(esum));
}

bool mmw_usable(const zkdl_mm_weights* p, size_t n) {
  static const bool off = getenv("ZKDL_NO_WFOLD") != nullptr;                               // tuning knob
  return p && p->small && !off && p->rows * p->cols == n;
}
static int eq_for(const zkdl_fr_t* u_host, size_t k, Scratch& ud, Scratch& E, cudaStream_t st) {
  int rc;
  if ((rc = ud.alloc(sizeof(Fr) * (k ? k : 1), st))) return rc;
  if (k) ZK_CUDA(cudaMemcpyAsync(ud.p, u_host, sizeof(Fr) * k, cudaMemcpyHostToDevice, st));
  if ((rc = E.alloc(sizeof(Fr) * ((size_t)1 << k), st))) return rc;
  return build_eq_table(ud.as<Fr>(), u_host, (int)k, 0, E.as<Fr>(), st);
}
// the table as [n / 2^k][2^k], folded over its columns (FrTensor::partial_me(u, 1)): out has n / 2^k entries
int wfold_cols(const zkdl_mm_weights* p, const zkdl_fr_t* u_host, size_t k, Fr* out, cudaStream_t st) {
  const size_t n = p->rows * p->cols, cols = (size_t)1 << k;
  ZK_REQUIRE(k < 31 && n % cols == 0, ZK_ERR_DIM, "Incompatible dimensions");
  const size_t rows = n / cols;
  Scratch ud, E; int rc;
  if ((rc = eq_for(u_host, k, ud, E, st))) return rc;
  ZK_LAUNCH(k_wfold_cols<<<(unsigned)((rows + 7) / 8 < 4096 ? (rows + 7) / 8 : 4096), 256, 0, st>>>(p->w32, E.as<Fr>(), rows, cols, out));
  return ZK_OK;
}
// the table as [rows][window], folded over its rows (FrTensor::partial_me(u, window)), rows <= 2^k: out has `window` entries
int wfold_rows(const zkdl_mm_weights* p, size_t window, const zkdl_fr_t* u_host, size_t k, Fr* out, cudaStream_t st) {
  const size_t n = p->rows * p->cols;
  ZK_REQUIRE(k < 31 && window > 0 && n % window == 0 && n / window <= ((size_t)1 << k), ZK_ERR_DIM, "Incompatible dimensions");
  const size_t rows = n / window;
  Scratch ud, E, part; int rc;
  if ((rc = eq_for(u_host, k, ud, E, st))) return rc;
  int slices = rows >= 32 ? 32 : (int)rows;
  if ((rc = part.alloc(sizeof(ISum) * slices * window, st))) return rc;
  dim3 grid(div_up(window, 128), slices);
  ZK_LAUNCH(k_wfold_rows<<<grid, 128, 0, st>>>(p->w32, E.as<Fr>(), rows, window, slices, part.as<ISum>()));
  ZK_LAUNCH(k_wfold_rows_final<<<div_up(window, 128), 128, 0, st>>>(part.as<ISum>(), E.as<Fr>(), rows, window, slices, rows == ((size_t)1 << k) ? 1 : 0, out));
  return ZK_OK;
}

// zkReLU::operator() (zkrelu.cu:44-52) on a finished product, unless the tcgen05 kernel already applied it in its epilogue
// (*skip != 0: the device-side route took the tensor-core path)
__global__ void __launch_bounds__(THREADS) k_relu_after(const Fr* __restrict__ Z, ReluOut ro, size_t n, const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    ReluParts p = relu_decompose(Z[i]);
    if (p.out_of_range && ro.bad) atomicAdd(ro.bad, 1u);
    ro.sign[i] = p.positive ? Fr::one() : Fr::zero();
    Fr q = Fr::zero(); q.v[0] = p.q;
    ro.act[i] = p.positive ? to_mont(q) : Fr::zero();
    ro.mag[i] = p.q; ro.rem[i] = p.r;
  }
}

static int matmul_run(const Fr* A, const Fr* W, const zkdl_mm_weights* prep, Fr* C, size_t rowsA, size_t colsA, size_t colsB, cudaStream_t st,
                      const ReluOut* relu = nullptr) {
  Scratch ai, wi, info, ap; int rc;
  const ReluOut none = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool fused = false;
  if (relu && relu->bad) ZK_CUDA(cudaMemsetAsync(relu->bad, 0, sizeof(uint32_t), st));
  const bool tc_shape = prep && prep->planes && rowsA % TC_M == 0;
  if ((rc = ai.alloc(sizeof(int32_t) * rowsA * colsA, st))) return rc;
  if ((rc = info.alloc(sizeof(uint32_t) * 8, st))) return rc;
  static const uint32_t force128 = getenv("ZKDL_MM_FORCE128") ? 1u : 0u;                     // tuning knobs
  static const bool no_tc = getenv("ZKDL_MM_NO_TC") != nullptr;
  uint32_t* inf = info.as<uint32_t>();
  ZK_CUDA(cudaMemsetAsync(inf, 0, sizeof(uint32_t) * 8, st));
  if (force128) ZK_CUDA(cudaMemsetAsync(inf + 3, 1, 1, st));
  const int32_t* w32;
  if (prep) {
    ZK_CUDA(cudaMemcpyAsync(inf, prep->info, sizeof(uint32_t) * 3, cudaMemcpyDeviceToDevice, st));   // weight flag and max|w|
    w32 = prep->w32;
  } else {
    if ((rc = wi.alloc(sizeof(int32_t) * colsA * colsB, st))) return rc;
    ZK_LAUNCH(k_fr_to_i32<0><<<stream_grid(colsA * colsB, THREADS), THREADS, 0, st>>>(W, wi.as<int32_t>(), colsA * colsB, inf, inf + 2, nullptr));
    w32 = wi.as<int32_t>();
  }
  if (tc_shape && !no_tc) {
    if ((rc = ap.alloc(TC_A_PLANES * rowsA * colsA, st))) return rc;
    ZK_LAUNCH(k_fr_to_i32<TC_A_PLANES><<<stream_grid(rowsA * colsA, THREADS), THREADS, 0, st>>>(A, ai.as<int32_t>(), rowsA * colsA, inf, inf + 1, ap.as<uint8_t>()));
    ZK_LAUNCH(k_mm_route<<<1, 1, 0, st>>>(inf, 1));
    // tcgen05 / TMEM / TMA kernel (matmul_umma.cu) when the shape tiles by 128 x 64 x 128, else the mma.sync kernel
    fused = relu && umma_matmul_shape_ok(rowsA, colsA, colsB);
    if (!umma_matmul_shape_ok(rowsA, colsA, colsB) ||
        umma_matmul_launch(ap.as<uint8_t>(), prep->planes, C, rowsA, colsA, colsB, inf, relu ? *relu : none, st) != ZK_OK) {
      fused = false;
      dim3 tgrid(div_up(colsB, TC_N), div_up(rowsA, TC_M));
      ZK_LAUNCH(k_tc_matmul<<<tgrid, 256, 0, st>>>(ap.as<uint8_t>(), prep->planes, C, rowsA, colsA, colsB, inf));
    }
  } else {
    ZK_LAUNCH(k_fr_to_i32<0><<<stream_grid(rowsA * colsA, THREADS), THREADS, 0, st>>>(A, ai.as<int32_t>(), rowsA * colsA, inf, inf + 1, nullptr));
  }
  dim3 igrid(div_up(colsB, IM_T), div_up(rowsA, IM_T));
  ZK_LAUNCH(k_i32_matmul<<<igrid, 256, 0, st>>>(ai.as<int32_t>(), w32, C, rowsA, colsA, colsB, inf));
  dim3 grid(div_up(colsB, MM_TILE), div_up(rowsA, MM_TILE));
  ZK_LAUNCH(k_fr_matmul<<<grid, MM_TILE * MM_TILE, 0, st>>>(A, W, C, rowsA, colsA, colsB, inf));
  if (relu) {
    const size_t n = rowsA * colsB;
    ZK_LAUNCH(k_relu_after<<<stream_grid(n, THREADS), THREADS, 0, st>>>(C, *relu, n, fused ? inf + 4 : nullptr));
  }
  return ZK_OK;
}

}  // namespace zk

using namespace zk;

extern "C" {

int zkdl_fr_matmul(const zkdl_fr_t* A, const zkdl_fr_t* B, zkdl_fr_t* C, size_t rowsA, size_t colsA, size_t colsB, void* stream) {
  if (rowsA == 0 || colsB == 0) return ZK_OK;
  ZK_REQUIRE(A && B && C, ZK_ERR_ARG, "null argument");
  return matmul_run(F(A), F(B), nullptr, F(C), rowsA, colsA, colsB, S(stream));
}

int zkdl_mm_weights_create(const zkdl_fr_t* W, size_t rows, size_t cols, zkdl_mm_weights** out, void* stream) {
  cudaStream_t st = S(stream);
  ZK_REQUIRE(W && out && rows > 0 && cols > 0, ZK_ERR_ARG, "bad weight arguments");
  zkdl_mm_weights* p = new zkdl_mm_weights();
  p->rows = rows; p->cols = cols; p->w32 = nullptr; p->planes = nullptr; p->info = nullptr; p->small = false;
  const bool tiles = rows % TC_K == 0 && cols % TC_N == 0 && rows <= TC_MAX_K;
  cudaError_t e = cudaMalloc(&p->w32, sizeof(int32_t) * rows * cols);
  if (e == cudaSuccess) e = cudaMalloc(&p->info, sizeof(uint32_t) * 4);
  if (e == cudaSuccess && tiles) e = cudaMalloc(&p->planes, (size_t)TC_W_PLANES * rows * cols);
  if (e != cudaSuccess) { zkdl_mm_weights_destroy(p); set_last_error("cudaMalloc weights: %s", cudaGetErrorString(e)); return ZK_ERR_CUDA; }
  e = cudaMemsetAsync(p->info, 0, sizeof(uint32_t) * 4, st);
  if (e != cudaSuccess) { zkdl_mm_weights_destroy(p); set_last_error("memset: %s", cudaGetErrorString(e)); return ZK_ERR_CUDA; }
  k_fr_to_i32<0><<<stream_grid(rows * cols, THREADS), THREADS, 0, st>>>(F(W), p->w32, rows * cols, p->info, p->info + 2, nullptr);
  if (tiles) k_w_planes<<<stream_grid(rows * cols, THREADS), THREADS, 0, st>>>(p->w32, p->planes, rows, cols);
  zk::g_launches.fetch_add(tiles ? 2 : 1);
  uint32_t flag = 1;                                            // set-up time: one read-back tells the host which fold path applies
  e = cudaMemcpyAsync(&flag, p->info, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { zkdl_mm_weights_destroy(p); set_last_error("weight kernels: %s", cudaGetErrorString(e)); return ZK_ERR_CUDA; }
  p->small = flag == 0;
  *out = p;
  return ZK_OK;
}

int zkdl_mm_weights_destroy(zkdl_mm_weights* p) {
  if (!p) return ZK_OK;
  cudaFree(p->w32); cudaFree(p->planes); cudaFree(p->info);
  delete p;
  return ZK_OK;
}

int zkdl_fr_matmul_prepared(const zkdl_fr_t* A, const zkdl_fr_t* W, const zkdl_mm_weights* prep, zkdl_fr_t* C, size_t rowsA, void* stream) {
  ZK_REQUIRE(prep, ZK_ERR_ARG, "null argument");
  if (rowsA == 0) return ZK_OK;
  ZK_REQUIRE(A && W && C, ZK_ERR_ARG, "null argument");
  return matmul_run(F(A), F(W), prep, F(C), rowsA, prep->rows, prep->cols, S(stream));
}

int zkdl_fr_matmul_prepared_relu(const zkdl_fr_t* A, const zkdl_fr_t* W, const zkdl_mm_weights* prep, zkdl_fr_t* Z, size_t rowsA,
                                 zkdl_fr_t* act, zkdl_fr_t* sign, uint32_t* mag_packed, uint16_t* rem_packed, uint32_t* out_of_range, void* stream) {
  ZK_REQUIRE(prep, ZK_ERR_ARG, "null argument");
  if (rowsA == 0) return ZK_OK;
  ZK_REQUIRE(A && W && Z && act && sign && mag_packed && rem_packed, ZK_ERR_ARG, "null argument");
  const ReluOut ro = {F(act), F(sign), mag_packed, rem_packed, out_of_range};
  return matmul_run(F(A), F(W), prep, F(Z), rowsA, prep->rows, prep->cols, S(stream), &ro);
}

}  // extern "C"
