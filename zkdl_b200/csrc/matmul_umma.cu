// matmul_umma.cu — the quantised forward product zkFC::operator() (/root/reference/zkfc.cu:6-47,117-126) on the
// 5th-generation tensor cores: tcgen05.mma kind::i8 with TMEM accumulators, operands staged by TMA.
//
// Same arithmetic as k_tc_matmul (matmul.cu): a = a2 2^16 + a1 2^8 + a0 (a2 signed, a1, a0 unsigned bytes; |a| < 2^23),
// w = w1 2^8 + w0 (w1 signed, w0 unsigned; |w| < 2^15); the six byte-plane products are accumulated per byte shift in FOUR
// s32 accumulators (every partial sum < 2^31 for K <= 16384), recombined exactly in the epilogue, reduced mod p and
// written as Montgomery Fr: bit-identical to the reference's Montgomery dot products.
//
// One CTA = one 128 x 64 output tile, 512 threads:
//   warp 0 / lane 0  TMA producer: per 128-byte k-block five 2-D tile loads (3 A planes 128 x 128 B, 2 W^T planes 64 x 128 B,
//                    SWIZZLE_128B) into a 3-stage shared-memory ring, completion on the stage's `full` mbarrier
//   warp 1 / lane 0  MMA issuer: 4 k-steps (UMMA_K = 32 bytes) x 6 tcgen05.mma.kind::i8 (M 128, N 64) per k-block into four
//                    64-column TMEM accumulators (u8 x u8, u8 x s8, s8 x u8, s8 x s8 instruction descriptors),
//                    tcgen05.commit frees the stage / signals the epilogue
//   warps 0-15       epilogue: tcgen05.ld 32x32b (warp w reads TMEM lanes 32 (w % 4) .. +31 = output rows, columns
//                    16 (w / 4) .. +15), recombination, to_mont, 32-byte stores.  (With 4 warps the epilogue - 64 dependent
//                    Montgomery products per thread, one warp per scheduler - took 4x longer than the whole main loop.)
//                    Integers go to Montgomery form with to_mont_u64 (two CIOS rows).  With a ReluOut the epilogue also
//                    applies zkReLU::operator() to the exact integer: activation, sign and the packed decomposition
//                    are written from the same registers, so the separate relu pass (re-read Z, from_mont) disappears.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "fr_device.cuh"
#include "../../include/zkdl_b200.h"

namespace zk {
namespace umma {

constexpr int BM = 128, BN = 64, BK = 128, STAGES = 3, NA = 3, NW = 2;
constexpr int A_BYTES = BM * BK, W_BYTES = BN * BK;                  // one plane of a stage
constexpr int STAGE_BYTES = NA * A_BYTES + NW * W_BYTES;             // 64 KiB
constexpr int TMEM_COLS = 256;                                       // 4 accumulators x 64 columns
constexpr int NTHREADS = 512;                                        // 16 warps: 4 lane quarters x 4 column chunks of 16 in the epilogue
static_assert(BN == 16 * (NTHREADS / 128), "one 16-column chunk per epilogue warp");
constexpr int EPI_ROW = 4 * 32 + 16, EPI_BUF = 32 * EPI_ROW;         // epilogue transpose: 32 rows x 4 elements, padded rows
static_assert((size_t)(NTHREADS / 32) * 2 * EPI_BUF <= (size_t)STAGES * STAGE_BYTES, "the epilogue slices live in the idle stage ring");
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers + TMEM slot */;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// spins on try_wait; a wait that lasts seconds is a protocol bug: trap instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO); LBO is unused for swizzled K-major
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;                                             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                                             // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::i8: D = S32, A / B unsigned (0) or signed (1) 8-bit, both K-major, N = 64, M = 128
__host__ __device__ constexpr uint32_t idesc(uint32_t a_signed, uint32_t b_signed) {
  return (2u << 4) | (a_signed << 7) | (b_signed << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t id, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(id), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16-byte shared-memory accesses by 32-bit shared address: the staging pointer is derived through an integer round-up, so the
// compiler would otherwise emit generic LD / ST (L1TEX address translation, long-scoreboard latency) instead of LDS / STS
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1) k_umma_matmul(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                                                        Fr* __restrict__ C, uint32_t M, uint32_t K, uint32_t N, const uint32_t* __restrict__ info,
                                                        ReluOut ro) {
  if (!info[4]) return;                                               // routed elsewhere (operands not small enough)
  extern __shared__ uint8_t umma_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(umma_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum = smem_u32(bars + 2 * STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t nkb = K / BK;

  if (warp == 0) {
    if (lane == 0) {                                                  // ---- TMA producer
      for (uint32_t kb = 0; kb < nkb; ++kb) {
        const uint32_t s = kb % STAGES, ph = (kb / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t st = smem_base + s * STAGE_BYTES, full = full0 + 8 * s;
        mbar_expect_tx(full, STAGE_BYTES);
        for (int p = 0; p < NA; ++p) tma_load_2d(st + p * A_BYTES, &mapA, full, (int)(kb * BK), (int)(p * M + row0));
        for (int p = 0; p < NW; ++p) tma_load_2d(st + NA * A_BYTES + p * W_BYTES, &mapW, full, (int)(kb * BK), (int)(p * N + col0));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {                                                  // ---- MMA issuer
      constexpr uint32_t IUU = idesc(0, 0), IUS = idesc(0, 1), ISU = idesc(1, 0), ISS = idesc(1, 1);
      for (uint32_t kb = 0; kb < nkb; ++kb) {
        const uint32_t s = kb % STAGES, ph = (kb / STAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_base + s * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < BK / 32; ++ks) {
          const uint64_t a0 = smem_desc(st + ks * 32), a1 = smem_desc(st + A_BYTES + ks * 32), a2 = smem_desc(st + 2 * A_BYTES + ks * 32);
          const uint64_t b0 = smem_desc(st + NA * A_BYTES + ks * 32), b1 = smem_desc(st + NA * A_BYTES + W_BYTES + ks * 32);
          const uint32_t acc = (kb | ks) ? 1u : 0u;                   // the first product into an accumulator overwrites it
          mma_i8(tmem + 0 * BN, a0, b0, IUU, acc);                    // shift 0:  a0 w0
          mma_i8(tmem + 1 * BN, a1, b0, IUU, acc);                    // shift 8:  a1 w0 + a0 w1
          mma_i8(tmem + 1 * BN, a0, b1, IUS, 1u);
          mma_i8(tmem + 2 * BN, a2, b0, ISU, acc);                    // shift 16: a2 w0 + a1 w1
          mma_i8(tmem + 2 * BN, a1, b1, IUS, 1u);
          mma_i8(tmem + 3 * BN, a2, b1, ISS, acc);                    // shift 24: a2 w1
        }
        mma_commit(empty0 + 8 * s);                                   // the stage may be refilled once these MMAs have read it
      }
      mma_commit(accum);
    }
    __syncwarp();
  }

  // ---- epilogue: TMEM lane = output row, column = output column
  mbar_wait(accum, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int quarter = warp & 3, c = (warp >> 2) * 16;                 // a warp may only read the TMEM lanes 32 (warp % 4) .. +31
  const size_t grow = (size_t)row0 + quarter * 32 + lane;
  const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
  {
    uint32_t acc[4][16];
#pragma unroll
    for (int s = 0; s < 4; ++s) tmem_ld16(lane_base + s * BN + c, acc[s]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // A TMEM lane is an output ROW, so a thread's 32-byte results lie N * 32 bytes apart across the warp: written directly,
    // every store instruction touches 32 cache lines and the epilogue - not the MMAs - bounds the kernel.  The stage ring is
    // idle now (every MMA has completed): each warp transposes 4 columns at a time through its own padded slice of it
    // (row stride 144 B: conflict-free 16-byte accesses both ways) and writes 128-byte row segments, 4 lines per instruction.
    const uint32_t zb = smem_base + warp * (2 * EPI_BUF), ab = zb + EPI_BUF;
    const Fr one = Fr::one();
    const uint4 one_lo = make_uint4(one.v[0], one.v[1], one.v[2], one.v[3]), one_hi = make_uint4(one.v[4], one.v[5], one.v[6], one.v[7]);
    const size_t tile_row = (size_t)row0 + quarter * 32;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t smask = 0;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = 4 * q + jj;
        long long v = 0;
#pragma unroll
        for (int s = 3; s >= 0; --s) v = v * 256 + (long long)(int32_t)acc[s][j];   // |v| < 2^29 * 2^24 * 1.01
        const bool negative = v < 0;
        Fr r = to_mont_u64(negative ? (unsigned long long)(-v) : (unsigned long long)v);
        if (negative) r = neg(r);
        const uint32_t zs = zb + lane * EPI_ROW + jj * 32;
        sts128(zs, r.v[0], r.v[1], r.v[2], r.v[3]); sts128(zs + 16, r.v[4], r.v[5], r.v[6], r.v[7]);
        if (ro.act) {                                                  // zkReLU::operator() on the exact integer (zkrelu.cu:11-52)
          const ReluParts p = relu_decompose_i64(v);
          if (p.out_of_range && ro.bad) atomicAdd(ro.bad, 1u);
          const Fr a = p.positive ? to_mont_u32(p.q) : Fr::zero();
          const uint32_t as = ab + lane * EPI_ROW + jj * 32;
          sts128(as, a.v[0], a.v[1], a.v[2], a.v[3]); sts128(as + 16, a.v[4], a.v[5], a.v[6], a.v[7]);
          smask |= (p.positive ? 1u : 0u) << jj;
          acc[0][j] = p.q; acc[1][j] = p.r;                           // the accumulators are consumed: reuse them as staging
        }
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int id = it * 32 + lane, row = id >> 3, ch = id & 7;    // 8 x 16 B = one row's 4 elements
        const size_t g = ((tile_row + row) * N + col0 + c + 4 * q) * sizeof(Fr) + ch * 16;
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(C) + g) = lds128(zb + row * EPI_ROW + ch * 16);
        if (ro.act) {
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(ro.act) + g) = lds128(ab + row * EPI_ROW + ch * 16);
          const uint32_t m = __shfl_sync(0xffffffffu, smask, row);
          const bool pos = (m >> (ch >> 1)) & 1u;
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(ro.sign) + g) = pos ? ((ch & 1) ? one_hi : one_lo) : make_uint4(0, 0, 0, 0);
        }
      }
      __syncwarp();
    }
    if (ro.act) {                                                     // 16 consecutive packed values per thread: 64 + 32 aligned bytes
      const size_t idx = grow * N + col0 + c;
      uint4* q4 = reinterpret_cast<uint4*>(ro.mag + idx);
#pragma unroll
      for (int j = 0; j < 4; ++j) q4[j] = make_uint4(acc[0][4 * j], acc[0][4 * j + 1], acc[0][4 * j + 2], acc[0][4 * j + 3]);
      uint4* r4 = reinterpret_cast<uint4*>(ro.rem + idx);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        r4[j] = make_uint4(acc[1][8 * j] | (acc[1][8 * j + 1] << 16), acc[1][8 * j + 2] | (acc[1][8 * j + 3] << 16),
                           acc[1][8 * j + 4] | (acc[1][8 * j + 5] << 16), acc[1][8 * j + 6] | (acc[1][8 * j + 7] << 16));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// planes: [nplanes][rows][K] bytes, K contiguous -> 2-D map {K, nplanes * rows}, box {128, box_rows}, 128-byte swizzle
static bool make_map(CUtensorMap* map, const uint8_t* planes, size_t rows_total, size_t K, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows_total};
  cuuint64_t strides[1] = {(cuuint64_t)K};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace umma

bool umma_matmul_shape_ok(size_t M, size_t K, size_t N) {
  static const bool off = getenv("ZKDL_MM_NO_UMMA") != nullptr;       // A/B knob: fall back to the mma.sync kernel
  return !off && M % umma::BM == 0 && N % umma::BN == 0 && K % umma::BK == 0 && K <= 16384 && ((uint64_t)3 * M | (uint64_t)2 * N) < (1ull << 31);
}

// Ap: [3][M][K] byte planes of A, Wp: [2][N][K] byte planes of W^T (same buffers as k_tc_matmul).  Returns -1 if the tensor
// maps cannot be built (the caller then launches the mma.sync kernel).
int umma_matmul_launch(const uint8_t* Ap, const uint8_t* Wp, Fr* C, size_t M, size_t K, size_t N, const uint32_t* info, const ReluOut& ro, cudaStream_t st) {
  CUtensorMap mapA, mapW;
  if (!umma::make_map(&mapA, Ap, 3 * M, K, umma::BM) || !umma::make_map(&mapW, Wp, 2 * N, K, umma::BN)) return -1;
  static bool attr_set = false;
  if (!attr_set) {
    ZK_CUDA(cudaFuncSetAttribute(umma::k_umma_matmul, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)umma::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((unsigned)(N / umma::BN), (unsigned)(M / umma::BM));
  ZK_LAUNCH(umma::k_umma_matmul<<<grid, umma::NTHREADS, umma::SMEM_BYTES, st>>>(mapA, mapW, C, (uint32_t)M, (uint32_t)K, (uint32_t)N, info, ro));
  return ZK_OK;
}

}  // namespace zk
