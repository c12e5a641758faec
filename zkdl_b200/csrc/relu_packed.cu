// relu_packed.cu — zkReLU proof on the bit-packed auxiliary input.
//
// The reference stores the ReLU decomposition as 0/1 Fr tables (mag_bin: 32 cells, rem_bin: 16 cells per activation,
// 1.5 KB per element, /root/reference/zkrelu.cu:30-38) and runs the generic binary sumcheck over them
// (/root/reference/zkrelu.cu:91-94 -> proof.cu:152-200): the largest HBM consumer of the whole proof.
// The cells are exactly Scalar_ONE / Scalar_ZERO, so the same proof elements can be produced from 48 bits per element:
//   * rounds 0..2 of the binary sumcheck only ever see tables whose entries are functions of 2 / 4 / 8 original bits.
//     With eq(u[j+1:], g) = eq_lo(j, group-in-element) * eq_hi(element), round j's three coefficients are
//     sum_elem eq_hi[elem] * sum_groups LUT_j[group][bit pattern], i.e. table look-ups + additions and 5 products per
//     ELEMENT for all three rounds together (only c1 and c2 are summed: c0 follows from the running claim, which starts
//     at 0 for 0/1 cells - k_bin_packed_finish), instead of 6 products per CELL PAIR per round.
//   * after three rounds each table entry is V3[byte]: round 3 pairs two of them, i.e. its coefficients are a function
//     of 16 bits (a 65536-entry look-up table in L2, pre-weighted by the in-element eq factor), and the table after
//     round 3 is V4[16-bit half].  For mag_bin (32 cells) round 4 pairs the two halves of an element: 2 + 4 products per
//     ELEMENT for rounds 3 and 4.  k_bin_r34 does these rounds straight from the packed words and writes the folded table with ONE entry
//     per element (a^(5) for mag_bin, a^(4) for rem_bin); only then do the generic single-pass rounds (fr_kernels.cu)
//     take over.  The tables a^(3), a^(4) (8 and 4 entries per element) are never materialised.
//   * mag_bin.partial_me(u, 32) / rem_bin.partial_me(u, 16) (zkrelu.cu:92,94) are per-bit sums of eq(u, elem).
// All values are the same field elements the reference computes, so the proof is bit-identical (tests/test_gpu_parity.py).
#include <atomic>
#include "common.cuh"
#include "fr_device.cuh"
#include "../../include/zkdl_b200.h"

namespace zk {

int build_eq_table(const Fr* q_dev, const zkdl_fr_t* q_host, int t, int rev, Fr* E, cudaStream_t st);
int bin_sumcheck_continue(const Fr* a, size_t n, const zkdl_fr_t* u_host, const zkdl_fr_t* v_host, size_t k, const Fr* claim_dev, Fr* proof, cudaStream_t st);

template <int Q> struct PackLayout {
  static constexpr int LOGQ = Q == 32 ? 5 : 4;
  static constexpr int N0 = Q / 2, N1 = Q / 4, N2 = Q / 8;
  static constexpr int OFF_E0 = 0;                         // eq_lo of round 0: N0 entries
  static constexpr int OFF_L1 = OFF_E0 + N0;               // [N1][16][3]
  static constexpr int OFF_L2 = OFF_L1 + N1 * 16 * 3;      // [N2][256][3]
  static constexpr int OFF_V3 = OFF_L2 + N2 * 256 * 3;     // [256]
  static constexpr int TOTAL = OFF_V3 + 256;               // shared-memory part (rounds 0..2)
  // L2-resident part (rounds 3..4), after the shared part in the same buffer
  static constexpr int NH = Q / 16;                         // 16-bit halves per element
  static constexpr int OFF_L3 = TOTAL;                      // [NH][65536][3], pre-weighted by eq(u[4:LOGQ], half)
  static constexpr int OFF_V4 = OFF_L3 + NH * 65536 * 3;    // [65536]
  static constexpr int TOTAL_ALL = OFF_V4 + 65536;
  static constexpr int ROUNDS = LOGQ;                       // packed rounds: all in-element variables
};

// unweighted Fr_bin_sc_step coefficients of a pair (proof.cu:152-163)
__device__ __forceinline__ void bin_coeffs(const Fr& x0, const Fr& x1, Fr* c) {
  Fr d = sub(x1, x0);
  c[0] = sub(mul(x0, x0), x0);
  c[1] = sub(mul(dbl(x0), d), d);
  c[2] = mul(d, d);
}
// eq(q[0..t), idx): q[0] binds bit 0
__device__ __forceinline__ Fr eq_point(const Fr* q, int t, unsigned idx) {
  Fr r = Fr::one();
  for (int i = 0; i < t; ++i) r = mul(r, ((idx >> i) & 1u) ? q[i] : sub(Fr::one(), q[i]));
  return r;
}

// Builds the look-up tables of the packed rounds from the challenges.  CTA 0: the shared-memory tables of rounds 0..2;
// CTAs 1..256: 256 entries each of the round-3 tables (every CTA recomputes the 4 + 16 + 256 folded cell values).
template <int Q>
__global__ void __launch_bounds__(256) k_bin_luts(const Fr* __restrict__ u, const Fr* __restrict__ v, Fr* __restrict__ lut) {
  using PL = PackLayout<Q>;
  __shared__ Fr V1[4], V2[16], V3[256], e1[PL::N1], e2[PL::N2], e3[PL::NH];
  const int tid = threadIdx.x;
  if (tid < 4) {
    Fr x0 = (tid & 1) ? Fr::one() : Fr::zero(), x1 = (tid >> 1) ? Fr::one() : Fr::zero();
    V1[tid] = fold_pair(x0, x1, v[0]);
  }
  if (blockIdx.x == 0) {
    if (tid < PL::N0) lut[PL::OFF_E0 + tid] = eq_point(u + 1, PL::LOGQ - 1, tid);
    if (tid < PL::N1) e1[tid] = eq_point(u + 2, PL::LOGQ - 2, tid);
    if (tid < PL::N2) e2[tid] = eq_point(u + 3, PL::LOGQ - 3, tid);
  } else if (tid < PL::NH) e3[tid] = eq_point(u + 4, PL::LOGQ - 4, tid);
  __syncthreads();
  if (tid < 16) V2[tid] = fold_pair(V1[tid & 3], V1[tid >> 2], v[1]);
  __syncthreads();
  {                                                       // V3[byte] = fold of (V2[lo nibble], V2[hi nibble]) with v[2]
    Fr x0 = V2[tid & 15], x1 = V2[tid >> 4];
    V3[tid] = fold_pair(x0, x1, v[2]);
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    // round 1 tables: pattern = nibble -> pair (V1[lo 2 bits], V1[hi 2 bits])
    for (int i = tid; i < PL::N1 * 16; i += blockDim.x) {
      int t = i / 16, nib = i % 16; Fr c[3];
      bin_coeffs(V1[nib & 3], V1[nib >> 2], c);
      for (int k = 0; k < 3; ++k) lut[PL::OFF_L1 + i * 3 + k] = mul(e1[t], c[k]);
    }
    // round 2 tables: pattern = byte -> pair (V2[lo nibble], V2[hi nibble])
    Fr c[3];
    bin_coeffs(V2[tid & 15], V2[tid >> 4], c);
    for (int t = 0; t < PL::N2; ++t)
      for (int k = 0; k < 3; ++k) lut[PL::OFF_L2 + (t * 256 + tid) * 3 + k] = mul(e2[t], c[k]);
    lut[PL::OFF_V3 + tid] = V3[tid];
    return;
  }
  // round 3: pattern = 16 bits -> pair (V3[lo byte], V3[hi byte]); V4[pattern] = fold of that pair with v[3]
  const int h = (blockIdx.x - 1) * 256 + tid;
  Fr x0 = V3[h & 255], x1 = V3[h >> 8], c[3];
  bin_coeffs(x0, x1, c);
  for (int t = 0; t < PL::NH; ++t)
    for (int k = 0; k < 3; ++k) lut[PL::OFF_L3 + ((size_t)t * 65536 + h) * 3 + k] = PL::NH == 1 ? c[k] : mul(e3[t], c[k]);
  lut[PL::OFF_V4 + h] = fold_pair(x0, x1, v[3]);
}

extern __shared__ __align__(16) unsigned char pk_smem[];

// One pass over the packed words: rounds 0, 1, 2 of the binary sumcheck.
// partials[blockIdx][7] = {S0, S1[3], S2[3]}
template <int Q, class T>
__global__ void __launch_bounds__(512) k_bin_packed3(const T* __restrict__ packed, size_t n, const Fr* __restrict__ e_hi, const Fr* __restrict__ lut,
                                                     Fr* __restrict__ partials) {
  using PL = PackLayout<Q>;
  Fr* sm = reinterpret_cast<Fr*>(pk_smem);
  for (int i = threadIdx.x; i < PL::OFF_V3; i += blockDim.x) sm[i] = lut[i];
  __syncthreads();
  const Fr* E0 = sm + PL::OFF_E0; const Fr* L1 = sm + PL::OFF_L1; const Fr* L2 = sm + PL::OFF_L2;
  Fr acc[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) acc[k] = Fr::zero();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t w = packed[i];
    Fr eh = e_hi[i];
    Fr s0 = Fr::zero();
    uint32_t diff = (w ^ (w >> 1));                       // bit 2t set <=> cells 2t, 2t+1 differ
#pragma unroll 4
    for (int t = 0; t < PL::N0; ++t) if ((diff >> (2 * t)) & 1u) s0 = add(s0, E0[t]);
    // only c1 and c2 of rounds 1 and 2 are summed: c0 follows from the running claim (k_bin_packed_finish), see bin_pair_c12
    Fr s1[3] = {Fr::zero(), Fr::zero(), Fr::zero()}, s2[3] = {Fr::zero(), Fr::zero(), Fr::zero()};
#pragma unroll 2
    for (int t = 0; t < PL::N1; ++t) {
      const Fr* e = L1 + (t * 16 + ((w >> (4 * t)) & 15u)) * 3;
      s1[1] = add(s1[1], e[1]); s1[2] = add(s1[2], e[2]);
    }
#pragma unroll
    for (int t = 0; t < PL::N2; ++t) {
      uint32_t byte = (w >> (8 * t)) & 255u;
      const Fr* e = L2 + (t * 256 + byte) * 3;
      s2[1] = add(s2[1], e[1]); s2[2] = add(s2[2], e[2]);
    }
    acc[0] = add(acc[0], mul(eh, s0));
#pragma unroll
    for (int k = 1; k < 3; ++k) { acc[1 + k] = add(acc[1 + k], mul(eh, s1[k])); acc[4 + k] = add(acc[4 + k], mul(eh, s2[k])); }
  }
  __shared__ Fr red[3 * 32];
  block_reduce_fr<3>(acc, red);
  __syncthreads();
  block_reduce_fr<3>(acc + 3, red);
  __syncthreads();
  block_reduce_fr<1>(acc + 6, red);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 7; ++k) partials[blockIdx.x * 7 + k] = acc[k];
}
// Rounds 3 (and 4 for 32-cell elements) from the packed words + the folded table with one entry per element.
// partials[blockIdx][3 * (ROUNDS - 3)] = {S3[3] (, S4[3])}
__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fr r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <int Q, class T>
__global__ void __launch_bounds__(256, 2) k_bin_r34(const T* __restrict__ packed, size_t n, const Fr* __restrict__ e_hi, const Fr* __restrict__ lut, Fr v4,
                                                 Fr* __restrict__ out, Fr* __restrict__ partials) {
  using PL = PackLayout<Q>;
  constexpr int NACC = 3 * (PL::ROUNDS - 3);
  const Fr* L3 = lut + PL::OFF_L3; const Fr* V4 = lut + PL::OFF_V4;
  Fr acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = Fr::zero();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t w = packed[i];
    Fr eh = e_hi[i];
    const Fr* e = L3 + (size_t)(w & 0xffffu) * 3;
    Fr s3[3] = {Fr::zero(), ldg_fr(e + 1), ldg_fr(e + 2)};             // c1, c2 only (c0 is derived from the running claim)
    Fr x0 = ldg_fr(V4 + (w & 0xffffu));
    if (Q == 32) {
      const Fr* e1 = L3 + ((size_t)65536 + (w >> 16)) * 3;
      s3[1] = add(s3[1], ldg_fr(e1 + 1)); s3[2] = add(s3[2], ldg_fr(e1 + 2));
      Fr x1 = ldg_fr(V4 + (w >> 16)), c[3];
      out[i] = bin_pair_c12(x0, x1, eh, v4, c);            // round 4: the element's two halves
#pragma unroll
      for (int k = 1; k < 3; ++k) acc[3 + k] = add(acc[3 + k], c[k]);
    } else {
      out[i] = x0;
    }
#pragma unroll
    for (int k = 1; k < 3; ++k) acc[k] = add(acc[k], mul(eh, s3[k]));
  }
  __shared__ Fr red[3 * 32];
  block_reduce_fr<3>(acc, red);
  if (NACC > 3) { __syncthreads(); block_reduce_fr<3>(acc + (NACC > 3 ? 3 : 0), red); }
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < NACC; ++k) partials[blockIdx.x * NACC + k] = acc[k];
}
// proof[0 .. 3 * rounds) from the per-CTA partials.  The cells are exactly 0 / 1, so the sumcheck's claim starts at 0 and
// round 0 is (0, -S0, S0); for every later packed round only c1 and c2 were summed and c0 = claim_j - u_j (c1 + c2) with
// claim_{j+1} = c0 + v_j (c1 + v_j c2)  (the verifier's own identities, exact in the field).  claim_out = the claim the generic
// rounds continue from.
__global__ void __launch_bounds__(256) k_bin_packed_finish(const Fr* __restrict__ partials, unsigned nparts, const Fr* __restrict__ partials2, unsigned nparts2,
                                                           int nacc2, const Fr* __restrict__ u, const Fr* __restrict__ v, Fr* __restrict__ proof,
                                                           Fr* __restrict__ claim_out) {
  __shared__ Fr red[3 * 32];
  __shared__ Fr sums[13];
  Fr acc[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) acc[k] = Fr::zero();
  for (unsigned i = threadIdx.x; i < nparts; i += blockDim.x)
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = add(acc[k], partials[(size_t)i * 7 + k]);
  block_reduce_fr<3>(acc, red);
  __syncthreads();
  block_reduce_fr<3>(acc + 3, red);
  __syncthreads();
  block_reduce_fr<1>(acc + 6, red);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 7; ++k) sums[k] = acc[k];
  __syncthreads();
  Fr b[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) b[k] = Fr::zero();
  for (unsigned i = threadIdx.x; i < nparts2; i += blockDim.x)
    for (int k = 0; k < nacc2; ++k) b[k] = add(b[k], partials2[(size_t)i * nacc2 + k]);
  block_reduce_fr<3>(b, red);
  __syncthreads();
  block_reduce_fr<3>(b + 3, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) sums[7 + k] = b[k];
    // sums: [0] S0 | [1..3] round 1 | [4..6] round 2 | [7..9] round 3 | [10..12] round 4   (c0 slots unused)
    const int rounds = 3 + nacc2 / 3;
    proof[0] = Fr::zero(); proof[1] = neg(sums[0]); proof[2] = sums[0];
    Fr claim = add(proof[0], mul(v[0], add(proof[1], mul(v[0], proof[2]))));
    for (int r = 1; r < rounds; ++r) {
      const Fr c1 = sums[3 * r - 1], c2 = sums[3 * r];
      const Fr c0 = sub(claim, mul(u[r], add(c1, c2)));
      proof[3 * r] = c0; proof[3 * r + 1] = c1; proof[3 * r + 2] = c2;
      claim = add(c0, mul(v[r], add(c1, mul(v[r], c2))));
    }
    *claim_out = claim;
  }
}

// partial_me(u, Q) of a 0/1 table: out[bit] = sum over elements with that bit set of eq(u, elem).  Lane = bit.
template <int Q, class T>
__global__ void __launch_bounds__(256) k_packed_recover(const T* __restrict__ packed, size_t n, const Fr* __restrict__ e, Fr* __restrict__ partials) {
  constexpr int EPW = 32 / Q;                               // elements per warp iteration
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int bit = lane % Q, sub = lane / Q;
  Fr acc = Fr::zero();
  size_t wid = (size_t)blockIdx.x * nwarps + warp, nw = (size_t)gridDim.x * nwarps;
  constexpr int UN = 4;                                     // independent loads in flight per lane
  for (size_t base = wid * EPW * UN; base < n; base += nw * EPW * UN) {
    uint32_t w[UN]; Fr ev[UN];
#pragma unroll
    for (int r = 0; r < UN; ++r) {
      size_t i = base + r * EPW + sub;
      w[r] = i < n ? (uint32_t)packed[i] : 0u;
      ev[r] = i < n ? e[i] : Fr::zero();
    }
#pragma unroll
    for (int r = 0; r < UN; ++r) if ((w[r] >> bit) & 1u) acc = add(acc, ev[r]);
  }
  __shared__ Fr sm[8][32];
  sm[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    Fr s = sm[0][lane];
    for (int w2 = 1; w2 < nwarps; ++w2) s = add(s, sm[w2][lane]);
    if (EPW == 2) {                                         // combine the two sub-elements of a 16-bit table
      Fr o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = __shfl_down_sync(0xffffffffu, s.v[k], 16);
      s = add(s, o);
    }
    if (lane < Q) partials[(size_t)blockIdx.x * Q + lane] = s;
  }
}
// out[b] = sum_g partials[g][b]; 8 row groups per column, combined through shared memory
__global__ void __launch_bounds__(256) k_colsum(const Fr* __restrict__ partials, unsigned nparts, int Q, Fr* __restrict__ out) {
  __shared__ Fr sm[8][32];
  const int b = threadIdx.x & 31, grp = threadIdx.x >> 5;
  Fr s = Fr::zero();
  if (b < Q)
    for (unsigned g = grp; g < nparts; g += 8) s = add(s, partials[(size_t)g * Q + b]);
  sm[grp][b] = s;
  __syncthreads();
  if (grp == 0 && b < Q) {
    for (int w2 = 1; w2 < 8; ++w2) s = add(s, sm[w2][b]);
    out[b] = s;
  }
}

template <int Q, class T>
static int packed_bin_and_recover(const T* packed, size_t n, size_t L, const zkdl_fr_t* u_host, const zkdl_fr_t* v_host, const Fr* erec,
                                  Fr* proof_sc, Fr* proof_rec, cudaStream_t st) {
  using PL = PackLayout<Q>;
  const size_t k = L + PL::LOGQ;
  int rc;
  Scratch ud, vd, lut, ehi, a3, parts, rparts;
  if ((rc = ud.alloc(sizeof(Fr) * k, st))) return rc;
  if ((rc = vd.alloc(sizeof(Fr) * k, st))) return rc;
  ZK_CUDA(cudaMemcpyAsync(ud.p, u_host, sizeof(Fr) * k, cudaMemcpyHostToDevice, st));
  ZK_CUDA(cudaMemcpyAsync(vd.p, v_host, sizeof(Fr) * k, cudaMemcpyHostToDevice, st));
  if ((rc = lut.alloc(sizeof(Fr) * PL::TOTAL_ALL, st))) return rc;
  ZK_LAUNCH(k_bin_luts<Q><<<257, 256, 0, st>>>(ud.as<Fr>(), vd.as<Fr>(), lut.as<Fr>()));
  if ((rc = ehi.alloc(sizeof(Fr) * n, st))) return rc;
  if ((rc = build_eq_table(ud.as<Fr>() + PL::LOGQ, u_host + PL::LOGQ, (int)L, 0, ehi.as<Fr>(), st))) return rc;
  if ((rc = a3.alloc(sizeof(Fr) * n, st))) return rc;      // the folded table after the packed rounds: one entry per element
  unsigned grid = (unsigned)num_sms();                     // one CTA per SM (the look-up tables take 56-111 KB of shared memory)
  if ((size_t)grid * 512 > n) grid = div_up(n, 512);
  if ((rc = parts.alloc(sizeof(Fr) * 7 * grid, st))) return rc;
  size_t smem = sizeof(Fr) * PL::OFF_V3;
  ZK_CUDA(cudaFuncSetAttribute(k_bin_packed3<Q, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // rounds 0..2 over Q n cells of 32 B by SURVEY.md §8d's Fr-cell model: 96 Q n (1 - 1/8) B; real traffic (sizeof(T) + 32) n B; 5 products per element
  ZK_LAUNCH_P(st, 96.0 * Q * n * 0.875, 5.0 * n, 0.0, k_bin_packed3<Q, T><<<grid, 512, smem, st>>>(packed, n, ehi.as<Fr>(), lut.as<Fr>(), parts.as<Fr>()));
  constexpr int NACC = 3 * (PL::ROUNDS - 3);
  unsigned grid2 = (unsigned)num_sms() * 4;
  if ((size_t)grid2 * 256 > n) grid2 = div_up(n, 256);
  Scratch parts2;
  if ((rc = parts2.alloc(sizeof(Fr) * NACC * grid2, st))) return rc;
  Fr v4; for (int i = 0; i < 8; ++i) v4.v[i] = v_host[4].val[i];
  // rounds 3 (and 4): tables of Q n / 8 (and Q n / 16) cells in the Fr-cell model; 2 (+ 4) products per element
  ZK_LAUNCH_P(st, 48.0 * (Q * n / 8) * (Q == 32 ? 1.5 : 1.0), (Q == 32 ? 6.0 : 2.0) * n, 0.0,
              k_bin_r34<Q, T><<<grid2, 256, 0, st>>>(packed, n, ehi.as<Fr>(), lut.as<Fr>(), v4, a3.as<Fr>(), parts2.as<Fr>()));
  Scratch claim;
  if ((rc = claim.alloc(sizeof(Fr), st))) return rc;
  ZK_LAUNCH(k_bin_packed_finish<<<1, 256, 0, st>>>(parts.as<Fr>(), grid, parts2.as<Fr>(), grid2, NACC, ud.as<Fr>(), vd.as<Fr>(), proof_sc, claim.as<Fr>()));
  // remaining rounds on the folded table: binary_sumcheck(a, u[R:], v[R:]) has exactly those rounds and the final a(0); they
  // continue from the running claim, so they too sum only c1 and c2
  constexpr int R = PL::ROUNDS;
  if ((rc = bin_sumcheck_continue(a3.as<Fr>(), n, u_host + R, v_host + R, k - R, claim.as<Fr>(), proof_sc + 3 * R, st))) return rc;
  // partial_me(u_recover, Q)
  unsigned rgrid = (unsigned)num_sms() * 2;
  if ((size_t)rgrid * 8 * 4 > n) rgrid = div_up(n, 32);
  if ((rc = rparts.alloc(sizeof(Fr) * Q * rgrid, st))) return rc;
  ZK_LAUNCH(k_packed_recover<Q, T><<<rgrid, 256, 0, st>>>(packed, n, erec, rparts.as<Fr>()));
  ZK_LAUNCH(k_colsum<<<1, 256, 0, st>>>(rparts.as<Fr>(), rgrid, Q, proof_rec));
  return ZK_OK;
}

}  // namespace zk

using namespace zk;

extern "C" {

int zkdl_zkrelu_prove_packed(const zkdl_fr_t* X, const zkdl_fr_t* sign, const uint32_t* mag_packed, const uint16_t* rem_packed, size_t n,
                             const zkdl_fr_t* u_z_host, const zkdl_fr_t* v_z_host, const zkdl_fr_t* u_r_host, const zkdl_fr_t* v_r_host,
                             const zkdl_fr_t* u_rec_host, const zkdl_fr_t* u_hp_host, const zkdl_fr_t* v_hp_host,
                             zkdl_fr_t* proof_fr, void* stream) {
  return zkdl_zkrelu_prove_packed_parts(X, sign, mag_packed, rem_packed, n, u_z_host, v_z_host, u_r_host, v_r_host, u_rec_host,
                                        u_hp_host, v_hp_host, proof_fr, ZKDL_RELU_MAG | ZKDL_RELU_REM | ZKDL_RELU_HP, stream);
}

int zkdl_zkrelu_prove_packed_parts(const zkdl_fr_t* X, const zkdl_fr_t* sign, const uint32_t* mag_packed, const uint16_t* rem_packed, size_t n,
                                   const zkdl_fr_t* u_z_host, const zkdl_fr_t* v_z_host, const zkdl_fr_t* u_r_host, const zkdl_fr_t* v_r_host,
                                   const zkdl_fr_t* u_rec_host, const zkdl_fr_t* u_hp_host, const zkdl_fr_t* v_hp_host,
                                   zkdl_fr_t* proof_fr, unsigned parts, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ZK_REQUIRE(X && sign && mag_packed && rem_packed && proof_fr, ZK_ERR_ARG, "null argument");
  ZK_REQUIRE(parts && !(parts & ~(ZKDL_RELU_MAG | ZKDL_RELU_REM | ZKDL_RELU_HP)), ZK_ERR_ARG, "bad parts mask");
  size_t L = 0; while (((size_t)1 << L) < n) ++L;
  ZK_REQUIRE(((size_t)1 << L) == n && L >= 1 && L < 27, ZK_ERR_DIM, "Incompatible dimensions");
  int rc;
  Scratch urd, erec;
  if (parts & (ZKDL_RELU_MAG | ZKDL_RELU_REM)) {
    if ((rc = urd.alloc(sizeof(Fr) * L, st))) return rc;
    ZK_CUDA(cudaMemcpyAsync(urd.p, u_rec_host, sizeof(Fr) * L, cudaMemcpyHostToDevice, st));
    if ((rc = erec.alloc(sizeof(Fr) * n, st))) return rc;
    if ((rc = build_eq_table(urd.as<Fr>(), u_rec_host, (int)L, 0, erec.as<Fr>(), st))) return rc;
  }
  Fr* p = reinterpret_cast<Fr*>(proof_fr);
  Fr* p_mag_sc = p;                       p += 3 * (L + 5) + 1;
  Fr* p_mag_rec = p;                      p += 32;
  Fr* p_rem_sc = p;                       p += 3 * (L + 4) + 1;
  Fr* p_rem_rec = p;                      p += 16;
  // the three sumchecks are independent: mag_bin on the caller's stream, rem_bin and the Hadamard product on side streams
  SideStream& s1 = side_stream(1, st); SideStream& s2 = side_stream(2, st);
  ForkScope f1(s1, st), f2(s2, st);
  if ((rc = f1.fork())) return rc;
  if ((rc = f2.fork())) return rc;
  if ((parts & ZKDL_RELU_MAG) &&
      (rc = packed_bin_and_recover<32, uint32_t>(mag_packed, n, L, u_z_host, v_z_host, erec.as<Fr>(), p_mag_sc, p_mag_rec, st))) return rc;
  if ((parts & ZKDL_RELU_REM) &&
      (rc = packed_bin_and_recover<16, uint16_t>(rem_packed, n, L, u_r_host, v_r_host, erec.as<Fr>(), p_rem_sc, p_rem_rec, s1.stream))) return rc;
  if ((parts & ZKDL_RELU_HP) &&
      (rc = zkdl_hp_sumcheck(X, sign, n, u_hp_host, v_hp_host, L, reinterpret_cast<zkdl_fr_t*>(p), reinterpret_cast<void*>(s2.stream)))) return rc;   // zkrelu.cu:99
  if ((rc = f1.join())) return rc;
  return f2.join();
}

}  // extern "C"
