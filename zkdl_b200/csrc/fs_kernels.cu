// fs_kernels.cu — the three sumchecks with a Fiat-Shamir transcript ON THE DEVICE.
//
// The reference draws every challenge from std::random_device before the sumcheck starts (random_vec, /root/reference/
// proof.cu:3-11; zkfc.cu:135-137, zkrelu.cu:85-98), so its proofs bind to nothing.  Here the fold challenge of round j is
//     S_{j+1} = SHA-256(S_j || c0 || c1 || c2),    x_j = limbs(S_{j+1}) with the top limb % 0x73eda753
// (the same "value < p, read as Montgomery form" recipe as random_vec), computed by the last CTA of round j's kernel right
// after it has summed that round's three coefficients.  No host round trip: the next launch reads x_j from device memory.
//
// A challenge that depends on the round's own coefficients cannot be used in the pass that computes them, so the rounds
// are shifted by half a step relative to fr_kernels.cu: launch j FOLDS table T_{j-1} with x_{j-1} on the fly (writing
// T_j) and evaluates round j's coefficients on the folded pairs - still one pass over HBM per round.  Same proof layout
// and the same field elements as inner_product / hadamard_product / binary_sumcheck (proof.cu:55-200) would produce for
// the challenges the transcript yields (tests/test_fiat_shamir_gpu.py checks exactly that against the injected-challenge
// kernels).  The eq point u of the Hadamard / binary sumchecks is fixed before round 0 and comes from the host.
#include "common.cuh"
#include "fr_device.cuh"
#include "../../include/zkdl_b200.h"

namespace zk {
int build_eq_table(const Fr* q_dev, const zkdl_fr_t* q_host, int t, int rev, Fr* E, cudaStream_t st);

namespace fs {
enum { FS_IP = 0, FS_HP = 1, FS_BIN = 2 };
static constexpr int THREADS = 256;

// ---- SHA-256 (FIPS 180-4), one thread
__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ void sha256_compress(uint32_t* st, const uint32_t* blk) {
  static const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74,
      0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d,
      0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e,
      0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5,
      0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t w[64];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = blk[i];
#pragma unroll
  for (int i = 16; i < 64; ++i) {
    uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 8
  for (int i = 0; i < 64; ++i) {
    uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
    uint32_t t1 = h + S1 + ch + K[i] + w[i];
    uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
__device__ __forceinline__ uint32_t bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
// state <- SHA-256(state bytes || c0 || c1 || c2)  (128 message bytes: 2 blocks + the padding block); returns the challenge.
// A field element is hashed as its 32 bytes in memory (8 little-endian u32 limbs, Montgomery form as stored in the proof).
__device__ Fr transcript_round(uint32_t* state, const Fr& c0, const Fr& c1, const Fr& c2) {
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  uint32_t blk[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) { blk[i] = state[i]; blk[8 + i] = bswap(c0.v[i]); }
  sha256_compress(h, blk);
#pragma unroll
  for (int i = 0; i < 8; ++i) { blk[i] = bswap(c1.v[i]); blk[8 + i] = bswap(c2.v[i]); }
  sha256_compress(h, blk);
#pragma unroll
  for (int i = 0; i < 16; ++i) blk[i] = 0;
  blk[0] = 0x80000000u; blk[15] = 128 * 8;
  sha256_compress(h, blk);
  Fr x;
#pragma unroll
  for (int i = 0; i < 8; ++i) { state[i] = h[i]; x.v[i] = bswap(h[i]); }      // limb i = little-endian u32 of digest bytes 4i..4i+3
  x.v[7] %= 1944954707u;
  return x;
}

// entry i of the CURRENT table: the previous table folded with the previous challenge (zero padding as fr-tensor.cu:404-408)
template <bool FIRST>
__device__ __forceinline__ Fr cur_entry(const Fr* __restrict__ prev, size_t prev_n, size_t i, const Fr& xp) {
  if (FIRST) return i < prev_n ? prev[i] : Fr::zero();
  const size_t i0 = 2 * i, i1 = 2 * i + 1;
  Fr p0 = i0 < prev_n ? prev[i0] : Fr::zero();
  Fr p1 = i1 < prev_n ? prev[i1] : Fr::zero();
  return fold_pair(p0, p1, xp);
}
__device__ __forceinline__ Fr ldcg(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldcg(q), b = __ldcg(q + 1);
  Fr r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// Round j.  prev_*: T_{j-1} (FIRST: T_0 itself), prev_n entries; cur_*: T_j is written here (not when FIRST); cur_n entries.
// e_in: eq(u[j+1:], .) indexed by the pairs of T_j (esize entries, HP / BIN); e_out: the next round's table (pair sums) or null.
template <int KIND, bool FIRST>
__global__ void __launch_bounds__(THREADS) k_fs_round(const Fr* __restrict__ prev_a, const Fr* __restrict__ prev_b, size_t prev_n, const Fr* __restrict__ x_prev,
                                                      Fr* __restrict__ cur_a, Fr* __restrict__ cur_b, size_t cur_n, const Fr* __restrict__ e_in,
                                                      Fr* __restrict__ e_out, size_t H, Fr* __restrict__ partials, unsigned* __restrict__ counter,
                                                      Fr* __restrict__ proof3, uint32_t* __restrict__ state, Fr* __restrict__ x_out) {
  __shared__ Fr sm[3 * 32];
  __shared__ bool is_last;
  const Fr xp = FIRST ? Fr::zero() : *x_prev;
  const size_t pairs = (cur_n + 1) / 2;
  Fr acc[3] = {Fr::zero(), Fr::zero(), Fr::zero()};
  for (size_t h = blockIdx.x * (size_t)blockDim.x + threadIdx.x; h < H; h += (size_t)gridDim.x * blockDim.x) {
    Fr e0 = Fr::zero(), e1 = Fr::zero();
    if (KIND != FS_IP) {
      e0 = e_in[2 * h];
      if (e_out) { e1 = e_in[2 * h + 1]; e_out[h] = add(e0, e1); }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const size_t g = 2 * h + t;
      if (g >= pairs) break;
      Fr a0 = cur_entry<FIRST>(prev_a, prev_n, 2 * g, xp), a1 = cur_entry<FIRST>(prev_a, prev_n, 2 * g + 1, xp);
      if (!FIRST) { cur_a[2 * g] = a0; if (2 * g + 1 < cur_n) cur_a[2 * g + 1] = a1; }
      const Fr& e = t ? e1 : e0;
      Fr da = sub(a1, a0), c0, c1, c2;
      if (KIND == FS_BIN) {                                  // proof.cu:152-163: a0^2 - a0, 2 a0 d - d, d^2, eq-weighted
        Fr ea0 = mul(e, a0), ed = mul(e, da), one = Fr::one();
        c0 = mul(ea0, sub(a0, one)); c1 = mul(ed, sub(dbl(a0), one)); c2 = mul(ed, da);
      } else {                                               // proof.cu:55-70: a0 b0, a0 (b1-b0) + b0 (a1-a0), (a1-a0)(b1-b0)
        Fr b0 = cur_entry<FIRST>(prev_b, prev_n, 2 * g, xp), b1 = cur_entry<FIRST>(prev_b, prev_n, 2 * g + 1, xp);
        if (!FIRST) { cur_b[2 * g] = b0; if (2 * g + 1 < cur_n) cur_b[2 * g + 1] = b1; }
        Fr db = sub(b1, b0);
        Fr wa0 = KIND == FS_HP ? mul(e, a0) : a0, wda = KIND == FS_HP ? mul(e, da) : da;
        c0 = mul(wa0, b0); c2 = mul(wda, db);
        c1 = sub(sub(mul(add(wa0, wda), b1), c0), c2);
      }
      acc[0] = add(acc[0], c0); acc[1] = add(acc[1], c1); acc[2] = add(acc[2], c2);
    }
  }
  block_reduce_fr<3>(acc, sm);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 3 + 0] = acc[0]; partials[blockIdx.x * 3 + 1] = acc[1]; partials[blockIdx.x * 3 + 2] = acc[2];
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  acc[0] = acc[1] = acc[2] = Fr::zero();
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    acc[0] = add(acc[0], ldcg(partials + (size_t)i * 3)); acc[1] = add(acc[1], ldcg(partials + (size_t)i * 3 + 1)); acc[2] = add(acc[2], ldcg(partials + (size_t)i * 3 + 2));
  }
  __syncthreads();
  block_reduce_fr<3>(acc, sm);
  if (threadIdx.x == 0) {
    proof3[0] = acc[0]; proof3[1] = acc[1]; proof3[2] = acc[2];
    *x_out = transcript_round(state, acc[0], acc[1], acc[2]);     // this round's challenge: binds to everything hashed so far
    *counter = 0;
  }
}
// a(0) (and b(0)) = the last table (<= 2 entries) folded with the last challenge
__global__ void k_fs_final(const Fr* __restrict__ a, const Fr* __restrict__ b, size_t n, const Fr* __restrict__ x_last, int nfin, Fr* __restrict__ out) {
  const Fr x = *x_last;
  out[0] = fold_pair(a[0], n > 1 ? a[1] : Fr::zero(), x);
  if (nfin > 1) out[1] = fold_pair(b[0], n > 1 ? b[1] : Fr::zero(), x);
}

static inline unsigned fs_grid(size_t items) {
  size_t blocks = (items + THREADS - 1) / THREADS, cap = (size_t)num_sms() * 8;
  return (unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

template <int KIND>
static int run(const Fr* a, const Fr* b, size_t n, const zkdl_fr_t* u_host, size_t k, const uint8_t* state_in_host, Fr* proof, Fr* xs, uint32_t* state,
               cudaStream_t st) {
  const int nfin = KIND == FS_BIN ? 1 : 2;
  // the transcript state travels as the 8 big-endian words of the 32 digest bytes
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) h[i] = ((uint32_t)state_in_host[4 * i] << 24) | ((uint32_t)state_in_host[4 * i + 1] << 16) | ((uint32_t)state_in_host[4 * i + 2] << 8) | state_in_host[4 * i + 3];
  ZK_CUDA(cudaMemcpyAsync(state, h, sizeof(h), cudaMemcpyHostToDevice, st));
  if (k == 0) {                                               // proof.cu:75-79 / 114-118 / 168-171
    ZK_CUDA(cudaMemcpyAsync(proof, a, sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    if (nfin > 1) ZK_CUDA(cudaMemcpyAsync(proof + 1, b, sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    return ZK_OK;
  }
  int rc;
  Scratch uq, E0, E1, A0, A1, B0, B1, parts, counter;
  size_t esize = 1;
  if (KIND != FS_IP) {
    esize = (size_t)1 << (k - 1);
    if ((rc = uq.alloc(sizeof(Fr) * k, st))) return rc;
    if (k > 1) ZK_CUDA(cudaMemcpyAsync(uq.p, u_host + 1, sizeof(Fr) * (k - 1), cudaMemcpyHostToDevice, st));
    if ((rc = E0.alloc(sizeof(Fr) * esize, st))) return rc;
    if ((rc = E1.alloc(sizeof(Fr) * (esize / 2 + 1), st))) return rc;
    if ((rc = build_eq_table(uq.as<Fr>(), u_host + 1, (int)k - 1, 0, E0.as<Fr>(), st))) return rc;
  }
  const size_t half = (n + 1) / 2;
  if ((rc = A0.alloc(sizeof(Fr) * half, st))) return rc;
  if ((rc = A1.alloc(sizeof(Fr) * half, st))) return rc;
  if (nfin > 1) { if ((rc = B0.alloc(sizeof(Fr) * half, st))) return rc; if ((rc = B1.alloc(sizeof(Fr) * half, st))) return rc; }
  size_t maxH = (half + 1) / 2; if (KIND != FS_IP && esize / 2 > maxH) maxH = esize / 2;
  if ((rc = parts.alloc(sizeof(Fr) * 3 * fs_grid(maxH), st))) return rc;
  if ((rc = counter.alloc(sizeof(unsigned), st))) return rc;
  ZK_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned), st));
  const Fr *pa = a, *pb = b; size_t prev_n = n;
  Fr *abuf[2] = {A0.as<Fr>(), A1.as<Fr>()}, *bbuf[2] = {B0.as<Fr>(), B1.as<Fr>()}, *ebuf[2] = {E0.as<Fr>(), E1.as<Fr>()};
  int which = 0, ewhich = 0;
  for (size_t j = 0; j < k; ++j) {
    const size_t cur_n = j ? (prev_n + 1) / 2 : prev_n, pairs = (cur_n + 1) / 2;
    const bool fold_e = KIND != FS_IP && esize >= 2;
    const size_t H = fold_e ? esize / 2 : (pairs + 1) / 2;
    const unsigned grid = fs_grid(H);
    if (j == 0)
      ZK_LAUNCH(k_fs_round<KIND, true><<<grid, THREADS, 0, st>>>(pa, pb, prev_n, nullptr, nullptr, nullptr, cur_n, ebuf[ewhich], fold_e ? ebuf[ewhich ^ 1] : nullptr, H,
                                                                 parts.as<Fr>(), counter.as<unsigned>(), proof + 3 * j, state, xs + j));
    else {
      ZK_LAUNCH(k_fs_round<KIND, false><<<grid, THREADS, 0, st>>>(pa, pb, prev_n, xs + j - 1, abuf[which], bbuf[which], cur_n, ebuf[ewhich],
                                                                  fold_e ? ebuf[ewhich ^ 1] : nullptr, H, parts.as<Fr>(), counter.as<unsigned>(), proof + 3 * j, state, xs + j));
      pa = abuf[which]; pb = bbuf[which]; which ^= 1; prev_n = cur_n;
    }
    if (fold_e) { ewhich ^= 1; esize /= 2; }
  }
  ZK_LAUNCH(k_fs_final<<<1, 1, 0, st>>>(pa, pb, prev_n, xs + k - 1, nfin, proof + 3 * k));
  return ZK_OK;
}

}  // namespace fs
}  // namespace zk

using namespace zk;

extern "C" {
int zkdl_sumcheck_fs(int kind, const zkdl_fr_t* a, const zkdl_fr_t* b, size_t n, const zkdl_fr_t* u_host, size_t k, const uint8_t* state_in_host,
                     zkdl_fr_t* proof, zkdl_fr_t* challenges, uint32_t* state_out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ZK_REQUIRE(a && state_in_host && proof && state_out && (k == 0 || challenges), ZK_ERR_ARG, "null argument");
  ZK_REQUIRE(kind >= 0 && kind <= 2 && (kind == ZKDL_FS_BIN || b) && (kind == ZKDL_FS_IP || k == 0 || u_host), ZK_ERR_ARG, "bad arguments");
  ZK_REQUIRE(k < 32 && !(n <= (((size_t)1 << k) / 2) || n > ((size_t)1 << k)), ZK_ERR_DIM, "Incompatible dimensions");      // proof.cu:102-104, 143-147, 193-196
  const Fr* A = reinterpret_cast<const Fr*>(a); const Fr* B = reinterpret_cast<const Fr*>(b);
  Fr* P = reinterpret_cast<Fr*>(proof); Fr* X = reinterpret_cast<Fr*>(challenges);
  if (kind == ZKDL_FS_IP) return fs::run<fs::FS_IP>(A, B, n, u_host, k, state_in_host, P, X, state_out, st);
  if (kind == ZKDL_FS_HP) return fs::run<fs::FS_HP>(A, B, n, u_host, k, state_in_host, P, X, state_out, st);
  return fs::run<fs::FS_BIN>(A, nullptr, n, u_host, k, state_in_host, P, X, state_out, st);
}
}
