// field.cuh — BLS12-381 Fr (8x32-bit limbs) and Fq (12x32-bit limbs) Montgomery arithmetic for sm_100a.
//
// Replaces the reference's blstrs__scalar__Scalar_* / blstrs__fp__Fp_* device functions
// (/root/reference/bls12-381.cu:213-608 and :612-1015).  Same value semantics: Montgomery form with
// R = 2^256 (Fr) / 2^384 (Fq), every result canonical (< p), so results are bit-comparable with the reference.
//
// Multiplication is an interleaved (CIOS) Montgomery product on two register accumulators, one aligned to even
// limb positions and one to odd positions, so that every 32x32->64 partial product lands on an aligned register
// pair and ptxas fuses each mad.lo.cc/madc.hi.cc pair into one IMAD.WIDE.U32(.X).  2*N^2 wide multiplies per
// product + 2N for the reduction factors: 136 (Fr) / 300 (Fq) IMAD-class instructions.
//
// The limb primitives below have a host emulation (carry flag in a thread-local) so that the exact same template
// code is unit-tested on the CPU against the oracle (tests/test_field_host.py) before any GPU time is spent.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_D __device__ __forceinline__
#else
#define ZK_HD inline
#define ZK_D inline
#endif

namespace zk {

// ---------------------------------------------------------------- carry-chain primitives
#if defined(__CUDA_ARCH__)
ZK_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
ZK_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
ZK_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// host emulation of the PTX condition-code register (tests only)
static thread_local uint32_t zk_cc = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; zk_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + zk_cc; zk_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + zk_cc; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; zk_cc = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - zk_cc; zk_cc = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - zk_cc; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(a * b) + c; zk_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(a * b) + c + zk_cc; zk_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + zk_cc; zk_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)(((uint64_t)a * b) >> 32) + c + zk_cc; }
#endif

// ---------------------------------------------------------------- field parameters
// Montgomery products are emitted as real calls (not inlined) where a kernel chains many of them: a G1 mixed add is
// 10 Fq products = ~60 KB of straight-line SASS when inlined, far beyond the 32 KB L1.5 instruction cache, and the
// first profile showed the MSM kernels instruction-fetch bound (profiles/r1_v1_layer_launches.csv).  Arguments travel
// in registers (by value), so a call costs ~36 MOVs against ~370 arithmetic instructions.
#ifndef ZK_FR_OUTLINE_MUL
#define ZK_FR_OUTLINE_MUL 0
#endif
#ifndef ZK_FQ_OUTLINE_MUL
#define ZK_FQ_OUTLINE_MUL 1
#endif
struct FrParams {
  static constexpr bool OUTLINE_MUL = ZK_FR_OUTLINE_MUL != 0;
  static constexpr int N = 8;
  static constexpr uint32_t INV = 0xffffffffu;               // -p^-1 mod 2^32  (bls12-381.cuh:119)
  ZK_HD static uint32_t P(int i) {
    constexpr uint32_t p[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    return p[i];
  }
  ZK_HD static uint32_t ONE(int i) {                          // R mod p   (bls12-381.cu:3)
    constexpr uint32_t v[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
    return v[i];
  }
  ZK_HD static uint32_t R2(int i) {                           // R^2 mod p (bls12-381.cu:5)
    constexpr uint32_t v[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
    return v[i];
  }
};
struct FqParams {
  static constexpr bool OUTLINE_MUL = ZK_FQ_OUTLINE_MUL != 0;
  static constexpr int N = 12;
  static constexpr uint32_t INV = 0xfffcfffdu;               // bls12-381.cuh:221
  ZK_HD static uint32_t P(int i) {
    constexpr uint32_t p[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return p[i];
  }
  ZK_HD static uint32_t ONE(int i) {                          // bls12-381.cu:8
    constexpr uint32_t v[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
    return v[i];
  }
  ZK_HD static uint32_t R2(int i) {                           // bls12-381.cu:10
    constexpr uint32_t v[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
    return v[i];
  }
};

// ---------------------------------------------------------------- field element
template <class PR>
struct alignas(16) Fe {
  static constexpr int N = PR::N;
  uint32_t v[N];

  ZK_HD static Fe zero() { Fe r; _Pragma("unroll") for (int i = 0; i < N; ++i) r.v[i] = 0; return r; }
  ZK_HD static Fe one() { Fe r; _Pragma("unroll") for (int i = 0; i < N; ++i) r.v[i] = PR::ONE(i); return r; }
  ZK_HD static Fe r2() { Fe r; _Pragma("unroll") for (int i = 0; i < N; ++i) r.v[i] = PR::R2(i); return r; }
  ZK_HD static Fe modulus() { Fe r; _Pragma("unroll") for (int i = 0; i < N; ++i) r.v[i] = PR::P(i); return r; }

  ZK_HD bool is_zero() const { uint32_t o = 0; _Pragma("unroll") for (int i = 0; i < N; ++i) o |= v[i]; return o == 0; }
  ZK_HD bool operator==(const Fe& b) const { uint32_t o = 0; _Pragma("unroll") for (int i = 0; i < N; ++i) o |= v[i] ^ b.v[i]; return o == 0; }
  ZK_HD bool operator!=(const Fe& b) const { return !(*this == b); }
};

// r = a - p if a >= p (a < 2p)
template <class PR>
ZK_HD void final_sub(Fe<PR>& a) {
  constexpr int N = PR::N;
  uint32_t t[N];
  t[0] = sub_cc(a.v[0], PR::P(0));
  _Pragma("unroll") for (int i = 1; i < N; ++i) t[i] = subc_cc(a.v[i], PR::P(i));
  uint32_t borrow = subc(0u, 0u);                         // 0xffffffff if a < p
  _Pragma("unroll") for (int i = 0; i < N; ++i) a.v[i] = borrow ? a.v[i] : t[i];
}

template <class PR>
ZK_HD Fe<PR> add(const Fe<PR>& a, const Fe<PR>& b) {        // Scalar_add / Fp_add (bls12-381.cu:296-300)
  constexpr int N = PR::N;
  Fe<PR> r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
  _Pragma("unroll") for (int i = 1; i < N - 1; ++i) r.v[i] = addc_cc(a.v[i], b.v[i]);
  r.v[N - 1] = addc(a.v[N - 1], b.v[N - 1]);               // p < 2^(32N-1): no carry out
  final_sub(r);
  return r;
}

template <class PR>
ZK_HD Fe<PR> sub(const Fe<PR>& a, const Fe<PR>& b) {        // Scalar_sub / Fp_sub (bls12-381.cu:289-293)
  constexpr int N = PR::N;
  Fe<PR> r;
  r.v[0] = sub_cc(a.v[0], b.v[0]);
  _Pragma("unroll") for (int i = 1; i < N; ++i) r.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t mask = subc(0u, 0u);                           // all ones if borrow
  r.v[0] = add_cc(r.v[0], PR::P(0) & mask);
  _Pragma("unroll") for (int i = 1; i < N - 1; ++i) r.v[i] = addc_cc(r.v[i], PR::P(i) & mask);
  r.v[N - 1] = addc(r.v[N - 1], PR::P(N - 1) & mask);
  return r;
}

template <class PR>
ZK_HD Fe<PR> neg(const Fe<PR>& a) { return sub(Fe<PR>::zero(), a); }
template <class PR>
ZK_HD Fe<PR> dbl(const Fe<PR>& a) { return add(a, a); }

// ---- Montgomery product.  T = EV + 2^32 * OD; see DESIGN.md "Field arithmetic" for the invariant.
// acc pairs (acc[j],acc[j+1]) += x[j]*w for j = 0,2,..,N-2 in one carry chain; leaves the carry-out in CC.
template <int N>
ZK_HD void cmad_n(uint32_t* acc, const uint32_t* x, uint32_t w) {
  acc[0] = mad_lo_cc(x[0], w, acc[0]);
  acc[1] = madc_hi_cc(x[0], w, acc[1]);
  _Pragma("unroll") for (int j = 2; j < N; j += 2) {
    acc[j] = madc_lo_cc(x[j], w, acc[j]);
    acc[j + 1] = madc_hi_cc(x[j], w, acc[j + 1]);
  }
}
// same chain for the modulus (compile-time limbs, offset `off`)
template <class PR, int off>
ZK_HD void cmad_p(uint32_t* acc, uint32_t w) {
  constexpr int N = PR::N;
  acc[0] = mad_lo_cc(PR::P(off), w, acc[0]);
  acc[1] = madc_hi_cc(PR::P(off), w, acc[1]);
  _Pragma("unroll") for (int j = 2; j < N; j += 2) {
    acc[j] = madc_lo_cc(PR::P(j + off), w, acc[j]);
    acc[j + 1] = madc_hi_cc(PR::P(j + off), w, acc[j + 1]);
  }
}

// One CIOS row: (ev, od) hold T with od still carrying the previous row's un-shifted even limbs.
template <class PR, bool first>
ZK_HD void mont_row(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi) {
  constexpr int N = PR::N;
  if (first) {
    _Pragma("unroll") for (int j = 0; j < N; j += 2) {
      ev[j] = mul_lo(a[j], bi); ev[j + 1] = mul_hi(a[j], bi);
      od[j] = mul_lo(a[j + 1], bi); od[j + 1] = mul_hi(a[j + 1], bi);
    }
  } else {
    // T' = T / 2^32: the old odd array becomes the even one (ev here), the old even array shifted down by two
    // limbs becomes the odd one; old even limb 1 is pending at limb 0 and its carry enters the odd chain.
    ev[0] = add_cc(ev[0], od[1]);
    _Pragma("unroll") for (int j = 0; j < N - 2; j += 2) {
      od[j] = madc_lo_cc(a[j + 1], bi, od[j + 2]);
      od[j + 1] = madc_hi_cc(a[j + 1], bi, od[j + 3]);
    }
    od[N - 2] = madc_lo_cc(a[N - 1], bi, 0u);
    od[N - 1] = madc_hi(a[N - 1], bi, 0u);
    cmad_n<N>(ev, a, bi);
    od[N - 1] = addc(od[N - 1], 0u);
  }
  uint32_t m = mul_lo(ev[0], PR::INV);
  cmad_p<PR, 1>(od, m);
  cmad_p<PR, 0>(ev, m);
  od[N - 1] = addc(od[N - 1], 0u);
}

template <class PR>
ZK_HD Fe<PR> mul_impl(const Fe<PR>& a, const Fe<PR>& b) {   // Scalar_mul / Fp_mul (bls12-381.cu:462-494, 869-901)
  constexpr int N = PR::N;
  uint32_t ev[N], od[N];
  mont_row<PR, true>(ev, od, a.v, b.v[0]);
  _Pragma("unroll") for (int i = 1; i < N; i += 2) {
    mont_row<PR, false>(od, ev, a.v, b.v[i]);
    if (i + 1 < N) mont_row<PR, false>(ev, od, a.v, b.v[i + 1]);
  }
  // N is even: the last row ran with roles (od, ev), so the value is od_as_even... see below
  // After an odd number of role swaps the "even" array of the *next* row would be `ev`, pending `od` shifted.
  Fe<PR> r;
  r.v[0] = add_cc(ev[0], od[1]);
  _Pragma("unroll") for (int i = 1; i < N - 1; ++i) r.v[i] = addc_cc(ev[i], od[i + 1]);
  r.v[N - 1] = addc(ev[N - 1], 0u);
  final_sub(r);
  return r;
}

#if defined(__CUDA_ARCH__)
template <class PR>
__device__ __noinline__ Fe<PR> mul_outlined(Fe<PR> a, Fe<PR> b) { return mul_impl(a, b); }
#endif
template <class PR>
ZK_HD Fe<PR> mul(const Fe<PR>& a, const Fe<PR>& b) {
#if defined(__CUDA_ARCH__)
  if (PR::OUTLINE_MUL) return mul_outlined<PR>(a, b);
#endif
  return mul_impl(a, b);
}
template <class PR>
ZK_HD Fe<PR> sqr(const Fe<PR>& a) { return mul(a, a); }
template <class PR>
ZK_HD Fe<PR> to_mont(const Fe<PR>& a) { return mul(a, Fe<PR>::r2()); }     // Scalar_mont (bls12-381.cu:585-587)
template <class PR>
ZK_HD Fe<PR> from_mont(const Fe<PR>& a) {                                   // Scalar_unmont (bls12-381.cu:589-593)
  Fe<PR> one = Fe<PR>::zero(); one.v[0] = 1; return mul(a, one);
}
template <class PR>
ZK_HD bool gte(const Fe<PR>& a, const Fe<PR>& b) {                          // Scalar_gte (bls12-381.cu:244-252)
  constexpr int N = PR::N;
  uint32_t t = sub_cc(a.v[0], b.v[0]); (void)t;
  _Pragma("unroll") for (int i = 1; i < N; ++i) { t = subc_cc(a.v[i], b.v[i]); (void)t; }
  return subc(0u, 0u) == 0u;
}

// ---- lazy reduction: sums of products accumulated unreduced, one Montgomery reduction at the end.
// NOTE (measured, profiles/README.md): a fold kernel built on this (8 products + 1 reduction instead of 7 Montgomery
// products) ran 22 % SLOWER on B200 than k_fr_fold_multi<3>: propagating carries to the top of the 18-limb accumulator
// moves ~110 IADD3 per product onto the ALU pipe and lengthens the dependent chains.  Kept (host-tested) for the matmul /
// dot-product shapes where k is large; not used on the proving path.
// A dot product sum_t a_t * b_t of k Montgomery values costs k * N^2 wide multiplies + one reduction instead of
// k * (2 N^2 + N).  The accumulator is the exact integer sum (2N+2 limbs, carries propagated to the top), so the result
// is the same canonical field element as the sum of the k Montgomery products.
template <class PR>
struct WideAcc {
  static constexpr int N = PR::N, L = 2 * PR::N + 2;
  uint32_t ev[L], od[L];                       // value = EV + 2^32 * OD
};
template <class PR>
ZK_HD void wide_zero(WideAcc<PR>& w) { _Pragma("unroll") for (int i = 0; i < WideAcc<PR>::L; ++i) { w.ev[i] = 0; w.od[i] = 0; } }
// acc[0..N) += x[xoff], x[xoff+2], ... (N/2 limbs) * w as aligned pairs, carry propagated through acc[N..top)
template <int N>
ZK_HD void wide_chain(uint32_t* acc, const uint32_t* x, uint32_t w, const int TOP) {   // TOP: limbs from acc to the top (constant after unrolling)
  acc[0] = mad_lo_cc(x[0], w, acc[0]);
  acc[1] = madc_hi_cc(x[0], w, acc[1]);
  _Pragma("unroll") for (int j = 2; j < N; j += 2) {
    acc[j] = madc_lo_cc(x[j], w, acc[j]);
    acc[j + 1] = madc_hi_cc(x[j], w, acc[j + 1]);
  }
  _Pragma("unroll") for (int j = N; j < 2 * N + 2; ++j) {
    if (j < TOP - 1) acc[j] = addc_cc(acc[j], 0u);
    else if (j == TOP - 1) acc[j] = addc(acc[j], 0u);
  }
}
template <class PR>
ZK_HD void wide_mac(WideAcc<PR>& w, const Fe<PR>& a, const Fe<PR>& b) {          // w += a * b (integers)
  constexpr int N = PR::N, L = WideAcc<PR>::L;
  _Pragma("unroll") for (int i = 0; i < N; i += 2) {
    // b limb i (even): even limbs of a land on even positions i+j -> ev[i+j]; odd limbs on odd positions -> od[i+j-1]
    wide_chain<N>(w.ev + i, a.v, b.v[i], L - i);
    wide_chain<N>(w.od + i, a.v + 1, b.v[i], L - i);
    // b limb i+1 (odd): even limbs of a land on odd positions i+1+j -> od[i+j]; odd limbs on even positions -> ev[i+1+j]
    wide_chain<N>(w.od + i, a.v, b.v[i + 1], L - i);
    wide_chain<N>(w.ev + i + 2, a.v + 1, b.v[i + 1], L - i - 2);
  }
}
// Montgomery-reduce the accumulated integer T: T / R mod p, canonical.  Valid for Fr (R < 5p) and up to 2^32 products.
template <class PR>
ZK_HD Fe<PR> wide_reduce(const WideAcc<PR>& w) {
  constexpr int N = PR::N, L = WideAcc<PR>::L;
  uint32_t t[L];
  t[0] = w.ev[0];
  t[1] = add_cc(w.ev[1], w.od[0]);
  _Pragma("unroll") for (int i = 2; i < L - 1; ++i) t[i] = addc_cc(w.ev[i], w.od[i - 1]);
  t[L - 1] = addc(w.ev[L - 1], w.od[L - 2]);
  Fe<PR> lo, hi, one = Fe<PR>::zero();
  one.v[0] = 1;
  _Pragma("unroll") for (int i = 0; i < N; ++i) { lo.v[i] = t[i]; hi.v[i] = t[N + i]; }
  // T = t[2N] * R^2 + hi * R + lo, so T / R = t[2N] * R + hi + lo / R  (mod p).  R mod p is the Montgomery ONE; t[2N] is
  // at most (number of products) / 4 (p^2 < R^2 / 4), t[2N+1] is zero for any sum of fewer than 2^32 products.
  _Pragma("unroll") for (int i = 0; i < 5; ++i) final_sub(hi);          // hi < R < 5p (Fr: R = 4.4p; Fq callers: R = 9.9p, not used)
  Fe<PR> r = add(hi, mul_impl(lo, one));                               // lo / R mod p
  const Fe<PR> rmodp = Fe<PR>::one();
  for (uint32_t c = 0; c < t[2 * N]; ++c) r = add(r, rmodp);
  return r;
}

typedef Fe<FrParams> Fr;
typedef Fe<FqParams> Fq;

}  // namespace zk
