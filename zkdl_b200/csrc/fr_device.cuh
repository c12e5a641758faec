// fr_device.cuh — per-element math of the Fr kernels as host/device functions, so the CPU test-suite can run the
// exact code the kernels instantiate (tests/host_shim.cpp) against the oracle.
#pragma once
#include <math.h>
#include "field.cuh"

namespace zk {

// Fr_me_step / Fr_partial_me_step pair rule (/root/reference/fr-tensor.cu:404-408): a0 + x (a1 - a0); a missing a1 is 0,
// for which the reference's a0 - x*a0 is the same field element.
ZK_HD Fr fold_pair(const Fr& a0, const Fr& a1, const Fr& x) { return add(a0, mul(x, sub(a1, a0))); }

// Fr_ip_sc_step coefficients (/root/reference/proof.cu:55-70), optionally pre-weighted by e (Hadamard sumcheck,
// proof.cu:120-129 evaluates the three coefficient vectors at u[1:], i.e. an eq-weighted sum), plus the two folds.
//   c0 = a0 b0, c1 = a0 (b1-b0) + b0 (a1-a0), c2 = (a1-a0)(b1-b0);   c1 via a1 b1 - c0 - c2 (3 products instead of 4)
template <bool WEIGHTED>
ZK_HD void ip_pair(const Fr& a0, const Fr& a1, const Fr& b0, const Fr& b1, const Fr& e, const Fr& x, Fr* c, Fr& a_out, Fr& b_out) {
  Fr da = sub(a1, a0), db = sub(b1, b0);
  Fr wa0 = WEIGHTED ? mul(e, a0) : a0;
  Fr wda = WEIGHTED ? mul(e, da) : da;
  c[0] = mul(wa0, b0);
  c[2] = mul(wda, db);
  Fr t = mul(add(wa0, wda), b1);
  c[1] = sub(sub(t, c[0]), c[2]);
  a_out = add(a0, mul(x, da));
  b_out = add(b0, mul(x, db));
}

// Fr_bin_sc_step coefficients (/root/reference/proof.cu:152-163) weighted by e, plus the fold with x:
//   c0 = a0^2 - a0 = a0 (a0 - 1), c1 = 2 a0 d - d = d (2 a0 - 1), c2 = d^2, d = a1 - a0
ZK_HD Fr bin_pair(const Fr& a0, const Fr& a1, const Fr& e, const Fr& x, Fr* c) {
  Fr d = sub(a1, a0);
  Fr ea0 = mul(e, a0), ed = mul(e, d);
  Fr one = Fr::one();
  c[0] = mul(ea0, sub(a0, one));
  c[1] = mul(ed, sub(dbl(a0), one));
  c[2] = mul(ed, d);
  return add(a0, mul(x, d));
}

// The same pair without c0 (4 products instead of 6): every round's verifier identity  claim_j = c0 + u_j (c1 + c2)  fixes c0
// once the running claim is known, so only c1 and c2 have to be SUMMED over the table; the kernel that finishes a round
// derives c0 = claim_j - u_j (c1 + c2) and the next claim c0 + v_j (c1 + v_j c2).  Exact field identities: same proof elements.
ZK_HD Fr bin_pair_c12(const Fr& a0, const Fr& a1, const Fr& e, const Fr& x, Fr* c) {
  Fr d = sub(a1, a0);
  Fr ed = mul(e, d);
  c[1] = mul(ed, sub(dbl(a0), Fr::one()));
  c[2] = mul(ed, d);
  return add(a0, mul(x, d));
}

// ---- sums of (small integer) x (field element): an unsigned 320-bit accumulator, reduced once at the end.
// Used to fold a table of quantised weights against an eq table without a Montgomery product per entry:
// sum_i w_i * E_i mod p with |w_i| < 2^32 is the same field element as the Montgomery sum of products mont(w_i) (x) E_i.
struct ISum { uint32_t v[10]; };
ZK_HD void isum_zero(ISum& a) { for (int i = 0; i < 10; ++i) a.v[i] = 0; }
ZK_HD void isum_mac(ISum& a, uint32_t s, const Fr& e) {            // a += s * e   (up to 2^32 terms before overflow)
  uint64_t c = 0;
  for (int j = 0; j < 8; ++j) { uint64_t t = (uint64_t)s * e.v[j] + a.v[j] + c; a.v[j] = (uint32_t)t; c = t >> 32; }
  uint64_t t = (uint64_t)a.v[8] + c; a.v[8] = (uint32_t)t; a.v[9] += (uint32_t)(t >> 32);
}
ZK_HD void isum_add(ISum& a, const ISum& b) {
  uint64_t c = 0;
  for (int j = 0; j < 10; ++j) { uint64_t t = (uint64_t)a.v[j] + b.v[j] + c; a.v[j] = (uint32_t)t; c = t >> 32; }
}
ZK_HD Fr isum_reduce(const ISum& a) {                               // a mod p
  Fr lo; for (int j = 0; j < 8; ++j) lo.v[j] = a.v[j];
  for (int i = 0; i < 5; ++i) final_sub(lo);                        // 2^256 < 5p
  Fr hi = Fr::zero(); hi.v[0] = a.v[8]; hi.v[1] = a.v[9];
  return add(lo, to_mont(hi));                                      // hi * 2^256 mod p
}

// float_to_Fr (/root/reference/zkfc.cu:63-78): round-half-away(|x| * 2^16) as u32 (saturating, NaN -> 0), sign from the
// sign bit; NOT Montgomery.
ZK_HD Fr float_to_fr(float x) {
  x = x * 65536.0f;
  float ax = roundf(fabsf(x));
  bool negative = signbit(x);
  uint32_t v;
#if defined(__CUDA_ARCH__)
  v = __float2uint_rz(ax);
#else
  if (ax != ax) v = 0; else if (ax >= 4294967296.0f) v = 0xffffffffu; else v = (uint32_t)ax;
#endif
  Fr r = Fr::zero(); r.v[0] = v;
  return negative ? sub(Fr::zero(), r) : r;
}

// relu_kernel decomposition (/root/reference/zkrelu.cu:11-41)
struct ReluParts { uint32_t q; uint16_t r; bool positive; bool out_of_range; };
// where zkReLU::operator()'s outputs go when it is applied in the forward product's epilogue (act == nullptr: not applied)
struct ReluOut { Fr* act; Fr* sign; uint32_t* mag; uint16_t* rem; uint32_t* bad; };
ZK_HD ReluParts relu_decompose(const Fr& X) {
  Fr x = from_mont(X);
  ReluParts p; p.positive = false; p.out_of_range = false;
  uint64_t mag = 0;
  bool hi_zero = (x.v[2] | x.v[3] | x.v[4] | x.v[5] | x.v[6] | x.v[7]) == 0;
  if (hi_zero && x.v[1] <= 32767u) {                        // x <= 2^47 - 1
    p.positive = true;
    mag = (uint64_t)x.v[0] | ((uint64_t)x.v[1] << 32);
  } else {
    Fr lim = Fr::modulus();                                 // p - 2^47 = {1, 0xffff7fff, p2, ...} (zkrelu.cu:23)
    lim.v[1] -= 32768u;
    if (gte(x, lim)) {
      Fr t = Fr::zero(); t.v[1] = 32768u;
      Fr s = add(x, t);                                     // x + 2^47 - p = 2^47 - |x|
      mag = (uint64_t)s.v[0] | ((uint64_t)s.v[1] << 32);
    } else {
      p.out_of_range = true;                                // undefined in the reference (uninitialised mag/sign)
    }
  }
  bool rem_sign = (mag & 32768ull) != 0;
  uint32_t rem_mag = (uint32_t)(mag & 32767ull);
  int32_t rem = rem_sign ? (int32_t)rem_mag - 32768 : (int32_t)rem_mag;
  p.q = (uint32_t)((mag - (uint64_t)(int64_t)rem) >> 16);
  p.r = (uint16_t)(rem_mag | (rem_sign ? 0x8000u : 0u));
  return p;
}


// m * R mod p for a 64-bit integer m (the Montgomery image of a small integer) in TWO CIOS rows instead of eight:
// mont(a, b) with a two-limb b is a * b * 2^-64, so a = 2^320 mod p gives m * 2^256.  4x fewer wide multiplies than
// to_mont(); used where an epilogue converts exact integer results (matmul_umma.cu).
ZK_HD Fr to_mont_u64(uint64_t m) {
  constexpr int N = 8;
  const uint32_t a[N] = {0x0121c884u, 0xc98da28eu, 0xc7363c67u, 0xe6f4f4a0u, 0xe92e7df1u, 0xb2d6ebc4u, 0x9d26242au, 0x19ae5794u};   // 2^320 mod p
  uint32_t ev[N], od[N];
  mont_row<FrParams, true>(ev, od, a, (uint32_t)m);
  mont_row<FrParams, false>(od, ev, a, (uint32_t)(m >> 32));
  Fr r;
  r.v[0] = add_cc(ev[0], od[1]);
  _Pragma("unroll") for (int i = 1; i < N - 1; ++i) r.v[i] = addc_cc(ev[i], od[i + 1]);
  r.v[N - 1] = addc(ev[N - 1], 0u);
  final_sub(r);
  return r;
}

// the same for a 32-bit integer: ONE row against 2^288 mod p (the rescaled ReLU magnitudes are 32-bit)
ZK_HD Fr to_mont_u32(uint32_t m) {
  constexpr int N = 8;
  const uint32_t a[N] = {0xcaaf6b13u, 0x355094eau, 0x69a568efu, 0xf6b10cb3u, 0x40cc3869u, 0xe2c926a6u, 0xed269aadu, 0x736a6d3bu};   // 2^288 mod p
  uint32_t ev[N], od[N];
  mont_row<FrParams, true>(od, ev, a, m);                 // roles as after the last row of an even-length product
  Fr r;
  r.v[0] = add_cc(ev[0], od[1]);
  _Pragma("unroll") for (int i = 1; i < N - 1; ++i) r.v[i] = addc_cc(ev[i], od[i + 1]);
  r.v[N - 1] = addc(ev[N - 1], 0u);
  final_sub(r);
  return r;
}

// relu_decompose for an exact integer pre-activation (the forward product's accumulator): same parts as
// relu_decompose(to_mont(v mod p)) without the field round trip.
ZK_HD ReluParts relu_decompose_i64(long long v) {
  ReluParts p;
  p.positive = v >= 0 && v < (1ll << 47);
  const bool negative = v < 0 && v >= -(1ll << 47);
  p.out_of_range = !p.positive && !negative;
  const uint64_t mag = p.positive ? (uint64_t)v : negative ? (uint64_t)((1ll << 47) + v) : 0ull;
  const bool rem_sign = (mag & 32768ull) != 0;
  const uint32_t rem_mag = (uint32_t)(mag & 32767ull);
  const int32_t rem = rem_sign ? (int32_t)rem_mag - 32768 : (int32_t)rem_mag;
  p.q = (uint32_t)((mag - (uint64_t)(int64_t)rem) >> 16);
  p.r = (uint16_t)(rem_mag | (rem_sign ? 0x8000u : 0u));
  return p;
}

}  // namespace zk
