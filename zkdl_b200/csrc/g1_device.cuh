// g1_device.cuh — per-element logic of the MSM pipeline (scalar normalisation, signed-digit recoding, Fq inversion)
// as host/device functions so the CPU suite can exercise it through tests/host_shim.cpp.
#pragma once
#include "g1.cuh"

namespace zk {

// Scalar -> (|s| mod r, sign).  mont = true: Montgomery Fr is un-Montgomery'd first (Commitment::commit,
// /root/reference/commitment.cu:33); false: the raw limbs are the integer (G1Jacobian_mul, g1-tensor.cu:422-430).
// The 256-bit integer is reduced mod r (G1 has prime order r, so [x]P = [x mod r]P) and mapped to the shorter of
// x and r - x with a sign, which leaves one spare bit for the signed-digit carry.
ZK_HD void scalar_prepare(const Fr& s_in, bool mont, Fr& mag, bool& negative) {
  Fr s;
  if (mont) s = from_mont(s_in);
  else { s = s_in; final_sub(s); final_sub(s); }            // < 2^256 < 3r
  Fr t = sub(Fr::zero(), s);                                // r - s  (0 for s = 0)
  if (!s.is_zero() && !gte(t, s)) { mag = t; negative = true; } else { mag = s; negative = false; }
}

// bits [pos, pos + c) of a 256-bit little-endian integer, c <= 24
ZK_HD uint32_t get_bits(const Fr& a, int pos, int c) {
  if (pos >= 256) return 0;
  int limb = pos >> 5, off = pos & 31;
  uint32_t lo = a.v[limb] >> off;
  if (off + c > 32 && limb + 1 < 8) lo |= a.v[limb + 1] << (32 - off);
  return lo & ((1u << c) - 1u);
}
// signed digit of the c-bit window starting at bit pos: value in [-2^(c-1)+1, 2^(c-1)], carry threaded by the caller
ZK_HD int32_t next_digit_at(const Fr& mag, int pos, int c, uint32_t& carry) {
  uint32_t raw = get_bits(mag, pos, c) + carry;
  if (raw > (1u << (c - 1))) { carry = 1; return (int32_t)raw - (int32_t)(1u << c); }
  carry = 0;
  return (int32_t)raw;
}
ZK_HD int32_t next_digit(const Fr& mag, int w, int c, uint32_t& carry) { return next_digit_at(mag, w * c, c, carry); }

// a^(p-2) in Fq (the reference has no inversion; used for table normalisation and zkdl_g1_normalize)
ZK_HD Fq fq_inv(const Fq& a) {
  Fq acc = a;                                               // bit 380 of p-2 is set
  for (int i = 379; i >= 0; --i) {
    acc = sqr(acc);
    uint32_t limb = FqParams::P(i >> 5);
    if ((i >> 5) == 0) limb -= 2u;                          // p - 2: low limb 0xffffaaab - 2, no borrow
    if ((limb >> (i & 31)) & 1u) acc = mul(acc, a);
  }
  return acc;
}

ZK_HD G1Affine xyzz_to_affine_with_inv(const G1XYZZ& p, const Fq& izzz) {   // izzz = 1/ZZZ
  Fq zinv = mul(p.zz, izzz);                                // ZZ/ZZZ = 1/Z
  Fq izz = sqr(zinv);
  G1Affine r; r.x = mul(p.x, izz); r.y = mul(p.y, izzz); return r;
}

}  // namespace zk
