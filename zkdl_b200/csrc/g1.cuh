// g1.cuh — BLS12-381 G1 group law for the MSM / commitment kernels.
//
// Replaces blstrs__g1__G1Affine_{double,add_mixed,add} (/root/reference/bls12-381.cu:1331-1435) and the bit-serial
// G1Jacobian_mul ladder (/root/reference/g1-tensor.cu:422-430).  The reference keeps Jacobian triples; here bucket
// accumulators are XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; EFD "xyzz" madd-2008-s / add-2008-s / dbl-2008-s-1),
// which makes the dominant operation (accumulator += affine base) 8M+2S instead of madd-2007-bl's 7M+4S with a
// shorter dependency chain, and needs no special Z handling.  API-edge points stay the reference's PODs:
// affine {x,y} (96 B) and Jacobian {x,y,z} (144 B, infinity <=> z == 0).  All formulas are complete: P+P, P+(-P)
// and infinity operands are handled, so results are the same group elements the reference computes.
#pragma once
#include "field.cuh"

namespace zk {

struct G1Affine { Fq x, y; };                 // infinity encoded as (0,0) (not on y^2 = x^3 + 4)
struct G1Jac { Fq x, y, z; };                 // reference layout, infinity <=> z == 0
struct G1XYZZ { Fq x, y, zz, zzz; };          // infinity <=> zz == 0

ZK_HD bool is_inf(const G1Affine& p) { return p.x.is_zero() && p.y.is_zero(); }
ZK_HD bool is_inf(const G1Jac& p) { return p.z.is_zero(); }
ZK_HD bool is_inf(const G1XYZZ& p) { return p.zz.is_zero(); }

ZK_HD G1XYZZ xyzz_inf() { G1XYZZ r; r.x = Fq::zero(); r.y = Fq::one(); r.zz = Fq::zero(); r.zzz = Fq::zero(); return r; }
ZK_HD G1Jac jac_inf() { G1Jac r; r.x = Fq::zero(); r.y = Fq::one(); r.z = Fq::zero(); return r; }   // G1Affine_ZERO (.cuh:419)

ZK_HD G1XYZZ xyzz_from_affine(const G1Affine& p) {
  if (is_inf(p)) return xyzz_inf();
  G1XYZZ r; r.x = p.x; r.y = p.y; r.zz = Fq::one(); r.zzz = Fq::one(); return r;
}
ZK_HD G1XYZZ xyzz_from_jac(const G1Jac& p) {
  if (is_inf(p)) return xyzz_inf();
  G1XYZZ r; r.x = p.x; r.y = p.y; r.zz = sqr(p.z); r.zzz = mul(r.zz, p.z); return r;
}
// Jacobian representative with Z = ZZ: (X*ZZ, Y*ZZZ, ZZ)  [x = X/ZZ = X*ZZ/ZZ^2, y = Y/ZZZ = Y*ZZZ/ZZ^3]
ZK_HD G1Jac xyzz_to_jac(const G1XYZZ& p) {
  if (is_inf(p)) return jac_inf();
  G1Jac r; r.x = mul(p.x, p.zz); r.y = mul(p.y, p.zzz); r.z = p.zz; return r;
}
ZK_HD G1XYZZ xyzz_neg(const G1XYZZ& p) { G1XYZZ r = p; r.y = neg(p.y); return r; }

// dbl-2008-s-1 (a = 0): 6M + 3S
ZK_HD G1XYZZ xyzz_dbl(const G1XYZZ& p) {
  if (is_inf(p)) return p;
  Fq u = dbl(p.y);
  Fq v = sqr(u);
  Fq w = mul(u, v);
  Fq s = mul(p.x, v);
  Fq xx = sqr(p.x);
  Fq m = add(dbl(xx), xx);
  G1XYZZ r;
  r.x = sub(sub(sqr(m), s), s);
  r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
  r.zz = mul(v, p.zz);
  r.zzz = mul(w, p.zzz);
  return r;
}
// mdbl-2008-s-1: affine -> XYZZ double
ZK_HD G1XYZZ xyzz_dbl_affine(const G1Affine& p) {
  if (is_inf(p)) return xyzz_inf();
  Fq u = dbl(p.y);
  G1XYZZ r;
  r.zz = sqr(u);
  r.zzz = mul(u, r.zz);
  Fq s = mul(p.x, r.zz);
  Fq xx = sqr(p.x);
  Fq m = add(dbl(xx), xx);
  r.x = sub(sub(sqr(m), s), s);
  r.y = sub(mul(m, sub(s, r.x)), mul(r.zzz, p.y));
  return r;
}

// madd-2008-s: acc += (neg ? -b : b), b affine.  8M + 2S on the common path.
ZK_HD void xyzz_madd(G1XYZZ& a, const G1Affine& b_in, bool negate) {
  if (is_inf(b_in)) return;
  G1Affine b = b_in;
  if (negate) b.y = neg(b.y);
  if (is_inf(a)) { a.x = b.x; a.y = b.y; a.zz = Fq::one(); a.zzz = Fq::one(); return; }
  Fq p = sub(mul(b.x, a.zz), a.x);
  Fq r = sub(mul(b.y, a.zzz), a.y);
  if (p.is_zero()) {
    if (r.is_zero()) a = xyzz_dbl_affine(b); else a = xyzz_inf();
    return;
  }
  Fq pp = sqr(p);
  Fq ppp = mul(p, pp);
  Fq q = mul(a.x, pp);
  Fq x3 = sub(sub(sub(sqr(r), ppp), q), q);
  a.y = sub(mul(r, sub(q, x3)), mul(a.y, ppp));
  a.x = x3;
  a.zz = mul(a.zz, pp);
  a.zzz = mul(a.zzz, ppp);
}

// add-2008-s: 12M + 2S
ZK_HD G1XYZZ xyzz_add(const G1XYZZ& a, const G1XYZZ& b) {
  if (is_inf(a)) return b;
  if (is_inf(b)) return a;
  Fq u1 = mul(a.x, b.zz), u2 = mul(b.x, a.zz);
  Fq s1 = mul(a.y, b.zzz), s2 = mul(b.y, a.zzz);
  Fq p = sub(u2, u1), r = sub(s2, s1);
  if (p.is_zero()) {
    if (r.is_zero()) return xyzz_dbl(a);
    return xyzz_inf();
  }
  Fq pp = sqr(p);
  Fq ppp = mul(p, pp);
  Fq q = mul(u1, pp);
  G1XYZZ o;
  o.x = sub(sub(sub(sqr(r), ppp), q), q);
  o.y = sub(mul(r, sub(q, o.x)), mul(s1, ppp));
  o.zz = mul(mul(a.zz, b.zz), pp);
  o.zzz = mul(mul(a.zzz, b.zzz), ppp);
  return o;
}

// [k] P for a small non-negative k (< 2^31): MSB-first double-and-add; used for bucket-segment offsets.
ZK_HD G1XYZZ xyzz_mul_small(const G1XYZZ& p, uint32_t k) {
  G1XYZZ r = xyzz_inf();
  if (k == 0) return r;
  int top = 31;
  while (!((k >> top) & 1)) --top;
  for (int i = top; i >= 0; --i) {
    r = xyzz_dbl(r);
    if ((k >> i) & 1) r = xyzz_add(r, p);
  }
  return r;
}

}  // namespace zk
