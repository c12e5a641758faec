// Host-compiled shim over zkdl_b200/csrc/*.cuh: the device templates with the PTX carry chain emulated on the host
// (field.cuh).  Lets the CPU test-suite run the exact arithmetic the kernels instantiate against the oracle.
#include "../zkdl_b200/csrc/field.cuh"
#include "../zkdl_b200/csrc/fr_device.cuh"
#include "../zkdl_b200/csrc/g1_device.cuh"
#include <cstddef>
using namespace zk;
extern "C" {
void hs_fr_mul(const Fr* a, const Fr* b, Fr* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = mul(a[i], b[i]); }
void hs_fr_add(const Fr* a, const Fr* b, Fr* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = add(a[i], b[i]); }
void hs_fr_sub(const Fr* a, const Fr* b, Fr* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = sub(a[i], b[i]); }
void hs_fr_mont(const Fr* a, Fr* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = to_mont(a[i]); }
void hs_fr_unmont(const Fr* a, Fr* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = from_mont(a[i]); }
void hs_fr_gte(const Fr* a, const Fr* b, int* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = gte(a[i], b[i]); }
void hs_fq_mul(const Fq* a, const Fq* b, Fq* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = mul(a[i], b[i]); }
void hs_fq_add(const Fq* a, const Fq* b, Fq* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = add(a[i], b[i]); }
void hs_fq_sub(const Fq* a, const Fq* b, Fq* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = sub(a[i], b[i]); }
void hs_fq_inv(const Fq* a, Fq* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = fq_inv(a[i]); }
// lazy reduction: out[i] = sum_{t<k} a[i*k+t] * b[i*k+t] / R mod p through the unreduced wide accumulator
void hs_fr_dot_lazy(const Fr* a, const Fr* b, Fr* o, size_t n, size_t k) {
  for (size_t i = 0; i < n; ++i) {
    WideAcc<FrParams> w; wide_zero(w);
    for (size_t t = 0; t < k; ++t) wide_mac(w, a[i * k + t], b[i * k + t]);
    o[i] = wide_reduce(w);
  }
}
// per-pair sumcheck math
void hs_fold_pair(const Fr* a0, const Fr* a1, const Fr* x, Fr* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = fold_pair(a0[i], a1[i], *x); }
void hs_ip_pair(const Fr* a0, const Fr* a1, const Fr* b0, const Fr* b1, const Fr* e, const Fr* x, int weighted, Fr* c, Fr* ao, Fr* bo, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    if (weighted) ip_pair<true>(a0[i], a1[i], b0[i], b1[i], e[i], *x, c + 3 * i, ao[i], bo[i]);
    else ip_pair<false>(a0[i], a1[i], b0[i], b1[i], e[i], *x, c + 3 * i, ao[i], bo[i]);
  }
}
void hs_bin_pair(const Fr* a0, const Fr* a1, const Fr* e, const Fr* x, Fr* c, Fr* ao, size_t n) {
  for (size_t i = 0; i < n; ++i) ao[i] = bin_pair(a0[i], a1[i], e[i], *x, c + 3 * i);
}
void hs_bin_pair_c12(const Fr* a0, const Fr* a1, const Fr* e, const Fr* x, Fr* c, Fr* ao, size_t n) {
  for (size_t i = 0; i < n; ++i) ao[i] = bin_pair_c12(a0[i], a1[i], e[i], *x, c + 3 * i);
}
void hs_to_mont_u64(const uint64_t* m, Fr* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = to_mont_u64(m[i]); }
void hs_to_mont_u32(const uint32_t* m, Fr* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = to_mont_u32(m[i]); }
void hs_relu_i64(const long long* v, uint32_t* q, uint16_t* r, int* pos, int* bad, size_t n) {
  for (size_t i = 0; i < n; ++i) { ReluParts p = relu_decompose_i64(v[i]); q[i] = p.q; r[i] = p.r; pos[i] = p.positive; bad[i] = p.out_of_range; }
}
void hs_float_to_fr(const float* f, Fr* o, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = float_to_fr(f[i]); }
void hs_relu(const Fr* x, uint32_t* q, uint16_t* r, int* pos, int* bad, size_t n) {
  for (size_t i = 0; i < n; ++i) { ReluParts p = relu_decompose(x[i]); q[i] = p.q; r[i] = p.r; pos[i] = p.positive; bad[i] = p.out_of_range; }
}
// MSM scalar preparation + signed-digit recoding: digits[i*W + w], returns via sign[i]
void hs_digits(const Fr* s, int mont, int c, int W, int32_t* digits, int* sign, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    Fr mag; bool neg; scalar_prepare(s[i], mont != 0, mag, neg); sign[i] = neg;
    uint32_t carry = 0;
    for (int w = 0; w < W; ++w) digits[i * W + w] = next_digit(mag, w, c, carry);
    if (carry) digits[i * W + W - 1] = 0x7fffffff;   // must never happen (W = ceil(255/c))
  }
}
// mixed-width recoding of the full-table MSM: windows [0, nfull) are c bits wide, the rest ctop bits (msm.cu, MsmCfg)
void hs_digits_mixed(const Fr* s, int mont, int c, int nfull, int ctop, int W, int32_t* digits, int* sign, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    Fr mag; bool neg; scalar_prepare(s[i], mont != 0, mag, neg); sign[i] = neg;
    uint32_t carry = 0;
    for (int w = 0; w < W; ++w) {
      int bit = w < nfull ? w * c : nfull * c + (w - nfull) * ctop;
      digits[i * W + w] = next_digit_at(mag, bit, w < nfull ? c : ctop, carry);
    }
    if (carry) digits[i * W + W - 1] = 0x7fffffff;
  }
}
// integer-weighted sums: out[i] = sum_t w[i*k+t] * e[i*k+t] mod p through the 320-bit accumulators (matmul.cu weight folds)
void hs_isum(const int32_t* w, const Fr* e, Fr* o, size_t n, size_t k) {
  for (size_t i = 0; i < n; ++i) {
    ISum pos, neg, pos2; isum_zero(pos); isum_zero(neg); isum_zero(pos2);
    for (size_t t = 0; t < k; ++t) {
      int32_t v = w[i * k + t];
      if (v >= 0) isum_mac((t & 1) ? pos2 : pos, (uint32_t)v, e[i * k + t]); else isum_mac(neg, 0u - (uint32_t)v, e[i * k + t]);
    }
    isum_add(pos, pos2);
    o[i] = sub(isum_reduce(pos), isum_reduce(neg));
  }
}
// G1: XYZZ formulas through the Jacobian PODs
static G1Jac J(const uint32_t* p) { G1Jac r; for (int i = 0; i < 12; ++i) { r.x.v[i] = p[i]; r.y.v[i] = p[12 + i]; r.z.v[i] = p[24 + i]; } return r; }
static void S(uint32_t* p, const G1Jac& r) { for (int i = 0; i < 12; ++i) { p[i] = r.x.v[i]; p[12 + i] = r.y.v[i]; p[24 + i] = r.z.v[i]; } }
void hs_g1_add(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) {
  for (size_t i = 0; i < n; ++i) S(o + 36 * i, xyzz_to_jac(xyzz_add(xyzz_from_jac(J(a + 36 * i)), xyzz_from_jac(J(b + 36 * i)))));
}
void hs_g1_dbl(const uint32_t* a, uint32_t* o, size_t n) { for (size_t i = 0; i < n; ++i) S(o + 36 * i, xyzz_to_jac(xyzz_dbl(xyzz_from_jac(J(a + 36 * i))))); }
void hs_g1_madd(const uint32_t* a, const uint32_t* b_aff, int negate, uint32_t* o, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    G1XYZZ acc = xyzz_from_jac(J(a + 36 * i)); G1Affine b;
    for (int k = 0; k < 12; ++k) { b.x.v[k] = b_aff[24 * i + k]; b.y.v[k] = b_aff[24 * i + 12 + k]; }
    xyzz_madd(acc, b, negate != 0); S(o + 36 * i, xyzz_to_jac(acc));
  }
}
void hs_g1_mul_small(const uint32_t* a, const uint32_t* k, uint32_t* o, size_t n) {
  for (size_t i = 0; i < n; ++i) S(o + 36 * i, xyzz_to_jac(xyzz_mul_small(xyzz_from_jac(J(a + 36 * i)), k[i])));
}
void hs_g1_to_affine(const uint32_t* a, uint32_t* o_aff, size_t n) {     // via fq_inv(ZZZ), as k_batch_affine does
  for (size_t i = 0; i < n; ++i) {
    G1XYZZ p = xyzz_from_jac(J(a + 36 * i));
    G1Affine r; r.x = Fq::zero(); r.y = Fq::zero();
    if (!is_inf(p)) r = xyzz_to_affine_with_inv(p, fq_inv(p.zzz));
    for (int k = 0; k < 12; ++k) { o_aff[24 * i + k] = r.x.v[k]; o_aff[24 * i + 12 + k] = r.y.v[k]; }
  }
}
}
