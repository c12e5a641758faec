/* zkdl_oracle.c — CPU restatement of the zkDL reference's FC-layer proof path (see zkdl_oracle.h header note:
 * TEST INFRASTRUCTURE ONLY; pinned by reference-generated fixtures in tests/golden, not by reference tests —
 * the reference has none).  Plain C, 64-bit limbs + unsigned __int128.  Build: make -C oracle
 */
#include "zkdl_oracle.h"
#include <string.h>
#include <stdlib.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------------
 * constants (bls12-381.cu:3-11, bls12-381.cuh:119,221) re-derived with Python big-ints, see tests
 * ---------------------------------------------------------------------------------------------- */
static const uint64_t FR_P[4]  = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t FR_ONE[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL};
static const uint64_t FR_R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL};
static const uint64_t FR_INV = 0xfffffffeffffffffULL;

static const uint64_t FQ_P[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                 0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t FQ_ONE[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                                   0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
static const uint64_t FQ_INV = 0x89f3fffcfffcfffdULL;
/* g1-tensor.cuh:28-57 (Montgomery form of the standard generator) */
static const uint64_t G1_GEN_X[6] = {0x5cb38790fd530c16ULL, 0x7817fc679976fff5ULL, 0x154f95c7143ba1c1ULL,
                                     0xf0ae6acdf3d0e747ULL, 0xedce6ecc21dbf440ULL, 0x120177419e0bfb75ULL};
static const uint64_t G1_GEN_Y[6] = {0xbaac93d50ce72271ULL, 0x8c22631a7918fd8eULL, 0xdd595f13570725ceULL,
                                     0x51ac582950405194ULL, 0x0e1c8c3fad0059c0ULL, 0x0bbc3efc5008a26aULL};

/* ------------------------------------------------------------------------------------------------
 * generic N-limb Montgomery field (CIOS, bls12-381.cu:505-535 `_mul_default`; the CUDA `_mul_nvidia`
 * path :462-494 computes the same canonical value a*b*R^-1 mod p)
 * ---------------------------------------------------------------------------------------------- */
static inline int ge_n(const uint64_t* a, const uint64_t* b, int N) {           /* _gte :244-252 */
  for (int i = N - 1; i >= 0; --i) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
  return 1;
}
static inline uint64_t add_n(uint64_t* r, const uint64_t* a, const uint64_t* b, int N) {
  u128 c = 0;
  for (int i = 0; i < N; ++i) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static inline uint64_t sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b, int N) {
  uint64_t br = 0;
  for (int i = 0; i < N; ++i) {
    u128 d = (u128)a[i] - b[i] - br; r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1;
  }
  return br;
}
static inline void modadd_n(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* p, int N) { /* :296-300 */
  uint64_t t[6]; add_n(t, a, b, N);           /* p < 2^(64N-1): no carry out for canonical inputs */
  if (ge_n(t, p, N)) sub_n(t, t, p, N);
  memcpy(r, t, 8 * N);
}
static inline void modsub_n(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* p, int N) { /* :289-293 */
  uint64_t t[6]; uint64_t br = sub_n(t, a, b, N);
  if (br) add_n(t, t, p, N);
  memcpy(r, t, 8 * N);
}
static inline void montmul_n(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* p, uint64_t inv, int N) {
  uint64_t t[8] = {0};
  for (int i = 0; i < N; ++i) {
    u128 c = 0;
    for (int j = 0; j < N; ++j) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[N]; t[N] = (uint64_t)c; t[N + 1] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * inv;
    c = (u128)m * p[0] + t[0]; c >>= 64;
    for (int j = 1; j < N; ++j) { c += (u128)m * p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[N]; t[N - 1] = (uint64_t)c; t[N] = t[N + 1] + (uint64_t)(c >> 64);
  }
  if (t[N] || ge_n(t, p, N)) sub_n(t, t, p, N);
  memcpy(r, t, 8 * N);
}

/* ---- Fr ---- */
static inline void fr_add(ofr_t* r, const ofr_t* a, const ofr_t* b) { modadd_n(r->l, a->l, b->l, FR_P, 4); }
static inline void fr_sub(ofr_t* r, const ofr_t* a, const ofr_t* b) { modsub_n(r->l, a->l, b->l, FR_P, 4); }
static inline void fr_mul(ofr_t* r, const ofr_t* a, const ofr_t* b) { montmul_n(r->l, a->l, b->l, FR_P, FR_INV, 4); }
static inline void fr_dbl(ofr_t* r, const ofr_t* a) { fr_add(r, a, a); }       /* :544-550 same value */
static inline void fr_mont(ofr_t* r, const ofr_t* a) { ofr_t r2; memcpy(r2.l, FR_R2, 32); fr_mul(r, a, &r2); }   /* :585-587 */
static inline void fr_unmont(ofr_t* r, const ofr_t* a) { ofr_t one = {{1, 0, 0, 0}}; fr_mul(r, a, &one); }        /* :589-593 */
static const ofr_t FR_ZERO_C = {{0, 0, 0, 0}};

void orc_fr_add(const ofr_t* a, const ofr_t* b, ofr_t* out, size_t n) { for (size_t i = 0; i < n; ++i) fr_add(out + i, a + i, b + i); }
void orc_fr_sub(const ofr_t* a, const ofr_t* b, ofr_t* out, size_t n) { for (size_t i = 0; i < n; ++i) fr_sub(out + i, a + i, b + i); }
void orc_fr_mul(const ofr_t* a, const ofr_t* b, ofr_t* out, size_t n) { for (size_t i = 0; i < n; ++i) fr_mul(out + i, a + i, b + i); }
void orc_fr_mont(const ofr_t* a, ofr_t* out, size_t n) { for (size_t i = 0; i < n; ++i) fr_mont(out + i, a + i); }
void orc_fr_unmont(const ofr_t* a, ofr_t* out, size_t n) { for (size_t i = 0; i < n; ++i) fr_unmont(out + i, a + i); }
void orc_fr_neg(const ofr_t* a, ofr_t* out, size_t n) { for (size_t i = 0; i < n; ++i) fr_sub(out + i, &FR_ZERO_C, a + i); } /* fr-tensor.cu:36-41 */
void orc_fr_bcast(const ofr_t* a, const ofr_t* x, int op, ofr_t* out, size_t n) {   /* fr-tensor.cu:28-33,51-56,82-87 */
  for (size_t i = 0; i < n; ++i) {
    if (op == 0) fr_add(out + i, a + i, x); else if (op == 1) fr_sub(out + i, a + i, x); else fr_mul(out + i, a + i, x);
  }
}
void orc_fr_sum(const ofr_t* a, size_t n, ofr_t* out) {                              /* fr-tensor.cu:240-292: sum mod p */
  ofr_t s = FR_ZERO_C;
  for (size_t i = 0; i < n; ++i) fr_add(&s, &s, a + i);
  *out = s;
}

/* fold(T,x)[g] = T[2g] + x (T[2g+1] - T[2g]); missing entries are 0 (fr-tensor.cu:399-409) */
static void fr_me_step(const ofr_t* in, ofr_t* out, const ofr_t* x, size_t in_size, size_t out_size) {
#pragma omp parallel for schedule(static) if (out_size > 4096)
  for (size_t g = 0; g < out_size; ++g) {
    size_t g0 = 2 * g, g1 = 2 * g + 1; ofr_t t;
    if (g1 < in_size) { fr_sub(&t, in + g1, in + g0); fr_mul(&t, x, &t); fr_add(out + g, in + g0, &t); }
    else if (g0 < in_size) { fr_mul(&t, x, in + g0); fr_sub(out + g, in + g0, &t); }
    else out[g] = FR_ZERO_C;
  }
}
void orc_fr_me(const ofr_t* a, size_t n, const ofr_t* u, size_t k, ofr_t* out) {     /* fr-tensor.cu:411-418 */
  ofr_t* cur = (ofr_t*)malloc(sizeof(ofr_t) * (n ? n : 1));
  memcpy(cur, a, sizeof(ofr_t) * n);
  size_t sz = n;
  for (size_t j = 0; j < k; ++j) {
    size_t o = (sz + 1) / 2;
    ofr_t* nx = (ofr_t*)malloc(sizeof(ofr_t) * (o ? o : 1));
    fr_me_step(cur, nx, u + j, sz, o);
    free(cur); cur = nx; sz = o;
  }
  *out = cur[0];
  free(cur);
}
size_t orc_fr_partial_me(const ofr_t* a, size_t n, const ofr_t* u, size_t k, size_t w, ofr_t* out) { /* fr-tensor.cu:420-443 */
  ofr_t* cur = (ofr_t*)malloc(sizeof(ofr_t) * (n ? n : 1));
  memcpy(cur, a, sizeof(ofr_t) * n);
  size_t sz = n;
  for (size_t j = 0; j < k; ++j) {
    size_t nw = (sz + 2 * w - 1) / (2 * w), o = w * nw;
    ofr_t* nx = (ofr_t*)malloc(sizeof(ofr_t) * (o ? o : 1));
#pragma omp parallel for schedule(static) if (o > 4096)
    for (size_t g = 0; g < o; ++g) {
      size_t wid = g / w, idx = g % w, g0 = 2 * wid * w + idx, g1 = (2 * wid + 1) * w + idx; ofr_t t;
      if (g1 < sz) { fr_sub(&t, cur + g1, cur + g0); fr_mul(&t, u + j, &t); fr_add(nx + g, cur + g0, &t); }
      else if (g0 < sz) { fr_mul(&t, u + j, cur + g0); fr_sub(nx + g, cur + g0, &t); }
      else nx[g] = FR_ZERO_C;
    }
    free(cur); cur = nx; sz = o;
  }
  memcpy(out, cur, sizeof(ofr_t) * sz);
  free(cur);
  return sz;
}

/* per-pair coefficient vectors (proof.cu:55-70) */
static void ip_step(const ofr_t* a, const ofr_t* b, ofr_t* o0, ofr_t* o1, ofr_t* o2, size_t in_size, size_t out_size) {
#pragma omp parallel for schedule(static) if (out_size > 4096)
  for (size_t g = 0; g < out_size; ++g) {
    size_t g0 = 2 * g, g1 = 2 * g + 1;
    ofr_t a0 = g0 < in_size ? a[g0] : FR_ZERO_C, b0 = g0 < in_size ? b[g0] : FR_ZERO_C;
    ofr_t a1 = g1 < in_size ? a[g1] : FR_ZERO_C, b1 = g1 < in_size ? b[g1] : FR_ZERO_C;
    ofr_t da, db, t1, t2;
    fr_sub(&da, &a1, &a0); fr_sub(&db, &b1, &b0);
    fr_mul(o0 + g, &a0, &b0);
    fr_mul(&t1, &a0, &db); fr_mul(&t2, &b0, &da); fr_add(o1 + g, &t1, &t2);
    fr_mul(o2 + g, &da, &db);
  }
}
static void bin_step(const ofr_t* a, ofr_t* o0, ofr_t* o1, ofr_t* o2, size_t in_size, size_t out_size) { /* proof.cu:152-163 */
#pragma omp parallel for schedule(static) if (out_size > 4096)
  for (size_t g = 0; g < out_size; ++g) {
    ofr_t a0 = 2 * g < in_size ? a[2 * g] : FR_ZERO_C, a1 = 2 * g + 1 < in_size ? a[2 * g + 1] : FR_ZERO_C;
    ofr_t t, d, d2;
    fr_mul(&t, &a0, &a0); fr_sub(o0 + g, &t, &a0);
    fr_sub(&d, &a1, &a0);
    fr_dbl(&d2, &a0); fr_mul(&t, &d2, &d); fr_sub(o1 + g, &t, &d);
    fr_mul(o2 + g, &d, &d);
  }
}

void orc_ip_sumcheck(const ofr_t* a, const ofr_t* b, size_t n, const ofr_t* u, size_t k, ofr_t* proof) { /* proof.cu:72-108 */
  size_t cap = n ? n : 1;
  ofr_t* ca = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* cb = (ofr_t*)malloc(sizeof(ofr_t) * cap);
  ofr_t* o0 = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* o1 = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* o2 = (ofr_t*)malloc(sizeof(ofr_t) * cap);
  ofr_t* na = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* nb = (ofr_t*)malloc(sizeof(ofr_t) * cap);
  memcpy(ca, a, sizeof(ofr_t) * n); memcpy(cb, b, sizeof(ofr_t) * n);
  size_t sz = n, pi = 0;
  for (size_t j = 0; j < k; ++j) {
    size_t o = (sz + 1) / 2;
    ip_step(ca, cb, o0, o1, o2, sz, o);
    orc_fr_sum(o0, o, proof + pi++); orc_fr_sum(o1, o, proof + pi++); orc_fr_sum(o2, o, proof + pi++);
    fr_me_step(ca, na, u + j, sz, o); fr_me_step(cb, nb, u + j, sz, o);
    ofr_t* t = ca; ca = na; na = t; t = cb; cb = nb; nb = t; sz = o;
  }
  proof[pi++] = ca[0]; proof[pi++] = cb[0];
  free(ca); free(cb); free(o0); free(o1); free(o2); free(na); free(nb);
}
void orc_hp_sumcheck(const ofr_t* a, const ofr_t* b, size_t n, const ofr_t* u, const ofr_t* v, size_t k, ofr_t* proof) { /* proof.cu:110-150 */
  size_t cap = n ? n : 1;
  ofr_t* ca = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* cb = (ofr_t*)malloc(sizeof(ofr_t) * cap);
  ofr_t* o0 = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* o1 = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* o2 = (ofr_t*)malloc(sizeof(ofr_t) * cap);
  ofr_t* na = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* nb = (ofr_t*)malloc(sizeof(ofr_t) * cap);
  memcpy(ca, a, sizeof(ofr_t) * n); memcpy(cb, b, sizeof(ofr_t) * n);
  size_t sz = n, pi = 0;
  for (size_t j = 0; j < k; ++j) {
    size_t o = (sz + 1) / 2;
    ip_step(ca, cb, o0, o1, o2, sz, o);
    orc_fr_me(o0, o, u + j + 1, k - j - 1, proof + pi++);
    orc_fr_me(o1, o, u + j + 1, k - j - 1, proof + pi++);
    orc_fr_me(o2, o, u + j + 1, k - j - 1, proof + pi++);
    fr_me_step(ca, na, v + j, sz, o); fr_me_step(cb, nb, v + j, sz, o);
    ofr_t* t = ca; ca = na; na = t; t = cb; cb = nb; nb = t; sz = o;
  }
  proof[pi++] = ca[0]; proof[pi++] = cb[0];
  free(ca); free(cb); free(o0); free(o1); free(o2); free(na); free(nb);
}
void orc_bin_sumcheck(const ofr_t* a, size_t n, const ofr_t* u, const ofr_t* v, size_t k, ofr_t* proof) { /* proof.cu:165-200 */
  size_t cap = n ? n : 1;
  ofr_t* ca = (ofr_t*)malloc(sizeof(ofr_t) * cap);
  ofr_t* o0 = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* o1 = (ofr_t*)malloc(sizeof(ofr_t) * cap); ofr_t* o2 = (ofr_t*)malloc(sizeof(ofr_t) * cap);
  ofr_t* na = (ofr_t*)malloc(sizeof(ofr_t) * cap);
  memcpy(ca, a, sizeof(ofr_t) * n);
  size_t sz = n, pi = 0;
  for (size_t j = 0; j < k; ++j) {
    size_t o = (sz + 1) / 2;
    bin_step(ca, o0, o1, o2, sz, o);
    orc_fr_me(o0, o, u + j + 1, k - j - 1, proof + pi++);
    orc_fr_me(o1, o, u + j + 1, k - j - 1, proof + pi++);
    orc_fr_me(o2, o, u + j + 1, k - j - 1, proof + pi++);
    fr_me_step(ca, na, v + j, sz, o);
    ofr_t* t = ca; ca = na; na = t; sz = o;
  }
  proof[pi++] = ca[0];
  free(ca); free(o0); free(o1); free(o2); free(na);
}

/* ---- mt19937 + random_vec (proof.cu:3-11): libstdc++ uniform_int_distribution<unsigned>(0,UINT_MAX) on
 * mt19937 returns the raw 32-bit draw; braced-init evaluates left to right ---- */
typedef struct { uint32_t mt[624]; int idx; } mt_t;
static void mt_seed(mt_t* s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; ++i) s->mt[i] = 1812433253U * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->idx = 624;
}
static uint32_t mt_next(mt_t* s) {
  if (s->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (s->mt[i] & 0x80000000U) | (s->mt[(i + 1) % 624] & 0x7fffffffU);
      s->mt[i] = s->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfU : 0);
    }
    s->idx = 0;
  }
  uint32_t y = s->mt[s->idx++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680U; y ^= (y << 15) & 0xefc60000U; y ^= y >> 18;
  return y;
}
void orc_random_vec(uint32_t seed, size_t len, ofr_t* out) {
  mt_t s; mt_seed(&s, seed);
  for (size_t i = 0; i < len; ++i) {
    uint32_t w[8];
    for (int j = 0; j < 8; ++j) w[j] = mt_next(&s);
    w[7] %= 1944954707U;
    for (int j = 0; j < 4; ++j) out[i].l[j] = (uint64_t)w[2 * j] | ((uint64_t)w[2 * j + 1] << 32);
  }
}
uint32_t orc_ceil_log2(uint32_t num) {                                             /* proof.cu:13-31 */
  if (num == 0) return 0;
  num--; uint32_t r = 0;
  while (num > 0) { num >>= 1; r++; }
  return r;
}

/* ---- quantise / matmul / relu ---- */
static ofr_t float_to_fr(float x) {                                                 /* zkfc.cu:63-78 */
  x = x * 65536.0f;
  float ax = roundf(fabsf(x));
  int neg = signbit(x) ? 1 : 0;                                                      /* copysign(1,x) < 0 */
  uint32_t v;
  if (isnan(ax)) v = 0; else if (ax >= 4294967296.0f) v = 0xffffffffU; else v = (uint32_t)ax;  /* cvt.rzi.u32.f32 saturates */
  ofr_t r = {{v, 0, 0, 0}};
  if (neg) fr_sub(&r, &FR_ZERO_C, &r);
  return r;
}
void orc_float_to_fr(const float* fs, ofr_t* frs, uint32_t fs_rows, uint32_t frs_rows, uint32_t fs_cols, uint32_t frs_cols) { /* zkfc.cu:80-88 */
  for (uint32_t r = 0; r < frs_rows; ++r)
    for (uint32_t c = 0; c < frs_cols; ++c)
      frs[(size_t)r * frs_cols + c] = (r < fs_rows && c < fs_cols) ? float_to_fr(fs[(size_t)r * fs_cols + c]) : FR_ZERO_C;
}
void orc_fr_matmul(const ofr_t* A, const ofr_t* B, ofr_t* C, size_t rowsA, size_t colsA, size_t colsB) { /* zkfc.cu:6-47 */
#pragma omp parallel for schedule(static)
  for (size_t r = 0; r < rowsA; ++r)
    for (size_t c = 0; c < colsB; ++c) {
      ofr_t s = FR_ZERO_C, t;
      for (size_t k = 0; k < colsA; ++k) { fr_mul(&t, A + r * colsA + k, B + k * colsB + c); fr_add(&s, &s, &t); }
      C[r * colsB + c] = s;
    }
}
size_t orc_relu(const ofr_t* X, ofr_t* Z, ofr_t* sign, ofr_t* mag_bin, ofr_t* rem_bin, size_t n) { /* zkrelu.cu:11-41 */
  static const uint64_t POS_MAX[4] = {0x00007fffffffffffULL, 0, 0, 0};                /* {4294967295,32767,0..} = 2^47-1 */
  uint64_t NEG_MIN[4]; { uint64_t t[4] = {1ULL << 47, 0, 0, 0}; sub_n(NEG_MIN, FR_P, t, 4); }  /* p - 2^47 (zkrelu.cu:23) */
  ofr_t one; memcpy(one.l, FR_ONE, 32);
  size_t bad = 0;
  for (size_t i = 0; i < n; ++i) {
    ofr_t x; fr_unmont(&x, X + i);
    uint64_t mag = 0; ofr_t sg = FR_ZERO_C;
    if (ge_n(POS_MAX, x.l, 4)) { sg = one; mag = x.l[0]; }
    else if (ge_n(x.l, NEG_MIN, 4)) { ofr_t t = {{1ULL << 47, 0, 0, 0}}, s; fr_add(&s, &x, &t); mag = s.l[0]; }
    else { bad++; }                                   /* reference: uninitialised (App. B9); here: sign=0, mag=0 */
    sign[i] = sg;
    int rem_sign = (mag & 32768ULL) != 0;
    uint32_t rem_mag = (uint32_t)(mag & 32767ULL);
    int rem = rem_sign ? ((int)rem_mag - (1 << 15)) : (int)rem_mag;
    uint32_t q = (uint32_t)((mag - (uint64_t)(int64_t)rem) >> 16);
    for (int k = 0; k < 32; ++k) mag_bin[i * 32 + k] = ((q >> k) & 1) ? one : FR_ZERO_C;
    for (int k = 0; k < 15; ++k) rem_bin[i * 16 + k] = ((rem_mag >> k) & 1) ? one : FR_ZERO_C;
    rem_bin[i * 16 + 15] = rem_sign ? one : FR_ZERO_C;
    ofr_t qf = {{q, 0, 0, 0}}, qm; fr_mont(&qm, &qf); fr_mul(Z + i, &qm, &sg);
  }
  return bad;
}

/* ---- Fq ---- */
static inline void fq_add(ofq_t* r, const ofq_t* a, const ofq_t* b) { modadd_n(r->l, a->l, b->l, FQ_P, 6); }
static inline void fq_sub(ofq_t* r, const ofq_t* a, const ofq_t* b) { modsub_n(r->l, a->l, b->l, FQ_P, 6); }
static inline void fq_mul(ofq_t* r, const ofq_t* a, const ofq_t* b) { montmul_n(r->l, a->l, b->l, FQ_P, FQ_INV, 6); }
static inline void fq_sqr(ofq_t* r, const ofq_t* a) { fq_mul(r, a, a); }          /* bls12-381.cu:951-953 */
static inline void fq_dbl(ofq_t* r, const ofq_t* a) { fq_add(r, a, a); }
static inline int fq_is_zero(const ofq_t* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3] | a->l[4] | a->l[5]) == 0; }
static inline int fq_eq(const ofq_t* a, const ofq_t* b) { return memcmp(a, b, sizeof(ofq_t)) == 0; }
static const ofq_t FQ_ZERO_C = {{0, 0, 0, 0, 0, 0}};
void orc_fq_mul(const ofq_t* a, const ofq_t* b, ofq_t* out, size_t n) { for (size_t i = 0; i < n; ++i) fq_mul(out + i, a + i, b + i); }

static og1j_t g1_zero(void) { og1j_t z; z.x = FQ_ZERO_C; memcpy(z.y.l, FQ_ONE, 48); z.z = FQ_ZERO_C; return z; } /* .cuh:419 */

static og1j_t g1_double(og1j_t p) {                                                 /* bls12-381.cu:1332-1359 dbl-2009-l */
  if (fq_is_zero(&p.z)) return p;
  ofq_t a, b, c, d, e, f, t;
  fq_sqr(&a, &p.x); fq_sqr(&b, &p.y); fq_sqr(&c, &b);
  fq_add(&d, &p.x, &b); fq_sqr(&d, &d); fq_sub(&d, &d, &a); fq_sub(&d, &d, &c); fq_dbl(&d, &d);
  fq_dbl(&e, &a); fq_add(&e, &e, &a);
  fq_sqr(&f, &e);
  fq_mul(&p.z, &p.y, &p.z); fq_dbl(&p.z, &p.z);
  fq_sub(&p.x, &f, &d); fq_sub(&p.x, &p.x, &d);
  fq_dbl(&c, &c); fq_dbl(&c, &c); fq_dbl(&c, &c);
  fq_sub(&t, &d, &p.x); fq_mul(&t, &t, &e); fq_sub(&p.y, &t, &c);
  return p;
}
static og1j_t g1_add_mixed(og1j_t a, const og1a_t* b) {                             /* :1362-1400 madd-2007-bl */
  if (fq_is_zero(&a.z)) { a.x = b->x; a.y = b->y; memcpy(a.z.l, FQ_ONE, 48); return a; }
  ofq_t z1z1, u2, s2, h, hh, i, j, r, v, t; og1j_t ret;
  fq_sqr(&z1z1, &a.z); fq_mul(&u2, &b->x, &z1z1);
  fq_mul(&s2, &b->y, &a.z); fq_mul(&s2, &s2, &z1z1);
  if (fq_eq(&a.x, &u2) && fq_eq(&a.y, &s2)) return g1_double(a);
  fq_sub(&h, &u2, &a.x); fq_sqr(&hh, &h);
  fq_dbl(&i, &hh); fq_dbl(&i, &i);
  fq_mul(&j, &h, &i);
  fq_sub(&r, &s2, &a.y); fq_dbl(&r, &r);
  fq_mul(&v, &a.x, &i);
  fq_sqr(&t, &r); fq_sub(&t, &t, &j); { ofq_t v2; fq_dbl(&v2, &v); fq_sub(&ret.x, &t, &v2); }
  fq_mul(&j, &a.y, &j); fq_dbl(&j, &j);
  fq_sub(&t, &v, &ret.x); fq_mul(&t, &t, &r); fq_sub(&ret.y, &t, &j);
  fq_add(&t, &a.z, &h); fq_sqr(&t, &t); fq_sub(&t, &t, &z1z1); fq_sub(&ret.z, &t, &hh);
  return ret;
}
static og1j_t g1_add(og1j_t a, og1j_t b) {                                          /* :1403-1435 add-2007-bl */
  if (fq_is_zero(&a.z)) return b;
  if (fq_is_zero(&b.z)) return a;
  ofq_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
  fq_sqr(&z1z1, &a.z); fq_sqr(&z2z2, &b.z);
  fq_mul(&u1, &a.x, &z2z2); fq_mul(&u2, &b.x, &z1z1);
  fq_mul(&s1, &a.y, &b.z); fq_mul(&s1, &s1, &z2z2);
  fq_mul(&s2, &b.y, &a.z); fq_mul(&s2, &s2, &z1z1);
  if (fq_eq(&u1, &u2) && fq_eq(&s1, &s2)) return g1_double(a);
  fq_sub(&h, &u2, &u1);
  fq_dbl(&i, &h); fq_sqr(&i, &i);
  fq_mul(&j, &h, &i);
  fq_sub(&r, &s2, &s1); fq_dbl(&r, &r);
  fq_mul(&v, &u1, &i);
  fq_sqr(&t, &r); fq_sub(&t, &t, &j); fq_sub(&t, &t, &v); fq_sub(&a.x, &t, &v);
  fq_sub(&t, &v, &a.x); fq_mul(&a.y, &t, &r);
  fq_mul(&s1, &s1, &j); fq_dbl(&s1, &s1);
  fq_sub(&a.y, &a.y, &s1);
  fq_add(&t, &a.z, &b.z); fq_sqr(&t, &t); fq_sub(&t, &t, &z1z1); fq_sub(&t, &t, &z2z2);
  fq_mul(&a.z, &t, &h);
  return a;
}
static og1j_t g1_neg(og1j_t a) { fq_sub(&a.y, &FQ_ZERO_C, &a.y); return a; }         /* g1-tensor.cu:9-19 */

static og1j_t g1_mul(og1j_t a, const ofr_t* x) {                                     /* g1-tensor.cu:422-430 */
  og1j_t out = g1_zero();
  for (int i = 0; i < 256; ++i) {
    if ((x->l[i / 64] >> (i % 64)) & 1) out = g1_add(out, a);
    a = g1_double(a);
  }
  return out;
}
static og1j_t g1_mul_fast(og1j_t a, const ofr_t* x) {          /* same group element; 4-bit fixed window, MSB first */
  og1j_t tab[16]; tab[0] = g1_zero(); tab[1] = a;
  for (int i = 2; i < 16; ++i) tab[i] = (i & 1) ? g1_add(tab[i - 1], a) : g1_double(tab[i / 2]);
  og1j_t out = g1_zero(); int started = 0;
  for (int w = 63; w >= 0; --w) {
    unsigned d = (unsigned)((x->l[w / 16] >> ((w % 16) * 4)) & 15);
    if (started) { out = g1_double(out); out = g1_double(out); out = g1_double(out); out = g1_double(out); }
    if (d) { out = g1_add(out, tab[d]); started = 1; }
  }
  return out;
}

void orc_g1_double(const og1j_t* a, og1j_t* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = g1_double(a[i]); }
void orc_g1_add(const og1j_t* a, const og1j_t* b, og1j_t* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = g1_add(a[i], b[i]); }
void orc_g1_add_mixed(const og1j_t* a, const og1a_t* b, og1j_t* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = g1_add_mixed(a[i], b + i); }
void orc_g1_neg(const og1j_t* a, og1j_t* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = g1_neg(a[i]); }
void orc_g1_mul(const og1j_t* P, size_t np, const ofr_t* x, size_t n, og1j_t* out) {
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t i = 0; i < n; ++i) out[i] = g1_mul(P[i % np], x + i);
}
void orc_g1_mul_fast(const og1j_t* P, size_t np, const ofr_t* x, size_t n, og1j_t* out) {
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t i = 0; i < n; ++i) out[i] = g1_mul_fast(P[i % np], x + i);
}
/* G1TensorJacobian::sum with the reference's exact launch structure (g1-tensor.cu:368-420): blocks of 64 threads,
 * each block covers 128 inputs, grid = ceil(cur/64) so the upper half of the blocks emit ZERO. */
void orc_g1_sum(const og1j_t* a, size_t n, og1j_t* out) {
  if (n == 0) { *out = g1_zero(); return; }
  size_t cap = n + 64;
  og1j_t* in = (og1j_t*)malloc(sizeof(og1j_t) * cap);
  og1j_t* ou = (og1j_t*)malloc(sizeof(og1j_t) * cap);
  memcpy(in, a, sizeof(og1j_t) * n);
  size_t cur = n;
  while (cur > 1) {
    size_t grid = (cur + 63) / 64;
    for (size_t b = 0; b < grid; ++b) {
      og1j_t s[64];
      for (size_t tid = 0; tid < 64; ++tid) {
        size_t i = b * 128 + tid;
        s[tid] = i < cur ? in[i] : g1_zero();
        if (i + 64 < cur) s[tid] = g1_add(s[tid], in[i + 64]);
      }
      for (size_t st = 32; st > 0; st >>= 1) for (size_t tid = 0; tid < st; ++tid) s[tid] = g1_add(s[tid], s[tid + st]);
      ou[b] = s[0];
    }
    og1j_t* t = in; in = ou; ou = t;
    cur = grid;
  }
  *out = in[0];
  free(in); free(ou);
}
void orc_g1_me(const og1j_t* a, size_t n, const ofr_t* u, size_t k, og1j_t* out) {     /* g1-tensor.cu:463-484 */
  og1j_t* cur = (og1j_t*)malloc(sizeof(og1j_t) * (n ? n : 1));
  memcpy(cur, a, sizeof(og1j_t) * n);
  size_t sz = n;
  for (size_t j = 0; j < k; ++j) {
    size_t o = (sz + 1) / 2;
    og1j_t* nx = (og1j_t*)malloc(sizeof(og1j_t) * (o ? o : 1));
    ofr_t xu; fr_unmont(&xu, u + j);
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t g = 0; g < o; ++g) {
      size_t g0 = 2 * g, g1 = 2 * g + 1;
      if (g1 < sz) nx[g] = g1_add(cur[g0], g1_mul(g1_add(cur[g1], g1_neg(cur[g0])), &xu));
      else if (g0 < sz) nx[g] = g1_add(cur[g0], g1_neg(g1_mul(cur[g0], &xu)));
      else nx[g] = g1_zero();
    }
    free(cur); cur = nx; sz = o;
  }
  *out = cur[0];
  free(cur);
}

/* Fq inversion by Fermat (the reference has no inversion; used only to compare points as affine) */
static void fq_inv(ofq_t* r, const ofq_t* a) {
  uint64_t e[6]; uint64_t two[6] = {2, 0, 0, 0, 0, 0}; sub_n(e, FQ_P, two, 6);
  ofq_t acc; memcpy(acc.l, FQ_ONE, 48);
  for (int i = 383; i >= 0; --i) {
    fq_sqr(&acc, &acc);
    if ((e[i / 64] >> (i % 64)) & 1) fq_mul(&acc, &acc, a);
  }
  *r = acc;
}
void orc_g1_to_affine(const og1j_t* a, og1a_t* out, uint8_t* is_inf, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    if (fq_is_zero(&a[i].z)) { is_inf[i] = 1; out[i].x = FQ_ZERO_C; out[i].y = FQ_ZERO_C; continue; }
    is_inf[i] = 0;
    ofq_t zi, zi2, zi3; fq_inv(&zi, &a[i].z); fq_sqr(&zi2, &zi); fq_mul(&zi3, &zi2, &zi);
    fq_mul(&out[i].x, &a[i].x, &zi2); fq_mul(&out[i].y, &a[i].y, &zi3);
  }
}
int orc_g1_eq(const og1j_t* a, const og1j_t* b) {
  int ia = fq_is_zero(&a->z), ib = fq_is_zero(&b->z);
  if (ia || ib) return ia && ib;
  ofq_t z1z1, z2z2, u1, u2, s1, s2;
  fq_sqr(&z1z1, &a->z); fq_sqr(&z2z2, &b->z);
  fq_mul(&u1, &a->x, &z2z2); fq_mul(&u2, &b->x, &z1z1);
  fq_mul(&s1, &a->y, &b->z); fq_mul(&s1, &s1, &z2z2);
  fq_mul(&s2, &b->y, &a->z); fq_mul(&s2, &s2, &z1z1);
  return fq_eq(&u1, &u2) && fq_eq(&s1, &s2);
}
int orc_g1_on_curve(const og1j_t* a) {       /* Y^2 = X^3 + 4 Z^6 */
  if (fq_is_zero(&a->z)) return 1;
  ofq_t y2, x3, z2, z6, four, t; ofq_t one; memcpy(one.l, FQ_ONE, 48);
  fq_sqr(&y2, &a->y); fq_sqr(&x3, &a->x); fq_mul(&x3, &x3, &a->x);
  fq_sqr(&z2, &a->z); fq_sqr(&z6, &z2); fq_mul(&z6, &z6, &z2);
  fq_dbl(&four, &one); fq_dbl(&four, &four);
  fq_mul(&t, &four, &z6); fq_add(&t, &t, &x3);
  return fq_eq(&t, &y2);
}
void orc_g1_generator(og1j_t* out) { memcpy(out->x.l, G1_GEN_X, 48); memcpy(out->y.l, G1_GEN_Y, 48); memcpy(out->z.l, FQ_ONE, 48); }

/* ---- Commitment ---- */
void orc_commit(const og1j_t* G, size_t ng, const ofr_t* t, size_t nt, og1j_t* com, int fast) { /* commitment.cu:29-41 (intended) */
  size_t m = nt / ng;
  ofr_t* tu = (ofr_t*)malloc(sizeof(ofr_t) * nt);
  og1j_t* tmp = (og1j_t*)malloc(sizeof(og1j_t) * nt);
  orc_fr_unmont(t, tu, nt);
  if (fast) orc_g1_mul_fast(G, ng, tu, nt, tmp); else orc_g1_mul(G, ng, tu, nt, tmp);
  for (size_t r = 0; r < m; ++r) orc_g1_sum(tmp + r * ng, ng, com + r);
  free(tu); free(tmp);
}
void orc_commit_as_written(const og1j_t* G, size_t ng, const ofr_t* t, size_t nt, og1j_t* com) { /* commitment.cu:3-41 literal */
  size_t m = nt / ng, n = ng;
  ofr_t* tu = (ofr_t*)malloc(sizeof(ofr_t) * nt);
  og1j_t* tmp = (og1j_t*)malloc(sizeof(og1j_t) * nt);
  orc_fr_unmont(t, tu, nt);
  orc_g1_mul(G, ng, tu, nt, tmp);
  for (size_t r = 0; r < m; ++r) com[r] = g1_zero();
  for (size_t b = 0; b < (m + 63) / 64; ++b) {
    og1j_t s[64];
    for (size_t l = 0; l < 64; ++l) {
      size_t gid = b * 64 + l;
      if (gid >= m) { s[l] = g1_zero(); continue; }        /* reference: uninitialised smem (fact 5) */
      og1j_t sum = tmp[gid * n + l];
      for (size_t i = l + 64; i < n; i += 64) sum = g1_add(sum, tmp[gid * n + i]);
      s[l] = sum;
    }
    for (size_t st = 32; st > 0; st >>= 1) for (size_t l = 0; l < st; ++l) s[l] = g1_add(s[l], s[l + st]);
    com[b * 64] = s[0];
  }
  free(tu); free(tmp);
}
void orc_me_open(const ofr_t* t, const og1j_t* G, size_t n, const ofr_t* u, size_t k, og1j_t* proof, ofr_t* ret, int fast) { /* commitment.cu:43-81 */
  ofr_t* s = (ofr_t*)malloc(sizeof(ofr_t) * (n ? n : 1)); og1j_t* g = (og1j_t*)malloc(sizeof(og1j_t) * (n ? n : 1));
  memcpy(s, t, sizeof(ofr_t) * n); memcpy(g, G, sizeof(og1j_t) * n);
  size_t sz = n, pi = 0;
  og1j_t (*mulf)(og1j_t, const ofr_t*) = fast ? g1_mul_fast : g1_mul;
  for (size_t j = 0; j < k; ++j) {
    size_t ns = sz / 2;
    ofr_t* s2 = (ofr_t*)malloc(sizeof(ofr_t) * (ns ? ns : 1)); og1j_t* g2 = (og1j_t*)malloc(sizeof(og1j_t) * (ns ? ns : 1));
    og1j_t* T = (og1j_t*)malloc(sizeof(og1j_t) * (ns ? ns : 1)); og1j_t* T0 = (og1j_t*)malloc(sizeof(og1j_t) * (ns ? ns : 1));
    og1j_t* T1 = (og1j_t*)malloc(sizeof(og1j_t) * (ns ? ns : 1));
    ofr_t uu; fr_unmont(&uu, u + j);
#pragma omp parallel for schedule(dynamic, 2)
    for (size_t i = 0; i < ns; ++i) {
      size_t g0 = 2 * i, g1 = 2 * i + 1; ofr_t d;
      fr_sub(&d, s + g1, s + g0); fr_mul(&d, u + j, &d); fr_add(s2 + i, s + g0, &d);
      g2[i] = g1_add(g[g1], mulf(g1_add(g[g0], g1_neg(g[g1])), &uu));
      T[i] = g1_add(mulf(g[g0], s + g0), mulf(g[g1], s + g1));     /* raw Montgomery limbs as the scalar (defect B4) */
      T0[i] = mulf(g[g1], s + g0);
      T1[i] = mulf(g[g0], s + g1);
    }
    orc_g1_sum(T, ns, proof + pi++); orc_g1_sum(T0, ns, proof + pi++); orc_g1_sum(T1, ns, proof + pi++);
    free(s); free(g); free(T); free(T0); free(T1);
    s = s2; g = g2; sz = ns;
  }
  proof[pi++] = g[0];
  *ret = s[0];
  free(s); free(g);
}

/* ---- CPU baseline Pippenger (not a restatement of reference code: the reference has no CPU path) ---- */
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_msm_pippenger(const og1a_t* bases, const ofr_t* sc, size_t n, og1j_t* out, int threads) {
  int c = 4; while ((1ULL << (c + 4)) < n && c < 16) ++c;      /* ~ log2(n) - 4 */
  int W = (255 + c - 1) / c;
  size_t nb = (size_t)1 << c;
  og1j_t* wsum = (og1j_t*)malloc(sizeof(og1j_t) * W);
  (void)threads;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int w = 0; w < W; ++w) {
    og1j_t* bk = (og1j_t*)malloc(sizeof(og1j_t) * nb);
    for (size_t b = 0; b < nb; ++b) bk[b] = g1_zero();
    for (size_t i = 0; i < n; ++i) {
      int bit = w * c; uint64_t d = sc[i].l[bit / 64] >> (bit % 64);
      if (bit % 64 + c > 64 && bit / 64 + 1 < 4) d |= sc[i].l[bit / 64 + 1] << (64 - bit % 64);
      d &= nb - 1;
      if (d) bk[d] = g1_add_mixed(bk[d], bases + i);
    }
    og1j_t run = g1_zero(), acc = g1_zero();
    for (size_t b = nb - 1; b >= 1; --b) { run = g1_add(run, bk[b]); acc = g1_add(acc, run); }
    wsum[w] = acc;
    free(bk);
  }
  og1j_t r = g1_zero();
  for (int w = W - 1; w >= 0; --w) { for (int i = 0; i < c; ++i) r = g1_double(r); r = g1_add(r, wsum[w]); }
  *out = r;
  free(wsum);
}
