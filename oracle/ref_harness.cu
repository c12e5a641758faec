// ref_harness.cu — OUR driver around the UNMODIFIED reference (compiled from /root/reference by oracle/build_ref.sh,
// linked against the reference's own objects).  Test infrastructure only.
//
// It feeds fixed inputs / injected challenges / fixed generators through the reference's public C++ API
// (FrTensor, G1TensorJacobian, Commitment, zkFC, zkReLU and the proof.cuh free functions) and dumps every result, so
// that (a) tests/golden/*.bin are produced by the reference's own CUDA code on a B200 and (b) the GPU test-suite can
// diff the new kernels against the reference live.  It also provides the Baseline-A timing loops of BASELINE.md.
//
// The same source also compiles, unchanged, against the drop-in headers of zkdl_b200/host (-DZKDL_HOST_BUILD,
// -Izkdl_b200/host): that binary (zkdl_b200/host/zk_harness) is the API-level parity and timing twin of this one.
//
//   ref_harness run  <in.bin> <out.bin>     execute every case present in the input container
//   ref_harness time <what> <args...>       print one JSON line with reference timings
//
// Container format (little endian): "ZKH1", u32 count, then per entry: u32 name_len, name, u64 n_words, u32 words[].
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <map>
#include <string>
#include <vector>
#include <chrono>
#include <iostream>
#include <fstream>
#include <random>
#include <stdexcept>
#include <cuda_runtime.h>
#include <iomanip>
#include <utility>
#ifndef ZKDL_HOST_BUILD      // building against the reference: its headers need these (zkfc.cuh pulls in LibTorch)
#include <curand_kernel.h>
#include <torch/torch.h>
#include <torch/script.h>
#endif
// the reference keeps device pointers private; the harness needs bulk copies (reference files are not modified)
#define private public
#define protected public
#include "fr-tensor.cuh"
#include "g1-tensor.cuh"
#include "commitment.cuh"
#include "proof.cuh"
#include "zkfc.cuh"
#include "zkrelu.cuh"
#undef private
#undef protected

typedef std::map<std::string, std::vector<uint32_t>> Box;

static Box read_box(const char* path) {
  Box b; FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
  char magic[4]; uint32_t cnt;
  if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "ZKH1", 4) || fread(&cnt, 4, 1, f) != 1) { fprintf(stderr, "bad container\n"); exit(2); }
  for (uint32_t i = 0; i < cnt; ++i) {
    uint32_t nl; uint64_t nw;
    if (fread(&nl, 4, 1, f) != 1) exit(2);
    std::string name(nl, 0); if (fread(&name[0], 1, nl, f) != nl) exit(2);
    if (fread(&nw, 8, 1, f) != 1) exit(2);
    std::vector<uint32_t> w(nw); if (nw && fread(w.data(), 4, nw, f) != nw) exit(2);
    b[name] = std::move(w);
  }
  fclose(f); return b;
}
static void write_box(const char* path, const Box& b) {
  FILE* f = fopen(path, "wb"); if (!f) { perror(path); exit(2); }
  uint32_t cnt = (uint32_t)b.size(); fwrite("ZKH1", 1, 4, f); fwrite(&cnt, 4, 1, f);
  for (auto& kv : b) {
    uint32_t nl = (uint32_t)kv.first.size(); uint64_t nw = kv.second.size();
    fwrite(&nl, 4, 1, f); fwrite(kv.first.data(), 1, nl, f); fwrite(&nw, 8, 1, f);
    if (nw) fwrite(kv.second.data(), 4, nw, f);
  }
  fclose(f);
}
static bool has(const Box& b, const std::string& k) { return b.find(k) != b.end(); }

static FrTensor fr_in(const std::vector<uint32_t>& w) { return FrTensor((uint)(w.size() / 8), reinterpret_cast<const Fr_t*>(w.data())); }
static std::vector<Fr_t> fr_vec(const std::vector<uint32_t>& w) {
  std::vector<Fr_t> v(w.size() / 8); if (!v.empty()) memcpy(v.data(), w.data(), w.size() * 4); return v;
}
static std::vector<uint32_t> fr_out(const FrTensor& t) {
  std::vector<uint32_t> w((size_t)t.size * 8);
  cudaMemcpy(w.data(), t.gpu_data, w.size() * 4, cudaMemcpyDeviceToHost); return w;
}
static std::vector<uint32_t> fr_out(const std::vector<Fr_t>& v) {
  std::vector<uint32_t> w(v.size() * 8); if (!v.empty()) memcpy(w.data(), v.data(), w.size() * 4); return w;
}
static std::vector<uint32_t> fr_out(const Fr_t& x) { std::vector<uint32_t> w(8); memcpy(w.data(), &x, 32); return w; }
static G1TensorJacobian g1_in(const std::vector<uint32_t>& w) { return G1TensorJacobian((uint)(w.size() / 36), reinterpret_cast<const G1Jacobian_t*>(w.data())); }
static std::vector<uint32_t> g1_out(const G1TensorJacobian& t) {
  std::vector<uint32_t> w((size_t)t.size * 36);
  cudaMemcpy(w.data(), t.gpu_data, w.size() * 4, cudaMemcpyDeviceToHost); return w;
}
static std::vector<uint32_t> g1_out(const std::vector<G1Jacobian_t>& v) {
  std::vector<uint32_t> w(v.size() * 36); if (!v.empty()) memcpy(w.data(), v.data(), w.size() * 4); return w;
}
static std::vector<uint32_t> g1_out(const G1Jacobian_t& x) { std::vector<uint32_t> w(36); memcpy(w.data(), &x, 144); return w; }

// ------------------------------------------------------------------------------------------------ run
static void run_cases(const Box& in, Box& out) {
  // ---- Fr elementwise / sum (fr-tensor.cu)
  if (has(in, "ops.a")) {
    FrTensor a = fr_in(in.at("ops.a")), b = fr_in(in.at("ops.b"));
    Fr_t x = fr_vec(in.at("ops.x"))[0];
    out["ops.add"] = fr_out(a + b); out["ops.sub"] = fr_out(a - b); out["ops.mul"] = fr_out(a * b); out["ops.neg"] = fr_out(-a);
    { FrTensor t(a); out["ops.mont"] = fr_out(t.mont()); }
    { FrTensor t(a); out["ops.unmont"] = fr_out(t.unmont()); }
    out["ops.badd"] = fr_out(a + x); out["ops.bsub"] = fr_out(a - x); out["ops.bmul"] = fr_out(a * x);
    out["ops.sum"] = fr_out(a.sum());
  }
  // ---- folds (fr-tensor.cu:295-443)
  if (has(in, "fold.a")) {
    FrTensor a = fr_in(in.at("fold.a"));
    auto u = fr_vec(in.at("fold.u"));
    uint w = in.at("fold.w")[0], kk = in.at("fold.w")[1];
    out["fold.me"] = fr_out(a(u));
    out["fold.pm"] = fr_out(a.partial_me(std::vector<Fr_t>(u.begin(), u.begin() + kk), w));
  }
  // ---- sumchecks (proof.cu)
  if (has(in, "sc.a")) {
    FrTensor a = fr_in(in.at("sc.a")), b = fr_in(in.at("sc.b"));
    auto u = fr_vec(in.at("sc.u")), v = fr_vec(in.at("sc.v"));
    out["sc.ip"] = fr_out(inner_product_sumcheck(a, b, u));
    out["sc.hp"] = fr_out(hadamard_product_sumcheck(a, b, u, v));
    out["sc.bin"] = fr_out(binary_sumcheck(a, u, v));
  }
  // ---- G1 tensors (g1-tensor.cu)
  if (has(in, "g1.p")) {
    G1TensorJacobian p = g1_in(in.at("g1.p")), q = g1_in(in.at("g1.q"));
    FrTensor x = fr_in(in.at("g1.x"));
    auto u = fr_vec(in.at("g1.u"));
    out["g1.add"] = g1_out(p + q); out["g1.sub"] = g1_out(p - q); out["g1.neg"] = g1_out(-p);
    out["g1.mul"] = g1_out(p * x);
    out["g1.sum"] = g1_out(p.sum());
    out["g1.me"] = g1_out(p(u));
  }
  // ---- Commitment (commitment.cu)
  if (has(in, "com.g")) {
    G1TensorJacobian g0 = g1_in(in.at("com.g"));
    Commitment G(g0.size, reinterpret_cast<const G1Jacobian_t*>(in.at("com.g").data()));
    FrTensor t = fr_in(in.at("com.t"));
    out["com.as_written"] = g1_out(G.commit(t));                        // sum_axis_n_optimized as is
    uint m = t.size / G.size;
    std::vector<G1Jacobian_t> rows;
    for (uint r = 0; r < m; ++r) {                                      // intended: (G * unmont(row)).sum()
      std::vector<uint32_t> roww(in.at("com.t").begin() + (size_t)r * G.size * 8, in.at("com.t").begin() + (size_t)(r + 1) * G.size * 8);
      FrTensor row = fr_in(roww); row.unmont();
      rows.push_back((G * row).sum());
    }
    out["com.rows"] = g1_out(rows);
    if (has(in, "com.u")) {                                             // me_open on the first row-size scalars
      auto u = fr_vec(in.at("com.u"));
      std::vector<uint32_t> sw(in.at("com.s").begin(), in.at("com.s").end());
      FrTensor s = fr_in(sw);
      std::vector<G1Jacobian_t> proof;
      Fr_t ret = Commitment::me_open(s, G, u.begin(), u.end(), proof);
      out["com.open_proof"] = g1_out(proof); out["com.open_ret"] = fr_out(ret);
    }
    if (has(in, "com.uo")) {                                            // Commitment::open pieces (commitment.cu:83-92)
      auto u = fr_vec(in.at("com.uo"));
      G1TensorJacobian com((uint)rows.size(), rows.data());
      uint k = ceilLog2(com.size);
      const std::vector<Fr_t> u_out(u.end() - k, u.end()), u_in(u.begin(), u.end() - k);
      out["com.eval"] = g1_out(com(u_out));
      std::vector<G1Jacobian_t> proof;
      Fr_t ret = Commitment::me_open(t.partial_me(u_out, 1 << u_in.size()), G, u_in.begin(), u_in.end(), proof);
      out["com.full_proof"] = g1_out(proof); out["com.full_ret"] = fr_out(ret);
      out["com.open_api_ret"] = fr_out(G.open(t, com, u));
    }
  }
  // ---- zkFC forward + zkReLU (zkfc.cu, zkrelu.cu)
  if (has(in, "fc.w")) {
    uint B = in.at("fc.dims")[0], I = in.at("fc.dims")[1], O = in.at("fc.dims")[2];
    float *dw, *dx;
    cudaMalloc(&dw, sizeof(float) * I * O); cudaMalloc(&dx, sizeof(float) * B * I);
    cudaMemcpy(dw, in.at("fc.w").data(), sizeof(float) * I * O, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, in.at("fc.x").data(), sizeof(float) * B * I, cudaMemcpyHostToDevice);
    FrTensor Wq = zkFC::load_float_gpu_input(I, O, dw);                  // same float_to_Fr_kernel as from_float_gpu_ptr
    out["fc.wq"] = fr_out(Wq);
    FrTensor X = zkFC::load_float_gpu_input(B, I, dx);
    out["fc.xq"] = fr_out(X);
    uint ng = in.at("fc.dims")[3];
    Commitment G(ng, G1Jacobian_generator);
    zkFC fc = zkFC::from_float_gpu_ptr(I, O, dw, G);
    X.mont();
    FrTensor Z = fc(X);
    out["fc.z"] = fr_out(Z);
    zkReLU relu;
    FrTensor A = relu(Z);
    out["relu.a"] = fr_out(A); out["relu.sign"] = fr_out(*relu.sign_ptr);
    out["relu.mag"] = fr_out(*relu.mag_bin_ptr); out["relu.rem"] = fr_out(*relu.rem_bin_ptr);
    cudaFree(dw); cudaFree(dx);
  }
  out["cuda_status"] = std::vector<uint32_t>(1, (uint32_t)cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ time
static double now() { return std::chrono::duration<double>(std::chrono::high_resolution_clock::now().time_since_epoch()).count(); }
static FrTensor rand_small(uint n, uint bits, unsigned seed) {
  std::mt19937 mt(seed); std::vector<Fr_t> h(n);
  for (uint i = 0; i < n; ++i) { h[i] = {(uint)(mt() & ((1u << bits) - 1)), 0, 0, 0, 0, 0, 0, 0}; }
  FrTensor t(n, h.data()); t.mont(); return t;
}
static std::vector<Fr_t> seeded_vec(uint len, unsigned seed) {      // random_vec recipe with a fixed seed (proof.cu:3-11)
  std::mt19937 mt(seed); std::uniform_int_distribution<unsigned int> dist(0, UINT_MAX);
  std::vector<Fr_t> out(len);
  for (uint i = 0; i < len; ++i) out[i] = {dist(mt), dist(mt), dist(mt), dist(mt), dist(mt), dist(mt), dist(mt), dist(mt) % 1944954707};
  return out;
}

// ------------------------------------------------------------------------------------------------ full
// Full-size differential cases (BASELINE.json configs 2-5): inputs are regenerated from seeds on both sides (std::mt19937
// here, numpy's legacy MT19937 seeding in tests/test_fullsize_reference.py), so only the (small) results travel.
// Signed small integers: v = (mt() & (2^bits - 1)) - 2^(bits-1), Montgomery form.
static FrTensor rand_signed(uint n, uint bits, unsigned seed) {
  std::mt19937 mt(seed); std::vector<Fr_t> pos(n), neg(n);
  const uint mask = (1u << bits) - 1, half = 1u << (bits - 1);
  for (uint i = 0; i < n; ++i) {
    uint r = (uint)mt() & mask;
    pos[i] = {r >= half ? r - half : 0u, 0, 0, 0, 0, 0, 0, 0};
    neg[i] = {r >= half ? 0u : half - r, 0, 0, 0, 0, 0, 0, 0};
  }
  FrTensor P(n, pos.data()), N(n, neg.data());
  FrTensor t = P - N; t.mont(); return t;
}
static G1TensorJacobian seeded_points(uint n, unsigned seed) {      // [k_i] g, k_i = raw limbs of seeded_vec (demo.cu:81-82 pattern)
  G1TensorJacobian G(n, G1Jacobian_generator);
  auto ks = seeded_vec(n, seed); FrTensor kt(n, ks.data()); G *= kt; return G;
}
static FrTensor forward_product(const FrTensor& X, const FrTensor& W, uint B, uint I, uint O) {
#ifndef ZKDL_HOST_BUILD
  // the reference's own kernel (zkfc.cu:6-47), launched as zkFC::operator() does (zkfc.cu:117-126) without paying for
  // zkFC's constructor (it commits the weights with |W| bit-serial ladders)
  FrTensor out(B * O);
  dim3 blockSize(TILE_WIDTH, TILE_WIDTH), gridSize((O + TILE_WIDTH - 1) / TILE_WIDTH, (B + TILE_WIDTH - 1) / TILE_WIDTH);
  matrixMultiplyOptimized<<<gridSize, blockSize>>>(X.gpu_data, W.gpu_data, out.gpu_data, B, I, O);
  cudaDeviceSynchronize();
  return out;
#else
  Commitment G1c(64, G1Jacobian_generator);
  zkFC fc(I, O, W, G1c);
  return fc(X);
#endif
}
static std::vector<Fr_t> cat(const std::vector<Fr_t>& a, const std::vector<Fr_t>& b) { std::vector<Fr_t> r(a); r.insert(r.end(), b.begin(), b.end()); return r; }

static void full_fc(Box& out, uint I, uint O, uint B, bool with_open, uint wbits) {
  FrTensor W = rand_signed(I * O, wbits, 1), X = rand_signed(B * I, 17, 2);
  FrTensor Z = forward_product(X, W, B, I, O);
  uint kb = ceilLog2(B), ki = ceilLog2(I), ko = ceilLog2(O);
  auto u_bs = seeded_vec(kb, 101), u_in = seeded_vec(ki, 102), u_out = seeded_vec(ko, 103);
  FrTensor Xr = X.partial_me(u_bs, I), Wr = W.partial_me(u_out, 1);                // zkfc.cu:139
  out["fc.xr"] = fr_out(Xr); out["fc.wr"] = fr_out(Wr);
  out["fc.ip"] = fr_out(inner_product_sumcheck(Xr, Wr, u_in));
  out["fc.zu"] = fr_out(Z(cat(u_out, u_bs)));                                       // zkfc.cu:141-143
  out["fc.z_sum"] = fr_out(Z.sum());
  if (with_open) {
    uint ng = 1u << ((ceilLog2(I * O) + 1) / 2);                                    // demo.cu:81
    G1TensorJacobian g0 = seeded_points(ng, 7);
    std::vector<uint32_t> gw = g1_out(g0);
    Commitment G(ng, reinterpret_cast<const G1Jacobian_t*>(gw.data()));
    uint ncom = I * O / ng;
    G1TensorJacobian com = seeded_points(ncom, 9);                                  // any points: com(u_hi) does not depend on W
    auto u = cat(u_out, u_in);
    uint k = ceilLog2(ncom);
    const std::vector<Fr_t> u_hi(u.end() - k, u.end()), u_lo(u.begin(), u.end() - k);
    out["open.com_eval"] = g1_out(com(u_hi));                                       // commitment.cu:88
    FrTensor tf = W.partial_me(u_hi, 1u << u_lo.size());
    out["open.tf"] = fr_out(tf);
    std::vector<G1Jacobian_t> proof;
    Fr_t ret = Commitment::me_open(tf, G, u_lo.begin(), u_lo.end(), proof);         // commitment.cu:91
    out["open.proof"] = g1_out(proof); out["open.ret"] = fr_out(ret);
    std::vector<G1Jacobian_t> rows;                                                 // true row commitments of the first rows
    for (uint r = 0; r < 2; ++r) {
      std::vector<uint32_t> roww((size_t)ng * 8);
      cudaMemcpy(roww.data(), W.gpu_data + (size_t)r * ng, roww.size() * 4, cudaMemcpyDeviceToHost);
      FrTensor row = fr_in(roww); row.unmont();
      rows.push_back((G * row).sum());
    }
    out["open.com_rows"] = g1_out(rows);
  }
}
static void full_relu(Box& out, uint I, uint O, uint B, uint wbits) {
  FrTensor W = rand_signed(I * O, wbits, 1), X = rand_signed(B * I, 17, 2);
  FrTensor Z = forward_product(X, W, B, I, O);
  zkReLU relu; FrTensor A = relu(Z);
  uint L = ceilLog2(B * O);
  out["relu.a_me"] = fr_out(A(seeded_vec(L, 201)));
  out["relu.sign_me"] = fr_out((*relu.sign_ptr)(seeded_vec(L, 202)));
  out["relu.mag_me"] = fr_out((*relu.mag_bin_ptr)(seeded_vec(L + 5, 203)));
  out["relu.rem_me"] = fr_out((*relu.rem_bin_ptr)(seeded_vec(L + 4, 204)));
  auto u_z = seeded_vec(L + 5, 211), v_z = seeded_vec(L + 5, 212), u_r = seeded_vec(L + 4, 213), v_r = seeded_vec(L + 4, 214);
  auto u_rec = seeded_vec(L, 215), u_hp = seeded_vec(L, 216), v_hp = seeded_vec(L, 217);
  out["relu.mag_sc"] = fr_out(binary_sumcheck(*relu.mag_bin_ptr, u_z, v_z));        // zkrelu.cu:91-94
  out["relu.mag_rec"] = fr_out(relu.mag_bin_ptr->partial_me(u_rec, 32));
  out["relu.rem_sc"] = fr_out(binary_sumcheck(*relu.rem_bin_ptr, u_r, v_r));
  out["relu.rem_rec"] = fr_out(relu.rem_bin_ptr->partial_me(u_rec, 16));
  out["relu.hp"] = fr_out(hadamard_product_sumcheck(Z, *relu.sign_ptr, u_hp, v_hp));   // zkrelu.cu:99
}
static void full_msm(Box& out, uint k) {
  uint n = 1u << k;
  G1TensorJacobian G = seeded_points(n, 7);
  auto sv = seeded_vec(n, 8); FrTensor s(n, sv.data());
  out["msm.full"] = g1_out((G * s).sum());                                          // 255-bit scalars
  FrTensor w = rand_signed(n, 16, 3); w.unmont();                                   // the quantised-weight distribution
  out["msm.small"] = g1_out((G * w).sum());
}

int main(int argc, char** argv) {
  if (argc >= 6 && !strcmp(argv[1], "full")) {
    Box out; std::string what = argv[2];
    double t0 = now();
    if (what == "layer") {              // full <layer> <log_dim> <batch> <out.bin>: one hidden layer, zkFC + zkReLU pieces
      uint k = atoi(argv[3]), B = atoi(argv[4]);
      full_fc(out, 1u << k, 1u << k, B, true, 13);
      full_relu(out, 1u << k, 1u << k, B, 13);
    } else if (what == "fc") {          // full fc <log_dim> <batch> <out.bin>: config 2, the zkFC sumcheck set
      uint k = atoi(argv[3]), B = atoi(argv[4]);
      full_fc(out, 1u << k, 1u << k, B, false, 16);
    } else if (what == "msm") {         // full msm <log_n> 0 <out.bin>: config 3, (G * s).sum()
      full_msm(out, atoi(argv[3]));
    } else if (what == "split") {       // full split <log_n> <window> <out.bin>: FrTensor::split on a ragged table (fr-tensor.cu:376-397)
      uint n = (1u << atoi(argv[3])) + 5;
      FrTensor a = rand_signed(n, 20, 4);
      for (uint w : {(uint)atoi(argv[4]), 1u, 7u, n - 1}) {
        auto pr = a.split(w);
        out["split." + std::to_string(w) + ".first"] = fr_out(pr.first); out["split." + std::to_string(w) + ".second"] = fr_out(pr.second);
      }
    } else if (what == "random") {      // full random <n> <seed> <out.bin>: FrTensor::random / random_int's kernels with a fixed seed
      uint n = atoi(argv[3]); unsigned long seed = strtoul(argv[4], nullptr, 10);
      FrTensor a(n), b(n);
#ifndef ZKDL_HOST_BUILD
      random_kernel<<<(n + 255) / 256, 256>>>(a.gpu_data, n - 1, seed);                // `tid > n` guard (fr-tensor.cu:342): n - 1 writes exactly n elements
      random_int_kernel<<<(n + 255) / 256, 256>>>(b.gpu_data, 13, n, seed + 1);
      cudaDeviceSynchronize();
#else
      zkdl_fr_random(a.gpu_data, n, seed, 0); zkdl_fr_random_int(b.gpu_data, 13, n, seed + 1, 0);
      cudaDeviceSynchronize();
#endif
      out["random.fr"] = fr_out(a); out["random.int13"] = fr_out(b);
    } else { fprintf(stderr, "unknown full case\n"); return 1; }
    out["cuda_status"] = std::vector<uint32_t>(1, (uint32_t)cudaGetLastError());
    write_box(argv[5], out);
    printf("{\"what\":\"full_%s\",\"seconds\":%.3f,\"cuda_status\":%d}\n", what.c_str(), now() - t0, (int)out["cuda_status"][0]);
    return 0;
  }
  if (argc >= 4 && !strcmp(argv[1], "run")) {
    Box in = read_box(argv[2]), out;
    run_cases(in, out);
    write_box(argv[3], out);
    return 0;
  }
  if (argc >= 3 && !strcmp(argv[1], "time")) {
    std::string what = argv[2];
    int reps = argc > 4 ? atoi(argv[4]) : 3;
    if (what == "fold") {            // partial_me(W, window 1) + Fr_me at size 2^k: the M3 metric on the reference
      uint k = atoi(argv[3]); uint n = 1u << k;
      FrTensor W = rand_small(n, 16, 1);
      auto u = seeded_vec(k, 2);
      for (int r = 0; r < reps + 1; ++r) {
        cudaDeviceSynchronize(); double t0 = now();
        Fr_t v = W(u); (void)v;
        cudaDeviceSynchronize(); double t1 = now();
        if (r) printf("{\"what\":\"ref_fr_me\",\"log_n\":%u,\"seconds\":%.6f}\n", k, t1 - t0);
      }
    } else if (what == "fc") {       // config 2: zkFC sumcheck set, I=O=2^k, batch B (no opening)
      uint k = atoi(argv[3]); uint B = argc > 5 ? atoi(argv[5]) : 256; uint I = 1u << k, O = 1u << k;
      FrTensor W = rand_small(I * O, 16, 1), X = rand_small(B * I, 16, 2);
      auto u_bs = seeded_vec(ceilLog2(B), 3), u_in = seeded_vec(k, 4), u_out = seeded_vec(k, 5);
      for (int r = 0; r < reps + 1; ++r) {
        cudaDeviceSynchronize(); double t0 = now();
        auto Xr = X.partial_me(u_bs, I); auto Wr = W.partial_me(u_out, 1);
        auto p = inner_product_sumcheck(Xr, Wr, u_in);
        cudaDeviceSynchronize(); double t1 = now();
        if (r) printf("{\"what\":\"ref_fc_sumcheck\",\"log_dim\":%u,\"batch\":%u,\"seconds\":%.6f}\n", k, B, t1 - t0);
      }
    } else if (what == "msm") {      // (G * s).sum(): the reference's correct MSM composition
      uint k = atoi(argv[3]); uint n = 1u << k;
      Commitment G(n, G1Jacobian_generator);
      { auto ks = seeded_vec(n, 7); FrTensor kt(n, ks.data()); G *= kt; }
      auto sv = seeded_vec(n, 8); FrTensor s(n, sv.data());
      for (int r = 0; r < reps + 1; ++r) {
        cudaDeviceSynchronize(); double t0 = now();
        G1Jacobian_t v = (G * s).sum(); (void)v;
        cudaDeviceSynchronize(); double t1 = now();
        if (r) printf("{\"what\":\"ref_msm\",\"log_n\":%u,\"seconds\":%.6f,\"mpts_per_s\":%.6f}\n", k, t1 - t0, n / (t1 - t0) / 1e6);
      }
    } else if (what == "open") {     // Commitment::open on a 2^k x 2^k weight matrix (the #1 cost of the timed path)
      uint k = atoi(argv[3]); uint n = 1u << k;
      Commitment G(n, G1Jacobian_generator);
      { auto ks = seeded_vec(n, 7); FrTensor kt(n, ks.data()); G *= kt; }
      FrTensor W = rand_small(n * n, 13, 1);
      G1TensorJacobian com(n, G1Jacobian_generator);
      { auto ks = seeded_vec(n, 9); FrTensor kt(n, ks.data()); com *= kt; }
      auto u = seeded_vec(2 * k, 10);
      for (int r = 0; r < reps + 1; ++r) {
        cudaDeviceSynchronize(); double t0 = now();
        Fr_t v = G.open(W, com, u); (void)v;
        cudaDeviceSynchronize(); double t1 = now();
        if (r) printf("{\"what\":\"ref_open\",\"log_dim\":%u,\"seconds\":%.6f}\n", k, t1 - t0);
      }
    } else if (what == "layer") {    // one hidden layer of the demo MLP: zkReLU::prove + zkFC::prove, 2^k x 2^k weights, batch B
      uint k = atoi(argv[3]); uint B = argc > 5 ? atoi(argv[5]) : 256; uint I = 1u << k, O = 1u << k;
      uint ng = 1u << ((ceilLog2(I * O) + 1) / 2);
      Commitment G(ng, G1Jacobian_generator);
      { auto ks = seeded_vec(ng, 7); FrTensor kt(ng, ks.data()); G *= kt; }
      FrTensor Wt = rand_small(I * O, 13, 1);
      zkFC fc(I, O, Wt, G);                                            // commits the weights (untimed, as in demo.cu)
      FrTensor X = rand_small(B * I, 16, 2);
      FrTensor Z = fc(X);
      zkReLU relu; FrTensor A = relu(Z);
      for (int r = 0; r < reps + 1; ++r) {
        cudaDeviceSynchronize(); double t0 = now();
        relu.prove(Z, A);
        cudaDeviceSynchronize(); double t1 = now();
        fc.prove(X, Z, G);
        cudaDeviceSynchronize(); double t2 = now();
        if (r) printf("{\"what\":\"layer_prove\",\"log_dim\":%u,\"batch\":%u,\"relu_seconds\":%.6f,\"fc_seconds\":%.6f,\"seconds\":%.6f}\n", k, B, t1 - t0, t2 - t1, t2 - t0);
        fflush(stdout);
      }
    } else if (what == "relu") {     // zkReLU::prove at n = 2^k
      uint k = atoi(argv[3]); uint n = 1u << k;
      FrTensor Z = rand_small(n, 30, 3);
      zkReLU relu; FrTensor A = relu(Z);
      for (int r = 0; r < reps + 1; ++r) {
        cudaDeviceSynchronize(); double t0 = now();
        relu.prove(Z, A);
        cudaDeviceSynchronize(); double t1 = now();
        if (r) printf("{\"what\":\"ref_zkrelu_prove\",\"log_n\":%u,\"seconds\":%.6f}\n", k, t1 - t0);
      }
    }
    printf("{\"cuda_status\":%d}\n", (int)cudaGetLastError());
    return 0;
  }
  fprintf(stderr, "usage: ref_harness run <in> <out> | time <fold|fc|msm|open|relu|layer> <log_n> [reps] [batch] | full <layer|fc|msm> <log_n> <batch> <out>\n");
  return 1;
}
