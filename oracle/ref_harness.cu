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

int main(int argc, char** argv) {
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
  fprintf(stderr, "usage: ref_harness run <in> <out> | time <fold|fc|msm|open|relu|layer> <log_n> [reps] [batch]\n");
  return 1;
}
