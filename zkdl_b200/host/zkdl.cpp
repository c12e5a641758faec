// zkdl.cpp — implementation of the host-side API mirror (zkdl.hpp) over the C ABI.  Plain C++ (g++), no kernels.
#include "zkdl.hpp"
#include <atomic>
#include <climits>
#include <random>

using std::vector;
using namespace zkdl_host;

// ------------------------------------------------------------------------------------------------ constants
const Fp_t G1_generator_x_mont = {{4250078230u, 1555269520u, 2574712821u, 2014837863u, 339452353u, 357537223u, 4090554183u, 4037962445u, 568063040u, 3989728972u, 2651585397u, 302085953u}};
const Fp_t G1_generator_y_mont = {{216474225u, 3131872213u, 2031680910u, 2351063834u, 1460086222u, 3713621779u, 1346392468u, 1370249257u, 2902481344u, 236751935u, 1342743146u, 196886268u}};
const Fp_t G1_ONE = {{196605u, 1980301312u, 3289120770u, 3958636555u, 1405573306u, 1598593111u, 1884444485u, 2010011731u, 2723605613u, 1543969431u, 4202751123u, 368467651u}};
const G1Affine_t G1Affine_generator = {G1_generator_x_mont, G1_generator_y_mont};
const G1Jacobian_t G1Jacobian_generator = {G1_generator_x_mont, G1_generator_y_mont, G1_ONE};
const Fr_t Fr_ONE_mont = {{4294967294u, 1u, 215042u, 1485092858u, 3971764213u, 2576109551u, 2898593135u, 405057881u}};
const Fr_t Fr_ZERO = {{0, 0, 0, 0, 0, 0, 0, 0}};

// ------------------------------------------------------------------------------------------------ streams / memory
static thread_local cudaStream_t t_stream = 0;
cudaStream_t zkdl_host::cur_stream() { return t_stream; }
void zkdl_host::set_thread_stream(cudaStream_t s) { t_stream = s; }
void* zkdl_host::dev_alloc_bytes(size_t bytes) {
  void* p = nullptr;
  if (t_stream == 0) cuda_check(cudaMalloc(&p, bytes));                          // the reference's behaviour (fr-tensor.cu:92-95)
  else cuda_check(cudaMallocAsync(&p, bytes, t_stream));
  return p;
}
void zkdl_host::dev_free(void* p) {
  if (!p) return;
  if (t_stream == 0) cudaFree(p); else cudaFreeAsync(p, t_stream);
}
void zkdl_host::copy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  if (t_stream == 0) { cuda_check(cudaMemcpy(dst, src, bytes, kind)); return; }
  cuda_check(cudaMemcpyAsync(dst, src, bytes, kind, t_stream));
  cuda_check(cudaStreamSynchronize(t_stream));
}

static std::atomic<bool> g_async_prove{false};
void zkdl_host::set_async_prove(bool on) { g_async_prove = on; }
bool zkdl_host::async_prove() { return g_async_prove.load(); }

// ------------------------------------------------------------------------------------------------ challenge seeds
static std::atomic<uint32_t> g_seed_base{0}, g_seed_ctr{0};
static thread_local uint32_t t_seed_base = 0, t_seed_ctr = 0;
void set_challenge_seed(uint32_t seed) { g_seed_base = seed; g_seed_ctr = 0; }
void set_thread_challenge_seed(uint32_t seed) { t_seed_base = seed; t_seed_ctr = 0; }
uint32_t zkdl_host::next_challenge_seed() {
  if (t_seed_base) return t_seed_base + t_seed_ctr++;
  if (g_seed_base.load() == 0) { std::random_device rd; return rd(); }          // proof.cu:5-6
  return g_seed_base.load() + g_seed_ctr.fetch_add(1);
}

// ------------------------------------------------------------------------------------------------ printing
std::ostream& operator<<(std::ostream& os, const Fr_t& x) {                       // fr-tensor.cu:5-13
  os << "0x" << std::hex;
  for (uint i = 8; i > 0; --i) os << std::setfill('0') << std::setw(8) << x.val[i - 1];
  return os << std::dec << std::setw(0) << std::setfill(' ');
}
std::ostream& operator<<(std::ostream& os, const Fp_t& x) {                       // g1-tensor.cu:21-29
  os << "0x" << std::hex;
  for (uint i = 12; i > 0; --i) os << std::setfill('0') << std::setw(8) << x.val[i - 1];
  return os << std::dec << std::setw(0) << std::setfill(' ');
}
std::ostream& operator<<(std::ostream& os, const G1Affine_t& g) { return os << "(" << g.x << ", " << g.y << ")"; }
std::ostream& operator<<(std::ostream& os, const G1Jacobian_t& g) { return os << "(" << g.x << ", " << g.y << ", " << g.z << ")"; }
std::ostream& operator<<(std::ostream& os, const FrTensor& A) {                   // fr-tensor.cu:445-451 (one bulk copy instead of size copies)
  vector<Fr_t> h(A.size);
  copy(h.data(), A.gpu_data, sizeof(Fr_t) * A.size, cudaMemcpyDeviceToHost);
  os << '[';
  for (uint i = 0; i + 1 < A.size; ++i) os << h[i] << '\n';
  if (A.size) os << h[A.size - 1];
  return os << ']';
}

// ------------------------------------------------------------------------------------------------ FrTensor
FrTensor::FrTensor(uint size) : gpu_data(dev_alloc<Fr_t>(size)), size(size) {}
FrTensor::FrTensor(uint size, const Fr_t* cpu_data) : gpu_data(dev_alloc<Fr_t>(size)), size(size) {
  copy(gpu_data, cpu_data, sizeof(Fr_t) * size, cudaMemcpyHostToDevice);
}
FrTensor::FrTensor(const FrTensor& t) : gpu_data(dev_alloc<Fr_t>(t.size)), size(t.size) {
  copy(gpu_data, t.gpu_data, sizeof(Fr_t) * size, cudaMemcpyDeviceToDevice);
}
FrTensor::~FrTensor() { dev_free(gpu_data); gpu_data = nullptr; }
Fr_t FrTensor::operator()(uint idx) const {
  Fr_t out; copy(&out, gpu_data + idx, sizeof(Fr_t), cudaMemcpyDeviceToHost); return out;
}
static FrTensor fr_binary(int op, const FrTensor& a, const FrTensor& b) {
  if (a.size != b.size) throw std::runtime_error("Incompatible dimensions");
  FrTensor out(a.size); check(zkdl_fr_elementwise(op, a.gpu_data, b.gpu_data, out.gpu_data, a.size, st())); sync(); return out;
}
static FrTensor fr_bcast(int op, const FrTensor& a, const Fr_t& x) {
  FrTensor out(a.size); check(zkdl_fr_broadcast(op, a.gpu_data, &x, out.gpu_data, a.size, st())); sync(); return out;
}
FrTensor FrTensor::operator+(const FrTensor& t) const { return fr_binary(ZKDL_OP_ADD, *this, t); }
FrTensor FrTensor::operator+(const Fr_t& x) const { return fr_bcast(ZKDL_OP_ADD, *this, x); }
FrTensor& FrTensor::operator+=(const FrTensor& t) {
  if (size != t.size) throw std::runtime_error("Incompatible dimensions");
  check(zkdl_fr_elementwise(ZKDL_OP_ADD, gpu_data, t.gpu_data, gpu_data, size, st())); sync(); return *this;
}
FrTensor& FrTensor::operator+=(const Fr_t& x) { check(zkdl_fr_broadcast(ZKDL_OP_ADD, gpu_data, &x, gpu_data, size, st())); sync(); return *this; }
FrTensor FrTensor::operator-() const { FrTensor out(size); check(zkdl_fr_elementwise(ZKDL_OP_NEG, gpu_data, nullptr, out.gpu_data, size, st())); sync(); return out; }
FrTensor FrTensor::operator-(const FrTensor& t) const { return fr_binary(ZKDL_OP_SUB, *this, t); }
FrTensor FrTensor::operator-(const Fr_t& x) const { return fr_bcast(ZKDL_OP_SUB, *this, x); }
FrTensor& FrTensor::operator-=(const FrTensor& t) {
  if (size != t.size) throw std::runtime_error("Incompatible dimensions");
  check(zkdl_fr_elementwise(ZKDL_OP_SUB, gpu_data, t.gpu_data, gpu_data, size, st())); sync(); return *this;
}
FrTensor& FrTensor::operator-=(const Fr_t& x) { check(zkdl_fr_broadcast(ZKDL_OP_SUB, gpu_data, &x, gpu_data, size, st())); sync(); return *this; }
FrTensor& FrTensor::mont() { check(zkdl_fr_elementwise(ZKDL_OP_MONT, gpu_data, nullptr, gpu_data, size, st())); sync(); return *this; }
FrTensor& FrTensor::unmont() { check(zkdl_fr_elementwise(ZKDL_OP_UNMONT, gpu_data, nullptr, gpu_data, size, st())); sync(); return *this; }
FrTensor FrTensor::operator*(const FrTensor& t) const { return fr_binary(ZKDL_OP_MUL, *this, t); }
FrTensor FrTensor::operator*(const Fr_t& x) const { return fr_bcast(ZKDL_OP_MUL, *this, x); }
FrTensor& FrTensor::operator*=(const FrTensor& t) {
  if (size != t.size) throw std::runtime_error("Incompatible dimensions");
  check(zkdl_fr_elementwise(ZKDL_OP_MUL, gpu_data, t.gpu_data, gpu_data, size, st())); sync(); return *this;
}
FrTensor& FrTensor::operator*=(const Fr_t& x) { check(zkdl_fr_broadcast(ZKDL_OP_MUL, gpu_data, &x, gpu_data, size, st())); sync(); return *this; }
Fr_t FrTensor::sum() const {
  FrTensor out(1); check(zkdl_fr_sum(gpu_data, size, out.gpu_data, st())); sync(); return out(0);
}
Fr_t FrTensor::operator()(const vector<Fr_t>& u) const {
  FrTensor out(1); check(zkdl_fr_me(gpu_data, size, u.data(), u.size(), out.gpu_data, st())); sync(); return out(0);
}
Fr_t Fr_me(const FrTensor& t, vector<Fr_t>::const_iterator begin, vector<Fr_t>::const_iterator end) {
  // the reference's Fr_me applies end-begin folds without the size guard of operator()(u) (fr-tensor.cu:411-418)
  size_t k = end - begin;
  if (k == 0) return t(0);
  FrTensor out((uint)zkdl_partial_me_size(t.size, k, 1));
  check(zkdl_fr_partial_me(t.gpu_data, t.size, &*begin, k, 1, out.gpu_data, st())); sync();
  return out(0);
}
std::pair<FrTensor, FrTensor> FrTensor::split(uint window_size) const {           // fr-tensor.cu:376-397
  if (window_size < 1 || window_size >= size) throw std::runtime_error("Invalid window size.");
  uint out_size = (size + 1) / 2;
  std::pair<FrTensor, FrTensor> out{out_size, out_size};
  cuda_check(cudaMemsetAsync(out.first.gpu_data, 0, sizeof(Fr_t) * out_size, cur_stream()));
  cuda_check(cudaMemsetAsync(out.second.gpu_data, 0, sizeof(Fr_t) * out_size, cur_stream()));
  sync();
  // even windows -> first, odd windows -> second: two strided 2-D copies for the whole windows (source pitch 2 w, destination
  // pitch w), then at most two ragged windows each by plain copies (the reference uses one kernel, fr-tensor.cu:376-397)
  const size_t w = window_size, full_dst = out_size / w;
  size_t full0 = size >= w ? (size - w) / (2 * w) + 1 : 0, full1 = size / (2 * w);
  if (full0 > full_dst) full0 = full_dst;
  if (full1 > full_dst) full1 = full_dst;
  if (full0) cuda_check(cudaMemcpy2DAsync(out.first.gpu_data, sizeof(Fr_t) * w, gpu_data, sizeof(Fr_t) * 2 * w, sizeof(Fr_t) * w, full0, cudaMemcpyDeviceToDevice, cur_stream()));
  if (full1) cuda_check(cudaMemcpy2DAsync(out.second.gpu_data, sizeof(Fr_t) * w, gpu_data + w, sizeof(Fr_t) * 2 * w, sizeof(Fr_t) * w, full1, cudaMemcpyDeviceToDevice, cur_stream()));
  for (size_t wid = full1; wid * w < out_size; ++wid) {
    size_t dst = wid * w, n0 = std::min<size_t>(w, out_size - dst);
    size_t s0 = 2 * wid * w, s1 = s0 + w;
    if (wid >= full0 && s0 < size) cuda_check(cudaMemcpyAsync(out.first.gpu_data + dst, gpu_data + s0, sizeof(Fr_t) * std::min<size_t>(n0, size - s0), cudaMemcpyDeviceToDevice, cur_stream()));
    if (s1 < size) cuda_check(cudaMemcpyAsync(out.second.gpu_data + dst, gpu_data + s1, sizeof(Fr_t) * std::min<size_t>(n0, size - s1), cudaMemcpyDeviceToDevice, cur_stream()));
  }
  sync();
  return out;
}
FrTensor FrTensor::partial_me(vector<Fr_t> u, uint window_size) const {
  FrTensor out((uint)zkdl_partial_me_size(size, u.size(), window_size));
  check(zkdl_fr_partial_me(gpu_data, size, u.data(), u.size(), window_size, out.gpu_data, st())); sync();
  return out;
}
FrTensor Fr_partial_me(const FrTensor& t, vector<Fr_t>::const_iterator begin, vector<Fr_t>::const_iterator end, uint window_size) {
  if (begin >= end) return t;
  return t.partial_me(vector<Fr_t>(begin, end), window_size);
}
// The reference's generator (curand XORWOW, curand_init(seed, index, 0)) through the C ABI; the 64-bit seed comes from
// std::random_device + mt19937_64 as in the reference (fr-tensor.cu:319-329), or from the injected seed stream.
static unsigned long random_seed64() {
  std::mt19937_64 rng(next_challenge_seed());
  std::uniform_int_distribution<unsigned long> distribution(0, ULONG_MAX);
  return distribution(rng);
}
FrTensor FrTensor::random_int(uint size, uint num_bits) {                          // fr-tensor.cu:302-335
  FrTensor out(size);
  check(zkdl_fr_random_int(out.gpu_data, num_bits, size, random_seed64(), st())); sync();
  return out;
}
FrTensor FrTensor::random(uint size) {                                              // fr-tensor.cu:337-368
  FrTensor out(size);
  check(zkdl_fr_random(out.gpu_data, size, random_seed64(), st())); sync();
  return out;
}

// ------------------------------------------------------------------------------------------------ G1TensorAffine
G1TensorAffine::G1TensorAffine(const G1TensorAffine& t) : G1Tensor(t.size), gpu_data(dev_alloc<G1Affine_t>(t.size)) {
  copy(gpu_data, t.gpu_data, sizeof(G1Affine_t) * size, cudaMemcpyDeviceToDevice);
}
G1TensorAffine::G1TensorAffine(uint size) : G1Tensor(size), gpu_data(dev_alloc<G1Affine_t>(size)) {}
G1TensorAffine::G1TensorAffine(uint size, const G1Affine_t& g) : G1Tensor(size), gpu_data(dev_alloc<G1Affine_t>(size)) {
  vector<G1Affine_t> h(size, g);
  copy(gpu_data, h.data(), sizeof(G1Affine_t) * size, cudaMemcpyHostToDevice);
}
G1TensorAffine::G1TensorAffine(uint size, const G1Affine_t* cpu_data) : G1Tensor(size), gpu_data(dev_alloc<G1Affine_t>(size)) {
  copy(gpu_data, cpu_data, sizeof(G1Affine_t) * size, cudaMemcpyHostToDevice);
}
G1TensorAffine::~G1TensorAffine() { dev_free(gpu_data); gpu_data = nullptr; }
G1Affine_t G1TensorAffine::operator()(uint idx) const {
  G1Affine_t out; copy(&out, gpu_data + idx, sizeof(G1Affine_t), cudaMemcpyDeviceToHost); return out;
}
G1TensorAffine G1TensorAffine::operator-() const {                                  // y -> p - y via the Jacobian kernel
  G1TensorJacobian j(*this);
  G1TensorJacobian nj = -j;
  vector<G1Jacobian_t> h(size); vector<G1Affine_t> a(size);
  copy(h.data(), nj.gpu_data, sizeof(G1Jacobian_t) * size, cudaMemcpyDeviceToHost);
  for (uint i = 0; i < size; ++i) { a[i].x = h[i].x; a[i].y = h[i].y; }
  return G1TensorAffine(size, a.data());
}

// ------------------------------------------------------------------------------------------------ G1TensorJacobian
G1TensorJacobian::G1TensorJacobian(const G1TensorJacobian& t) : G1Tensor(t.size), gpu_data(dev_alloc<G1Jacobian_t>(t.size)), table_(t.table_) {
  copy(gpu_data, t.gpu_data, sizeof(G1Jacobian_t) * size, cudaMemcpyDeviceToDevice);
}
G1TensorJacobian::TableHolder::~TableHolder() { if (t) { cudaDeviceSynchronize(); zkdl_g1_table_destroy(t); } }
G1TensorJacobian::G1TensorJacobian(uint size) : G1Tensor(size), gpu_data(dev_alloc<G1Jacobian_t>(size)) {}
G1TensorJacobian::G1TensorJacobian(uint size, const G1Jacobian_t& g) : G1Tensor(size), gpu_data(dev_alloc<G1Jacobian_t>(size)) {
  vector<G1Jacobian_t> h(size, g);
  copy(gpu_data, h.data(), sizeof(G1Jacobian_t) * size, cudaMemcpyHostToDevice);
}
G1TensorJacobian::G1TensorJacobian(uint size, const G1Jacobian_t* cpu_data) : G1Tensor(size), gpu_data(dev_alloc<G1Jacobian_t>(size)) {
  copy(gpu_data, cpu_data, sizeof(G1Jacobian_t) * size, cudaMemcpyHostToDevice);
}
G1TensorJacobian::G1TensorJacobian(const G1TensorAffine& a) : G1Tensor(a.size), gpu_data(dev_alloc<G1Jacobian_t>(a.size)) {
  check(zkdl_g1_affine_to_jacobian(a.gpu_data, gpu_data, size, st())); sync();
}
G1TensorJacobian::~G1TensorJacobian() { invalidate_table(); dev_free(gpu_data); gpu_data = nullptr; }
void G1TensorJacobian::invalidate_table() const { table_.reset(); }
const zkdl_g1_table* G1TensorJacobian::table() const {
  if (!table_) {
    auto h = std::make_shared<TableHolder>();
    check(zkdl_g1_table_create(gpu_data, size, 0, 1, &h->t, st())); sync();
    table_ = h;
  }
  return table_->t;
}
G1Jacobian_t G1TensorJacobian::operator()(uint idx) const {
  G1Jacobian_t out; copy(&out, gpu_data + idx, sizeof(G1Jacobian_t), cudaMemcpyDeviceToHost); return out;
}
G1TensorJacobian G1TensorJacobian::binary(int op, const void* b, size_t nb) const {
  G1TensorJacobian out(size); check(zkdl_g1_elementwise(op, gpu_data, b, nb, out.gpu_data, size, st())); sync(); return out;
}
G1TensorJacobian& G1TensorJacobian::binary_inplace(int op, const void* b, size_t nb) {
  invalidate_table();
  check(zkdl_g1_elementwise(op, gpu_data, b, nb, gpu_data, size, st())); sync(); return *this;
}
template <class T> struct DevOne {     // a single point uploaded for the broadcast forms
  T* p; DevOne(const T& v) : p(dev_alloc<T>(1)) { copy(p, &v, sizeof(T), cudaMemcpyHostToDevice); } ~DevOne() { dev_free(p); }
};
static void same(uint a, uint b) { if (a != b) throw std::runtime_error("Incompatible dimensions"); }
G1TensorJacobian G1TensorJacobian::operator-() const { return binary(ZKDL_G1_NEG, nullptr, 0); }
G1TensorJacobian G1TensorJacobian::operator+(const G1TensorJacobian& t) const { same(size, t.size); return binary(ZKDL_G1_ADD, t.gpu_data, t.size); }
G1TensorJacobian G1TensorJacobian::operator+(const G1TensorAffine& t) const { same(size, t.size); return binary(ZKDL_G1_MADD, t.gpu_data, t.size); }
G1TensorJacobian G1TensorJacobian::operator+(const G1Jacobian_t& x) const { DevOne<G1Jacobian_t> d(x); return binary(ZKDL_G1_ADD, d.p, 1); }
G1TensorJacobian G1TensorJacobian::operator+(const G1Affine_t& x) const { DevOne<G1Affine_t> d(x); return binary(ZKDL_G1_MADD, d.p, 1); }
G1TensorJacobian& G1TensorJacobian::operator+=(const G1TensorJacobian& t) { same(size, t.size); return binary_inplace(ZKDL_G1_ADD, t.gpu_data, t.size); }
G1TensorJacobian& G1TensorJacobian::operator+=(const G1TensorAffine& t) { same(size, t.size); return binary_inplace(ZKDL_G1_MADD, t.gpu_data, t.size); }
G1TensorJacobian& G1TensorJacobian::operator+=(const G1Jacobian_t& x) { DevOne<G1Jacobian_t> d(x); return binary_inplace(ZKDL_G1_ADD, d.p, 1); }
G1TensorJacobian& G1TensorJacobian::operator+=(const G1Affine_t& x) { DevOne<G1Affine_t> d(x); return binary_inplace(ZKDL_G1_MADD, d.p, 1); }
G1TensorJacobian G1TensorJacobian::operator-(const G1TensorJacobian& t) const { same(size, t.size); return binary(ZKDL_G1_SUB, t.gpu_data, t.size); }
G1TensorJacobian G1TensorJacobian::operator-(const G1TensorAffine& t) const { same(size, t.size); return binary(ZKDL_G1_MSUB, t.gpu_data, t.size); }
G1TensorJacobian G1TensorJacobian::operator-(const G1Jacobian_t& x) const { DevOne<G1Jacobian_t> d(x); return binary(ZKDL_G1_SUB, d.p, 1); }
G1TensorJacobian G1TensorJacobian::operator-(const G1Affine_t& x) const { DevOne<G1Affine_t> d(x); return binary(ZKDL_G1_MSUB, d.p, 1); }
G1TensorJacobian& G1TensorJacobian::operator-=(const G1TensorJacobian& t) { same(size, t.size); return binary_inplace(ZKDL_G1_SUB, t.gpu_data, t.size); }
G1TensorJacobian& G1TensorJacobian::operator-=(const G1TensorAffine& t) { same(size, t.size); return binary_inplace(ZKDL_G1_MSUB, t.gpu_data, t.size); }
G1TensorJacobian& G1TensorJacobian::operator-=(const G1Jacobian_t& x) { DevOne<G1Jacobian_t> d(x); return binary_inplace(ZKDL_G1_SUB, d.p, 1); }
G1TensorJacobian& G1TensorJacobian::operator-=(const G1Affine_t& x) { DevOne<G1Affine_t> d(x); return binary_inplace(ZKDL_G1_MSUB, d.p, 1); }
G1Jacobian_t G1TensorJacobian::sum() const {
  G1TensorJacobian out(1); check(zkdl_g1_sum(gpu_data, size, out.gpu_data, st())); sync(); return out(0);
}
G1TensorJacobian G1TensorJacobian::operator*(const FrTensor& s) const {             // g1-tensor.cu:447-454
  if (s.size % size != 0) throw std::runtime_error("Incompatible dimensions");
  G1TensorJacobian out(s.size); check(zkdl_g1_mul(gpu_data, size, s.gpu_data, s.size, out.gpu_data, st())); sync(); return out;
}
G1TensorJacobian& G1TensorJacobian::operator*=(const FrTensor& s) {                 // g1-tensor.cu:456-461
  if (size != s.size) throw std::runtime_error("Incompatible dimensions 01");
  invalidate_table();
  check(zkdl_g1_mul(gpu_data, size, s.gpu_data, s.size, gpu_data, st())); sync(); return *this;
}
G1Jacobian_t G1TensorJacobian::operator()(const vector<Fr_t>& u) const {
  G1TensorJacobian out(1); check(zkdl_g1_me(gpu_data, size, u.data(), u.size(), out.gpu_data, st())); sync(); return out(0);
}
G1Jacobian_t G1_me(const G1TensorJacobian& t, vector<Fr_t>::const_iterator begin, vector<Fr_t>::const_iterator end) {
  return t(vector<Fr_t>(begin, end));
}

// ------------------------------------------------------------------------------------------------ Commitment
G1TensorJacobian Commitment::commit(const FrTensor& t) const {                       // commitment.cu:29-41 (intended semantics)
  if (t.size % size != 0) throw std::runtime_error("Incompatible dimensions");
  G1TensorJacobian out(t.size / size);
  check(zkdl_commit(table(), t.gpu_data, t.size, out.gpu_data, st())); sync();
  return out;
}
Fr_t Commitment::me_open(const FrTensor& t, const Commitment& generators, vector<Fr_t>::const_iterator begin, vector<Fr_t>::const_iterator end,
                         vector<G1Jacobian_t>& proof) {                              // commitment.cu:62-81
  if (t.size != generators.size) throw std::runtime_error("Incompatible dimensions");
  size_t k = end - begin;
  G1TensorJacobian pr((uint)(3 * k + 1)); FrTensor ret(1);
  check(zkdl_me_open(generators.table(), t.gpu_data, t.size, k ? &*begin : nullptr, k, pr.gpu_data, ret.gpu_data, st())); sync();
  vector<G1Jacobian_t> h(3 * k + 1);
  copy(h.data(), pr.gpu_data, sizeof(G1Jacobian_t) * h.size(), cudaMemcpyDeviceToHost);
  proof.insert(proof.end(), h.begin(), h.end());
  return ret(0);
}
Fr_t Commitment::open_with_proof(const FrTensor& t, const G1TensorJacobian& c, const vector<Fr_t>& u, vector<G1Jacobian_t>& proof) const {
  uint khi = ceilLog2(c.size);
  if (u.size() < khi) throw std::runtime_error("Incompatible dimensions");
  size_t klo = u.size() - khi;
  G1TensorJacobian pr((uint)(3 * klo + 2)); FrTensor ret(1);
  check(zkdl_open(table(), c.table(), t.gpu_data, t.size, u.data(), u.size(), pr.gpu_data, pr.gpu_data + 1, ret.gpu_data, st())); sync();
  vector<G1Jacobian_t> h(3 * klo + 2);
  copy(h.data(), pr.gpu_data, sizeof(G1Jacobian_t) * h.size(), cudaMemcpyDeviceToHost);
  proof.insert(proof.end(), h.begin(), h.end());
  return ret(0);
}
Fr_t Commitment::open(const FrTensor& t, const G1TensorJacobian& c, const vector<Fr_t>& u) const {   // commitment.cu:83-92
  vector<G1Jacobian_t> proof;
  return open_with_proof(t, c, u, proof);
}

// ------------------------------------------------------------------------------------------------ proof.cuh
vector<Fr_t> random_vec(uint len) {
  vector<Fr_t> out(len);
  zkdl_random_vec_host(next_challenge_seed(), len, out.data());
  return out;
}
uint ceilLog2(uint num) { return zkdl_ceil_log2(num); }

static vector<Fr_t> download(const FrTensor& t) {
  vector<Fr_t> h(t.size);
  copy(h.data(), t.gpu_data, sizeof(Fr_t) * t.size, cudaMemcpyDeviceToHost);
  return h;
}
vector<Fr_t> inner_product_sumcheck(const FrTensor& a, const FrTensor& b, vector<Fr_t> u) {
  if (a.size != b.size) throw std::runtime_error("Incompatible dimensions");
  FrTensor proof((uint)(3 * u.size() + 2));
  check(zkdl_ip_sumcheck(a.gpu_data, b.gpu_data, a.size, u.data(), u.size(), proof.gpu_data, st())); sync();
  return download(proof);
}
vector<Fr_t> hadamard_product_sumcheck(const FrTensor& a, const FrTensor& b, vector<Fr_t> u, vector<Fr_t> v) {
  if (u.size() != v.size()) throw std::runtime_error("Incompatible dimensions 1");
  if (a.size != b.size) throw std::runtime_error("Incompatible dimensions 2");
  FrTensor proof((uint)(3 * u.size() + 2));
  check(zkdl_hp_sumcheck(a.gpu_data, b.gpu_data, a.size, u.data(), v.data(), u.size(), proof.gpu_data, st())); sync();
  return download(proof);
}
vector<Fr_t> binary_sumcheck(const FrTensor& a, vector<Fr_t> u, vector<Fr_t> v) {
  if (u.size() != v.size()) throw std::runtime_error("Incompatible dimensions");
  FrTensor proof((uint)(3 * u.size() + 1));
  check(zkdl_bin_sumcheck(a.gpu_data, a.size, u.data(), v.data(), u.size(), proof.gpu_data, st())); sync();
  return download(proof);
}
// iterator forms (the reference's recursion entry points): same element order, no size guards beyond a.size == b.size
void Fr_ip_sc(const FrTensor& a, const FrTensor& b, vector<Fr_t>::const_iterator begin, vector<Fr_t>::const_iterator end, vector<Fr_t>& proof) {
  auto p = inner_product_sumcheck(a, b, vector<Fr_t>(begin, end)); proof.insert(proof.end(), p.begin(), p.end());
}
void Fr_hp_sc(const FrTensor& a, const FrTensor& b, vector<Fr_t>::const_iterator u_begin, vector<Fr_t>::const_iterator u_end,
              vector<Fr_t>::const_iterator v_begin, vector<Fr_t>::const_iterator v_end, vector<Fr_t>& proof) {
  auto p = hadamard_product_sumcheck(a, b, vector<Fr_t>(u_begin, u_end), vector<Fr_t>(v_begin, v_end)); proof.insert(proof.end(), p.begin(), p.end());
}
void Fr_bin_sc(const FrTensor& a, vector<Fr_t>::const_iterator u_begin, vector<Fr_t>::const_iterator u_end,
               vector<Fr_t>::const_iterator v_begin, vector<Fr_t>::const_iterator v_end, vector<Fr_t>& proof) {
  auto p = binary_sumcheck(a, vector<Fr_t>(u_begin, u_end), vector<Fr_t>(v_begin, v_end)); proof.insert(proof.end(), p.begin(), p.end());
}

// ------------------------------------------------------------------------------------------------ zkFC
zkFC::zkFC(uint input_size, uint output_size, const FrTensor& t, const Commitment& c)
    : weights(t), com(c.commit(t)), inputSize(input_size), outputSize(output_size) {        // zkfc.cu:102-104
  if (t.size != input_size * output_size) throw std::runtime_error("Incompatible dimensions");
  com.table();                       // fixed-base tables of the commitment vector: setup work, like commit() itself
}
zkFC zkFC::from_float_gpu_ptr(uint input_size, uint output_size, float* float_gpu_ptr, const Commitment& generators) {   // zkfc.cu:90-100
  uint I = 1u << ceilLog2(input_size), O = 1u << ceilLog2(output_size);
  FrTensor w(I * O);
  check(zkdl_float_to_fr(float_gpu_ptr, w.gpu_data, input_size, I, output_size, O, st())); sync();
  return zkFC(I, O, w.mont(), generators);
}
FrTensor zkFC::load_float_gpu_input(uint batch_size, uint input_dim, float* input_ptr) {                                // zkfc.cu:106-115
  uint B = 1u << ceilLog2(batch_size), I = 1u << ceilLog2(input_dim);
  FrTensor t(B * I);
  check(zkdl_float_to_fr(input_ptr, t.gpu_data, batch_size, B, input_dim, I, st())); sync();
  return t;
}
FrTensor zkFC::operator()(const FrTensor& X) const {                                                                    // zkfc.cu:117-126
  if (X.size % inputSize != 0) throw std::runtime_error("Incompatible dimensions");
  uint B = X.size / inputSize;
  FrTensor out(B * outputSize);
  if (!mm_) {                                     // weights never change after construction: derive the integer copy once
    mm_ = std::make_shared<MMHolder>();
    check(zkdl_mm_weights_create(weights.gpu_data, inputSize, outputSize, &mm_->w, st()));
  }
  check(zkdl_fr_matmul_prepared(X.gpu_data, weights.gpu_data, mm_->w, out.gpu_data, B, st())); sync();
  return out;
}
zkFC::MMHolder::~MMHolder() { if (w) zkdl_mm_weights_destroy(w); }
void zkFC::prove(const FrTensor& X, const FrTensor& Z, Commitment& generators) const {                                  // zkfc.cu:128-145
  if (X.size % inputSize != 0) throw std::runtime_error("Incompatible dimensions 1");
  uint B = X.size / inputSize;
  auto u_bs = random_vec(ceilLog2(B));
  auto u_in = random_vec(ceilLog2(inputSize));
  auto u_out = random_vec(ceilLog2(outputSize));
  size_t nfr, ng1;
  zkdl_zkfc_proof_sizes(B, inputSize, outputSize, generators.size, &nfr, &ng1);
  if (!pbuf_ || pbuf_->nfr != nfr || pbuf_->ng1 != ng1) {               // first proof of this object (or a new shape): plain cudaMalloc, kept
    pbuf_ = std::make_shared<ProofBuf>();
    cuda_check(cudaMalloc(&pbuf_->fr, sizeof(Fr_t) * nfr)); cuda_check(cudaMalloc(&pbuf_->g1, sizeof(G1Jacobian_t) * ng1));
    pbuf_->nfr = nfr; pbuf_->ng1 = ng1;
  }
  // the integer copy of the weights exists once operator() has run (the forward pass precedes the proof, demo.cu:116-138)
  check(zkdl_zkfc_prove_parts(X.gpu_data, weights.gpu_data, mm_ ? mm_->w : nullptr, Z.gpu_data, B, inputSize, outputSize,
                              generators.table(), com.table(), u_bs.data(), u_in.data(), u_out.data(), pbuf_->fr, pbuf_->g1,
                              ZKDL_FC_SUMCHECK | ZKDL_FC_OPENING, st()));
  pbuf_->stream = cur_stream(); pbuf_->pending = true;
  if (!async_prove()) fetch_proof();
}
zkFC::ProofBuf::~ProofBuf() { cudaFree(fr); cudaFree(g1); }
void zkFC::fetch_proof() const {
  if (!pbuf_ || !pbuf_->pending) return;
  cuda_check(cudaStreamSynchronize(pbuf_->stream));
  proof_fr_.resize(pbuf_->nfr); proof_g1_.resize(pbuf_->ng1);
  cuda_check(cudaMemcpy(proof_fr_.data(), pbuf_->fr, sizeof(Fr_t) * pbuf_->nfr, cudaMemcpyDeviceToHost));
  cuda_check(cudaMemcpy(proof_g1_.data(), pbuf_->g1, sizeof(G1Jacobian_t) * pbuf_->ng1, cudaMemcpyDeviceToHost));
  pbuf_->pending = false;
}

// ------------------------------------------------------------------------------------------------ zkReLU
bool zkReLU::materialize_tables = true;
void zkReLU::reset_ptrs(uint size) {                                                                                    // zkrelu.cu:53-62
  delete sign_ptr; delete mag_bin_ptr; delete rem_bin_ptr; mag_bin_ptr = rem_bin_ptr = nullptr;
  dev_free(mag_packed_); dev_free(rem_packed_);
  sign_ptr = new FrTensor(size);
  mag_packed_ = dev_alloc<uint32_t>(size); rem_packed_ = dev_alloc<uint16_t>(size); n_ = size;
  if (materialize_tables) { mag_bin_ptr = new FrTensor(size * 32); rem_bin_ptr = new FrTensor(size * 16); }
}
zkReLU::~zkReLU() {
  delete sign_ptr; delete mag_bin_ptr; delete rem_bin_ptr; sign_ptr = mag_bin_ptr = rem_bin_ptr = nullptr;
  dev_free(mag_packed_); dev_free(rem_packed_); mag_packed_ = nullptr; rem_packed_ = nullptr;
  cudaFree(dev_proof_); dev_proof_ = nullptr;
}
FrTensor zkReLU::operator()(const FrTensor& X) {                                                                        // zkrelu.cu:44-51
  reset_ptrs(X.size);
  FrTensor out(X.size);
  check(zkdl_relu_packed(X.gpu_data, out.gpu_data, sign_ptr->gpu_data, mag_packed_, rem_packed_, X.size, nullptr, st()));
  if (materialize_tables) check(zkdl_relu_expand(mag_packed_, rem_packed_, mag_bin_ptr->gpu_data, rem_bin_ptr->gpu_data, X.size, st()));
  sync();
  return out;
}
void zkReLU::prove(const FrTensor& X, const FrTensor& Z) {                                                              // zkrelu.cu:79-100
  if (X.size != Z.size) throw std::runtime_error("Incompatible dimensions");
  if (!sign_ptr || sign_ptr->size != X.size) throw std::runtime_error("Incompatible dimensions");
  uint L = ceilLog2(X.size);
  auto u_z = random_vec(L + 5), v_z = random_vec(L + 5), u_r = random_vec(L + 4), v_r = random_vec(L + 4), u_rec = random_vec(L);
  auto u_hp = random_vec(L), v_hp = random_vec(L);
  const size_t np = zkdl_zkrelu_proof_size(X.size);
  if (!dev_proof_ || dev_proof_n_ != np) { cudaFree(dev_proof_); cuda_check(cudaMalloc(&dev_proof_, sizeof(Fr_t) * np)); dev_proof_n_ = np; }
  check(zkdl_zkrelu_prove_packed(X.gpu_data, sign_ptr->gpu_data, mag_packed_, rem_packed_, X.size, u_z.data(), v_z.data(), u_r.data(),
                                 v_r.data(), u_rec.data(), u_hp.data(), v_hp.data(), dev_proof_, st()));
  proof_stream_ = cur_stream(); proof_pending_ = true;
  if (!async_prove()) fetch_proof();
}
void zkReLU::fetch_proof() {
  if (!proof_pending_) return;
  cuda_check(cudaStreamSynchronize(proof_stream_));
  proof_.resize(dev_proof_n_);
  cuda_check(cudaMemcpy(proof_.data(), dev_proof_, sizeof(Fr_t) * dev_proof_n_, cudaMemcpyDeviceToHost));
  proof_pending_ = false;
}
