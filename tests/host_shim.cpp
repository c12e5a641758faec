// Host-compiled shim over zkdl_b200/csrc/*.cuh (device templates with the PTX carry chain emulated on the host).
// Lets the CPU test-suite check the exact template code the kernels instantiate against the oracle.
#include "../zkdl_b200/csrc/field.cuh"
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
}
