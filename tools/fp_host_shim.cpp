// Host build of ecfft_b200/csrc/fp.cuh for CPU unit tests (tests/test_fp_host.py).
// Test infrastructure only: lets the limb-level logic of the device arithmetic be
// checked against Python big integers without a GPU.
#include "../ecfft_b200/csrc/fp.cuh"
#include <string.h>
using namespace ecfft;
extern "C" {
void fph_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_mul(x, y); memcpy(r, &z, 32); }
void fph_mul_lazy(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_mul_lazy(x, y); memcpy(r, &z, 32); }
void fph_dot2(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* r) {
  Fp x, y, u, v; memcpy(&x, a, 32); memcpy(&y, b, 32); memcpy(&u, c, 32); memcpy(&v, d, 32);
  Fp z = fp_dot2_lazy(x, y, u, v); memcpy(r, &z, 32); }
void fph_muladd(const uint32_t* x0, const uint32_t* a, const uint32_t* b, uint32_t* r) {
  Fp x, y, u; memcpy(&x, x0, 32); memcpy(&y, a, 32); memcpy(&u, b, 32);
  Fp z = fp_muladd_lazy(x, y, u); memcpy(r, &z, 32); }
void fph_canon(const uint32_t* a, uint32_t* r) { Fp x; memcpy(&x, a, 32); Fp z = fp_canon(x); memcpy(r, &z, 32); }
void fph_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_add(x, y); memcpy(r, &z, 32); }
void fph_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_sub(x, y); memcpy(r, &z, 32); }
void fph_mont_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_mont_mul(x, y); memcpy(r, &z, 32); }
void fph_inv(const uint32_t* a, uint32_t* r) { Fp x; memcpy(&x, a, 32); Fp z = fp_inv(x); memcpy(r, &z, 32); }
void fph_pow(const uint32_t* a, uint64_t e, uint32_t* r) { Fp x; memcpy(&x, a, 32); Fp z = fp_pow_u64(x, e); memcpy(r, &z, 32); }
void fph_consts(uint32_t* R, uint32_t* RINV) { Fp a = fp_const_R(), b = fp_const_RINV(); memcpy(R, &a, 32); memcpy(RINV, &b, 32); }
}
extern "C" void fph_sqrt(const uint32_t* a, uint32_t* r) { Fp x; memcpy(&x, a, 32); Fp z = fp_sqrt_candidate(x); memcpy(r, &z, 32); }
extern "C" void fph_sub_lazy(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_sub_lazy(x, y); memcpy(r, &z, 32); }
extern "C" void fph_add_lazy(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_add_lazy(x, y); memcpy(r, &z, 32); }
extern "C" void fph_sub_lazy2(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_sub_lazy2(x, y); memcpy(r, &z, 32); }
extern "C" void fph_add_lazy_f(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_add_lazy_f(x, y); memcpy(r, &z, 32); }
extern "C" void fph_sub_lazy2_f(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); Fp z = fp_sub_lazy2_f(x, y); memcpy(r, &z, 32); }
