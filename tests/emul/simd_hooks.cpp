// TEST INFRASTRUCTURE: C entry points over the host path's streaming kernels (deodr_b200/csrc/host_simd.cpp) so that
// tests/test_host_simd.py can drive them on the CPU (they are host-only code of the product; no GPU needed).
#include "../../deodr_b200/csrc/host_simd.h"

extern "C" {
void hook_f64_to_f32(float *dst, const double *src, long n) { deodr_simd_f64_to_f32(dst, src, (size_t)n); }
void hook_f32_to_f64(double *dst, const float *src, long n) { deodr_simd_f32_to_f64(dst, src, (size_t)n); }
void hook_f32_add_f64(double *dst, const float *src, long n) { deodr_simd_f32_add_f64(dst, src, (size_t)n); }
void hook_copy(void *dst, const void *src, long bytes) { deodr_simd_copy(dst, src, (size_t)bytes); }
void hook_zero(void *dst, long bytes) { deodr_simd_zero(dst, (size_t)bytes); }
int hook_equal_f32(const double *user, const float *mirror, long n) { return deodr_simd_equal_f32(user, mirror, (size_t)n); }
}
