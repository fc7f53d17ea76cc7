// Streaming host kernels used by host_api.cu (see host_simd.cpp).
#pragma once
#include <cstddef>

void deodr_simd_f64_to_f32(float *dst, const double *src, size_t n);
void deodr_simd_f32_to_f64(double *dst, const float *src, size_t n);
void deodr_simd_f32_add_f64(double *dst, const float *src, size_t n);  // dst[i] += src[i]
void deodr_simd_copy(void *dst, const void *src, size_t bytes);
void deodr_simd_zero(void *dst, size_t bytes);
int deodr_simd_equal_f32(const double *user, const float *mirror, size_t n);  // (float)user[i] bit-equal mirror[i]
