// Wrapper translation unit: compiles the reference's own AVX distance kernels
// (pkg/distance/simd/cpp/avx.cpp — the C++ source the Go assembly in
// pkg/distance/simd/avx/AVX_amd64.s was generated from) UNMODIFIED, from where the file
// lies under $(REF).  No reference source is copied into this repository; the output goes
// to oracle/_ref/ (git-ignored).  TEST INFRASTRUCTURE ONLY (see coltt_oracle.cpp header).
//
// The #define dodges the clash between the reference's `inline float abs(float)`
// (avx.cpp:11-13) and libstdc++; <cstddef> supplies size_t.
#include <immintrin.h>
#include <cstddef>
#define abs coltt_ref_abs
#include REF_AVX_CPP
#undef abs

extern "C" {
__attribute__((visibility("default"))) void ref_cosine_similarity_dot_norm(size_t len, float* a, float* b, float* dot, float* n2) {
  cosine_similarity_dot_norm(len, a, b, dot, n2);
}
__attribute__((visibility("default"))) void ref_euclidean_distance_squared(size_t len, float* a, float* b, float* r) {
  euclidean_distance_squared(len, a, b, r);
}
__attribute__((visibility("default"))) void ref_manhattan_distance(size_t len, float* a, float* b, float* r) {
  manhattan_distance(len, a, b, r);
}
}
