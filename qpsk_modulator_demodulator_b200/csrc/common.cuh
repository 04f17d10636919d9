// common.cuh — shared host/device helpers for libqpskcuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/qpskcuda.h"

namespace qpsk {

// ---- error plumbing ------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);

#define QPSK_CUDA_TRY(expr)                                          \
  do {                                                               \
    cudaError_t _e = (expr);                                         \
    if (_e != cudaSuccess) {                                         \
      ::qpsk::set_cuda_error(_e, #expr, __FILE__, __LINE__);         \
      return (_e == cudaErrorMemoryAllocation) ? QPSK_ERR_NOMEM : QPSK_ERR_CUDA; \
    }                                                                \
  } while (0)

#define QPSK_TRY(expr)                 \
  do {                                 \
    int _s = (expr);                   \
    if (_s != QPSK_OK) return _s;      \
  } while (0)

// after a kernel launch
#define QPSK_LAUNCH_CHECK()                      \
  do {                                           \
    ::qpsk::count_launch();                      \
    QPSK_CUDA_TRY(cudaGetLastError());           \
  } while (0)

int current_device();          // ordinal chosen by qpsk_set_device (default 0)
int ensure_device();           // cudaSetDevice(current) + arch check; status code
int device_sm_count();

// RAII device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  int alloc(size_t count) {
    release();
    if (count == 0) return QPSK_OK;
    QPSK_CUDA_TRY(cudaMalloc((void**)&p, count * sizeof(T)));
    n = count;
    return QPSK_OK;
  }
  int ensure(size_t count) { return (count <= n) ? QPSK_OK : alloc(count); }
  int zero(cudaStream_t s) {
    if (p) QPSK_CUDA_TRY(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    return QPSK_OK;
  }
};

// ---- counter RNG (spec: DESIGN.md "counter RNG"; the oracle restates the same function) -----
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t rng_u64(uint64_t seed, uint64_t stream, uint64_t counter) {
  uint64_t a = mix64(seed + 0x9E3779B97F4A7C15ULL * (stream + 1));
  return mix64(a + 0xD1B54A32D192ED03ULL * (counter + 1));
}
__host__ __device__ __forceinline__ double rng_double(uint64_t seed, uint64_t stream, uint64_t counter) {
  return (double)(rng_u64(seed, stream, counter) >> 11) * (1.0 / 9007199254740992.0);
}

// ---- device math helpers -------------------------------------------------------------------
#ifdef __CUDACC__
// packed fp32x2 FMA (Blackwell FFMA2): d = a*b + c on both halves, one issue slot.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
// MathF.Sin/Cos model: the correctly rounded fp32 value, i.e. fp64 evaluation rounded once.  The
// oracle uses the same definition (cr_sinf/cr_cosf), so results agree except where the fp64 value
// lies within an fp64 ulp of an fp32 rounding boundary (probability ~2^-28 per call).
__device__ __forceinline__ void sincos_f32_exact(float x, float* s, float* c) {
  double sd, cd;
  sincos((double)x, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
}
#endif

}  // namespace qpsk
