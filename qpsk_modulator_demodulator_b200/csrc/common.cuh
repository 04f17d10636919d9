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
// Branch-free fp64 sincos for |x| < 1e5: 3-term Cody-Waite reduction to [-pi/4, pi/4] and Taylor
// polynomials (Estrin form) with DFMA.  Within 1 ulp of glibc's sin/cos (checked on 4e8 arguments on
// the host with fma(): identical after the cast to fp32), like CUDA's own sincos() — but straight-line
// code, so the scheduler can interleave it with the independent filter arithmetic of the loop kernels.
__device__ __forceinline__ void sincos_fast_f64(double x, double* sn, double* cs) {
  const double k = rint(x * 0.63661977236758134308);
  double r = fma(-k, 1.5707963267948966e+00, x);
  r = fma(-k, 6.123233995736766e-17, r);
  r = fma(-k, -1.4973849048591698e-33, r);
  const double z = r * r, z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
  // Estrin's scheme: dependent depth 4 instead of 10.  (Immediate constants on purpose: fetching them
  // from the constant bank measured slower in the loop kernels.)
  // sin(r) = r + r*z*S(z),  S = sum_{i<10} (-1)^(i+1) z^i / (2i+3)!
  const double a01 = fma(1.0 / 120.0, z, -1.0 / 6.0);
  const double a23 = fma(1.0 / 362880.0, z, -1.0 / 5040.0);
  const double a45 = fma(1.0 / 6227020800.0, z, -1.0 / 39916800.0);
  const double a67 = fma(1.0 / 355687428096000.0, z, -1.0 / 1307674368000.0);
  const double a89 = fma(1.0 / 51090942171709440000.0, z, -1.0 / 121645100408832000.0);
  const double S = fma(a89, z8, fma(fma(a67, z2, a45), z4, fma(a23, z2, a01)));
  const double sr = fma(r * z, S, r);
  // cos(r) = 1 - z/2 + z^2*C(z),  C = sum_{i<10} (-1)^i z^i / (2i+4)!
  const double d01 = fma(-1.0 / 720.0, z, 1.0 / 24.0);
  const double d23 = fma(-1.0 / 3628800.0, z, 1.0 / 40320.0);
  const double d45 = fma(-1.0 / 87178291200.0, z, 1.0 / 479001600.0);
  const double d67 = fma(-1.0 / 6402373705728000.0, z, 1.0 / 20922789888000.0);
  const double d89 = fma(-1.0 / 1124000727777607680000.0, z, 1.0 / 2432902008176640000.0);
  const double Cc = fma(d89, z8, fma(fma(d67, z2, d45), z4, fma(d23, z2, d01)));
  const double hz = 0.5 * z;
  const double w = 1.0 - hz;
  const double cr = w + (((1.0 - w) - hz) + z2 * Cc);     // compensated 1 - z/2 + z^2*C
  const int q = (int)k;
  const double s0 = (q & 1) ? cr : sr;
  const double c0 = (q & 1) ? sr : cr;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
}
// MathF.Sin/Cos model through the fast path (|x| small: loop phases are bounded)
__device__ __forceinline__ void sincos_f32_fast(float x, float* s, float* c) {
  double sd, cd;
  sincos_fast_f64((double)x, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
}
#endif

}  // namespace qpsk
