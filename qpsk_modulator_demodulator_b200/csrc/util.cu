// util.cu — counter-RNG fills and the FP32 FMA-pipe micro-benchmark used as a roofline denominator.
#include "common.cuh"

namespace qpsk {

__global__ void fill_uniform_kernel(uint64_t seed, uint64_t stream, long long first, long long n, float* out) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    out[k] = (float)(rng_double(seed, stream, (uint64_t)(first + k)) * 2.0 - 1.0);
}

// 8 independent FFMA2 chains per thread, register resident: measures the packed-FP32 FMA issue peak.
__global__ void __launch_bounds__(256) fma_peak_kernel(float2* out, int iters, float seedv) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seedv + i, seedv - i);
  const float2 m = make_float2(1.0000001f, 0.9999999f);
  const float2 c = make_float2(1e-7f, -1e-7f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ffma2(a[i], m, c);
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
  if (s.x == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace qpsk

using namespace qpsk;

extern "C" {

int qpsk_fill_uniform_dev(uint64_t seed, uint64_t stream_id, int64_t first, int64_t n, float* d_out, void* stream) {
  if (n < 0) return QPSK_ERR_RANGE;
  if (n == 0) return QPSK_OK;
  if (!d_out) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device());
  long long blocks = (n + 255) / 256;
  const long long cap = 16LL * device_sm_count();
  if (blocks > cap) blocks = cap;
  fill_uniform_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(seed, stream_id, first, n, d_out);
  QPSK_LAUNCH_CHECK();
  return QPSK_OK;
}

int qpsk_measure_fma_peak(double* tflops) {
  if (!tflops) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device());
  const int blocks = device_sm_count() * 8, threads = 256, iters = 4096;
  DevBuf<float2> sink;
  QPSK_TRY(sink.alloc((size_t)blocks * threads));
  cudaEvent_t e0, e1;
  QPSK_CUDA_TRY(cudaEventCreate(&e0));
  QPSK_CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    QPSK_CUDA_TRY(cudaEventRecord(e0, 0));
    fma_peak_kernel<<<blocks, threads>>>(sink.p, iters, 1.0f);
    QPSK_LAUNCH_CHECK();
    QPSK_CUDA_TRY(cudaEventRecord(e1, 0));
    QPSK_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    QPSK_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (double)blocks * threads * iters * 64.0 * 2.0 * 2.0;  // 64 FFMA2 x 2 lanes x 2 flop
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return QPSK_OK;
}

}  // extern "C"
