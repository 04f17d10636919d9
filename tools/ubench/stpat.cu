// stpat.cu — write-bandwidth micro-benchmark for two store patterns of one warp instruction:
//   A: 32 lanes x 8 B contiguous (256 B)          B: 8 segments of 32 B, 256 B apart (mod_shape_kernel's pattern)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench/stpat.cu -o tools/ubench/stpat
#include <cstdio>
#include <cuda_runtime.h>
template <int PAT>
__global__ void k(float2* out, long long n_groups) {   // group = 32 samples (256 B); a warp handles 8 groups per pass
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long g0 = warp * 8; g0 + 8 <= n_groups; g0 += nwarps * 8) {
    float2* base = out + g0 * 32;
    if (PAT == 0) {
#pragma unroll
      for (int r = 0; r < 8; ++r) base[r * 32 + lane] = make_float2((float)r, (float)lane);
    } else {
      const int gl = lane >> 2, p = lane & 3;
#pragma unroll
      for (int r = 0; r < 8; ++r) base[gl * 32 + r * 4 + p] = make_float2((float)r, (float)lane);
    }
  }
}
int main() {
  const long long n = 1LL << 31;   // float2 -> 17.2 GB
  float2* d;
  cudaMalloc(&d, n * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int pat = 0; pat < 2; ++pat)
    for (int blocks : {148 * 4, 148 * 8, 148 * 32}) {
      for (int it = 0; it < 2; ++it) { if (pat == 0) k<0><<<blocks, 256>>>(d, n / 32); else k<1><<<blocks, 256>>>(d, n / 32); }
      cudaEventRecord(e0);
      for (int it = 0; it < 3; ++it) { if (pat == 0) k<0><<<blocks, 256>>>(d, n / 32); else k<1><<<blocks, 256>>>(d, n / 32); }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
      printf("pattern %c blocks %5d: %.3f ms  %.0f GB/s\n", pat ? 'B' : 'A', blocks, ms, n * 8 / ms / 1e6);
    }
  return 0;
}
