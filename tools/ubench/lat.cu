// Dependent-chain latency probe for the instructions the serial loop kernels sit on (sm_100a).
// One warp, one CTA: cycles per dependent op = (clock64 delta) / chain length.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void lat(double* out, double a, double b, float fa, float fb, int nwarps_dummy) {
  __shared__ float sm[64];
  sm[threadIdx.x & 63] = fa;
  __syncthreads();
  double x = a + threadIdx.x;
  float f = fa + threadIdx.x;
  int idx = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, b, a);
    if (OP == 1) x = x * b;
    if (OP == 2) x = x + b;
    if (OP == 3) f = fmaf(f, fb, fa);
    if (OP == 4) f = __fadd_rn(f, fb);
    if (OP == 5) { x = (double)f; f = (float)(x) ; f = __fadd_rn(f, fb);}          // F2F both ways + FADD
    if (OP == 6) f = __shfl_sync(0xffffffffu, f, (idx + 1) & 31);
    if (OP == 7) { idx = (int)sm[idx] ; }                                            // LDS + F2I
    if (OP == 8) x = rint(x * b);
    if (OP == 9) { x = fma(x, b, a); f = fmaf(f, fb, fa); }                          // overlap check
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / N;
  out[1 + threadIdx.x] = x + f + idx;
}
// throughput: many warps, independent chains
template <int OP>
__global__ void thr(double* out, double a, double b) {
  double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < N; ++i) {
    x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
    x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
  double* d; cudaMalloc(&d, 1 << 24);
  const char* names[] = {"DFMA", "DMUL", "DADD", "FFMA", "FADD", "F2F.64<-32 + F2F.32<-64 + FADD", "SHFL", "LDS+F2I", "DMUL+rint", "DFMA||FFMA"};
  double h;
#define RUN(OP) lat<OP><<<1, 32>>>(d, 1.0000001, 0.9999999, 1.0001f, 0.9999f, 0); cudaDeviceSynchronize(); \
  lat<OP><<<1, 32>>>(d, 1.0000001, 0.9999999, 1.0001f, 0.9999f, 0); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("%-36s %.2f cyc/iter\n", names[OP], h);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 1; warps <= 32; warps *= 2) {
    thr<0><<<148, 32 * warps>>>(d, 1.0000001, 0.9999999); cudaDeviceSynchronize();
    cudaEventRecord(e0); thr<0><<<148, 32 * warps>>>(d, 1.0000001, 0.9999999); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dfma = 148.0 * 32 * warps * 8.0 * N;
    printf("DFMA throughput %2d warps/SM: %.1f DFMA/clk/SM (at 1.965 GHz)  %.2f TFLOP/s\n", warps, dfma / (ms * 1e-3) / 148 / 1.965e9, 2 * dfma / (ms * 1e-3) / 1e12);
  }
  return 0;
}
