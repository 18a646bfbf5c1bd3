// FP64 add latency / throughput on one SM and on the whole GPU (B200): nvcc -arch=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double x) {
  double a = out[0];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a = __dadd_rn(a, x);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void thr(double* out, long long* cyc, double x) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = out[k];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = __dadd_rn(a[k], x);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 24); cudaMemset(out, 0, 1 << 24); cudaMalloc(&cyc, 64);
  long long h;
  for (int warps : {1, 4, 8, 16, 32}) {
    lat<<<1, 32 * warps>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    lat<<<1, 32 * warps>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DADD chain, %2d warps on one SM: %.1f cycles per DADD per warp\n", warps, h / 4096.0);
  }
  for (int warps : {1, 4, 8, 16, 32}) {
    thr<<<1, 32 * warps>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    thr<<<1, 32 * warps>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("8 independent DADD chains, %2d warps on one SM: %.2f cycles per warp-DADD -> %.1f DADD lanes / clk / SM\n", warps,
           h / 4096.0, 32.0 * warps * 4096.0 / h);
  }
  return 0;
}
