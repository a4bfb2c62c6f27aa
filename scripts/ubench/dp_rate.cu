// FP64 / FP32 FMA issue rate and dependent-chain latency on the device (one-off probe).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/dp_rate scripts/ubench/dp_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int CH>
__global__ void k(T *out, int iters, T a, T b) {
  T v[CH];
  for (int i = 0; i < CH; i++) v[i] = (T)threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = v[i] * a + b;
  }
  T s = 0;
  for (int i = 0; i < CH; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T, int CH> void run(const char *name, int blocks, int threads, int iters) {
  T *d; cudaMalloc(&d, sizeof(T) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<T, CH><<<blocks, threads>>>(d, iters, (T)1.0000001, (T)1e-9);
  cudaEventRecord(e0);
  k<T, CH><<<blocks, threads>>>(d, iters, (T)1.0000001, (T)1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)blocks * threads * iters * CH;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("%s CH=%d blocks=%d thr=%d: %.3f ms, %.2f TFMA/s, %.1f FMA/clk/SM (at %d MHz), cycles/iter/warp-chain=%.1f\n", name, CH, blocks, threads, ms,
         fma / ms / 1e9, fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000, ms * 1e-3 * khz * 1e3 / iters);
  cudaFree(d);
}
int main() {
  run<double, 8>("f64 throughput", 148 * 4, 512, 20000);
  run<double, 1>("f64 latency  ", 148, 32, 20000);
  run<double, 1>("f64 1chain 16w/SM", 148, 512, 20000);
  run<double, 2>("f64 2chain 16w/SM", 148, 512, 20000);
  run<float, 8>("f32 throughput", 148 * 4, 512, 20000);
  run<float, 1>("f32 latency  ", 148, 32, 20000);
  return 0;
}
