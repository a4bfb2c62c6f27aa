// tc_probe_h.cu -- can ONE half-precision frame tile in shared memory feed both GEMMs of the accumulate pass?
//   tile X[32 frames][80 columns], byte(f, k) = (f/8)*1280 + (k/8)*128 + (f%8)*16 + (k%8)*2
//   GEMM1  L[128][32] = W[128][80] X^T     B = X as a K-major operand   (N = frames,  K = columns)
//   GEMM2  S[128][80] = w[128][32] X       B = X as an MN-major operand (N = columns, K = frames)
// A operands from tensor memory (two halves per 32-bit column).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I speech_recognition_hmm_continuous_b200/csrc -I include scripts/tc_probe_h.cu -o scripts/build/tc_probe_h
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "tc_kernels.cuh"
#include "ws_kernels.cuh"
using namespace hmmk;

// variant: 0 = GEMM1; 1 = GEMM2 with (LBO, SBO) = (1280, 128); 2 = GEMM2 with (128, 1280)
__global__ void probe(const float *W, const float *X, const float *Wt, float *Dout, int variant) {
  __shared__ __align__(1024) uint8_t sX[32 * 80 * 2];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) mbar_init(&mbar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s, trow = (uint32_t)(32 * warp) << 16;
  for (int idx = tid; idx < 32 * 80; idx += blockDim.x) {
    const int f = idx / 80, k = idx % 80;
    *reinterpret_cast<__half *>(sX + (f / 8) * 1280 + (k / 8) * 128 + (f % 8) * 16 + (k % 8) * 2) = __float2half_rn(X[idx]);
  }
  {  // A operands: W (80 halves = 40 columns) at column 256, w (32 halves = 16 columns) at column 320
    uint32_t r[16];
    for (int c0 = 0; c0 < 48; c0 += 16) {
      for (int c = 0; c < 16; c++) {
        const int k = 2 * (c0 + c);
        const __half2 h = __floats2half2_rn(k < 80 ? W[tid * 80 + k] : 0.f, k + 1 < 80 ? W[tid * 80 + k + 1] : 0.f);
        r[c] = *reinterpret_cast<const uint32_t *>(&h);
      }
      tmem_st16(tm + trow + 256 + c0, r);
    }
    for (int c = 0; c < 16; c++) {
      const __half2 h = __floats2half2_rn(Wt[tid * 32 + 2 * c], Wt[tid * 32 + 2 * c + 1]);
      r[c] = *reinterpret_cast<const uint32_t *>(&h);
    }
    tmem_st16(tm + trow + 320, r);
    tmem_wait_st();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  const int ND = variant == 0 ? 32 : 80;
  if (tid == 0) {
    tc_fence_after();
    if (variant == 0) {
      const uint32_t idesc = make_idesc_f16(128, 32);
      const uint64_t b = make_smem_desc2(smem_u32(sX), 128, 1280);
      for (int j = 0; j < 5; j++) tc_mma_f16_ts(tm, tm + 256 + j * 8, b + (uint64_t)(j * 16), idesc, j > 0);
    } else {
      const uint32_t idesc = make_idesc_f16(128, 80) | (1u << 16);
      const uint64_t b = variant == 1 ? make_smem_desc2(smem_u32(sX), 1280, 128) : make_smem_desc2(smem_u32(sX), 128, 1280);
      for (int j = 0; j < 2; j++) tc_mma_f16_ts(tm, tm + 320 + j * 8, b + (uint64_t)(j * 160), idesc, j > 0);
    }
    tc_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < ND; c0 += 8) {
    float v[8];
    tmem_ld8(tm + trow + c0, v);
    for (int j = 0; j < 8; j++) Dout[tid * ND + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  std::vector<float> W(128 * 80), X(32 * 80), Wt(128 * 32), D(128 * 80);
  for (size_t i = 0; i < W.size(); i++) W[i] = (float)((i * 7 + 3) % 11) - 5.f;
  for (size_t i = 0; i < X.size(); i++) X[i] = (float)((i * 5 + 1) % 13) - 6.f;
  for (size_t i = 0; i < Wt.size(); i++) Wt[i] = (float)((i * 3 + 2) % 7) - 3.f;
  float *dW, *dX, *dWt, *dD;
  cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dWt, Wt.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dWt, Wt.data(), Wt.size() * 4, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 3; variant++) {
    const int ND = variant == 0 ? 32 : 80;
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128>>>(dW, dX, dWt, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int nz = 0;
    for (int m = 0; m < 128; m++)
      for (int n = 0; n < ND; n++) {
        double want = 0;
        if (variant == 0) for (int k = 0; k < 80; k++) want += (double)W[m * 80 + k] * X[n * 80 + k];
        else for (int f = 0; f < 32; f++) want += (double)Wt[m * 32 + f] * X[f * 80 + n];
        maxerr = fmax(maxerr, fabs(want - D[m * ND + n]));
        nz += D[m * ND + n] != 0.f;
      }
    printf("variant %d: max err %g, nonzero %d / %d; D[0][0..3] = %g %g %g %g\n", variant, maxerr, nz, 128 * ND, D[0], D[1], D[2], D[3]);
  }
  return 0;
}
