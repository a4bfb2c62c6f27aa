// tc_probe.cu -- which tcgen05 operand forms work for kind::tf32 on this GPU?  One CTA, one MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I speech_recognition_hmm_continuous_b200/csrc -I include scripts/tc_probe.cu -o scripts/build/tc_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_kernels.cuh"
using namespace hmmk;

// variant bit0: A from TMEM; bit1: B MN-major
__global__ void probe(const float *A, const float *B, float *Dout, int N, int variant) {
  __shared__ __align__(1024) uint8_t sA[4096];
  __shared__ __align__(1024) uint8_t sB[8192];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) mbar_init(&mbar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const bool a_tmem = variant & 1, b_mn = variant & 2;
  // A: 128 x 8
  for (int k = 0; k < 8; k++) {
    const float v = A[tid * 8 + k];
    *reinterpret_cast<float *>(sA + (tid & 7) * 16 + (k & 3) * 4 + (k >> 2) * 128 + (tid >> 3) * 256) = v;
  }
  if (a_tmem) {
    uint32_t r[16];
    for (int k = 0; k < 16; k++) r[k] = __float_as_uint(k < 8 ? A[tid * 8 + k] : 0.f);
    tmem_st16(tm + ((uint32_t)(32 * warp) << 16) + 256, r);
    tmem_wait_st();
  }
  // B: N x 8
  for (int idx = tid; idx < N * 8; idx += blockDim.x) {
    const int n = idx / 8, k = idx % 8;
    uint32_t o = b_mn ? (uint32_t)((k & 7) * 16 + (n & 3) * 4 + (n >> 2) * 128)
                      : (uint32_t)((n & 7) * 16 + (k & 3) * 4 + (k >> 2) * 128 + (n >> 3) * 256);
    *reinterpret_cast<float *>(sB + o) = B[idx];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    uint32_t idesc = make_idesc_tf32(128, N) | (b_mn ? (1u << 16) : 0u);
    uint64_t bdesc = b_mn ? make_smem_desc2(smem_u32(sB), 2048, 128) : make_smem_desc2(smem_u32(sB), 128, 256);
    if (a_tmem) tc_mma_tf32_ts(tm, tm + 256, bdesc, idesc, 0);
    else tc_mma_tf32(tm, make_smem_desc2(smem_u32(sA), 128, 256), bdesc, idesc, 0);
    tc_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    float v[8];
    tmem_ld8(tm + ((uint32_t)(32 * warp) << 16) + c0, v);
    for (int j = 0; j < 8; j++) Dout[tid * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  const int N = 80;
  std::vector<float> A(128 * 8), B(N * 8), D(128 * N);
  for (int i = 0; i < 128 * 8; i++) A[i] = (float)((i * 7 + 3) % 11 - 5);
  for (int i = 0; i < N * 8; i++) B[i] = (float)((i * 5 + 1) % 13 - 6);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 4; variant++) {
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128>>>(dA, dB, dD, N, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int nz = 0;
    for (int m = 0; m < 128; m++)
      for (int n = 0; n < N; n++) {
        double want = 0;
        for (int k = 0; k < 8; k++) want += (double)A[m * 8 + k] * B[n * 8 + k];
        maxerr = fmax(maxerr, fabs(want - D[m * N + n]));
        nz += D[m * N + n] != 0.f;
      }
    printf("variant %d (A %s, B %s): max err %g, nonzero %d / %d; D[0][0..3] = %g %g %g %g\n", variant, (variant & 1) ? "tmem" : "smem",
           (variant & 2) ? "MN-major" : "K-major", maxerr, nz, 128 * N, D[0], D[1], D[2], D[3]);
  }
  return 0;
}
