// mma_rate.cu -- cycles per tcgen05.mma (cta_group::1, M = 128) for kind::tf32 and kind::f16 (bf16),
// A from shared memory (SS) or tensor memory (TS), B K-major SWIZZLE_NONE in shared memory, as a
// function of N.  One CTA per SM; every CTA issues REP MMAs back to back into one accumulator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I speech_recognition_hmm_continuous_b200/csrc -I include scripts/ubench/mma_rate.cu -o scripts/ubench/mma_rate
#include <cstdio>
#include <vector>
#include "tc_kernels.cuh"
using namespace hmmk;

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 = tf32 SS, 1 = tf32 TS, 2 = bf16 SS, 3 = bf16 TS
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int rep, int nacc, long long *out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (4096 + 256 * 32 * 4) / 4; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = 0x3f800000u;
  if (warp == 0) tmem_alloc(&tslot, 512);
  if (tid == 0) mbar_init(&mbar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    const uint32_t tb = __shfl_sync(0xffffffffu, tm, 0);
    t0 = clock64();
    if (elect_one_sync()) {
      const uint64_t ad = make_smem_desc2(smem_u32(sm), 128, 256);
      const uint32_t idesc_tf = make_idesc_tf32(128, N);
      const uint32_t idesc_bf = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // descriptors precomputed; the loop body is 8 MMAs and nothing else
      uint64_t bd[4];
      for (int q = 0; q < 4; q++) bd[q] = make_smem_desc2(smem_u32(sm + 4096 + q * 256 * 32), 128, 256);
      const uint32_t d0 = tb, d1 = tb + (nacc > 1 ? 128 : 0);
      const uint32_t a0 = tb + 384, a1 = tb + 392;
      auto one = [&](uint32_t dd, uint32_t aa, uint64_t bb, uint32_t acc) {
        if (mode == 0) tc_mma_tf32(dd, ad, bb, idesc_tf, acc);
        else if (mode == 1) tc_mma_tf32_ts(dd, aa, bb, idesc_tf, acc);
        else if (mode == 2) mma_f16_ss(dd, ad, bb, idesc_bf, acc);
        else mma_f16_ts(dd, aa, bb, idesc_bf, acc);
      };
      one(d0, a0, bd[0], 0); one(d1, a1, bd[1], 0);
      for (int r = 0; r < rep; r += 8) {
        one(d0, a0, bd[0], 1); one(d1, a1, bd[1], 1); one(d0, a0, bd[2], 1); one(d1, a1, bd[3], 1);
        one(d0, a1, bd[0], 1); one(d1, a0, bd[1], 1); one(d0, a1, bd[2], 1); one(d1, a0, bd[3], 1);
      }
      tc_commit(&mbar);
    }
    __syncwarp();
    t1 = clock64();
  }
  mbar_wait(&mbar, 0);
  t2 = clock64();
  tc_fence_after();
  if (tid == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long *d; cudaMalloc(&d, 148 * 2 * 8);
  const size_t smem = 4096 + 4 * 256 * 32 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char *names[] = {"tf32 SS", "tf32 TS", "bf16 SS", "bf16 TS"};
  const int rep = 2000;
  for (int nacc : {1, 2}) {
    for (int mode = 0; mode < 4; mode++) {
      for (int N : {32, 64, 80, 96, 128, 192, 256}) {
        if (nacc > 1 && N > 128) continue;
        const int grid = 148;
        k<<<grid, 128, smem>>>(mode, N, rep, nacc, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[mode], N, cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(grid * 2);
        cudaMemcpy(h.data(), d, grid * 16, cudaMemcpyDeviceToHost);
        double issue = 0, total = 0;
        for (int b = 0; b < grid; b++) { issue += h[2 * b]; total += h[2 * b + 1]; }
        issue /= grid * (double)rep; total /= grid * (double)rep;
        printf("nacc %d %s M=128 N=%3d K=8/16: issue %.1f cyc/MMA, retire %.1f cyc/MMA -> %.0f flop/clk/SM\n", nacc, names[mode], N, issue, total,
               2.0 * 128 * N * (mode < 2 ? 8 : 16) / total);
      }
    }
  }
  return 0;
}
