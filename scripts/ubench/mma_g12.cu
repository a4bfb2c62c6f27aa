// mma_g12.cu -- the two MMA sequences of k_accum_ws in isolation (same shapes, descriptors and TMEM
// columns), optionally with other warps storing to shared memory / loading and storing tensor memory at
// the same time.  Prints cycles per MMA.
#include <cstdio>
#include <vector>
#include "ws_kernels.cuh"
using namespace hmmk;

// bg bit0: 4 warps st.shared.v4 in a loop; bit1: 4 warps tcgen05.ld/st in a loop; bit2: 4 warps FMA spin
__global__ void __launch_bounds__(288, 1) k(int which, int rep, int bg, long long *out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 160 * 1024 / 4; i += 288) reinterpret_cast<uint32_t *>(sm)[i] = 0x3f800000u;
  if (warp == 8) tmem_alloc(&tslot, 512);
  if (tid == 0) { mbar_init(&mbar, 1); stop = 0; }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  const int KP = 80, NSLAB = 10;
  const uint32_t PX = (KP / 4) * 128;
  if (warp == 8) {
    const uint32_t tb = __shfl_sync(0xffffffffu, tm, 0);
    const long long t0 = clock64();
    if (elect_one_sync()) {
      if (which == 1) {
        const uint32_t idesc1 = make_idesc_tf32(128, 64);
        const uint64_t bh = make_smem_desc2(smem_u32(sm), 128, PX), bl = make_smem_desc2(smem_u32(sm) + 8 * PX, 128, PX);
        for (int r = 0; r < rep; r++) {
          uint32_t accf = 0;
          for (int p = 0; p < 3; p++) {
            const uint32_t a0 = tb + 352 + ((p == 1) ? KP : 0);
            const uint64_t b0 = (p == 2) ? bl : bh;
#pragma unroll
            for (int j = 0; j < NSLAB; j++) { tc_mma_tf32_ts(tb + (r & 1) * 64, a0 + j * 8, b0 + (uint64_t)(j * 16), idesc1, accf); accf = 1; }
          }
        }
      } else {
        const uint32_t idesc2 = make_idesc_tf32(128, 80);
        const uint64_t bh = make_smem_desc2(smem_u32(sm) + 16 * PX, 128, 2048), bl = make_smem_desc2(smem_u32(sm) + 16 * PX + 10 * 2048, 128, 2048);
        for (int r = 0; r < rep; r++) {
          uint32_t accf = 0;
          for (int p = 0; p < 3; p++) {
            const uint32_t a0 = tb + (r & 1) * 64 + ((p == 1) ? 128 : 0);
            const uint64_t b0 = (p == 2) ? bl : bh;
#pragma unroll
            for (int kk = 0; kk < 8; kk++) { tc_mma_tf32_ts(tb + 256, a0 + kk * 8, b0 + (uint64_t)(kk * 16), idesc2, accf); accf = 1; }
          }
        }
      }
      tc_commit(&mbar);
    }
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(&mbar, 0);
    const long long t2 = clock64();
    stop = 1;
    if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  } else if (warp < 4 && (bg & 1)) {
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    uint32_t base = smem_u32(sm) + 100 * 1024 + tid * 16;
    while (!stop) {
#pragma unroll
      for (int q = 0; q < 16; q++) st_shared_v4(base + q * 2048, v);
    }
  } else if (warp >= 4 && warp < 8 && (bg & 2)) {
    uint32_t r[16];
    const uint32_t ta = tm + ((uint32_t)(32 * (warp & 3)) << 16) + 448;
    for (int q = 0; q < 16; q++) r[q] = q;
    while (!stop) {
      tmem_st16(ta, r);
      tmem_wait_st();
      tmem_ld16(ta + 16, r);
    }
    if (r[0] == 12345) out[0] = 0;
  } else if (warp < 4 && (bg & 4)) {
    float a = tid, b = 1.0001f;
    while (!stop) {
#pragma unroll
      for (int q = 0; q < 64; q++) a = fmaf(a, b, 0.5f);
    }
    if (a == 12345.f) out[0] = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc(tm, 512);
}

__device__ __forceinline__ void dummy() {}

int main() {
  long long *d; cudaMalloc(&d, 148 * 2 * 8);
  const size_t smem = 161 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int rep = 100, grid = 148;
  for (int which = 1; which <= 2; which++)
    for (int bg : {0, 1, 2, 3, 4, 7}) {
      k<<<grid, 288, smem>>>(which, rep, bg, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid * 2);
      cudaMemcpy(h.data(), d, grid * 16, cudaMemcpyDeviceToHost);
      double issue = 0, total = 0;
      const int per = which == 1 ? 30 : 24;
      for (int b = 0; b < grid; b++) { issue += h[2 * b]; total += h[2 * b + 1]; }
      printf("GEMM%d (N=%d) bg=%d: issue %.1f cyc/MMA, retire %.1f cyc/MMA\n", which, which == 1 ? 64 : 80, bg, issue / grid / rep / per, total / grid / rep / per);
    }
  return 0;
}
