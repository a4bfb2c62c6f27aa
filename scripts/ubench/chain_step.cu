// Cycles per step of the forward-backward chains of k_fb_res (res_chain, fbres_kernels.cuh) under the kernel's
// conditions: one chain warp per team with two lanes per utterance, every lane on its own rows of shared memory,
// the team's other warps parked at the team barrier, 227 KB of shared memory per CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I speech_recognition_hmm_continuous_b200/csrc -I include \
//             -o scripts/ubench/chain_step scripts/ubench/chain_step.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fbres_kernels.cuh"
using namespace hmmk;
constexpr int NS = 5;


template <int VAR>
__global__ void __launch_bounds__(512, 1)
k(long long *out, const double *A, int count, int T, int nteams, int mapping) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int32_t sT[4][16], swoff[4][16], smodel[4][16];
  __shared__ double sphi[4][16], slp[4][16];
  __shared__ uint32_t sdummy[4][32 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = mapping ? warp >> 2 : warp & 3, role = mapping ? warp & 3 : warp >> 2;
  const bool chain_warp = mapping ? role == team : role == 0;
  constexpr int SW = res_slot_words<NS>();
  uint32_t *slot = reinterpret_cast<uint32_t *>(smem) + (size_t)team * SW;
  const int tt = role * 32 + lane;
  for (int i = tt; i < 256; i += 128) sdummy[team][i] = d32_pack(0.5);
  if (tt < 16) { sT[team][tt] = T - (tt & 3); swoff[team][tt] = tt * 2 * NS * T; smodel[team][tt] = 0; }
  for (int i = tt; i < 2 * NS * T * count; i += 128) slot[i] = d32_pack(0.25 + 0.5 * ((i * 2654435761u >> 8) & 0xffff) / 65536.0);
  __syncthreads();
  long long t0 = clock64();
  if (team < nteams && chain_warp) res_chain<NS>(lane, count, sT[team], swoff[team], smodel[team], A, slot, sdummy[team], sphi[team], slp[team]);
  long long t1 = clock64();
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(128) : "memory");
  if (chain_warp && lane == 0) { out[2 * team] = t1 - t0; out[2 * team + 1] = (long long)(slp[team][0] * 1e3); }
}
template <int VAR> void run(const char *name, long long *d, const double *A) {
  long long h[8];
  const int T = 300;
  cudaFuncSetAttribute(k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kResSmemBytes - 8192);
  for (int nteams = 1; nteams <= 4; nteams += 3)
    for (int count = 1; count <= 16; count = count == 1 ? 4 : count * 4) {
      if (count * 2 * NS * T > res_slot_words<NS>() - 2048) continue;
      for (int rep = 0; rep < 2; rep++) k<VAR><<<1, 512, kResSmemBytes - 8192>>>(d, A, count, T, nteams, 1);
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-52s %d team(s), %2d utt/team: %6.1f cycles/step (team 0), %6.1f (last)  check %lld  %s\n", name, nteams, count, (double)h[0] / T,
             (double)h[2 * (nteams - 1)] / T, h[1], cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
  long long *d;
  double *A, hA[25] = {0};
  for (int i = 0; i < 5; i++) { hA[i * 5 + i] = i < 4 ? 0.98 : 1.0; if (i < 4) hA[i * 5 + i + 1] = 0.02; }
  cudaMalloc(&d, 64); cudaMalloc(&A, sizeof(hA)); cudaMemcpy(A, hA, sizeof(hA), cudaMemcpyHostToDevice);
  // Variants measured while the loop was shaped (B200, 1965 MHz, cycles per step, one chain warp per sub-partition):
  //   per-lane bounds inside the loop (if (s < T) around every step)      244
  //   uniform loop over the common length, exact scaling before the next step 114
  //   ... stores issued one step late                                      108
  //   ... scaling folded into the next step's b~ (this kernel)             99
  run<0>("res_chain as in the kernel", d, A);
  return 0;
}
