// Cycles per step of the forward-backward chain of k_fb_res (one warp, `lanes` active lanes), in variants.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I speech_recognition_hmm_continuous_b200/csrc -I include \
//             -o scripts/ubench/chain_step scripts/ubench/chain_step.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fbres_kernels.cuh"
using namespace hmmk;
constexpr int NS = 5;
// variant 0: the kernel's step (exact power-of-two scaling every step, in place in shared memory)
// variant 1: no scaling at all (arithmetic floor)          variant 2: scaling factor taken from the previous step
// variant 3: variant 0 without the stores                   variant 4: float arithmetic, exact scaling
template <int VAR>
__global__ void k(long long *out, int T, int lanes, double a_self, double a_next) {
  extern __shared__ uint32_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t *p = sm + (warp * 32 + lane) * 0 + warp * (NS * (T + 4)) ;
  // every lane walks the same rows (bank-conflict free broadcast); enough for timing the dependent chain
  for (int i = threadIdx.x; i < NS * (T + 4); i += 32) p[i] = d32_pack(0.25 + 0.5 * ((i * 2654435761u >> 8) & 0xffff) / 65536.0);
  __syncwarp();
  double cs[NS], cn[NS], z[NS];
  for (int i = 0; i < NS; i++) { cs[i] = a_self; cn[i] = i ? a_next : 0.0; z[i] = i == 0 ? 1.0 : 0.0; }
  int esum = 0;
  long long t0 = clock64();
  if (lane < lanes) {
    if (VAR == 0 || VAR == 3) {
      double b0[NS], b1[NS];
      for (int i = 0; i < NS; i++) b0[i] = d32_unpack(p[NS + i]);
      for (int s = 1; s < T; s += 2) {
        for (int i = 0; i < NS; i++) b1[i] = d32_unpack(p[2 * NS + i]);
        if (VAR == 0) res_step<NS>(p + NS, b0, z, cs, cn, esum);
        else {
          double raw[NS];
          for (int i = 0; i < NS; i++) { double aux = z[i] * cs[i]; if (i > 0) aux = fma(z[i - 1], cn[i], aux); raw[i] = aux * b0[i]; }
          int e; const double r = pow2_scale_max<NS>(raw, e);
          for (int i = 0; i < NS; i++) z[i] = raw[i] * r;
          esum += e;
        }
        for (int i = 0; i < NS; i++) b0[i] = d32_unpack(p[3 * NS + i]);
        if (VAR == 0) res_step<NS>(p + 2 * NS, b1, z, cs, cn, esum);
        else {
          double raw[NS];
          for (int i = 0; i < NS; i++) { double aux = z[i] * cs[i]; if (i > 0) aux = fma(z[i - 1], cn[i], aux); raw[i] = aux * b1[i]; }
          int e; const double r = pow2_scale_max<NS>(raw, e);
          for (int i = 0; i < NS; i++) z[i] = raw[i] * r;
          esum += e;
        }
        p += 2 * NS;
      }
    } else if (VAR == 1) {
      for (int s = 1; s < T; s++) {
        double b[NS], raw[NS];
        for (int i = 0; i < NS; i++) b[i] = d32_unpack(p[NS + i]);
        for (int i = 0; i < NS; i++) { double aux = z[i] * cs[i]; if (i > 0) aux = fma(z[i - 1], cn[i], aux); raw[i] = aux * b[i]; }
        for (int i = 0; i < NS; i++) { z[i] = raw[i] * 1.7; p[NS + i] = d32_pack(z[i]); }
        p += NS;
      }
    } else if (VAR == 2) {
      double r = 1.0;
      for (int s = 1; s < T; s++) {
        double b[NS], raw[NS];
        for (int i = 0; i < NS; i++) b[i] = d32_unpack(p[NS + i]);
        for (int i = 0; i < NS; i++) { double aux = z[i] * cs[i]; if (i > 0) aux = fma(z[i - 1], cn[i], aux); raw[i] = aux * (b[i] * r); }
        int e; r = pow2_scale_max<NS>(raw, e);   // used by the NEXT step
        esum += e;
        for (int i = 0; i < NS; i++) { z[i] = raw[i]; p[NS + i] = d32_pack(z[i]); }
        p += NS;
      }
    } else if (VAR == 4) {
      float zf[NS], csf[NS], cnf[NS];
      for (int i = 0; i < NS; i++) { zf[i] = (float)z[i]; csf[i] = (float)cs[i]; cnf[i] = (float)cn[i]; }
      for (int s = 1; s < T; s++) {
        float b[NS], raw[NS];
        for (int i = 0; i < NS; i++) b[i] = __uint_as_float(p[NS + i] >> 3);
        for (int i = 0; i < NS; i++) { float aux = zf[i] * csf[i]; if (i > 0) aux = fmaf(zf[i - 1], cnf[i], aux); raw[i] = aux * b[i]; }
        int be = __float_as_int(raw[0]) >> 23;
        for (int i = 1; i < NS; i++) be = max(be, __float_as_int(raw[i]) >> 23);
        const bool ok = be > 0 && be < 255;
        const float r = __int_as_float(ok ? (254 - be) << 23 : 0x3f800000);
        esum += ok ? be - 127 : 0;
        for (int i = 0; i < NS; i++) { zf[i] = raw[i] * r; p[NS + i] = __float_as_uint(zf[i]); }
        p += NS;
      }
      for (int i = 0; i < NS; i++) z[i] = zf[i];
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NS; i++) s += z[i];
  if (lane == 0) { out[2 * (blockIdx.x * (blockDim.x >> 5) + warp)] = t1 - t0; out[2 * (blockIdx.x * (blockDim.x >> 5) + warp) + 1] = (long long)(s * 1e6) + esum; }
}
template <int VAR> void run(const char *name, int warps, int lanes) {
  const int T = 1000;
  long long *d, h[64];
  cudaMalloc(&d, sizeof(h));
  const size_t smem = sizeof(uint32_t) * warps * NS * (T + 4);
  cudaFuncSetAttribute(k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; rep++) k<VAR><<<1, 32 * warps, smem>>>(d, T, lanes, 0.98, 0.02);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-44s warps=%d lanes=%2d: %7.1f cycles/step (check %lld)  %s\n", name, warps, lanes, (double)h[0] / (T - 1), h[1], cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  run<0>("f64 exact scaling, in place (kernel)", 1, 8);
  run<0>("f64 exact scaling, in place (kernel)", 1, 32);
  run<0>("f64 exact scaling, 4 warps (one per SMSP?)", 4, 8);
  run<0>("f64 exact scaling, 8 warps", 8, 8);
  run<3>("f64 exact scaling, no stores", 1, 8);
  run<1>("f64 no scaling", 1, 8);
  run<2>("f64 scaling lagged one step", 1, 8);
  run<4>("f32 exact scaling", 1, 8);
  return 0;
}
