// ws_kernels.cuh -- warp-specialised, mbarrier-pipelined tcgen05 kernels for sm_100a.
//
// k_emis_ws: Gaussian-mixture emission log-likelihoods (calc_symbol_probab + calc_gaus, T-FS:1749-1841,
// R-FS:860-947) as a dense contraction on the tensor pipe with the per-state log-sum-exp fused into
// the epilogue.  One persistent CTA per SM, nine warps in three roles that only meet at mbarriers:
//
//   warps 4-7  LOADERS   read 128 raw frames (fp32, centred; prefetched three levels ahead), form
//                        [x | x^2], split every value into TF32 hi + lo and write the A operand with
//                        tcgen05.st into TENSOR MEMORY stage a (one frame per lane); (re)load the W
//                        image into shared memory when the unit's image changes
//   warp  8    MMA       one thread issues 3*KP/8 tcgen05.mma.kind::tf32 (hi*hi + lo*hi + hi*lo) per
//                        unit, A from TMEM, B = W from shared memory, into TMEM accumulator stage a,
//                        commits to free the operand stage and to publish the accumulator
//   warps 0-3  EPILOGUE  tcgen05.ld the accumulator (one frame per thread), log-sum-exp over the M
//                        mixtures of each state, store log b
//
// Why the frame operand lives in TMEM: with both operands in shared memory a K = 8 TF32 MMA of
// 128 x 80 reads (128 + 80) x 32 B = 6.6 KB for 40 cycles of math -- more than the 128 B/cycle the
// shared memory delivers; measured 120 cycles per MMA (scripts/debug_ws_timeline.py).  From TMEM the
// A operand costs no shared-memory bandwidth, the loaders need no generic->async proxy fence for it,
// and 160 KB of shared memory are freed.
//
// Two operand stages and two accumulator stages in TMEM keep the tensor pipe busy while the next
// frame tile is being expanded and the previous one reduced.  The additive constant kc[g] of every
// Gaussian is added in the epilogue in FP32 (round to nearest): folding it into the contraction was
// tried and rejected -- the tensor pipe's FP32 accumulation truncates, and a partial sum that starts
// at |kc| ~ 150 loses ~5e-5 of log-likelihood per frame, systematically (posteriors then sum to 1 - 5e-5).
// The epilogue forms log2(c_g N_g(x)) = fma(acc, log2 e, kc2[g]) with exactly the expression the
// accumulate kernel uses, so that the posteriors it recomputes are consistent with log b.
//
// Units are (W image, 128-frame tile) pairs ordered image-major, each CTA takes a contiguous range:
// training: explicit list (frames of one model gathered through frame_ids); decode: unit u = (u / ntiles,
// u % ntiles) over the contiguous frames of the current utterance batch.
//
// W image in shared memory (SWIZZLE_NONE, K-major, 16-byte chunks of 4 values), rows g, columns k:
//   byte(g, k) = (g%8)*16 + (k%4)*4 + (k/4)*128 + (g/8)*P,   P = (KP/4)*128;   LBO = 128, SBO = P,
//   K-step j starts at +256 j;  image = [hi: (TN/8) P][lo: (TN/8) P][kc2: TN floats].
// TMEM columns: operand stage a at 160 a: [x_hi (DP) | x2_hi (DP) | x_lo (DP) | x2_lo (DP)] (KP <= 80);
//               accumulator stage a at 320 + 96 a (TN <= 96).
#pragma once
#include "tc_kernels.cuh"

namespace hmmk {

constexpr int kWsThreads = 288;   // 9 warps
constexpr int kWsMaxTN = 96;      // Gaussians (columns) per W image

// W image in global and shared memory: [hi: (TN/8) P][lo: (TN/8) P][kc2: TN floats]
__host__ __device__ inline size_t ws_image_bytes(int TN, int KP) { return (size_t)2 * (TN / 8) * (KP / 4) * 128 + (size_t)TN * 4; }
__host__ __device__ inline size_t ws_emis_smem_bytes(int TN, int KP) { return ws_image_bytes(TN, KP) + 1024 + 272; }

// Mixtures per state as laid out in the W image: padded to a power of two (M <= 16) or to a multiple of
// 16, so that state boundaries fall on the 16-column chunks the epilogue reads; pad mixtures have W = 0
// and kc = -inf (density 0).
__host__ __device__ inline int ws_pad_m(int M) {
  if (M > 16) return (M + 15) / 16 * 16;
  int p = 1;
  while (p < M) p <<= 1;
  return p;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// TF32 hi / lo split: hi = v rounded to nearest (ties away) to 11 significant bits with two integer
// instructions, lo = v - hi exactly (FP32); the tensor core ignores the low 13 mantissa bits of lo, which
// costs at most 2^-23 |v|, unbiased because lo has either sign.  (cvt.rna.tf32.f32 compiles to a longer
// NaN-safe sequence on sm_100a, and a rounded lo doubles the load on the ALU pipe, which is what bounds
// the loader and epilogue warps.)  The values here are finite.
__device__ __forceinline__ void split_tf32_fast(float v, float &hi, float &lo) {
  hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
  lo = v - hi;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const float4 &v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// non-blocking poll (try_wait may suspend the thread for a long time when the phase is not complete)
__device__ __forceinline__ bool mbar_try(uint64_t *mbar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(mbar)), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

// W images for k_emis_ws: image i covers global states [img_state0[i], +img_nstates[i]), TN rows (whole
// states; rows beyond are zero), followed by kc2[TN] = log2(e) (ln c - 0.5 (D ln 2pi + ln|det|) - 0.5 sum mu^2 iv),
// -inf for a Gaussian with c == 0 or det == 0 (density 0 in the reference) and for the pad rows.
__global__ void k_pack_w_ws(const double *__restrict__ mu, const double *__restrict__ iv, const float *__restrict__ kc2all,
                            const double *__restrict__ ctr, int M, int MP, int D, int DP, int TN,
                            const int32_t *__restrict__ img_state0, const int32_t *__restrict__ img_nstates,
                            float *__restrict__ images) {
  const int img = blockIdx.y;
  const int KP = 2 * DP;
  const uint32_t P = (uint32_t)(KP / 4) * 128;
  const size_t img_floats = ws_image_bytes(TN, KP) / 4;
  float *hi = images + (size_t)img * img_floats;
  float *lo = hi + (size_t)(TN / 8) * (P / 4);
  float *kc2 = lo + (size_t)(TN / 8) * (P / 4);
  const int64_t s0g = img_state0[img];
  const int nst = img_nstates[img];
  // image row n = (state n / MP, mixture n % MP) -> Gaussian index, or -1 for a pad row
  auto gauss_of = [&](int n) -> int64_t {
    const int st = n / MP, m = n - st * MP;
    return (st < nst && m < M) ? (s0g + st) * M + m : -1;
  };
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < TN * KP; idx += gridDim.x * blockDim.x) {
    const int n = idx / KP, k = idx - n * KP;
    const int part = k / DP, d = k - part * DP;
    const int64_t g = gauss_of(n);
    float val = 0.f;
    if (g >= 0 && d < D) {
      const double m = mu[g * D + d] - ctr[d], w = iv[g * D + d];
      val = (float)(part == 0 ? m * w : -0.5 * w);
    }
    float h, l;
    split_tf32(val, h, l);
    const size_t o = ((size_t)(n & 7) * 16 + (k & 3) * 4 + (size_t)(k >> 2) * 128 + (size_t)(n >> 3) * P) / 4;
    hi[o] = h;
    lo[o] = l;
    if (k == 0) kc2[n] = (g >= 0) ? kc2all[g] : kNegInf;
  }
}

// units == nullptr: decode, unit u = (image u / ntiles_dec, frames [128 (u % ntiles_dec), ...) of the batch, nframes_dec in all)
// MP: padded mixtures per state when <= 16 (1, 2, 4, 8, 16); 0 = a multiple of 16 given at run time (M)
template <bool TRAIN, int MP>
__global__ void __launch_bounds__(kWsThreads, 1)
k_emis_ws(const TcTile *__restrict__ units, int nunits, int ntiles_dec, int nframes_dec, const int32_t *__restrict__ frame_ids,
          const float *__restrict__ x32, const float *__restrict__ images, int N, int M, int DP, int TN, float *__restrict__ logb,
          int64_t fbase, int64_t ldb, int S_total, int SCt, long long *__restrict__ tdbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int KP = 2 * DP, NSLAB = KP / 8;
  // optional timeline of CTA 0 (diagnostics): tdbg[unit][8] clock stamps
  auto stamp = [&](int i, int slot) {
    if (tdbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && i < 64) tdbg[i * 8 + slot] = clock64();
  };
  const uint32_t P = (uint32_t)(KP / 4) * 128;
  const uint32_t w_bytes = 2 * (uint32_t)(TN / 8) * P, img_bytes = w_bytes + (uint32_t)TN * 4;
  uint8_t *sm = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *Ws = sm;                        // [hi | lo]
  uint64_t *bars = reinterpret_cast<uint64_t *>(Ws + w_bytes);
  uint64_t *full = bars, *empty = bars + 2, *dfull = bars + 4, *dempty = bars + 6;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(&full[s], 128);    // every loader thread arrives
      mbar_init(&empty[s], 1);     // tcgen05.commit
      mbar_init(&dfull[s], 1);     // tcgen05.commit
      mbar_init(&dempty[s], 128);  // every epilogue thread arrives
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;

  const int per = (nunits + gridDim.x - 1) / gridDim.x;
  const int u_begin = blockIdx.x * per, u_end = min(nunits, u_begin + per);
  auto get_unit = [&](int ui) -> TcTile {
    if (TRAIN) return units[ui];
    const int img = ui / ntiles_dec, t = ui - img * ntiles_dec;
    return TcTile{t * kTcRows, min(kTcRows, nframes_dec - t * kTcRows), img, img * SCt, 0, 0};
  };

  if (warp >= 4 && warp < 8) {
    // =================================== LOADERS ===================================
    const int r = tid - 128;  // frame row of the tile
    const uint32_t rbase = (uint32_t)(r & 7) * 16 + (uint32_t)(r >> 3) * P;
    constexpr int kQ = 10;  // float4 per row held in registers (DP <= 40)
    const int nq = DP / 4;
    // Global loads run three levels ahead of the expansion so that no latency is exposed per unit:
    // unit descriptor (i+3) -> frame id (i+2) -> feature row (i+1), while unit i is split and stored.
    auto unit_at = [&](int ui) -> TcTile { return ui < u_end ? get_unit(ui) : TcTile{0, 0, -1, 0, 0, 0}; };
    auto frame_of = [&](const TcTile &u) -> int64_t {
      if (r >= u.nrows) return -1;
      return TRAIN ? (int64_t)__ldg(frame_ids + u.row0 + r) : fbase + u.row0 + r;
    };
    auto load_row = [&](int64_t f, float4 (&xv)[kQ]) {
      const float4 *src = reinterpret_cast<const float4 *>(x32 + (f < 0 ? 0 : f) * DP);
#pragma unroll
      for (int j = 0; j < kQ; j++) xv[j] = (f >= 0 && j < nq) ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    TcTile d0 = unit_at(u_begin), d1 = unit_at(u_begin + 1), d2 = unit_at(u_begin + 2);
    int64_t f1 = frame_of(d1);
    float4 xv[kQ];
    load_row(frame_of(d0), xv);
    int cur_img = -1;
    for (int ui = u_begin, i = 0; ui < u_end; ui++, i++) {
      const TcTile unit = d0;
      const int s = i & 1;
      float4 xn[kQ];
      load_row(f1, xn);                      // rows of unit i+1
      const int64_t f2 = frame_of(d2);       // frame id of unit i+2
      const TcTile d3 = unit_at(ui + 3);     // descriptor of unit i+3
      if (unit.img != cur_img) {
        // the tensor pipe may still be reading the old image: wait for the previous unit's MMAs
        if (i >= 1) mbar_wait(&empty[(i - 1) & 1], ((i - 1) >> 1) & 1);
        const float4 *wsrc = reinterpret_cast<const float4 *>(images + (size_t)unit.img * (img_bytes / 4));
        float4 *wdst = reinterpret_cast<float4 *>(Ws);
        for (int k = r; k < (int)(w_bytes / 16); k += 128) wdst[k] = __ldg(wsrc + k);
        cur_img = unit.img;
      }
      if (warp == 4) stamp(i, 0);
      mbar_wait(&empty[s], ((i >> 1) & 1) ^ 1);  // operand stage s is free (unit i-2 has been multiplied)
      if (warp == 4) stamp(i, 1);
      tc_fence_after();
      const uint32_t xa = tmem0 + (uint32_t)s * 160 + ((uint32_t)(32 * (warp & 3)) << 16);  // my lane, stage s
      // 16 columns at a time: chunk c of [x | x^2] covers float4 4c .. 4c+3 of the doubled row
#pragma unroll
      for (int c = 0; c < 2 * kQ / 4; c++) {
        if (c * 16 < KP) {
          uint32_t vh[16], vl[16];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int j = c * 4 + q;  // float4 index in [x | x^2]; static after unrolling
            float4 xx = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < nq) xx = xv[j < kQ ? j : 0];
            else if (j < 2 * nq) {
              const float4 t = xv[(j - nq) < kQ && (j - nq) >= 0 ? (j - nq) : 0];
              xx = make_float4(t.x * t.x, t.y * t.y, t.z * t.z, t.w * t.w);
            }
            float h, l;
            split_tf32_fast(xx.x, h, l); vh[q * 4 + 0] = __float_as_uint(h); vl[q * 4 + 0] = __float_as_uint(l);
            split_tf32_fast(xx.y, h, l); vh[q * 4 + 1] = __float_as_uint(h); vl[q * 4 + 1] = __float_as_uint(l);
            split_tf32_fast(xx.z, h, l); vh[q * 4 + 2] = __float_as_uint(h); vl[q * 4 + 2] = __float_as_uint(l);
            split_tf32_fast(xx.w, h, l); vh[q * 4 + 3] = __float_as_uint(h); vl[q * 4 + 3] = __float_as_uint(l);
          }
          tmem_st16(xa + c * 16, vh);
          tmem_st16(xa + 80 + c * 16, vl);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      fence_async_smem();  // the W image (generic-proxy writes) -> visible to the tensor core
      mbar_arrive(&full[s]);
      if (warp == 4) stamp(i, 2);
#pragma unroll
      for (int j = 0; j < kQ; j++) xv[j] = xn[j];
      d0 = d1; d1 = d2; d2 = d3; f1 = f2;
    }
  } else if (warp == 8) {
    // =================================== MMA ISSUER ===================================
    const uint32_t idesc = make_idesc_tf32(kTcRows, TN);
    for (int ui = u_begin, i = 0; ui < u_end; ui++, i++) {
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      mbar_wait(&full[s], ph);          // operands landed
      stamp(i, 3);
      mbar_wait(&dempty[s], ph ^ 1);    // accumulator stage drained by the epilogue (unit i-2)
      stamp(i, 4);
      tc_fence_after();
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem0, 0);
      if (elect_one_sync()) {
        const uint32_t xh = tb + (uint32_t)s * 160, xl = xh + 80;
        const uint64_t wh = make_smem_desc2(smem_u32(Ws), 128, P), wl = make_smem_desc2(smem_u32(Ws) + (uint32_t)(TN / 8) * P, 128, P);
        const uint32_t d = tb + 320 + (uint32_t)s * 96;
        uint32_t acc = 0;
        for (int p = 0; p < 3; p++) {  // Xh*Wh, Xl*Wh, Xh*Wl
          const uint32_t a0 = (p == 1) ? xl : xh;
          const uint64_t b0 = (p == 2) ? wl : wh;
          for (int j = 0; j < NSLAB; j++) {
            tc_mma_tf32_ts(d, a0 + j * 8, b0 + (uint64_t)(j * 16), idesc, acc);  // +256 B per K-step, in 16-byte units
            acc = 1;
          }
        }
        tc_commit(&empty[s]);   // operand stage (and, for the loaders' image switch, W) free when these MMAs retire
        tc_commit(&dfull[s]);   // accumulator ready
      }
      __syncwarp();
      stamp(i, 5);
    }
  } else {
    // =================================== EPILOGUE ===================================
    const int row = 32 * warp + lane;  // warps 0..3 <-> TMEM lanes 32w..32w+31
    const uint32_t trow = (uint32_t)(32 * warp) << 16;
    for (int ui = u_begin, i = 0; ui < u_end; ui++, i++) {
      const TcTile unit = get_unit(ui);
      const int s = i & 1;
      mbar_wait(&dfull[s], (i >> 1) & 1);
      if (warp == 0) stamp(i, 6);
      tc_fence_after();
      const int st_lim = TRAIN ? N : S_total;
      const int nst = max(0, min(SCt, st_lim - unit.state0));  // states present in this image
      const int mp = MP ? MP : M;  // M here is the padded count
      const int ncols = nst * mp;
      const bool live = row < unit.nrows;
      int64_t f = 0;
      if (live) f = TRAIN ? (int64_t)frame_ids[unit.row0 + row] : fbase + unit.row0 + row;
      float *lrow = TRAIN ? logb + f * N + unit.state0 : logb + (f - fbase) * ldb + unit.state0;
      const uint32_t d = tmem0 + 320 + (uint32_t)s * 96 + trow;
      // kc2 of this unit's image from global memory (L1-resident, same address for the whole warp): the
      // shared-memory image may already belong to a later unit
      const float4 *kc4 = reinterpret_cast<const float4 *>(images + (size_t)unit.img * (img_bytes / 4) + w_bytes / 4);
      float mx = kNegInf, sum = 0.f;  // running state (MP == 0 only)
      int chunks = 0, st = 0;
      constexpr int kMaxCh = kWsMaxTN / 16;
      uint32_t v[kMaxCh][16];
#pragma unroll
      for (int c = 0; c < kMaxCh; c++)  // every accumulator column of my frame in flight at once
        if (c * 16 < ncols) tmem_ld16_nowait(d + c * 16, v[c]);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < kMaxCh; c++) {
        if (c * 16 < ncols) {
          const float4 k0 = __ldg(kc4 + c * 4), k1 = __ldg(kc4 + c * 4 + 1), k2 = __ldg(kc4 + c * 4 + 2), k3 = __ldg(kc4 + c * 4 + 3);
          const float kc[16] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w, k2.x, k2.y, k2.z, k2.w, k3.x, k3.y, k3.z, k3.w};
          float val[16];  // log2(c_g N_g(x)); -inf for a Gaussian of density 0 and for the pad columns
#pragma unroll
          for (int j = 0; j < 16; j++) val[j] = fmaf(__uint_as_float(v[c][j]), 1.4426950408889634f, kc[j]);
          if (MP) {  // 16 / MP whole states in this chunk, each reduced independently
#pragma unroll
            for (int g = 0; g < 16 / (MP ? MP : 16); g++) {
              float m = val[g * MP];
#pragma unroll
              for (int j = 1; j < MP; j++) m = fmaxf(m, val[g * MP + j]);
              const float ms = (m > kNegInf) ? m : 0.f;
              float sm_ = 0.f;
#pragma unroll
              for (int j = 0; j < MP; j++) sm_ += ex2_approx(val[g * MP + j] - ms);
              const float lb = (m > kNegInf) ? (ms + __log2f(sm_)) * 0.6931471805599453f : kNegInf;
              if (live && st + g < nst) lrow[st + g] = lb;
            }
            st += 16 / (MP ? MP : 16);
          } else {  // one state spans M/16 chunks: online log-sum-exp across chunks
            float m = val[0];
#pragma unroll
            for (int j = 1; j < 16; j++) m = fmaxf(m, val[j]);
            const float mn = fmaxf(mx, m);
            const float ms = (mn > kNegInf) ? mn : 0.f;
            float sm_ = 0.f;
#pragma unroll
            for (int j = 0; j < 16; j++) sm_ += ex2_approx(val[j] - ms);
            sum = fmaf(sum, ex2_approx(((mx > kNegInf) ? mx : ms) - ms), sm_);
            mx = mn;
            if (++chunks == mp / 16) {
              const float lb = (mx > kNegInf) ? (mx + __log2f(sum)) * 0.6931471805599453f : kNegInf;
              if (live) lrow[st] = lb;
              st++; chunks = 0; mx = kNegInf; sum = 0.f;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&dempty[s]);
      if (warp == 0) stamp(i, 7);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc(tmem0, 512);
}


// ================================================================================================
// k_accum_ws: mixture accumulators (calc_mix_param, T-FS:1691-1727), warp-specialised and pipelined.
// Same mathematics as k_accum_tc (tc_kernels.cuh): per sub-tile of 64 frames of one model and block of
// 128 Gaussians
//   GEMM1  L[g][f]  = sum_k W[g][k] Xaug[f][k]           A = W (TMEM, parked per image), B = X  (smem)
//   w[g][f] = gamma_f(s(g)) exp(L + kc[g] - logb_f(s(g)))  epilogue, in place in TMEM (hi) + beside it (lo)
//   GEMM2  S[g][k] += sum_f w[g][f] Xaug[f][k]           A = w (TMEM), B = XT (smem)
// but the activities run concurrently on different warps and meet only at mbarriers (17 warps):
//   warps  8-11 X LOADERS   features three levels ahead in registers; X (frames x columns), TF32 hi / lo,
//                           and the per-frame weight exponents cfs
//   warps 12-15 XT LOADERS  the same tile transposed (columns x frames), TF32 hi / lo
//   warp  16    MMA         issues whichever of GEMM1(next unit) / GEMM2(oldest unit) has its inputs,
//                           so neither the loaders nor the epilogue wait on the other's hand-off
//   warps  0-7  EPILOGUE    L -> w with tcgen05.ld / tcgen05.st (warp w: TMEM lanes 32 (w%4).., frames
//                           32 (w/4)..); every kAccDrain sub-tiles S moves from TMEM (FP32, truncating
//                           accumulation) to FP32 registers (round to nearest, 40 columns per thread);
//                           double atomics when the CTA's (model, Gaussian block) changes; loads the next
//                           image's W into TMEM
// Two shared-memory stages (X, XT, cfs) and two TMEM stages (L / w_hi, w_lo).
// TMEM columns: [0,128) L / w_hi x2 | [128,256) w_lo x2 | [256, 256+KP2) S | [352, 352+2KP) W_hi, W_lo
// Shared-memory layouts (SWIZZLE_NONE K-major, 16-byte chunks):
//   X  : byte(f, k) = (f%8)*16 + (k%4)*4 + (k/4)*128 + (f/8)*PX     PX = (KP/4)*128   (8 frame groups)
//   XT : byte(k, f) = (k%8)*16 + (f%4)*4 + (f/4)*128 + (k/8)*2048   (KP2/8 column groups)
// ================================================================================================
constexpr int kAccSub = 64;        // frames per sub-tile
constexpr int kAccDrain = 8;       // sub-tiles accumulated in TMEM before S moves to registers
constexpr int kAccWsThreads = 544; // 17 warps

__host__ __device__ inline size_t ws_acc_stage_bytes(int KP) { return (size_t)2 * 8 * (KP / 4) * 128 + (size_t)2 * (tc_kp2(KP) / 8) * 2048 + 64 * 8 * 4; }
__host__ __device__ inline size_t ws_acc_smem_bytes(int KP) { return 2 * ws_acc_stage_bytes(KP) + 1024 + 256; }

__global__ void __launch_bounds__(kAccWsThreads, 1)
k_accum_ws(const TcTile *__restrict__ units, int nunits, const int32_t *__restrict__ frame_ids, const float *__restrict__ x32,
           const float *__restrict__ images, const float *__restrict__ kcT, const float *__restrict__ logb,
           const float *__restrict__ gamma, int N, int M, int G, int D, int DP, double *__restrict__ stats,
           int64_t stats_stride, int64_t off_S0, int64_t off_S1, int64_t off_S2, long long *__restrict__ tdbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  auto stamp = [&](int i, int slot) {  // latest arrival over the warps of a role
    if (tdbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && i < 64) atomicMax((unsigned long long *)&tdbg[i * 8 + slot], (unsigned long long)clock64());
  };
  const int KP = 2 * DP, KP2 = tc_kp2(KP), NSLAB = KP / 8, nq = DP / 4;
  const uint32_t PX = (uint32_t)(KP / 4) * 128;
  const uint32_t x_bytes = 8 * PX, xt_bytes = (uint32_t)(KP2 / 8) * 2048;
  const uint32_t stage_bytes = 2 * x_bytes + 2 * xt_bytes + 64 * 8 * 4;
  uint8_t *sm = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm + 2 * stage_bytes);
  uint64_t *x_full = bars, *x_free = bars + 2, *d1_full = bars + 4, *w_full = bars + 6, *s_full = bars + 8, *s_free = bars + 9,
           *wimg_full = bars + 10;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 11);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(&x_full[s], 256);   // X and XT loaders
      mbar_init(&x_free[s], 1);     // tcgen05.commit after GEMM2
      mbar_init(&d1_full[s], 1);    // tcgen05.commit after GEMM1
      mbar_init(&w_full[s], 256);   // epilogue threads
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(wimg_full, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;

  const int per = (nunits + gridDim.x - 1) / gridDim.x;
  const int u_begin = blockIdx.x * per, u_end = min(nunits, u_begin + per);
  const int n_my = max(0, u_end - u_begin);
  auto img_at = [&](int ui) -> int { return (ui >= u_begin && ui < u_end) ? __ldg(&units[ui].img) : -1; };
  auto unit_at = [&](int ui) -> TcTile { return ui < u_end ? units[ui] : TcTile{0, 0, -1, 0, 0, 0}; };

  if (warp >= 8 && warp < 12) {
    // =================================== X LOADERS ===================================
    const int t = tid - 256;
    const int xr = t & 63, xq0 = (t >> 6) * ((nq + 1) / 2);  // row xr, float4 [xq0, xq1)
    const int xq1 = min(nq, xq0 + (nq + 1) / 2);
    constexpr int kXQ = 5;                                   // DP <= 40
    struct Pre { float4 x[kXQ]; float gm[8], lb[8]; };
    auto fid_of = [&](const TcTile &u) -> int { return (xr < u.nrows) ? __ldg(frame_ids + u.row0 + xr) : -1; };
    auto load_pre = [&](int f, Pre &p) {
      const float4 *src = reinterpret_cast<const float4 *>(x32 + (int64_t)(f < 0 ? 0 : f) * DP);
#pragma unroll
      for (int j = 0; j < kXQ; j++) p.x[j] = (f >= 0 && xq0 + j < xq1) ? __ldg(src + xq0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const bool ok = t < 64 && f >= 0 && s < N;
        p.gm[s] = ok ? __ldg(gamma + (int64_t)f * N + s) : 0.f;
        p.lb[s] = ok ? __ldg(logb + (int64_t)f * N + s) : 0.f;
      }
    };
    TcTile d1 = unit_at(u_begin + 1), d2 = unit_at(u_begin + 2);
    int f1 = fid_of(d1);
    Pre cur, nxt;
    load_pre(fid_of(unit_at(u_begin)), cur);
    const uint32_t rbase = (uint32_t)(xr & 7) * 16 + (uint32_t)(xr >> 3) * PX;
    for (int i = 0; i < n_my; i++) {
      const int s = i & 1;
      load_pre(f1, nxt);                               // operands of unit i+1
      const int f2 = fid_of(d2);                       // frame id of unit i+2
      const TcTile d3 = unit_at(u_begin + i + 3);      // descriptor of unit i+3
      if (warp == 8) stamp(i, 0);
      mbar_wait(&x_free[s], ((i >> 1) & 1) ^ 1);       // stage s: GEMM2 of unit i-2 has retired
      if (warp == 8) stamp(i, 1);
      const uint32_t Xh = smem_u32(sm) + (uint32_t)s * stage_bytes, Xl = Xh + x_bytes;
      float *cfs = reinterpret_cast<float *>(sm + (size_t)s * stage_bytes + 2 * x_bytes + 2 * xt_bytes);
#pragma unroll
      for (int j = 0; j < kXQ; j++) {
        if (xq0 + j < xq1) {
          const float4 xx = cur.x[j];
          float4 h, l;
          split_tf32_fast(xx.x, h.x, l.x); split_tf32_fast(xx.y, h.y, l.y); split_tf32_fast(xx.z, h.z, l.z); split_tf32_fast(xx.w, h.w, l.w);
          uint32_t o = rbase + (uint32_t)(xq0 + j) * 128;
          st_shared_v4(Xh + o, h);
          st_shared_v4(Xl + o, l);
          split_tf32_fast(xx.x * xx.x, h.x, l.x); split_tf32_fast(xx.y * xx.y, h.y, l.y);
          split_tf32_fast(xx.z * xx.z, h.z, l.z); split_tf32_fast(xx.w * xx.w, h.w, l.w);
          o += (uint32_t)nq * 128;
          st_shared_v4(Xh + o, h);
          st_shared_v4(Xl + o, l);
        }
      }
      if (t < 64) {  // weight exponent of frame t per state: log2(gamma) - logb log2(e); -inf = no weight
#pragma unroll
        for (int st = 0; st < 8; st++) {
          float cf = kNegInf;
          if (st < N && cur.gm[st] > 0.f && cur.lb[st] > kNegInf) cf = __log2f(cur.gm[st]) - cur.lb[st] * 1.4426950408889634f;
          cfs[t * 8 + st] = cf;
        }
      }
      fence_async_smem();
      mbar_arrive(&x_full[s]);
      stamp(i, 2);
      cur = nxt;
      d1 = d2; d2 = d3; f1 = f2;
    }
  } else if (warp >= 12 && warp < 16) {
    // =================================== XT LOADERS ===================================
    const int t = tid - 384;
    const int fg = t >> 3, nl = t & 7;  // frames 4fg..4fg+3, columns nl + 8k
    constexpr int kTK = 5;              // DP <= 40
    struct Fids { int ft[4]; };
    auto fids_of = [&](const TcTile &u) -> Fids {
      Fids f;
#pragma unroll
      for (int j = 0; j < 4; j++) f.ft[j] = (4 * fg + j < u.nrows) ? __ldg(frame_ids + u.row0 + 4 * fg + j) : -1;
      return f;
    };
    struct Pre { float xt[kTK][4]; };
    auto load_pre = [&](const Fids &f, Pre &p) {
#pragma unroll
      for (int k = 0; k < kTK; k++)
#pragma unroll
        for (int j = 0; j < 4; j++)
          p.xt[k][j] = (f.ft[j] >= 0 && nl + 8 * k < DP) ? __ldg(x32 + (int64_t)f.ft[j] * DP + nl + 8 * k) : 0.f;
    };
    TcTile d1 = unit_at(u_begin + 1), d2 = unit_at(u_begin + 2);
    Fids f1 = fids_of(d1);
    Pre cur, nxt;
    load_pre(fids_of(unit_at(u_begin)), cur);
    for (int i = 0; i < n_my; i++) {
      const int s = i & 1;
      load_pre(f1, nxt);
      const Fids f2 = fids_of(d2);
      const TcTile d3 = unit_at(u_begin + i + 3);
      mbar_wait(&x_free[s], ((i >> 1) & 1) ^ 1);
      const uint32_t XTh = smem_u32(sm) + (uint32_t)s * stage_bytes + 2 * x_bytes, XTl = XTh + xt_bytes;
#pragma unroll
      for (int k = 0; k < kTK; k++) {  // rows n (x) and n + DP (x^2), 16-byte chunk = frames 4fg..4fg+3
        const int n = nl + 8 * k;
        if (n < DP) {
          float4 h, l;
          split_tf32_fast(cur.xt[k][0], h.x, l.x); split_tf32_fast(cur.xt[k][1], h.y, l.y);
          split_tf32_fast(cur.xt[k][2], h.z, l.z); split_tf32_fast(cur.xt[k][3], h.w, l.w);
          uint32_t o = (uint32_t)(n & 7) * 16 + (uint32_t)fg * 128 + (uint32_t)(n >> 3) * 2048;
          st_shared_v4(XTh + o, h);
          st_shared_v4(XTl + o, l);
          split_tf32_fast(cur.xt[k][0] * cur.xt[k][0], h.x, l.x); split_tf32_fast(cur.xt[k][1] * cur.xt[k][1], h.y, l.y);
          split_tf32_fast(cur.xt[k][2] * cur.xt[k][2], h.z, l.z); split_tf32_fast(cur.xt[k][3] * cur.xt[k][3], h.w, l.w);
          const int n2 = n + DP;
          o = (uint32_t)(n2 & 7) * 16 + (uint32_t)fg * 128 + (uint32_t)(n2 >> 3) * 2048;
          st_shared_v4(XTh + o, h);
          st_shared_v4(XTl + o, l);
        }
      }
      if (KP2 > KP && nl == 0) {  // pad rows of GEMM2's N
        for (int n = KP; n < KP2; n++) {
          const uint32_t o = (uint32_t)(n & 7) * 16 + (uint32_t)fg * 128 + (uint32_t)(n >> 3) * 2048;
          st_shared_v4(XTh + o, make_float4(0.f, 0.f, 0.f, 0.f));
          st_shared_v4(XTl + o, make_float4(0.f, 0.f, 0.f, 0.f));
        }
      }
      fence_async_smem();
      mbar_arrive(&x_full[s]);
      stamp(i, 2);
      cur = nxt;
      d1 = d2; d2 = d3; f1 = f2;
    }
  } else if (warp == 16) {
    // =================================== MMA ISSUER ===================================
    const uint32_t idesc1 = make_idesc_tf32(128, kAccSub), idesc2 = make_idesc_tf32(128, KP2);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem0, 0);
    // GEMM1(i) needs x_full(i) (and the image's W); GEMM2(j) needs w_full(j) (and S drained).  Whichever
    // is ready goes first; GEMM1 may run at most one unit ahead of GEMM2 (two TMEM / smem stages).
    int cnt = 0, ndrain = 0, nimg = 0, g1 = 0, g2 = 0;
    bool need_s_free = false;
    // image boundaries of the units at the two cursors, kept in registers (one load per advance)
    int img_g1m = -1, img_g1 = img_at(u_begin), img_g2 = img_g1, img_g2p = img_at(u_begin + 1);
    auto issue_g1 = [&](int i) {
      const int s = i & 1;
      tc_fence_after();
      if (elect_one_sync()) {
        const uint8_t *Xh = sm + (size_t)s * stage_bytes;
        const uint64_t bh = make_smem_desc2(smem_u32(Xh), 128, PX), bl = make_smem_desc2(smem_u32(Xh + x_bytes), 128, PX);
        uint32_t accf = 0;
        for (int p = 0; p < 3; p++) {  // Wh*Xh, Wl*Xh, Wh*Xl
          const uint32_t a0 = tb + kAccTmW + ((p == 1) ? KP : 0);
          const uint64_t b0 = (p == 2) ? bl : bh;
          for (int j = 0; j < NSLAB; j++) {
            tc_mma_tf32_ts(tb + (uint32_t)s * 64, a0 + j * 8, b0 + (uint64_t)(j * 16), idesc1, accf);
            accf = 1;
          }
        }
        tc_commit(&d1_full[s]);
      }
      __syncwarp();
    };
    auto issue_g2 = [&](int j) {
      const int sj = j & 1;
      tc_fence_after();
      cnt++;
      const bool drain = img_g2p != img_g2 || cnt == kAccDrain;  // last unit of its image (or of this CTA), or S is due
      if (elect_one_sync()) {
        const uint8_t *XTh = sm + (size_t)sj * stage_bytes + 2 * x_bytes;
        const uint64_t bh = make_smem_desc2(smem_u32(XTh), 128, 2048), bl = make_smem_desc2(smem_u32(XTh + xt_bytes), 128, 2048);
        uint32_t accf = cnt > 1 ? 1u : 0u;
        for (int p = 0; p < 3; p++) {  // wh*Xh, wl*Xh, wh*Xl
          const uint32_t a0 = tb + (uint32_t)sj * 64 + ((p == 1) ? 128 : 0);
          const uint64_t b0 = (p == 2) ? bl : bh;
          for (int k = 0; k < kAccSub / 8; k++) {
            tc_mma_tf32_ts(tb + 256, a0 + k * 8, b0 + (uint64_t)(k * 16), idesc2, accf);
            accf = 1;
          }
        }
        tc_commit(&x_free[sj]);
        if (drain) tc_commit(s_full);
      }
      __syncwarp();
      if (drain) { ndrain++; need_s_free = true; cnt = 0; }
    };
    while (g2 < n_my) {
      bool did = false;
      const bool first_g1 = img_g1 != img_g1m;
      // a new image's first GEMM1 only after the old image's last GEMM2 has been issued
      if (g1 < n_my && g1 - g2 <= 1 && (!first_g1 || g1 == g2)) {
        bool ok = mbar_try(&x_full[g1 & 1], (g1 >> 1) & 1);
        if (ok && first_g1) ok = mbar_try(wimg_full, nimg & 1);
        ok = __all_sync(0xffffffffu, ok);
        if (ok) {
          if (first_g1) nimg++;
          stamp(g1, 3);
          issue_g1(g1);
          stamp(g1, 4);
          g1++;
          img_g1m = img_g1;
          img_g1 = img_at(u_begin + g1);
          did = true;
        }
      }
      if (!did && g2 < g1) {
        bool ok = mbar_try(&w_full[g2 & 1], (g2 >> 1) & 1);
        if (ok && need_s_free) ok = mbar_try(s_free, (ndrain - 1) & 1);
        ok = __all_sync(0xffffffffu, ok);
        if (ok) {
          need_s_free = false;
          issue_g2(g2);
          stamp(g2, 5);
          g2++;
          img_g2 = img_g2p;
          img_g2p = img_at(u_begin + g2 + 1);
        }
      }
    }
  } else {
    // =================================== EPILOGUE (warps 0-7) ===================================
    const int q = warp & 3, hb = warp >> 2;  // TMEM lane quarter; frame half (and column half of S)
    const int row = 32 * q + lane;
    const uint32_t trow = (uint32_t)(32 * q) << 16;
    constexpr int kMaxCol = 40;  // KP2 <= 80 (tc_acc_fits): 5 groups of 8 columns per thread
    const int nc8 = KP2 / 8;
    const int c8_beg = hb ? (nc8 + 1) / 2 : 0, c8_end = hb ? nc8 : (nc8 + 1) / 2;
    float acc[kMaxCol];
#pragma unroll
    for (int k = 0; k < kMaxCol; k++) acc[k] = 0.f;
    int cnt = 0, ndrain = 0;
    float kcr = kNegInf;
    int st = 0, cur_v = 0, cur_rb = 0;
    for (int i = 0; i < n_my; i++) {
      const int ui = u_begin + i, s = i & 1;
      const int img = img_at(ui);
      const bool first = (i == 0) || img != img_at(ui - 1);
      const bool last = (i == n_my - 1) || img != img_at(ui + 1);
      if (first) {  // W (hi | lo) of this (model, Gaussian block) -> TMEM, one Gaussian per lane; halves split the columns
        const TcTile unit = units[ui];
        cur_v = unit.v; cur_rb = unit.pad;
        const float *im = images + (size_t)img * (tc_accT_image_bytes(KP) / 4) + (size_t)row * 2 * KP;
        for (int ch = hb; ch < 2 * KP / 16; ch += 2) {
          uint32_t r[16];
#pragma unroll
          for (int qq = 0; qq < 4; qq++) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(im + ch * 16 + qq * 4));
            r[qq * 4 + 0] = __float_as_uint(a.x); r[qq * 4 + 1] = __float_as_uint(a.y); r[qq * 4 + 2] = __float_as_uint(a.z); r[qq * 4 + 3] = __float_as_uint(a.w);
          }
          tmem_st16(tmem0 + kAccTmW + trow + ch * 16, r);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(wimg_full);
        const int g = cur_rb * 128 + row;
        kcr = (g < G) ? __ldg(kcT + (size_t)img * 128 + row) : kNegInf;
        st = min(g / M, 7);
      }
      mbar_wait(&d1_full[s], (i >> 1) & 1);
      if (warp == 0) stamp(i, 6);
      tc_fence_after();
      const float *cfs = reinterpret_cast<const float *>(sm + (size_t)s * stage_bytes + 2 * x_bytes + 2 * xt_bytes);
      {  // my 32 frames: accumulator columns [32 hb, 32 hb + 32)
        uint32_t v[2][16];
#pragma unroll
        for (int c = 0; c < 2; c++) tmem_ld16_nowait(tmem0 + (uint32_t)s * 64 + trow + (hb * 2 + c) * 16, v[c]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int c0 = (hb * 2 + c) * 16;
          uint32_t vl[16];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const float y = fmaf(__uint_as_float(v[c][j]), 1.4426950408889634f, kcr + cfs[(c0 + j) * 8 + st]);
            const float w = ex2_approx(y);  // ex2(-inf) = +0 and underflow flushes to 0: no weight; y is never NaN
            float h, l;
            split_tf32_fast(w, h, l);
            v[c][j] = __float_as_uint(h);
            vl[j] = __float_as_uint(l);
          }
          tmem_st16(tmem0 + (uint32_t)s * 64 + trow + c0, v[c]);
          tmem_st16(tmem0 + 128 + (uint32_t)s * 64 + trow + c0, vl);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&w_full[s]);
      stamp(i, 7);
      cnt++;
      if (last || cnt == kAccDrain) {  // S: TMEM (FP32, truncating) -> registers (round to nearest)
        mbar_wait(s_full, ndrain & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kMaxCol / 8; c++) {
          if (c8_beg + c < c8_end) {
            float v[8];
            tmem_ld8(tmem0 + 256 + trow + (c8_beg + c) * 8, v);
#pragma unroll
            for (int j = 0; j < 8; j++) acc[c * 8 + j] += v[j];
          }
        }
        tc_fence_before();
        mbar_arrive(s_free);
        ndrain++; cnt = 0;
        if (last) {  // the CTA leaves this (model, Gaussian block): registers -> statistics
          const int g = cur_rb * 128 + row;
          double *stp = stats + (int64_t)cur_v * stats_stride;
#pragma unroll
          for (int c = 0; c < kMaxCol / 8; c++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const int k = (c8_beg + c) * 8 + j;
              if (c8_beg + c < c8_end && g < G) {
                const double a = (double)acc[c * 8 + j];
                if (k < D) atomicAdd(stp + off_S1 + (int64_t)g * D + k, a);
                else if (k == D) atomicAdd(stp + off_S0 + g, a);
                else if (k >= DP && k < DP + D) atomicAdd(stp + off_S2 + (int64_t)g * D + (k - DP), a);
              }
              acc[c * 8 + j] = 0.f;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 16) tmem_dealloc(tmem0, 512);
}

}  // namespace hmmk
