// fb_kernels.cuh -- scaled forward / backward for Baum-Welch training (calc_alpha, calc_beta,
// calc_transition_probab, calc_den_mix_coef, calc_probability; T-FS:1380-1664), latency-first.
//
// The recursion over time is a dependent chain; what bounds it is the latency of one step, not
// throughput.  So one THREAD owns one utterance and keeps the whole state vector in registers
// (no shuffles on the chain), the forward and the backward chains of an utterance run concurrently in
// two different warps, and each step rescales by an exact power of two taken from the largest exponent
// field of the step's components (integer instructions only; no division, no logarithm, no sum):
//   forward :  z_t = ((z_{t-1} A) o b~_t) 2^-e_t      b~_i(t) = exp(logb_i(t) - m_t), m_t = max_i logb_i(t)
//   backward:  w_t = (A (b~_{t+1} o w_{t+1})) 2^-e'_t  w_{T-1} = [0,..,0,1]   (final state only, T-FS:1484)
// Any positive per-frame scaling gives the same posteriors; the reference's c_t-scaled quantities
// are recovered exactly from per-frame normalisation:
//   alpha^_t(i) beta^_t(i) / c_t          = phi * z_t(i) w_t(i) / sum_j z_t(j) w_t(j)
//   alpha^_t(i) a_ij b_j(t+1) beta^_t+1(j) = phi * z_t(i) a_ij q_j / sum_kl z_t(k) a_kl q_l,  q = b~_{t+1} o w_{t+1}
// with phi = alpha^_{T-1}(N-1) (the reference's beta^ starts from the final state only, so its
// gammas sum to phi, not to 1), and
//   log P = sum_t m_t + ln2 sum_t e_t + log z_{T-1}(N-1)      (calc_probability, T-FS:1546-1549).
// The chain threads never touch global memory for their operands: the six other warps of the CTA stage
// the NEXT window of kFbWin frames of every utterance of the CTA into shared memory as b~ (double) while
// the chains consume the current one (two buffers) -- forward window from the start of the utterance,
// backward window from its end -- so a chain step is NS LDS.64, the banded matvec, the power-of-two
// scaling and two 16-byte stores of the scaled vector in single precision (rows of 8 floats).
// A second phase (one warp per utterance, lanes over frames) forms gamma, the transition sums and sum m_t.
#pragma once
#include "kernels.cuh"

namespace hmmk {

constexpr int kFbUtts = 8;        // utterances per CTA
constexpr int kFbThreads = 256;   // 8 warps: warp 0 = forward chains, warp 1 = backward chains, warps 2-7 stage; then all combine
constexpr int kFbWin = 64;        // frames per staged window
constexpr int kFbRow = 8;         // floats per row of the alpha / beta workspaces (NS <= 8)

// exp(y) for y <= 0 as a double: single-precision mantissa accuracy (the log-densities it is fed
// are single precision), double-precision range.  y < -700 -> 0.
__device__ __forceinline__ double exp_scaled(float y) {
  // Only ONE operation on the transcendental / conversion unit (ex2): the rounding of y log2(e) to an integer uses
  // the 1.5 * 2^23 trick on the FMA pipe, and the double is assembled from the float's bits with integer
  // instructions (a cvt.f64.f32 and a cvt.rni per value made the decode scorers wait on that unit).
  const float yc = fmaxf(y, -700.f);  // branch-free: the chain's five exponentials must overlap
  const float t = fmaf(yc, 1.4426950408889634f, 12582912.f);         // 1.5 * 2^23 + rint(yc log2 e)
  const float n = t - 12582912.f;
  const int ni = __float_as_int(t) - 0x4B400000;                      // the same integer
  float f = fmaf(yc, 1.4426950216293335f, -n);  // (float)log2(e)
  f = fmaf(yc, 1.9259629911e-8f, f);            // log2(e) - (float)log2(e)
  float p;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(f));  // |f| <= 0.5: 2 ulp; p in [0.70, 1.42] is a normal float
  const unsigned pb = (unsigned)__float_as_int(p);
  const int hi = (int)(pb >> 3) + ((1023 - 127 + ni) << 20);          // exponent re-biased, 20 of the 23 mantissa bits
  const int lo = (int)(pb << 29);                                     // the other 3
  return (y < -700.f) ? 0.0 : __hiloint2double(hi, lo);
}

// r = 2^-e with e the unbiased exponent of s (so that s*r is in [1,2)); e returned.  s == 0, denormal,
// inf or NaN: r = 1, e = 0.
__device__ __forceinline__ double pow2_scale(double s, int &e) {
  const int be = (__double2hiint(s) >> 20) & 0x7ff;
  if (be == 0 || be == 0x7ff) { e = 0; return 1.0; }
  e = be - 1023;
  return __hiloint2double((2046 - be) << 20, 0);
}

template <int NS>
__device__ __forceinline__ void load_lb(const float *__restrict__ p, float (&l)[NS]) {
#pragma unroll
  for (int i = 0; i < NS; i++) l[i] = __ldg(p + i);
}

__host__ __device__ inline size_t fb_smem_bytes(int NS) {
  return (size_t)2 * 2 * kFbUtts * (kFbWin * NS + 2) * sizeof(double);  // [buffer][direction][utterance][window]
}

// 2^-(E - 1023) for the largest exponent field E of the non-negative components v[]; 1 when that field is 0
// (all zero / denormal) or 0x7ff (inf / NaN).  e = E - 1023 (0 in the degenerate cases).
template <int NS>
__device__ __forceinline__ double pow2_scale_max(const double (&v)[NS], int &e) {
  int be = __double2hiint(v[0]) >> 20;
#pragma unroll
  for (int i = 1; i < NS; i++) be = max(be, __double2hiint(v[i]) >> 20);
  const bool ok = be > 0 && be < 0x7ff;
  e = ok ? be - 1023 : 0;
  return __hiloint2double(ok ? (2046 - be) << 20 : 0x3ff00000, 0);
}

template <int NS>
__device__ __forceinline__ void store_row(float *__restrict__ p, const double (&z)[NS]) {
  float f[kFbRow];
#pragma unroll
  for (int i = 0; i < kFbRow; i++) f[i] = (i < NS) ? (float)z[i] : 0.f;
  *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
  if (NS > 4) *reinterpret_cast<float4 *>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
template <int NS>
__device__ __forceinline__ void load_row(const float *__restrict__ p, double (&z)[NS]) {
  const float4 a = *reinterpret_cast<const float4 *>(p);
  float f[kFbRow] = {a.x, a.y, a.z, a.w, 0.f, 0.f, 0.f, 0.f};
  if (NS > 4) {
    const float4 b = *reinterpret_cast<const float4 *>(p + 4);
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
#pragma unroll
  for (int i = 0; i < NS; i++) z[i] = (double)f[i];
}

// Outputs: gamma32[F][N] (the reference's alpha^ beta^ / c, T-FS:1709); per-model statistics head
// (num_trans, den_trans, den_mix, sum_logp, n_utt) by double atomics, one set per utterance;
// logp_utt[U] (0 for masked utterances).  alpha_ws / beta_ws: float [F][kFbRow] workspaces.
template <int NS, bool BANDED>
__global__ void __launch_bounds__(kFbThreads)
k_fb(const float *__restrict__ logb, const int64_t *__restrict__ off, const int32_t *__restrict__ u2m,
     const double *__restrict__ Aall, int U, float *__restrict__ alpha_ws, float *__restrict__ beta_ws,
     float *__restrict__ gamma, double *__restrict__ stats, int64_t stats_stride, int64_t off_sumlogp,
     double *__restrict__ logp_utt) {
  extern __shared__ __align__(16) uint8_t fb_smem[];
  constexpr int US = kFbWin * NS + 2;  // per-utterance stride, padded: the 8 chain lanes hit 8 different banks
  double *bf = reinterpret_cast<double *>(fb_smem);  // [2 buffers][2 directions][kFbUtts][US]  b~
  __shared__ double sA[kFbUtts][NS * NS];
  __shared__ double sphi[kFbUtts], slp[kFbUtts];
  __shared__ int64_t sbase[kFbUtts];
  __shared__ int sT[kFbUtts];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = blockIdx.x * kFbUtts;
  for (int idx = tid; idx < kFbUtts * NS * NS; idx += kFbThreads) {
    const int uu = idx / (NS * NS), k = idx - uu * NS * NS;
    const int u = u0 + uu;
    const int v = (u < U) ? u2m[u] : -1;
    sA[uu][k] = (v >= 0) ? Aall[(int64_t)v * NS * NS + k] : 0.0;
  }
  if (tid < kFbUtts) {
    const int u = u0 + tid;
    const bool live = u < U && u2m[u] >= 0;
    sbase[tid] = live ? off[u] : 0;
    sT[tid] = live ? (int)(off[u + 1] - off[u]) : 0;
  }
  __syncthreads();
  int Tmax = 0;
#pragma unroll
  for (int uu = 0; uu < kFbUtts; uu++) Tmax = max(Tmax, sT[uu]);

  // Staging of window w0 into buffer `buf` by `nthr` threads (index t0): item <-> (direction, utterance, frame).
  // Three items per pass so that their loads are in flight together.
  auto stage = [&](int w0, int buf, int t0, int nthr) {
    constexpr int NIT = 2 * kFbUtts * kFbWin, PASS = 3;
    double *dstb = bf + (size_t)buf * 2 * kFbUtts * US;
    for (int it0 = t0; it0 < NIT; it0 += PASS * nthr) {
      float l[PASS][NS];
      double *dst[PASS];
#pragma unroll
      for (int p = 0; p < PASS; p++) {
        const int it = it0 + p * nthr;
        dst[p] = nullptr;
        if (it < NIT) {
          const int dir = it / (kFbUtts * kFbWin), rem = it - dir * (kFbUtts * kFbWin);
          const int uu = rem / kFbWin, k = rem - uu * kFbWin;
          const int T = sT[uu];
          if (w0 + k < T) {
            const int t = dir == 0 ? w0 + k : T - 1 - (w0 + k);
            load_lb<NS>(logb + (sbase[uu] + t) * NS, l[p]);
            dst[p] = dstb + (size_t)(dir * kFbUtts + uu) * US + k * NS;
          }
        }
      }
#pragma unroll
      for (int p = 0; p < PASS; p++) {
        if (dst[p]) {
          float m = l[p][0];
#pragma unroll
          for (int i = 1; i < NS; i++) m = fmaxf(m, l[p][i]);
          const float ms = (m > kNegInf) ? m : 0.f;  // all states at -inf: every b~ is 0 (not NaN)
#pragma unroll
          for (int i = 0; i < NS; i++) dst[p][i] = exp_scaled(l[p][i] - ms);
        }
      }
    }
  };

  // ---------------- phase 1: the two chains of each utterance, one thread each ----------------
  const bool chain = warp < 2 && lane < kFbUtts && sT[lane & (kFbUtts - 1)] > 0;
  const int myT = chain ? sT[lane] : 0;
  const int64_t mybase = chain ? sbase[lane] : 0;
  double a[NS * NS];
#pragma unroll
  for (int k = 0; k < NS * NS; k++) a[k] = chain ? sA[lane][k] : 0.0;
  double z[NS];  // forward: z_t; backward: w_t
#pragma unroll
  for (int i = 0; i < NS; i++) z[i] = (warp == 1 && i == NS - 1) ? 1.0 : 0.0;
  int esum = 0;
  if (chain && warp == 1) store_row<NS>(beta_ws + (mybase + myT - 1) * kFbRow, z);  // final state only, T-FS:1484-1490
  stage(0, 0, tid, kFbThreads);
  __syncthreads();
  for (int w0 = 0, wi = 0; w0 < Tmax; w0 += kFbWin, wi++) {
    const double *bcur = bf + (size_t)(wi & 1) * 2 * kFbUtts * US;
    if (warp >= 2) {
      if (w0 + kFbWin < Tmax) stage(w0 + kFbWin, (wi + 1) & 1, tid - 64, kFbThreads - 64);
    } else if (chain && warp == 0) {  // forward: frames w0 .. w0+kFbWin-1
      const double *bw = bcur + (size_t)lane * US;
      const int kend = min(kFbWin, myT - w0);
      float *ap = alpha_ws + (mybase + w0) * kFbRow;
      for (int k = 0; k < kend; k++) {
        double b[NS], raw[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) b[i] = bw[k * NS + i];
        if (w0 + k == 0) {
#pragma unroll
          for (int i = 0; i < NS; i++) raw[i] = (i == 0) ? b[0] : 0.0;  // pi = [1,0,..,0]  T-FS:232-234
        } else {
#pragma unroll
          for (int i = 0; i < NS; i++) {
            double aux;
            if (BANDED) {
              aux = z[i] * a[i * NS + i];
              if (i > 0) aux = fma(z[i - 1], a[(i - 1) * NS + i], aux);
            } else {
              aux = 0.0;
#pragma unroll
              for (int j = 0; j < NS; j++) aux = fma(z[j], a[j * NS + i], aux);
            }
            raw[i] = aux * b[i];
          }
        }
        int e;
        const double r = pow2_scale_max<NS>(raw, e);
#pragma unroll
        for (int i = 0; i < NS; i++) z[i] = raw[i] * r;
        store_row<NS>(ap + (size_t)k * kFbRow, z);
        esum += e;
      }
    } else if (chain && warp == 1) {  // backward: step j = w0 + k turns beta~_{T-1-j} into beta~_{T-2-j} with b~ of frame T-1-j
      const double *bw = bcur + (size_t)(kFbUtts + lane) * US;
      const int kend = min(kFbWin, myT - 1 - w0);
      for (int k = 0; k < kend; k++) {
        double q[NS], raw[NS];
#pragma unroll
        for (int j = 0; j < NS; j++) q[j] = bw[k * NS + j] * z[j];
#pragma unroll
        for (int i = 0; i < NS; i++) {
          double aux;
          if (BANDED) {
            aux = a[i * NS + i] * q[i];
            if (i + 1 < NS) aux = fma(a[i * NS + i + 1], q[i + 1], aux);
          } else {
            aux = 0.0;
#pragma unroll
            for (int j = 0; j < NS; j++) aux = fma(a[i * NS + j], q[j], aux);
          }
          raw[i] = aux;
        }
        int e;
        const double r = pow2_scale_max<NS>(raw, e);
#pragma unroll
        for (int i = 0; i < NS; i++) z[i] = raw[i] * r;
        store_row<NS>(beta_ws + (mybase + myT - 2 - (w0 + k)) * kFbRow, z);
      }
    }
    __syncthreads();  // window consumed, next one staged
  }
  if (chain && warp == 0) {
    double s = z[0];
#pragma unroll
    for (int i = 1; i < NS; i++) s += z[i];
    sphi[lane] = z[NS - 1] / s;                                       // alpha^_{T-1}(N-1)
    slp[lane] = 0.6931471805599453 * (double)esum + log(z[NS - 1]);  // + sum m_t, added in phase 2
  }
  __syncthreads();
  // ---------------- phase 2: one warp per utterance, lanes over frames ----------------
  for (int uu = warp; uu < kFbUtts; uu += kFbThreads / 32) {
    const int u = u0 + uu;
    if (u >= U) continue;
    const int v = u2m[u];
    if (v < 0) {  // masked utterance (its model has converged)
      if (lane == 0 && logp_utt) logp_utt[u] = 0.0;
      continue;
    }
    const int64_t base = off[u];
    const int T = (int)(off[u + 1] - base);
    const double *A = sA[uu];
    const double phi = sphi[uu];
    double acc_num[NS][2], acc_dt[NS], acc_dm[NS], acc_m = 0.0;
#pragma unroll
    for (int i = 0; i < NS; i++) { acc_num[i][0] = acc_num[i][1] = 0.0; acc_dt[i] = 0.0; acc_dm[i] = 0.0; }
    // the rows of frame t + 32 are fetched while frame t is being worked on (the loop is latency bound otherwise)
    struct Raw { float4 a0, a1, b0, b1, c0, c1; float l0[NS], l1[NS]; };
    auto fetch = [&](int t, Raw &r) {
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      r.a0 = r.a1 = r.b0 = r.b1 = r.c0 = r.c1 = z4;
#pragma unroll
      for (int i = 0; i < NS; i++) r.l0[i] = r.l1[i] = 0.f;
      if (t < T) {
        const float4 *pa = reinterpret_cast<const float4 *>(alpha_ws + (base + t) * kFbRow);
        const float4 *pb = reinterpret_cast<const float4 *>(beta_ws + (base + t) * kFbRow);
        r.a0 = pa[0]; r.b0 = pb[0];
        if (NS > 4) { r.a1 = pa[1]; r.b1 = pb[1]; }
        load_lb<NS>(logb + (base + t) * NS, r.l0);
        if (t < T - 1) {
          r.c0 = pb[2];  // beta row of frame t + 1 (rows are kFbRow = 8 floats)
          if (NS > 4) r.c1 = pb[3];
          load_lb<NS>(logb + (base + t + 1) * NS, r.l1);
        }
      }
    };
    auto unpack = [](const float4 &x0, const float4 &x1, double (&o)[NS]) {
      const float f[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < NS; i++) o[i] = (double)f[i];
    };
    Raw cur, nxt;
    fetch(lane, cur);
    for (int t = lane; t < T; t += 32) {
      fetch(t + 32, nxt);
      double al[NS], be[NS], g[NS], G = 0.0;
      unpack(cur.a0, cur.a1, al);
      unpack(cur.b0, cur.b1, be);
      {
        float m = cur.l0[0];
#pragma unroll
        for (int i = 1; i < NS; i++) m = fmaxf(m, cur.l0[i]);
        acc_m += (double)m;  // sum_t m_t  (calc_probability); -inf when a frame has no density at all
      }
#pragma unroll
      for (int i = 0; i < NS; i++) {
        g[i] = al[i] * be[i];
        G += g[i];
      }
      const double sc = (G > 0.0) ? phi / G : 0.0;  // unreachable final state: no occupancy, as the reference
#pragma unroll
      for (int i = 0; i < NS; i++) {
        g[i] *= sc;  // alpha^ beta^ / c   T-FS:1617,1658,1709
        gamma[(base + t) * NS + i] = (float)g[i];
        acc_dm[i] += g[i];
      }
      if (t < T - 1) {
        float m = cur.l1[0];
#pragma unroll
        for (int i = 1; i < NS; i++) m = fmaxf(m, cur.l1[i]);
        double q[NS], b1[NS];
        unpack(cur.c0, cur.c1, b1);
#pragma unroll
        for (int j = 0; j < NS; j++) q[j] = exp_scaled(cur.l1[j] - ((m > kNegInf) ? m : 0.f)) * b1[j];
        double Z = 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) {
          double rb = 0.0;
#pragma unroll
          for (int j = 0; j < NS; j++) rb = fma(A[i * NS + j], q[j], rb);
          Z = fma(al[i], rb, Z);
        }
        const double zs = (Z > 0.0) ? phi / Z : 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) {
          acc_dt[i] += g[i];
          acc_num[i][0] += al[i] * A[i * NS + i] * q[i] * zs;                              // band j = i   T-FS:1611
          if (i + 1 < NS) acc_num[i][1] += al[i] * A[i * NS + i + 1] * q[i + 1] * zs;      // j = i + 1
        }
      }
      cur = nxt;
    }
    double *st = stats + (int64_t)v * stats_stride;
#pragma unroll
    for (int i = 0; i < NS; i++) {
      const double n0 = warp_sum(acc_num[i][0]), n1 = warp_sum(acc_num[i][1]);
      const double dt = warp_sum(acc_dt[i]), dm = warp_sum(acc_dm[i]);
      if (lane == 0) {
        atomicAdd(st + i * NS + i, n0);
        if (i + 1 < NS) atomicAdd(st + i * NS + i + 1, n1);
        atomicAdd(st + NS * NS + i, dt);
        atomicAdd(st + NS * NS + NS + i, dm);
      }
    }
    const double msum = warp_sum(acc_m);
    if (lane == 0) {
      const double lp = slp[uu] + msum;  // calc_probability T-FS:1546-1549
      atomicAdd(st + off_sumlogp, lp);
      atomicAdd(st + off_sumlogp + 1, 1.0);
      if (logp_utt) logp_utt[u] = lp;
    }
  }
}

// ================================================================================================
// k_fb_seg: the same recursions, parallel in time.
//
// The chain of an utterance is a product of T linear maps z -> (z A) o b~_t; the map of a segment of frames can
// be built without knowing where the segment starts by pushing the N unit vectors through it.  So each
// utterance is cut into kSegs segments and the recursion runs in two passes of T / kSegs steps instead of one of T:
//   pass 1   segment 0 runs the true chain (and stores its rows); for every other segment, N threads push the
//            unit vectors e_0 .. e_{N-1} through it (own power-of-two scaling, cumulative exponent kept) and keep
//            only the end vectors;
//   boundary one thread per (utterance, direction) walks the segments: start vector of segment s+1 = sum_i a_i
//            y_i 2^{E_i}, renormalised; this also yields phi, and log P from pass-1 quantities only;
//   pass 2   every remaining segment runs its true chain from its start vector and stores its rows.
// Backward is the mirror image (segments from the end, w_{T-1} = e_{N-1}).  All b~ of the CTA's utterances are staged
// once in shared memory (4 utterances x T x N doubles) and shared by both directions.  The third phase (gamma, xi,
// den sums) is k_fb's, two warps per utterance.  Used when the utterances fit (fb_seg_fits); k_fb otherwise.
// ================================================================================================
constexpr int kSegs = 8;
constexpr int kSegUtts = 4;
constexpr int kSegThreads = 288;  // 9 warps: 2 x 36 chain lanes per utterance in pass 1

__host__ __device__ inline size_t fb_seg_smem_bytes(int NS, int Tmax) { return sizeof(double) * (size_t)kSegUtts * ((size_t)Tmax * NS + 2); }
__host__ __device__ inline bool fb_seg_fits(int NS, int Tmax) { return fb_seg_smem_bytes(NS, Tmax) <= 160 * 1024; }

__device__ __forceinline__ int seg_begin(int s, int T) { return min(T, max(s, (int)(((long long)s * T) / kSegs))); }
// v * 2^k for |k| possibly large (k < -1000 -> 0; the operands here never need k > 0 by much)
__device__ __forceinline__ double scale_pow2(double v, int k) {
  if (k < -1000) return 0.0;
  if (k > 1000) k = 1000;
  return v * __hiloint2double((1023 + k) << 20, 0);
}

template <int NS, bool BANDED>
__global__ void __launch_bounds__(kSegThreads, 2)
k_fb_seg(const float *__restrict__ logb, const int64_t *__restrict__ off, const int32_t *__restrict__ u2m,
         const double *__restrict__ Aall, int U, int Tcap, float *__restrict__ alpha_ws, float *__restrict__ beta_ws,
         float *__restrict__ gamma, double *__restrict__ stats, int64_t stats_stride, int64_t off_sumlogp,
         double *__restrict__ logp_utt) {
  extern __shared__ __align__(16) uint8_t fb_smem[];
  const int US = Tcap * NS + 2;                       // per-utterance stride of the staged b~
  double *bt = reinterpret_cast<double *>(fb_smem);  // [kSegUtts][US]   b~_i(t), t < T
  __shared__ double sA[kSegUtts][NS * NS];
  __shared__ double yend[2][kSegUtts][kSegs][NS][NS];  // pass 1: end vector of (direction, utterance, segment, unit vector)
  __shared__ int yexp[2][kSegUtts][kSegs][NS];         //         and its cumulative exponent
  __shared__ double zend[2][kSegUtts][NS];             // end vector of the first segment's true chain
  __shared__ int zexp[2][kSegUtts];
  __shared__ double sstart[2][kSegUtts][kSegs][NS];    // start vector of every segment (boundary pass)
  __shared__ double sphi[kSegUtts], slp[kSegUtts], smsum[kSegUtts];
  __shared__ int64_t sbase[kSegUtts];
  __shared__ int sT[kSegUtts];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = blockIdx.x * kSegUtts;
  for (int idx = tid; idx < kSegUtts * NS * NS; idx += kSegThreads) {
    const int uu = idx / (NS * NS), k = idx - uu * NS * NS;
    const int u = u0 + uu;
    const int v = (u < U) ? u2m[u] : -1;
    sA[uu][k] = (v >= 0) ? Aall[(int64_t)v * NS * NS + k] : 0.0;
  }
  if (tid < kSegUtts) {
    const int u = u0 + tid;
    const bool live = u < U && u2m[u] >= 0;
    sbase[tid] = live ? off[u] : 0;
    sT[tid] = live ? (int)(off[u + 1] - off[u]) : 0;
    smsum[tid] = 0.0;
  }
  __syncthreads();
  // ---------------- staging: b~ of every frame of the CTA's utterances ----------------
  for (int uu = 0; uu < kSegUtts; uu++) {
    const int T = sT[uu];
    const float *src = logb + sbase[uu] * NS;
    double *dst = bt + (size_t)uu * US;
    for (int t = tid; t < T; t += 2 * kSegThreads) {  // two frames per trip: their loads are in flight together
      const int t2 = t + kSegThreads;
      float l[NS], l2[NS];
      load_lb<NS>(src + (size_t)t * NS, l);
      if (t2 < T) load_lb<NS>(src + (size_t)t2 * NS, l2);
      {
        float m = l[0];
#pragma unroll
        for (int i = 1; i < NS; i++) m = fmaxf(m, l[i]);
        const float ms = (m > kNegInf) ? m : 0.f;  // all states at -inf: every b~ is 0 (not NaN)
#pragma unroll
        for (int i = 0; i < NS; i++) dst[(size_t)t * NS + i] = exp_scaled(l[i] - ms);
      }
      if (t2 < T) {
        float m = l2[0];
#pragma unroll
        for (int i = 1; i < NS; i++) m = fmaxf(m, l2[i]);
        const float ms = (m > kNegInf) ? m : 0.f;
#pragma unroll
        for (int i = 0; i < NS; i++) dst[(size_t)t2 * NS + i] = exp_scaled(l2[i] - ms);
      }
    }
  }
  __syncthreads();

  // one forward step z <- ((z A) o b) 2^-e (first = the step of frame 0: pi o b); one backward step w <- (A (b o w)) 2^-e
  auto fwd_step = [&](double (&z)[NS], const double (&a)[NS * NS], const double *b, bool first, int &e) {
    double raw[NS];
    if (first) {
#pragma unroll
      for (int i = 0; i < NS; i++) raw[i] = (i == 0) ? b[0] : 0.0;  // pi = [1,0,..,0]  T-FS:232-234
    } else {
#pragma unroll
      for (int i = 0; i < NS; i++) {
        double aux;
        if (BANDED) {
          aux = z[i] * a[i * NS + i];
          if (i > 0) aux = fma(z[i - 1], a[(i - 1) * NS + i], aux);
        } else {
          aux = 0.0;
#pragma unroll
          for (int j = 0; j < NS; j++) aux = fma(z[j], a[j * NS + i], aux);
        }
        raw[i] = aux * b[i];
      }
    }
    const double r = pow2_scale_max<NS>(raw, e);
#pragma unroll
    for (int i = 0; i < NS; i++) z[i] = raw[i] * r;
  };
  auto bwd_step = [&](double (&w)[NS], const double (&a)[NS * NS], const double *b, int &e) {
    double q[NS], raw[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) q[j] = b[j] * w[j];
#pragma unroll
    for (int i = 0; i < NS; i++) {
      double aux;
      if (BANDED) {
        aux = a[i * NS + i] * q[i];
        if (i + 1 < NS) aux = fma(a[i * NS + i + 1], q[i + 1], aux);
      } else {
        aux = 0.0;
#pragma unroll
        for (int j = 0; j < NS; j++) aux = fma(a[i * NS + j], q[j], aux);
      }
      raw[i] = aux;
    }
    const double r = pow2_scale_max<NS>(raw, e);
#pragma unroll
    for (int i = 0; i < NS; i++) w[i] = raw[i] * r;
  };

  // ---------------- pass 1: true chain of the first segment, unit vectors through the others ----------------
  constexpr int CPU_ = 1 + (kSegs - 1) * NS;  // chains per (utterance, direction)
  for (int cid = tid; cid < 2 * kSegUtts * CPU_; cid += kSegThreads) {  // one round for N <= 5
    const int dir = cid / (kSegUtts * CPU_), rem = cid - dir * (kSegUtts * CPU_);
    const int uu = rem / CPU_, k = rem - uu * CPU_;
    const int T = sT[uu];
    if (T > 0) {
      double a[NS * NS];
#pragma unroll
      for (int q = 0; q < NS * NS; q++) a[q] = sA[uu][q];
      const double *bw = bt + (size_t)uu * US;
      const int64_t base = sbase[uu];
      double z[NS];
      int esum = 0, e;
      if (dir == 0) {
        if (k == 0) {  // frames [0, fs_1): the true chain, rows stored
#pragma unroll
          for (int i = 0; i < NS; i++) z[i] = 0.0;
          const int t1 = seg_begin(1, T);
          for (int t = 0; t < t1; t++) {
            fwd_step(z, a, bw + (size_t)t * NS, t == 0, e);
            esum += e;
            store_row<NS>(alpha_ws + (base + t) * kFbRow, z);
          }
#pragma unroll
          for (int i = 0; i < NS; i++) zend[0][uu][i] = z[i];
          zexp[0][uu] = esum;
        } else {
          const int sg = 1 + (k - 1) / NS, bi = (k - 1) % NS;
#pragma unroll
          for (int i = 0; i < NS; i++) z[i] = (i == bi) ? 1.0 : 0.0;
          const int t0 = seg_begin(sg, T), t1 = seg_begin(sg + 1, T);
          for (int t = t0; t < t1; t++) {
            fwd_step(z, a, bw + (size_t)t * NS, false, e);
            esum += e;
          }
#pragma unroll
          for (int i = 0; i < NS; i++) yend[0][uu][sg][bi][i] = z[i];
          yexp[0][uu][sg][bi] = esum;
        }
      } else {
        // backward: w_t for t = T-2 .. 0; step t uses b~ of frame t + 1.  Segment s owns t in [fs_s, min(fs_{s+1}, T-1)).
        if (k == 0) {  // the last segment: true chain from w_{T-1} = e_{N-1} (final state only, T-FS:1484-1490)
#pragma unroll
          for (int i = 0; i < NS; i++) z[i] = (i == NS - 1) ? 1.0 : 0.0;
          store_row<NS>(beta_ws + (base + T - 1) * kFbRow, z);
          const int t0 = seg_begin(kSegs - 1, T);
          for (int t = T - 2; t >= t0; t--) {
            bwd_step(z, a, bw + (size_t)(t + 1) * NS, e);
            store_row<NS>(beta_ws + (base + t) * kFbRow, z);
          }
#pragma unroll
          for (int i = 0; i < NS; i++) zend[1][uu][i] = z[i];
        } else {
          const int sg = (k - 1) / NS, bi = (k - 1) % NS;  // segments 0 .. kSegs-2
#pragma unroll
          for (int i = 0; i < NS; i++) z[i] = (i == bi) ? 1.0 : 0.0;
          const int t0 = seg_begin(sg, T), t1 = min(seg_begin(sg + 1, T), T - 1);
          for (int t = t1 - 1; t >= t0; t--) {
            bwd_step(z, a, bw + (size_t)(t + 1) * NS, e);
            esum += e;
          }
#pragma unroll
          for (int i = 0; i < NS; i++) yend[1][uu][sg][bi][i] = z[i];
          yexp[1][uu][sg][bi] = esum;
        }
      }
    }
  }
  __syncthreads();
  // ---------------- boundaries: start vector of every segment; phi and log P (forward) ----------------
  // one warp per (direction, utterance), lane j holds component j of the running vector
  if (warp < 2 * kSegUtts) {
    const int dir = warp / kSegUtts, uu = warp - dir * kSegUtts;
    if (sT[uu] > 0) {
      const int j = lane < NS ? lane : 0;
      double av = (lane < NS) ? zend[dir][uu][j] : 0.0;
      int etot = (dir == 0) ? zexp[0][uu] : 0;
      for (int step = 1; step < kSegs; step++) {
        const int sg = (dir == 0) ? step : kSegs - 1 - step;  // forward: segments 1 .. ; backward: kSegs-2 .. 0
        if (lane < NS) sstart[dir][uu][sg][j] = av;
        // lane i: does unit vector i carry weight and survive the segment?  reference exponent = the largest such
        double ym = 0.0;
#pragma unroll
        for (int q = 0; q < NS; q++) ym = fmax(ym, yend[dir][uu][sg][j][q]);
        const bool alive = lane < NS && av != 0.0 && ym > 0.0;
        int ex = alive ? yexp[dir][uu][sg][j] : -(1 << 30);
        int eref = ex;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) eref = max(eref, __shfl_xor_sync(0xffffffffu, eref, o));  // NS <= 8: lanes 0..7
        const double ci = alive ? scale_pow2(av, ex - eref) : 0.0;
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) {
          const double c = __shfl_sync(0xffffffffu, ci, i);
          v = fma(c, yend[dir][uu][sg][i][j], v);
        }
        if (lane >= NS) v = 0.0;
        if (eref == -(1 << 30)) eref = 0;
        int be = __double2hiint(v) >> 20;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) be = max(be, __shfl_xor_sync(0xffffffffu, be, o));
        const bool ok = be > 0 && be < 0x7ff;
        av = v * __hiloint2double(ok ? (2046 - be) << 20 : 0x3ff00000, 0);
        etot += eref + (ok ? be - 1023 : 0);
      }
      if (dir == 0) {
        double sm = av;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
        const double last = __shfl_sync(0xffffffffu, av, NS - 1);
        if (lane == 0) {
          sphi[uu] = last / sm;                                           // alpha^_{T-1}(N-1)
          slp[uu] = 0.6931471805599453 * (double)etot + log(last);       // + sum m_t, added in phase 3
        }
      }
    }
  }
  __syncthreads();
  // ---------------- pass 2: true chains of the remaining segments, rows stored ----------------
  if (tid < 2 * kSegUtts * (kSegs - 1)) {
    const int dir = tid / (kSegUtts * (kSegs - 1)), rem = tid - dir * (kSegUtts * (kSegs - 1));
    const int uu = rem / (kSegs - 1), sx = rem - uu * (kSegs - 1);
    const int T = sT[uu];
    if (T > 0) {
      double a[NS * NS];
#pragma unroll
      for (int q = 0; q < NS * NS; q++) a[q] = sA[uu][q];
      const double *bw = bt + (size_t)uu * US;
      const int64_t base = sbase[uu];
      double z[NS];
      int e;
      if (dir == 0) {
        const int sg = 1 + sx;
#pragma unroll
        for (int i = 0; i < NS; i++) z[i] = sstart[0][uu][sg][i];
        const int t0 = seg_begin(sg, T), t1 = seg_begin(sg + 1, T);
        for (int t = t0; t < t1; t++) {
          fwd_step(z, a, bw + (size_t)t * NS, false, e);
          store_row<NS>(alpha_ws + (base + t) * kFbRow, z);
        }
      } else {
        const int sg = sx;  // segments 0 .. kSegs-2
#pragma unroll
        for (int i = 0; i < NS; i++) z[i] = sstart[1][uu][sg][i];
        const int t0 = seg_begin(sg, T), t1 = min(seg_begin(sg + 1, T), T - 1);
        for (int t = t1 - 1; t >= t0; t--) {
          bwd_step(z, a, bw + (size_t)(t + 1) * NS, e);
          store_row<NS>(beta_ws + (base + t) * kFbRow, z);
        }
      }
    }
  }
  __syncthreads();
  // ---------------- phase 3: gamma, transition and den sums; two warps per utterance, lanes over frames ----------------
  if (warp < 2 * kSegUtts) {
    const int uu = warp >> 1, half = warp & 1;
    const int u = u0 + uu;
    const int v = (u < U) ? u2m[u] : -1;
    if (u < U && v >= 0) {
      const int64_t base = sbase[uu];
      const int T = sT[uu];
      const double *A = sA[uu];
      const double phi = sphi[uu];
      double acc_num[NS][2], acc_dt[NS], acc_dm[NS], acc_m = 0.0;
#pragma unroll
      for (int i = 0; i < NS; i++) { acc_num[i][0] = acc_num[i][1] = 0.0; acc_dt[i] = 0.0; acc_dm[i] = 0.0; }
      for (int t = half * 32 + lane; t < T; t += 64) {
        double al[NS], be[NS], g[NS], G = 0.0;
        load_row<NS>(alpha_ws + (base + t) * kFbRow, al);
        load_row<NS>(beta_ws + (base + t) * kFbRow, be);
        {
          const double *b0 = bt + (size_t)uu * US + (size_t)t * NS;  // (b~ is at hand; m_t still comes from logb)
          (void)b0;
          float l0[NS];
          load_lb<NS>(logb + (base + t) * NS, l0);
          float m = l0[0];
#pragma unroll
          for (int i = 1; i < NS; i++) m = fmaxf(m, l0[i]);
          acc_m += (double)m;
        }
#pragma unroll
        for (int i = 0; i < NS; i++) {
          g[i] = al[i] * be[i];
          G += g[i];
        }
        const double sc = (G > 0.0) ? phi / G : 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) {
          g[i] *= sc;
          gamma[(base + t) * NS + i] = (float)g[i];
          acc_dm[i] += g[i];
        }
        if (t < T - 1) {
          double q[NS], b1[NS];
          load_row<NS>(beta_ws + (base + t + 1) * kFbRow, b1);
          const double *bn = bt + (size_t)uu * US + (size_t)(t + 1) * NS;  // b~ of frame t + 1, staged above
#pragma unroll
          for (int j = 0; j < NS; j++) q[j] = bn[j] * b1[j];
          double Z = 0.0;
#pragma unroll
          for (int i = 0; i < NS; i++) {
            double rb = 0.0;
#pragma unroll
            for (int j = 0; j < NS; j++) rb = fma(A[i * NS + j], q[j], rb);
            Z = fma(al[i], rb, Z);
          }
          const double zs = (Z > 0.0) ? phi / Z : 0.0;
#pragma unroll
          for (int i = 0; i < NS; i++) {
            acc_dt[i] += g[i];
            acc_num[i][0] += al[i] * A[i * NS + i] * q[i] * zs;
            if (i + 1 < NS) acc_num[i][1] += al[i] * A[i * NS + i + 1] * q[i + 1] * zs;
          }
        }
      }
      double *st = stats + (int64_t)v * stats_stride;
#pragma unroll
      for (int i = 0; i < NS; i++) {
        const double n0 = warp_sum(acc_num[i][0]), n1 = warp_sum(acc_num[i][1]);
        const double dt = warp_sum(acc_dt[i]), dm = warp_sum(acc_dm[i]);
        if (lane == 0) {
          atomicAdd(st + i * NS + i, n0);
          if (i + 1 < NS) atomicAdd(st + i * NS + i + 1, n1);
          atomicAdd(st + NS * NS + i, dt);
          atomicAdd(st + NS * NS + NS + i, dm);
        }
      }
      const double msum = warp_sum(acc_m);
      if (lane == 0) atomicAdd(&smsum[uu], msum);
    }
  }
  __syncthreads();
  if (tid < kSegUtts) {
    const int u = u0 + tid;
    if (u < U) {
      const int v = u2m[u];
      if (v < 0) {
        if (logp_utt) logp_utt[u] = 0.0;
      } else {
        const double lp = slp[tid] + smsum[tid];  // calc_probability T-FS:1546-1549
        double *st = stats + (int64_t)v * stats_stride;
        atomicAdd(st + off_sumlogp, lp);
        atomicAdd(st + off_sumlogp + 1, 1.0);
        if (logp_utt) logp_utt[u] = lp;
      }
    }
  }
}

// ================================================================================================
// k_fb_wide + k_fb_gamma: the recursions for MANY utterances (a C3 shard: 12,500 per GPU).
// With thousands of utterances there are enough independent chains to fill the machine without cutting them up:
// one thread per (utterance, direction), every lane of every warp busy, log-emissions read straight from global
// memory four frames ahead (as the decode scorers do), b~ formed on the fly.  The third phase (gamma, xi, den
// sums) is a kernel of its own, one warp per utterance.  At 1,000 utterances this is slower than k_fb_seg (2,000
// threads cannot hide a 300-step chain); hmmcu_estep picks by the number of utterances.
// ================================================================================================
constexpr int kWideThreads = 128;
constexpr int kWidePF = 4;

// Log-emissions are fetched in chunks of four frames aligned to the absolute frame index: 4 x NS floats = NS
// 16-byte loads (a lane per utterance makes every load instruction touch 32 sectors, so the fewer the better).
// The chunk of the next four frames is in flight while the current one is consumed; a chunk may reach a few
// frames beyond the utterance or the buffer (the allocation has slack), those frames are skipped.
template <int NS, bool BANDED>
__global__ void __launch_bounds__(kWideThreads)
k_fb_wide(const float *__restrict__ logb, const int64_t *__restrict__ off, const int32_t *__restrict__ u2m,
          const double *__restrict__ Aall, int U, float *__restrict__ alpha_ws, float *__restrict__ beta_ws,
          double *__restrict__ phi_utt, double *__restrict__ lp_utt) {
  const int cid = blockIdx.x * kWideThreads + threadIdx.x;
  if (cid >= 2 * U) return;
  const int dir = cid / U, u = cid - dir * U;  // forward chains first: the lanes of a warp share a direction
  const int v = u2m[u];
  if (v < 0) return;
  const int64_t base = off[u];
  const int T = (int)(off[u + 1] - base);
  double a[NS * NS];
#pragma unroll
  for (int k = 0; k < NS * NS; k++) a[k] = Aall[(int64_t)v * NS * NS + k];
  double z[NS];
  const float4 *lb4 = reinterpret_cast<const float4 *>(logb);
  auto load_chunk = [&](int64_t c, float (&buf)[4 * NS]) {
    const float4 *p = lb4 + c * NS;
#pragma unroll
    for (int q = 0; q < NS; q++) {
      const float4 x = __ldg(p + q);
      buf[4 * q] = x.x; buf[4 * q + 1] = x.y; buf[4 * q + 2] = x.z; buf[4 * q + 3] = x.w;
    }
  };
  float nxt[4 * NS], cur[4 * NS];
  if (dir == 0) {
#pragma unroll
    for (int i = 0; i < NS; i++) z[i] = 0.0;
    int esum = 0;
    const int64_t c0 = base >> 2, c1 = (base + T - 1) >> 2;
    load_chunk(c0, nxt);
    for (int64_t c = c0; c <= c1; c++) {
#pragma unroll
      for (int q = 0; q < 4 * NS; q++) cur[q] = nxt[q];
      if (c < c1) load_chunk(c + 1, nxt);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int64_t f = 4 * c + k;
        if (f >= base && f < base + T) {
          const float *l = cur + k * NS;
          float m = l[0];
#pragma unroll
          for (int i = 1; i < NS; i++) m = fmaxf(m, l[i]);
          const float ms = (m > kNegInf) ? m : 0.f;
          double raw[NS];
          if (f == base) {
#pragma unroll
            for (int i = 0; i < NS; i++) raw[i] = (i == 0) ? exp_scaled(l[0] - ms) : 0.0;  // pi = [1,0,..,0]
          } else {
#pragma unroll
            for (int i = 0; i < NS; i++) {
              double aux;
              if (BANDED) {
                aux = z[i] * a[i * NS + i];
                if (i > 0) aux = fma(z[i - 1], a[(i - 1) * NS + i], aux);
              } else {
                aux = 0.0;
#pragma unroll
                for (int j = 0; j < NS; j++) aux = fma(z[j], a[j * NS + i], aux);
              }
              raw[i] = aux * exp_scaled(l[i] - ms);
            }
          }
          int e;
          const double r = pow2_scale_max<NS>(raw, e);
#pragma unroll
          for (int i = 0; i < NS; i++) z[i] = raw[i] * r;
          esum += e;
          store_row<NS>(alpha_ws + f * kFbRow, z);
        }
      }
    }
    double sm = z[0];
#pragma unroll
    for (int i = 1; i < NS; i++) sm += z[i];
    phi_utt[u] = z[NS - 1] / sm;                                            // alpha^_{T-1}(N-1)
    lp_utt[u] = 0.6931471805599453 * (double)esum + log(z[NS - 1]);        // + sum m_t, added by k_fb_gamma
  } else {
    // backward: w_{T-1} = e_{N-1}; frame f = T-1 .. 1 (relative) turns w_f into w_{f-1} with the emissions of frame f
#pragma unroll
    for (int i = 0; i < NS; i++) z[i] = (i == NS - 1) ? 1.0 : 0.0;
    store_row<NS>(beta_ws + (base + T - 1) * kFbRow, z);
    if (T > 1) {
      const int64_t chi = (base + T - 1) >> 2, clo = (base + 1) >> 2;
      load_chunk(chi, nxt);
      for (int64_t c = chi; c >= clo; c--) {
#pragma unroll
        for (int q = 0; q < 4 * NS; q++) cur[q] = nxt[q];
        if (c > clo) load_chunk(c - 1, nxt);
#pragma unroll
        for (int k = 3; k >= 0; k--) {
          const int64_t f = 4 * c + k;
          if (f >= base + 1 && f <= base + T - 1) {
            const float *l = cur + k * NS;
            float m = l[0];
#pragma unroll
            for (int i = 1; i < NS; i++) m = fmaxf(m, l[i]);
            const float ms = (m > kNegInf) ? m : 0.f;
            double q[NS], raw[NS];
#pragma unroll
            for (int j = 0; j < NS; j++) q[j] = exp_scaled(l[j] - ms) * z[j];
#pragma unroll
            for (int i = 0; i < NS; i++) {
              double aux;
              if (BANDED) {
                aux = a[i * NS + i] * q[i];
                if (i + 1 < NS) aux = fma(a[i * NS + i + 1], q[i + 1], aux);
              } else {
                aux = 0.0;
#pragma unroll
                for (int j = 0; j < NS; j++) aux = fma(a[i * NS + j], q[j], aux);
              }
              raw[i] = aux;
            }
            int e;
            const double r = pow2_scale_max<NS>(raw, e);
#pragma unroll
            for (int i = 0; i < NS; i++) z[i] = raw[i] * r;
            store_row<NS>(beta_ws + (f - 1) * kFbRow, z);
          }
        }
      }
    }
  }
}

// third phase for k_fb_wide: one warp per utterance, lanes over frames (the same arithmetic as k_fb's phase 2)
constexpr int kGammaWarps = 8;
template <int NS>
__global__ void __launch_bounds__(kGammaWarps * 32, 2)
k_fb_gamma(const float *__restrict__ logb, const int64_t *__restrict__ off, const int32_t *__restrict__ u2m,
           const double *__restrict__ Aall, int U, const float *__restrict__ alpha_ws, const float *__restrict__ beta_ws,
           const double *__restrict__ phi_utt, const double *__restrict__ lp_utt, float *__restrict__ gamma,
           double *__restrict__ stats, int64_t stats_stride, int64_t off_sumlogp, double *__restrict__ logp_utt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * kGammaWarps + warp;
  if (u >= U) return;
  const int v = u2m[u];
  if (v < 0) {  // masked utterance (its model has converged)
    if (lane == 0 && logp_utt) logp_utt[u] = 0.0;
    return;
  }
  const int64_t base = off[u];
  const int T = (int)(off[u + 1] - base);
  double A[NS * NS];
#pragma unroll
  for (int k = 0; k < NS * NS; k++) A[k] = Aall[(int64_t)v * NS * NS + k];
  const double phi = phi_utt[u];
  double acc_num[NS][2], acc_dt[NS], acc_dm[NS], acc_m = 0.0;
#pragma unroll
  for (int i = 0; i < NS; i++) { acc_num[i][0] = acc_num[i][1] = 0.0; acc_dt[i] = 0.0; acc_dm[i] = 0.0; }
  for (int t = lane; t < T; t += 32) {
    double al[NS], be[NS], g[NS], G = 0.0;
    load_row<NS>(alpha_ws + (base + t) * kFbRow, al);
    load_row<NS>(beta_ws + (base + t) * kFbRow, be);
    {
      float l0[NS];
      load_lb<NS>(logb + (base + t) * NS, l0);
      float m = l0[0];
#pragma unroll
      for (int i = 1; i < NS; i++) m = fmaxf(m, l0[i]);
      acc_m += (double)m;
    }
#pragma unroll
    for (int i = 0; i < NS; i++) {
      g[i] = al[i] * be[i];
      G += g[i];
    }
    const double sc = (G > 0.0) ? phi / G : 0.0;  // unreachable final state: no occupancy, as the reference
#pragma unroll
    for (int i = 0; i < NS; i++) {
      g[i] *= sc;  // alpha^ beta^ / c   T-FS:1617,1658,1709
      gamma[(base + t) * NS + i] = (float)g[i];
      acc_dm[i] += g[i];
    }
    if (t < T - 1) {
      float l1[NS];
      load_lb<NS>(logb + (base + t + 1) * NS, l1);
      float m = l1[0];
#pragma unroll
      for (int i = 1; i < NS; i++) m = fmaxf(m, l1[i]);
      double q[NS], b1[NS];
      load_row<NS>(beta_ws + (base + t + 1) * kFbRow, b1);
#pragma unroll
      for (int j = 0; j < NS; j++) q[j] = exp_scaled(l1[j] - ((m > kNegInf) ? m : 0.f)) * b1[j];
      double Z = 0.0;
#pragma unroll
      for (int i = 0; i < NS; i++) {
        double rb = 0.0;
#pragma unroll
        for (int j = 0; j < NS; j++) rb = fma(A[i * NS + j], q[j], rb);
        Z = fma(al[i], rb, Z);
      }
      const double zs = (Z > 0.0) ? phi / Z : 0.0;
#pragma unroll
      for (int i = 0; i < NS; i++) {
        acc_dt[i] += g[i];
        acc_num[i][0] += al[i] * A[i * NS + i] * q[i] * zs;                              // band j = i   T-FS:1611
        if (i + 1 < NS) acc_num[i][1] += al[i] * A[i * NS + i + 1] * q[i + 1] * zs;      // j = i + 1
      }
    }
  }
  double *st = stats + (int64_t)v * stats_stride;
#pragma unroll
  for (int i = 0; i < NS; i++) {
    const double n0 = warp_sum(acc_num[i][0]), n1 = warp_sum(acc_num[i][1]);
    const double dt = warp_sum(acc_dt[i]), dm = warp_sum(acc_dm[i]);
    if (lane == 0) {
      atomicAdd(st + i * NS + i, n0);
      if (i + 1 < NS) atomicAdd(st + i * NS + i + 1, n1);
      atomicAdd(st + NS * NS + i, dt);
      atomicAdd(st + NS * NS + NS + i, dm);
    }
  }
  const double msum = warp_sum(acc_m);
  if (lane == 0) {
    const double lp = lp_utt[u] + msum;  // calc_probability T-FS:1546-1549
    atomicAdd(st + off_sumlogp, lp);
    atomicAdd(st + off_sumlogp + 1, 1.0);
    if (logp_utt) logp_utt[u] = lp;
  }
}

}  // namespace hmmk
