// init_kernels.cuh -- creating_initial_model (T-FS:732-1317) on the device, for V words at
// once: uniform segmentation, one centroid per state, LBG splitting (x1.005 / x0.995), three k-means passes per
// split level with the empty-cell rule, then per-cluster variances and weights.
//
// Results are BIT-IDENTICAL to the host restatement hmmh_init_model() (which is pinned against the compiled
// reference): every floating-point operation is the same IEEE double operation in the same order.
//   * distances are formed with explicit __dmul_rn / __dadd_rn (no FMA contraction), dimensions in order;
//   * every sum the reference accumulates frame by frame (centroid sums, distortions, squared deviations) is
//     accumulated sequentially in the reference's frame order by ONE thread per (word, state, cell, dimension):
//     the frames of a (word, state) pair are listed once on the host in that order (utterance by utterance,
//     segment k of each), a warp per (word, state, cell) walks the list with its lanes over the dimensions.
//     A list has ~T*U/N entries, so the sequential walk is some thousands of steps -- microseconds, not the
//     seconds the single-threaded host loop takes over all frames;
//   * the classification itself (nearest centroid, strict '<', first minimum wins) is one thread per frame.
// One difference, documented: the reference's `which` is left untouched when no distance beats 1e20 (it then
// keeps the previous frame's cell, T-FS:1179-1215); here such a frame goes to cell 0.  That needs distances
// >= 1e20 or NaNs and does not occur for finite features.
#pragma once
#include "kernels.cuh"

namespace hmmk {

// entry e of the lists: frame id lst[e]; the entries of (word v, state k) are [lst_off[v*N+k], lst_off[v*N+k+1])
// cent / sum: [V][N][M][D]; cnt / dist: [V][N][M]

// nearest of `have` centroids of each entry's (v, k)   (classifying, T-FS:1179-1215)
// DT = D at compile time keeps the frame in registers (it is compared with up to M centroids; re-reading it through
// the cache, a line per lane, made this kernel L1-bound); DT = 0 is the general form.
template <int DT>
__global__ void k_init_classify(const double *__restrict__ x, const int32_t *__restrict__ lst, const int32_t *__restrict__ ent_vk,
                                int64_t E, const double *__restrict__ cent, int M, int D, int have, uint8_t *__restrict__ idx,
                                double *__restrict__ dd) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const double *xr = x + (int64_t)lst[e] * D;
  const double *c = cent + (int64_t)ent_vk[e] * M * D;
  double xv[DT > 0 ? DT : 1];
  if (DT > 0) {
#pragma unroll
    for (int d = 0; d < DT; d++) xv[d] = xr[d];
  }
  double best = 1.0e20;
  int which = 0;
  for (int i = 0; i < have; i++) {
    double dist = 0.0;
    if (DT > 0) {
#pragma unroll
      for (int d = 0; d < DT; d++) {
        const double a = __dsub_rn(c[(int64_t)i * DT + d], xv[d]);
        dist = __dadd_rn(dist, __dmul_rn(a, a));
      }
    } else {
      for (int d = 0; d < D; d++) {
        const double a = __dsub_rn(c[(int64_t)i * D + d], xr[d]);
        dist = __dadd_rn(dist, __dmul_rn(a, a));
      }
    }
    if (dist < best) { best = dist; which = i; }
  }
  idx[e] = (uint8_t)which;
  dd[e] = best;
}

// One warp per (v, k, cell j): sequential sums over the list of (v, k), lanes over the dimensions (two per lane).
//   mode 0: sum[d] += x[d]                    (+ dist += dd, cnt++)         k-means pass / single-mean seed
//   mode 1: sum[d] += (x[d] - cent[d])^2      (+ cnt++)                      init_mix_param, T-FS:864-932
// all == true: every entry belongs to cell 0 (the seed pass, no classification yet)
__global__ void k_init_accumulate(const double *__restrict__ x, const int32_t *__restrict__ lst, const int32_t *__restrict__ lst_off,
                                  const uint8_t *__restrict__ idx, const double *__restrict__ dd, const double *__restrict__ cent,
                                  int VN, int M, int D, int have, int mode, bool all, double *__restrict__ sum,
                                  double *__restrict__ dist, double *__restrict__ cnt) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= VN * have) return;
  const int vk = w / have, j = w - vk * have;
  const int64_t g = (int64_t)vk * M + j;
  const int e0 = lst_off[vk], e1 = lst_off[vk + 1];
  const int d0 = lane, d1 = lane + 32;
  double s0 = 0.0, s1 = 0.0, ds = 0.0, n_acc = 0.0;
  const double c0 = (mode == 1 && d0 < D) ? cent[g * D + d0] : 0.0, c1 = (mode == 1 && d1 < D) ? cent[g * D + d1] : 0.0;
  for (int eb = e0; eb < e1; eb += 32) {
    // this block of 32 entries: cell and distortion of entry eb + lane, frame id likewise (one coalesced load each)
    const int e = eb + lane;
    const int my_cell = (e < e1) ? (all ? 0 : (int)idx[e]) : -1;
    const int my_frame = (e < e1) ? lst[e] : 0;
    const double my_dd = (e < e1 && dd && !all) ? dd[e] : 0.0;
    unsigned hit = __ballot_sync(0xffffffffu, my_cell == j);
    // The additions stay in entry order (bit-exactness); the feature loads do not depend on them, so the rows of up to
    // sixteen hits are fetched together before their sums are formed (a dependent load per addition cost ~500 cycles each).
    while (hit) {  // `hit` is warp-uniform
      constexpr int kB = 16;
      double x0[kB], x1[kB], dl[kB];
      int n = 0;
#pragma unroll
      for (int q = 0; q < kB; q++) {
        x0[q] = x1[q] = dl[q] = 0.0;
        if (hit) {
          const int l = __ffs(hit) - 1;
          hit &= hit - 1;
          const int f = __shfl_sync(0xffffffffu, my_frame, l);
          dl[q] = __shfl_sync(0xffffffffu, my_dd, l);
          const double *xr = x + (int64_t)f * D;
          if (d0 < D) x0[q] = xr[d0];
          if (d1 < D) x1[q] = xr[d1];
          n = q + 1;
        }
      }
#pragma unroll
      for (int q = 0; q < kB; q++) {
        if (q < n) {  // in entry order
          if (mode == 0) {
            if (d0 < D) s0 = __dadd_rn(s0, x0[q]);
            if (d1 < D) s1 = __dadd_rn(s1, x1[q]);
            ds = __dadd_rn(ds, dl[q]);
          } else {
            if (d0 < D) { const double a = __dsub_rn(x0[q], c0); s0 = __dadd_rn(s0, __dmul_rn(a, a)); }
            if (d1 < D) { const double a = __dsub_rn(x1[q], c1); s1 = __dadd_rn(s1, __dmul_rn(a, a)); }
          }
          n_acc += 1.0;
        }
      }
    }
  }
  if (d0 < D) sum[g * D + d0] = s0;
  if (d1 < D) sum[g * D + d1] = s1;
  if (lane == 0) {
    cnt[g] = n_acc;
    if (mode == 0 && dist) dist[g] = ds;
  }
}

__device__ inline void init_order_desc(const double *keys, int *idx, int n) {  // order_desc of hmm_host.c (T-FS:1289-1317)
  for (int i = 0; i < n; i++) idx[i] = i;
  for (int again = 1; again;) {
    again = 0;
    for (int i = 0; i + 1 < n; i++)
      if (keys[idx[i]] < keys[idx[i + 1]]) {
        const int t = idx[i]; idx[i] = idx[i + 1]; idx[i + 1] = t;
        again = 1;
      }
  }
}

// One warp per (v, k), lanes over the dimensions (every vector operation of the reference is elementwise in d, and a
// lane keeps the reference's statement order for its own d): new centroids = sum / count, the empty-cell rule
// (T-FS:1236-1269), and -- when this was the last pass of a level and more cells are wanted -- the next split
// (T-FS:1120-1158).  seed: cent[.][0] = sum / cnt only.  have_next = cells after the optional split (== have: no split).
__global__ void k_init_update(double *__restrict__ cent, const double *__restrict__ sum, const double *__restrict__ dist,
                              const double *__restrict__ cnt, int VN, int M, int D, int have, int have_next, int *__restrict__ ord_ws) {
  const int vk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (vk >= VN) return;
  double *ck = cent + (int64_t)vk * M * D;
  const double *sk = sum + (int64_t)vk * M * D, *dk = dist + (int64_t)vk * M, *nk = cnt + (int64_t)vk * M;
  int *ord = ord_ws + (int64_t)vk * M;
  for (int j = 0; j < have; j++)
    for (int d = lane; d < D; d += 32) ck[(int64_t)j * D + d] = sk[(int64_t)j * D + d] / nk[j];
  if (have > 1 || have_next > have) {
    if (lane == 0) init_order_desc(dk, ord, have);
    __syncwarp();
  }
  if (have > 1) {  // an empty cell is refilled from the most distorted ones
    int nxt = 0;
    for (int j = 0; j < have; j++)
      if (nk[j] == 0.0) {
        const int src = ord[nxt++];
        for (int d = lane; d < D; d += 32) {
          ck[(int64_t)j * D + d] = ck[(int64_t)src * D + d] * (1.005);
          ck[(int64_t)src * D + d] = ck[(int64_t)src * D + d] * (0.995);
        }
      }
  }
  if (have_next > have) {
    if (have_next == 2 * have && 2 * have < M) {  // doubling
      for (int j = 0; j < have; j++)
        for (int d = lane; d < D; d += 32) {
          ck[(int64_t)(have + j) * D + d] = ck[(int64_t)j * D + d] * (1.005);
          ck[(int64_t)j * D + d] = ck[(int64_t)j * D + d] * (0.995);
        }
    } else {  // split the M - have most distorted cells
      for (int j = 0; j < have_next - have; j++) {
        const int src = ord[j];
        for (int d = lane; d < D; d += 32) {
          ck[(int64_t)(have + j) * D + d] = ck[(int64_t)src * D + d] * (1.005);
          ck[(int64_t)src * D + d] = ck[(int64_t)src * D + d] * (0.995);
        }
      }
    }
  }
}

// Final parameters (init_mix_param, T-FS:864-932): one thread per (v, k).  sq = sum of squared deviations,
// cnt = cluster sizes, dur[vk] = frames of the state.  A = uniform band (T-FS:774-795) by thread k == 0.
__global__ void k_init_finish(const double *__restrict__ cent, const double *__restrict__ sq, const double *__restrict__ cnt,
                              const int32_t *__restrict__ lst_off, int V, int N, int M, int D, double *__restrict__ A,
                              double *__restrict__ c, double *__restrict__ mu, double *__restrict__ iv, double *__restrict__ det) {
  const int vk = blockIdx.x * blockDim.x + threadIdx.x;
  if (vk >= V * N) return;
  const int v = vk / N, i = vk - v * N;
  const double dur = (double)(lst_off[vk + 1] - lst_off[vk]);
  for (int j = 0; j < M; j++) {
    const int64_t g = (int64_t)vk * M + j;
    double dt = 1.0;
    for (int d = 0; d < D; d++) {
      double var = sq[g * D + d] / cnt[g];
      if (var < 1.0e-5) var = 1.0e-5;  // FINITE_PROBAB, T-FS:904-905
      iv[g * D + d] = var;
      mu[g * D + d] = cent[g * D + d];
    }
    for (int d = 0; d < D; d++) dt *= iv[g * D + d];             // calc_det
    for (int d = 0; d < D; d++) iv[g * D + d] = 1.0 / iv[g * D + d];  // inv_matrix
    det[g] = dt;
    c[g] = cnt[g] / dur;
  }
  // changing_zero_coef, T-FS:1338-1359
  double s = 0.0;
  for (int j = 0; j < M; j++) {
    const int64_t g = (int64_t)vk * M + j;
    if (c[g] < 1.0e-5) c[g] = 1.0e-5;
    s += c[g];
  }
  for (int j = 0; j < M; j++) c[(int64_t)vk * M + j] /= s;
  for (int j = 0; j < N; j++) {
    double a = 0.0;
    if (j >= i && j <= i + 1) a = (2 > N - i) ? 1.0 / (double)(N - i) : 1.0 / 2.0;  // DELTA = 1
    A[((int64_t)v * N + i) * N + j] = a;
  }
}

}  // namespace hmmk
