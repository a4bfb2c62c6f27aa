/*
 * oracle/hmm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, double-precision restatement of the diagonal-covariance continuous-HMM hot path
 * of edielsonpf/speech-recognition-hmm-continuous.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file; the product (the CUDA
 * library behind include/hmm_cuda.h) never does.
 *
 * "T-FS" = /root/reference/train/source/hmm-fs/hmm_continuous_fs.c
 * "R-FS" = /root/reference/test/source/recognition-fs/recognition_continuous_fs.c
 *
 * Parity pinning: every function below is checked in tests/test_oracle_vs_ref.py against the
 * reference's own object code (oracle/_ref/libref_*.so, built by oracle/build_ref.sh from the
 * sources where they lie) and against tests/golden/ (KATs produced by the reference binaries:
 * tests/golden/make_golden.py).  The one exception is orc_viterbi(): the reference contains no
 * Viterbi decoder at all, so that function is PARITY UNPINNED -- it is
 * our statement of the algorithm under the reference's conventions (pi = [1,0,..], full A
 * matrix, final-state termination, lowest index wins ties).
 *
 * Layout: everything is flat, row-major, time-major:
 *   x[T][D]  A[N][N]  c[N][M]  mu[N][M][D]  iv[N][M][D] (INVERSE variance, as the reference's
 *   `cov_matrix` holds after init / M-step)  det[N][M] (product of variances)
 *   b[T][N]  post[T][N][M]  alpha[T][N]  beta[T][N]  scale[T]
 * The reference stores [state][time]; the arithmetic and its order are the same.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_THRESHOLD 1.0e-3   /* T-FS:37 THRESHOULD */
#define ORC_DELTA 1            /* T-FS:38 */
#define ORC_FLOOR 1.0e-5       /* T-FS:39 FINITE_PROBAB */

/* ---- A2: one diagonal Gaussian density, linear domain.  T-FS:1804-1841 == R-FS:910-947 ---- */
double orc_gauss(int D, const double *x, const double *mu, const double *iv, double det) {
  double norm = pow(2.0 * M_PI, D / 2.0); /* recomputed per call in the reference */
  double q = 0.0, g = 0.0;                /* the reference returns garbage when det == 0; we return 0 */
  if (det != 0) {
    double sd = pow(fabs(det), 0.5);
    for (int d = 0; d < D; d++) {
      double dif = x[d] - mu[d];
      q += dif * iv[d] * dif;
    }
    q *= -0.5;
    g = exp(q) / (norm * sd);
  }
  return g;
}

/* ---- A3 / A3': per-state mixture density (+ normalised per-mixture posterior when post != NULL)
 * trainer flavour T-FS:1749-1783, recogniser flavour R-FS:860-889 (post == NULL). ---- */
void orc_emissions(int N, int M, int D, int T, const double *x, const double *c, const double *mu,
                   const double *iv, const double *det, double *b, double *post) {
  double *g = (double *)malloc(sizeof(double) * (size_t)M);
  for (int t = 0; t < T; t++) {
    for (int i = 0; i < N; i++) {
      double s = 0.0;
      for (int m = 0; m < M; m++) {
        size_t k = (size_t)i * M + m;
        g[m] = orc_gauss(D, x + (size_t)t * D, mu + k * D, iv + k * D, det[k]);
        g[m] *= c[k];
        s += g[m];
      }
      b[(size_t)t * N + i] = s;
      if (post) {
        double *p = post + ((size_t)t * N + i) * M;
        if (s != 0.0)
          for (int m = 0; m < M; m++) p[m] = g[m] / s;
        else
          for (int m = 0; m < M; m++) p[m] = 0.0;
      }
    }
  }
  free(g);
}

/* ---- A4: scaled forward recursion, pi = [1,0,...,0].  T-FS:1380-1443 == R-FS:739-799 ---- */
void orc_forward(int N, int T, const double *A, const double *b, double *alpha, double *scale) {
  double sum = 0.0;
  for (int i = 0; i < N; i++) {
    alpha[i] = (i == 0 ? 1 : 0) * b[i];
    sum += alpha[i];
  }
  scale[0] = 1.0 / sum;
  for (int i = 0; i < N; i++) alpha[i] *= scale[0];
  for (int t = 1; t < T; t++) {
    const double *ap = alpha + (size_t)(t - 1) * N;
    double *at = alpha + (size_t)t * N;
    sum = 0.0;
    for (int i = 0; i < N; i++) {
      double aux = 0.0;
      for (int j = 0; j < N; j++) aux += ap[j] * A[j * N + i]; /* ALL predecessors, not the band */
      at[i] = aux * b[(size_t)t * N + i];
      sum += at[i];
    }
    scale[t] = 1.0 / sum;
    for (int i = 0; i < N; i++) at[i] *= scale[t];
  }
}

/* ---- A5: scaled backward recursion, FINAL STATE ONLY initialisation.  T-FS:1463-1516 ---- */
void orc_backward(int N, int T, const double *A, const double *b, const double *scale, double *beta) {
  double *bl = beta + (size_t)(T - 1) * N;
  for (int i = 0; i < N - 1; i++) bl[i] = 0.0;
  bl[N - 1] = 1.0;
  bl[N - 1] *= scale[T - 1];
  for (int t = T - 2; t >= 0; t--) {
    const double *bn = beta + (size_t)(t + 1) * N;
    double *bt = beta + (size_t)t * N;
    for (int i = 0; i < N; i++) {
      double aux = 0.0;
      for (int j = 0; j < N; j++) aux += bn[j] * A[i * N + j] * b[(size_t)(t + 1) * N + j];
      bt[i] = aux;
    }
    for (int i = 0; i < N; i++) bt[i] *= scale[t];
  }
}

/* ---- A6: log P(O, q_T = N-1) = -sum log c_t + log alpha^[T-1][N-1].  T-FS:1536-1553 ---- */
double orc_logprob(int N, int T, const double *alpha, const double *scale) {
  double p = 0.0;
  for (int t = 0; t < T; t++) p -= log(scale[t]);
  p += log(alpha[(size_t)(T - 1) * N + (N - 1)]);
  return p;
}

/* ---- A7 + A8 + A9: Baum-Welch accumulators of ONE utterance, added into the running sums.
 * T-FS:1577-1620 (transitions, band i <= j <= i+DELTA only), 1642-1664 (den_mix, includes the
 * last frame), 1691-1727 (per-mixture S0/S1 and the variance around the OLD mean). ---- */
void orc_accumulate(int N, int M, int D, int T, const double *x, const double *A, const double *mu,
                    const double *b, const double *post, const double *alpha, const double *beta,
                    const double *scale, double *num_trans, double *den_trans, double *den_mix,
                    double *S0, double *S1, double *S2c) {
  for (int i = 0; i < N; i++) {
    for (int j = 0; j < N; j++) {
      if (j >= i && j < i + ORC_DELTA + 1) {
        double aux = 0.0;
        for (int t = 0; t < T - 1; t++)
          aux += alpha[(size_t)t * N + i] * A[i * N + j] * b[(size_t)(t + 1) * N + j] *
                 beta[(size_t)(t + 1) * N + j];
        num_trans[i * N + j] += aux;
      }
    }
    for (int t = 0; t < T - 1; t++)
      den_trans[i] += alpha[(size_t)t * N + i] * beta[(size_t)t * N + i] / scale[t];
  }
  for (int i = 0; i < N; i++)
    for (int t = 0; t < T; t++) {
      double aux = alpha[(size_t)t * N + i] * beta[(size_t)t * N + i] / scale[t];
      den_mix[i] += aux;
    }
  for (int t = 0; t < T; t++) {
    const double *xt = x + (size_t)t * D;
    for (int i = 0; i < N; i++) {
      double gam = alpha[(size_t)t * N + i] * beta[(size_t)t * N + i] / scale[t];
      for (int m = 0; m < M; m++) {
        size_t k = (size_t)i * M + m;
        double w = gam * post[((size_t)t * N + i) * M + m];
        S0[k] += w;
        for (int d = 0; d < D; d++) {
          S1[k * D + d] += w * xt[d];
          double dif = xt[d] - mu[k * D + d];
          dif *= dif;
          S2c[k * D + d] += w * dif;
        }
      }
    }
  }
}

/* T-FS:1338-1359 */
static void orc_floor_weights(int M, double *c) {
  double sum = 0.0;
  for (int m = 0; m < M; m++) {
    if (c[m] < ORC_FLOOR) c[m] = ORC_FLOOR;
    sum += c[m];
  }
  for (int m = 0; m < M; m++) c[m] /= sum;
}

/* ---- A10: M-step.  T-FS:1862-1889 (transitions), 1911-1955 (mixtures), then det / inverse for
 * EVERY mixture exactly as main() does at T-FS:339-346 -- including states whose den_mix is 0,
 * whose `iv` therefore still holds the old inverse and gets inverted again (a reference quirk
 * that is reproduced, not fixed). ---- */
void orc_mstep(int N, int M, int D, const double *num_trans, const double *den_trans,
               const double *den_mix, const double *S0, const double *S1, const double *S2c,
               double *A, double *c, double *mu, double *iv, double *det) {
  for (int i = 0; i < N; i++)
    if (den_trans[i] != 0.0)
      for (int j = 0; j < N; j++) A[i * N + j] = num_trans[i * N + j] / den_trans[i];
  for (int i = 0; i < N; i++)
    if (den_mix[i] != 0.0)
      for (int m = 0; m < M; m++) {
        size_t k = (size_t)i * M + m;
        c[k] = S0[k] / den_mix[i];
        for (int d = 0; d < D; d++) {
          mu[k * D + d] = S1[k * D + d] / S0[k];
          iv[k * D + d] = S2c[k * D + d] / S0[k]; /* variance for now */
          if (iv[k * D + d] < ORC_FLOOR) iv[k * D + d] = ORC_FLOOR;
        }
      }
  for (int i = 0; i < N; i++) orc_floor_weights(M, c + (size_t)i * M);
  for (size_t k = 0; k < (size_t)N * M; k++) {
    double p = 1.0;
    for (int d = 0; d < D; d++) p *= iv[k * D + d]; /* calc_det T-FS:1976-1991 */
    det[k] = p;
    for (int d = 0; d < D; d++) iv[k * D + d] = 1.0 / iv[k * D + d]; /* inv_matrix T-FS:2012-2022 */
  }
}

/* ---- One E-step over a whole training set (the body of the do-loop, T-FS:240-321).
 * Returns sum_u logP_u.  off[U+1] are frame offsets into x[F][D]. ---- */
double orc_estep(int N, int M, int D, int U, const long long *off, const double *x, const double *A,
                 const double *c, const double *mu, const double *iv, const double *det,
                 double *num_trans, double *den_trans, double *den_mix, double *S0, double *S1,
                 double *S2c, double *logp_utt /* U or NULL */) {
  size_t nm = (size_t)N * M;
  memset(num_trans, 0, sizeof(double) * N * N);
  memset(den_trans, 0, sizeof(double) * N);
  memset(den_mix, 0, sizeof(double) * N);
  memset(S0, 0, sizeof(double) * nm);
  memset(S1, 0, sizeof(double) * nm * D);
  memset(S2c, 0, sizeof(double) * nm * D);
  int Tmax = 0;
  for (int u = 0; u < U; u++)
    if ((int)(off[u + 1] - off[u]) > Tmax) Tmax = (int)(off[u + 1] - off[u]);
  double *b = (double *)malloc(sizeof(double) * (size_t)Tmax * N);
  double *post = (double *)malloc(sizeof(double) * (size_t)Tmax * nm);
  double *alpha = (double *)malloc(sizeof(double) * (size_t)Tmax * N);
  double *beta = (double *)malloc(sizeof(double) * (size_t)Tmax * N);
  double *scale = (double *)malloc(sizeof(double) * (size_t)Tmax);
  double total = 0.0;
  for (int u = 0; u < U; u++) {
    int T = (int)(off[u + 1] - off[u]);
    const double *xu = x + (size_t)off[u] * D;
    orc_emissions(N, M, D, T, xu, c, mu, iv, det, b, post);
    orc_forward(N, T, A, b, alpha, scale);
    orc_backward(N, T, A, b, scale, beta);
    orc_accumulate(N, M, D, T, xu, A, mu, b, post, alpha, beta, scale, num_trans, den_trans, den_mix,
                   S0, S1, S2c);
    double lp = orc_logprob(N, T, alpha, scale);
    if (logp_utt) logp_utt[u] = lp;
    total += lp;
  }
  free(b); free(post); free(alpha); free(beta); free(scale);
  return total;
}

/* ---- A11: the EM control loop, T-FS:238-361.  The model is updated only while the relative
 * change exceeds 1e-3, so the model written is the one that produced the last `probab`.
 * Returns the number of iterations; *mean_logp = probab / U of the last E-step. ---- */
int orc_train(int N, int M, int D, int U, const long long *off, const double *x, double *A, double *c,
              double *mu, double *iv, double *det, double *mean_logp, int max_iter) {
  size_t nm = (size_t)N * M;
  double *num_trans = (double *)malloc(sizeof(double) * N * N);
  double *den_trans = (double *)malloc(sizeof(double) * N);
  double *den_mix = (double *)malloc(sizeof(double) * N);
  double *S0 = (double *)malloc(sizeof(double) * nm);
  double *S1 = (double *)malloc(sizeof(double) * nm * D);
  double *S2c = (double *)malloc(sizeof(double) * nm * D);
  double old = 1.0, probab = 0.0, var;
  int it = 0;
  do {
    it++;
    probab = orc_estep(N, M, D, U, off, x, A, c, mu, iv, det, num_trans, den_trans, den_mix, S0, S1,
                       S2c, NULL);
    var = fabs((old - probab) / old);
    if (var > ORC_THRESHOLD) {
      old = probab;
      orc_mstep(N, M, D, num_trans, den_trans, den_mix, S0, S1, S2c, A, c, mu, iv, det);
    }
  } while (var > ORC_THRESHOLD && (max_iter <= 0 || it < max_iter));
  *mean_logp = probab / (double)U;
  free(num_trans); free(den_trans); free(den_mix); free(S0); free(S1); free(S2c);
  return it;
}

/* ---- R1 (one cell): forward score of one utterance against one model, R-FS:349-367 ---- */
double orc_forward_score(int N, int M, int D, int T, const double *x, const double *A, const double *c,
                         const double *mu, const double *iv, const double *det) {
  double *b = (double *)malloc(sizeof(double) * (size_t)T * N);
  double *alpha = (double *)malloc(sizeof(double) * (size_t)T * N);
  double *scale = (double *)malloc(sizeof(double) * (size_t)T);
  orc_emissions(N, M, D, T, x, c, mu, iv, det, b, NULL);
  orc_forward(N, T, A, b, alpha, scale);
  double lp = orc_logprob(N, T, alpha, scale);
  free(b); free(alpha); free(scale);
  return lp;
}

/* ---- R2: bubble sort of the index array, descending, strict '<' (stable; a NaN never moves).
 * R-FS:968-995.  index[0] is the recognised label (R3), index[1] the "second candidate". ---- */
void orc_rank(int V, const double *score, int *index) {
  int done = 0;
  for (int i = 0; i < V; i++) index[i] = i;
  while (!done) {
    done = 1;
    for (int i = 0; i < V - 1; i++)
      if (score[index[i]] < score[index[i + 1]]) {
        int t = index[i];
        index[i] = index[i + 1];
        index[i + 1] = t;
        done = 0;
      }
  }
}

/* ---- V1: Viterbi with back-pointers.  NOT IN THE REFERENCE (parity unpinned).  Conventions:
 * emissions are log of the reference's linear densities, log 0 = -inf, pi = [1,0,..], the full A
 * matrix, termination in the final state N-1 (as A5/A6), lowest predecessor index wins a tie.
 * Returns the best-path log score; path[T] receives the state sequence. ---- */
double orc_viterbi(int N, int T, const double *A, const double *b, int *path) {
  double *delta = (double *)malloc(sizeof(double) * (size_t)N * 2);
  unsigned char *psi = (unsigned char *)malloc((size_t)T * N);
  double *la = (double *)malloc(sizeof(double) * N * N);
  for (int k = 0; k < N * N; k++) la[k] = log(A[k]);
  double *cur = delta, *nxt = delta + N;
  for (int i = 0; i < N; i++) {
    cur[i] = (i == 0 ? 0.0 : -INFINITY) + log(b[i]);
    psi[i] = 0;
  }
  for (int t = 1; t < T; t++) {
    for (int j = 0; j < N; j++) {
      double best = cur[0] + la[0 * N + j];
      int arg = 0;
      for (int i = 1; i < N; i++) {
        double v = cur[i] + la[i * N + j];
        if (v > best) { best = v; arg = i; }
      }
      nxt[j] = best + log(b[(size_t)t * N + j]);
      psi[(size_t)t * N + j] = (unsigned char)arg;
    }
    double *tmp = cur; cur = nxt; nxt = tmp;
  }
  double score = cur[N - 1];
  int s = N - 1;
  for (int t = T - 1; t >= 0; t--) {
    path[t] = s;
    s = psi[(size_t)t * N + s];
  }
  free(delta); free(psi); free(la);
  return score;
}

/* =====================  initial-model builder  ===================== */

/* T-FS:1289-1317 */
static void orc_sort_desc(const double *v, int *index, int n) {
  int done = 0;
  for (int i = 0; i < n; i++) index[i] = i;
  while (!done) {
    done = 1;
    for (int i = 0; i < n - 1; i++) {
      int j = index[i], k = index[i + 1];
      if (v[j] < v[k]) { index[i] = k; index[i + 1] = j; done = 0; }
    }
  }
}

/* T-FS:1179-1215.  *index keeps its previous value when no distance beats 1e20. */
static double orc_classify(const double *x, int nmix, int D, const double *mean, int *index) {
  double best = 1.0e20;
  for (int i = 0; i < nmix; i++) {
    double dist = 0.0;
    for (int d = 0; d < D; d++) {
      double a = mean[(size_t)i * D + d] - x[d];
      dist += a * a;
    }
    if (dist < best) { best = dist; *index = i; }
  }
  return best;
}

/* T-FS:1120-1158 */
static int orc_split(int D, int old, int M, double *mean, const double *distortion) {
  if (2 * old < M) {
    for (int k = 0; k < old; k++) {
      for (int d = 0; d < D; d++) mean[(size_t)(old + k) * D + d] = mean[(size_t)k * D + d] * (1.005);
      for (int d = 0; d < D; d++) mean[(size_t)k * D + d] = mean[(size_t)k * D + d] * (0.995);
    }
    return 2 * old;
  }
  int *index = (int *)malloc(sizeof(int) * (size_t)(old > 0 ? old : 1));
  orc_sort_desc(distortion, index, old);
  int dif = M - old;
  for (int k = 0; k < dif; k++) {
    int i = index[k];
    for (int d = 0; d < D; d++) mean[(size_t)(old + k) * D + d] = mean[(size_t)i * D + d] * (1.005);
    for (int d = 0; d < D; d++) mean[(size_t)i * D + d] = mean[(size_t)i * D + d] * (0.995);
  }
  free(index);
  return dif + old;
}

/* T-FS:1236-1269 */
static void orc_new_means(int D, int nmix, const double *sum, const int *count, double *mean,
                          const double *distortion) {
  int *index = (int *)malloc(sizeof(int) * (size_t)nmix);
  for (int j = 0; j < nmix; j++)
    for (int d = 0; d < D; d++) mean[(size_t)j * D + d] = sum[(size_t)j * D + d] / (double)count[j];
  orc_sort_desc(distortion, index, nmix);
  int i = 0;
  for (int j = 0; j < nmix; j++)
    if (count[j] == 0) {
      int l = index[i++];
      for (int d = 0; d < D; d++) mean[(size_t)j * D + d] = mean[(size_t)l * D + d] * (1.005);
      for (int d = 0; d < D; d++) mean[(size_t)l * D + d] = mean[(size_t)l * D + d] * (0.995);
    }
  free(index);
}

/* uniform segmentation used by every init pass, T-FS:1005-1013 */
static void orc_segment(int T, int N, int k, int *begin, int *end) {
  int q = T / N, r = T % N, e = 0, b = 0;
  for (int s = 0; s <= k; s++) {
    b = e;
    e += (s < r) ? q + 1 : q;
  }
  *begin = b;
  *end = e;
}

/* ---- creating_initial_model, T-FS:732-1317: uniform left-to-right A, uniform segmentation,
 * LBG-style splitting with 3 k-means passes per split, per-cluster variances and weights.
 * `distortion[k][index] += classifying(.., &index)` (T-FS:1076) has unspecified evaluation order in
 * C; gcc 13 -O2 on x86-64 reads `index` AFTER the call (credit goes to the cell the frame was just
 * assigned to) -- pinned against oracle/_ref in tests/test_oracle_vs_ref.py. ---- */
void orc_init_model(int N, int M, int D, int U, const long long *off, const double *x, double *A,
                    double *c, double *mu, double *iv, double *det) {
  size_t nm = (size_t)N * M;
  /* init_transition_probab T-FS:774-795 */
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      if (j > ORC_DELTA + i || j < i) A[i * N + j] = 0;
      else if (ORC_DELTA + 1 > N - i) A[i * N + j] = 1.0 / (double)(N - i);
      else A[i * N + j] = 1.0 / (double)(ORC_DELTA + 1);
    }
  double *mean = (double *)calloc(nm * D, sizeof(double));
  double *sum = (double *)calloc(nm * D, sizeof(double));
  double *distortion = (double *)calloc(nm, sizeof(double));
  int *count = (int *)calloc(nm, sizeof(int));
  int index = 0;
  /* single-mean seed T-FS:996-1030 */
  for (int u = 0; u < U; u++) {
    int T = (int)(off[u + 1] - off[u]);
    const double *xu = x + (size_t)off[u] * D;
    for (int k = 0; k < N; k++) {
      int b, e;
      orc_segment(T, N, k, &b, &e);
      for (int j = b; j < e; j++) {
        for (int d = 0; d < D; d++) mean[((size_t)k * M) * D + d] += xu[(size_t)j * D + d];
        count[(size_t)k * M]++;
      }
    }
  }
  for (int k = 0; k < N; k++)
    for (int d = 0; d < D; d++) mean[((size_t)k * M) * D + d] /= (double)count[(size_t)k * M];
  int old = 1;
  while (old < M) {
    int nm_new = old;
    for (int k = 0; k < N; k++)
      nm_new = orc_split(D, old, M, mean + (size_t)k * M * D, distortion + (size_t)k * M);
    old = nm_new;
    for (int ite = 0; ite < 3; ite++) {
      for (int k = 0; k < N; k++)
        for (int i = 0; i < old; i++) {
          count[(size_t)k * M + i] = 0;
          distortion[(size_t)k * M + i] = 0.0;
          for (int d = 0; d < D; d++) sum[((size_t)k * M + i) * D + d] = 0.0;
        }
      for (int u = 0; u < U; u++) {
        int T = (int)(off[u + 1] - off[u]);
        const double *xu = x + (size_t)off[u] * D;
        for (int k = 0; k < N; k++) {
          int b, e;
          orc_segment(T, N, k, &b, &e);
          for (int j = b; j < e; j++) {
            double dist = orc_classify(xu + (size_t)j * D, old, D, mean + (size_t)k * M * D, &index);
            distortion[(size_t)k * M + index] += dist;
            count[(size_t)k * M + index]++;
            for (int d = 0; d < D; d++) sum[((size_t)k * M + index) * D + d] += xu[(size_t)j * D + d];
          }
        }
      }
      for (int k = 0; k < N; k++)
        orc_new_means(D, old, sum + (size_t)k * M * D, count + (size_t)k * M, mean + (size_t)k * M * D,
                      distortion + (size_t)k * M);
    }
  }
  /* init_mix_param T-FS:864-932 */
  int *dur = (int *)calloc((size_t)N, sizeof(int));
  for (size_t k = 0; k < nm; k++) c[k] = 0.0;
  for (size_t k = 0; k < nm * D; k++) iv[k] = 0.0;
  for (int u = 0; u < U; u++) {
    int T = (int)(off[u + 1] - off[u]);
    const double *xu = x + (size_t)off[u] * D;
    for (int k = 0; k < N; k++) {
      int b, e;
      orc_segment(T, N, k, &b, &e);
      for (int j = b; j < e; j++) {
        orc_classify(xu + (size_t)j * D, M, D, mean + (size_t)k * M * D, &index);
        for (int d = 0; d < D; d++) {
          double a = xu[(size_t)j * D + d] - mean[((size_t)k * M + index) * D + d];
          iv[((size_t)k * M + index) * D + d] += a * a;
        }
        c[(size_t)k * M + index]++;
      }
      dur[k] += e - b;
    }
  }
  for (size_t k = 0; k < nm; k++) {
    for (int d = 0; d < D; d++) {
      iv[k * D + d] /= c[k];
      if (iv[k * D + d] < ORC_FLOOR) iv[k * D + d] = ORC_FLOOR;
    }
    double p = 1.0;
    for (int d = 0; d < D; d++) p *= iv[k * D + d];
    det[k] = p;
    for (int d = 0; d < D; d++) iv[k * D + d] = 1.0 / iv[k * D + d];
    for (int d = 0; d < D; d++) mu[k * D + d] = mean[k * D + d];
  }
  for (int i = 0; i < N; i++)
    for (int m = 0; m < M; m++) c[(size_t)i * M + m] /= (double)dur[i];
  for (int i = 0; i < N; i++) orc_floor_weights(M, c + (size_t)i * M);
  free(mean); free(sum); free(distortion); free(count); free(dur);
}
