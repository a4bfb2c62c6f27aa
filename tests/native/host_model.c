/* TEST INFRASTRUCTURE: the single-threaded host model code (initial-model builder, M-step, .hmm files with one and two
 * streams) under AddressSanitizer + UndefinedBehaviorSanitizer.  argv[1] = scratch directory. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hmm_cuda.h"
static double rnd(unsigned *s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) & 0xffff) / 65536.0 - 0.5; }
int main(int argc, char **argv) {
  unsigned seed = 7;
  const int U = 6, D = 9, N = 5;
  int64_t off[7] = {0};
  for (int u = 0; u < U; u++) off[u + 1] = off[u] + 5 + 13 * u;   /* includes an utterance as short as the state chain */
  double *x = malloc(sizeof(double) * off[U] * D);
  for (int64_t k = 0; k < off[U] * D; k++) x[k] = rnd(&seed) + (k % D);
  int bad = 0;
  for (int M = 1; M <= 7; M++) {  /* doubling, partial split, more cells than some states have frames */
    hmmh_model m;
    if (hmmh_model_alloc(&m, N, M, D)) return 2;
    hmmh_init_model(&m, x, off, U);
    double *st = calloc((size_t)(N * N + 2 * N + N * M + 2 * N * M * D + 2), sizeof(double));
    for (int k = 0; k < N * N + 2 * N + N * M + 2 * N * M * D; k++) st[k] = 1.0 + fabs(rnd(&seed));
    hmmh_mstep(&m, st);
    strcpy(m.word, "word");
    char p[512];
    snprintf(p, sizeof p, "%s/m%d.hmm", argv[1], M);
    hmmh_model back[2];
    int P = 0;
    if (hmmh_write_model(p, &m) || hmmh_read_model(p, &back[0], 0)) bad++;
    else {
      if (memcmp(back[0].mu, m.mu, sizeof(double) * N * M * D) || strcmp(back[0].word, "word")) bad++;
      hmmh_model two[2] = {m, back[0]};
      snprintf(p, sizeof p, "%s/p%d.hmm", argv[1], M);
      hmmh_model rd[2];
      if (hmmh_write_model_streams(p, two, 2) || hmmh_read_model_streams(p, rd, 2, &P, 0) || P != 2) bad++;
      else {
        if (memcmp(rd[1].inv_var, m.inv_var, sizeof(double) * N * M * D)) bad++;
        if (hmmh_read_model(p, &back[1], 0) == HMMCU_OK) bad++;  /* a two-stream file is refused by the one-stream reader */
        hmmh_model_free(&rd[0]); hmmh_model_free(&rd[1]);
      }
      hmmh_model_free(&back[0]);
    }
    free(st);
    hmmh_model_free(&m);
  }
  free(x);
  printf("host model code: bad=%d\n", bad);
  return bad != 0;
}
