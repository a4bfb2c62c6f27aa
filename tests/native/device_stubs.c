/* TEST INFRASTRUCTURE: link-time stand-ins for the device entry points the host pools reference (no GPU in a sanitizer run). */
#include <stdio.h>
#include "hmm_cuda.h"
int hmmcu_features_begin(hmmcu_ctx *c, const int64_t *o, int U, int D){return 0;}
int hmmcu_features_append(hmmcu_ctx *c, const double *x, int64_t f, int64_t n, int *t){return 0;}
int hmmcu_features_wait(hmmcu_ctx *c, int t){return 0;}
int hmmcu_features_end(hmmcu_ctx *c){return 0;}
void *hmmcu_staging(hmmcu_ctx *c, int s, uint64_t b){return 0;}
int hmmcu_set_models(hmmcu_ctx *ctx, int V, int N, int M, int D, const double *A, const double *c, const double *mu, const double *inv_var, const double *det){return 0;}
#include <stdio.h>
int hmmh_write_features(const char *path, const double *x, int T, int D) {
  FILE *f = fopen(path, "wb"); if (!f) return 5;
  fwrite(&D, sizeof(int), 1, f); fwrite(x, sizeof(double) * D, (size_t)T, f); fclose(f); return 0; }
