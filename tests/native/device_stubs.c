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
#ifndef WITH_HMM_HOST
int hmmh_write_features(const char *path, const double *x, int T, int D) {
  FILE *f = fopen(path, "wb"); if (!f) return 5;
  fwrite(&D, sizeof(int), 1, f); fwrite(x, sizeof(double) * D, (size_t)T, f); fclose(f); return 0; }
#endif
/* the EM loop's device calls (never reached by the sanitizer drivers) */
int hmmcu_em_reset(hmmcu_ctx *c){return 0;}
int hmmcu_estep(hmmcu_ctx *c, const int32_t *m, double *s, double *l){return 0;}
double *hmmcu_stats_device(hmmcu_ctx *c, int64_t *n){return 0;}
void *hmmcu_stream(hmmcu_ctx *c){return 0;}
int hmmcu_mstep(hmmcu_ctx *c, double t, double *a, double *b, int32_t *u){return 0;}
int hmmcu_get_models(hmmcu_ctx *ctx, double *A, double *c, double *mu, double *inv_var, double *det){return 0;}
int hmmcu_link_streams(hmmcu_ctx *p, hmmcu_ctx *const *o, int n){return 0;}
