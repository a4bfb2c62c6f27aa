/*
 * train_main.c -- drop-in for the reference trainer's main() (T-FS:101-391):
 *   hmm_continuous_fs word N P M_1..M_P list_1..list_P out.hmm [init.hmm]
 * Same argv, same list / feature / .hmm / .txt formats, same exit codes (usage or any I/O error:
 * message on stdout, exit(1)).  The E-step runs on the GPU through hmm_cuda.h.
 * Differences, all deliberate: P (feature streams) must be 1 (SURVEY #16); the optional initial
 * model is read from argv[argc-1] (the reference reads argv[argc] == NULL and would crash,
 * T-FS:216-222); capacity limits are runtime values.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/times.h>
#include <time.h>

#include "hmm_cuda.h"

#define NAME_SIZE 100

static void die(const char *fmt, const char *arg) {
  printf(fmt, arg);
  exit(1);
}

/* writing_text T-FS:2168-2265 */
static void write_report(const char *txt, const char *hmm, const char *word, int N, int M, const char *list,
                         const char *t0, const char *t1, const char *cpu, int exemplars, double mean, int iters) {
  FILE *f = fopen(txt, "w");
  if (!f) die("can't open file %s \n", txt);
  fprintf(f, "Continuous HMM created using forward backward algorithm (diagonal covariance matrix). It is considered a final state.\n");
  fprintf(f, "model file: %s \n", hmm);
  fprintf(f, "word: %s \n", word);
  fprintf(f, "number of states: %d \n", N);
  fprintf(f, "number of parameters: %d \n", 1);
  fprintf(f, "number of mixtures %d: %d \n", 1, M);
  fprintf(f, "parameter %d: %s \n", 1, list);
  fprintf(f, "threshould to finish training: %f \n", 1.0e-3);
  fprintf(f, "number of exemplars in training sequence: %d \n", exemplars);
  fprintf(f, "mean probability: %f \n", mean);
  fprintf(f, "number of iterations: %d \n", iters);
  fprintf(f, "starting time: %s \n", t0);
  fprintf(f, "ending time: %s \n", t1);
  fprintf(f, "cpu time: %s \n", cpu);
  if (ferror(f)) die("writing error on file %s \n", txt);
  fclose(f);
}

int hmmh_train_main(int argc, char **argv) {
  char t0[100], t1[100], cpu[100];
  time_t start, end;
  time(&start);
  strftime(t0, sizeof(t0), "%d-%h-%Y %X", localtime(&start));
  if (argc < 7) {
    puts("Usage: hmm_continuous_fs word states_number param_number mix_number1 ... mix_numberN  input_file1 ... input_fileN output_file [initial_model]");
    puts("word: word that will be represented by the model");
    puts("states_number: number of states");
    puts("param_number: number of parameters to train the model");
    puts("mix_number1: number of mixtures per state (parameter 1)");
    puts("mix_numberN: number of mixtures per state (parameter N)");
    puts("input_file1: name of file with names of files with parameters 1 ");
    puts("input_fileN: name of file with names of files with parameters N");
    puts("output_file: output file name");
    puts("initial_model: name of initial model, if there is one");
    exit(1);
  }
  const char *word = argv[1];
  const int N = atoi(argv[2]), P = atoi(argv[3]);
  if (P != 1) die("param_number %s is not supported: this build handles one feature stream \n", argv[3]);
  const int M = atoi(argv[4]);
  const char *list = argv[5], *out = argv[6];
  if (N < 1 || N > HMMCU_MAX_STATES || M < 1) die("bad states/mixtures number (%s) \n", argv[2]);
  char txt[NAME_SIZE + 8];
  strncpy(txt, out, NAME_SIZE);
  txt[NAME_SIZE - 1] = 0;
  strtok(txt, "."); /* T-FS:205-207 */
  strcat(txt, ".txt");

  /* the whole training list, once (the reference re-reads every file twice per iteration): the many-files
   * reader streams it into HBM through pinned staging, no host copy is kept (ingest.c) */
  char **paths = NULL;
  int U = 0, D = 0, bad = -1;
  if (hmmh_read_list(list, &paths, &U) != HMMCU_OK) die("file %s not found \n", list);
  if (U == 0) die("file %s not found \n", list);
  for (int u = 0; u < U; u++) printf("\r\nOpenning %s", paths[u]);
  int64_t *off = (int64_t *)malloc(sizeof(int64_t) * ((size_t)U + 1));
  if (!off) die("error on allocating memory. %s\n", "");

  hmmcu_ctx *ctx = NULL;
  if (hmmcu_create(0, &ctx) != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(NULL));
  int rc = hmmh_ingest(ctx, (const char *const *)paths, U, 0, off, &D, &bad, NULL);
  if (rc == HMMCU_EIO) die("file %s not found \n", bad >= 0 ? paths[bad] : list);
  if (rc != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(ctx));
  int32_t *u2m = (int32_t *)calloc((size_t)U, sizeof(int32_t));

  hmmh_model m;
  memset(&m, 0, sizeof(m));
  if (argc == 2 * P + 6) {
    if (hmmh_read_model(argv[argc - 1], &m, 0) != HMMCU_OK) die("reading error on file %s \n", argv[argc - 1]);
    if (m.D != D) die("reading error on file %s \n", argv[argc - 1]);
  } else {
    if (hmmh_model_alloc(&m, N, M, D) != HMMCU_OK) die("error on allocating memory. %s\n", "");
    /* creating_initial_model (T-FS:226, 732-1317): on the device (bit-identical to the host version, which stays
     * as the path for feature widths / mixture counts the device builder does not take) */
    if (D <= 64 && M <= 255 && N <= 8) {
      if (hmmcu_init_models(ctx, u2m, 1, N, M) != HMMCU_OK || hmmcu_get_models(ctx, m.A, m.c, m.mu, m.inv_var, m.det) != HMMCU_OK)
        die("GPU error: %s \n", hmmcu_last_error(ctx));
    } else { /* shapes the device builder does not take: the host builder, from a host copy of the corpus */
      double *x = (double *)malloc(sizeof(double) * (size_t)off[U] * D);
      if (!x) die("error on allocating memory. %s\n", "");
      for (int u = 0; u < U; u++) {
        double *xu; int T, d;
        if (hmmh_read_features(paths[u], &xu, &T, &d) != HMMCU_OK || d != D || T != (int)(off[u + 1] - off[u])) die("reading error on file %s \n", paths[u]);
        memcpy(x + off[u] * D, xu, sizeof(double) * (size_t)T * D);
        free(xu);
      }
      hmmh_init_model(&m, x, off, U);
      free(x);
    }
  }
  memset(m.word, 0, sizeof(m.word));
  strncpy(m.word, word, sizeof(m.word) - 1);

  double mean = 0.0;
  int iters = 0;
  printf("\r\nCreating HMM using Forward-Backward algorithm (Baum-Welch)");
  if (hmmh_train(ctx, &m, 1, u2m, U, &mean, &iters, 0, NULL, NULL) != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(ctx));
  hmmcu_destroy(ctx);

  /* cpu time exactly as the reference formats it (T-FS:364-369) */
  struct tms tb;
  times(&tb);
  time_t cpu_s = (int)(tb.tms_utime / 60.0);
  struct tm *ct = gmtime(&cpu_s);
  ct->tm_mday -= 1;
  strftime(cpu, sizeof(cpu), "%d %X", ct);
  time(&end);
  strftime(t1, sizeof(t1), "%d-%h-%Y %X", localtime(&end));

  if (hmmh_write_model(out, &m) != HMMCU_OK) die("can't open file %s \n", out);
  write_report(txt, out, word, m.N, m.M, list, t0, t1, cpu, U, mean, iters);
  printf("\r\nmean probability: %f, iterations: %d\r\n", mean, iters);
  hmmh_model_free(&m);
  hmmh_free_list(paths, U);
  free(off); free(u2m);
  return 0;
}
