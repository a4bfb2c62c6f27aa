/*
 * train_main.c -- drop-in for the reference trainer's main() (T-FS:101-391):
 *   hmm_continuous_fs word N P M_1..M_P list_1..list_P out.hmm [init.hmm]
 * Same argv, same list / feature / .hmm / .txt formats, same exit codes (usage or any I/O error:
 * message on stdout, exit(1)).  The E-step runs on the GPU through hmm_cuda.h.
 * P feature streams are one context each, linked for the E-step (hmmcu_link_streams).
 * Differences, all deliberate: the optional initial
 * model is read from argv[argc-1] (the reference reads argv[argc] == NULL and would crash,
 * T-FS:216-222); capacity limits are runtime values.
 * One extension: `hmm_continuous_fs @jobs.txt` runs one job per line of jobs.txt (each line = the arguments of one
 * ordinary invocation) in ONE process, so that the CUDA context (most of the wall clock of a small job) is created once.
 * HMMCU_TRACE=1 in the environment prints the wall clock of the phases on stderr.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/times.h>
#include <time.h>

#include "hmm_cuda.h"

#define NAME_SIZE 100

/* contexts kept across the jobs of a job file (one per feature stream) */
static hmmcu_ctx *g_keep[HMMH_MAX_STREAMS];
static int g_batch = 0;

static void trace(const char *what) {
  static double last = 0.0;
  static int on = -1;
  if (on < 0) on = getenv("HMMCU_TRACE") != NULL;
  if (!on) return;
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  const double now = ts.tv_sec + 1e-9 * ts.tv_nsec;
  fprintf(stderr, "[hmmcu] %-22s +%.1f ms\n", what, last > 0.0 ? (now - last) * 1e3 : 0.0);
  last = now;
}

static void die(const char *fmt, const char *arg) {
  printf(fmt, arg);
  exit(1);
}

/* writing_text T-FS:2168-2265 */
static void write_report(const char *txt, const char *hmm, const char *word, int N, int P, const int *M, const char *const *list,
                         const char *t0, const char *t1, const char *cpu, int exemplars, double mean, int iters) {
  FILE *f = fopen(txt, "w");
  if (!f) die("can't open file %s \n", txt);
  fprintf(f, "Continuous HMM created using forward backward algorithm (diagonal covariance matrix). It is considered a final state.\n");
  fprintf(f, "model file: %s \n", hmm);
  fprintf(f, "word: %s \n", word);
  fprintf(f, "number of states: %d \n", N);
  fprintf(f, "number of parameters: %d \n", P);
  for (int p = 0; p < P; p++) fprintf(f, "number of mixtures %d: %d \n", p + 1, M[p]);
  for (int p = 0; p < P; p++) fprintf(f, "parameter %d: %s \n", p + 1, list[p]);
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

static int train_jobs(const char *prog, const char *jobfile) {
  FILE *f = fopen(jobfile, "r");
  if (!f) die("file %s not found \n", jobfile);
  char line[8192];
  g_batch = 1;
  while (fgets(line, sizeof(line), f)) {
    char *av[64];
    int n = 0;
    av[n++] = (char *)prog;
    for (char *tok = strtok(line, " \t\r\n"); tok && n < 63; tok = strtok(NULL, " \t\r\n")) av[n++] = tok;
    av[n] = NULL;
    if (n > 1) hmmh_train_main(n, av);
  }
  fclose(f);
  g_batch = 0;
  for (int p = HMMH_MAX_STREAMS - 1; p >= 0; p--) {
    if (g_keep[p]) hmmcu_destroy(g_keep[p]);
    g_keep[p] = NULL;
  }
  return 0;
}

int hmmh_train_main(int argc, char **argv) {
  if (argc == 2 && argv[1][0] == '@' && !g_batch) return train_jobs(argv[0], argv[1] + 1);
  trace("start");
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
  if (P < 1 || P > HMMH_MAX_STREAMS) die("bad param_number (%s) \n", argv[3]);
  if (argc < 2 * P + 5) die("param_number %s does not match the argument list \n", argv[3]);
  int M[HMMH_MAX_STREAMS];
  const char *list[HMMH_MAX_STREAMS];
  for (int p = 0; p < P; p++) {
    M[p] = atoi(argv[4 + p]);
    list[p] = argv[4 + P + p];
    if (M[p] < 1) die("bad states/mixtures number (%s) \n", argv[4 + p]);
  }
  const char *out = argv[4 + 2 * P];
  if (N < 1) die("bad states/mixtures number (%s) \n", argv[2]);
  if (N > HMMCU_MAX_STATES) {
    printf("states_number %d is beyond this build's limit of %d states per model (HMMCU_MAX_STATES) \n", N, HMMCU_MAX_STATES);
    exit(1);
  }
  char txt[NAME_SIZE + 8];
  strncpy(txt, out, NAME_SIZE);
  txt[NAME_SIZE - 1] = 0;
  strtok(txt, "."); /* T-FS:205-207 */
  strcat(txt, ".txt");

  /* the whole training list of every feature stream, once (the reference re-reads every file twice per iteration):
   * the many-files reader streams it into HBM through pinned staging, no host copy is kept (ingest.c).  One context
   * per stream; line i of every list is the same utterance (T-FS:272-290). */
  char **paths[HMMH_MAX_STREAMS];
  hmmcu_ctx *ctxs[HMMH_MAX_STREAMS];
  int64_t *off = NULL;
  int U = 0, D[HMMH_MAX_STREAMS];
  for (int p = 0; p < P; p++) {
    int Up = 0, bad = -1;
    if (hmmh_read_list(list[p], &paths[p], &Up) != HMMCU_OK || Up == 0) die("file %s not found \n", list[p]);
    if (p == 0) U = Up;
    if (Up < U) die("reading error on file %s \n", list[p]);
    for (int u = 0; u < U; u++) printf("\r\nOpenning %s", paths[p][u]);
    int64_t *offp = (int64_t *)malloc(sizeof(int64_t) * ((size_t)U + 1));
    if (!offp) die("error on allocating memory. %s\n", "");
    ctxs[p] = g_batch ? g_keep[p] : NULL;
    if (!ctxs[p] && hmmcu_create(0, &ctxs[p]) != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(NULL));
    if (g_batch) g_keep[p] = ctxs[p];
    trace("context");
    hmmh_ingest_stats ist;
    memset(&ist, 0, sizeof(ist));
    int rc = hmmh_ingest(ctxs[p], (const char *const *)paths[p], U, 0, offp, &D[p], &bad, &ist);
    if (getenv("HMMCU_TRACE"))
      fprintf(stderr, "[hmmcu]   ingest: scan %.1f ms, staging buffers %.1f ms, read + copy %.1f ms, pack %.1f ms\n", ist.scan_s * 1e3, ist.stage_s * 1e3,
              ist.read_s * 1e3, (ist.total_s - ist.scan_s - ist.stage_s - ist.read_s) * 1e3);
    if (rc == HMMCU_EIO) die("file %s not found \n", bad >= 0 ? paths[p][bad] : list[p]);
    if (rc != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(ctxs[p]));
    trace("ingest");
    if (p == 0) off = offp;
    else {
      if (memcmp(off, offp, sizeof(int64_t) * ((size_t)U + 1)) != 0) die("reading error on file %s (the streams of an utterance differ in length) \n", list[p]);
      free(offp);
    }
  }
  hmmcu_ctx *ctx = ctxs[0];
  int32_t *u2m = (int32_t *)calloc((size_t)U, sizeof(int32_t));

  hmmh_model m[HMMH_MAX_STREAMS];
  memset(m, 0, sizeof(m));
  if (argc == 2 * P + 6) {
    int Pf = 0;
    if (hmmh_read_model_streams(argv[argc - 1], m, P, &Pf, 0) != HMMCU_OK || Pf != P) die("reading error on file %s \n", argv[argc - 1]);
    for (int p = 0; p < P; p++) {
      if (m[p].D != D[p]) die("reading error on file %s \n", argv[argc - 1]);
      M[p] = m[p].M; /* the model's own sizes go into the report */
    }
  } else {
    for (int p = 0; p < P; p++) {
      if (hmmh_model_alloc(&m[p], N, M[p], D[p]) != HMMCU_OK) die("error on allocating memory. %s\n", "");
      /* creating_initial_model (T-FS:226, 732-1317), stream by stream: on the device (bit-identical to the host version,
       * which stays as the path for feature widths / mixture counts the device builder does not take) */
      if (D[p] <= 64 && M[p] <= 255 && N <= 8) {
        if (hmmcu_init_models(ctxs[p], u2m, 1, N, M[p]) != HMMCU_OK ||
            hmmcu_get_models(ctxs[p], m[p].A, m[p].c, m[p].mu, m[p].inv_var, m[p].det) != HMMCU_OK)
          die("GPU error: %s \n", hmmcu_last_error(ctxs[p]));
      } else { /* the host builder, from a host copy of the stream */
        double *x = (double *)malloc(sizeof(double) * (size_t)off[U] * D[p]);
        if (!x) die("error on allocating memory. %s\n", "");
        for (int u = 0; u < U; u++) {
          double *xu; int T, d;
          if (hmmh_read_features(paths[p][u], &xu, &T, &d) != HMMCU_OK || d != D[p] || T != (int)(off[u + 1] - off[u])) die("reading error on file %s \n", paths[p][u]);
          memcpy(x + off[u] * D[p], xu, sizeof(double) * (size_t)T * D[p]);
          free(xu);
        }
        hmmh_init_model(&m[p], x, off, U);
        free(x);
      }
    }
  }
  for (int p = 0; p < P; p++) {
    memset(m[p].word, 0, sizeof(m[p].word));
    strncpy(m[p].word, word, sizeof(m[p].word) - 1);
  }

  trace("initial model");
  double mean = 0.0;
  int iters = 0;
  printf("\r\nCreating HMM using Forward-Backward algorithm (Baum-Welch)");
  if (hmmh_train_streams(ctxs, P, m, 1, u2m, U, &mean, &iters, 0, NULL, NULL) != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(ctx));
  trace("EM loop");
  if (!g_batch)
    for (int p = P - 1; p >= 0; p--) hmmcu_destroy(ctxs[p]);
  trace("context released");

  /* cpu time exactly as the reference formats it (T-FS:364-369) */
  struct tms tb;
  times(&tb);
  time_t cpu_s = (int)(tb.tms_utime / 60.0);
  struct tm *ct = gmtime(&cpu_s);
  ct->tm_mday -= 1;
  strftime(cpu, sizeof(cpu), "%d %X", ct);
  time(&end);
  strftime(t1, sizeof(t1), "%d-%h-%Y %X", localtime(&end));

  if (hmmh_write_model_streams(out, m, P) != HMMCU_OK) die("can't open file %s \n", out);
  write_report(txt, out, word, m[0].N, P, M, list, t0, t1, cpu, U, mean, iters);
  printf("\r\nmean probability: %f, iterations: %d\r\n", mean, iters);
  for (int p = 0; p < P; p++) {
    hmmh_model_free(&m[p]);
    hmmh_free_list(paths[p], U);
  }
  free(off); free(u2m);
  trace("model + report written");
  return 0;
}
