/*
 * test_main.c -- drop-in for the reference recogniser's main() (R-FS:87-428):
 *   recognition_continuous_fs K modelslist_1..K weight_1..K featlist_1.. wordsfile resultfile
 * Same argv, list / feature / .hmm formats and result-file text.  All (utterance, model) forward
 * scores are computed on the GPU in one batch; ranking follows sorting_probab including its NaN
 * behaviour.  K model sets per word (weighted sum of their log-probabilities, R-FS:326-364) are taken although the
 * reference is compiled with MAX_MODELS_NUMBER 1 (R-FS:40).  Deliberate limits: the models of one set share one
 * topology; a set whose models carry P feature streams takes P feature lists and is scored through linked contexts
 * (product of the streams' densities, R-FS:341-364).
 * Reproduced quirks: the weight is printed through an int* with "%.2d" (R-FS:1016,1029); the last
 * per-word block lists wrong words only for the first `models_number` vocabulary entries (R-FS:400).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/times.h>
#include <time.h>

#include "hmm_cuda.h"

#define STR 100
#define WSTR 50

/* HMMCU_TRACE=1 in the environment prints the wall clock of the phases on stderr (as in the trainer) */
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

static FILE *g_out;

/* writing_result_word R-FS:1110-1150 */
static void write_word_block(int correct, int error, int second, int nwords, const char *spoken, const int *wrong,
                             char (*words)[64], double cpu_time, int frames) {
  int sum = correct + error;
  double per = (double)correct / (double)sum;
  cpu_time /= sum;
  frames /= sum;
  fprintf(g_out, "\nResults: \n");
  fprintf(g_out, "Spoken word: %s\n", spoken);
  fprintf(g_out, "Correct words: %d\n", correct);
  fprintf(g_out, "Errors: %d\n", error);
  fprintf(g_out, "Percentagen correct : %.2f%%\n", (per * 100.0));
  fprintf(g_out, "Second candidate: %d\n", second);
  if (error != 0) {
    fprintf(g_out, "Wrong words: \n");
    for (int i = 0; i < nwords; i++)
      if (wrong[i] != 0) fprintf(g_out, "%s: %d time%s\n", words[i], wrong[i], wrong[i] == 1 ? "" : "s");
  }
  fprintf(g_out, "Average recognition time: %.2f sec. \n", cpu_time);
  fprintf(g_out, "Average word length: %d frames \n", frames);
}

int hmmh_test_main(int argc, char **argv) {
  char date[100];
  time_t now;
  time(&now);
  strftime(date, sizeof(date), "%d-%h-%Y %X", localtime(&now));
  if (argc < 7) {
    puts("Usage: recognition_continuous_fs models_number  model1 ... modelN coef_model1 ... coef_modelN input_file1 ... input_fileM  word_file output_file");
    puts("models_number: number of model");
    puts("model1: name of file with the name of model 1");
    puts("modelN: name of file with the name of model N");
    puts("coef_model1: weighting coefficient of model 1");
    puts("coef_modelN: weighting coefficient of model N");
    puts("input_file1: name of file with name of files with parameters 1 ");
    puts("input_fileM: name of file with name of files with parameters M ");
    puts("word_file: name of file with the spoken words ");
    puts("output_file: name of output file ");
    exit(1);
  }
  /* K model sets per word, each with its own feature list (one stream each) and weighting coefficient; the
   * score of word k is sum_j coef_j * logP_j(u, k), accumulated in set order from 0.0 (R-FS:284, 326-364) */
  const int K = atoi(argv[1]);
  if (K < 1 || K > 16 || argc < 3 * K + 4) die("models_number %s does not match the argument list \n", argv[1]);
  double weight[16];
  for (int j = 0; j < K; j++) weight[j] = atof(argv[K + 2 + j]);
  const char *words_file = argv[argc - 2], *result = argv[argc - 1];

  /* the reference's order of opening: model lists (R-FS:203), feature lists (R-FS:259-266), words file, result file */
  for (int j = 0; j < K; j++) {
    FILE *f = fopen(argv[2 + j], "rb");
    if (!f) die("file %s not found \n", argv[2 + j]);
    fclose(f);
  }
  for (int a = 2 + 2 * K; a < argc - 2; a++) {
    FILE *f = fopen(argv[a], "r");
    if (!f) die("file %s not found \n", argv[a]);
    fclose(f);
  }
  /* the spoken words: line i pairs with line i of every feature list */
  FILE *fw = fopen(words_file, "r");
  if (!fw) die("file %s not found \n", words_file);
  char (*spoken)[WSTR] = NULL;
  int U = 0;
  char w[WSTR];
  while (fscanf(fw, "%49s", w) == 1) {
    spoken = (char (*)[WSTR])realloc(spoken, (size_t)(U + 1) * WSTR);
    strncpy(spoken[U++], w, WSTR);
  }
  fclose(fw);
  g_out = fopen(result, "w");
  if (!g_out) die("can't open file %s \n", result);

  /* writing_header R-FS:1014-1031.  The reference hands its double coef_model[] to an int* parameter: entry i
   * printed with "%.2d" is the i-th 32-bit word of the array, reproduced here. */
  int wbits[32];
  memcpy(wbits, weight, sizeof(double) * (size_t)K);
  fprintf(g_out, "Isolated word recognition using Continuous HMM (diagonal covariance matrix). It is considered a final state. \n");
  fprintf(g_out, "Algorithm used for recognition: Forward \n");
  fprintf(g_out, "Number of models: %d  \n", K);
  for (int j = 0; j < K; j++) {
    fprintf(g_out, "Model name %d: %s\n", j + 1, argv[2 + j]);
    fprintf(g_out, "Weighting coefficient of model %d:%.2d\n", j + 1, wbits[j]);
  }
  fprintf(g_out, "Date and time: %s \n\n", date);

  struct tms tb;
  times(&tb);
  double old_aux = tb.tms_utime / 60.0;

  /* the hot path, per model set: all files of its feature list into HBM (read once; the reference re-reads every
   * test file for each model, R-FS:341-369), the whole model list in one go (modelset.c), every (utterance, model)
   * forward score in one batch; then the ranking rule on the weighted sums */
  char (*words)[64] = NULL; /* the words of the LAST set name the vocabulary (R-FS:231 overwrites word[]) */
  int V = 0;
  double *probab = NULL, *logp = NULL;
  int64_t *off = (int64_t *)calloc((size_t)U + 2, sizeof(int64_t)), *offq = (int64_t *)calloc((size_t)U + 2, sizeof(int64_t));
  int32_t *label = (int32_t *)malloc(sizeof(int32_t) * (U > 0 ? U : 1)), *second = (int32_t *)malloc(sizeof(int32_t) * (U > 0 ? U : 1));
  hmmcu_ctx *ctx = NULL;
  trace("start");
  if (U > 0 && hmmcu_create(0, &ctx) != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(NULL));
  trace("context");
  int fl = 2 + 2 * K; /* argv index of the set's first feature list: a set with P streams takes P lists (R-FS:259-266) */
  for (int j = 0; j < K; j++) {
    const char *models_list = argv[2 + j];
    char **mpaths = NULL;
    int Vj = 0, badm = -1, Pj = 0;
    printf("\r\nLoading Models\r\n");
    if (hmmh_read_list(models_list, &mpaths, &Vj) != HMMCU_OK || Vj == 0) die("file %s not found \n", models_list);
    hmmh_model probe[HMMH_MAX_STREAMS];
    memset(probe, 0, sizeof(probe));
    if (hmmh_read_model_streams(mpaths[0], probe, HMMH_MAX_STREAMS, &Pj, 0) != HMMCU_OK) die("file %s not found \n", mpaths[0]);
    if (fl + Pj > argc - 2) die("the argument list is too short for the feature streams of %s \n", models_list);
    if (probe[0].N > HMMCU_MAX_STATES) { /* refused up front: the kernels are built for HMMCU_MAX_STATES states */
      printf("model %s has %d states: beyond this build's limit of %d states per model (HMMCU_MAX_STATES) \n", mpaths[0], probe[0].N, HMMCU_MAX_STATES);
      exit(1);
    }
    if (j == 0) {
      V = Vj;
      probab = (double *)calloc((size_t)(U > 0 ? U : 1) * V, sizeof(double)); /* probab[i] = 0.0, R-FS:284 */
      logp = (double *)malloc(sizeof(double) * (size_t)(U > 0 ? U : 1) * V);
    } else if (Vj != V) {
      die("model list %s has a different number of words \n", models_list);
    }
    words = (char (*)[64])realloc(words, (size_t)V * 64);
    hmmh_model_set models;
    memset(&models, 0, sizeof(models));
    hmmh_model *sm = NULL; /* [Pj][V] when the set has several streams */
    if (Pj == 1) { /* the whole model list in one go, parsed into the layout hmmcu_set_models takes (modelset.c) */
      int mrc = hmmh_read_model_set((const char *const *)mpaths, V, 0, &models, &badm);
      if (mrc == HMMCU_EINVAL) die("model %s has a different topology: not supported \n", badm >= 0 ? mpaths[badm] : models_list);
      if (mrc != HMMCU_OK) die("file %s not found \n", badm >= 0 ? mpaths[badm] : models_list);
      memcpy(words, models.word, (size_t)V * 64);
    } else {
      sm = (hmmh_model *)calloc((size_t)Pj * V, sizeof(hmmh_model));
      for (int v = 0; v < V; v++) {
        hmmh_model tmp[HMMH_MAX_STREAMS];
        int Pv = 0;
        memset(tmp, 0, sizeof(tmp));
        if (hmmh_read_model_streams(mpaths[v], tmp, HMMH_MAX_STREAMS, &Pv, 0) != HMMCU_OK) die("file %s not found \n", mpaths[v]);
        if (Pv != Pj) die("model %s has a different topology: not supported \n", mpaths[v]);
        for (int p = 0; p < Pj; p++) {
          if (tmp[p].N != probe[p].N || tmp[p].M != probe[p].M || tmp[p].D != probe[p].D) die("model %s has a different topology: not supported \n", mpaths[v]);
          sm[(size_t)p * V + v] = tmp[p];
        }
        memcpy(words[v], tmp[0].word, 64);
      }
    }
    hmmh_free_list(mpaths, Vj);
    if (U > 0) {
      hmmcu_ctx *cx[HMMH_MAX_STREAMS];
      cx[0] = ctx;
      for (int p = 0; p < Pj; p++) { /* every stream's feature list into its own context (read once, ingest.c) */
        const char *feat_list = argv[fl + p];
        char **paths = NULL;
        int nf = 0, d = 0, bad = -1;
        if (p > 0 && hmmcu_create(0, &cx[p]) != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(NULL));
        if (hmmh_read_list(feat_list, &paths, &nf) != HMMCU_OK) die("file %s not found \n", feat_list);
        if (nf < U) die("reading error on file %s \n", feat_list);
        int rc = hmmh_ingest(cx[p], (const char *const *)paths, U, 0, p == 0 ? off : offq, &d, &bad, NULL);
        if (rc == HMMCU_EIO) die("file %s not found \n", bad >= 0 ? paths[bad] : feat_list);
        if (rc != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(cx[p]));
        if (d != probe[p].D) die("reading error on file %s \n", paths[0]);
        trace("ingest");
        if (p > 0 && memcmp(off, offq, sizeof(int64_t) * ((size_t)U + 1)) != 0) die("reading error on file %s (the streams of an utterance differ in length) \n", feat_list);
        rc = Pj == 1 ? hmmh_upload_model_set(cx[p], &models) : hmmh_upload_models(cx[p], sm + (size_t)p * V, V);
        if (rc != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(cx[p]));
        hmmh_free_list(paths, nf);
      }
      /* every (utterance, model) forward score in one batch; several streams: the product of their densities */
      if ((Pj > 1 && hmmcu_link_streams(ctx, cx + 1, Pj - 1) != HMMCU_OK) || hmmcu_forward_scores(ctx, logp, 1) != HMMCU_OK)
        die("GPU error: %s \n", hmmcu_last_error(ctx));
      trace("models + scores");
      for (int p = 1; p < Pj; p++) hmmcu_destroy(cx[p]); /* unlinks */
      for (size_t k = 0; k < (size_t)U * V; k++) probab[k] += weight[j] * logp[k];
    }
    for (int p = 0; p < Pj; p++) hmmh_model_free(&probe[p]);
    if (sm) {
      for (size_t k = 0; k < (size_t)Pj * V; k++) hmmh_model_free(&sm[k]);
      free(sm);
    }
    hmmh_model_set_free(&models);
    fl += Pj;
  }
  if (fl != argc - 2) die("models_number %s does not match the argument list \n", argv[1]);
  if (U > 0) {
    if (hmmcu_rank(ctx, probab, U, V, 1.0, label, second) != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(ctx));
    trace("ranking");
    hmmcu_destroy(ctx);
    trace("context released");
  }
  times(&tb);
  const double batch_cpu = tb.tms_utime / 60.0 - old_aux; /* spread evenly over the utterances */

  /* bookkeeping and report, R-FS:283-412 */
  int correct = 0, error = 0, nsecond = 0, sum_correct = 0, sum_error = 0, sum_second = 0, word_frames = 0, total_frames = 0;
  int *wrong = (int *)calloc((size_t)V, sizeof(int));
  double sum_cpu = 0.0;
  char last[WSTR] = " ";
  printf("\r\nStarting Tests\r\n");
  for (int u = 0; u < U; u++) {
    if (strncmp(last, spoken[u], WSTR) != 0) {
      if (strcmp(last, " ") != 0) {
        double cpu_time = batch_cpu * (correct + error) / U;
        sum_cpu += cpu_time;
        write_word_block(correct, error, nsecond, V, last, wrong, words, cpu_time, word_frames);
        sum_correct += correct; sum_error += error; sum_second += nsecond; total_frames += word_frames;
        word_frames = correct = error = nsecond = 0;
        for (int i = 0; i < V; i++) wrong[i] = 0;
      }
      fprintf(g_out, "\nSpoken word: %s\n", spoken[u]);
    }
    word_frames += (int)(off[u + 1] - off[u]);
    printf("\r\nSpoken word: %s -> %s : %f\r\n", spoken[u], words[label[u]], probab[(size_t)u * V + label[u]]);
    if (strncmp(spoken[u], words[label[u]], WSTR) == 0) correct++;
    else {
      error++;
      wrong[label[u]]++;
      if (V > 1 && strncmp(spoken[u], words[second[u]], WSTR) == 0) nsecond++;
    }
    memcpy(last, spoken[u], WSTR);
  }
  printf("\r\nEnding Tests\r\n");
  if (U > 0) {
    double cpu_time = batch_cpu * (correct + error) / U;
    sum_cpu += cpu_time;
    write_word_block(correct, error, nsecond, K /* sic, R-FS:400 */, last, wrong, words, cpu_time, word_frames);
    sum_correct += correct; sum_error += error; sum_second += nsecond; total_frames += word_frames;
    /* writing_total_result R-FS:1170-1194 */
    int sum = sum_correct + sum_error;
    double per = (double)sum_correct / (double)sum;
    fprintf(g_out, "\nConsidering all the words: \n");
    fprintf(g_out, "Results: \n");
    fprintf(g_out, "Correct words: %d\n", sum_correct);
    fprintf(g_out, "Errors: %d\n", sum_error);
    fprintf(g_out, "Percentagen correct : %.2f%%\n", (per * 100.0));
    fprintf(g_out, "Second candidate: %d\n", sum_second);
    fprintf(g_out, "Average recognition time: %.2f sec. \n", sum_cpu / sum);
    fprintf(g_out, "Average word length: %d frames \n", total_frames / sum);
  }
  fclose(g_out);
  free(words); free(offq);
  free(off); free(spoken); free(logp); free(probab); free(label); free(second); free(wrong);
  return 0;
}
