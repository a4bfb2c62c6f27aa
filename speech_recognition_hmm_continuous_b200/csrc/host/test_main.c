/*
 * test_main.c -- drop-in for the reference recogniser's main() (R-FS:87-428):
 *   recognition_continuous_fs K modelslist_1..K weight_1..K featlist_1.. wordsfile resultfile
 * Same argv, list / feature / .hmm formats and result-file text.  All (utterance, model) forward
 * scores are computed on the GPU in one batch; ranking follows sorting_probab including its NaN
 * behaviour.  Deliberate limits: K (model sets) is 1, as MAX_MODELS_NUMBER is (R-FS:40); all models
 * must share one topology; one feature stream.
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

static void die(const char *fmt, const char *arg) {
  printf(fmt, arg);
  exit(1);
}

static FILE *g_out;

/* writing_result_word R-FS:1110-1150 */
static void write_word_block(int correct, int error, int second, int nwords, const char *spoken, const int *wrong,
                             hmmh_model *models, double cpu_time, int frames) {
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
      if (wrong[i] != 0) fprintf(g_out, "%s: %d time%s\n", models[i].word, wrong[i], wrong[i] == 1 ? "" : "s");
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
  const int K = atoi(argv[1]);
  if (K != 1) die("models_number %s is not supported: one model set per word \n", argv[1]);
  const double weight = atof(argv[3]);
  const char *models_list = argv[2], *feat_list = argv[4], *words_file = argv[argc - 2], *result = argv[argc - 1];

  /* model set */
  FILE *fm = fopen(models_list, "rb");
  if (!fm) die("file %s not found \n", models_list);
  hmmh_model *models = NULL;
  int V = 0;
  char name[STR];
  printf("\r\nLoading Models\r\n");
  while (fscanf(fm, "%99s", name) == 1) {
    models = (hmmh_model *)realloc(models, sizeof(hmmh_model) * (V + 1));
    memset(&models[V], 0, sizeof(hmmh_model));
    if (hmmh_read_model(name, &models[V], 0) != HMMCU_OK) die("file %s not found \n", name);
    if (V > 0 && (models[V].N != models[0].N || models[V].M != models[0].M || models[V].D != models[0].D))
      die("model %s has a different topology: not supported \n", name);
    V++;
  }
  fclose(fm);
  if (V == 0) die("file %s not found \n", models_list);

  /* test utterances: line i of the feature list pairs with line i of the words file */
  FILE *ff = fopen(feat_list, "r"), *fw = fopen(words_file, "r");
  if (!ff) die("file %s not found \n", feat_list);
  if (!fw) die("file %s not found \n", words_file);
  g_out = fopen(result, "w");
  if (!g_out) die("can't open file %s \n", result);
  char (*spoken)[WSTR] = NULL;
  char **paths = NULL;
  int U = 0, D = models[0].D;
  char w[WSTR], path[STR];
  while (fscanf(fw, "%49s", w) == 1) {
    if (fscanf(ff, "%99s", path) != 1) die("reading error on file %s \n", feat_list);
    spoken = (char (*)[WSTR])realloc(spoken, (size_t)(U + 1) * WSTR);
    strncpy(spoken[U], w, WSTR);
    paths = (char **)realloc(paths, sizeof(char *) * (size_t)(U + 1));
    paths[U++] = strdup(path);
  }
  fclose(ff); fclose(fw);
  int64_t *off = (int64_t *)calloc((size_t)U + 2, sizeof(int64_t));

  /* writing_header R-FS:1014-1031 (weight printed through an int*, as the reference does) */
  int wbits;
  memcpy(&wbits, &weight, sizeof(int));
  fprintf(g_out, "Isolated word recognition using Continuous HMM (diagonal covariance matrix). It is considered a final state. \n");
  fprintf(g_out, "Algorithm used for recognition: Forward \n");
  fprintf(g_out, "Number of models: %d  \n", K);
  fprintf(g_out, "Model name %d: %s\n", 1, models_list);
  fprintf(g_out, "Weighting coefficient of model %d:%.2d\n", 1, wbits);
  fprintf(g_out, "Date and time: %s \n\n", date);

  struct tms tb;
  times(&tb);
  double old_aux = tb.tms_utime / 60.0;

  /* the hot path: every (utterance, model) forward score, then the ranking rule */
  double *logp = (double *)malloc(sizeof(double) * (size_t)(U > 0 ? U : 1) * V);
  int32_t *label = (int32_t *)malloc(sizeof(int32_t) * (U > 0 ? U : 1)), *second = (int32_t *)malloc(sizeof(int32_t) * (U > 0 ? U : 1));
  if (U > 0) {
    hmmcu_ctx *ctx = NULL;
    if (hmmcu_create(0, &ctx) != HMMCU_OK) die("GPU error: %s \n", hmmcu_last_error(NULL));
    /* every test file is read once (the reference re-reads it for each model, R-FS:341-369), by the many-files reader */
    int d = 0, bad = -1;
    int rc = hmmh_ingest(ctx, (const char *const *)paths, U, 0, off, &d, &bad, NULL);
    if (rc == HMMCU_EIO) die("file %s not found \n", bad >= 0 ? paths[bad] : feat_list);
    if (rc == HMMCU_OK && d != D) die("reading error on file %s \n", paths[0]);
    if (rc != HMMCU_OK || hmmh_upload_models(ctx, models, V) != HMMCU_OK ||
        hmmcu_forward_scores(ctx, logp, 1) != HMMCU_OK || hmmcu_rank(ctx, logp, U, V, weight, label, second) != HMMCU_OK)
      die("GPU error: %s \n", hmmcu_last_error(ctx));
    hmmcu_destroy(ctx);
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
        write_word_block(correct, error, nsecond, V, last, wrong, models, cpu_time, word_frames);
        sum_correct += correct; sum_error += error; sum_second += nsecond; total_frames += word_frames;
        word_frames = correct = error = nsecond = 0;
        for (int i = 0; i < V; i++) wrong[i] = 0;
      }
      fprintf(g_out, "\nSpoken word: %s\n", spoken[u]);
    }
    word_frames += (int)(off[u + 1] - off[u]);
    printf("\r\nSpoken word: %s -> %s : %f\r\n", spoken[u], models[label[u]].word, weight * logp[(size_t)u * V + label[u]]);
    if (strncmp(spoken[u], models[label[u]].word, WSTR) == 0) correct++;
    else {
      error++;
      wrong[label[u]]++;
      if (V > 1 && strncmp(spoken[u], models[second[u]].word, WSTR) == 0) nsecond++;
    }
    strncpy(last, spoken[u], WSTR);
  }
  printf("\r\nEnding Tests\r\n");
  if (U > 0) {
    double cpu_time = batch_cpu * (correct + error) / U;
    sum_cpu += cpu_time;
    write_word_block(correct, error, nsecond, K /* sic, R-FS:400 */, last, wrong, models, cpu_time, word_frames);
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
  for (int v = 0; v < V; v++) hmmh_model_free(&models[v]);
  for (int u = 0; u < U; u++) free(paths[u]);
  free(paths);
  free(models); free(off); free(spoken); free(logp); free(label); free(second); free(wrong);
  return 0;
}
