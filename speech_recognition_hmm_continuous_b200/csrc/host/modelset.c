/*
 * modelset.c -- bulk .hmm I/O: 1000-2000 model files <-> the struct-of-arrays layout that
 * hmmcu_set_models takes.  The reference loads its vocabulary one model at a time with one fread per field
 * (reading_model, R-FS:612-712, called from the list walk at R-FS:214-238) into malloc'd `struct model` nodes;
 * here every file is read with ONE read into a buffer by a pool of threads and parsed straight into its slice
 * of A[V][N][N], c[V][N][M], mu[V][N][M][D], inv_var[V][N][M][D], det[V][N][M].  Same bytes in and out as
 * hmmh_read_model / hmmh_write_model (writer T-FS:2058-2144), including the 4-byte length header of the files
 * shipped with the reference.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "hmm_cuda.h"

int hmmh_model_set_alloc(hmmh_model_set *s, int V, int N, int M, int D) {
  if (!s || V < 1 || N < 1 || M < 1 || D < 1) return HMMCU_EINVAL;
  const size_t g = (size_t)V * N * M;
  memset(s, 0, sizeof(*s));
  s->V = V; s->N = N; s->M = M; s->D = D;
  s->word = (char (*)[64])calloc((size_t)V, 64);
  s->A = (double *)calloc((size_t)V * N * N, sizeof(double));
  s->c = (double *)calloc(g, sizeof(double));
  s->mu = (double *)calloc(g * D, sizeof(double));
  s->inv_var = (double *)calloc(g * D, sizeof(double));
  s->det = (double *)calloc(g, sizeof(double));
  if (!s->word || !s->A || !s->c || !s->mu || !s->inv_var || !s->det) { hmmh_model_set_free(s); return HMMCU_ENOMEM; }
  return HMMCU_OK;
}

void hmmh_model_set_free(hmmh_model_set *s) {
  if (!s) return;
  free(s->word); free(s->A); free(s->c); free(s->mu); free(s->inv_var); free(s->det);
  memset(s, 0, sizeof(*s));
}

/* whole file into a malloc'd buffer */
static int slurp(const char *path, unsigned char **buf, size_t *len) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) return HMMCU_EIO;
  struct stat sb;
  if (fstat(fd, &sb) != 0 || sb.st_size < 24) { close(fd); return HMMCU_EIO; }
  unsigned char *b = (unsigned char *)malloc((size_t)sb.st_size);
  if (!b) { close(fd); return HMMCU_ENOMEM; }
  size_t got = 0;
  while (got < (size_t)sb.st_size) {
    ssize_t r = pread(fd, b + got, (size_t)sb.st_size - got, (off_t)got);
    if (r <= 0) { free(b); close(fd); return HMMCU_EIO; }
    got += (size_t)r;
  }
  close(fd);
  *buf = b; *len = got;
  return HMMCU_OK;
}

/* header: returns the offset of A, or 0 on a malformed file */
static size_t parse_header(const unsigned char *b, size_t len, char word[64], int *N, int *M, int *D) {
  unsigned long long len8 = 0;
  unsigned int len4 = 0;
  memcpy(&len8, b, 8); memcpy(&len4, b, 4);
  const int hb = (len8 < 64) ? 8 : 4; /* a 4-byte header followed by text never reads as a small u64 */
  const size_t wl = hb == 8 ? (size_t)len8 : (size_t)len4;
  if (wl >= 64 || hb + wl + 16 > len) return 0;
  memset(word, 0, 64);
  memcpy(word, b + hb, wl);
  int P = 0;
  size_t o = hb + wl;
  memcpy(N, b + o, 4); memcpy(&P, b + o + 4, 4);
  if (P != 1 || *N < 1 || *N > HMMH_MAX_FILE_STATES) return 0;
  memcpy(M, b + o + 8, 4); memcpy(D, b + o + 12, 4);
  if (*M < 1 || *D < 1 || *M > 4096 || *D > 4096) return 0;
  return o + 16;
}

static size_t body_bytes(int N, int M, int D) { return sizeof(double) * ((size_t)N * N + (size_t)N * M * (2 * (size_t)D + 2)); }

static void parse_body(const unsigned char *p, hmmh_model_set *s, int v) {
  const int N = s->N, M = s->M, D = s->D;
  memcpy(s->A + (size_t)v * N * N, p, sizeof(double) * N * N);
  p += sizeof(double) * N * N;
  for (int i = 0; i < N; i++) {
    const size_t g0 = ((size_t)v * N + i) * M;
    memcpy(s->c + g0, p, sizeof(double) * M);
    p += sizeof(double) * M;
    for (int j = 0; j < M; j++) {
      memcpy(s->mu + (g0 + j) * D, p, sizeof(double) * D);
      p += sizeof(double) * D;
      memcpy(s->det + g0 + j, p, sizeof(double));
      p += sizeof(double);
      memcpy(s->inv_var + (g0 + j) * D, p, sizeof(double) * D);
      p += sizeof(double) * D;
    }
  }
}

typedef struct {
  const char *const *paths;
  hmmh_model_set *s;
  int next, err, bad, write;
  pthread_mutex_t mu;
} set_job;

static int read_one(const char *path, hmmh_model_set *s, int v) {
  unsigned char *b = NULL;
  size_t len = 0;
  int rc = slurp(path, &b, &len);
  if (rc) return rc;
  int N, M, D;
  const size_t o = parse_header(b, len, s->word[v], &N, &M, &D);
  if (!o) rc = HMMCU_EIO;
  else if (N != s->N || M != s->M || D != s->D) rc = HMMCU_EINVAL; /* one topology per set */
  else if (len < o + body_bytes(N, M, D)) rc = HMMCU_EIO;
  else parse_body(b + o, s, v);
  free(b);
  return rc;
}

static int write_one(const char *path, const hmmh_model_set *s, int v) {
  const int N = s->N, M = s->M, D = s->D;
  const size_t wl = strnlen(s->word[v], 63);
  const size_t len = sizeof(size_t) + wl + 16 + body_bytes(N, M, D);
  unsigned char *b = (unsigned char *)malloc(len), *p = b;
  if (!b) return HMMCU_ENOMEM;
  const int P = 1;
  memcpy(p, &wl, sizeof(size_t)); p += sizeof(size_t);
  memcpy(p, s->word[v], wl); p += wl;
  memcpy(p, &N, 4); memcpy(p + 4, &P, 4); memcpy(p + 8, &M, 4); memcpy(p + 12, &D, 4); p += 16;
  memcpy(p, s->A + (size_t)v * N * N, sizeof(double) * N * N); p += sizeof(double) * N * N;
  for (int i = 0; i < N; i++) {
    const size_t g0 = ((size_t)v * N + i) * M;
    memcpy(p, s->c + g0, sizeof(double) * M); p += sizeof(double) * M;
    for (int j = 0; j < M; j++) {
      memcpy(p, s->mu + (g0 + j) * D, sizeof(double) * D); p += sizeof(double) * D;
      memcpy(p, s->det + g0 + j, sizeof(double)); p += sizeof(double);
      memcpy(p, s->inv_var + (g0 + j) * D, sizeof(double) * D); p += sizeof(double) * D;
    }
  }
  int rc = HMMCU_OK;
  int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) rc = HMMCU_EIO;
  else {
    size_t put = 0;
    while (put < len) {
      ssize_t r = write(fd, b + put, len - put);
      if (r <= 0) { rc = HMMCU_EIO; break; }
      put += (size_t)r;
    }
    if (close(fd) != 0) rc = HMMCU_EIO;
  }
  free(b);
  return rc;
}

static void *set_worker(void *arg) {
  set_job *j = (set_job *)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    const int v = j->next < j->s->V && !j->err ? j->next++ : j->s->V;
    pthread_mutex_unlock(&j->mu);
    if (v >= j->s->V) return NULL;
    const int rc = j->write ? write_one(j->paths[v], j->s, v) : read_one(j->paths[v], j->s, v);
    if (rc) {
      pthread_mutex_lock(&j->mu);
      if (!j->err || v < j->bad) { j->bad = v; j->err = rc; }
      pthread_mutex_unlock(&j->mu);
    }
  }
}

static int run_pool(set_job *j, int nthreads, int first) {
  if (nthreads <= 0) {
    const char *e = getenv("HMMCU_INGEST_THREADS");
    nthreads = e ? atoi(e) : 0;
    if (nthreads <= 0) {
      long n = sysconf(_SC_NPROCESSORS_ONLN);
      nthreads = n > 16 ? 16 : (n < 1 ? 1 : (int)n);
    }
  }
  if (nthreads > 64) nthreads = 64;
  if (nthreads > j->s->V - first) nthreads = j->s->V - first;
  j->next = first;
  pthread_mutex_init(&j->mu, NULL);
  pthread_t th[64];
  int started = 0;
  for (int k = 1; k < nthreads; k++)
    if (pthread_create(&th[started], NULL, set_worker, j) == 0) started++;
  set_worker(j);
  for (int k = 0; k < started; k++) pthread_join(th[k], NULL);
  pthread_mutex_destroy(&j->mu);
  return j->err;
}

int hmmh_read_model_set(const char *const *paths, int V, int nthreads, hmmh_model_set *s, int *bad_file) {
  if (!paths || V < 1 || !s) return HMMCU_EINVAL;
  if (bad_file) *bad_file = -1;
  /* the first file fixes the topology */
  unsigned char *b = NULL;
  size_t len = 0;
  int rc = slurp(paths[0], &b, &len);
  if (rc) { if (bad_file) *bad_file = 0; return rc; }
  char word[64];
  int N, M, D;
  const size_t o = parse_header(b, len, word, &N, &M, &D);
  free(b);
  if (!o) { if (bad_file) *bad_file = 0; return HMMCU_EIO; }
  rc = hmmh_model_set_alloc(s, V, N, M, D);
  if (rc) return rc;
  set_job j;
  memset(&j, 0, sizeof(j));
  j.paths = paths; j.s = s; j.bad = -1;
  rc = run_pool(&j, nthreads, 0);
  if (rc) {
    if (bad_file) *bad_file = j.bad;
    hmmh_model_set_free(s);
  }
  return rc;
}

int hmmh_write_model_set(const char *const *paths, const hmmh_model_set *s, int nthreads, int *bad_file) {
  if (!paths || !s || s->V < 1) return HMMCU_EINVAL;
  if (bad_file) *bad_file = -1;
  set_job j;
  memset(&j, 0, sizeof(j));
  j.paths = paths; j.s = (hmmh_model_set *)s; j.bad = -1; j.write = 1;
  const int rc = run_pool(&j, nthreads, 0);
  if (rc && bad_file) *bad_file = j.bad;
  return rc;
}

int hmmh_upload_model_set(hmmcu_ctx *ctx, const hmmh_model_set *s) {
  if (!ctx || !s) return HMMCU_EINVAL;
  return hmmcu_set_models(ctx, s->V, s->N, s->M, s->D, s->A, s->c, s->mu, s->inv_var, s->det);
}
