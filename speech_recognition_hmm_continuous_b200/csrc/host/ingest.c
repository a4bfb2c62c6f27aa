/*
 * ingest.c -- many-files feature reader.  The reference opens every feature file twice per EM
 * iteration and issues one fread per frame (reading_coef, T-FS:527-548; the list walk at T-FS:272-321); the
 * recogniser re-reads each test file once per model (R-FS:341-369).  Here the whole list is read ONCE:
 *   scan    -- fstat of every file, in parallel: frame counts from the file sizes (T = (bytes - 4) / (8 D), a
 *              trailing partial frame dropped exactly as hmmh_read_features does), prefix sums = frame offsets;
 *   read    -- a pool of threads, one preadv per file, straight into its slot of a staging buffer (three
 *              buffers of consecutive utterances; pinned when the sink is the device);
 *   hand-off-- the calling thread passes every full staging buffer to the sink (hmmcu_features_append: an
 *              asynchronous host-to-device copy) while the pool fills the next two.
 * No fp64 copy of the corpus is kept on the host; at C3 scale (9.4 GB) it only ever exists in HBM.
 * The sink is an interface so that the CPU tests can drive the same pipeline into plain memory.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <time.h>
#include <unistd.h>

#include "hmm_cuda.h"

#define NSTAGE 3

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* whitespace-separated tokens of at most 99 characters, as fscanf("%s") into char[100] reads them */
int hmmh_read_list(const char *list_path, char ***paths_out, int *n_out) {
  FILE *f = fopen(list_path, "r");
  if (!f) return HMMCU_EIO;
  char **paths = NULL;
  int n = 0, cap = 0;
  char tok[100];
  while (fscanf(f, "%99s", tok) == 1) {
    if (n == cap) {
      cap = cap ? cap * 2 : 256;
      char **p = (char **)realloc(paths, sizeof(char *) * (size_t)cap);
      if (!p) { fclose(f); hmmh_free_list(paths, n); return HMMCU_ENOMEM; }
      paths = p;
    }
    paths[n] = strdup(tok);
    if (!paths[n]) { fclose(f); hmmh_free_list(paths, n); return HMMCU_ENOMEM; }
    n++;
  }
  fclose(f);
  *paths_out = paths;
  *n_out = n;
  return HMMCU_OK;
}

void hmmh_free_list(char **paths, int n) {
  if (!paths) return;
  for (int i = 0; i < n; i++) free(paths[i]);
  free(paths);
}

/* ------------------------------------------------------------------------ scan ------------- */
typedef struct {
  const char *const *paths;
  int U, D;
  int64_t *bytes; /* payload bytes of every file (header excluded), -1 = error */
  int next;
  int bad; /* index of the first file that failed, -1 = none */
  pthread_mutex_t mu;
} scan_job;

static void *scan_worker(void *arg) {
  scan_job *j = (scan_job *)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    const int u = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (u >= j->U) return NULL;
    struct stat sb;
    if (stat(j->paths[u], &sb) != 0 || sb.st_size < (off_t)sizeof(int)) {
      pthread_mutex_lock(&j->mu);
      if (j->bad < 0 || u < j->bad) j->bad = u;
      pthread_mutex_unlock(&j->mu);
      j->bytes[u] = -1;
    } else {
      j->bytes[u] = (int64_t)sb.st_size - (int64_t)sizeof(int);
    }
  }
}

static int clamp_threads(int nthreads, int U) {
  if (nthreads <= 0) {
    const char *e = getenv("HMMCU_INGEST_THREADS");
    nthreads = e ? atoi(e) : 0;
    if (nthreads <= 0) {
      long n = sysconf(_SC_NPROCESSORS_ONLN);
      nthreads = n > 16 ? 16 : (n < 1 ? 1 : (int)n);
    }
  }
  if (nthreads > 64) nthreads = 64;
  if (nthreads > U) nthreads = U > 0 ? U : 1;
  return nthreads;
}

int hmmh_scan_features(const char *const *paths, int U, int nthreads, int *D_out, int64_t *frame_off, int *bad_file) {
  if (!paths || U < 1 || !D_out || !frame_off) return HMMCU_EINVAL;
  if (bad_file) *bad_file = -1;
  /* D from the header of the first file; every other header is checked when its file is read */
  int D = 0;
  {
    int fd = open(paths[0], O_RDONLY);
    if (fd < 0 || pread(fd, &D, sizeof(int), 0) != (ssize_t)sizeof(int) || D < 1 || D > 4096) {
      if (fd >= 0) close(fd);
      if (bad_file) *bad_file = 0;
      return HMMCU_EIO;
    }
    close(fd);
  }
  scan_job j;
  memset(&j, 0, sizeof(j));
  j.paths = paths; j.U = U; j.D = D; j.bad = -1;
  j.bytes = (int64_t *)malloc(sizeof(int64_t) * (size_t)U);
  if (!j.bytes) return HMMCU_ENOMEM;
  pthread_mutex_init(&j.mu, NULL);
  nthreads = clamp_threads(nthreads, U);
  pthread_t th[64];
  int started = 0;
  for (int k = 1; k < nthreads; k++)
    if (pthread_create(&th[started], NULL, scan_worker, &j) == 0) started++;
  scan_worker(&j);
  for (int k = 0; k < started; k++) pthread_join(th[k], NULL);
  pthread_mutex_destroy(&j.mu);
  int rc = HMMCU_OK;
  if (j.bad >= 0) {
    if (bad_file) *bad_file = j.bad;
    rc = HMMCU_EIO;
  } else {
    frame_off[0] = 0;
    for (int u = 0; u < U; u++) frame_off[u + 1] = frame_off[u] + j.bytes[u] / ((int64_t)sizeof(double) * D);
    *D_out = D;
  }
  free(j.bytes);
  return rc;
}

/* ------------------------------------------------------------------------ read + hand-off -- */
typedef struct {
  const char *const *paths;
  const int64_t *off;
  int U, D;
  const int *batch_of;      /* utterance -> batch */
  const int64_t *batch_f0;  /* first frame of every batch */
  double *stage[NSTAGE];
  int *remaining; /* files still to read, per batch */
  int next;       /* next utterance to read */
  int freed;      /* batches [0, freed) have left their staging buffers */
  int err, bad;
  pthread_mutex_t mu;
  pthread_cond_t cv;
} read_job;

static int read_file_into(const char *path, int D, double *dst, int64_t frames) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) return HMMCU_EIO;
  int d = 0;
  size_t want = sizeof(double) * (size_t)frames * D, got = 0;
  struct iovec iv[2];
  iv[0].iov_base = &d; iv[0].iov_len = sizeof(int);
  iv[1].iov_base = dst; iv[1].iov_len = want;
  ssize_t r = preadv(fd, iv, 2, 0);
  if (r < (ssize_t)sizeof(int) || d != D) { close(fd); return HMMCU_EIO; }
  got = (size_t)r - sizeof(int);
  while (got < want) { /* short reads */
    r = pread(fd, (char *)dst + got, want - got, (off_t)(sizeof(int) + got));
    if (r <= 0) { close(fd); return HMMCU_EIO; }
    got += (size_t)r;
  }
  close(fd);
  return HMMCU_OK;
}

static void *read_worker(void *arg) {
  read_job *j = (read_job *)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    const int u = j->next < j->U ? j->next++ : j->U;
    int b = 0;
    if (u < j->U) {
      b = j->batch_of[u];
      while (!j->err && b >= j->freed + NSTAGE) pthread_cond_wait(&j->cv, &j->mu);
    }
    const int stop = j->err || u >= j->U;
    pthread_mutex_unlock(&j->mu);
    if (stop) return NULL;
    const int64_t frames = j->off[u + 1] - j->off[u];
    int rc = read_file_into(j->paths[u], j->D, j->stage[b % NSTAGE] + (j->off[u] - j->batch_f0[b]) * j->D, frames);
    pthread_mutex_lock(&j->mu);
    if (rc) {
      if (!j->err || u < j->bad) j->bad = u;
      j->err = rc;
    }
    j->remaining[b]--;
    pthread_cond_broadcast(&j->cv);
    pthread_mutex_unlock(&j->mu);
  }
}

int hmmh_ingest_to(const hmmh_sink *sink, const char *const *paths, int U, int nthreads, int64_t stage_frames, int64_t *frame_off,
                   int *D_out, int *bad_file, hmmh_ingest_stats *stats) {
  if (!sink || !sink->begin || !sink->append || !sink->wait || !sink->end || !paths || U < 1 || !frame_off || !D_out) return HMMCU_EINVAL;
  const double t0 = now_s();
  int D = 0;
  int rc = hmmh_scan_features(paths, U, nthreads, &D, frame_off, bad_file);
  if (rc) return rc;
  for (int u = 0; u < U; u++)
    if (frame_off[u + 1] == frame_off[u]) { /* an empty file: the reference would divide by a zero length */
      if (bad_file) *bad_file = u;
      return HMMCU_EIO;
    }
  const double t1 = now_s();
  *D_out = D;
  const int64_t F = frame_off[U];
  /* batches of consecutive utterances, each at most stage_frames frames (default 8 MiB of doubles) */
  if (stage_frames <= 0) stage_frames = ((int64_t)8 << 20) / ((int64_t)sizeof(double) * D);
  for (int u = 0; u < U; u++)
    if (frame_off[u + 1] - frame_off[u] > stage_frames) stage_frames = frame_off[u + 1] - frame_off[u];
  if (stage_frames > F) stage_frames = F;
  int *batch_of = (int *)malloc(sizeof(int) * (size_t)U);
  int64_t *batch_f0 = (int64_t *)malloc(sizeof(int64_t) * ((size_t)U + 1));
  int *remaining = (int *)calloc((size_t)U, sizeof(int));
  if (!batch_of || !batch_f0 || !remaining) { free(batch_of); free(batch_f0); free(remaining); return HMMCU_ENOMEM; }
  int B = 0;
  batch_f0[0] = 0;
  for (int u = 0; u < U; u++) {
    if (frame_off[u + 1] - batch_f0[B] > stage_frames) batch_f0[++B] = frame_off[u];
    batch_of[u] = B;
    remaining[B]++;
  }
  B++;
  batch_f0[B] = F;

  read_job j;
  memset(&j, 0, sizeof(j));
  j.paths = paths; j.off = frame_off; j.U = U; j.D = D; j.batch_of = batch_of; j.batch_f0 = batch_f0; j.remaining = remaining;
  j.bad = -1;
  const size_t stage_bytes = sizeof(double) * (size_t)stage_frames * D;
  int nstage = 0;
  for (; nstage < NSTAGE; nstage++) {
    j.stage[nstage] = (double *)(sink->stage_alloc ? sink->stage_alloc(sink->user, nstage, stage_bytes) : malloc(stage_bytes));
    if (!j.stage[nstage]) break;
  }
  const double t1b = now_s();
  int tickets[NSTAGE] = {-1, -1, -1};
  int started = 0, begun = 0;
  pthread_t th[64];
  if (nstage < NSTAGE) { rc = HMMCU_ENOMEM; goto done; }
  rc = sink->begin(sink->user, frame_off, U, D);
  if (rc) goto done;
  begun = 1;
  pthread_mutex_init(&j.mu, NULL);
  pthread_cond_init(&j.cv, NULL);
  nthreads = clamp_threads(nthreads, U);
  for (int k = 0; k < nthreads; k++)
    if (pthread_create(&th[started], NULL, read_worker, &j) == 0) started++;
  if (started == 0) { rc = HMMCU_ENOMEM; j.err = rc; }
  for (int b = 0; b < B && !rc; b++) {
    pthread_mutex_lock(&j.mu);
    while (!j.err && j.remaining[b] > 0) pthread_cond_wait(&j.cv, &j.mu);
    rc = j.err;
    pthread_mutex_unlock(&j.mu);
    if (rc) break;
    rc = sink->append(sink->user, j.stage[b % NSTAGE], batch_f0[b], batch_f0[b + 1] - batch_f0[b], &tickets[b % NSTAGE]);
    if (!rc && b >= 1) rc = sink->wait(sink->user, tickets[(b - 1) % NSTAGE]);
    pthread_mutex_lock(&j.mu);
    if (rc) j.err = rc;
    j.freed = b; /* batches before b have been copied out of their buffers */
    pthread_cond_broadcast(&j.cv);
    pthread_mutex_unlock(&j.mu);
  }
  if (!rc) rc = sink->wait(sink->user, tickets[(B - 1) % NSTAGE]);
  pthread_mutex_lock(&j.mu);
  if (rc && !j.err) j.err = rc;
  pthread_cond_broadcast(&j.cv);
  pthread_mutex_unlock(&j.mu);
  for (int k = 0; k < started; k++) pthread_join(th[k], NULL);
  pthread_mutex_destroy(&j.mu);
  pthread_cond_destroy(&j.cv);
  if (rc && bad_file) *bad_file = j.bad;
  {
    const double t2 = now_s();
    if (!rc) rc = sink->end(sink->user);
    if (stats) {
      stats->scan_s = t1 - t0;
      stats->stage_s = t1b - t1;
      stats->read_s = t2 - t1b;
      stats->total_s = now_s() - t0;
      stats->bytes = (int64_t)sizeof(double) * F * D;
      stats->batches = B;
      stats->threads = nthreads;
    }
  }
done:
  (void)begun;
  for (int k = 0; k < nstage; k++) {
    if (!sink->stage_alloc) free(j.stage[k]); /* a sink's own buffers stay the sink's */
  }
  free(batch_of); free(batch_f0); free(remaining);
  return rc;
}

/* ------------------------------------------------------------------------ the device sink -- */
static int dev_begin(void *user, const int64_t *off, int U, int D) { return hmmcu_features_begin((hmmcu_ctx *)user, off, U, D); }
static int dev_append(void *user, const double *x, int64_t f0, int64_t n, int *ticket) {
  return hmmcu_features_append((hmmcu_ctx *)user, x, f0, n, ticket);
}
static int dev_wait(void *user, int ticket) { return hmmcu_features_wait((hmmcu_ctx *)user, ticket); }
static int dev_end(void *user) { return hmmcu_features_end((hmmcu_ctx *)user); }
static void *dev_alloc(void *user, int slot, size_t bytes) { return hmmcu_staging((hmmcu_ctx *)user, slot, bytes); }

int hmmh_ingest(hmmcu_ctx *ctx, const char *const *paths, int U, int nthreads, int64_t *frame_off, int *D_out, int *bad_file,
                hmmh_ingest_stats *stats) {
  if (!ctx) return HMMCU_EINVAL;
  hmmh_sink s = {ctx, dev_begin, dev_append, dev_wait, dev_end, dev_alloc};
  return hmmh_ingest_to(&s, paths, U, nthreads, 0, frame_off, D_out, bad_file, stats);
}
