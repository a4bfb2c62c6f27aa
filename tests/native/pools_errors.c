/* TEST INFRASTRUCTURE: drives the host-side reader pools (ingest.c, modelset.c) under ThreadSanitizer / AddressSanitizer.
 * Built and run by tests/test_host_sanitizers.py; argv[1] = scratch directory. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "hmm_cuda.h"
static double *X; static int Dg;
static int b(void *u, const int64_t *off, int U, int D){ Dg=D; X=calloc((size_t)off[U]*D,sizeof(double)); return 0; }
static int a(void *u, const double *x, int64_t f0, int64_t n, int *t){ memcpy(X+f0*Dg,x,sizeof(double)*n*Dg); *t=0; return 0; }
static int afail(void *u, const double *x, int64_t f0, int64_t n, int *t){ return f0 > 3000 ? 3 : a(u,x,f0,n,t); }
static int w(void *u, int t){ return 0; }
static int e(void *u){ return 0; }
int main(int argc, char **argv){
  int U=200; char **paths=malloc(sizeof(char*)*U);
  for(int u=0;u<U;u++){ char p[512]; snprintf(p,sizeof p,"%s/f%03d.bin",argv[1],u); paths[u]=strdup(p);}    
  int64_t *off=malloc(sizeof(int64_t)*(U+1)); int D=0,bad=-1;
  /* a sink that fails in the middle */
  hmmh_sink s1={0,b,afail,w,e,0};
  int rc=hmmh_ingest_to(&s1,(const char*const*)paths,U,8,50,off,&D,&bad,NULL); printf("sink failure: rc=%d\n",rc); free(X); X=NULL;
  /* a file of another width in the middle: truncate/rewrite file 120 with D=9 */
  { double z[9]={0}; hmmh_write_features(paths[120], z, 1, 9); }
  hmmh_sink s2={0,b,a,w,e,0};
  rc=hmmh_ingest_to(&s2,(const char*const*)paths,U,8,50,off,&D,&bad,NULL); printf("wrong width: rc=%d bad=%d\n",rc,bad); free(X); X=NULL;
  unlink(paths[120]);
  rc=hmmh_ingest_to(&s2,(const char*const*)paths,U,8,50,off,&D,&bad,NULL); printf("missing: rc=%d bad=%d\n",rc,bad);
  for(int u=0;u<U;u++) free(paths[u]); free(paths); free(off); free(X);
  return 0; }
