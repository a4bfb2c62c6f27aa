/* TEST INFRASTRUCTURE: drives the host-side reader pools (ingest.c, modelset.c) under ThreadSanitizer / AddressSanitizer.
 * Built and run by tests/test_host_sanitizers.py; argv[1] = scratch directory. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hmm_cuda.h"
static double *X; static int Dg; static int64_t got;
static int b(void *u, const int64_t *off, int U, int D){ Dg=D; X=calloc((size_t)off[U]*D,sizeof(double)); got=0; return 0; }
static int a(void *u, const double *x, int64_t f0, int64_t n, int *t){ memcpy(X+f0*Dg,x,sizeof(double)*n*Dg); got+=n; *t=0; return 0; }
static int w(void *u, int t){ return 0; }
static int e(void *u){ return 0; }
int main(int argc, char **argv){
  int U=200, D=7; char **paths=malloc(sizeof(char*)*U); double sum=0;
  for(int u=0;u<U;u++){ char p[512]; snprintf(p,sizeof p,"%s/f%03d.bin",argv[1],u); paths[u]=strdup(p); int T=1+(u*37)%300; double *x=malloc(sizeof(double)*T*D); for(int k=0;k<T*D;k++){x[k]=u+k*1e-3; sum+=x[k];} hmmh_write_features(p,x,T,D); free(x);}    
  hmmh_sink s={0,b,a,w,e,0};
  int64_t *off=malloc(sizeof(int64_t)*(U+1)); int Dd=0,bad=-1; hmmh_ingest_stats st;
  for(int rep=0;rep<3;rep++){
    int rc=hmmh_ingest_to(&s,(const char*const*)paths,U,rep==0?1:8,rep==2?50:0,off,&Dd,&bad,&st);
    double s2=0; for(int64_t k=0;k<off[U]*D;k++) s2+=X[k];
    printf("rc=%d D=%d frames=%lld got=%lld batches=%d sum_ok=%d\n",rc,Dd,(long long)off[U],(long long)got,st.batches,s2==sum);
    free(X);
  }
  /* model set pool */
  hmmh_model_set ms; hmmh_model_set_alloc(&ms,40,5,3,9);
  for(int v=0;v<40;v++){ snprintf(ms.word[v],64,"w%d",v); for(int k=0;k<5*3*9;k++) ms.mu[(size_t)v*135+k]=v+k; }
  char **mp=malloc(sizeof(char*)*40); for(int v=0;v<40;v++){ char p[512]; snprintf(p,sizeof p,"%s/m%02d.hmm",argv[1],v); mp[v]=strdup(p);}    
  int rc=hmmh_write_model_set((const char*const*)mp,&ms,8,&bad); hmmh_model_set back; int rc2=hmmh_read_model_set((const char*const*)mp,40,8,&back,&bad);
  printf("write rc=%d read rc=%d eq=%d word=%s\n",rc,rc2,memcmp(back.mu,ms.mu,sizeof(double)*40*135)==0,back.word[39]);
  hmmh_model_set_free(&ms); hmmh_model_set_free(&back);
  return 0; }
