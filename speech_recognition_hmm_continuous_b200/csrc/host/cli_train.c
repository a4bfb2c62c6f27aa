/* hmm_continuous_fs: the reference trainer's program name, GPU E-step behind it.
 * The process ends with _exit after flushing its files: the orderly unloading of the CUDA runtime (destruction of the
 * primary context by the exit handlers) costs 0.3-0.9 s here, several times the work of a small job; the driver
 * reclaims a dead process's resources without it. */
#include <stdio.h>
#include <unistd.h>
#include "hmm_cuda.h"
int main(int argc, char **argv) {
  const int rc = hmmh_train_main(argc, argv);
  fflush(NULL);
  _exit(rc);
}
