/* recognition_continuous_fs: the reference recogniser's program name, GPU scoring behind it.
 * Ends with _exit after flushing its files, like the trainer (cli_train.c): no orderly CUDA runtime teardown. */
#include <stdio.h>
#include <unistd.h>
#include "hmm_cuda.h"
int main(int argc, char **argv) {
  const int rc = hmmh_test_main(argc, argv);
  fflush(NULL);
  _exit(rc);
}
