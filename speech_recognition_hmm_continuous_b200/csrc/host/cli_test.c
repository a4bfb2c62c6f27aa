/* recognition_continuous_fs: the reference recogniser's program name, GPU scoring behind it. */
#include "hmm_cuda.h"
int main(int argc, char **argv) { return hmmh_test_main(argc, argv); }
