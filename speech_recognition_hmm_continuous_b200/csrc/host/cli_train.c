/* hmm_continuous_fs: the reference trainer's program name, GPU E-step behind it. */
#include "hmm_cuda.h"
int main(int argc, char **argv) { return hmmh_train_main(argc, argv); }
