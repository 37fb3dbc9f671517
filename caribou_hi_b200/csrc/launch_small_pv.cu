// pseudo-Voigt half of the k_eval_small instantiations (see launch_small.cu)
#define CHI_SMALL_PV 1
#include "launch_small.cu"
