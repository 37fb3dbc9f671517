// general-baseline half of the k_moments_abs instantiations (see launch_abs.cu)
#define CHI_SB 0
#include "launch_abs.cu"
