// general-baseline half of the k_moments_em instantiations (see launch_em.cu)
#define CHI_SB 0
#include "launch_em.cu"
