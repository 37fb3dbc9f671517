// pseudo-Voigt half of launch_em_tame.cu
#define CHI_EM_TAME_PV 1
#include "launch_em_tame.cu"
