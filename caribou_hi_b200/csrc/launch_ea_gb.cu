// general-baseline half of the fused-kernel instantiations (see launch_ea.cu)
#define CHI_EA_SB 0
#include "launch_ea.cu"
