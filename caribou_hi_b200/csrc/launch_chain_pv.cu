#define CHI_CHAIN_PV 1
#include "launch_chain.cu"
