// placeholder until the tcgen05 kernel lands: every shape reports "not supported"
#include "common.cuh"
int turtle_gemm_tc(const TurtleGemmArgs *, void *) { return TURTLE_ENOTSUP; }
