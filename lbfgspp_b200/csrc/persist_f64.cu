// persist_f64.cu -- fp64 instantiations of the device-resident solve (persist.cuh) + its C ABI (persist_host.cuh).
// A translation unit of its own so that the three big kernel families of liblbfgs_b200.so compile in parallel.
#include "internal.cuh"
#include "two_loop_gram.cuh"
#include "persist.cuh"
#define LBFGS_B200_PERSIST_F64 1
#include "persist_host.cuh"
