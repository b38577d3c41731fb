// persist_f32.cu -- fp32 instantiations of the device-resident solve (see persist_f64.cu).
#include "internal.cuh"
#include "two_loop_gram.cuh"
#include "persist.cuh"
#define LBFGS_B200_PERSIST_F32 1
#include "persist_host.cuh"
