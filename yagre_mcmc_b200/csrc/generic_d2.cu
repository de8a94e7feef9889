// generic_d2.cu -- the one-chain-per-thread kernels (generic_kernel.cuh) for parameter-dimension capacity 2.
#include "generic_kernel.cuh"
YG_GENERIC_UNIT(2)
