// generic_d8.cu -- the one-chain-per-thread kernels (generic_kernel.cuh) for parameter-dimension capacity 8.
#ifndef YG_DEV_22       /* the dev build (make dev) covers d <= 2 only */
#include "generic_kernel.cuh"
YG_GENERIC_UNIT(8)
#endif
