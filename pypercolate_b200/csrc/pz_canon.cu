// pz_canon.cu -- placeholder (binomial pmf / contraction kernels land next)
#include "pz_common.cuh"
#include "pz_internal.h"
