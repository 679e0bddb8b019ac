// Shared definitions for the sm_100a GP-SLC kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define GPSLC_OK 0
#define GPSLC_ERR_CUDA 1
#define GPSLC_ERR_ARG 2
#define GPSLC_ERR_NOT_PD 3
#define GPSLC_ERR_UNSUPPORTED 4
#define GPSLC_ERR_NO_DEVICE 5

namespace gpslc {

#ifndef GPSLC_DMAX
#define GPSLC_DMAX 64
#endif
constexpr int DMAX = GPSLC_DMAX;        // max feature dimensions of one covariance factor (nU + nX + 1); beyond CF_DIMS (24) the column
                                // features of a panel are read from global memory instead of shared memory (slower, same values)
constexpr double LOG_2PI = 1.8378770664093454835606594728112;

struct Ctx;  // defined in context.cuh

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace gpslc
