// Floating-point type selection; same switches as the reference (driver-includes/defs.hpp:40-84).
#pragma once

#define DAWN_SINGLE_PRECISION 0
#define DAWN_DOUBLE_PRECISION 1

#ifndef DAWN_PRECISION
#if defined(GT_FLOAT_PRECISION) && GT_FLOAT_PRECISION == 4
#define DAWN_PRECISION DAWN_SINGLE_PRECISION
#else
#define DAWN_PRECISION DAWN_DOUBLE_PRECISION
#endif
#endif

namespace dawn {
#if DAWN_PRECISION == DAWN_SINGLE_PRECISION
using float_type = float;
#else
using float_type = double;
#endif
} // namespace dawn

// pair of adjacent values: what generated tile kernels read from shared memory in one LDS.128 / LDS.64
#if defined(__CUDACC__) || defined(B200_EMU_VECTOR_TYPES)
#if defined(__CUDACC__)
#include <vector_types.h>
#include <vector_functions.h>
#endif
namespace dawn {
#if DAWN_PRECISION == DAWN_SINGLE_PRECISION
using float_type2 = float2;
__host__ __device__ inline float2 make_float_type2(float x, float y) { return make_float2(x, y); }
#else
using float_type2 = double2;
__host__ __device__ inline double2 make_float_type2(double x, double y) { return make_double2(x, y); }
#endif
} // namespace dawn
#endif

#ifndef GRIDTOOLS_DAWN_HALO_EXTENT
#define GRIDTOOLS_DAWN_HALO_EXTENT 3
#endif
