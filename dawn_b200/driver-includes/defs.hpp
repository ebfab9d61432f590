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

#if DAWN_PRECISION == DAWN_SINGLE_PRECISION
#error "the b200 backend currently emits double precision only (b200_field_t carries doubles)"
#endif

#ifndef GRIDTOOLS_DAWN_HALO_EXTENT
#define GRIDTOOLS_DAWN_HALO_EXTENT 3
#endif
