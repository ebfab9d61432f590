// Umbrella header of the B200 run-time headers.  Same include name as the reference's
// dawn/src/driver-includes/gridtools_includes.hpp so that generated code and user drivers written
// against the reference (`#include <driver-includes/gridtools_includes.hpp>`, `using namespace
// gridtools::dawn;`) compile unchanged -- but nothing here depends on GridTools: storages are
// B200-native (pinned host mirror + HBM buffer behind the C-ABI runtime, include/dawn_b200_rt.h).
#pragma once
#include <cassert>
#include <exception>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "defs.hpp"
#include "domain.hpp"
#include "extent.hpp"
#include "math.hpp"
#include "b200_runtime.hpp"
#include "storage.hpp"
#include "timer_b200.hpp"
#include "b200_tma.hpp"
#include <cstdint>
