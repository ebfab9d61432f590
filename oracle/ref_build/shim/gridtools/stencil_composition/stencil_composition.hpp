// TEST INFRASTRUCTURE ONLY (oracle/_ref build): empty stand-in, see ../storage/storage_facility.hpp.
// The reference's generated *naive* code needs nothing from GridTools' stencil composition layer.
#pragma once
#include "../storage/storage_facility.hpp"
