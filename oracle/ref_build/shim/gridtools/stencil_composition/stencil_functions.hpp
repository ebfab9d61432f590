// TEST INFRASTRUCTURE ONLY (oracle/_ref build): empty stand-in, see ../storage/storage_facility.hpp.
#pragma once
#include "../storage/storage_facility.hpp"
