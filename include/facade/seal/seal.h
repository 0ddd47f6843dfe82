// Drop-in for `#include "seal/seal.h"` (M/include.hpp:10): put `include/facade` of this repository
// BEFORE stock SEAL on the include path and the reference's module code binds to the B200 backend.
#pragma once
#include "../../moai_b200_seal.hpp"
