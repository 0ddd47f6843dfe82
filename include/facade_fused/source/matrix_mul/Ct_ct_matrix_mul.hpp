// Shadows M/source/matrix_mul/Ct_ct_matrix_mul.hpp when include/facade_fused precedes the reference on the include path:
// the same functions (names, signatures, results), each one fused device pipeline of libmoai_b200.so.
#pragma once
#include "../../../moai_b200_fused_modules.hpp"
