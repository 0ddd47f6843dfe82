// MOAI module functions over the batched Evaluator (see modules.cu).
#pragma once
#include "evaluator.hpp"

namespace moai
{
    Ct gelu_v2(const Evaluator &ev, const Ct &x, const Keys &keys);
    Ct layernorm(const Evaluator &ev, const Ct &x, const std::vector<double> &gamma, const std::vector<double> &beta,
                 const std::vector<int> &bias_vec, const Keys &keys, int variant);
    Ct exp_128(const Evaluator &ev, const Ct &x, const Keys &keys);
    Ct inverse(const Evaluator &ev, const Ct &x, const Keys &keys, int iter);
    Ct ct_ct_matrix_mul_colpacking(const Evaluator &ev, const Ct &X, const Ct &W, const Keys &keys, int col_X, int row_X,
                                   int col_W, int row_W, int num_batch);
    Ct ct_ct_matrix_mul_diagpacking(const Evaluator &ev, const Ct &X, const Ct &W, const Keys &keys, int col_X,
                                    int row_X, int col_W, int row_W, int num_batch);
} // namespace moai
