// MOAI module functions over the batched Evaluator (see modules.cu).
#pragma once
#include "bootstrap.hpp"
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
    // ct_pt matmul wrappers on Ct (csrc/matmul.cu)
    Ct ct_pt_matrix_mul_wo_pre(const Evaluator &ev, const Ct &X, const std::vector<double> &W, int col_W);
    Ct softmax_boot(const Evaluator &ev, const Ct &X, const std::vector<int> &bias_vec, int input_num, const Keys &keys,
                    int iter, Bootstrapper &boot, int layer_id);
    Ct single_att_block(const Evaluator &ev, const Ct &X, const std::vector<double> &WQ, const std::vector<double> &WK,
                        const std::vector<double> &WV, const std::vector<double> &bQ, const std::vector<double> &bK,
                        const std::vector<double> &bV, const std::vector<int> &bias_vec, int input_num,
                        const Keys &keys, Bootstrapper &boot, int num_batch, int iter, int layer_id);
} // namespace moai
