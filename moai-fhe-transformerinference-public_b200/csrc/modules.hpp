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

    // Weights of one BERT-base encoder layer as the reference driver reads them
    // (M/test/test_full_scheme.hpp:94-337): all matrices row-major [in][out] doubles.
    struct LayerWeights
    {
        int hidden = 768, heads = 12, head_dim = 64, inter = 3072;
        std::vector<std::vector<double>> WQ, WK, WV; // [heads] each hidden x head_dim
        std::vector<std::vector<double>> bQ, bK, bV; // [heads] each head_dim
        std::vector<double> selfoutput, selfoutput_bias;     // hidden x hidden, hidden
        std::vector<double> ln1_gamma, ln1_beta;             // hidden
        std::vector<double> inter_weight, inter_bias;        // hidden x inter, inter
        std::vector<double> final_weight, final_bias;        // inter x hidden, hidden
        std::vector<double> ln2_gamma, ln2_beta;             // hidden
    };

    // One encoder layer of all_layer_test (M/test/test_full_scheme.hpp:484-1087): x = 768 column
    // ciphertexts at chain_index 20; returns the layer output at chain_index 20 (after the 4th
    // bootstrapping), ready to be the next layer's input.  reuse_input: x's storage may be overwritten
    // once the first residual has consumed it (saves one 15.75 GiB buffer at the repo's size).
    Ct encoder_layer(const Evaluator &ev, const Ct &x, const LayerWeights &w, const std::vector<int> &bias_vec,
                     int input_num, const Keys &keys, Bootstrapper &boot, int num_batch, int layer_id,
                     long long boot_chunk, bool reuse_input = false);
    // One bootstrap-delimited quarter of the layer (stage 0..3) on two persistent buffers; see modules.cu.
    void encoder_layer_stage(const Evaluator &ev, int stage, Ct &x, Ct &aux, const LayerWeights &w,
                             const std::vector<int> &bias_vec, int input_num, const Keys &keys, Bootstrapper &boot,
                             int num_batch, int layer_id, long long boot_chunk);
} // namespace moai
