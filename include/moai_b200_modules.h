/*
 * moai_b200_modules.h — module-level (fused) entry points of libmoai_b200.so: the MOAI free
 * functions of SURVEY §8(a) rows B/C that take whole vectors of ciphertexts, implemented as
 * device pipelines.  Same conventions as moai_b200.h (device pointers, SEAL layout, status codes).
 * M/ = include/ of the reference.
 */
#ifndef MOAI_B200_MODULES_H
#define MOAI_B200_MODULES_H

#include "moai_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* B1/B2: ct_pt_matrix_mul_wo_pre and ct_pt_matrix_mul_wo_pre_large
 * (M/source/matrix_mul/Ct_pt_matrix_mul.hpp:4-49, 51-101; they differ only in OpenMP tiling):
 *   out[i] = rescale_to_next( sum_j enc_X[j] * encode(W[j][i], scale) ),  i < col_W, j < row_W.
 * enc_X: device [row_W][2][limbs][N]; W: HOST row-major row_W x col_W doubles (the reference's
 * vector<vector<double>>); scale = enc_X[0].scale(); out: device [col_W][2][limbs-1][N].
 * The caller sets out[i].scale() = scale like the reference does (Ct_pt_matrix_mul.hpp:41).      */
int32_t moai_ct_pt_matrix_mul_wo_pre(moai_context *ctx, const uint64_t *enc_X, const double *W, int32_t col_X,
                                     int32_t col_W, int32_t row_W, int32_t limbs, double scale, uint64_t *out);

/* B3: ct_pt_matrix_mul_wo_pre_w_mask (M/source/matrix_mul/Ct_pt_matrix_mul.hpp:103-170):
 * the plaintext of weight w is encode(w * mask) with mask = bias_vec (HOST, N/2 ints, 1 = valid
 * token slot).  An all-ones mask takes the scalar path above (bit-identical); any other mask is
 * evaluated exactly: one device FFT + `limbs` NTTs per weight.                                   */
int32_t moai_ct_pt_matrix_mul_wo_pre_w_mask(moai_context *ctx, const uint64_t *enc_X, const double *W,
                                            const int32_t *bias_vec, int32_t col_X, int32_t col_W, int32_t row_W,
                                            int32_t limbs, double scale, uint64_t *out);

/* ---- evaluation keys: SEAL's RelinKeys / GaloisKeys (S/relinkeys.h, S/galoiskeys.h) as device
 * pointers; each key is one KSwitchKeys entry laid out [key_limbs-1][2][key_limbs][N].           */
typedef struct moai_keys moai_keys;
int32_t moai_keys_create(moai_context *ctx, moai_keys **out);
int32_t moai_keys_destroy(moai_keys *keys);
int32_t moai_keys_set_relin(moai_keys *keys, const uint64_t *ksk);
int32_t moai_keys_add_galois(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk);

/* Evaluator::rotate_vector incl. SEAL's NAF fallback for missing keys (S/evaluator.cpp:2667-2722) */
int32_t moai_rotate_vector(moai_context *ctx, moai_keys *keys, const uint64_t *in, uint64_t *out, int64_t batch,
                           int32_t limbs, int32_t steps);

/* Module outputs: `out` must hold as many size-2 ciphertexts as the function returns, at the input
 * limb count (the result is written at *out_limbs <= limbs limbs, contiguously); *out_scale is the
 * scale() the reference's result carries.                                                        */

/* B9: gelu_v2 (M/source/non_linear_func/gelu_others.hpp:4-154), batched over `batch` ciphertexts */
int32_t moai_gelu_v2(moai_context *ctx, moai_keys *keys, const uint64_t *x, int64_t batch, int32_t limbs,
                     double scale, uint64_t *out, int32_t *out_limbs, double *out_scale);
/* B8: layernorm (variant 1) / layernorm2 (variant 2) (M/source/non_linear_func/layernorm.hpp:157-547);
 * gamma, beta: host arrays of num_ct doubles; bias_vec: host N/2 ints                            */
int32_t moai_layernorm(moai_context *ctx, moai_keys *keys, const uint64_t *x, int32_t num_ct, int32_t limbs,
                       double scale, const double *gamma, const double *beta, const int32_t *bias_vec,
                       int32_t variant, uint64_t *out, int32_t *out_limbs, double *out_scale);
/* B7: exp and inverse (M/source/non_linear_func/softmax.hpp:9-47, 49-82), batched                */
int32_t moai_exp(moai_context *ctx, moai_keys *keys, const uint64_t *x, int64_t batch, int32_t limbs, double scale,
                 uint64_t *out, int32_t *out_limbs, double *out_scale);
int32_t moai_inverse(moai_context *ctx, moai_keys *keys, const uint64_t *x, int64_t batch, int32_t limbs,
                     double scale, int32_t iter, uint64_t *out, int32_t *out_limbs, double *out_scale);
/* B4/B5: ct_ct_matrix_mul_colpacking / _diagpacking (M/source/matrix_mul/Ct_ct_matrix_mul.hpp:5-156) */
int32_t moai_ct_ct_matrix_mul_colpacking(moai_context *ctx, moai_keys *keys, const uint64_t *enc_X,
                                         const uint64_t *enc_W, int32_t limbs, double scale_X, double scale_W,
                                         int32_t col_X, int32_t row_X, int32_t col_W, int32_t row_W,
                                         int32_t num_batch, uint64_t *out, int32_t *out_limbs, double *out_scale);
int32_t moai_ct_ct_matrix_mul_diagpacking(moai_context *ctx, moai_keys *keys, const uint64_t *enc_X,
                                          const uint64_t *enc_W, int32_t limbs, double scale_X, double scale_W,
                                          int32_t col_X, int32_t row_X, int32_t col_W, int32_t row_W,
                                          int32_t num_batch, uint64_t *out, int32_t *out_limbs, double *out_scale);

#ifdef __cplusplus
}
#endif
#endif /* MOAI_B200_MODULES_H */
