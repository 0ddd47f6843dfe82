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

#ifdef __cplusplus
}
#endif
#endif /* MOAI_B200_MODULES_H */
