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

/* B1/B2 with HOST buffers — the shape of the reference's own call, whose vector<Ciphertext> lives in host
 * memory (Ct_pt_matrix_mul.hpp:4-6): host_enc_X [row_W][2][limbs][N] (pinned memory for full PCIe speed) ->
 * host_out [col_W][2][limbs-1][N], col_W <= 768.  Upload, GEMM and download are pipelined per
 * (polynomial, limb) slice; returns when host_out is complete.                                     */
int32_t moai_ct_pt_matrix_mul_wo_pre_host(moai_context *ctx, const uint64_t *host_enc_X, const double *W,
                                          int32_t col_X, int32_t col_W, int32_t row_W, int32_t limbs, double scale,
                                          uint64_t *host_out);

/* B3: ct_pt_matrix_mul_wo_pre_w_mask (M/source/matrix_mul/Ct_pt_matrix_mul.hpp:103-170):
 * the plaintext of weight w is encode(w * mask) with mask = bias_vec (HOST, N/2 ints, 1 = valid
 * token slot).  An all-ones mask takes the scalar path above (bit-identical); any other mask is
 * evaluated exactly: one device FFT + `limbs` NTTs per weight.                                   */
int32_t moai_ct_pt_matrix_mul_wo_pre_w_mask(moai_context *ctx, const uint64_t *enc_X, const double *W,
                                            const int32_t *bias_vec, int32_t col_X, int32_t col_W, int32_t row_W,
                                            int32_t limbs, double scale, uint64_t *out);
/* Fast-mode variant of the same module (same arguments and output level / scale; needs scale >= 2^44).  The reference
 * encodes one plaintext per weight, encode(w * mask) (Ct_pt_matrix_mul.hpp:127-134): K * C encodings, which the entry
 * point above reproduces bit for bit (2.8 s for 768 x 768 at N = 65536).  This one factorises it — [sum_j round(w_ji
 * scale / 2^26) X_j] (.) encode(mask at 2^26) — into ONE tensor-core GEMM and ONE plaintext: 15-30 ms, decrypted result
 * within 1e-4 relative of the float64 product like the exact path (tests/test_gpu_fullsize.py). */
int32_t moai_ct_pt_matrix_mul_wo_pre_w_mask_fast(moai_context *ctx, const uint64_t *enc_X, const double *W,
                                                 const int32_t *bias_vec, int32_t col_X, int32_t col_W, int32_t row_W,
                                                 int32_t limbs, double scale, uint64_t *out);

/* ---- evaluation keys: SEAL's RelinKeys / GaloisKeys (S/relinkeys.h, S/galoiskeys.h) as device
 * pointers; each key is one KSwitchKeys entry laid out [key_limbs-1][2][key_limbs][N].           */
typedef struct moai_keys moai_keys;
int32_t moai_keys_create(moai_context *ctx, moai_keys **out);
int32_t moai_keys_destroy(moai_keys *keys);
int32_t moai_keys_set_relin(moai_keys *keys, const uint64_t *ksk);
int32_t moai_keys_add_galois(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk);

/* ---- fast mode (NOT SEAL's residues: same plaintext, different but
 * equally small key-switching noise; judged on decrypted tolerance).
 * moai_key_prepare: device-side re-layout of one Galois key of SEAL's KeyGenerator
 * (S/keygenerator.cpp:303-336, [key_limbs-1][2][key_limbs][N]) into a level-truncated key
 * [max_limbs][2][max_limbs+1][N] (first max_limbs digits / data limbs + the special prime); with
 * pre_permute != 0 every limb is additionally mapped through the NTT-domain automorphism of
 * galois_elt^-1 (S/util/galois.cpp:192-218), K' = sigma^-1(K), so that
 *     rotate(ct) = sigma((c0, 0) + ModDown(sum_J digit_J(c1) (.) K'_J))
 * needs the digit decomposition of the UNROTATED c1 only: every rotation of one ciphertext shares
 * one decomposition ("hoisting"), which SEAL's rotate (automorphism first, S/evaluator.cpp:2635-2657)
 * cannot do.  moai_keys_add_galois_fast registers such a key (key_limbs = max_limbs + 1; several
 * truncations of one element may coexist, the smallest that covers the level is used);
 * moai_rotate_vector / the modules use it when no SEAL-layout key of that element is registered.
 * moai_rotate_many: out[s][batch][2][limbs][N] = rotate(in, steps[s]) for n_steps steps, hoisted
 * when every step has a pre-permuted key.                                                        */
int32_t moai_key_prepare(moai_context *ctx, const uint64_t *ksk_in, uint32_t galois_elt, int32_t max_limbs,
                         int32_t pre_permute, uint64_t *ksk_out);
int32_t moai_keys_add_galois_fast(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk_pre, int32_t key_limbs);
/* Seeded components.  A stock client ships the uniform half `a` of every key-switching-key digit
 * and of a symmetric ciphertext as a 64-byte seed (Serializable<T>::save; S/keygenerator.cpp:164-232,
 * S/util/rlwe.cpp:328-368); SEAL's loader regenerates it with sample_poly_uniform(Blake2xbPRNG(seed))
 * (S/util/rlwe.cpp:137-166, S/randomgen.cpp:176-211).  moai_expand_seeds does that on the device, bit for bit:
 * seeds = count x 8 uint64 (HOST memory, prng_seed_type), out[i] = device [limbs][N] residues modulo the first `limbs`
 * primes of the key-level list, consecutive polynomials out_stride_words apart (so the `a` halves of all digits of a
 * key can be written straight into its [digits][2][key_limbs][N] buffer). */
int32_t moai_expand_seeds(moai_context *ctx, const uint64_t *seeds, int64_t count, int32_t limbs, uint64_t *out,
                          int64_t out_stride_words);

/* Grouped-digit keys (csrc/ksgroup.hpp): the fast-mode replacement of SEAL's one-digit-per-prime key switch
 * (S/evaluator.cpp:2724-3021, l (l + 1) forward NTTs at l limbs).  K_G = sum_{J in G} K_J of a stock SEAL key is a
 * hybrid key-switching key for the special modulus P' = p * (the k_extra top data primes, unused by a ciphertext at
 * l <= L - k_extra limbs): digits * (l + k_extra + 1) NTTs with digits ~ l / (k_extra + 1).  Same plaintext, different
 * (negligible) noise; derived on the device from the keys the client already ships.
 *   moai_ksg_best_extra      the k_extra the library's cost model prefers at `limbs` (0 = SEAL's digits)
 *   moai_ksg_key_shape       layout of a grouped key: [digits][2][key_limbs = max_limbs + k_extra + 1][N]
 *   moai_key_prepare_grouped SEAL-layout key -> grouped key usable at every level <= max_limbs (pre_permute as in
 *                            moai_key_prepare; 0 for the relinearisation key)
 *   moai_keys_add_grouped    registers it: galois_elt = 0 for a relinearisation key.  Several variants per element
 *                            may coexist; every key switch picks the cheapest one that covers its level.          */
int32_t moai_ksg_best_extra(moai_context *ctx, int32_t limbs, int32_t *k_extra);
int32_t moai_ksg_key_shape(moai_context *ctx, int32_t k_extra, int32_t max_limbs, int32_t *digits, int32_t *key_limbs);
int32_t moai_key_prepare_grouped(moai_context *ctx, const uint64_t *ksk_in, uint32_t galois_elt, int32_t k_extra,
                                 int32_t max_limbs, int32_t pre_permute, uint64_t *ksk_out);
int32_t moai_keys_add_grouped(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk_grouped, int32_t k_extra,
                              int32_t max_limbs);
/* Single-digit keys for rotations of a MOD-RAISED ciphertext (Bootstrapper::modraise_inplace, Bootstrapper.cpp:2938-2992):
 * its c1 is the centred lift of residues modulo q_0, i.e. as an integer polynomial it is ONE digit below the special
 * prime, and sum_J K_J encrypts p * s': the key switch needs no decomposition at all (one NTT for the special prime).
 * moai_key_prepare_single: SEAL-layout key -> [1][2][key_limbs][N]; used by the first CoeffToSlot stage in hoisting
 * mode 2 (moai_bootstrapper_set_hoisting), which then runs on baby steps only.  Register keys prepared with
 * pre_permute = 0: the stage's fused kernel forms sigma(ext) (.) K by gathering the digit, not by permuting the key. */
int32_t moai_key_prepare_single(moai_context *ctx, const uint64_t *ksk_in, uint32_t galois_elt, int32_t pre_permute,
                                uint64_t *ksk_out);
int32_t moai_keys_add_single(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk_single);
/* SEAL-exact rotations with a level-truncated key (moai_key_prepare with pre_permute = 0): the residues are
 * SEAL's bit for bit — a key switch at l limbs never reads digits or limbs beyond l — at (L/35)^2 of the
 * memory; a rotation above L limbs with such a key is rejected.                                    */
int32_t moai_keys_add_galois_truncated(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk, int32_t key_limbs);
int32_t moai_rotate_many(moai_context *ctx, moai_keys *keys, const uint64_t *in, int64_t batch, int32_t limbs,
                         const int32_t *steps, int32_t n_steps, uint64_t *out);

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

/* ---- C1-C5: Bootstrapper (M/source/bootstrapping/Bootstrapper.{h,cpp}) --------------------------
 * create = constructor + prepare_mod_polynomial + generate_LT_coefficient_3 (Bootstrapper.h:15-221;
 * driver: M/test/test_full_scheme.hpp:413-448): total_limbs = data limbs after ModRaise (35),
 * final_scale = 2^46, boundary_K = 25, deg = 59, double_angles = scale_factor = 2, log_width = loge = 10.
 * required_steps = addLeftRotKeys_Linear_to_vector_3 (Bootstrapper.cpp:89-185): the rotation steps
 * whose Galois keys make the linear transforms run without SEAL's NAF fallback.
 * moai_bootstrap = Bootstrapper::bootstrap_3 (Bootstrapper.cpp:3496-3502) on a batch: in
 * [batch][2][1][N] at chain_index 0 -> out [batch][2][total_limbs-14][N], *out_scale = final_scale. */
typedef struct moai_bootstrapper moai_bootstrapper;
int32_t moai_bootstrapper_create(moai_context *ctx, int32_t total_limbs, double final_scale, int32_t boundary_K,
                                 int32_t deg, int32_t double_angles, int32_t log_width, moai_bootstrapper **out);
int32_t moai_bootstrapper_destroy(moai_bootstrapper *b);
/* fast mode: on = 1 plans the linear stages for hoisted baby steps (16 baby x 4 giant instead of 8 x 8); on = 2
 * additionally keeps the baby-step rotations in the key-switch basis (one mod-down per giant step instead of one per
 * baby step) and runs the first CoeffToSlot stage on single-digit keys with baby steps only; call before
 * moai_bootstrapper_required_steps[_levels] and register the keys with moai_keys_add_galois_fast / _grouped / _single.
 * moai_bootstrapper_required_step_levels reports the steps that want a single-digit key with limbs = 0.           */
int32_t moai_bootstrapper_set_hoisting(moai_bootstrapper *b, int32_t on);
int32_t moai_bootstrapper_required_steps(moai_bootstrapper *b, int32_t *steps, int32_t capacity, int32_t *count);
/* the same steps with the level (limb count) each is used at; step 0 = the complex conjugation
 * (Evaluator::complex_conjugate, S/evaluator.cpp:2635-2657): what a level-truncated or grouped key must cover */
int32_t moai_bootstrapper_required_step_levels(moai_bootstrapper *b, int32_t *steps, int32_t *limbs, int32_t capacity,
                                               int32_t *count);
/* Evaluator::relinearize (S/evaluator.cpp:1345-1400) / complex_conjugate through a key handle: SEAL-layout key, or the
 * cheapest fast-mode key registered for the level */
int32_t moai_relinearize_keys(moai_context *ctx, moai_keys *keys, const uint64_t *in3, uint64_t *out2, int64_t batch,
                              int32_t limbs);
/* rescale_to_next(relinearize(in3)) — the pair every module function ends a ciphertext product with
 * (S/evaluator.cpp:1345-1400 then :1402-1481) — out2 at limbs - 1.  With a grouped-digit relinearisation key (fast mode)
 * the key switch's division by P' and the rescale's division by q_last are ONE division by P' q_last
 * (csrc/ksgroup.cu: ksg_moddown_rescale); with SEAL's key it is the two exact calls. */
int32_t moai_relin_rescale_keys(moai_context *ctx, moai_keys *keys, const uint64_t *in3, uint64_t *out2, int64_t batch,
                                int32_t limbs);
int32_t moai_complex_conjugate_keys(moai_context *ctx, moai_keys *keys, const uint64_t *in, uint64_t *out, int64_t batch,
                                    int32_t limbs);
int32_t moai_bootstrap(moai_context *ctx, moai_bootstrapper *b, moai_keys *keys, const uint64_t *in, int64_t batch,
                       double scale, uint64_t *out, int32_t *out_limbs, double *out_scale);
/* Bootstrapping of REAL-slot messages (all of MOAI's activations), two per bootstrapping: z = a + i b is one
 * full-slot message, so ceil(batch/2) bootstrappings refresh `batch` ciphertexts; the halves are separated by
 * one conjugation at the output level (needs the conjugation key, which bootstrap_3 needs anyway).  Same
 * arguments as moai_bootstrap plus chunk_pairs (pairs bootstrapped together; <= 0: 32).  The imaginary parts
 * of the inputs must be zero (they are mixed into the partner otherwise).  moai_encoder_layer uses this for its
 * 4 x 768 bootstrappings (M/test/test_full_scheme.hpp:654-660, 758-764, 991-995, 1081-1085) unless the
 * environment sets MOAI_BOOT_PAIR=0.                                                                        */
/* Diagnostics: the same pipeline stopped after one phase, for per-phase decrypted-error reports (bootstrap_full_3's
 * phases, Bootstrapper.cpp:3231-3251).  stop_after 1: ModRaise -> batch ciphertexts at total_limbs, scale q0 (plaintext
 * t = m + q0 I); 2: CoeffToSlot -> 2 x batch ciphertexts (real halves, then imaginary halves) whose slots hold
 * t[bitrev(j)] / (K q0) and t[bitrev(j) + N/2] / (K q0); 3: EvalMod -> 2 x batch, slots ~ sin(2 pi t / q0).
 * out must hold 2 x batch ciphertexts at total_limbs. */
int32_t moai_bootstrap_phase_debug(moai_context *ctx, moai_bootstrapper *b, moai_keys *keys, const uint64_t *in,
                                   int64_t batch, double scale, int32_t stop_after, uint64_t *out, int64_t *out_count,
                                   int32_t *out_limbs, double *out_scale);
int32_t moai_bootstrap_real(moai_context *ctx, moai_bootstrapper *b, moai_keys *keys, const uint64_t *in,
                            int64_t batch, double scale, int64_t chunk_pairs, uint64_t *out, int32_t *out_limbs,
                            double *out_scale);
/* host-only inspection of the plan (no GPU): one linear stage's diagonals and the cosine coefficients */
int32_t moai_bootstrap_plan_debug(int32_t log_n, const uint64_t *primes, int32_t n_key_limbs, int32_t total_limbs,
                                  int32_t dir, int32_t stage, int32_t *n_diags, int32_t *offsets, double *diag_values,
                                  double *cheb, int32_t *n_cheb);

/* host-only: the EvalMod cosine fit for (boundary_K, deg, double_angles, log_width); *poly_levels = the levels the
 * polynomial spends, ceil(log2(deg + 1)).  moai_bootstrapper_create accepts any pair with
 * poly_levels + double_angles == 8 (the level budget of bootstrap_3): the reference's (59, 2); (31, 3) would save two
 * relinearizations per EvalMod but its fit error is 3e-4 after the double angles (2.7e-9 for (59, 2)): not advisable. */
int32_t moai_bootstrap_cosine_fit_debug(int32_t boundary_K, int32_t deg, int32_t double_angles, int32_t log_width,
                                        double *cheb, int32_t *n_cheb, int32_t *poly_levels);

/* ---- B7/B6: softmax_boot (M/source/non_linear_func/softmax.hpp:308-581) and single_att_block
 * (M/source/att_block/single_att_block.hpp:10-207); weights row-major num_col x col_W doubles.     */
int32_t moai_softmax_boot(moai_context *ctx, moai_keys *keys, moai_bootstrapper *b, const uint64_t *enc_X,
                          int32_t num, int32_t limbs, double scale, const int32_t *bias_vec, int32_t input_num,
                          int32_t iter, int32_t layer_id, uint64_t *out, int32_t *out_limbs, double *out_scale);
int32_t moai_single_att_block(moai_context *ctx, moai_keys *keys, moai_bootstrapper *b, const uint64_t *enc_X,
                              int32_t num_col, int32_t limbs, double scale, const double *WQ, const double *WK,
                              const double *WV, const double *bQ, const double *bK, const double *bV, int32_t col_W,
                              const int32_t *bias_vec, int32_t input_num, int32_t num_batch, int32_t iter,
                              int32_t layer_id, uint64_t *out, int32_t *out_limbs, double *out_scale);

/* ---- one encoder layer of all_layer_test (M/test/test_full_scheme.hpp:484-1087): attention
 * (12 heads) -> self-output matmul -> bootstrap -> residual + LayerNorm -> bootstrap ->
 * intermediate matmul -> GELU -> final matmul -> bootstrap -> residual + LayerNorm2 -> bootstrap.
 * x: [hidden][2][limbs][N] at chain_index 20; out: the next layer's input at chain_index 20.
 * All weight pointers are HOST row-major [in][out] doubles as the driver reads them (:94-337).
 * `out` may equal `x` (same shape): the layer then runs in place — x's storage receives the second
 * bootstrapping's output once the first residual has consumed the input, and finally the layer
 * output — which saves two 15.75 GiB buffers at the repo's size (the next layer's input is this
 * layer's output anyway, :1081-1087).                                                             */
typedef struct moai_layer_weights
{
    int32_t hidden, heads, head_dim, inter;
    const double *WQ, *WK, *WV;          /* [heads][hidden][head_dim] */
    const double *bQ, *bK, *bV;          /* [heads][head_dim] */
    const double *selfoutput, *selfoutput_bias;
    const double *ln1_gamma, *ln1_beta;
    const double *inter_weight, *inter_bias;
    const double *final_weight, *final_bias;
    const double *ln2_gamma, *ln2_beta;
} moai_layer_weights;
int32_t moai_encoder_layer(moai_context *ctx, moai_keys *keys, moai_bootstrapper *b, const uint64_t *x, int32_t limbs,
                           double scale, const moai_layer_weights *w, const int32_t *bias_vec, int32_t input_num,
                           int32_t num_batch, int32_t layer_id, int64_t boot_chunk, uint64_t *out,
                           int32_t *out_limbs, double *out_scale);
/* The same layer one bootstrap-delimited quarter at a time, on two caller-owned buffers of the layer's shape
 * ([hidden][2][limbs][N]): x holds the layer input (and, after stage 3, the layer output = the next layer's input), aux
 * the other live activation.  stage 0: attention + self-output matmul + bootstrapping (test_full_scheme.hpp:496-660)
 * reads x, writes aux; 1: residual + LayerNorm + bootstrapping (:686-773) reads both, writes x; 2: intermediate matmul
 * + GELU + final matmul + bootstrapping (:807-995) reads x, writes aux; 3: residual + LayerNorm2 + bootstrapping
 * (:1016-1087) reads both, writes x.  moai_encoder_layer is exactly stages 0..3 in order.  A serving loop uses this to
 * interleave several packed batches, to checkpoint between bootstrappings, and bench.py to time the layer in steps. */
int32_t moai_encoder_layer_stage(moai_context *ctx, moai_keys *keys, moai_bootstrapper *b, int32_t stage, uint64_t *x,
                                 uint64_t *aux, int32_t limbs, double scale, const moai_layer_weights *w,
                                 const int32_t *bias_vec, int32_t input_num, int32_t num_batch, int32_t layer_id,
                                 int64_t boot_chunk);
/* "name:ms:count;" for every profiled phase (see moai_profile_enable) */
/* ---- one packed batch over several GPUs of one node ("partitioning A": the independent units of the layer are column
 * ciphertexts — the 768 bootstrappings of a stage, test_full_scheme.hpp:654-660; the 12 heads, :530-533; the 3072
 * intermediate columns with their GELUs, :807-888).  One process per GPU; each process creates its context on its device
 * and joins the communicator: rank 0 obtains an id (moai_comm_unique_id) and ships the 128 bytes to the others by any
 * means (bench.py: torch.distributed), then every rank calls moai_comm_init.  From then on moai_encoder_layer[_stage] on
 * that context computes this rank's share of every sharded loop and exchanges the shares with an NCCL all-gather of raw
 * uint64 limbs over NVLink (sums are modular: no reduction collective); LayerNorm and the two K = 768 / 3072 -> 768 matmuls
 * are replicated (1 % of the layer).  Every rank ends each stage with the complete activations, bit-identical to a
 * single-GPU run.  NCCL is resolved at run time (dlopen of libnccl.so.2): the library loads without it. */
int32_t moai_comm_unique_id(uint8_t *out128);
int32_t moai_comm_init(moai_context *ctx, const uint8_t *id128, int32_t rank, int32_t world);
int32_t moai_comm_destroy(moai_context *ctx);
int32_t moai_comm_stats(moai_context *ctx, uint64_t *gathers, uint64_t *received_bytes);

int32_t moai_profile_dump(moai_context *ctx, char *buf, int32_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* MOAI_B200_MODULES_H */
