// Host-callable launchers of the evaluation kernels (all stream-ordered on Context::stream).
// Buffers are device pointers in SEAL's layout: a batch of B ciphertexts is
// [B][polys][limbs][n] uint64, a plaintext [limbs][n].
#pragma once
#include "context.hpp"

namespace moai
{
    enum EwOp
    {
        EW_ADD = 0,
        EW_SUB = 1,
        EW_NEG = 2,
    };

    // b_broadcast: b is ONE ciphertext applied to every batch item of a
    void ew_addsub(Context *c, int op, const u64 *a, const u64 *b, u64 *out, long long batch, int polys, int limbs,
                   bool b_broadcast = false);
    void sum_batch(Context *c, const u64 *a, u64 *out, long long batch, int polys, int limbs);
    // mode 0: out3 = sum_j a[j] (x) b[j];  mode 1: out3 = sum_j (a[j] - b)^2 with b one ciphertext
    void inner_product(Context *c, const u64 *a, const u64 *b, u64 *out3, long long batch, int limbs, int mode);
    // ct (+/-) pt on poly 0; pt_stride = elements between the plaintexts of consecutive batch items (0 = broadcast)
    void ew_addsub_plain(Context *c, int op, const u64 *ct, const u64 *pt, u64 *out, long long batch, int polys,
                         int limbs, long long pt_stride);
    void ew_multiply_plain(Context *c, const u64 *ct, const u64 *pt, u64 *out, long long batch, int polys, int limbs,
                           long long pt_stride);
    // per-limb scalar constants: out = ct * k[l]  (scalar-encoded plaintext, S/ckks.cpp:131-153)
    void ew_multiply_scalar(Context *c, const u64 *ct, const u64 *h_consts, u64 *out, long long batch, int polys,
                            int limbs);
    void ew_add_scalar(Context *c, const u64 *ct, const u64 *h_consts, u64 *out, long long batch, int polys, int limbs);
    // out = 2 ct (+ consts on poly 0): double_inplace + add_const in one pass
    void ew_double_add_scalar(Context *c, const u64 *ct, const u64 *h_consts, u64 *out, long long batch, int polys, int limbs);
    // acc3 (size 3) += x2 (size 2) on the first two polynomials
    void ew_add_into3(Context *c, u64 *acc3, const u64 *x2, long long batch, int limbs);
    // out = sum_j in[j] * k[j][l] over n_terms <= 8 batches; in[j] has in_limbs[j] >= limbs limbs per polynomial and
    // only its first `limbs` are read; h_consts is [n_terms][limbs] (host)
    void ew_lincomb_scalar(Context *c, int n_terms, const u64 *const *in, const int *in_limbs, const u64 *h_consts,
                           u64 *out, long long batch, int polys, int limbs);
    // a_limbs / b_limbs (0 = limbs): limbs stored per polynomial of the operands; an operand at a higher level is read
    // in place (its implicit mod-switch)
    void ew_multiply(Context *c, const u64 *a, const u64 *b, u64 *out3, long long batch, int limbs, bool accumulate,
                     bool b_broadcast = false, int a_limbs = 0, int b_limbs = 0);
    void ew_square(Context *c, const u64 *a, u64 *out3, long long batch, int limbs);

    // fused BSGS inner sums of one linear stage: out[g] = sum_j pt[g * n_baby + j] (.) rot[j]
    // (pt entries may be nullptr = absent diagonal; plaintexts broadcast over the batch)
    constexpr int BSGS_MAX_BABY = 16, BSGS_MAX_GIANT = 8;
    void bsgs_inner(Context *c, const u64 *const *rot, int n_baby, const u64 *const *pt, int n_giant, u64 *const *out,
                    long long batch, int limbs);

    // the same inner sums with the rotations left in the key-switch basis (lazy mod-down; csrc/ops.cu)
    struct KsShape;
    void bsgs_ext(Context *c, const u64 *const *acc, const uint32_t *const *perm, int n_baby, const u64 *const *pt,
                  int n_giant, u64 *const *out, const u64 *cP, long long batch, const KsShape &sh, bool accumulate);
    // the first CoeffToSlot stage on single-digit keys in natural order (one fused pass per 16 rotations; csrc/ops.cu)
    void bsgs_single(Context *c, const u64 *ext, const u64 *const *key, const uint32_t *const *perm, const u64 *const *pt,
                     int n_rot, int key_kl, u64 *out, const u64 *cP, long long batch, const KsShape &sh, bool accumulate);
    void moddown_special(Context *c, const u64 *in, long long P, int limbs, const u64 *addend, u64 *out,
                         bool addend_even_only);

    void rescale(Context *c, const u64 *in, u64 *out, long long batch, int polys, int limbs);
    void mod_switch_drop(Context *c, const u64 *in, u64 *out, long long batch, int polys, int limbs_in, int limbs_out);
    void mod_raise(Context *c, const u64 *in, u64 *out, long long batch, int polys, int limbs_out);

    void apply_galois_ntt(Context *c, const u64 *in, u64 *out, long long count_polys_limbs, uint32_t elt);
    // key_kl = limbs stored per key polynomial: 0 / c->kl for SEAL's layout [kl-1][2][kl][n], L + 1 for
    // a key truncated to L levels by key_prepare ([L][2][L+1][n])
    // k_extra > 0: `ksk` is a grouped-digit key of ksg_key_prepare (csrc/ksgroup.hpp; fast mode, not SEAL's residues)
    void switch_key(Context *c, u64 *ct, const u64 *target, long long batch, int limbs, const u64 *ksk, int key_kl = 0,
                    int k_extra = 0);
    void relinearize(Context *c, const u64 *in3, u64 *out2, long long batch, int limbs, const u64 *ksk,
                     int key_kl = 0, int k_extra = 0);
    // grouped keys (k_extra > 0) only: rescale_to_next(relinearize(in3)) in one division, out2 at limbs - 1
    void relinearize_rescale(Context *c, const u64 *in3, u64 *out2, long long batch, int limbs, const u64 *ksk, int key_kl,
                             int k_extra);
    void apply_galois(Context *c, const u64 *in, u64 *out, long long batch, int limbs, uint32_t elt, const u64 *ksk,
                      int key_kl = 0);
    // the two halves of a key switch, exposed for hoisting (one decomposition, many rotations)
    size_t ks_ext_bytes_per_ct(Context *c, int limbs);
    size_t ks_ext_budget();
    long long ks_chunk(Context *c, int limbs, long long batch, size_t budget_bytes);
    void ks_decompose(Context *c, const u64 *target, long long batch, int limbs, u64 *ext, long long target_stride = 0);
    void ks_mac_moddown(Context *c, const u64 *ext, long long batch, int limbs, const u64 *ksk, int key_kl,
                        const u64 *addend, bool addend_c0_only, u64 *out);
    void rotate_hoisted(Context *c, const u64 *ct, const u64 *ext, long long batch, int limbs, uint32_t elt,
                        const u64 *ksk_pre, int key_kl, u64 *out);
    // one rotation with a pre-permuted key and its own decomposition (fused key-switch kernel)
    void rotate_prepermuted(Context *c, const u64 *ct, long long batch, int limbs, uint32_t elt, const u64 *ksk_pre,
                            int key_kl, u64 *out, int k_extra = 0);
    // pieces shared with csrc/ksgroup.cu
    void ks_mac_int(Context *c, const u64 *ext, const u64 *ksk, u64 *acc, long long batch, const KsShape &sh, int key_kl,
                    int I);
    void divround_finish(Context *c, const u64 *in, const u64 *u, const u64 *addend, u64 *out, long long P, int targets,
                         int limbs_in, const Twiddle *d_inv, bool addend_even_only, int addend_group = 2,
                         const Twiddle *d_addend_mul = nullptr, int addend_limbs = 0);
    // up to KSM_R (csrc/ntt.cuh) hoisted rotations in one pass over the extended digits
    bool ks_multi_enabled(Context *c, int limbs);
    void rotate_hoisted_multi(Context *c, const u64 *ct, const u64 *ext, long long batch, int limbs, int n_rot,
                              const uint32_t *elts, const u64 *const *ksk_pre, const int *key_kl, u64 *const *outs);
    void key_prepare(Context *c, const u64 *in, uint32_t elt, int max_limbs, bool pre_permute, u64 *out);
    // seeded key / ciphertext components regenerated on the device (csrc/seedexpand.cu); h_seeds: count x 8 words (host)
    void expand_seeds(Context *c, const u64 *h_seeds, long long count, int limbs, u64 *d_out, long long out_stride);

    // fused module: out[C][2][limbs-1][n] = rescale(sum_j X[j] * encode_scalar(W[j][i]))
    void ct_pt_matmul_scalar(Context *c, const u64 *X, const double *h_W, int K, int C, int limbs, double scale,
                             u64 *out, const u64 *post_pt = nullptr);

    // tcgen05 (5th-generation tensor core) version of the byte-plane GEMM (csrc/matmul_tc5.cu)
    size_t tc5_packed_weight_bytes(int K, int C, int limbs, int np);
    void tc5_pack_weights(Context *c, const double *dW, unsigned char *Bp, int K, int C, int limbs, int np, double scale);
    void tc5_gemm(Context *c, const u64 *X, const unsigned char *Bp, u64 *Y, int K, int C_total, int c0, int cn, int np,
                  int limbs, int pl_first, int pl_count, cudaStream_t stream);

    // the same module with HOST buffers: upload, tensor-core GEMM and download pipelined per (polynomial, limb) slice
    void ct_pt_matmul_scalar_host(Context *c, const u64 *h_X, const double *h_W, int K, int C, int limbs, double scale,
                                  u64 *h_out);

    // CKKSEncoder::encode(vector) on the device; values: DEVICE [count][n_vals] complex (re, im)
    void encode_vector(Context *c, const double *d_values, long long count, int n_vals, double scale, int limbs,
                       u64 *out);
    void encode_masked_weights(Context *c, const double *d_w, const int *d_mask, long long count, double scale,
                               int limbs, u64 *out);
    // out[C][2][limbs-1][n] = rescale(sum_j X[j] * encode_vector(W[j][i] * mask))  (exact general-mask path)
    void ct_pt_matmul_masked(Context *c, const u64 *X, const double *h_W, const int *h_mask, int K, int C, int limbs,
                             double scale, u64 *out);
    // fast mode: one scalar GEMM with weights at scale / 2^26 times ONE mask plaintext at 2^26 (csrc/matmul.cu)
    bool ct_pt_matmul_masked_fast_ok(double scale);
    void ct_pt_matmul_masked_fast(Context *c, const u64 *X, const double *h_W, const int *h_mask, int K, int C, int limbs,
                                  double scale, u64 *out);
} // namespace moai
