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
    void ew_multiply(Context *c, const u64 *a, const u64 *b, u64 *out3, long long batch, int limbs, bool accumulate,
                     bool b_broadcast = false);
    void ew_square(Context *c, const u64 *a, u64 *out3, long long batch, int limbs);

    void rescale(Context *c, const u64 *in, u64 *out, long long batch, int polys, int limbs);
    void mod_switch_drop(Context *c, const u64 *in, u64 *out, long long batch, int polys, int limbs_in, int limbs_out);
    void mod_raise(Context *c, const u64 *in, u64 *out, long long batch, int polys, int limbs_out);

    void apply_galois_ntt(Context *c, const u64 *in, u64 *out, long long count_polys_limbs, uint32_t elt);
    void switch_key(Context *c, u64 *ct, const u64 *target, long long batch, int limbs, const u64 *ksk);
    void relinearize(Context *c, const u64 *in3, u64 *out2, long long batch, int limbs, const u64 *ksk);
    void apply_galois(Context *c, const u64 *in, u64 *out, long long batch, int limbs, uint32_t elt, const u64 *ksk);

    // fused module: out[C][2][limbs-1][n] = rescale(sum_j X[j] * encode_scalar(W[j][i]))
    void ct_pt_matmul_scalar(Context *c, const u64 *X, const double *h_W, int K, int C, int limbs, double scale,
                             u64 *out);

    // CKKSEncoder::encode(vector) on the device; values: DEVICE [count][n_vals] complex (re, im)
    void encode_vector(Context *c, const double *d_values, long long count, int n_vals, double scale, int limbs,
                       u64 *out);
    void encode_masked_weights(Context *c, const double *d_w, const int *d_mask, long long count, double scale,
                               int limbs, u64 *out);
    // out[C][2][limbs-1][n] = rescale(sum_j X[j] * encode_vector(W[j][i] * mask))  (exact general-mask path)
    void ct_pt_matmul_masked(Context *c, const u64 *X, const double *h_W, const int *h_mask, int K, int C, int limbs,
                             double scale, u64 *out);
} // namespace moai
