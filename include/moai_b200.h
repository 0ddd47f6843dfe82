/*
 * moai_b200.h — C ABI of libmoai_b200.so, the B200-native CKKS evaluation backend behind the
 * SEAL call surface that MOAI's module code uses.
 *
 * Conventions (modelled on SEAL's own flat wrapper, S/c/evaluator.h:16-83, S/c/defines.h:33-58):
 *   - every function returns a status (0 = OK); moai_last_error() gives the message of the last
 *     failure on the calling thread; no function throws across the boundary;
 *   - buffers are plain DEVICE pointers to uint64 residues in SEAL's own layout
 *     (S/ciphertext.h:339-370): a batch of B ciphertexts is [B][size][limbs][N], a plaintext
 *     [limbs][N], a key-switching key [digit][2][key_limbs][N] (S/kswitchkeys.h:335-340), all in
 *     NTT form; "limbs" is coeff_modulus_size of the operand's level (chain_index + 1);
 *   - every evaluator entry point is BATCHED over independent ciphertexts (the reference's only
 *     parallelism, `#pragma omp parallel for` over ciphertexts, M/source/matrix_mul/
 *     Ct_pt_matrix_mul.hpp:19) — batch = 1 gives the single-ciphertext SEAL call;
 *   - work is stream-ordered on the context's CUDA stream (moai_set_stream); results are
 *     canonical residues in [0, q), bit-identical to SEAL-4.1-bs on the same inputs;
 *   - there is no CPU fallback: without a CUDA device every call fails with MOAI_CUDA_ERROR.
 *
 * S/ = thirdparty/SEAL-4.1-bs/native/src/seal/, M/ = include/ of the reference.
 */
#ifndef MOAI_B200_H
#define MOAI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct moai_context moai_context;

enum
{
    MOAI_OK = 0,
    MOAI_INVALID_ARGUMENT = 1, /* std::invalid_argument in SEAL */
    MOAI_LOGIC_ERROR = 2,      /* std::logic_error in SEAL */
    MOAI_CUDA_ERROR = 3,
    MOAI_OUT_OF_MEMORY = 4
};

const char *moai_last_error(void);
int32_t moai_version(void);

/* ---- context: replaces SEALContext::ContextData tables (S/context.cpp:422-524) ------------- */
/* primes = coeff_modulus of the key level: data primes q_0..q_{L-1} then the special prime.    */
int32_t moai_context_create(int32_t log_n, const uint64_t *primes, int32_t n_key_limbs, int32_t device,
                            moai_context **out);
int32_t moai_context_destroy(moai_context *ctx);
/* A LANE of `ctx` for one more host thread: same tables (shared, immutable), its own CUDA stream and therefore its own
 * arena.  SEAL's Evaluator is called from many OpenMP threads at once (M/test/test_full_scheme.hpp:654-660 and every
 * module header's `#pragma omp parallel for`); with one lane per thread those calls are issued concurrently instead of
 * queueing on one stream.  Key handles and device buffers may be used from any lane; a buffer written on one lane must
 * be complete (moai_synchronize on that lane) before another lane reads it.  Destroy lanes (moai_context_destroy)
 * before their parent. */
int32_t moai_context_fork(moai_context *ctx, moai_context **lane);
int32_t moai_set_stream(moai_context *ctx, void *cuda_stream);
int32_t moai_synchronize(moai_context *ctx);

/* ---- measurement hooks used by bench.py ----------------------------------------------------
 * profile: when enabled, named device phases (e.g. "ctpt_gemm") are bracketed by CUDA events on
 * the launching stream and their elapsed ms accumulated; launch_count: kernels launched so far. */
int32_t moai_profile_enable(moai_context *ctx, int32_t on);
int32_t moai_profile_get(moai_context *ctx, const char *name, double *ms, int64_t *count);
int32_t moai_launch_count(moai_context *ctx, uint64_t *count);

/* ---- device memory (plumbing for hosts without their own allocator) ------------------------
 * The library keeps its temporaries in a stream-ordered best-fit block cache (the role SEAL's
 * MemoryPoolMT plays on the host, S/util/mempool.h:228): freed blocks are reused by later work on
 * the same stream and returned to the driver only when an allocation fails or on
 * moai_release_cached_memory. */
int32_t moai_malloc(moai_context *ctx, uint64_t bytes, void **out);
int32_t moai_free(moai_context *ctx, void *ptr);
int32_t moai_release_cached_memory(moai_context *ctx);
int32_t moai_memcpy_h2d(moai_context *ctx, void *dst, const void *src, uint64_t bytes);
int32_t moai_memcpy_d2h(moai_context *ctx, void *dst, const void *src, uint64_t bytes);
int32_t moai_memcpy_d2d(moai_context *ctx, void *dst, const void *src, uint64_t bytes);

/* ---- A1/A2: ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey (S/util/ntt.cpp:394-475) --
 * data: [batch][polys][limbs][N] in place; limb l uses prime l.                                */
int32_t moai_ntt_forward(moai_context *ctx, uint64_t *data, int64_t batch, int32_t polys, int32_t limbs);
int32_t moai_ntt_inverse(moai_context *ctx, uint64_t *data, int64_t batch, int32_t polys, int32_t limbs);
/* `count` polynomials that all use prime `limb` (Evaluator::transform_{to,from}_ntt on one limb) */
int32_t moai_ntt_forward_limb(moai_context *ctx, uint64_t *data, int64_t count, int32_t limb);
int32_t moai_ntt_inverse_limb(moai_context *ctx, uint64_t *data, int64_t count, int32_t limb);

/* ---- A5: Evaluator::add/sub/negate (S/evaluator.cpp:130-350) ------------------------------- */
int32_t moai_add(moai_context *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch, int32_t size,
                 int32_t limbs);
int32_t moai_sub(moai_context *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch, int32_t size,
                 int32_t limbs);
int32_t moai_negate(moai_context *ctx, const uint64_t *a, uint64_t *out, int64_t batch, int32_t size, int32_t limbs);
/* add_plain / sub_plain (S/evaluator.cpp:1938-2152); pt_stride = uint64 elements between the
 * plaintexts of consecutive batch items, 0 = one plaintext broadcast to the batch             */
int32_t moai_add_plain(moai_context *ctx, const uint64_t *ct, const uint64_t *pt, uint64_t *out, int64_t batch,
                       int32_t size, int32_t limbs, int64_t pt_stride);
int32_t moai_sub_plain(moai_context *ctx, const uint64_t *ct, const uint64_t *pt, uint64_t *out, int64_t batch,
                       int32_t size, int32_t limbs, int64_t pt_stride);

/* ---- A4: Evaluator::multiply_plain (S/evaluator.cpp:2154-2198, 2336-2373) ------------------ */
int32_t moai_multiply_plain(moai_context *ctx, const uint64_t *ct, const uint64_t *pt, uint64_t *out, int64_t batch,
                            int32_t size, int32_t limbs, int64_t pt_stride);

/* ---- A10: Evaluator::multiply / square for size-2 inputs (S/evaluator.cpp:770-909,1223-1282)
 * out: [batch][3][limbs][N].  accumulate != 0 adds into out (the size-3 running sum of
 * ct_ct_matrix_mul_colpacking, M/source/matrix_mul/Ct_ct_matrix_mul.hpp:33-41)                 */
int32_t moai_multiply(moai_context *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch,
                      int32_t limbs, int32_t accumulate);
int32_t moai_square(moai_context *ctx, const uint64_t *a, uint64_t *out, int64_t batch, int32_t limbs);

/* ---- A8: Evaluator::rescale_to_next (S/evaluator.cpp:1682-1720; S/util/rns.cpp:830-901) ----
 * in [batch][size][limbs][N] -> out [batch][size][limbs-1][N]; the caller divides the scale.   */
int32_t moai_rescale_to_next(moai_context *ctx, const uint64_t *in, uint64_t *out, int64_t batch, int32_t size,
                             int32_t limbs);
/* ---- A9: Evaluator::mod_switch_to_next / mod_switch_to (S/evaluator.cpp:1483-1652) --------- */
int32_t moai_mod_switch_to(moai_context *ctx, const uint64_t *in, uint64_t *out, int64_t batch, int32_t size,
                           int32_t limbs_in, int32_t limbs_out);

/* ---- A6/A7: key switching (S/evaluator.cpp:2724-3021) --------------------------------------
 * ksk: device key [key_limbs-1][2][key_limbs][N] exactly as SEAL stores one KSwitchKeys entry.  */
int32_t moai_galois_elt_from_step(moai_context *ctx, int32_t step, uint32_t *elt); /* S/util/galois.cpp:53-95 */
/* naf() + the skip rule of Evaluator::rotate_internal (S/evaluator.cpp:2699-2721): the power-of-two
 * steps SEAL applies, in order, when the Galois key of `steps` itself is absent.                */
int32_t moai_rotate_naf_steps(moai_context *ctx, int32_t steps, int32_t *out_steps, int32_t *out_count);
/* Evaluator::apply_galois (rotate_vector / complex_conjugate with the key present), out != in   */
int32_t moai_apply_galois(moai_context *ctx, const uint64_t *in, uint64_t *out, int64_t batch, int32_t limbs,
                          uint32_t galois_elt, const uint64_t *ksk);
/* Evaluator::relinearize (size 3 -> 2, S/evaluator.cpp:1345-1400)                               */
int32_t moai_relinearize(moai_context *ctx, const uint64_t *in3, uint64_t *out2, int64_t batch, int32_t limbs,
                         const uint64_t *ksk);
/* raw Evaluator::switch_key_inplace: ct[batch][2][limbs][N] += keyswitch(target[batch][limbs][N]) */
int32_t moai_switch_key(moai_context *ctx, uint64_t *ct, const uint64_t *target, int64_t batch, int32_t limbs,
                        const uint64_t *ksk);

/* ---- A11: CKKSEncoder::encode(double) (S/ckks.cpp:77-216): per-limb constants -------------- */
int32_t moai_encode_scalar_consts(moai_context *ctx, double value, double scale, int32_t limbs, uint64_t *host_out);
/* fork ops Evaluator::multiply_const / add_const (S/evaluator.cpp:395-409) with pre-encoded
 * per-limb constants (host array of `limbs` reduced residues)                                   */
int32_t moai_multiply_scalar(moai_context *ctx, const uint64_t *ct, const uint64_t *host_consts, uint64_t *out,
                             int64_t batch, int32_t size, int32_t limbs);
int32_t moai_add_scalar(moai_context *ctx, const uint64_t *ct, const uint64_t *host_consts, uint64_t *out,
                        int64_t batch, int32_t size, int32_t limbs);

/* ---- A12: CKKSEncoder::encode(vector<complex<double>>) (S/ckks.h:457-638), batched ---------
 * values: DEVICE [count][n_vals] complex numbers (interleaved re, im), n_vals <= N/2 (remaining
 * slots are zero); out: [count][limbs][N] NTT form.  Bit-identical to SEAL (same FFT order).    */
int32_t moai_encode_vector(moai_context *ctx, const double *values, int64_t count, int32_t n_vals, double scale,
                           int32_t limbs, uint64_t *out);

/* ---- C1: Bootstrapper::modraise_inplace (M/source/bootstrapping/Bootstrapper.cpp:2938-2992) -
 * in [batch][size][1][N] (limb q0) -> out [batch][size][limbs_out][N]                           */
int32_t moai_mod_raise(moai_context *ctx, const uint64_t *in, uint64_t *out, int64_t batch, int32_t size,
                       int32_t limbs_out);

#ifdef __cplusplus
}
#endif
#endif /* MOAI_B200_H */
