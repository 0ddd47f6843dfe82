// C ABI of libmoai_b200.so (declared in include/moai_b200.h).  Thin: argument checks with the
// reference's error semantics, then the stream-ordered launchers of ops.cuh / ntt.cuh.
#include "../../include/moai_b200.h"
#include "../../include/moai_b200_modules.h"
#include "ntt.cuh"
#include "bootstrap.hpp"
#include "modules.hpp"
#include "comm.hpp"
#include "ops.cuh"
#include <cmath>
#include <cstring>

using namespace moai;

struct moai_context
{
    Context *c;
};

struct moai_keys
{
    Keys k;
    int kl = 0; // key-level limb count of the owning context (SEAL-layout keys)
};

struct moai_bootstrapper
{
    Bootstrapper *b;
};

#define API_BEGIN                                                                                                      \
    try                                                                                                                \
    {
#define API_END                                                                                                        \
    }                                                                                                                  \
    catch (const StatusError &e)                                                                                       \
    {                                                                                                                  \
        set_last_error(e.msg);                                                                                         \
        return e.code;                                                                                                 \
    }                                                                                                                  \
    catch (const std::exception &e)                                                                                    \
    {                                                                                                                  \
        set_last_error(e.what());                                                                                      \
        return MOAI_LOGIC_ERROR;                                                                                       \
    }                                                                                                                  \
    return MOAI_OK;

static Context *get(moai_context *ctx)
{
    if (!ctx || !ctx->c)
    {
        throw StatusError{ INVALID_ARGUMENT, "null context" };
    }
    MOAI_CUDA_CHECK(cudaSetDevice(ctx->c->device));
    return ctx->c;
}

static void check_shape(Context *c, long long batch, int size, int limbs)
{
    MOAI_REQUIRE(batch >= 0, "negative batch");
    MOAI_REQUIRE(size >= 1 && size <= 3, "ciphertext size must be 1..3");
    MOAI_REQUIRE(limbs >= 1 && limbs <= c->kl, "encrypted is not valid for encryption parameters");
}

extern "C"
{
    const char *moai_last_error(void)
    {
        return last_error().c_str();
    }

    int32_t moai_version(void)
    {
        return 100;
    }

    int32_t moai_context_create(int32_t log_n, const uint64_t *primes, int32_t n_key_limbs, int32_t device,
                                moai_context **out)
    {
        API_BEGIN
        MOAI_REQUIRE(primes && out, "null argument");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
        {
            throw StatusError{ CUDA_ERROR, "no CUDA device: libmoai_b200 has no CPU fallback" };
        }
        Context *c = context_create(log_n, reinterpret_cast<const u64 *>(primes), n_key_limbs, device);
        *out = new moai_context{ c };
        API_END
    }

    int32_t moai_context_fork(moai_context *ctx, moai_context **lane)
    {
        API_BEGIN
        MOAI_REQUIRE(ctx && ctx->c && lane, "null argument");
        *lane = new moai_context{ context_fork(ctx->c) };
        API_END
    }

    int32_t moai_context_destroy(moai_context *ctx)
    {
        API_BEGIN
        if (ctx)
        {
            if (ctx->c)
            {
                cudaSetDevice(ctx->c->device);
                if (ctx->c->parent)
                {
                    delete ctx->c; // a lane: drains and destroys its own stream only
                }
                else
                {
                    cudaDeviceSynchronize();
                    device_release_cached();
                    delete ctx->c;
                }
            }
            delete ctx;
        }
        API_END
    }

    int32_t moai_set_stream(moai_context *ctx, void *cuda_stream)
    {
        API_BEGIN
        get(ctx)->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
        API_END
    }

    int32_t moai_synchronize(moai_context *ctx)
    {
        API_BEGIN
        MOAI_CUDA_CHECK(cudaStreamSynchronize(get(ctx)->stream));
        API_END
    }

    int32_t moai_profile_enable(moai_context *ctx, int32_t on)
    {
        API_BEGIN
        Context *c = get(ctx);
        c->kernel_timers_collect(true);
        c->profiling = on != 0;
        if (on)
        {
            c->prof.clear();
        }
        API_END
    }

    int32_t moai_profile_get(moai_context *ctx, const char *name, double *ms, int64_t *count)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(name && ms && count, "null argument");
        c->kernel_timers_collect(true);
        auto it = c->prof.find(name);
        *ms = it == c->prof.end() ? 0.0 : it->second.first;
        *count = it == c->prof.end() ? 0 : it->second.second;
        API_END
    }

    int32_t moai_launch_count(moai_context *ctx, uint64_t *count)
    {
        API_BEGIN
        MOAI_REQUIRE(count, "null argument");
        *count = get(ctx)->launches;
        API_END
    }

    int32_t moai_malloc(moai_context *ctx, uint64_t bytes, void **out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(out, "null argument");
        *out = device_alloc(bytes, c->stream);
        API_END
    }

    int32_t moai_free(moai_context *ctx, void *ptr)
    {
        API_BEGIN
        Context *c = get(ctx);
        device_free(ptr, c->stream);
        API_END
    }

    int32_t moai_release_cached_memory(moai_context *ctx)
    {
        API_BEGIN
        get(ctx);
        device_release_cached();
        API_END
    }

    int32_t moai_memcpy_h2d(moai_context *ctx, void *dst, const void *src, uint64_t bytes)
    {
        API_BEGIN
        MOAI_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, get(ctx)->stream));
        API_END
    }

    int32_t moai_memcpy_d2h(moai_context *ctx, void *dst, const void *src, uint64_t bytes)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        API_END
    }

    int32_t moai_memcpy_d2d(moai_context *ctx, void *dst, const void *src, uint64_t bytes)
    {
        API_BEGIN
        MOAI_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, get(ctx)->stream));
        API_END
    }

    int32_t moai_ntt_forward(moai_context *ctx, uint64_t *data, int64_t batch, int32_t polys, int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, polys, limbs);
        ntt_forward(c, reinterpret_cast<u64 *>(data), batch * polys * limbs, c->d_ids, limbs);
        API_END
    }

    int32_t moai_ntt_inverse(moai_context *ctx, uint64_t *data, int64_t batch, int32_t polys, int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, polys, limbs);
        ntt_inverse(c, reinterpret_cast<u64 *>(data), batch * polys * limbs, c->d_ids, limbs);
        API_END
    }

    int32_t moai_ntt_forward_limb(moai_context *ctx, uint64_t *data, int64_t count, int32_t limb)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(limb >= 0 && limb < c->kl, "limb out of range");
        ntt_forward(c, reinterpret_cast<u64 *>(data), count, c->d_ids + limb, 1);
        API_END
    }

    int32_t moai_ntt_inverse_limb(moai_context *ctx, uint64_t *data, int64_t count, int32_t limb)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(limb >= 0 && limb < c->kl, "limb out of range");
        ntt_inverse(c, reinterpret_cast<u64 *>(data), count, c->d_ids + limb, 1);
        API_END
    }

#define U(p) reinterpret_cast<u64 *>(p)
#define CU(p) reinterpret_cast<const u64 *>(p)

    int32_t moai_add(moai_context *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch,
                     int32_t size, int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs);
        ew_addsub(c, EW_ADD, CU(a), CU(b), U(out), batch, size, limbs);
        API_END
    }

    int32_t moai_sub(moai_context *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch,
                     int32_t size, int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs);
        ew_addsub(c, EW_SUB, CU(a), CU(b), U(out), batch, size, limbs);
        API_END
    }

    int32_t moai_negate(moai_context *ctx, const uint64_t *a, uint64_t *out, int64_t batch, int32_t size,
                        int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs);
        ew_addsub(c, EW_NEG, CU(a), CU(a), U(out), batch, size, limbs);
        API_END
    }

    int32_t moai_add_plain(moai_context *ctx, const uint64_t *ct, const uint64_t *pt, uint64_t *out, int64_t batch,
                           int32_t size, int32_t limbs, int64_t pt_stride)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs);
        ew_addsub_plain(c, EW_ADD, CU(ct), CU(pt), U(out), batch, size, limbs, pt_stride);
        API_END
    }

    int32_t moai_sub_plain(moai_context *ctx, const uint64_t *ct, const uint64_t *pt, uint64_t *out, int64_t batch,
                           int32_t size, int32_t limbs, int64_t pt_stride)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs);
        ew_addsub_plain(c, EW_SUB, CU(ct), CU(pt), U(out), batch, size, limbs, pt_stride);
        API_END
    }

    int32_t moai_multiply_plain(moai_context *ctx, const uint64_t *ct, const uint64_t *pt, uint64_t *out,
                                int64_t batch, int32_t size, int32_t limbs, int64_t pt_stride)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs);
        ew_multiply_plain(c, CU(ct), CU(pt), U(out), batch, size, limbs, pt_stride);
        API_END
    }

    int32_t moai_multiply(moai_context *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch,
                          int32_t limbs, int32_t accumulate)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 2, limbs);
        ew_multiply(c, CU(a), CU(b), U(out), batch, limbs, accumulate != 0);
        API_END
    }

    int32_t moai_square(moai_context *ctx, const uint64_t *a, uint64_t *out, int64_t batch, int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 2, limbs);
        ew_square(c, CU(a), U(out), batch, limbs);
        API_END
    }

    int32_t moai_rescale_to_next(moai_context *ctx, const uint64_t *in, uint64_t *out, int64_t batch, int32_t size,
                                 int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        // limbs == key limbs is the key level: dividing by the special prime with rounding is how a fresh
        // public-key encryption comes down to the first data level (S/encryptor.cpp:122-150)
        check_shape(c, batch, size, limbs);
        rescale(c, CU(in), U(out), batch, size, limbs);
        API_END
    }

    int32_t moai_mod_switch_to(moai_context *ctx, const uint64_t *in, uint64_t *out, int64_t batch, int32_t size,
                               int32_t limbs_in, int32_t limbs_out)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs_in);
        mod_switch_drop(c, CU(in), U(out), batch, size, limbs_in, limbs_out);
        API_END
    }

    int32_t moai_galois_elt_from_step(moai_context *ctx, int32_t step, uint32_t *elt)
    {
        API_BEGIN
        MOAI_REQUIRE(elt, "null argument");
        *elt = get(ctx)->elt_from_step(step);
        API_END
    }

    int32_t moai_rotate_naf_steps(moai_context *ctx, int32_t steps, int32_t *out_steps, int32_t *out_count)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(out_steps && out_count, "null argument");
        // non-adjacent form, least significant term first (S/util/numth.h:22-42)
        const bool neg = steps < 0;
        long long v = neg ? -(long long)steps : steps;
        int cnt = 0;
        for (int i = 0; v; i++)
        {
            int zi = (v & 1) ? 2 - (int)(v & 3) : 0;
            v = (v - zi) >> 1;
            if (zi)
            {
                long long term = (long long)(neg ? -zi : zi) * (1LL << i);
                long long mag = term < 0 ? -term : term;
                if ((size_t)mag != (c->n >> 1)) // a term of N/2 slots is the identity rotation
                {
                    out_steps[cnt++] = (int32_t)term;
                }
            }
        }
        *out_count = cnt;
        API_END
    }

    int32_t moai_apply_galois(moai_context *ctx, const uint64_t *in, uint64_t *out, int64_t batch, int32_t limbs,
                              uint32_t galois_elt, const uint64_t *ksk)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 2, limbs);
        MOAI_REQUIRE(limbs <= c->kl - 1, "encrypted is not valid for encryption parameters");
        MOAI_REQUIRE(ksk, "Galois key not present");
        apply_galois(c, CU(in), U(out), batch, limbs, galois_elt, CU(ksk));
        API_END
    }

    int32_t moai_relinearize(moai_context *ctx, const uint64_t *in3, uint64_t *out2, int64_t batch, int32_t limbs,
                             const uint64_t *ksk)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 3, limbs);
        MOAI_REQUIRE(limbs <= c->kl - 1, "encrypted is not valid for encryption parameters");
        MOAI_REQUIRE(ksk, "not enough relinearization keys");
        relinearize(c, CU(in3), U(out2), batch, limbs, CU(ksk));
        API_END
    }

    int32_t moai_switch_key(moai_context *ctx, uint64_t *ct, const uint64_t *target, int64_t batch, int32_t limbs,
                            const uint64_t *ksk)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 2, limbs);
        MOAI_REQUIRE(limbs <= c->kl - 1, "encrypted is not valid for encryption parameters");
        switch_key(c, U(ct), CU(target), batch, limbs, CU(ksk));
        API_END
    }

    int32_t moai_encode_scalar_consts(moai_context *ctx, double value, double scale, int32_t limbs,
                                      uint64_t *host_out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(limbs >= 1 && limbs <= c->kl && host_out, "bad arguments");
        MOAI_REQUIRE(scale > 0, "scale out of bounds");
        // CKKSEncoder::encode_internal(double): round(value*scale), sign-magnitude, residue per limb
        double v = value * scale;
        int bit_count = (int)(std::log2(std::fabs(v))) + 2;
        double r = std::round(v);
        const bool neg = std::signbit(r);
        r = std::fabs(r);
        MOAI_REQUIRE(bit_count <= 128, "encoded value is too large");
        unsigned __int128 mag;
        if (bit_count <= 64)
        {
            mag = (u64)r;
        }
        else
        {
            const double two64 = std::pow(2.0, 64);
            mag = (((unsigned __int128)(u64)(r / two64)) << 64) | (u64)std::fmod(r, two64);
        }
        for (int l = 0; l < limbs; l++)
        {
            u64 res = (u64)(mag % c->q[l]);
            host_out[l] = neg ? (res ? c->q[l] - res : 0) : res;
        }
        API_END
    }

    int32_t moai_multiply_scalar(moai_context *ctx, const uint64_t *ct, const uint64_t *host_consts, uint64_t *out,
                                 int64_t batch, int32_t size, int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs);
        ew_multiply_scalar(c, CU(ct), CU(host_consts), U(out), batch, size, limbs);
        API_END
    }

    int32_t moai_add_scalar(moai_context *ctx, const uint64_t *ct, const uint64_t *host_consts, uint64_t *out,
                            int64_t batch, int32_t size, int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs);
        ew_add_scalar(c, CU(ct), CU(host_consts), U(out), batch, size, limbs);
        API_END
    }

    int32_t moai_mod_raise(moai_context *ctx, const uint64_t *in, uint64_t *out, int64_t batch, int32_t size,
                           int32_t limbs_out)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, size, limbs_out);
        mod_raise(c, CU(in), U(out), batch, size, limbs_out);
        API_END
    }

    int32_t moai_ct_pt_matrix_mul_wo_pre(moai_context *ctx, const uint64_t *enc_X, const double *W, int32_t col_X,
                                         int32_t col_W, int32_t row_W, int32_t limbs, double scale, uint64_t *out)
    {
        API_BEGIN
        Context *c = get(ctx);
        // "ERROR: bad dimensions of X or W" (Ct_pt_matrix_mul.hpp:11-14)
        MOAI_REQUIRE(col_X == row_W, "bad dimensions of X or W");
        MOAI_REQUIRE(enc_X && W && out, "null argument");
        ct_pt_matmul_scalar(c, CU(enc_X), W, row_W, col_W, limbs, scale, U(out));
        API_END
    }

    int32_t moai_ct_pt_matrix_mul_wo_pre_host(moai_context *ctx, const uint64_t *host_enc_X, const double *W,
                                              int32_t col_X, int32_t col_W, int32_t row_W, int32_t limbs, double scale,
                                              uint64_t *host_out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(col_X == row_W, "bad dimensions of X or W");
        MOAI_REQUIRE(host_enc_X && W && host_out, "null argument");
        ct_pt_matmul_scalar_host(c, CU(host_enc_X), W, row_W, col_W, limbs, scale, U(host_out));
        API_END
    }

    int32_t moai_ct_pt_matrix_mul_wo_pre_w_mask(moai_context *ctx, const uint64_t *enc_X, const double *W,
                                                const int32_t *bias_vec, int32_t col_X, int32_t col_W, int32_t row_W,
                                                int32_t limbs, double scale, uint64_t *out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(col_X == row_W, "bad dimensions of X or W");
        MOAI_REQUIRE(enc_X && W && bias_vec && out, "null argument");
        bool all_ones = true;
        for (size_t s = 0; s < c->n / 2 && all_ones; s++)
        {
            all_ones = bias_vec[s] == 1;
        }
        if (all_ones)
        {
            // encode(vector of a constant) == encode(scalar) exactly
            ct_pt_matmul_scalar(c, CU(enc_X), W, row_W, col_W, limbs, scale, U(out));
        }
        else
        {
            ct_pt_matmul_masked(c, CU(enc_X), W, reinterpret_cast<const int *>(bias_vec), row_W, col_W, limbs, scale,
                                U(out));
        }
        API_END
    }

    int32_t moai_ct_pt_matrix_mul_wo_pre_w_mask_fast(moai_context *ctx, const uint64_t *enc_X, const double *W,
                                                     const int32_t *bias_vec, int32_t col_X, int32_t col_W, int32_t row_W,
                                                     int32_t limbs, double scale, uint64_t *out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(col_X == row_W, "bad dimensions of X or W");
        MOAI_REQUIRE(enc_X && W && bias_vec && out, "null argument");
        MOAI_REQUIRE(ct_pt_matmul_masked_fast_ok(scale), "scale too small for the factorised masked matmul (needs >= 2^44)");
        ct_pt_matmul_masked_fast(c, CU(enc_X), W, reinterpret_cast<const int *>(bias_vec), row_W, col_W, limbs, scale, U(out));
        API_END
    }

    int32_t moai_encode_vector(moai_context *ctx, const double *values, int64_t count, int32_t n_vals, double scale,
                               int32_t limbs, uint64_t *out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(limbs >= 1 && limbs <= c->kl && out, "bad arguments");
        encode_vector(c, values, count, n_vals, scale, limbs, U(out));
        API_END
    }

    // ---- key handles ---------------------------------------------------------------------------
    int32_t moai_keys_create(moai_context *ctx, moai_keys **out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(out, "null argument");
        *out = new moai_keys();
        (*out)->kl = c->kl;
        API_END
    }

    int32_t moai_keys_destroy(moai_keys *keys)
    {
        API_BEGIN
        delete keys;
        API_END
    }

    int32_t moai_keys_set_relin(moai_keys *keys, const uint64_t *ksk)
    {
        API_BEGIN
        MOAI_REQUIRE(keys, "null argument");
        keys->k.relin = KeyRef{ CU(ksk), keys->kl };
        API_END
    }

    int32_t moai_keys_add_galois(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk)
    {
        API_BEGIN
        MOAI_REQUIRE(keys && ksk, "null argument");
        keys->k.galois[galois_elt] = KeyRef{ CU(ksk), keys->kl };
        API_END
    }

    int32_t moai_keys_add_galois_truncated(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk, int32_t key_limbs)
    {
        API_BEGIN
        MOAI_REQUIRE(keys && ksk, "null argument");
        MOAI_REQUIRE(key_limbs >= 2 && key_limbs <= keys->kl, "key_limbs out of range");
        keys->k.galois[galois_elt] = KeyRef{ CU(ksk), key_limbs };
        API_END
    }

    int32_t moai_keys_add_galois_fast(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk_pre, int32_t key_limbs)
    {
        API_BEGIN
        MOAI_REQUIRE(keys && ksk_pre, "null argument");
        MOAI_REQUIRE(key_limbs >= 2 && key_limbs <= keys->kl, "key_limbs out of range");
        keys->k.galois_fast[galois_elt].push_back(KeyRef{ CU(ksk_pre), key_limbs });
        API_END
    }

    int32_t moai_key_prepare(moai_context *ctx, const uint64_t *ksk_in, uint32_t galois_elt, int32_t max_limbs,
                             int32_t pre_permute, uint64_t *ksk_out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(ksk_in && ksk_out, "null argument");
        key_prepare(c, CU(ksk_in), galois_elt, max_limbs, pre_permute != 0, U(ksk_out));
        API_END
    }

    // ---- seeded components (csrc/seedexpand.cu) ---------------------------------------------------
    int32_t moai_expand_seeds(moai_context *ctx, const uint64_t *seeds, int64_t count, int32_t limbs, uint64_t *out,
                              int64_t out_stride_words)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(seeds && out && count >= 1, "null argument");
        MOAI_REQUIRE(out_stride_words >= (int64_t)limbs * (int64_t)c->n, "output polynomials overlap");
        expand_seeds(c, CU(seeds), count, limbs, U(out), out_stride_words);
        API_END
    }

    // ---- grouped-digit keys (csrc/ksgroup.hpp) ----------------------------------------------------
    int32_t moai_ksg_best_extra(moai_context *ctx, int32_t limbs, int32_t *k_extra)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(k_extra && limbs >= 1 && limbs <= c->kl - 1, "bad argument");
        *k_extra = ksg_best_k(c, limbs);
        API_END
    }

    int32_t moai_ksg_key_shape(moai_context *ctx, int32_t k_extra, int32_t max_limbs, int32_t *digits, int32_t *key_limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(digits && key_limbs, "null argument");
        *digits = ksg_digits(c, k_extra, max_limbs);
        *key_limbs = ksg_key_kl(k_extra, max_limbs);
        API_END
    }

    int32_t moai_key_prepare_grouped(moai_context *ctx, const uint64_t *ksk_in, uint32_t galois_elt, int32_t k_extra,
                                     int32_t max_limbs, int32_t pre_permute, uint64_t *ksk_out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(ksk_in && ksk_out, "null argument");
        ksg_key_prepare(c, CU(ksk_in), galois_elt, k_extra, max_limbs, pre_permute != 0, U(ksk_out));
        API_END
    }

    int32_t moai_keys_add_grouped(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk_grouped, int32_t k_extra,
                                  int32_t max_limbs)
    {
        API_BEGIN
        MOAI_REQUIRE(keys && ksk_grouped, "null argument");
        MOAI_REQUIRE(k_extra >= 1 && max_limbs >= 1 && max_limbs + k_extra <= keys->kl - 1, "bad grouped-key shape");
        const KeyRef ref{ CU(ksk_grouped), ksg_key_kl(k_extra, max_limbs), k_extra };
        if (galois_elt == 0)
        {
            keys->k.relin_fast.push_back(ref);
        }
        else
        {
            keys->k.galois_fast[galois_elt].push_back(ref);
        }
        API_END
    }

    static const Keys &getk(moai_keys *keys)
    {
        if (!keys)
        {
            throw StatusError{ INVALID_ARGUMENT, "null keys" };
        }
        return keys->k;
    }

    static void emit(Context *c, const Ct &r, uint64_t *out, int64_t out_capacity_cts, int32_t *out_limbs,
                     double *out_scale)
    {
        MOAI_REQUIRE(out && out_limbs && out_scale, "null argument");
        MOAI_REQUIRE(out_capacity_cts >= r.batch, "output buffer too small");
        MOAI_CUDA_CHECK(cudaMemcpyAsync(out, r.d, (size_t)r.batch * r.size * r.limbs * c->n * sizeof(u64),
                                        cudaMemcpyDeviceToDevice, c->stream));
        *out_limbs = r.limbs;
        *out_scale = r.scale;
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream)); // r's storage is released on return
    }

    int32_t moai_rotate_vector(moai_context *ctx, moai_keys *keys, const uint64_t *in, uint64_t *out, int64_t batch,
                               int32_t limbs, int32_t steps)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 2, limbs);
        Evaluator ev(c);
        Ct a = ev.wrap(const_cast<u64 *>(CU(in)), batch, 2, limbs, 1.0);
        Ct r = ev.rotate_vector(a, steps, getk(keys));
        MOAI_CUDA_CHECK(cudaMemcpyAsync(out, r.d, (size_t)batch * 2 * limbs * c->n * sizeof(u64),
                                        cudaMemcpyDeviceToDevice, c->stream));
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        API_END
    }

    int32_t moai_rotate_many(moai_context *ctx, moai_keys *keys, const uint64_t *in, int64_t batch, int32_t limbs,
                             const int32_t *steps, int32_t n_steps, uint64_t *out)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 2, limbs);
        MOAI_REQUIRE(steps && out && n_steps >= 0, "bad arguments");
        Evaluator ev(c);
        Ct a = ev.wrap(const_cast<u64 *>(CU(in)), batch, 2, limbs, 1.0);
        std::vector<Ct> r = ev.rotate_many(a, std::vector<int>(steps, steps + n_steps), getk(keys));
        const size_t per = (size_t)batch * 2 * limbs * c->n;
        for (int s = 0; s < n_steps; s++)
        {
            MOAI_CUDA_CHECK(cudaMemcpyAsync(out + s * per, r[s].d, per * sizeof(u64), cudaMemcpyDeviceToDevice,
                                            c->stream));
        }
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        API_END
    }

    int32_t moai_gelu_v2(moai_context *ctx, moai_keys *keys, const uint64_t *x, int64_t batch, int32_t limbs,
                         double scale, uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 2, limbs);
        Evaluator ev(c);
        Ct r = gelu_v2(ev, ev.wrap(const_cast<u64 *>(CU(x)), batch, 2, limbs, scale), getk(keys));
        emit(c, r, out, batch, out_limbs, out_scale);
        API_END
    }

    int32_t moai_layernorm(moai_context *ctx, moai_keys *keys, const uint64_t *x, int32_t num_ct, int32_t limbs,
                           double scale, const double *gamma, const double *beta, const int32_t *bias_vec,
                           int32_t variant, uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, num_ct, 2, limbs);
        MOAI_REQUIRE(gamma && beta && bias_vec, "null argument");
        MOAI_REQUIRE(variant == 1 || variant == 2, "variant must be 1 (layernorm) or 2 (layernorm2)");
        Evaluator ev(c);
        std::vector<double> g(gamma, gamma + num_ct), b(beta, beta + num_ct);
        std::vector<int> bv(bias_vec, bias_vec + c->n / 2);
        Ct r = layernorm(ev, ev.wrap(const_cast<u64 *>(CU(x)), num_ct, 2, limbs, scale), g, b, bv, getk(keys), variant);
        emit(c, r, out, num_ct, out_limbs, out_scale);
        API_END
    }

    int32_t moai_exp(moai_context *ctx, moai_keys *keys, const uint64_t *x, int64_t batch, int32_t limbs, double scale,
                     uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 2, limbs);
        Evaluator ev(c);
        Ct r = exp_128(ev, ev.wrap(const_cast<u64 *>(CU(x)), batch, 2, limbs, scale), getk(keys));
        emit(c, r, out, batch, out_limbs, out_scale);
        API_END
    }

    int32_t moai_inverse(moai_context *ctx, moai_keys *keys, const uint64_t *x, int64_t batch, int32_t limbs,
                         double scale, int32_t iter, uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, batch, 2, limbs);
        Evaluator ev(c);
        Ct r = inverse(ev, ev.wrap(const_cast<u64 *>(CU(x)), batch, 2, limbs, scale), getk(keys), iter);
        emit(c, r, out, batch, out_limbs, out_scale);
        API_END
    }

    int32_t moai_ct_ct_matrix_mul_colpacking(moai_context *ctx, moai_keys *keys, const uint64_t *enc_X,
                                             const uint64_t *enc_W, int32_t limbs, double scale_X, double scale_W,
                                             int32_t col_X, int32_t row_X, int32_t col_W, int32_t row_W,
                                             int32_t num_batch, uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, col_X, 2, limbs);
        Evaluator ev(c);
        Ct X = ev.wrap(const_cast<u64 *>(CU(enc_X)), col_X, 2, limbs, scale_X);
        Ct W = ev.wrap(const_cast<u64 *>(CU(enc_W)), col_W, 2, limbs, scale_W);
        Ct r = ct_ct_matrix_mul_colpacking(ev, X, W, getk(keys), col_X, row_X, col_W, row_W, num_batch);
        emit(c, r, out, row_X, out_limbs, out_scale);
        API_END
    }

    int32_t moai_ct_ct_matrix_mul_diagpacking(moai_context *ctx, moai_keys *keys, const uint64_t *enc_X,
                                              const uint64_t *enc_W, int32_t limbs, double scale_X, double scale_W,
                                              int32_t col_X, int32_t row_X, int32_t col_W, int32_t row_W,
                                              int32_t num_batch, uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, row_X, 2, limbs);
        Evaluator ev(c);
        Ct X = ev.wrap(const_cast<u64 *>(CU(enc_X)), row_X, 2, limbs, scale_X);
        Ct W = ev.wrap(const_cast<u64 *>(CU(enc_W)), col_W, 2, limbs, scale_W);
        Ct r = ct_ct_matrix_mul_diagpacking(ev, X, W, getk(keys), col_X, row_X, col_W, row_W, num_batch);
        emit(c, r, out, col_W, out_limbs, out_scale);
        API_END
    }

    // ---- bootstrapping ---------------------------------------------------------------------------
    int32_t moai_bootstrapper_create(moai_context *ctx, int32_t total_limbs, double final_scale, int32_t boundary_K,
                                     int32_t deg, int32_t double_angles, int32_t log_width, moai_bootstrapper **out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(out, "null argument");
        BootParams p;
        p.total_limbs = total_limbs;
        p.final_scale = final_scale;
        p.boundary_K = boundary_K;
        p.deg = deg;
        p.double_angles = double_angles;
        p.log_width = log_width;
        *out = new moai_bootstrapper{ new Bootstrapper(c, p) };
        API_END
    }

    int32_t moai_bootstrapper_destroy(moai_bootstrapper *b)
    {
        API_BEGIN
        if (b)
        {
            delete b->b;
            delete b;
        }
        API_END
    }

    int32_t moai_key_prepare_single(moai_context *ctx, const uint64_t *ksk_in, uint32_t galois_elt, int32_t pre_permute,
                                    uint64_t *ksk_out)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(ksk_in && ksk_out, "null argument");
        ks_key_prepare_single(c, CU(ksk_in), galois_elt, pre_permute != 0, U(ksk_out));
        API_END
    }

    int32_t moai_keys_add_single(moai_keys *keys, uint32_t galois_elt, const uint64_t *ksk_single)
    {
        API_BEGIN
        MOAI_REQUIRE(keys && ksk_single, "null argument");
        keys->k.galois_single[galois_elt] = KeyRef{ CU(ksk_single), keys->kl, KS_SINGLE };
        API_END
    }

    int32_t moai_bootstrapper_set_hoisting(moai_bootstrapper *b, int32_t on)
    {
        API_BEGIN
        MOAI_REQUIRE(b, "null argument");
        b->b->set_hoisting(on);
        API_END
    }

    int32_t moai_bootstrapper_required_steps(moai_bootstrapper *b, int32_t *steps, int32_t capacity, int32_t *count)
    {
        API_BEGIN
        MOAI_REQUIRE(b && steps && count, "null argument");
        auto v = b->b->required_steps();
        MOAI_REQUIRE((int)v.size() <= capacity, "steps buffer too small");
        for (size_t i = 0; i < v.size(); i++)
        {
            steps[i] = v[i];
        }
        *count = (int32_t)v.size();
        API_END
    }

    int32_t moai_bootstrapper_required_step_levels(moai_bootstrapper *b, int32_t *steps, int32_t *limbs, int32_t capacity,
                                                   int32_t *count)
    {
        API_BEGIN
        MOAI_REQUIRE(b && steps && limbs && count, "null argument");
        auto v = b->b->required_step_levels();
        MOAI_REQUIRE((int)v.size() <= capacity, "steps buffer too small");
        for (size_t i = 0; i < v.size(); i++)
        {
            steps[i] = v[i].first;
            limbs[i] = v[i].second;
        }
        *count = (int32_t)v.size();
        API_END
    }

    int32_t moai_relinearize_keys(moai_context *ctx, moai_keys *keys, const uint64_t *in3, uint64_t *out2, int64_t batch,
                                  int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(in3 && out2, "null argument");
        check_shape(c, batch, 3, limbs);
        Evaluator ev(c);
        Ct r = ev.relinearize(ev.wrap(const_cast<u64 *>(CU(in3)), batch, 3, limbs, 1.0), getk(keys));
        MOAI_CUDA_CHECK(cudaMemcpyAsync(out2, r.d, (size_t)batch * 2 * limbs * c->n * sizeof(u64), cudaMemcpyDeviceToDevice,
                                        c->stream));
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        API_END
    }

    int32_t moai_relin_rescale_keys(moai_context *ctx, moai_keys *keys, const uint64_t *in3, uint64_t *out2, int64_t batch,
                                    int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(in3 && out2, "null argument");
        MOAI_REQUIRE(limbs >= 2, "end of modulus switching chain reached");
        check_shape(c, batch, 3, limbs);
        Evaluator ev(c);
        Ct r = ev.relin_rescale(ev.wrap(const_cast<u64 *>(CU(in3)), batch, 3, limbs, 1.0), getk(keys));
        MOAI_CUDA_CHECK(cudaMemcpyAsync(out2, r.d, (size_t)batch * 2 * (limbs - 1) * c->n * sizeof(u64),
                                        cudaMemcpyDeviceToDevice, c->stream));
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        API_END
    }

    int32_t moai_complex_conjugate_keys(moai_context *ctx, moai_keys *keys, const uint64_t *in, uint64_t *out, int64_t batch,
                                        int32_t limbs)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(in && out && in != out, "null or aliased argument");
        check_shape(c, batch, 2, limbs);
        Evaluator ev(c);
        Ct r = ev.complex_conjugate(ev.wrap(const_cast<u64 *>(CU(in)), batch, 2, limbs, 1.0), getk(keys));
        MOAI_CUDA_CHECK(cudaMemcpyAsync(out, r.d, (size_t)batch * 2 * limbs * c->n * sizeof(u64), cudaMemcpyDeviceToDevice,
                                        c->stream));
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        API_END
    }

    int32_t moai_bootstrap(moai_context *ctx, moai_bootstrapper *b, moai_keys *keys, const uint64_t *in, int64_t batch,
                           double scale, uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(b && in && out, "null argument");
        check_shape(c, batch, 2, 1);
        Evaluator ev(c);
        Ct r = b->b->bootstrap(ev, ev.wrap(const_cast<u64 *>(CU(in)), batch, 2, 1, scale), getk(keys));
        MOAI_REQUIRE(out_limbs && out_scale, "null argument");
        MOAI_CUDA_CHECK(cudaMemcpyAsync(out, r.d, (size_t)r.batch * 2 * r.limbs * c->n * sizeof(u64),
                                        cudaMemcpyDeviceToDevice, c->stream));
        *out_limbs = r.limbs;
        *out_scale = r.scale;
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        API_END
    }

    int32_t moai_bootstrap_phase_debug(moai_context *ctx, moai_bootstrapper *b, moai_keys *keys, const uint64_t *in,
                                       int64_t batch, double scale, int32_t stop_after, uint64_t *out, int64_t *out_count,
                                       int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(b && in && out && out_count && out_limbs && out_scale, "null argument");
        MOAI_REQUIRE(stop_after >= 1 && stop_after <= 3, "stop_after must be 1 (ModRaise), 2 (CoeffToSlot) or 3 (EvalMod)");
        check_shape(c, batch, 2, 1);
        Evaluator ev(c);
        Ct r = b->b->bootstrap(ev, ev.wrap(const_cast<u64 *>(CU(in)), batch, 2, 1, scale), getk(keys), stop_after);
        MOAI_CUDA_CHECK(cudaMemcpyAsync(out, r.d, (size_t)r.batch * 2 * r.limbs * c->n * sizeof(u64),
                                        cudaMemcpyDeviceToDevice, c->stream));
        *out_count = r.batch;
        *out_limbs = r.limbs;
        *out_scale = r.scale;
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        API_END
    }

    int32_t moai_bootstrap_real(moai_context *ctx, moai_bootstrapper *b, moai_keys *keys, const uint64_t *in,
                                int64_t batch, double scale, int64_t chunk_pairs, uint64_t *out, int32_t *out_limbs,
                                double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(b && in && out && out_limbs && out_scale, "null argument");
        check_shape(c, batch, 2, 1);
        Evaluator ev(c);
        const int ol = b->b->prm.total_limbs - 14;
        Ct into = ev.wrap(reinterpret_cast<u64 *>(out), batch, 2, ol, b->b->prm.final_scale);
        Ct r = b->b->bootstrap_real_pairs(ev, ev.wrap(const_cast<u64 *>(CU(in)), batch, 2, 1, scale), getk(keys),
                                          chunk_pairs > 0 ? chunk_pairs : 32, &into);
        *out_limbs = r.limbs;
        *out_scale = r.scale;
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        API_END
    }

    // Host-only view of the bootstrapping plan (no GPU needed): the sparse-diagonal matrix of one
    // linear stage (dir 0 = CoeffToSlot, 1 = SlotToCoeff; stage 0..2) and the cosine coefficients.
    // Call with diag_values == NULL to query *n_diags.
    int32_t moai_bootstrap_plan_debug(int32_t log_n, const uint64_t *primes, int32_t n_key_limbs, int32_t total_limbs,
                                      int32_t dir, int32_t stage, int32_t *n_diags, int32_t *offsets,
                                      double *diag_values, double *cheb, int32_t *n_cheb)
    {
        API_BEGIN
        MOAI_REQUIRE(primes && n_diags && stage >= 0 && stage < 3, "bad arguments");
        Context host; // tables are not needed for the plan: only n and the prime list
        host.log_n = log_n;
        host.n = (size_t)1 << log_n;
        host.kl = n_key_limbs;
        host.q.assign(primes, primes + n_key_limbs);
        BootParams p;
        p.total_limbs = total_limbs;
        Bootstrapper b(&host, p);
        const LinearStage &st = b.stage(dir, stage);
        *n_diags = (int32_t)st.diags.size();
        if (diag_values && offsets)
        {
            size_t k = 0;
            const size_t n = host.n / 2;
            for (auto &kv : st.diags)
            {
                offsets[k] = kv.first;
                for (size_t i = 0; i < n; i++)
                {
                    diag_values[(k * n + i) * 2] = kv.second[i].real();
                    diag_values[(k * n + i) * 2 + 1] = kv.second[i].imag();
                }
                k++;
            }
        }
        if (cheb && n_cheb)
        {
            *n_cheb = (int32_t)b.cheb_coeffs().size();
            for (size_t i = 0; i < b.cheb_coeffs().size(); i++)
            {
                cheb[i] = b.cheb_coeffs()[i];
            }
        }
        API_END
    }

    // Host-only: the Chebyshev coefficients of the EvalMod cosine cos(2 pi (x - 1/4) / 2^double_angles) fitted on
    // the intervals [i - 2^-log_width, i + 2^-log_width], |i| < boundary_K, in y = x / boundary_K (no GPU needed).
    int32_t moai_bootstrap_cosine_fit_debug(int32_t boundary_K, int32_t deg, int32_t double_angles, int32_t log_width,
                                            double *cheb, int32_t *n_cheb, int32_t *poly_levels)
    {
        API_BEGIN
        MOAI_REQUIRE(cheb && n_cheb && poly_levels, "null argument");
        static const uint64_t dummy_primes[17] = { 0 };
        Context host;
        host.log_n = 4;
        host.n = 16;
        host.kl = 17;
        host.q.assign(dummy_primes, dummy_primes + 17);
        BootParams p;
        p.boundary_K = boundary_K;
        p.deg = deg;
        p.double_angles = double_angles;
        p.log_width = log_width;
        p.total_limbs = 16;
        Bootstrapper b(&host, p);
        *poly_levels = p.poly_levels();
        *n_cheb = (int32_t)b.cheb_coeffs().size();
        for (size_t i = 0; i < b.cheb_coeffs().size(); i++)
        {
            cheb[i] = b.cheb_coeffs()[i];
        }
        API_END
    }

    // ---- attention / encoder layer ----------------------------------------------------------------
    int32_t moai_softmax_boot(moai_context *ctx, moai_keys *keys, moai_bootstrapper *b, const uint64_t *enc_X,
                              int32_t num, int32_t limbs, double scale, const int32_t *bias_vec, int32_t input_num,
                              int32_t iter, int32_t layer_id, uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, num, 2, limbs);
        MOAI_REQUIRE(b && bias_vec, "null argument");
        Evaluator ev(c);
        std::vector<int> bv(bias_vec, bias_vec + c->n / 2);
        Ct r = softmax_boot(ev, ev.wrap(const_cast<u64 *>(CU(enc_X)), num, 2, limbs, scale), bv, input_num, getk(keys),
                            iter, *b->b, layer_id);
        emit(c, r, out, num, out_limbs, out_scale);
        API_END
    }

    int32_t moai_single_att_block(moai_context *ctx, moai_keys *keys, moai_bootstrapper *b, const uint64_t *enc_X,
                                  int32_t num_col, int32_t limbs, double scale, const double *WQ, const double *WK,
                                  const double *WV, const double *bQ, const double *bK, const double *bV, int32_t col_W,
                                  const int32_t *bias_vec, int32_t input_num, int32_t num_batch, int32_t iter,
                                  int32_t layer_id, uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        check_shape(c, num_col, 2, limbs);
        MOAI_REQUIRE(b && bias_vec && WQ && WK && WV && bQ && bK && bV, "null argument");
        Evaluator ev(c);
        std::vector<int> bv(bias_vec, bias_vec + c->n / 2);
        const size_t wn = (size_t)num_col * col_W;
        std::vector<double> wq(WQ, WQ + wn), wk(WK, WK + wn), wv(WV, WV + wn);
        std::vector<double> q(bQ, bQ + col_W), k(bK, bK + col_W), v(bV, bV + col_W);
        Ct r = single_att_block(ev, ev.wrap(const_cast<u64 *>(CU(enc_X)), num_col, 2, limbs, scale), wq, wk, wv, q, k, v,
                                bv, input_num, getk(keys), *b->b, num_batch, iter, layer_id);
        emit(c, r, out, col_W, out_limbs, out_scale);
        API_END
    }

    static LayerWeights to_layer_weights(const moai_layer_weights *w)
    {
        MOAI_REQUIRE(w->hidden == w->heads * w->head_dim, "hidden must equal heads * head_dim");
        LayerWeights lw;
        lw.hidden = w->hidden;
        lw.heads = w->heads;
        lw.head_dim = w->head_dim;
        lw.inter = w->inter;
        const size_t hw = (size_t)w->hidden * w->head_dim;
        for (int h = 0; h < w->heads; h++)
        {
            lw.WQ.emplace_back(w->WQ + h * hw, w->WQ + (h + 1) * hw);
            lw.WK.emplace_back(w->WK + h * hw, w->WK + (h + 1) * hw);
            lw.WV.emplace_back(w->WV + h * hw, w->WV + (h + 1) * hw);
            lw.bQ.emplace_back(w->bQ + (size_t)h * w->head_dim, w->bQ + (size_t)(h + 1) * w->head_dim);
            lw.bK.emplace_back(w->bK + (size_t)h * w->head_dim, w->bK + (size_t)(h + 1) * w->head_dim);
            lw.bV.emplace_back(w->bV + (size_t)h * w->head_dim, w->bV + (size_t)(h + 1) * w->head_dim);
        }
        const size_t H = w->hidden, I = w->inter;
        lw.selfoutput.assign(w->selfoutput, w->selfoutput + H * H);
        lw.selfoutput_bias.assign(w->selfoutput_bias, w->selfoutput_bias + H);
        lw.ln1_gamma.assign(w->ln1_gamma, w->ln1_gamma + H);
        lw.ln1_beta.assign(w->ln1_beta, w->ln1_beta + H);
        lw.inter_weight.assign(w->inter_weight, w->inter_weight + H * I);
        lw.inter_bias.assign(w->inter_bias, w->inter_bias + I);
        lw.final_weight.assign(w->final_weight, w->final_weight + I * H);
        lw.final_bias.assign(w->final_bias, w->final_bias + H);
        lw.ln2_gamma.assign(w->ln2_gamma, w->ln2_gamma + H);
        lw.ln2_beta.assign(w->ln2_beta, w->ln2_beta + H);
        return lw;
    }

    int32_t moai_encoder_layer(moai_context *ctx, moai_keys *keys, moai_bootstrapper *b, const uint64_t *x,
                               int32_t limbs, double scale, const moai_layer_weights *w, const int32_t *bias_vec,
                               int32_t input_num, int32_t num_batch, int32_t layer_id, int64_t boot_chunk,
                               uint64_t *out, int32_t *out_limbs, double *out_scale)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(b && w && bias_vec, "null argument");
        check_shape(c, w->hidden, 2, limbs);
        Evaluator ev(c);
        LayerWeights lw = to_layer_weights(w);
        std::vector<int> bv(bias_vec, bias_vec + c->n / 2);
        Ct r = encoder_layer(ev, ev.wrap(const_cast<u64 *>(CU(x)), w->hidden, 2, limbs, scale), lw, bv, input_num,
                             getk(keys), *b->b, num_batch, layer_id, boot_chunk > 0 ? boot_chunk : 32,
                             /*reuse_input=*/out == x);
        if (reinterpret_cast<const uint64_t *>(r.d) == out)
        {
            MOAI_REQUIRE(out_limbs && out_scale, "null argument");
            *out_limbs = r.limbs;
            *out_scale = r.scale;
            MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        }
        else
        {
            emit(c, r, out, w->hidden, out_limbs, out_scale);
        }
        API_END
    }

    int32_t moai_encoder_layer_stage(moai_context *ctx, moai_keys *keys, moai_bootstrapper *b, int32_t stage, uint64_t *x,
                                     uint64_t *aux, int32_t limbs, double scale, const moai_layer_weights *w,
                                     const int32_t *bias_vec, int32_t input_num, int32_t num_batch, int32_t layer_id,
                                     int64_t boot_chunk)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(b && w && bias_vec && x && aux && x != aux, "null or aliased argument");
        check_shape(c, w->hidden, 2, limbs);
        Evaluator ev(c);
        LayerWeights lw = to_layer_weights(w);
        std::vector<int> bv(bias_vec, bias_vec + c->n / 2);
        Ct cx = ev.wrap(reinterpret_cast<u64 *>(x), w->hidden, 2, limbs, scale);
        Ct ca = ev.wrap(reinterpret_cast<u64 *>(aux), w->hidden, 2, limbs, scale);
        encoder_layer_stage(ev, stage, cx, ca, lw, bv, input_num, getk(keys), *b->b, num_batch, layer_id,
                            boot_chunk > 0 ? boot_chunk : 32);
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream)); // temporaries of the stage are released on return
        API_END
    }

    // ---- one packed batch over several GPUs (comm.hpp) ------------------------------------------------------------
    int32_t moai_comm_unique_id(uint8_t *out128)
    {
        API_BEGIN
        MOAI_REQUIRE(out128, "null argument");
        comm_unique_id(out128);
        API_END
    }

    int32_t moai_comm_init(moai_context *ctx, const uint8_t *id128, int32_t rank, int32_t world)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(id128 || world == 1, "null argument");
        comm_init(c, id128, rank, world);
        API_END
    }

    int32_t moai_comm_destroy(moai_context *ctx)
    {
        API_BEGIN
        comm_destroy(get(ctx));
        API_END
    }

    int32_t moai_comm_stats(moai_context *ctx, uint64_t *gathers, uint64_t *received_bytes)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(gathers && received_bytes, "null argument");
        *gathers = c->comm ? c->comm->gathers : 0;
        *received_bytes = c->comm ? c->comm->gathered_bytes : 0;
        API_END
    }

    int32_t moai_profile_dump(moai_context *ctx, char *buf, int32_t capacity)
    {
        API_BEGIN
        Context *c = get(ctx);
        MOAI_REQUIRE(buf && capacity > 0, "bad arguments");
        c->kernel_timers_collect(true);
        std::string s;
        for (auto &kv : c->prof)
        {
            s += kv.first + ":" + std::to_string(kv.second.first) + ":" + std::to_string(kv.second.second) + ";";
        }
        const AllocStats as = alloc_stats();
        s += "alloc_host:" + std::to_string(as.host_ms) + ":" + std::to_string(as.calls) + ";";
        s += "alloc_cache_flushes:0:" + std::to_string(as.retries) + ";";
        s += "alloc_owned_GiB:" + std::to_string(as.owned_bytes / 1073741824.0) + ":1;";
        MOAI_REQUIRE((int)s.size() < capacity, "buffer too small");
        std::memcpy(buf, s.c_str(), s.size() + 1);
        API_END
    }
}
