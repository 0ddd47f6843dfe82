/* mock_cabi.c — TEST DOUBLE of libmoai_b200.so's C ABI for CPU-only host-logic tests.  NOT PRODUCT CODE.
 *
 * The header-only facade include/moai_b200_seal.hpp is host logic (metadata, SEAL's checks and
 * exception rules, operation sequencing) over the C ABI.  This container has no GPU, so the `-m "not gpu"`
 * tests bind the facade to THIS file instead of the CUDA library: every entry point the facade uses is
 * forwarded to the CPU oracle (oracle/ckks_oracle.c), "device" memory is malloc.  That lets the
 * reference's unmodified module headers be run through the facade on the CPU and compared bit for bit
 * with the same headers on the reference's real SEAL (tests/test_facade.py).  The `-m gpu` tests run
 * the identical driver against the real libmoai_b200.so.
 *
 * Built only by tests/facade_harness/build.py into oracle/_ref/ (git-ignored); nothing in the product package
 * or in libmoai_b200.so references it, and the product never falls back to it
 * (tests/test_abi.py::test_product_never_imports_oracle).  Entry points the facade does not need on the
 * CPU (GPU-only modules, fast-mode keys) return MOAI_LOGIC_ERROR; the bootstrapper is a plumbing fake (below).
 */
#include "moai_b200.h"
#include "moai_b200_modules.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef struct orc_ctx orc_ctx;
/* oracle/ckks_oracle.c */
extern orc_ctx *orc_create_from_primes(int log_n, const u64 *primes, int n_bits);
extern void orc_destroy(orc_ctx *c);
extern void orc_ntt(const orc_ctx *c, int limb, u64 *v);
extern void orc_intt(const orc_ctx *c, int limb, u64 *v);
extern void orc_addsub(const orc_ctx *c, int op, const u64 *a, const u64 *b, int polys, int limbs, u64 *out);
extern void orc_addsub_plain(const orc_ctx *c, int op, const u64 *ct, const u64 *pt, int polys, int limbs, u64 *out);
extern void orc_multiply_plain(const orc_ctx *c, const u64 *ct, const u64 *pt, int polys, int limbs, u64 *out);
extern void orc_multiply(const orc_ctx *c, const u64 *a, const u64 *b, int limbs, u64 *out);
extern void orc_square(const orc_ctx *c, const u64 *a, int limbs, u64 *out);
extern void orc_rescale(const orc_ctx *c, const u64 *in, int polys, int limbs, u64 *out);
extern uint32_t orc_elt_from_step(const orc_ctx *c, int step);
extern void orc_relinearize(const orc_ctx *c, const u64 *ct3, int limbs, const u64 *relin_key, u64 *out2);
extern void orc_apply_galois(const orc_ctx *c, const u64 *ct, int limbs, uint32_t elt, const u64 *gal_key, u64 *out);
extern int orc_naf_steps(const orc_ctx *c, int steps, int *out);
extern int orc_encode_scalar_consts(const orc_ctx *c, double value, double scale, int limbs, u64 *consts);
extern int orc_encode_vector(const orc_ctx *c, const double *values, int n_vals, double scale, int limbs, u64 *out);

struct moai_context
{
    orc_ctx *o;
    int log_n, kl;
    size_t n;
    u64 q[64];
};

#define MAX_GAL 4096
struct moai_keys
{
    moai_context *c;
    const u64 *relin;
    int n_gal;
    uint32_t elt[MAX_GAL];
    const u64 *key[MAX_GAL];
};

static _Thread_local char g_err[256];
static int fail(int code, const char *msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
#define REQ(cond, msg)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
            return fail(MOAI_INVALID_ARGUMENT, msg);                                                                   \
    } while (0)
#define UNSUPPORTED(name) return fail(MOAI_LOGIC_ERROR, name ": not available in the CPU test double")

const char *moai_last_error(void) { return g_err; }
int32_t moai_version(void) { return -1; /* marks the test double */ }

int32_t moai_context_create(int32_t log_n, const uint64_t *primes, int32_t n_key_limbs, int32_t device, moai_context **out)
{
    (void)device;
    REQ(primes && out && n_key_limbs >= 1 && n_key_limbs <= 64, "bad arguments");
    moai_context *c = (moai_context *)calloc(1, sizeof *c);
    c->o = orc_create_from_primes(log_n, primes, n_key_limbs);
    if (!c->o)
    {
        free(c);
        return fail(MOAI_INVALID_ARGUMENT, "encryption parameters are not set correctly");
    }
    c->log_n = log_n;
    c->kl = n_key_limbs;
    c->n = (size_t)1 << log_n;
    memcpy(c->q, primes, 8 * (size_t)n_key_limbs);
    *out = c;
    return MOAI_OK;
}
int32_t moai_context_fork(moai_context *c, moai_context **lane)
{
    /* the test double has no streams: a lane is an independent context over the same primes (thread-safe by construction) */
    REQ(c && lane, "bad arguments");
    return moai_context_create(c->log_n, c->q, c->kl, 0, lane);
}
int32_t moai_context_destroy(moai_context *c)
{
    if (c)
    {
        orc_destroy(c->o);
        free(c);
    }
    return MOAI_OK;
}
int32_t moai_set_stream(moai_context *c, void *s) { (void)c; (void)s; return MOAI_OK; }
int32_t moai_synchronize(moai_context *c) { (void)c; return MOAI_OK; }
int32_t moai_malloc(moai_context *c, uint64_t bytes, void **out)
{
    (void)c;
    *out = malloc(bytes ? bytes : 8);
    return *out ? MOAI_OK : fail(MOAI_OUT_OF_MEMORY, "out of memory");
}
int32_t moai_free(moai_context *c, void *p) { (void)c; free(p); return MOAI_OK; }
int32_t moai_release_cached_memory(moai_context *c) { (void)c; return MOAI_OK; }
int32_t moai_memcpy_h2d(moai_context *c, void *d, const void *s, uint64_t b) { (void)c; memcpy(d, s, b); return MOAI_OK; }
int32_t moai_memcpy_d2h(moai_context *c, void *d, const void *s, uint64_t b) { (void)c; memcpy(d, s, b); return MOAI_OK; }
int32_t moai_memcpy_d2d(moai_context *c, void *d, const void *s, uint64_t b) { (void)c; memmove(d, s, b); return MOAI_OK; }

static int shape_ok(moai_context *c, int64_t batch, int size, int limbs)
{
    return c && batch >= 0 && size >= 1 && limbs >= 1 && limbs <= c->kl;
}

int32_t moai_ntt_forward(moai_context *c, uint64_t *d, int64_t batch, int32_t polys, int32_t limbs)
{
    REQ(shape_ok(c, batch, polys, limbs), "bad shape");
    for (int64_t i = 0; i < batch * polys; i++)
        for (int l = 0; l < limbs; l++)
            orc_ntt(c->o, l, d + ((size_t)i * limbs + l) * c->n);
    return MOAI_OK;
}
int32_t moai_ntt_inverse(moai_context *c, uint64_t *d, int64_t batch, int32_t polys, int32_t limbs)
{
    REQ(shape_ok(c, batch, polys, limbs), "bad shape");
    for (int64_t i = 0; i < batch * polys; i++)
        for (int l = 0; l < limbs; l++)
            orc_intt(c->o, l, d + ((size_t)i * limbs + l) * c->n);
    return MOAI_OK;
}

static int32_t binop(moai_context *c, int op, const u64 *a, const u64 *b, u64 *out, int64_t batch, int size, int limbs)
{
    REQ(shape_ok(c, batch, size, limbs), "bad shape");
    size_t per = (size_t)size * limbs * c->n;
    for (int64_t i = 0; i < batch; i++)
        orc_addsub(c->o, op, a + i * per, b ? b + i * per : a + i * per, size, limbs, out + i * per);
    return MOAI_OK;
}
int32_t moai_add(moai_context *c, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch, int32_t size, int32_t limbs)
{
    return binop(c, 0, a, b, out, batch, size, limbs);
}
int32_t moai_sub(moai_context *c, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch, int32_t size, int32_t limbs)
{
    return binop(c, 1, a, b, out, batch, size, limbs);
}
int32_t moai_negate(moai_context *c, const uint64_t *a, uint64_t *out, int64_t batch, int32_t size, int32_t limbs)
{
    return binop(c, 2, a, NULL, out, batch, size, limbs);
}

static int32_t plainop(moai_context *c, int op, const u64 *ct, const u64 *pt, u64 *out, int64_t batch, int size, int limbs,
                       int64_t pt_stride)
{
    REQ(shape_ok(c, batch, size, limbs) && pt, "bad shape");
    size_t per = (size_t)size * limbs * c->n;
    for (int64_t i = 0; i < batch; i++)
    {
        if (op == 2)
            orc_multiply_plain(c->o, ct + i * per, pt + i * pt_stride, size, limbs, out + i * per);
        else
            orc_addsub_plain(c->o, op, ct + i * per, pt + i * pt_stride, size, limbs, out + i * per);
    }
    return MOAI_OK;
}
int32_t moai_add_plain(moai_context *c, const uint64_t *ct, const uint64_t *pt, uint64_t *out, int64_t batch, int32_t size,
                       int32_t limbs, int64_t pt_stride)
{
    return plainop(c, 0, ct, pt, out, batch, size, limbs, pt_stride);
}
int32_t moai_sub_plain(moai_context *c, const uint64_t *ct, const uint64_t *pt, uint64_t *out, int64_t batch, int32_t size,
                       int32_t limbs, int64_t pt_stride)
{
    return plainop(c, 1, ct, pt, out, batch, size, limbs, pt_stride);
}
int32_t moai_multiply_plain(moai_context *c, const uint64_t *ct, const uint64_t *pt, uint64_t *out, int64_t batch,
                            int32_t size, int32_t limbs, int64_t pt_stride)
{
    return plainop(c, 2, ct, pt, out, batch, size, limbs, pt_stride);
}

int32_t moai_multiply(moai_context *c, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch, int32_t limbs,
                      int32_t accumulate)
{
    REQ(shape_ok(c, batch, 2, limbs) && !accumulate, "bad shape");
    for (int64_t i = 0; i < batch; i++)
        orc_multiply(c->o, a + (size_t)i * 2 * limbs * c->n, b + (size_t)i * 2 * limbs * c->n, limbs,
                     out + (size_t)i * 3 * limbs * c->n);
    return MOAI_OK;
}
int32_t moai_square(moai_context *c, const uint64_t *a, uint64_t *out, int64_t batch, int32_t limbs)
{
    REQ(shape_ok(c, batch, 2, limbs), "bad shape");
    for (int64_t i = 0; i < batch; i++)
        orc_square(c->o, a + (size_t)i * 2 * limbs * c->n, limbs, out + (size_t)i * 3 * limbs * c->n);
    return MOAI_OK;
}
int32_t moai_rescale_to_next(moai_context *c, const uint64_t *in, uint64_t *out, int64_t batch, int32_t size, int32_t limbs)
{
    REQ(shape_ok(c, batch, size, limbs) && limbs >= 2, "end of modulus switching chain reached");
    for (int64_t i = 0; i < batch; i++)
        orc_rescale(c->o, in + (size_t)i * size * limbs * c->n, size, limbs, out + (size_t)i * size * (limbs - 1) * c->n);
    return MOAI_OK;
}
int32_t moai_mod_switch_to(moai_context *c, const uint64_t *in, uint64_t *out, int64_t batch, int32_t size, int32_t limbs_in,
                           int32_t limbs_out)
{
    REQ(shape_ok(c, batch, size, limbs_in) && limbs_out >= 1 && limbs_out <= limbs_in, "cannot switch to higher level modulus");
    for (int64_t p = 0; p < batch * size; p++)
        memmove(out + (size_t)p * limbs_out * c->n, in + (size_t)p * limbs_in * c->n, (size_t)limbs_out * c->n * 8);
    return MOAI_OK;
}
int32_t moai_galois_elt_from_step(moai_context *c, int32_t step, uint32_t *elt)
{
    *elt = orc_elt_from_step(c->o, step);
    return MOAI_OK;
}
int32_t moai_rotate_naf_steps(moai_context *c, int32_t steps, int32_t *out_steps, int32_t *out_count)
{
    *out_count = orc_naf_steps(c->o, steps, out_steps);
    return MOAI_OK;
}
int32_t moai_apply_galois(moai_context *c, const uint64_t *in, uint64_t *out, int64_t batch, int32_t limbs, uint32_t elt,
                          const uint64_t *ksk)
{
    REQ(shape_ok(c, batch, 2, limbs) && ksk, "bad shape");
    for (int64_t i = 0; i < batch; i++)
        orc_apply_galois(c->o, in + (size_t)i * 2 * limbs * c->n, limbs, elt, ksk, out + (size_t)i * 2 * limbs * c->n);
    return MOAI_OK;
}
int32_t moai_relinearize(moai_context *c, const uint64_t *in3, uint64_t *out2, int64_t batch, int32_t limbs, const uint64_t *ksk)
{
    REQ(shape_ok(c, batch, 3, limbs) && ksk, "bad shape");
    for (int64_t i = 0; i < batch; i++)
        orc_relinearize(c->o, in3 + (size_t)i * 3 * limbs * c->n, limbs, ksk, out2 + (size_t)i * 2 * limbs * c->n);
    return MOAI_OK;
}
int32_t moai_switch_key(moai_context *c, uint64_t *ct, const uint64_t *t, int64_t b, int32_t l, const uint64_t *k)
{
    (void)c; (void)ct; (void)t; (void)b; (void)l; (void)k;
    UNSUPPORTED("moai_switch_key");
}
int32_t moai_encode_scalar_consts(moai_context *c, double value, double scale, int32_t limbs, uint64_t *host_out)
{
    REQ(c && limbs >= 1 && limbs <= c->kl, "parms_id is not valid for encryption parameters");
    if (orc_encode_scalar_consts(c->o, value, scale, limbs, host_out))
        return fail(MOAI_INVALID_ARGUMENT, "encoded value is too large");
    return MOAI_OK;
}
static int32_t scalarop(moai_context *c, int mul, const u64 *ct, const u64 *k, u64 *out, int64_t batch, int size, int limbs)
{
    REQ(shape_ok(c, batch, size, limbs), "bad shape");
    /* a scalar plaintext is the same constant in every NTT slot of a limb */
    u64 *pt = (u64 *)malloc((size_t)limbs * c->n * 8);
    for (int l = 0; l < limbs; l++)
        for (size_t i = 0; i < c->n; i++)
            pt[(size_t)l * c->n + i] = k[l];
    int32_t rc = plainop(c, mul ? 2 : 0, ct, pt, out, batch, size, limbs, 0);
    free(pt);
    return rc;
}
int32_t moai_multiply_scalar(moai_context *c, const uint64_t *ct, const uint64_t *k, uint64_t *out, int64_t batch, int32_t size,
                             int32_t limbs)
{
    return scalarop(c, 1, ct, k, out, batch, size, limbs);
}
int32_t moai_add_scalar(moai_context *c, const uint64_t *ct, const uint64_t *k, uint64_t *out, int64_t batch, int32_t size,
                        int32_t limbs)
{
    return scalarop(c, 0, ct, k, out, batch, size, limbs);
}
int32_t moai_encode_vector(moai_context *c, const double *values, int64_t count, int32_t n_vals, double scale, int32_t limbs,
                           uint64_t *out)
{
    REQ(c && limbs >= 1 && limbs <= c->kl && n_vals >= 0 && (size_t)n_vals <= c->n / 2, "bad shape");
    for (int64_t i = 0; i < count; i++)
        if (orc_encode_vector(c->o, values + (size_t)i * 2 * n_vals, n_vals, scale, limbs, out + (size_t)i * limbs * c->n))
            return fail(MOAI_INVALID_ARGUMENT, "encoded values are too large");
    return MOAI_OK;
}
int32_t moai_mod_raise(moai_context *c, const uint64_t *in, uint64_t *out, int64_t b, int32_t s, int32_t l)
{
    (void)c; (void)in; (void)out; (void)b; (void)s; (void)l;
    UNSUPPORTED("moai_mod_raise");
}

/* ---- keys + rotate_vector with SEAL's NAF fallback (S/evaluator.cpp:2667-2722) ---- */
int32_t moai_keys_create(moai_context *c, moai_keys **out)
{
    moai_keys *k = (moai_keys *)calloc(1, sizeof *k);
    k->c = c;
    *out = k;
    return MOAI_OK;
}
int32_t moai_keys_destroy(moai_keys *k) { free(k); return MOAI_OK; }
int32_t moai_keys_set_relin(moai_keys *k, const uint64_t *ksk) { k->relin = ksk; return MOAI_OK; }
int32_t moai_keys_add_galois(moai_keys *k, uint32_t elt, const uint64_t *ksk)
{
    REQ(k->n_gal < MAX_GAL, "too many keys for the test double");
    k->elt[k->n_gal] = elt;
    k->key[k->n_gal++] = ksk;
    return MOAI_OK;
}
static const u64 *find_key(moai_keys *k, uint32_t elt)
{
    for (int i = 0; i < k->n_gal; i++)
        if (k->elt[i] == elt)
            return k->key[i];
    return NULL;
}
static int32_t rotate_one(moai_context *c, moai_keys *k, const u64 *in, u64 *out, int limbs, int steps)
{
    size_t words = (size_t)2 * limbs * c->n;
    if (steps == 0)
    {
        memmove(out, in, words * 8);
        return MOAI_OK;
    }
    const u64 *key = find_key(k, orc_elt_from_step(c->o, steps));
    if (key)
    {
        orc_apply_galois(c->o, in, limbs, orc_elt_from_step(c->o, steps), key, out);
        return MOAI_OK;
    }
    int terms[64];
    int cnt = orc_naf_steps(c->o, steps, terms);
    if (cnt == 1 && terms[0] == steps)
        return fail(MOAI_INVALID_ARGUMENT, "Galois key not present");
    u64 *cur = (u64 *)malloc(words * 8), *nxt = (u64 *)malloc(words * 8);
    memcpy(cur, in, words * 8);
    int32_t rc = MOAI_OK;
    for (int i = 0; i < cnt && rc == MOAI_OK; i++)
    {
        rc = rotate_one(c, k, cur, nxt, limbs, terms[i]);
        u64 *t = cur;
        cur = nxt;
        nxt = t;
    }
    if (rc == MOAI_OK)
        memcpy(out, cur, words * 8);
    free(cur);
    free(nxt);
    return rc;
}
int32_t moai_rotate_vector(moai_context *c, moai_keys *k, const uint64_t *in, uint64_t *out, int64_t batch, int32_t limbs,
                           int32_t steps)
{
    REQ(shape_ok(c, batch, 2, limbs) && k, "bad shape");
    for (int64_t i = 0; i < batch; i++)
    {
        int32_t rc = rotate_one(c, k, in + (size_t)i * 2 * limbs * c->n, out + (size_t)i * 2 * limbs * c->n, limbs, steps);
        if (rc != MOAI_OK)
            return rc;
    }
    return MOAI_OK;
}

/* ---- fused modules: the ct-pt matmuls exist in the oracle (enough to exercise the pack / unpack glue of
 * include/moai_b200_fused_modules.hpp on the CPU); the others are GPU-only ---- */
extern void orc_ct_pt_matmul_scalar(const orc_ctx *c, const u64 *X, const double *W, int K, int C, int limbs, double scale,
                                    int c_begin, int c_end, u64 *out);
extern void orc_ct_pt_matmul_masked(const orc_ctx *c, const u64 *X, const double *W, const int *mask, int K, int C, int limbs,
                                    double scale, int c_begin, int c_end, u64 *out);
int32_t moai_ct_pt_matrix_mul_wo_pre(moai_context *c, const uint64_t *X, const double *W, int32_t col_X, int32_t col_W,
                                     int32_t row_W, int32_t limbs, double scale, uint64_t *out)
{
    REQ(c && col_X == row_W && limbs >= 2 && limbs <= c->kl, "bad dimensions of X or W");
    orc_ct_pt_matmul_scalar(c->o, X, W, row_W, col_W, limbs, scale, 0, col_W, out);
    return MOAI_OK;
}
int32_t moai_ct_pt_matrix_mul_wo_pre_w_mask(moai_context *c, const uint64_t *X, const double *W, const int32_t *mask,
                                            int32_t col_X, int32_t col_W, int32_t row_W, int32_t limbs, double scale,
                                            uint64_t *out)
{
    REQ(c && col_X == row_W && limbs >= 2 && limbs <= c->kl && mask, "bad dimensions of X or W");
    orc_ct_pt_matmul_masked(c->o, X, W, mask, row_W, col_W, limbs, scale, 0, col_W, out);
    return MOAI_OK;
}
int32_t moai_ct_pt_matrix_mul_wo_pre_host(moai_context *c, const uint64_t *X, const double *W, int32_t a, int32_t b, int32_t d,
                                          int32_t l, double s, uint64_t *out)
{
    (void)c; (void)X; (void)W; (void)a; (void)b; (void)d; (void)l; (void)s; (void)out;
    UNSUPPORTED("moai_ct_pt_matrix_mul_wo_pre_host");
}
#define GPU_ONLY_MODULE(name, ...)                                                                                     \
    int32_t name(__VA_ARGS__) { UNSUPPORTED(#name); }
GPU_ONLY_MODULE(moai_gelu_v2, moai_context *c, moai_keys *k, const uint64_t *x, int64_t b, int32_t l, double s, uint64_t *o,
                int32_t *ol, double *os)
GPU_ONLY_MODULE(moai_layernorm, moai_context *c, moai_keys *k, const uint64_t *x, int32_t n, int32_t l, double s,
                const double *g, const double *be, const int32_t *bv, int32_t v, uint64_t *o, int32_t *ol, double *os)
GPU_ONLY_MODULE(moai_exp, moai_context *c, moai_keys *k, const uint64_t *x, int64_t b, int32_t l, double s, uint64_t *o,
                int32_t *ol, double *os)
GPU_ONLY_MODULE(moai_inverse, moai_context *c, moai_keys *k, const uint64_t *x, int64_t b, int32_t l, double s, int32_t it,
                uint64_t *o, int32_t *ol, double *os)
GPU_ONLY_MODULE(moai_ct_ct_matrix_mul_colpacking, moai_context *c, moai_keys *k, const uint64_t *x, const uint64_t *w,
                int32_t l, double sx, double sw, int32_t a, int32_t b, int32_t d, int32_t e, int32_t nb, uint64_t *o,
                int32_t *ol, double *os)
GPU_ONLY_MODULE(moai_ct_ct_matrix_mul_diagpacking, moai_context *c, moai_keys *k, const uint64_t *x, const uint64_t *w,
                int32_t l, double sx, double sw, int32_t a, int32_t b, int32_t d, int32_t e, int32_t nb, uint64_t *o,
                int32_t *ol, double *os)
GPU_ONLY_MODULE(moai_softmax_boot, moai_context *c, moai_keys *k, moai_bootstrapper *bt, const uint64_t *x, int32_t n,
                int32_t l, double s, const int32_t *bv, int32_t in, int32_t it, int32_t li, uint64_t *o, int32_t *ol,
                double *os)
GPU_ONLY_MODULE(moai_single_att_block, moai_context *c, moai_keys *k, moai_bootstrapper *bt, const uint64_t *x, int32_t nc,
                int32_t l, double s, const double *wq, const double *wk, const double *wv, const double *bq,
                const double *bk, const double *bv, int32_t cw, const int32_t *mask, int32_t in, int32_t nb, int32_t it,
                int32_t li, uint64_t *o, int32_t *ol, double *os)

/* ---- not needed by the CPU host-logic tests ---- */
int32_t moai_expand_seeds(moai_context *c, const uint64_t *s, int64_t n, int32_t l, uint64_t *o, int64_t st)
{
    (void)c; (void)s; (void)n; (void)l; (void)o; (void)st;
    UNSUPPORTED("moai_expand_seeds"); /* the facade expands seeds on the host when bound to the test double */
}
int32_t moai_key_prepare(moai_context *c, const uint64_t *a, uint32_t e, int32_t m, int32_t p, uint64_t *o)
{
    (void)c; (void)a; (void)e; (void)m; (void)p; (void)o;
    UNSUPPORTED("moai_key_prepare");
}
int32_t moai_keys_add_galois_fast(moai_keys *k, uint32_t e, const uint64_t *p, int32_t l)
{
    (void)k; (void)e; (void)p; (void)l;
    UNSUPPORTED("moai_keys_add_galois_fast");
}
/* ---- bootstrapper: a PLUMBING FAKE, not a bootstrapping.  It lets the CPU tests exercise the host logic around
 * moai_bootstrap / moai_bootstrap_real (the facade Bootstrapper's request combining: which ciphertext goes into which
 * batch and comes back to which caller).  The "result" of ciphertext b is a fixed function of ITS OWN input only:
 *   out[b][p][l][i] = in[b][p][0][i] mod q_l,  at total_limbs - 14 limbs with scale final_scale.                    */
struct moai_bootstrapper
{
    int total_limbs;
    double final_scale;
};
static int g_boot_calls = 0;
int moai_mock_bootstrap_calls(void) { return g_boot_calls; }
int32_t moai_bootstrapper_create(moai_context *c, int32_t total_limbs, double final_scale, int32_t K, int32_t deg, int32_t da,
                                 int32_t lw, moai_bootstrapper **out)
{
    (void)K; (void)deg; (void)da; (void)lw;
    REQ(c && out && total_limbs >= 15 && total_limbs <= c->kl - 1, "bad level budget");
    moai_bootstrapper *b = (moai_bootstrapper *)calloc(1, sizeof *b);
    b->total_limbs = total_limbs;
    b->final_scale = final_scale;
    *out = b;
    return MOAI_OK;
}
int32_t moai_bootstrapper_destroy(moai_bootstrapper *b) { free(b); return MOAI_OK; }
int32_t moai_bootstrapper_set_hoisting(moai_bootstrapper *b, int32_t on) { (void)b; (void)on; return MOAI_OK; }
int32_t moai_bootstrapper_required_steps(moai_bootstrapper *b, int32_t *s, int32_t cap, int32_t *n)
{
    (void)b;
    REQ(cap >= 2, "capacity");
    s[0] = 1;
    s[1] = 2;
    *n = 2;
    return MOAI_OK;
}
static int32_t fake_bootstrap(moai_context *c, moai_bootstrapper *b, const uint64_t *in, int64_t batch, uint64_t *out,
                              int32_t *ol, double *os)
{
    REQ(c && b && in && out && batch >= 0, "null argument");
    const int L = b->total_limbs - 14;
    for (int64_t i = 0; i < batch; i++)
        for (int p = 0; p < 2; p++)
            for (int l = 0; l < L; l++)
                for (size_t k = 0; k < c->n; k++)
                    out[(((size_t)i * 2 + p) * L + l) * c->n + k] = in[((size_t)i * 2 + p) * c->n + k] % c->q[l];
    *ol = L;
    *os = b->final_scale;
    __atomic_add_fetch(&g_boot_calls, 1, __ATOMIC_SEQ_CST);
    return MOAI_OK;
}
int32_t moai_bootstrap(moai_context *c, moai_bootstrapper *b, moai_keys *k, const uint64_t *in, int64_t batch, double scale,
                       uint64_t *out, int32_t *ol, double *os)
{
    (void)k; (void)scale;
    return fake_bootstrap(c, b, in, batch, out, ol, os);
}
int32_t moai_bootstrap_real(moai_context *c, moai_bootstrapper *b, moai_keys *k, const uint64_t *in, int64_t batch,
                            double scale, int64_t chunk_pairs, uint64_t *out, int32_t *ol, double *os)
{
    (void)k; (void)scale; (void)chunk_pairs;
    return fake_bootstrap(c, b, in, batch, out, ol, os);
}
