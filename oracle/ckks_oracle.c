/*
 * ckks_oracle.c — CPU restatement (plain C) of the reference's CKKS evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product library (libmoai_b200.so) never does and
 * has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks this file against
 *   (1) the known-answer vectors of the reference's own gtest suites
 *       (ST/util/ntt.cpp:53-101, ST/util/galois.cpp:19-115, ST/util/rns.cpp:1013-1073,
 *        ST/util/uintarithsmallmod.cpp) and
 *   (2) the reference's vendored SEAL-4.1-bs itself, compiled here into oracle/_ref
 *       (oracle/refbuild/Makefile), on identical seeded inputs, bit for bit.
 *
 * Citation shorthand: S/ = /root/reference/thirdparty/SEAL-4.1-bs/native/src/seal/,
 * M/ = /root/reference/include/.  Every function names the reference lines it restates.
 * Layouts are SEAL's: ciphertext = [poly][limb][coeff] uint64 (S/ciphertext.h:339-370),
 * key-switching key = [digit][poly 0..1][key limb][coeff] (S/kswitchkeys.h:335-340).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <complex.h>
#undef I

typedef unsigned __int128 u128;
typedef uint64_t u64;

#define ORC_MAX_LIMBS 64

typedef struct
{
    int log_n;
    size_t n;
    int n_key_limbs;          /* data primes + 1 special prime */
    u64 q[ORC_MAX_LIMBS];     /* primes, key level order */
    u64 *root_op[ORC_MAX_LIMBS], *root_quo[ORC_MAX_LIMBS];   /* psi powers, bit-reversed order */
    u64 *iroot_op[ORC_MAX_LIMBS], *iroot_quo[ORC_MAX_LIMBS]; /* psi^-1 powers, "scrambled" order */
    u64 inv_n_op[ORC_MAX_LIMBS], inv_n_quo[ORC_MAX_LIMBS];
    /* CKKS encoder tables (S/ckks.cpp:33-75) */
    size_t *index_map;
    double complex *fft_inv_roots;
    double complex *fft_roots;
} orc_ctx;

/* ---------------------------------------------------------------- modular primitives ---- */

/* canonical a*b mod q */
static inline u64 mulmod(u64 a, u64 b, u64 q)
{
    return (u64)(((u128)a * b) % q);
}

static u64 powmod(u64 a, u64 e, u64 q)
{
    u64 r = 1;
    a %= q;
    while (e)
    {
        if (e & 1)
            r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}

static u64 invmod(u64 a, u64 q) /* q prime */
{
    return powmod(a, q - 2, q);
}

/* Shoup quotient floor(operand * 2^64 / q): S/util/uintarithsmallmod.h:260-270 */
static inline u64 shoup_quotient(u64 operand, u64 q)
{
    return (u64)((((u128)operand) << 64) / q);
}

/* multiply_uint_mod_lazy: S/util/uintarithsmallmod.h:313-326; result in [0, 2q) */
static inline u64 mul_lazy(u64 x, u64 op, u64 quo, u64 q)
{
    u64 hi = (u64)(((u128)x * quo) >> 64);
    return op * x - hi * q;
}

/* Exposed for the uintarithsmallmod KATs (ST/util/uintarithsmallmod.cpp). */
u64 orc_mulmod(u64 a, u64 b, u64 q) { return mulmod(a, b, q); }
u64 orc_powmod(u64 a, u64 e, u64 q) { return powmod(a, e, q); }
u64 orc_invmod(u64 a, u64 q) { return invmod(a, q); }
u64 orc_shoup_quotient(u64 op, u64 q) { return shoup_quotient(op, q); }
u64 orc_mul_lazy(u64 x, u64 op, u64 quo, u64 q) { return mul_lazy(x, op, quo, q); }

/* barrett_reduce_128 with const_ratio = floor(2^128/q): S/util/uintarithsmallmod.h:166-204,
 * S/modulus.cpp:83-98.  Restated with the same three partial products. */
u64 orc_barrett_reduce_128(u64 lo, u64 hi, u64 q)
{
    /* const_ratio = floor(2^128 / q) as (cr1:cr0) */
    u128 top = ((u128)1 << 64) / q;                        /* floor(2^64/q) : high word */
    u128 rem = ((u128)1 << 64) % q;
    u64 cr1 = (u64)top;
    u64 cr0 = (u64)((rem << 64) / q);
    u64 carry = (u64)(((u128)lo * cr0) >> 64);
    u128 t2 = (u128)lo * cr1;
    u64 tmp1;
    u64 t2lo = (u64)t2, t2hi = (u64)(t2 >> 64);
    tmp1 = t2lo + carry;
    u64 tmp3 = t2hi + (tmp1 < t2lo);
    t2 = (u128)hi * cr0;
    t2lo = (u64)t2;
    t2hi = (u64)(t2 >> 64);
    u64 s = tmp1 + t2lo;
    carry = t2hi + (s < tmp1);
    tmp1 = hi * cr1 + tmp3 + carry;
    tmp3 = lo - tmp1 * q;
    return tmp3 >= q ? tmp3 - q : tmp3;
}

static inline u64 reverse_bits(u64 x, int bits)
{
    u64 r = 0;
    for (int i = 0; i < bits; i++)
    {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

/* ---------------------------------------------------------------- primes / roots ---- */

/* Deterministic Miller-Rabin for 64-bit (SEAL uses random bases, S/util/numth.cpp:180-277;
 * the set of primes is the same). */
static int is_prime_u64(u64 n)
{
    if (n < 2)
        return 0;
    static const u64 small[] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37 };
    for (int i = 0; i < 12; i++)
    {
        if (n == small[i])
            return 1;
        if (n % small[i] == 0)
            return 0;
    }
    u64 d = n - 1;
    int r = 0;
    while (!(d & 1))
    {
        d >>= 1;
        r++;
    }
    for (int i = 0; i < 12; i++)
    {
        u64 x = powmod(small[i], d, n);
        if (x == 1 || x == n - 1)
            continue;
        int ok = 0;
        for (int j = 1; j < r; j++)
        {
            x = mulmod(x, x, n);
            if (x == n - 1)
            {
                ok = 1;
                break;
            }
        }
        if (!ok)
            return 0;
    }
    return 1;
}

/* CoeffModulus::Create: S/modulus.cpp:143-184 + get_primes S/util/numth.cpp:278-311.
 * For each bit size the primes are found downward from 2^bits in steps of 2N; positions are
 * served from the BACK of that list (smallest first). */
int orc_create_primes(int log_n, const int *bits, int n_bits, u64 *out)
{
    u64 factor = (u64)2 << log_n;
    for (int i = 0; i < n_bits; i++)
        out[i] = 0;
    for (int i = 0; i < n_bits; i++)
    {
        if (out[i])
            continue;
        int bs = bits[i];
        int count = 0;
        for (int j = i; j < n_bits; j++)
            if (bits[j] == bs)
                count++;
        u64 *found = (u64 *)malloc(sizeof(u64) * (size_t)count);
        u64 value = ((((u64)1) << bs) - 1) / factor * factor + 1;
        u64 lower = ((u64)1) << (bs - 1);
        int k = 0;
        while (k < count && value > lower)
        {
            if (is_prime_u64(value))
                found[k++] = value;
            value -= factor;
        }
        if (k < count)
        {
            free(found);
            return -1;
        }
        int back = count - 1;
        for (int j = i; j < n_bits; j++)
            if (bits[j] == bs)
                out[j] = found[back--];
        free(found);
    }
    return 0;
}

/* try_minimal_primitive_root: S/util/numth.cpp:385-413 — the numerically smallest primitive
 * 2N-th root of unity.  Any primitive root generates the same set of odd powers. */
u64 orc_minimal_primitive_root(u64 degree, u64 q)
{
    u64 quot = (q - 1) / degree;
    u64 g = 0;
    for (u64 a = 2;; a++)
    {
        g = powmod(a, quot, q);
        if (powmod(g, degree >> 1, q) == q - 1)
            break;
    }
    u64 gsq = mulmod(g, g, q);
    u64 cur = g, best = g;
    for (u64 i = 0; i < degree; i += 2)
    {
        if (cur < best)
            best = cur;
        cur = mulmod(cur, gsq, q);
    }
    return best;
}

/* ---------------------------------------------------------------- context ---- */

/* ComplexRoots::get_root: S/util/croots.cpp:17-72 (octant symmetry of the 2N-th roots). */
static double complex croot(const double complex *tab, size_t m, size_t index)
{
    index &= m - 1;
    if (index <= m / 8)
        return tab[index];
    else if (index <= m / 4)
    {
        double complex a = tab[m / 4 - index];
        return CMPLX(cimag(a), creal(a));
    }
    else if (index <= m / 2)
        return -conj(croot(tab, m, m / 2 - index));
    else if (index <= 3 * m / 4)
        return -croot(tab, m, index - m / 2);
    else
        return conj(croot(tab, m, m - index));
}

void orc_destroy(orc_ctx *c);

/* NTTTables::initialize: S/util/ntt.cpp:241-300; CKKSEncoder ctor: S/ckks.cpp:20-75 */
orc_ctx *orc_create_from_primes(int log_n, const u64 *primes, int n_bits)
{
    if (n_bits > ORC_MAX_LIMBS)
        return NULL;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->log_n = log_n;
    c->n = (size_t)1 << log_n;
    c->n_key_limbs = n_bits;
    memcpy(c->q, primes, 8 * (size_t)n_bits);
    size_t n = c->n;
    for (int l = 0; l < n_bits; l++)
    {
        u64 q = c->q[l];
        u64 psi = orc_minimal_primitive_root(2 * n, q);
        u64 ipsi = invmod(psi, q);
        c->root_op[l] = (u64 *)malloc(n * 8);
        c->root_quo[l] = (u64 *)malloc(n * 8);
        c->iroot_op[l] = (u64 *)malloc(n * 8);
        c->iroot_quo[l] = (u64 *)malloc(n * 8);
        u64 power = psi;
        for (size_t i = 1; i < n; i++)
        {
            size_t r = reverse_bits(i, log_n);
            c->root_op[l][r] = power;
            c->root_quo[l][r] = shoup_quotient(power, q);
            power = mulmod(power, psi, q);
        }
        c->root_op[l][0] = 1;
        c->root_quo[l][0] = shoup_quotient(1, q);
        power = ipsi;
        for (size_t i = 1; i < n; i++)
        {
            size_t r = reverse_bits(i - 1, log_n) + 1;
            c->iroot_op[l][r] = power;
            c->iroot_quo[l][r] = shoup_quotient(power, q);
            power = mulmod(power, ipsi, q);
        }
        c->iroot_op[l][0] = 1;
        c->iroot_quo[l][0] = shoup_quotient(1, q);
        c->inv_n_op[l] = invmod((u64)n % q, q);
        c->inv_n_quo[l] = shoup_quotient(c->inv_n_op[l], q);
    }
    if (n < 4)
        return c;
    /* encoder tables */
    size_t slots = n >> 1;
    u64 m = (u64)n << 1;
    c->index_map = (size_t *)malloc(sizeof(size_t) * n);
    u64 pos = 1;
    for (size_t i = 0; i < slots; i++)
    {
        u64 index1 = (pos - 1) >> 1;
        u64 index2 = (m - pos - 1) >> 1;
        c->index_map[i] = reverse_bits(index1, log_n);
        c->index_map[slots | i] = reverse_bits(index2, log_n);
        pos = (pos * 5) & (m - 1);
    }
    double complex *tab = (double complex *)malloc(sizeof(double complex) * (m / 8 + 1));
    const double PI = 3.1415926535897932384626433832795028842;
    for (size_t i = 0; i <= m / 8; i++)
    {
        double ang = 2 * PI * (double)i / (double)m;
        tab[i] = CMPLX(cos(ang), sin(ang));
    }
    c->fft_roots = (double complex *)calloc(n, sizeof(double complex));
    c->fft_inv_roots = (double complex *)calloc(n, sizeof(double complex));
    for (size_t i = 1; i < n; i++)
    {
        c->fft_roots[i] = croot(tab, m, reverse_bits(i, log_n));
        c->fft_inv_roots[i] = conj(croot(tab, m, reverse_bits(i - 1, log_n) + 1));
    }
    free(tab);
    return c;
}

orc_ctx *orc_create(int log_n, const int *bits, int n_bits)
{
    u64 primes[ORC_MAX_LIMBS];
    if (n_bits > ORC_MAX_LIMBS || orc_create_primes(log_n, bits, n_bits, primes))
        return NULL;
    return orc_create_from_primes(log_n, primes, n_bits);
}

void orc_destroy(orc_ctx *c)
{
    if (!c)
        return;
    for (int l = 0; l < c->n_key_limbs; l++)
    {
        free(c->root_op[l]);
        free(c->root_quo[l]);
        free(c->iroot_op[l]);
        free(c->iroot_quo[l]);
    }
    free(c->index_map);
    free(c->fft_roots);
    free(c->fft_inv_roots);
    free(c);
}

int orc_n_key_limbs(const orc_ctx *c) { return c->n_key_limbs; }
void orc_primes(const orc_ctx *c, u64 *out) { memcpy(out, c->q, 8 * (size_t)c->n_key_limbs); }

void orc_ntt_tables(const orc_ctx *c, int limb, u64 *root_op, u64 *root_quo, u64 *iroot_op, u64 *iroot_quo,
                    u64 *inv_n_op_quo)
{
    memcpy(root_op, c->root_op[limb], 8 * c->n);
    memcpy(root_quo, c->root_quo[limb], 8 * c->n);
    memcpy(iroot_op, c->iroot_op[limb], 8 * c->n);
    memcpy(iroot_quo, c->iroot_quo[limb], 8 * c->n);
    inv_n_op_quo[0] = c->inv_n_op[limb];
    inv_n_op_quo[1] = c->inv_n_quo[limb];
}

void orc_fft_tables(const orc_ctx *c, double *roots_reim, double *inv_roots_reim, u64 *index_map)
{
    for (size_t i = 0; i < c->n; i++)
    {
        roots_reim[2 * i] = creal(c->fft_roots[i]);
        roots_reim[2 * i + 1] = cimag(c->fft_roots[i]);
        inv_roots_reim[2 * i] = creal(c->fft_inv_roots[i]);
        inv_roots_reim[2 * i + 1] = cimag(c->fft_inv_roots[i]);
        index_map[i] = c->index_map[i];
    }
}

/* ---------------------------------------------------------------- NTT ---- */

/* Forward negacyclic NTT, lazy: DWTHandler::transform_to_rev S/util/dwthandler.h:94-191 with
 * Arithmetic<> of S/util/ntt.h:30-61.  Input < 4q, output < 4q, natural -> bit-reversed. */
static void ntt_lazy(const orc_ctx *c, int limb, u64 *v)
{
    size_t n = c->n;
    u64 q = c->q[limb], two_q = q << 1;
    const u64 *rop = c->root_op[limb], *rquo = c->root_quo[limb];
    size_t gap = n >> 1, m = 1, ridx = 0;
    for (; m < n; m <<= 1)
    {
        size_t offset = 0;
        for (size_t i = 0; i < m; i++)
        {
            ++ridx;
            u64 *x = v + offset, *y = x + gap;
            for (size_t j = 0; j < gap; j++)
            {
                u64 u = x[j] >= two_q ? x[j] - two_q : x[j];
                u64 w = mul_lazy(y[j], rop[ridx], rquo[ridx], q);
                x[j] = u + w;
                y[j] = u + two_q - w;
            }
            offset += gap << 1;
        }
        gap >>= 1;
    }
}

/* ntt_negacyclic_harvey: S/util/ntt.cpp:408-437 — lazy transform then [0,4q) -> [0,q). */
void orc_ntt(const orc_ctx *c, int limb, u64 *v)
{
    ntt_lazy(c, limb, v);
    u64 q = c->q[limb], two_q = q << 1;
    for (size_t i = 0; i < c->n; i++)
    {
        if (v[i] >= two_q)
            v[i] -= two_q;
        if (v[i] >= q)
            v[i] -= q;
    }
}

/* Inverse NTT, lazy: transform_from_rev S/util/dwthandler.h:202-356 with scalar = N^-1 folded
 * into the last stage.  Input < 2q, output < 2q, bit-reversed -> natural. */
static void intt_lazy(const orc_ctx *c, int limb, u64 *v)
{
    size_t n = c->n;
    u64 q = c->q[limb], two_q = q << 1;
    const u64 *rop = c->iroot_op[limb], *rquo = c->iroot_quo[limb];
    size_t gap = 1, m = n >> 1, ridx = 0;
    for (; m > 1; m >>= 1)
    {
        size_t offset = 0;
        for (size_t i = 0; i < m; i++)
        {
            ++ridx;
            u64 *x = v + offset, *y = x + gap;
            for (size_t j = 0; j < gap; j++)
            {
                u64 u = x[j], w = y[j];
                u64 s = u + w;
                x[j] = s >= two_q ? s - two_q : s;
                y[j] = mul_lazy(u + two_q - w, rop[ridx], rquo[ridx], q);
            }
            offset += gap << 1;
        }
        gap <<= 1;
    }
    /* last stage, scaled by N^-1 (dwthandler.h:262-312) */
    ++ridx;
    u64 s_op = c->inv_n_op[limb], s_quo = c->inv_n_quo[limb];
    u64 sr_op = mulmod(rop[ridx], s_op, q);
    u64 sr_quo = shoup_quotient(sr_op, q);
    u64 *x = v, *y = v + gap;
    for (size_t j = 0; j < gap; j++)
    {
        u64 u = x[j] >= two_q ? x[j] - two_q : x[j];
        u64 w = y[j];
        u64 s = u + w;
        s = s >= two_q ? s - two_q : s;
        x[j] = mul_lazy(s, s_op, s_quo, q);
        y[j] = mul_lazy(u + two_q - w, sr_op, sr_quo, q);
    }
}

/* inverse_ntt_negacyclic_harvey: S/util/ntt.cpp:454-475 */
void orc_intt(const orc_ctx *c, int limb, u64 *v)
{
    intt_lazy(c, limb, v);
    u64 q = c->q[limb];
    for (size_t i = 0; i < c->n; i++)
        if (v[i] >= q)
            v[i] -= q;
}

/* ---------------------------------------------------------------- element-wise ---- */

/* add/sub/negate_poly_coeffmod: S/util/polyarithsmallmod.h:77-300; Evaluator::add_inplace
 * S/evaluator.cpp:155-240 etc.  op: 0 add, 1 sub, 2 negate(a).  polys*limbs*n elements. */
void orc_addsub(const orc_ctx *c, int op, const u64 *a, const u64 *b, int polys, int limbs, u64 *out)
{
    size_t n = c->n;
    for (int p = 0; p < polys; p++)
        for (int l = 0; l < limbs; l++)
        {
            u64 q = c->q[l];
            size_t base = ((size_t)p * limbs + l) * n;
            for (size_t i = 0; i < n; i++)
            {
                u64 x = a[base + i];
                if (op == 2)
                {
                    out[base + i] = x ? q - x : 0;
                    continue;
                }
                u64 y = b[base + i];
                if (op == 0)
                {
                    u64 s = x + y;
                    out[base + i] = s >= q ? s - q : s;
                }
                else
                {
                    out[base + i] = x >= y ? x - y : x + q - y;
                }
            }
        }
}

/* add_plain / sub_plain: S/evaluator.cpp:1938-2152 — the plaintext touches poly 0 only. */
void orc_addsub_plain(const orc_ctx *c, int op, const u64 *ct, const u64 *pt, int polys, int limbs, u64 *out)
{
    size_t n = c->n;
    if (out != ct)
        memcpy(out, ct, (size_t)polys * limbs * n * 8);
    orc_addsub(c, op, ct, pt, 1, limbs, out);
}

/* multiply_plain_ntt: S/evaluator.cpp:2336-2373 — every poly of ct times pt, per limb. */
void orc_multiply_plain(const orc_ctx *c, const u64 *ct, const u64 *pt, int polys, int limbs, u64 *out)
{
    size_t n = c->n;
    for (int p = 0; p < polys; p++)
        for (int l = 0; l < limbs; l++)
        {
            u64 q = c->q[l];
            size_t base = ((size_t)p * limbs + l) * n;
            const u64 *m = pt + (size_t)l * n;
            for (size_t i = 0; i < n; i++)
                out[base + i] = mulmod(ct[base + i], m[i], q);
        }
}

/* ckks_multiply for two size-2 inputs: S/evaluator.cpp:835-863 -> (c0d0, c0d1+c1d0, c1d1). */
void orc_multiply(const orc_ctx *c, const u64 *a, const u64 *b, int limbs, u64 *out)
{
    size_t n = c->n;
    for (int l = 0; l < limbs; l++)
    {
        u64 q = c->q[l];
        const u64 *a0 = a + (size_t)l * n, *a1 = a + ((size_t)limbs + l) * n;
        const u64 *b0 = b + (size_t)l * n, *b1 = b + ((size_t)limbs + l) * n;
        u64 *o0 = out + (size_t)l * n, *o1 = out + ((size_t)limbs + l) * n, *o2 = out + ((size_t)2 * limbs + l) * n;
        for (size_t i = 0; i < n; i++)
        {
            u64 x0 = a0[i], x1 = a1[i], y0 = b0[i], y1 = b1[i];
            u64 t = mulmod(x0, y1, q) + mulmod(x1, y0, q);
            o0[i] = mulmod(x0, y0, q);
            o1[i] = t >= q ? t - q : t;
            o2[i] = mulmod(x1, y1, q);
        }
    }
}

/* ckks_square: S/evaluator.cpp:1263-1277 -> (c0^2, 2 c0 c1, c1^2). */
void orc_square(const orc_ctx *c, const u64 *a, int limbs, u64 *out)
{
    orc_multiply(c, a, a, limbs, out);
}

/* ---------------------------------------------------------------- rescale / modswitch ---- */

/* RNSTool::divide_and_round_q_last_ntt_inplace: S/util/rns.cpp:830-901, applied to every poly
 * (mod_switch_scale_to_next S/evaluator.cpp:1402-1481).  in: [polys][limbs][n] -> out
 * [polys][limbs-1][n].  Destroys nothing (works on a copy). */
void orc_rescale(const orc_ctx *c, const u64 *in, int polys, int limbs, u64 *out)
{
    size_t n = c->n;
    int last = limbs - 1;
    u64 ql = c->q[last], half = ql >> 1;
    u64 *t = (u64 *)malloc(n * 8), *tmp = (u64 *)malloc(n * 8);
    for (int p = 0; p < polys; p++)
    {
        memcpy(t, in + ((size_t)p * limbs + last) * n, n * 8);
        orc_intt(c, last, t);
        for (size_t i = 0; i < n; i++)
        {
            u64 s = t[i] + half;
            t[i] = s >= ql ? s - ql : s;
        }
        for (int l = 0; l < last; l++)
        {
            u64 q = c->q[l];
            u64 neg_half_mod = q - (half % q);
            for (size_t i = 0; i < n; i++)
                tmp[i] = (q < ql ? t[i] % q : t[i]) + neg_half_mod;
            ntt_lazy(c, l, tmp);
            u64 qi_lazy = q << 2;
            u64 inv = invmod(ql % q, q);
            const u64 *src = in + ((size_t)p * limbs + l) * n;
            u64 *dst = out + ((size_t)p * last + l) * n;
            for (size_t i = 0; i < n; i++)
                dst[i] = mulmod((src[i] + qi_lazy - tmp[i]) % q, inv, q);
        }
    }
    free(t);
    free(tmp);
}

/* mod_switch_drop_to_next (CKKS): S/evaluator.cpp:1483-1546 — drop the last limb of every poly. */
void orc_mod_switch(const orc_ctx *c, const u64 *in, int polys, int limbs, u64 *out)
{
    size_t n = c->n;
    for (int p = 0; p < polys; p++)
        for (int l = 0; l < limbs - 1; l++)
            memmove(out + ((size_t)p * (limbs - 1) + l) * n, in + ((size_t)p * limbs + l) * n, n * 8);
}

/* ---------------------------------------------------------------- Galois ---- */

/* GaloisTool::get_elt_from_step: S/util/galois.cpp:53-95 (generator 5; step 0 = conjugation). */
uint32_t orc_elt_from_step(const orc_ctx *c, int step)
{
    uint32_t n = (uint32_t)c->n;
    u64 m = (u64)n << 1;
    if (step == 0)
        return (uint32_t)(m - 1);
    int neg = step < 0;
    uint32_t pos = (uint32_t)(neg ? -step : step);
    int s = neg ? (int)(n >> 1) - (int)pos : (int)pos;
    u64 e = 1;
    while (s--)
        e = (e * 5) & (m - 1);
    return (uint32_t)e;
}

/* GaloisTool::generate_table_ntt: S/util/galois.cpp:18-51 */
void orc_galois_table(const orc_ctx *c, uint32_t elt, uint32_t *table)
{
    size_t n = c->n;
    for (size_t i = n; i < 2 * n; i++)
    {
        u64 rev = reverse_bits(i, c->log_n + 1);
        u64 idx = (((u64)elt * rev) >> 1) & (n - 1);
        table[i - n] = (uint32_t)reverse_bits(idx, c->log_n);
    }
}

/* apply_galois_ntt on one limb: S/util/galois.cpp:192-218 */
static void apply_galois_limb(const uint32_t *table, size_t n, const u64 *in, u64 *out)
{
    for (size_t i = 0; i < n; i++)
        out[i] = in[table[i]];
}

/* ---------------------------------------------------------------- key switching ---- */

/* Evaluator::switch_key_inplace (CKKS branch): S/evaluator.cpp:2724-3021.
 *   ct     : [2][limbs][n], updated in place (ct += keyswitch(target))
 *   target : [limbs][n] NTT form
 *   key    : [limbs.. digits][2][n_key_limbs][n]
 */
void orc_switch_key(const orc_ctx *c, u64 *ct, const u64 *target, int limbs, const u64 *key)
{
    size_t n = c->n;
    int kl = c->n_key_limbs;
    int rns = limbs + 1;
    u64 *d = (u64 *)malloc((size_t)limbs * n * 8);          /* coefficient-form digits */
    u64 *prod = (u64 *)malloc((size_t)2 * rns * n * 8);     /* t_poly_prod [k][I][n] */
    u64 *op = (u64 *)malloc(n * 8);
    u128 *acc = (u128 *)malloc((size_t)2 * n * sizeof(u128));
    memcpy(d, target, (size_t)limbs * n * 8);
    for (int j = 0; j < limbs; j++)
        orc_intt(c, j, d + (size_t)j * n);
    for (int I = 0; I < rns; I++)
    {
        int ki = (I == limbs) ? kl - 1 : I;
        u64 m = c->q[ki];
        memset(acc, 0, (size_t)2 * n * sizeof(u128));
        for (int J = 0; J < limbs; J++)
        {
            const u64 *operand;
            if (I == J)
                operand = target + (size_t)J * n;
            else
            {
                const u64 *dj = d + (size_t)J * n;
                if (c->q[J] <= m)
                    memcpy(op, dj, n * 8);
                else
                    for (size_t i = 0; i < n; i++)
                        op[i] = dj[i] % m;
                ntt_lazy(c, ki, op);
                operand = op;
            }
            for (int k = 0; k < 2; k++)
            {
                const u64 *kp = key + (((size_t)J * 2 + k) * kl + ki) * n;
                u128 *a = acc + (size_t)k * n;
                for (size_t i = 0; i < n; i++)
                {
                    /* SEAL reduces every 256 summands; reducing always is the same function */
                    a[i] = (a[i] + (u128)operand[i] * kp[i]) % m;
                }
            }
        }
        for (int k = 0; k < 2; k++)
            for (size_t i = 0; i < n; i++)
                prod[((size_t)k * rns + I) * n + i] = (u64)acc[(size_t)k * n + i];
    }
    /* mod-down by the special prime (evaluator.cpp:2962-3018) */
    u64 qk = c->q[kl - 1], qk_half = qk >> 1;
    for (int k = 0; k < 2; k++)
    {
        u64 *t_last = prod + ((size_t)k * rns + limbs) * n;
        intt_lazy(c, kl - 1, t_last);
        for (size_t i = 0; i < n; i++)
            t_last[i] = (t_last[i] + qk_half) % qk;
        for (int l = 0; l < limbs; l++)
        {
            u64 qi = c->q[l];
            u64 fix = qi - (qk_half % qi);
            for (size_t i = 0; i < n; i++)
                op[i] = (qk > qi ? t_last[i] % qi : t_last[i]) + fix;
            ntt_lazy(c, l, op);
            u64 qi_lazy = qi << 2;
            u64 inv = invmod(qk % qi, qi);
            u64 *pl = prod + ((size_t)k * rns + l) * n;
            u64 *dst = ct + ((size_t)k * limbs + l) * n;
            for (size_t i = 0; i < n; i++)
            {
                u64 v = mulmod((pl[i] + qi_lazy - op[i]) % qi, inv, qi);
                u64 s = dst[i] + v;
                dst[i] = s >= qi ? s - qi : s;
            }
        }
    }
    free(d);
    free(prod);
    free(op);
    free(acc);
}

/* relinearize_internal (size 3 -> 2): S/evaluator.cpp:1345-1400 */
void orc_relinearize(const orc_ctx *c, const u64 *ct3, int limbs, const u64 *relin_key, u64 *out2)
{
    size_t n = c->n;
    memcpy(out2, ct3, (size_t)2 * limbs * n * 8);
    orc_switch_key(c, out2, ct3 + (size_t)2 * limbs * n, limbs, relin_key);
}

/* apply_galois_inplace (CKKS): S/evaluator.cpp:2563-2665 — c0 <- sigma(c0), target <- sigma(c1),
 * c1 <- 0, then key switch with the Galois key of `elt`. */
void orc_apply_galois(const orc_ctx *c, const u64 *ct, int limbs, uint32_t elt, const u64 *gal_key, u64 *out)
{
    size_t n = c->n;
    uint32_t *table = (uint32_t *)malloc(n * 4);
    u64 *target = (u64 *)malloc((size_t)limbs * n * 8);
    orc_galois_table(c, elt, table);
    for (int l = 0; l < limbs; l++)
    {
        apply_galois_limb(table, n, ct + (size_t)l * n, out + (size_t)l * n);
        apply_galois_limb(table, n, ct + ((size_t)limbs + l) * n, target + (size_t)l * n);
    }
    memset(out + (size_t)limbs * n, 0, (size_t)limbs * n * 8);
    orc_switch_key(c, out, target, limbs, gal_key);
    free(table);
    free(target);
}

/* naf(): S/util/numth.h:22-42, and the loop of rotate_internal S/evaluator.cpp:2699-2721
 * that turns a rotation whose key is missing into power-of-two rotations.  Writes the
 * non-zero NAF terms (with sign) in the order SEAL applies them, skipping |term| == n/2.
 * Returns the number of terms. */
int orc_naf_steps(const orc_ctx *c, int steps, int *out)
{
    int res[64], cnt = 0;
    int sign = steps < 0;
    int value = sign ? -steps : steps;
    for (int i = 0; value; i++)
    {
        int zi = (value & 1) ? 2 - (value & 3) : 0;
        value = (value - zi) >> 1;
        if (zi)
            res[cnt++] = (sign ? -zi : zi) * (1 << i);
    }
    int k = 0;
    for (int i = 0; i < cnt; i++)
    {
        int a = res[i] < 0 ? -res[i] : res[i];
        if ((size_t)a != (c->n >> 1))
            out[k++] = res[i];
    }
    return k;
}

/* ---------------------------------------------------------------- encoder ---- */

/* scalar encode: S/ckks.cpp:77-216 (<=64-bit and <=128-bit paths) -> [limbs] constants */
int orc_encode_scalar_consts(const orc_ctx *c, double value, double scale, int limbs, u64 *consts)
{
    value *= scale;
    int coeff_bit_count = (int)(log2(fabs(value))) + 2;
    double coeffd = round(value);
    int is_neg = signbit(coeffd);
    coeffd = fabs(coeffd);
    if (coeff_bit_count <= 64)
    {
        u64 cu = (u64)coeffd;
        for (int j = 0; j < limbs; j++)
        {
            u64 r = cu % c->q[j];
            consts[j] = is_neg ? (r ? c->q[j] - r : 0) : r;
        }
    }
    else if (coeff_bit_count <= 128)
    {
        double two64 = pow(2.0, 64);
        u128 cu = ((u128)(u64)(coeffd / two64) << 64) | (u64)fmod(coeffd, two64);
        for (int j = 0; j < limbs; j++)
        {
            u64 r = (u64)(cu % c->q[j]);
            consts[j] = is_neg ? (r ? c->q[j] - r : 0) : r;
        }
    }
    else
        return -1;
    return 0;
}

void orc_encode_scalar(const orc_ctx *c, double value, double scale, int limbs, u64 *out)
{
    u64 consts[ORC_MAX_LIMBS];
    orc_encode_scalar_consts(c, value, scale, limbs, consts);
    for (int j = 0; j < limbs; j++)
        for (size_t i = 0; i < c->n; i++)
            out[(size_t)j * c->n + i] = consts[j];
}

/* complex inverse DWT with scalar: FFTHandler::transform_from_rev S/util/dwthandler.h:202-356
 * instantiated with complex<double> (S/ckks.h:39-87: add, sub, mul_root = complex product,
 * mul_scalar = complex * double, guard = identity). */
static void fft_from_rev(const orc_ctx *c, double complex *v, double fix)
{
    size_t n = c->n;
    const double complex *roots = c->fft_inv_roots;
    size_t gap = 1, m = n >> 1, ridx = 0;
    for (; m > 1; m >>= 1)
    {
        size_t offset = 0;
        for (size_t i = 0; i < m; i++)
        {
            double complex r = roots[++ridx];
            double complex *x = v + offset, *y = x + gap;
            for (size_t j = 0; j < gap; j++)
            {
                double complex u = x[j], w = y[j];
                x[j] = u + w;
                y[j] = (u - w) * r;
            }
            offset += gap << 1;
        }
        gap <<= 1;
    }
    double complex r = roots[++ridx];
    double complex sr = r * fix;
    double complex *x = v, *y = v + gap;
    for (size_t j = 0; j < gap; j++)
    {
        double complex u = x[j], w = y[j];
        x[j] = (u + w) * fix;
        y[j] = (u - w) * sr;
    }
}

/* forward complex DWT (decode): transform_to_rev S/util/dwthandler.h:94-191 */
static void fft_to_rev(const orc_ctx *c, double complex *v)
{
    size_t n = c->n;
    const double complex *roots = c->fft_roots;
    size_t gap = n >> 1, m = 1, ridx = 0;
    for (; m < n; m <<= 1)
    {
        size_t offset = 0;
        for (size_t i = 0; i < m; i++)
        {
            double complex r = roots[++ridx];
            double complex *x = v + offset, *y = x + gap;
            for (size_t j = 0; j < gap; j++)
            {
                double complex u = x[j], w = y[j] * r;
                x[j] = u + w;
                y[j] = u - w;
            }
            offset += gap << 1;
        }
        gap >>= 1;
    }
}

/* vector encode: S/ckks.h:457-638.  values: n_vals complex numbers (interleaved re,im),
 * n_vals <= n/2.  Output [limbs][n] NTT form.  Returns -1 if coefficients exceed 128 bits. */
int orc_encode_vector(const orc_ctx *c, const double *values, int n_vals, double scale, int limbs, u64 *out)
{
    size_t n = c->n, slots = n >> 1;
    double complex *cv = (double complex *)calloc(n, sizeof(double complex));
    for (int i = 0; i < n_vals; i++)
    {
        double complex z = CMPLX(values[2 * i], values[2 * i + 1]);
        cv[c->index_map[i]] = z;
        cv[c->index_map[(size_t)i + slots]] = conj(z);
    }
    double fix = scale / (double)n;
    fft_from_rev(c, cv, fix);
    double max_coeff = 0;
    for (size_t i = 0; i < n; i++)
        max_coeff = fmax(max_coeff, fabs(creal(cv[i])));
    int bits = (int)ceil(log2(fmax(max_coeff, 1.0))) + 1;
    double two64 = pow(2.0, 64);
    int rc = 0;
    if (bits > 128)
        rc = -1;
    else
        for (size_t i = 0; i < n; i++)
        {
            double coeffd = round(creal(cv[i]));
            int is_neg = signbit(coeffd);
            coeffd = fabs(coeffd);
            u128 cu = bits <= 64 ? (u128)(u64)coeffd : (((u128)(u64)(coeffd / two64) << 64) | (u64)fmod(coeffd, two64));
            for (int j = 0; j < limbs; j++)
            {
                u64 r = (u64)(cu % c->q[j]);
                out[(size_t)j * n + i] = is_neg ? (r ? c->q[j] - r : 0) : r;
            }
        }
    free(cv);
    if (rc)
        return rc;
    for (int j = 0; j < limbs; j++)
        orc_ntt(c, j, out + (size_t)j * n);
    return 0;
}

/* decode: S/ckks.h:644-760, restated with an exact CRT lift in long double arithmetic for up
 * to 2 limbs-equivalent precision: centred residue composed through Garner's algorithm into a
 * long double.  Used only to read results back in tests (tolerance checks). */
void orc_decode(const orc_ctx *c, const u64 *pt, int limbs, double scale, double *out_reim)
{
    size_t n = c->n, slots = n >> 1;
    u64 *cp = (u64 *)malloc((size_t)limbs * n * 8);
    memcpy(cp, pt, (size_t)limbs * n * 8);
    for (int j = 0; j < limbs; j++)
        orc_intt(c, j, cp + (size_t)j * n);
    double complex *cv = (double complex *)malloc(n * sizeof(double complex));
    /* Garner mixed-radix: x = v0 + v1 q0 + v2 q0 q1 + ... ; centre by comparing with Q/2 */
    u64 inv[ORC_MAX_LIMBS][ORC_MAX_LIMBS];
    for (int j = 0; j < limbs; j++)
        for (int k = 0; k < j; k++)
            inv[k][j] = invmod(c->q[k] % c->q[j], c->q[j]);
    for (size_t i = 0; i < n; i++)
    {
        u64 v[ORC_MAX_LIMBS];
        for (int j = 0; j < limbs; j++)
        {
            u64 x = cp[(size_t)j * n + i];
            for (int k = 0; k < j; k++)
            {
                u64 vk = v[k] % c->q[j];
                x = mulmod(x >= vk ? x - vk : x + c->q[j] - vk, inv[k][j], c->q[j]);
            }
            v[j] = x;
        }
        /* sign: x > Q/2 iff mixed-radix digits (top first) exceed those of (Q-1)/2.  Compute the
         * value and Q in long double; for the negative case evaluate x - Q digit-wise to keep
         * precision: x - Q = sum (v_j - (q_j - 1)) * prod_{k<j} q_k  - 1 */
        long double val = 0, negval = -1.0L, radix = 1.0L;
        int gt_half = 0, decided = 0;
        for (int j = limbs - 1; j >= 0 && !decided; j--)
        {
            /* digits of (Q-1)/2 in mixed radix: (q_j - 1)/2 for every j (all q odd) */
            u64 h = (c->q[j] - 1) >> 1;
            if (v[j] > h)
            {
                gt_half = 1;
                decided = 1;
            }
            else if (v[j] < h)
            {
                gt_half = 0;
                decided = 1;
            }
        }
        for (int j = 0; j < limbs; j++)
        {
            val += (long double)v[j] * radix;
            negval += ((long double)v[j] - (long double)(c->q[j] - 1)) * radix;
            radix *= (long double)c->q[j];
        }
        cv[i] = CMPLX((double)((gt_half ? negval : val) / (long double)scale), 0.0);
    }
    fft_to_rev(c, cv);
    for (size_t i = 0; i < slots; i++)
    {
        double complex z = cv[c->index_map[i]];
        out_reim[2 * i] = creal(z);
        out_reim[2 * i + 1] = cimag(z);
    }
    free(cv);
    free(cp);
}

/* ---------------------------------------------------------------- ModRaise ---- */

/* Bootstrapper::modraise_inplace: M/source/bootstrapping/Bootstrapper.cpp:2938-2992.
 * in: [polys][1][n] (limb 0, NTT form) -> out: [polys][limbs][n] NTT form. */
void orc_modraise(const orc_ctx *c, const u64 *in, int polys, int limbs, u64 *out)
{
    size_t n = c->n;
    u64 q0 = c->q[0], half = q0 >> 1;
    u64 *d = (u64 *)malloc(n * 8);
    for (int p = 0; p < polys; p++)
    {
        memcpy(d, in + (size_t)p * n, n * 8);
        orc_intt(c, 0, d);
        for (int j = 0; j < limbs; j++)
        {
            u64 qj = c->q[j];
            u64 corr = qj - (q0 % qj);
            u64 *dst = out + ((size_t)p * limbs + j) * n;
            for (size_t i = 0; i < n; i++)
            {
                u64 x = d[i] % qj;
                if (d[i] > half)
                {
                    x += corr;
                    if (x >= qj)
                        x -= qj;
                }
                dst[i] = x;
            }
            orc_ntt(c, j, dst);
        }
    }
    free(d);
}

/* ---------------------------------------------------------------- client side (test inputs) ---- */
/* Key generation / encryption / decryption with a simple deterministic PRNG.  These are NOT
 * restatements of SEAL's samplers (identical randomness comes from oracle/_ref); they produce
 * valid RLWE material with the same algebraic shape (S/keygenerator.cpp:303-336,
 * S/util/rlwe.cpp) so that evaluation results can be decrypted in tests on the GPU box. */

static u64 sm64(u64 *s)
{
    u64 z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static u64 uniform_mod(u64 *s, u64 q)
{
    u64 lim = UINT64_MAX - (UINT64_MAX % q);
    u64 r;
    do
        r = sm64(s);
    while (r >= lim);
    return r % q;
}

/* sparse (hamming_weight > 0) or dense ternary secret, NTT form at all key limbs: [kl][n] */
void orc_gen_secret(const orc_ctx *c, u64 seed, int hamming_weight, u64 *sk)
{
    size_t n = c->n;
    int *s = (int *)calloc(n, sizeof(int));
    u64 st = seed ^ 0xA5A5A5A5ULL;
    if (hamming_weight > 0)
    {
        int placed = 0;
        while (placed < hamming_weight)
        {
            size_t idx = sm64(&st) % n;
            if (!s[idx])
            {
                s[idx] = (sm64(&st) & 1) ? 1 : -1;
                placed++;
            }
        }
    }
    else
        for (size_t i = 0; i < n; i++)
            s[i] = (int)(sm64(&st) % 3) - 1;
    for (int l = 0; l < c->n_key_limbs; l++)
    {
        u64 *dst = sk + (size_t)l * n;
        for (size_t i = 0; i < n; i++)
            dst[i] = s[i] < 0 ? c->q[l] - 1 : (u64)s[i];
        orc_ntt(c, l, dst);
    }
    free(s);
}

/* centred binomial noise (21 coin pairs like SEAL's default CBD, sigma ~ 3.2) */
static void sample_noise(u64 *st, size_t n, int *e)
{
    for (size_t i = 0; i < n; i++)
    {
        u64 r = sm64(st);
        e[i] = __builtin_popcountll(r & 0x1FFFFF) - __builtin_popcountll((r >> 21) & 0x1FFFFF);
    }
}

/* symmetric encryption of zero at `limbs` using key limb indices idx[]: (c0, c1) = (-(a s) + e, a) */
static void enc_zero_sym(const orc_ctx *c, const u64 *sk, u64 *st, const int *idx, int limbs, u64 *c0, u64 *c1)
{
    size_t n = c->n;
    int *e = (int *)malloc(n * sizeof(int));
    sample_noise(st, n, e);
    for (int j = 0; j < limbs; j++)
    {
        int l = idx[j];
        u64 q = c->q[l];
        u64 *a = c1 + (size_t)j * n, *b = c0 + (size_t)j * n;
        for (size_t i = 0; i < n; i++)
        {
            a[i] = uniform_mod(st, q);
            b[i] = e[i] < 0 ? q - (u64)(-e[i]) : (u64)e[i];
        }
        orc_ntt(c, l, b);
        const u64 *s = sk + (size_t)l * n;
        for (size_t i = 0; i < n; i++)
        {
            u64 as = mulmod(a[i], s[i], q);
            b[i] = b[i] >= as ? b[i] - as : b[i] + q - as;
        }
    }
    free(e);
}

/* encrypt a plaintext [limbs][n] (NTT form) symmetrically -> ct [2][limbs][n] */
void orc_encrypt_sym(const orc_ctx *c, const u64 *sk, u64 seed, const u64 *pt, int limbs, u64 *ct)
{
    size_t n = c->n;
    int idx[ORC_MAX_LIMBS];
    for (int j = 0; j < limbs; j++)
        idx[j] = j;
    u64 st = seed * 0x2545F4914F6CDD1DULL + 7;
    enc_zero_sym(c, sk, &st, idx, limbs, ct, ct + (size_t)limbs * n);
    for (int j = 0; j < limbs; j++)
    {
        u64 q = c->q[j];
        u64 *b = ct + (size_t)j * n;
        const u64 *m = pt + (size_t)j * n;
        for (size_t i = 0; i < n; i++)
        {
            u64 s = b[i] + m[i];
            b[i] = s >= q ? s - q : s;
        }
    }
}

/* decrypt: S/decryptor.cpp:154-187,299-381 — pt = c0 + c1 s + c2 s^2 (NTT form) */
void orc_decrypt(const orc_ctx *c, const u64 *sk, const u64 *ct, int polys, int limbs, u64 *pt)
{
    size_t n = c->n;
    for (int j = 0; j < limbs; j++)
    {
        u64 q = c->q[j];
        const u64 *s = sk + (size_t)j * n;
        for (size_t i = 0; i < n; i++)
        {
            u64 acc = ct[(size_t)j * n + i];
            u64 sp = s[i];
            for (int p = 1; p < polys; p++)
            {
                acc = (u64)(((u128)ct[((size_t)p * limbs + j) * n + i] * sp + acc) % q);
                sp = mulmod(sp, s[i], q);
            }
            pt[(size_t)j * n + i] = acc;
        }
    }
}

/* KeyGenerator::generate_one_kswitch_key: S/keygenerator.cpp:303-336 — digit J is an encryption
 * of zero at the key level with  p * new_key  added to limb J of component 0.
 * new_key: [kl][n] NTT form (s^2 for relin, sigma(s) for Galois).  out: [digits][2][kl][n]. */
void orc_gen_kswitch_key(const orc_ctx *c, const u64 *sk, u64 seed, const u64 *new_key, u64 *out)
{
    size_t n = c->n;
    int kl = c->n_key_limbs, digits = kl - 1;
    int idx[ORC_MAX_LIMBS];
    for (int j = 0; j < kl; j++)
        idx[j] = j;
    u64 p = c->q[kl - 1];
#pragma omp parallel for schedule(dynamic)
    for (int J = 0; J < digits; J++)
    {
        u64 st = (seed * 0x9E3779B97F4A7C15ULL + 11) ^ ((u64)(J + 1) * 0xD1B54A32D192ED03ULL);
        u64 *c0 = out + ((size_t)J * 2) * kl * n, *c1 = c0 + (size_t)kl * n;
        enc_zero_sym(c, sk, &st, idx, kl, c0, c1);
        u64 q = c->q[J];
        u64 f = p % q;
        u64 *b = c0 + (size_t)J * n;
        const u64 *nk = new_key + (size_t)J * n;
        for (size_t i = 0; i < n; i++)
        {
            u64 s = b[i] + mulmod(nk[i], f, q);
            b[i] = s >= q ? s - q : s;
        }
    }
}

/* s^2 at the key level (relin target key) */
void orc_secret_squared(const orc_ctx *c, const u64 *sk, u64 *out)
{
    for (int l = 0; l < c->n_key_limbs; l++)
        for (size_t i = 0; i < c->n; i++)
            out[(size_t)l * c->n + i] = mulmod(sk[(size_t)l * c->n + i], sk[(size_t)l * c->n + i], c->q[l]);
}

/* sigma_elt(s) at the key level (Galois target key): S/keygenerator.cpp:236-262 */
void orc_secret_galois(const orc_ctx *c, const u64 *sk, uint32_t elt, u64 *out)
{
    size_t n = c->n;
    uint32_t *table = (uint32_t *)malloc(n * 4);
    orc_galois_table(c, elt, table);
    for (int l = 0; l < c->n_key_limbs; l++)
        apply_galois_limb(table, n, sk + (size_t)l * n, out + (size_t)l * n);
    free(table);
}

/* ---------------------------------------------------------------- module level ---- */

/* ct_pt_matrix_mul_wo_pre / _wo_pre_large: M/source/matrix_mul/Ct_pt_matrix_mul.hpp:4-101.
 *   X   : K ciphertexts, each [2][limbs][n]
 *   W   : row-major K x C doubles
 *   out : C ciphertexts, each [2][limbs-1][n]  (after rescale_to_next)
 * out[i] = rescale( sum_j X[j] * encode_scalar(W[j][i], scale) ).  The _large variant differs
 * only in how the output loop is tiled over threads.  columns [c_begin, c_end) are produced. */
void orc_ct_pt_matmul_scalar(const orc_ctx *c, const u64 *X, const double *W, int K, int C, int limbs, double scale,
                             int c_begin, int c_end, u64 *out)
{
    size_t n = c->n;
    size_t ctsz = (size_t)2 * limbs * n;
    size_t outsz = (size_t)2 * (limbs - 1) * n;
#pragma omp parallel for schedule(dynamic)
    for (int i = c_begin; i < c_end; i++)
    {
        u64 *acc = (u64 *)calloc(ctsz, 8);
        u64 consts[ORC_MAX_LIMBS];
        for (int j = 0; j < K; j++)
        {
            orc_encode_scalar_consts(c, W[(size_t)j * C + i], scale, limbs, consts);
            const u64 *x = X + (size_t)j * ctsz;
            for (int p = 0; p < 2; p++)
                for (int l = 0; l < limbs; l++)
                {
                    u64 q = c->q[l], w = consts[l];
                    const u64 *src = x + ((size_t)p * limbs + l) * n;
                    u64 *dst = acc + ((size_t)p * limbs + l) * n;
                    for (size_t t = 0; t < n; t++)
                    {
                        u64 s = dst[t] + mulmod(src[t], w, q);
                        dst[t] = s >= q ? s - q : s;
                    }
                }
        }
        orc_rescale(c, acc, 2, limbs, out + (size_t)(i - c_begin) * outsz);
        free(acc);
    }
}

/* ct_pt_matrix_mul_wo_pre_w_mask: M/source/matrix_mul/Ct_pt_matrix_mul.hpp:103-170 — the
 * plaintext for weight w is encode_vector(w * mask) (mask = bias_vec 0/1 per slot). */
void orc_ct_pt_matmul_masked(const orc_ctx *c, const u64 *X, const double *W, const int *mask, int K, int C, int limbs,
                             double scale, int c_begin, int c_end, u64 *out)
{
    size_t n = c->n, slots = n >> 1;
    size_t ctsz = (size_t)2 * limbs * n;
    size_t outsz = (size_t)2 * (limbs - 1) * n;
#pragma omp parallel for schedule(dynamic)
    for (int i = c_begin; i < c_end; i++)
    {
        u64 *acc = (u64 *)calloc(ctsz, 8);
        u64 *pt = (u64 *)malloc((size_t)limbs * n * 8);
        u64 *tmp = (u64 *)malloc(ctsz * 8);
        double *vals = (double *)calloc(2 * slots, sizeof(double));
        for (int j = 0; j < K; j++)
        {
            double w = W[(size_t)j * C + i];
            for (size_t s = 0; s < slots; s++)
                vals[2 * s] = mask[s] == 1 ? w : 0.0;
            orc_encode_vector(c, vals, (int)slots, scale, limbs, pt);
            orc_multiply_plain(c, X + (size_t)j * ctsz, pt, 2, limbs, tmp);
            orc_addsub(c, 0, acc, tmp, 2, limbs, acc);
        }
        orc_rescale(c, acc, 2, limbs, out + (size_t)(i - c_begin) * outsz);
        free(acc);
        free(pt);
        free(tmp);
        free(vals);
    }
}
