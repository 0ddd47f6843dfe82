// Element-wise, rescale / mod-switch / mod-raise and key-switching kernels.
// Reference behaviour (canonical residues, SEAL layout) per SURVEY Appendix D:
//   add/sub/negate      S/util/polyarithsmallmod.h:77-300, S/evaluator.cpp:130-350
//   multiply_plain      S/evaluator.cpp:2336-2373
//   multiply / square   S/evaluator.cpp:770-909, 1223-1282
//   rescale             S/util/rns.cpp:830-901 (divide_and_round_q_last_ntt_inplace)
//   mod-switch          S/evaluator.cpp:1483-1546
//   key switch          S/evaluator.cpp:2724-3021
//   Galois (NTT form)   S/util/galois.cpp:192-218
//   ModRaise            M/source/bootstrapping/Bootstrapper.cpp:2938-2992
// All kernels are HBM-streaming: flat 1-D grids, 16-byte vector accesses, one limb constant
// lookup per thread (limb index is uniform per 2-element vector because n is even).
#include "ntt.cuh"
#include "ops.cuh"
#include "ksgroup.hpp"
#include "fpfield.cuh"
#include <cstdlib>

namespace moai
{
    namespace
    {
        constexpr int EW_THREADS = 256;

        inline unsigned grid_for(long long work_items)
        {
            return (unsigned)((work_items + EW_THREADS - 1) / EW_THREADS);
        }

        // ------------------------------------------------------------------ add / sub / negate
        // a, b and out may alias each other (in-place calls, double_inplace passes a, a, a): no __restrict__ on them
        __global__ void k_addsub(int op, const ulonglong2 *a, const ulonglong2 *b, ulonglong2 *out, long long total2, int log_n2, int limbs,
                                 const LimbConst *__restrict__ lcs, long long b_period2)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total2)
            {
                return;
            }
            const int limb = (int)((i >> log_n2) % limbs);
            const u64 q = lcs[limb].q;
            ulonglong2 x = a[i], r;
            if (op == EW_NEG)
            {
                r.x = negmod(x.x, q);
                r.y = negmod(x.y, q);
            }
            else
            {
                ulonglong2 y = b[b_period2 ? i % b_period2 : i];
                if (op == EW_ADD)
                {
                    r.x = addmod(x.x, y.x, q);
                    r.y = addmod(x.y, y.y, q);
                }
                else
                {
                    r.x = submod(x.x, y.x, q);
                    r.y = submod(x.y, y.y, q);
                }
            }
            out[i] = r;
        }

        // ct[b][p][l][i] (op) pt[b*stride][l][i] on poly 0, copy on the others
        __global__ void k_addsub_plain(int op, const ulonglong2 *ct, const ulonglong2 *__restrict__ pt,
                                       ulonglong2 *out /* may alias ct */, long long total2, int log_n2, int polys,
                                       int limbs,
                                       long long pt_stride2, const LimbConst *__restrict__ lcs)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total2)
            {
                return;
            }
            const long long lp = i >> log_n2; // (b * polys + p) * limbs + l
            const int limb = (int)(lp % limbs);
            const long long bp = lp / limbs;
            const int p = (int)(bp % polys);
            ulonglong2 x = ct[i];
            if (p == 0)
            {
                const long long b = bp / polys;
                const u64 q = lcs[limb].q;
                const long long within = i & (((long long)1 << log_n2) - 1);
                ulonglong2 y = pt[b * pt_stride2 + ((long long)limb << log_n2) + within];
                if (op == EW_ADD)
                {
                    x.x = addmod(x.x, y.x, q);
                    x.y = addmod(x.y, y.y, q);
                }
                else
                {
                    x.x = submod(x.x, y.x, q);
                    x.y = submod(x.y, y.y, q);
                }
            }
            out[i] = x;
        }

        // ct and out may alias (in-place multiply_plain)
        __global__ void k_multiply_plain(const ulonglong2 *ct, const ulonglong2 *__restrict__ pt,
                                         ulonglong2 *out, long long total2, int log_n2, int polys,
                                         int limbs, long long pt_stride2, const LimbConst *__restrict__ lcs)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total2)
            {
                return;
            }
            const long long lp = i >> log_n2;
            const int limb = (int)(lp % limbs);
            const long long b = lp / limbs / polys;
            const LimbConst lc = lcs[limb];
            const long long within = i & (((long long)1 << log_n2) - 1);
            ulonglong2 x = ct[i];
            ulonglong2 y = pt[b * pt_stride2 + ((long long)limb << log_n2) + within];
            x.x = mulmod(x.x, y.x, lc);
            x.y = mulmod(x.y, y.y, lc);
            out[i] = x;
        }

        // mode 0: out = ct * k[l] (all polys);  mode 1: out = ct + k[l] on poly 0;  mode 2: out = 2 ct (+ k[l] on poly 0)
        // ct and out may alias (add_plain_inplace with a scalar plaintext)
        __global__ void k_scalar(int mode, const ulonglong2 *ct, const Twiddle *__restrict__ consts,
                                 ulonglong2 *out, long long total2, int log_n2, int polys, int limbs,
                                 const LimbConst *__restrict__ lcs)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total2)
            {
                return;
            }
            const long long lp = i >> log_n2;
            const int limb = (int)(lp % limbs);
            const int p = (int)((lp / limbs) % polys);
            const u64 q = lcs[limb].q;
            const Twiddle k = consts[limb];
            ulonglong2 x = ct[i];
            if (mode == 0)
            {
                x.x = mul_shoup(x.x, k.w, k.wq, q);
                x.y = mul_shoup(x.y, k.w, k.wq, q);
            }
            else
            {
                if (mode == 2)
                {
                    x.x = addmod(x.x, x.x, q);
                    x.y = addmod(x.y, x.y, q);
                }
                if (p == 0)
                {
                    x.x = addmod(x.x, k.w, q);
                    x.y = addmod(x.y, k.w, q);
                }
            }
            out[i] = x;
        }

        // acc3[b][p][l][.] += x2[b][p][l][.] for p < 2: a size-2 ciphertext added into a size-3 one (SEAL's add pads the
        // smaller operand, S/evaluator.cpp:190-212)
        __global__ void k_add_into3(ulonglong2 *acc3, const ulonglong2 *__restrict__ x2, long long total2, int log_n2, int limbs,
                                    const LimbConst *__restrict__ lcs)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [batch][2][limbs][n/2]
            if (i >= total2)
            {
                return;
            }
            const long long lp = i >> log_n2;
            const int limb = (int)(lp % limbs);
            const long long bp = lp / limbs, b = bp >> 1, p = bp & 1;
            const u64 q = lcs[limb].q;
            ulonglong2 *dst = acc3 + ((((b * 3 + p) * limbs + limb) << log_n2) + (i & (((long long)1 << log_n2) - 1)));
            const ulonglong2 v = x2[i];
            ulonglong2 a = *dst;
            a.x = addmod(a.x, v.x, q);
            a.y = addmod(a.y, v.y, q);
            *dst = a;
        }

        // out = sum_j in_j * k[j][l]: a linear combination of up to LC_MAX ciphertext batches with per-limb scalar
        // constants, every input read at ITS OWN limb count (only the first `limbs` limbs are used: the mod-switch to
        // the output level costs nothing).  One pass instead of a copy + multiply + add per term — the leaves
        // sum_j c_j T_j of the Chebyshev evaluation in EvalMod (M/source/bootstrapping/common/Polynomial.cpp:322-495
        // does them with multiply_const + add per term).
        constexpr int LC_MAX = 8;
        struct LinCombArgs
        {
            const ulonglong2 *in[LC_MAX];
            int in_limbs[LC_MAX];
            int n_terms;
        };
        __global__ void k_lincomb(LinCombArgs a, const Twiddle *__restrict__ consts, ulonglong2 *__restrict__ out,
                                  long long total2, int log_n2, int limbs, const LimbConst *__restrict__ lcs)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [batch * polys][limbs][n/2]
            if (i >= total2)
            {
                return;
            }
            const long long lp = i >> log_n2;
            const int limb = (int)(lp % limbs);
            const long long bp = lp / limbs;
            const long long within = i & (((long long)1 << log_n2) - 1);
            const u64 q = lcs[limb].q;
            ulonglong2 acc{ 0, 0 };
#pragma unroll
            for (int j = 0; j < LC_MAX; j++)
            {
                if (j < a.n_terms)
                {
                    const ulonglong2 v = a.in[j][((bp * a.in_limbs[j] + limb) << log_n2) + within];
                    const Twiddle k = consts[j * limbs + limb];
                    acc.x = addmod(acc.x, mul_shoup(v.x, k.w, k.wq, q), q);
                    acc.y = addmod(acc.y, mul_shoup(v.y, k.w, k.wq, q), q);
                }
            }
            out[i] = acc;
        }

        // (a0, a1) x (b0, b1) -> (a0 b0, a0 b1 + a1 b0, a1 b1); one thread per coefficient pair of a limb
        __global__ void k_multiply(const ulonglong2 *__restrict__ a, const ulonglong2 *__restrict__ b,
                                   ulonglong2 *__restrict__ out, long long total2, int log_n2, int limbs,
                                   const LimbConst *__restrict__ lcs, int accumulate, int square, int b_bcast,
                                   int a_limbs, int b_limbs)
        {
            // a_limbs / b_limbs >= limbs: limbs stored per polynomial of a / b (an operand at a higher level is read in
            // place, which IS its mod-switch: S/evaluator.cpp:1583-1652 drops the trailing limbs)
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [batch][limbs][n/2]
            if (i >= total2)
            {
                return;
            }
            const long long lb = i >> log_n2; // b * limbs + l
            const int limb = (int)(lb % limbs);
            const long long bt = lb / limbs;
            const LimbConst lc = lcs[limb];
            const long long within = i & (((long long)1 << log_n2) - 1);
            const long long poly2 = (long long)limbs << log_n2;
            const long long apoly2 = (long long)a_limbs << log_n2, bpoly2 = (long long)b_limbs << log_n2;
            const long long off_a = bt * 2 * apoly2 + ((long long)limb << log_n2) + within;
            const long long off_out = bt * 3 * poly2 + ((long long)limb << log_n2) + within;
            ulonglong2 a0 = a[off_a], a1 = a[off_a + apoly2];
            const long long off_b = (b_bcast ? 0 : bt * 2 * bpoly2) + ((long long)limb << log_n2) + within;
            ulonglong2 b0 = square ? a0 : b[off_b], b1 = square ? a1 : b[off_b + bpoly2];
            ulonglong2 r0, r1, r2;
            r0.x = mulmod(a0.x, b0.x, lc);
            r0.y = mulmod(a0.y, b0.y, lc);
            r2.x = mulmod(a1.x, b1.x, lc);
            r2.y = mulmod(a1.y, b1.y, lc);
            u128 m;
            m = mul_wide(a0.x, b1.x);
            mac_wide(m, a1.x, b0.x);
            r1.x = barrett_reduce_wide(m, lc);
            m = mul_wide(a0.y, b1.y);
            mac_wide(m, a1.y, b0.y);
            r1.y = barrett_reduce_wide(m, lc);
            if (accumulate)
            {
                ulonglong2 o0 = out[off_out], o1 = out[off_out + poly2], o2 = out[off_out + 2 * poly2];
                r0.x = addmod(r0.x, o0.x, lc.q);
                r0.y = addmod(r0.y, o0.y, lc.q);
                r1.x = addmod(r1.x, o1.x, lc.q);
                r1.y = addmod(r1.y, o1.y, lc.q);
                r2.x = addmod(r2.x, o2.x, lc.q);
                r2.y = addmod(r2.y, o2.y, lc.q);
            }
            out[off_out] = r0;
            out[off_out + poly2] = r1;
            out[off_out + 2 * poly2] = r2;
        }

        // ------------------------------------------------------------------ divide-and-round by the last limb
        // in:  [P][limbs_in][n] with the divisor limb last (prime index `last_id`), targets 0..limbs_in-2
        // step 1 (host): t[P][n] = INTT_last(in[.][limbs_in-1])
        // step 2 (fused into the forward NTT's first pass, NttPrologue mode 2):
        //         u[P][limbs_in-1][n] = ((t + half) mod q_last) mod q_i + (q_i - half mod q_i)     (< 2 q_i)
        // step 3: u = NTT_i(u)
        // step 4: out[P][targets][n] = (in[P][limbs_in][n](limb i) - u) * q_last^-1 mod q_i  (+ addend)
        __global__ void k_divround_finish(const ulonglong2 *__restrict__ in, const ulonglong2 *__restrict__ u,
                                          const ulonglong2 *addend, ulonglong2 *out, // may alias each other
                                          long long total2, int log_n2, int targets, int limbs_in, int last_id, int kl,
                                          const LimbConst *__restrict__ lcs, const Twiddle *__restrict__ inv_last,
                                          int addend_even_only, int addend_group,
                                          const Twiddle *__restrict__ addend_mul = nullptr, int addend_limbs = 0)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [P][targets][n/2]
            if (i >= total2)
            {
                return;
            }
            const long long lp = i >> log_n2;
            const int limb = (int)(lp % targets);
            const long long p = lp / targets;
            const long long within = i & (((long long)1 << log_n2) - 1);
            const u64 q = lcs[limb].q;
            const Twiddle inv = inv_last[(size_t)last_id * kl + limb];
            ulonglong2 x = in[((p * limbs_in + limb) << log_n2) + within];
            ulonglong2 y = u[i];
            ulonglong2 r;
            r.x = mul_shoup(submod(x.x, y.x, q), inv.w, inv.wq, q);
            r.y = mul_shoup(submod(x.y, y.y, q), inv.w, inv.wq, q);
            if (addend && !(addend_even_only && (p & 1)))
            {
                // the addend's ciphertexts may hold more polynomials than the result's two (relinearize reads c0, c1
                // straight out of the size-3 input): polynomial p of the result <-> (p / 2) * addend_group + p % 2
                // merged mod-down + rescale: the addend's polynomials hold addend_limbs limbs and are multiplied by
                // addend_mul[limb] = q_last^-1 mod q_limb
                ulonglong2 z = addend[((((p >> 1) * addend_group + (p & 1)) * (addend_limbs ? addend_limbs : targets) + limb)
                                       << log_n2) + within];
                if (addend_mul)
                {
                    const Twiddle am = addend_mul[limb];
                    z.x = mul_shoup(z.x, am.w, am.wq, q);
                    z.y = mul_shoup(z.y, am.w, am.wq, q);
                }
                r.x = addmod(r.x, z.x, q);
                r.y = addmod(r.y, z.y, q);
            }
            out[i] = r;
        }

        // in [P][limbs_in][n] -> out [P][limbs_in-1][n] (+= addend of the same shape when given)
        // addend_even_only: add `addend` only to the even polynomials (c0 of size-2 ciphertexts)
        void divide_round_last(Context *c, const u64 *in, long long P, int limbs_in, int last_id, const u64 *addend,
                               u64 *out, bool addend_even_only = false, int addend_group = 2)
        {
            const size_t n = c->n;
            const int targets = limbs_in - 1;
            Scratch t(P * n * sizeof(u64), c->stream);
            Scratch u((size_t)P * targets * n * sizeof(u64), c->stream);
            ntt_inverse_from(c, in + (size_t)targets * n, 1, limbs_in, t.as<u64>(), P, c->d_ids + last_id, 1);
            const long long total2 = P * targets * (long long)(n / 2);
            // expansion (+half, reduce into every target prime, +fix) fused into the NTT's first pass
            NttPrologue pro;
            pro.src = t.as<u64>();
            pro.mode = 2;
            pro.last_id = last_id;
            // (c - u) q_last^-1 (+ addend) rides on the transform's second pass when the launch allows it
            FinishEpi fin;
            fin.in = in;
            fin.addend = addend;
            fin.out = out;
            fin.inv = c->d_inv_last + (size_t)last_id * c->kl;
            fin.limbs_in = limbs_in;
            fin.addend_even_only = addend_even_only ? 1 : 0;
            fin.addend_group = addend_group;
            if (ntt_forward(c, u.as<u64>(), P * targets, c->d_ids, targets, 1, &pro, 3, &fin))
            {
                return;
            }
            KernelTimer kt1(c, "k_divround_finish", 1);
            k_divround_finish<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
                reinterpret_cast<const ulonglong2 *>(in), u.as<ulonglong2>(),
                reinterpret_cast<const ulonglong2 *>(addend), reinterpret_cast<ulonglong2 *>(out), total2, c->log_n - 1,
                targets, limbs_in, last_id, c->kl, c->d_limb, c->d_inv_last, addend_even_only ? 1 : 0, addend_group);
            c->launches += 1;
            MOAI_CUDA_CHECK(cudaGetLastError());
        }

        // ------------------------------------------------------------------ mod raise
        // d[P][n] (coefficients mod q0) -> out[P][limbs][n]: centred lift into every prime
        __global__ void k_modraise_expand(const ulonglong2 *__restrict__ d, ulonglong2 *__restrict__ out,
                                          long long total2, int log_n2, int limbs, const LimbConst *__restrict__ lcs)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total2)
            {
                return;
            }
            const long long lp = i >> log_n2;
            const int limb = (int)(lp % limbs);
            const long long p = lp / limbs;
            const long long within = i & (((long long)1 << log_n2) - 1);
            const LimbConst lc = lcs[limb];
            const u64 q0 = lcs[0].q, half = q0 >> 1;
            const u64 corr = lc.q - lc.q0_mod; // -(q0) mod q_j
            ulonglong2 v = d[(p << log_n2) + within], r;
            r.x = reduce64(v.x, lc);
            r.y = reduce64(v.y, lc);
            if (v.x > half)
            {
                r.x = addmod(r.x, corr, lc.q);
            }
            if (v.y > half)
            {
                r.y = addmod(r.y, corr, lc.q);
            }
            out[i] = r;
        }

        // ------------------------------------------------------------------ Galois permutation
        __global__ void k_galois(const u64 *__restrict__ in, u64 *__restrict__ out, long long total, int log_n,
                                 const uint32_t *__restrict__ table)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total)
            {
                return;
            }
            const long long within = i & (((long long)1 << log_n) - 1);
            out[i] = in[i - within + table[within]];
        }

        // out[per_ct] = sum over the batch of a[b][per_ct]  (modular; lazy 64-bit partial sums)
        __global__ void k_sum_batch(const ulonglong2 *__restrict__ a, ulonglong2 *__restrict__ out, long long per_ct2,
                                    long long batch, int log_n2, int limbs, const LimbConst *__restrict__ lcs)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= per_ct2)
            {
                return;
            }
            const int limb = (int)((i >> log_n2) % limbs);
            const LimbConst lc = lcs[limb];
            u64 sx = 0, sy = 0;
            for (long long b = 0; b < batch; b++)
            {
                const ulonglong2 v = a[b * per_ct2 + i];
                sx += v.x; // q < 2^61: up to 8 canonical residues fit in 64 bits
                sy += v.y;
                if ((b & 7) == 7)
                {
                    sx = reduce64(sx, lc);
                    sy = reduce64(sy, lc);
                }
            }
            ulonglong2 r;
            r.x = reduce64(sx, lc);
            r.y = reduce64(sy, lc);
            out[i] = r;
        }

        // out3 = sum_j a[j] (x) b[j]  (size-2 x size-2 -> size-3, summed over the batch): the inner loop
        // of ct_ct_matrix_mul_colpacking (M/source/matrix_mul/Ct_ct_matrix_mul.hpp:33-41) without
        // materialising the 64 size-3 products.  128-bit lazy sums: 2 * batch * q^2 must stay < 2^128.
        // mode 1: b is ONE ciphertext m and the terms are (a[j] - m)^2  (LayerNorm's variance sum,
        // M/source/non_linear_func/layernorm.hpp:245-260).
        __global__ void k_inner_product(const ulonglong2 *__restrict__ a, const ulonglong2 *__restrict__ b,
                                        ulonglong2 *__restrict__ out, long long batch, int log_n2, int limbs, int mode,
                                        const LimbConst *__restrict__ lcs, const Twiddle *__restrict__ two64)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [limbs][n/2]
            const int limb = (int)(i >> log_n2);
            if (limb >= limbs)
            {
                return;
            }
            const LimbConst lc = lcs[limb];
            const Twiddle t64 = two64[limb];
            const long long poly2 = (long long)limbs << log_n2;
            u128 s0x{ 0, 0 }, s0y{ 0, 0 }, s1x{ 0, 0 }, s1y{ 0, 0 }, s2x{ 0, 0 }, s2y{ 0, 0 };
            ulonglong2 m0, m1;
            if (mode == 1)
            {
                m0 = b[i];
                m1 = b[i + poly2];
            }
            for (long long j = 0; j < batch; j++)
            {
                ulonglong2 a0 = a[j * 2 * poly2 + i], a1 = a[j * 2 * poly2 + poly2 + i];
                ulonglong2 b0, b1;
                if (mode == 1)
                {
                    a0.x = submod(a0.x, m0.x, lc.q);
                    a0.y = submod(a0.y, m0.y, lc.q);
                    a1.x = submod(a1.x, m1.x, lc.q);
                    a1.y = submod(a1.y, m1.y, lc.q);
                    b0 = a0;
                    b1 = a1;
                }
                else
                {
                    b0 = b[j * 2 * poly2 + i];
                    b1 = b[j * 2 * poly2 + poly2 + i];
                }
                mac_wide(s0x, a0.x, b0.x);
                mac_wide(s0y, a0.y, b0.y);
                mac_wide(s1x, a0.x, b1.x);
                mac_wide(s1x, a1.x, b0.x);
                mac_wide(s1y, a0.y, b1.y);
                mac_wide(s1y, a1.y, b0.y);
                mac_wide(s2x, a1.x, b1.x);
                mac_wide(s2y, a1.y, b1.y);
            }
            ulonglong2 r;
            r.x = barrett_reduce_acc(s0x, lc, t64.w, t64.wq);
            r.y = barrett_reduce_acc(s0y, lc, t64.w, t64.wq);
            out[i] = r;
            r.x = barrett_reduce_acc(s1x, lc, t64.w, t64.wq);
            r.y = barrett_reduce_acc(s1y, lc, t64.w, t64.wq);
            out[i + poly2] = r;
            r.x = barrett_reduce_acc(s2x, lc, t64.w, t64.wq);
            r.y = barrett_reduce_acc(s2y, lc, t64.w, t64.wq);
            out[i + 2 * poly2] = r;
        }


        // ------------------------------------------------------------------ BSGS inner sums
        // inner[i][b][p] = sum_j pt[i][j] (.) rot[j][b][p]  for every giant step i of one linear stage
        // (the plaintext-matrix inner loops of M/source/bootstrapping/Bootstrapper.cpp:2021-2044) in ONE
        // pass: each thread keeps the baby-step rotations of its coefficient pair in registers and
        // streams the pre-rotated diagonals; 128-bit lazy sums (<= 16 products below 2^122 each).
        struct BsgsArgs
        {
            const u64 *rot[BSGS_MAX_BABY];                  // [batch][2][limbs][n] each
            const u64 *pt[BSGS_MAX_GIANT][BSGS_MAX_BABY];   // [limbs][n] each, nullptr = absent diagonal
            u64 *out[BSGS_MAX_GIANT];                       // [batch][2][limbs][n] each
            int n_baby, n_giant;
        };

        __global__ void __launch_bounds__(EW_THREADS) k_bsgs_inner(BsgsArgs a, long long polys, int log_n2, int limbs,
                                                                   const LimbConst *__restrict__ lcs,
                                                                   const Twiddle *__restrict__ two64)
        {
            // grid: the (ciphertext, polynomial) index is FASTEST, then the coefficient block, then the limb:
            // CTAs that share a tile of the diagonals run together, so the plaintexts (63 x limbs x 512 KiB
            // per stage, far more than the ciphertext data) are streamed from HBM once per batch, not
            // once per ciphertext.
            const long long bp = blockIdx.x % polys; // b * 2 + p
            const long long rest = blockIdx.x / polys;
            const int cblks = (1 << log_n2) / EW_THREADS;
            const long long within = (rest % cblks) * EW_THREADS + threadIdx.x;
            const int limb = (int)(rest / cblks);
            const long long i = ((bp * limbs + limb) << log_n2) + within; // over [batch][2][limbs][n/2]
            const long long pt_off = ((long long)limb << log_n2) + within;
            const LimbConst lc = lcs[limb];
            const Twiddle t64 = two64[limb];
            ulonglong2 r[BSGS_MAX_BABY];
#pragma unroll
            for (int j = 0; j < BSGS_MAX_BABY; j++)
            {
                if (j < a.n_baby)
                {
                    r[j] = reinterpret_cast<const ulonglong2 *>(a.rot[j])[i];
                }
            }
            for (int g = 0; g < a.n_giant; g++)
            {
                u128 sx{ 0, 0 }, sy{ 0, 0 };
#pragma unroll
                for (int j = 0; j < BSGS_MAX_BABY; j++)
                {
                    if (j < a.n_baby && a.pt[g][j])
                    {
                        const ulonglong2 w = __ldg(reinterpret_cast<const ulonglong2 *>(a.pt[g][j]) + pt_off);
                        mac_wide(sx, r[j].x, w.x);
                        mac_wide(sy, r[j].y, w.y);
                    }
                }
                ulonglong2 o;
                o.x = barrett_reduce_acc(sx, lc, t64.w, t64.wq);
                o.y = barrett_reduce_acc(sy, lc, t64.w, t64.wq);
                reinterpret_cast<ulonglong2 *>(a.out[g])[i] = o;
            }
        }

        // ------------------------------------------------------------------ key switch pieces
        // ext[b][I][J][n] = d[b][J][n] mod m_I   (I over {q_0..q_{l-1}, p})
        __global__ void k_ks_expand(const ulonglong2 *__restrict__ d, ulonglong2 *__restrict__ ext, long long total2,
                                    int log_n2, int limbs, const int *__restrict__ ids_ks,
                                    const LimbConst *__restrict__ lcs)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [b][I][J][n/2]
            if (i >= total2)
            {
                return;
            }
            const long long lp = i >> log_n2;
            const int J = (int)(lp % limbs);
            const long long bi = lp / limbs;
            const int I = (int)(bi % (limbs + 1));
            const long long b = bi / (limbs + 1);
            const long long within = i & (((long long)1 << log_n2) - 1);
            const LimbConst lc = lcs[ids_ks[I]];
            ulonglong2 v = d[((b * limbs + J) << log_n2) + within];
            ulonglong2 r;
            r.x = reduce64(v.x, lc);
            r.y = reduce64(v.y, lc);
            ext[i] = r;
        }

        // acc[b][k][I][n] = sum_J ext[b][I][J][n] * ksk[J][k][ids_ks[I]][n] mod m_I
        __global__ void k_ks_mac(const ulonglong2 *__restrict__ ext, const ulonglong2 *__restrict__ ksk,
                                 ulonglong2 *__restrict__ acc, long long batch, int log_n2, int limbs, int rns, int n_data,
                                 int key_kl, const int *__restrict__ ids_ks, const LimbConst *__restrict__ lcs,
                                 const Twiddle *__restrict__ two64, int I0)
        {
            // grid: x = batch item (fastest, so that CTAs sharing a key tile run together and the evk
            // is streamed from HBM once per chunk), y = I, z = coefficient block
            const long long within = (long long)blockIdx.z * blockDim.x + threadIdx.x;
            const int I = blockIdx.y + I0; // I0: first target modulus of this launch
            const long long b = blockIdx.x;
            const int prime = ids_ks[I];
            // position of modulus I inside the key: data limbs first, the extra / special primes last
            // (a level-truncated key keeps fewer data limbs, see key_prepare; KsShape in ntt.cuh); limbs = digits
            const int key_limb = I < n_data ? I : I + key_kl - rns;
            const LimbConst lc = lcs[prime];
            const Twiddle t64 = two64[prime];
            const long long n2 = (long long)1 << log_n2;
            const ulonglong2 *e = ext + ((b * rns + I) * limbs << log_n2) + within;
            u128 a0x{ 0, 0 }, a0y{ 0, 0 }, a1x{ 0, 0 }, a1y{ 0, 0 };
            for (int J = 0; J < limbs; J++)
            {
                const ulonglong2 v = e[(long long)J << log_n2];
                const ulonglong2 k0 = __ldg(ksk + (((long long)J * 2 + 0) * key_kl + key_limb) * n2 + within);
                const ulonglong2 k1 = __ldg(ksk + (((long long)J * 2 + 1) * key_kl + key_limb) * n2 + within);
                mac_wide(a0x, v.x, k0.x);
                mac_wide(a0y, v.y, k0.y);
                mac_wide(a1x, v.x, k1.x);
                mac_wide(a1y, v.y, k1.y);
            }
            ulonglong2 r0, r1;
            r0.x = barrett_reduce_acc(a0x, lc, t64.w, t64.wq);
            r0.y = barrett_reduce_acc(a0y, lc, t64.w, t64.wq);
            r1.x = barrett_reduce_acc(a1x, lc, t64.w, t64.wq);
            r1.y = barrett_reduce_acc(a1y, lc, t64.w, t64.wq);
            acc[(((b * 2 + 0) * rns + I) << log_n2) + within] = r0;
            acc[(((b * 2 + 1) * rns + I) << log_n2) + within] = r1;
        }
    } // namespace

    // ====================================================================== launchers
    // 128-bit integer inner product of target modulus I (integer-path moduli of a key switch of shape `sh`)
    void ks_mac_int(Context *c, const u64 *ext, const u64 *ksk, u64 *acc, long long batch, const KsShape &sh, int key_kl,
                    int I)
    {
        dim3 grid((unsigned)batch, 1u, (unsigned)((c->n / 2) / EW_THREADS));
        KernelTimer kt(c, "k_ks_mac", 1);
        k_ks_mac<<<grid, EW_THREADS, 0, c->stream>>>(reinterpret_cast<const ulonglong2 *>(ext),
                                                     reinterpret_cast<const ulonglong2 *>(ksk),
                                                     reinterpret_cast<ulonglong2 *>(acc), batch, c->log_n - 1, sh.digits,
                                                     sh.rns, sh.n_data, key_kl, sh.ids, c->d_limb, c->d_two64, I);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    // out[P][targets][n] = (in[P][limbs_in][n](limb j < targets) - u[P][targets][n]) * inv[j] mod q_j (+ addend)
    void divround_finish(Context *c, const u64 *in, const u64 *u, const u64 *addend, u64 *out, long long P, int targets,
                         int limbs_in, const Twiddle *d_inv, bool addend_even_only, int addend_group,
                         const Twiddle *d_addend_mul, int addend_limbs)
    {
        const long long total2 = P * targets * (long long)(c->n / 2);
        KernelTimer kt(c, "k_divround_finish", 1);
        k_divround_finish<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
            reinterpret_cast<const ulonglong2 *>(in), reinterpret_cast<const ulonglong2 *>(u),
            reinterpret_cast<const ulonglong2 *>(addend), reinterpret_cast<ulonglong2 *>(out), total2, c->log_n - 1,
            targets, limbs_in, 0, 0, c->d_limb, d_inv, addend_even_only ? 1 : 0, addend_group, d_addend_mul, addend_limbs);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ew_addsub(Context *c, int op, const u64 *a, const u64 *b, u64 *out, long long batch, int polys, int limbs,
                   bool b_broadcast)
    {
        const long long total2 = batch * polys * limbs * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        KernelTimer kt2(c, "k_addsub", 1);
        k_addsub<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
            op, reinterpret_cast<const ulonglong2 *>(a), reinterpret_cast<const ulonglong2 *>(b),
            reinterpret_cast<ulonglong2 *>(out), total2, c->log_n - 1, limbs, c->d_limb,
            b_broadcast ? (long long)polys * limbs * (long long)(c->n / 2) : 0);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ew_addsub_plain(Context *c, int op, const u64 *ct, const u64 *pt, u64 *out, long long batch, int polys,
                         int limbs, long long pt_stride)
    {
        const long long total2 = batch * polys * limbs * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        k_addsub_plain<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
            op, reinterpret_cast<const ulonglong2 *>(ct), reinterpret_cast<const ulonglong2 *>(pt),
            reinterpret_cast<ulonglong2 *>(out), total2, c->log_n - 1, polys, limbs, pt_stride / 2, c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ew_multiply_plain(Context *c, const u64 *ct, const u64 *pt, u64 *out, long long batch, int polys, int limbs,
                           long long pt_stride)
    {
        const long long total2 = batch * polys * limbs * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        KernelTimer kt3(c, "k_multiply_plain", 1);
        k_multiply_plain<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
            reinterpret_cast<const ulonglong2 *>(ct), reinterpret_cast<const ulonglong2 *>(pt),
            reinterpret_cast<ulonglong2 *>(out), total2, c->log_n - 1, polys, limbs, pt_stride / 2, c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    static void scalar_op(Context *c, int mode, const u64 *ct, const u64 *h_consts, u64 *out, long long batch,
                          int polys, int limbs)
    {
        const long long total2 = batch * polys * limbs * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        std::vector<Twiddle> h(limbs);
        for (int l = 0; l < limbs; l++)
        {
            MOAI_REQUIRE(h_consts[l] < c->q[l], "scalar constant must be reduced");
            h[l].w = h_consts[l];
            h[l].wq = (u64)((((unsigned __int128)h_consts[l]) << 64) / c->q[l]);
        }
        Scratch d(limbs * sizeof(Twiddle), c->stream);
        MOAI_CUDA_CHECK(cudaMemcpyAsync(d.p, h.data(), limbs * sizeof(Twiddle), cudaMemcpyHostToDevice, c->stream));
        // h is pageable: the copy is staged before returning, so it may go out of scope
        KernelTimer kt4(c, "k_scalar", 1);
        k_scalar<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(mode, reinterpret_cast<const ulonglong2 *>(ct),
                                                                 d.as<Twiddle>(), reinterpret_cast<ulonglong2 *>(out),
                                                                 total2, c->log_n - 1, polys, limbs, c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ew_lincomb_scalar(Context *c, int n_terms, const u64 *const *in, const int *in_limbs, const u64 *h_consts,
                           u64 *out, long long batch, int polys, int limbs)
    {
        MOAI_REQUIRE(n_terms >= 1 && n_terms <= LC_MAX, "1..8 terms per linear combination");
        const long long total2 = batch * polys * limbs * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        LinCombArgs a{};
        a.n_terms = n_terms;
        std::vector<Twiddle> h((size_t)n_terms * limbs);
        for (int j = 0; j < n_terms; j++)
        {
            MOAI_REQUIRE(in_limbs[j] >= limbs, "a term is below the output level");
            a.in[j] = reinterpret_cast<const ulonglong2 *>(in[j]);
            a.in_limbs[j] = in_limbs[j];
            for (int l = 0; l < limbs; l++)
            {
                const u64 k = h_consts[(size_t)j * limbs + l];
                MOAI_REQUIRE(k < c->q[l], "scalar constant must be reduced");
                h[(size_t)j * limbs + l].w = k;
                h[(size_t)j * limbs + l].wq = (u64)((((unsigned __int128)k) << 64) / c->q[l]);
            }
        }
        Scratch d(h.size() * sizeof(Twiddle), c->stream);
        MOAI_CUDA_CHECK(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(Twiddle), cudaMemcpyHostToDevice, c->stream));
        KernelTimer kt(c, "k_lincomb", 1);
        k_lincomb<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(a, d.as<Twiddle>(), reinterpret_cast<ulonglong2 *>(out),
                                                                  total2, c->log_n - 1, limbs, c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ew_multiply_scalar(Context *c, const u64 *ct, const u64 *h_consts, u64 *out, long long batch, int polys,
                            int limbs)
    {
        scalar_op(c, 0, ct, h_consts, out, batch, polys, limbs);
    }

    void ew_double_add_scalar(Context *c, const u64 *ct, const u64 *h_consts, u64 *out, long long batch, int polys, int limbs)
    {
        scalar_op(c, 2, ct, h_consts, out, batch, polys, limbs);
    }

    void ew_add_into3(Context *c, u64 *acc3, const u64 *x2, long long batch, int limbs)
    {
        const long long total2 = batch * 2 * limbs * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        KernelTimer kt(c, "k_addsub", 1);
        k_add_into3<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(reinterpret_cast<ulonglong2 *>(acc3),
                                                                    reinterpret_cast<const ulonglong2 *>(x2), total2,
                                                                    c->log_n - 1, limbs, c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ew_add_scalar(Context *c, const u64 *ct, const u64 *h_consts, u64 *out, long long batch, int polys, int limbs)
    {
        scalar_op(c, 1, ct, h_consts, out, batch, polys, limbs);
    }

    void ew_multiply(Context *c, const u64 *a, const u64 *b, u64 *out3, long long batch, int limbs, bool accumulate,
                     bool b_broadcast, int a_limbs, int b_limbs)
    {
        a_limbs = a_limbs > 0 ? a_limbs : limbs;
        b_limbs = b_limbs > 0 ? b_limbs : limbs;
        MOAI_REQUIRE(a_limbs >= limbs && b_limbs >= limbs, "operand below the product's level");
        const long long total2 = batch * limbs * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        KernelTimer kt5(c, "k_multiply", 1);
        k_multiply<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
            reinterpret_cast<const ulonglong2 *>(a), reinterpret_cast<const ulonglong2 *>(b),
            reinterpret_cast<ulonglong2 *>(out3), total2, c->log_n - 1, limbs, c->d_limb, accumulate ? 1 : 0, 0,
            b_broadcast ? 1 : 0, a_limbs, b_limbs);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ew_square(Context *c, const u64 *a, u64 *out3, long long batch, int limbs)
    {
        const long long total2 = batch * limbs * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        KernelTimer kt6(c, "k_multiply", 1);
        k_multiply<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
            reinterpret_cast<const ulonglong2 *>(a), reinterpret_cast<const ulonglong2 *>(a),
            reinterpret_cast<ulonglong2 *>(out3), total2, c->log_n - 1, limbs, c->d_limb, 0, 1, 0, limbs, limbs);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void sum_batch(Context *c, const u64 *a, u64 *out, long long batch, int polys, int limbs)
    {
        const long long per_ct2 = (long long)polys * limbs * (long long)(c->n / 2);
        KernelTimer kt7(c, "k_sum_batch", 1);
        k_sum_batch<<<grid_for(per_ct2), EW_THREADS, 0, c->stream>>>(reinterpret_cast<const ulonglong2 *>(a),
                                                                   reinterpret_cast<ulonglong2 *>(out), per_ct2, batch,
                                                                   c->log_n - 1, limbs, c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void inner_product(Context *c, const u64 *a, const u64 *b, u64 *out3, long long batch, int limbs, int mode)
    {
        MOAI_REQUIRE(batch <= 1024, "inner product batch too large for the lazy accumulator");
        {
            // the 128-bit lazy sums hold 2 * batch products of two residues: 2 batch q^2 < 2^128 for the widest prime in use
            int bits = 0;
            for (int l = 0; l < limbs; l++)
            {
                int b = 0;
                while ((c->q[l] >> b) != 0)
                {
                    b++;
                }
                bits = b > bits ? b : bits;
            }
            int lb = 0;
            while ((1ll << lb) < 2 * batch)
            {
                lb++;
            }
            MOAI_REQUIRE(2 * bits + lb <= 127, "inner product batch too large for primes of this size");
        }
        const long long total2 = (long long)limbs * (long long)(c->n / 2);
        KernelTimer kt8(c, "k_inner_product", 1);
        k_inner_product<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
            reinterpret_cast<const ulonglong2 *>(a), reinterpret_cast<const ulonglong2 *>(b),
            reinterpret_cast<ulonglong2 *>(out3), batch, c->log_n - 1, limbs, mode, c->d_limb, c->d_two64);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }


    // ------------------------------------------------------------------ BSGS inner sums in the EXTENDED basis
    // Lazy mod-down ("double hoisting"): the baby-step rotations of a linear stage are left in the key-switch basis
    // {q_0..q_{l-1}} + extra primes (rns limbs, scaled by the special modulus P'):
    //     rot_r(ct) * P'  =  sigma_r( acc_r + (P' c0, 0) ),   acc_r = sum_g digit_g (.) K'_{r,g}   (no mod-down),
    // and the inner sums  out_i = sum_r pt[i][r] (.) sigma_r(...)  are formed there, so that a stage pays ONE mod-down
    // per giant step instead of one per baby step.  sigma_r is a gather through the NTT-domain permutation table;
    // the identity rotation contributes P' * ct (zero modulo the extra primes).  128-bit lazy sums like k_bsgs_inner.
    namespace
    {
        struct BsgsExtArgs
        {
            const u64 *acc[BSGS_MAX_BABY];                  // [batch][2][rns][n]; nullptr = the identity rotation
            const uint32_t *perm[BSGS_MAX_BABY];            // [n] permutation of rotation r
            const u64 *pt[BSGS_MAX_GIANT][BSGS_MAX_BABY];   // [rns][n], nullptr = absent diagonal
            u64 *out[BSGS_MAX_GIANT];                       // [batch][2][rns][n]
            const u64 *cP;                                  // [batch][2][n_data][n] = P' * ct
            const int *ids;                                 // [rns] prime index of limb I
            int n_baby, n_giant, rns, n_data, accumulate;
        };

        // value of baby step j at this thread's coefficient pair: sigma_j(acc_j + (P' c0, 0)), or P' ct for the identity
        __device__ __forceinline__ ulonglong2 bsgs_ext_value(const BsgsExtArgs &a, int j, long long bp, int I, long long within,
                                                             int log_n, bool data, bool poly0, u64 q)
        {
            ulonglong2 v = make_ulonglong2(0, 0);
            if (a.acc[j])
            {
                // The NTT-domain automorphism maps the aligned pair (2w, 2w + 1) onto an aligned pair (j, j ^ 1): the two
                // positions differ in the top bit of the bit-reversed index, i.e. by N in the exponent, and N e = N
                // (mod 2N) for odd e.  One 16-byte gather per pair, swapped when the image starts at the odd slot.
                const unsigned ix = __ldg(a.perm[j] + 2 * within);
                const long long at = (long long)(ix >> 1);
                const bool swap = (ix & 1u) != 0;
                const ulonglong2 w = reinterpret_cast<const ulonglong2 *>(a.acc[j] + ((bp * a.rns + I) << log_n))[at];
                v.x = swap ? w.y : w.x;
                v.y = swap ? w.x : w.y;
                if (data && poly0)
                {
                    // P' * c0 of this ciphertext
                    const ulonglong2 z = reinterpret_cast<const ulonglong2 *>(a.cP + (((bp & ~1ll) * a.n_data + I) << log_n))[at];
                    v.x = addmod(v.x, swap ? z.y : z.x, q);
                    v.y = addmod(v.y, swap ? z.x : z.y, q);
                }
            }
            else if (data)
            {
                v = reinterpret_cast<const ulonglong2 *>(a.cP + ((bp * a.n_data + I) << log_n))[within];
            }
            return v;
        }

        // FP64-path primes: exact products on the FP64 pipe (fpfield.cuh), sums reduced every 2 / 8 terms
        template <bool WIDE, int GM>
        __device__ __forceinline__ void bsgs_ext_fp(const FpField<WIDE> &f, const BsgsExtArgs &a, long long bp, int I,
                                                    long long within, int log_n2, bool data, bool poly0)
        {
            // Rotation-outer software pipeline.  ncu of the first version (values of 8 rotations fetched, then used):
            // 17 warps per issue stalled on the scoreboard at 19 % of the DRAM rate — the dependent chain
            // permutation index -> 16-byte gather -> diagonals was exposed.  Now the gathers run two rotations ahead of
            // their use (all indices first: they are L2 hits), the diagonals of a rotation are requested before its
            // value is converted, and the giants' sums stay in registers.
            const int log_n = log_n2 + 1;
            const long long o = ((bp * a.rns + I) << log_n2) + within;
            const long long pt_off = ((long long)I << log_n2) + within;
            const int red_every = WIDE ? 2 : 8;
            const u64 qq = f.pi;
            const bool add_c0 = data && poly0;
            const ulonglong2 *cp0 = reinterpret_cast<const ulonglong2 *>(a.cP + (((bp & ~1ll) * a.n_data + I) << log_n));
            double sx[GM], sy[GM];
#pragma unroll
            for (int g = 0; g < GM; g++)
            {
                sx[g] = 0.0;
                sy[g] = 0.0;
                if (g < a.n_giant && a.accumulate)
                {
                    const ulonglong2 prev = reinterpret_cast<const ulonglong2 *>(a.out[g])[o];
                    sx[g] = f.red(f.in_outer(prev.x));
                    sy[g] = f.red(f.in_outer(prev.y));
                }
            }
            // all permutation indices up front (L2 hits), parked in shared memory: the rotation loop below is NOT
            // unrolled (registers), so they are indexed dynamically
            __shared__ unsigned s_ix[BSGS_MAX_BABY][EW_THREADS];
            unsigned(*ix)[EW_THREADS] = s_ix;
            const int tx = threadIdx.x;
#pragma unroll
            for (int j = 0; j < BSGS_MAX_BABY; j++)
            {
                ix[j][tx] = (j < a.n_baby && a.acc[j]) ? __ldg(a.perm[j] + 2 * within) : 0u;
            }
            // raw operands of rotation j: the gathered pair of acc_j (and of P' c0), or P' * this polynomial for the identity
            auto issue = [&](int j, ulonglong2 &w, ulonglong2 &z) {
                w = make_ulonglong2(0, 0);
                z = make_ulonglong2(0, 0);
                if (j >= a.n_baby)
                {
                    return;
                }
                if (a.acc[j])
                {
                    const long long at = (long long)(ix[j][tx] >> 1);
                    w = reinterpret_cast<const ulonglong2 *>(a.acc[j] + ((bp * a.rns + I) << log_n))[at];
                    if (add_c0)
                    {
                        z = cp0[at];
                    }
                }
                else if (data)
                {
                    w = reinterpret_cast<const ulonglong2 *>(a.cP + ((bp * a.n_data + I) << log_n))[within];
                }
            };
            auto step = [&](int j, ulonglong2 &w, ulonglong2 &z) {
                // diagonals of this rotation first (independent loads), then the value
                ulonglong2 pw[GM];
#pragma unroll
                for (int g = 0; g < GM; g++)
                {
                    pw[g] = make_ulonglong2(0, 0);
                    if (g < a.n_giant && a.pt[g][j])
                    {
                        pw[g] = __ldg(reinterpret_cast<const ulonglong2 *>(a.pt[g][j]) + pt_off);
                    }
                }
                const bool swap = a.acc[j] && (ix[j][tx] & 1u);
                const u64 vx = addmod(swap ? w.y : w.x, swap ? z.y : z.x, qq);
                const u64 vy = addmod(swap ? w.x : w.y, swap ? z.x : z.y, qq);
                const double rx = f.red(f.in_outer(vx)), ry = f.red(f.in_outer(vy)); // centred: |r| <= p/2 + 1
                issue(j + 2, w, z); // this slot is free again: rotation j + 2 goes on its way
#pragma unroll
                for (int g = 0; g < GM; g++)
                {
                    if (g < a.n_giant && a.pt[g][j])
                    {
                        sx[g] = __dadd_rn(sx[g], f.mul_lazy(f.in_outer(pw[g].x), rx));
                        sy[g] = __dadd_rn(sy[g], f.mul_lazy(f.in_outer(pw[g].y), ry));
                    }
                }
                if ((j + 1) % red_every == 0) // at most red_every terms per sum since the last reduction
                {
#pragma unroll
                    for (int g = 0; g < GM; g++)
                    {
                        sx[g] = f.red(sx[g]);
                        sy[g] = f.red(sy[g]);
                    }
                }
            };
            ulonglong2 w0, z0, w1, z1;
            issue(0, w0, z0);
            issue(1, w1, z1);
#pragma unroll 1
            for (int j = 0; j < a.n_baby; j += 2)
            {
                step(j, w0, z0);
                if (j + 1 < a.n_baby)
                {
                    step(j + 1, w1, z1);
                }
            }
#pragma unroll
            for (int g = 0; g < GM; g++)
            {
                if (g < a.n_giant)
                {
                    ulonglong2 res;
                    res.x = f.canon(sx[g]);
                    res.y = f.canon(sy[g]);
                    reinterpret_cast<ulonglong2 *>(a.out[g])[o] = res;
                }
            }
        }

        template <int GM> // giants' sums kept in registers: GM >= n_giant
        __global__ void __launch_bounds__(EW_THREADS) k_bsgs_ext(BsgsExtArgs a, long long polys, int log_n2,
                                                                 const LimbConst *__restrict__ lcs,
                                                                 const Twiddle *__restrict__ two64)
        {
            // CTA order: coefficient block fastest, then (ciphertext, polynomial), limb slowest.  The gathers of one
            // (polynomial, limb) scatter over whole limbs of every acc_r (n_baby x 512 KiB) and the diagonals of one limb
            // (n_giant x n_baby x 512 KiB) are shared by all polynomials: both sets stay L2-resident in this order, so
            // the diagonals are streamed from HBM once per launch and the gathers never leave the L2.
            const int cblks = (1 << log_n2) / EW_THREADS;
            const long long within = (blockIdx.x % cblks) * EW_THREADS + threadIdx.x;
            const long long rest = blockIdx.x / cblks;
            const long long bp = rest % polys; // b * 2 + p
            const int I = (int)(rest / polys);
            const int prime = a.ids[I];
            const LimbConst lc = lcs[prime];
            const bool data = I < a.n_data;
            const bool poly0 = (bp & 1) == 0;
            if (lc.fp_class == 1)
            {
                bsgs_ext_fp<false, GM>(FpField<false>(lc), a, bp, I, within, log_n2, data, poly0);
                return;
            }
            if (lc.fp_class == 2)
            {
                bsgs_ext_fp<true, GM>(FpField<true>(lc), a, bp, I, within, log_n2, data, poly0);
                return;
            }
            // integer-path primes (the 58-bit special prime, 1 limb in 36): 128-bit lazy sums, one giant step at a time
            // (the rotations' values are gathered again for every giant step: few CTAs take this path, and it keeps the
            // kernel's register count at the FP64 path's)
            const Twiddle t64 = two64[prime];
            const long long o = ((bp * a.rns + I) << log_n2) + within;
            const long long pt_off = ((long long)I << log_n2) + within;
            for (int g = 0; g < a.n_giant; g++)
            {
                u128 sx{ 0, 0 }, sy{ 0, 0 };
                if (a.accumulate)
                {
                    const ulonglong2 prev = reinterpret_cast<const ulonglong2 *>(a.out[g])[o];
                    sx.lo = prev.x;
                    sy.lo = prev.y;
                }
                for (int j = 0; j < a.n_baby; j++)
                {
                    if (a.pt[g][j])
                    {
                        const ulonglong2 v = bsgs_ext_value(a, j, bp, I, within, log_n2 + 1, data, poly0, lc.q);
                        const ulonglong2 w = __ldg(reinterpret_cast<const ulonglong2 *>(a.pt[g][j]) + pt_off);
                        mac_wide(sx, v.x, w.x);
                        mac_wide(sy, v.y, w.y);
                    }
                }
                ulonglong2 res;
                res.x = barrett_reduce_acc(sx, lc, t64.w, t64.wq);
                res.y = barrett_reduce_acc(sy, lc, t64.w, t64.wq);
                reinterpret_cast<ulonglong2 *>(a.out[g])[o] = res;
            }
        }
    } // namespace

    void bsgs_ext(Context *c, const u64 *const *acc, const uint32_t *const *perm, int n_baby, const u64 *const *pt,
                  int n_giant, u64 *const *out, const u64 *cP, long long batch, const KsShape &sh, bool accumulate)
    {
        MOAI_REQUIRE(n_baby >= 1 && n_baby <= BSGS_MAX_BABY && n_giant >= 1 && n_giant <= BSGS_MAX_GIANT,
                     "BSGS plan exceeds the fused kernel's limits");
        BsgsExtArgs a;
        a.n_baby = n_baby;
        a.n_giant = n_giant;
        a.rns = sh.rns;
        a.n_data = sh.n_data;
        a.ids = sh.ids;
        a.cP = cP;
        a.accumulate = accumulate ? 1 : 0;
        for (int j = 0; j < n_baby; j++)
        {
            a.acc[j] = acc[j];
            a.perm[j] = perm[j];
        }
        for (int g = 0; g < n_giant; g++)
        {
            a.out[g] = out[g];
            for (int j = 0; j < n_baby; j++)
            {
                a.pt[g][j] = pt[g * n_baby + j];
            }
        }
        const long long ctas = batch * 2 * sh.rns * (long long)((c->n / 2) / EW_THREADS);
        MOAI_REQUIRE((c->n / 2) % EW_THREADS == 0 && ctas < (1ll << 31), "unsupported shape for the fused inner sums");
        KernelTimer kt(c, "k_bsgs_ext", 1);
        if (n_giant == 1)
        {
            k_bsgs_ext<1><<<(unsigned)ctas, EW_THREADS, 0, c->stream>>>(a, batch * 2, c->log_n - 1, c->d_limb, c->d_two64);
        }
        else if (n_giant <= 4)
        {
            k_bsgs_ext<4><<<(unsigned)ctas, EW_THREADS, 0, c->stream>>>(a, batch * 2, c->log_n - 1, c->d_limb, c->d_two64);
        }
        else
        {
            k_bsgs_ext<BSGS_MAX_GIANT><<<(unsigned)ctas, EW_THREADS, 0, c->stream>>>(a, batch * 2, c->log_n - 1, c->d_limb,
                                                                                    c->d_two64);
        }
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    // ------------------------------------------------------------------ single-digit stage (mod-raised input)
    // The first CoeffToSlot stage of the bootstrapping: the mod-raised c1 is ONE small digit (ksgroup.hpp), so the
    // key-switch inner product of rotation r is a single product per limb and is formed on the fly:
    //     out[b][k] = sum_r pt_r (.) sigma_r( ext[b] (.) K'_r[k] + (P c0, 0) )
    //               = sum_r pt_r (.) ( sigma_r(ext[b]) (.) K_r[k] + sigma_r(P c0, 0) ),     K_r in NATURAL order
    // (sigma_r is a ring automorphism: it distributes over the product, and sigma_r(K'_r) = K_r), i.e. one gather of the
    // 36 limbs of ext per rotation instead of materialising 2 x 36 limbs of inner product per rotation and ciphertext.
    namespace
    {
        struct BsgsSingleArgs
        {
            const u64 *ext;                       // [batch][rns][n]
            const u64 *key[BSGS_MAX_BABY];        // [2][key_kl][n] natural order; nullptr = the identity rotation
            const uint32_t *perm[BSGS_MAX_BABY];
            const u64 *pt[BSGS_MAX_BABY];         // [rns][n]
            u64 *out;                             // [batch][2][rns][n]
            const u64 *cP;                        // [batch][2][n_data][n] = P * ct
            const int *ids;
            int n_rot, rns, n_data, key_kl, accumulate;
        };

        template <bool WIDE>
        __device__ __forceinline__ void bsgs_single_fp(const FpField<WIDE> &f, const BsgsSingleArgs &a, long long bp, int I,
                                                       long long within, int log_n2, bool data, bool poly0)
        {
            const int log_n = log_n2 + 1;
            const long long b = bp >> 1;
            const int k = (int)(bp & 1);
            const int key_limb = I < a.n_data ? I : I + a.key_kl - a.rns;
            const long long o = ((bp * a.rns + I) << log_n2) + within;
            const long long lo = ((long long)I << log_n2) + within;
            const ulonglong2 *ext = reinterpret_cast<const ulonglong2 *>(a.ext + ((b * a.rns + I) << log_n));
            const ulonglong2 *cp0 = reinterpret_cast<const ulonglong2 *>(a.cP + (((b * 2) * a.n_data + I) << log_n));
            double sx = 0.0, sy = 0.0;
            if (a.accumulate)
            {
                const ulonglong2 prev = reinterpret_cast<const ulonglong2 *>(a.out)[o];
                sx = f.red(f.in_outer(prev.x));
                sy = f.red(f.in_outer(prev.y));
            }
            const int red_every = WIDE ? 2 : 8;
            int terms = 0;
#pragma unroll 4
            for (int r = 0; r < a.n_rot; r++)
            {
                if (!a.pt[r])
                {
                    continue;
                }
                double tx, ty;
                if (a.key[r])
                {
                    const unsigned ix = __ldg(a.perm[r] + 2 * within);
                    const long long at = (long long)(ix >> 1);
                    const bool swap = (ix & 1u) != 0;
                    const ulonglong2 e = ext[at];
                    const ulonglong2 kv = __ldg(reinterpret_cast<const ulonglong2 *>(a.key[r]) +
                                                (((long long)k * a.key_kl + key_limb) << log_n2) + within);
                    const double ex = f.red(f.in_outer(swap ? e.y : e.x)), ey = f.red(f.in_outer(swap ? e.x : e.y));
                    tx = f.mul_lazy(f.in_outer(kv.x), ex);
                    ty = f.mul_lazy(f.in_outer(kv.y), ey);
                    if (data && poly0)
                    {
                        const ulonglong2 z = cp0[at];
                        tx = __dadd_rn(tx, f.in_outer(swap ? z.y : z.x));
                        ty = __dadd_rn(ty, f.in_outer(swap ? z.x : z.y));
                    }
                    tx = f.red(tx);
                    ty = f.red(ty);
                }
                else
                {
                    if (!data)
                    {
                        continue;
                    }
                    const ulonglong2 z = reinterpret_cast<const ulonglong2 *>(a.cP + ((bp * a.n_data + I) << log_n))[within];
                    tx = f.red(f.in_outer(z.x));
                    ty = f.red(f.in_outer(z.y));
                }
                const ulonglong2 w = __ldg(reinterpret_cast<const ulonglong2 *>(a.pt[r]) + lo);
                sx = __dadd_rn(sx, f.mul_lazy(f.in_outer(w.x), tx));
                sy = __dadd_rn(sy, f.mul_lazy(f.in_outer(w.y), ty));
                if (++terms % red_every == 0)
                {
                    sx = f.red(sx);
                    sy = f.red(sy);
                }
            }
            ulonglong2 res;
            res.x = f.canon(sx);
            res.y = f.canon(sy);
            reinterpret_cast<ulonglong2 *>(a.out)[o] = res;
        }

        __global__ void __launch_bounds__(EW_THREADS) k_bsgs_single(BsgsSingleArgs a, long long polys, int log_n2,
                                                                    const LimbConst *__restrict__ lcs,
                                                                    const Twiddle *__restrict__ two64)
        {
            // CTA order as in k_bsgs_ext: coefficient block fastest, then (ciphertext, polynomial), limb slowest
            const int cblks = (1 << log_n2) / EW_THREADS;
            const long long within = (blockIdx.x % cblks) * EW_THREADS + threadIdx.x;
            const long long rest = blockIdx.x / cblks;
            const long long bp = rest % polys;
            const int I = (int)(rest / polys);
            const int prime = a.ids[I];
            const LimbConst lc = lcs[prime];
            const bool data = I < a.n_data;
            const bool poly0 = (bp & 1) == 0;
            if (lc.fp_class == 1)
            {
                bsgs_single_fp<false>(FpField<false>(lc), a, bp, I, within, log_n2, data, poly0);
                return;
            }
            if (lc.fp_class == 2)
            {
                bsgs_single_fp<true>(FpField<true>(lc), a, bp, I, within, log_n2, data, poly0);
                return;
            }
            // integer path (the special prime)
            const Twiddle t64 = two64[prime];
            const int log_n = log_n2 + 1;
            const long long b = bp >> 1;
            const int k = (int)(bp & 1);
            const int key_limb = I < a.n_data ? I : I + a.key_kl - a.rns;
            const long long o = ((bp * a.rns + I) << log_n2) + within;
            const long long lo = ((long long)I << log_n2) + within;
            const ulonglong2 *ext = reinterpret_cast<const ulonglong2 *>(a.ext + ((b * a.rns + I) << log_n));
            u128 sx{ 0, 0 }, sy{ 0, 0 };
            if (a.accumulate)
            {
                const ulonglong2 prev = reinterpret_cast<const ulonglong2 *>(a.out)[o];
                sx.lo = prev.x;
                sy.lo = prev.y;
            }
            for (int r = 0; r < a.n_rot; r++)
            {
                if (!a.pt[r])
                {
                    continue;
                }
                u64 tx, ty;
                if (a.key[r])
                {
                    const unsigned ix = __ldg(a.perm[r] + 2 * within);
                    const long long at = (long long)(ix >> 1);
                    const bool swap = (ix & 1u) != 0;
                    const ulonglong2 e = ext[at];
                    const ulonglong2 kv = __ldg(reinterpret_cast<const ulonglong2 *>(a.key[r]) +
                                                (((long long)k * a.key_kl + key_limb) << log_n2) + within);
                    tx = mulmod(swap ? e.y : e.x, kv.x, lc);
                    ty = mulmod(swap ? e.x : e.y, kv.y, lc);
                    if (data && poly0)
                    {
                        const ulonglong2 z = reinterpret_cast<const ulonglong2 *>(a.cP + (((b * 2) * a.n_data + I) << log_n))[at];
                        tx = addmod(tx, swap ? z.y : z.x, lc.q);
                        ty = addmod(ty, swap ? z.x : z.y, lc.q);
                    }
                }
                else
                {
                    if (!data)
                    {
                        continue;
                    }
                    const ulonglong2 z = reinterpret_cast<const ulonglong2 *>(a.cP + ((bp * a.n_data + I) << log_n))[within];
                    tx = z.x;
                    ty = z.y;
                }
                const ulonglong2 w = __ldg(reinterpret_cast<const ulonglong2 *>(a.pt[r]) + lo);
                mac_wide(sx, tx, w.x);
                mac_wide(sy, ty, w.y);
            }
            ulonglong2 res;
            res.x = barrett_reduce_acc(sx, lc, t64.w, t64.wq);
            res.y = barrett_reduce_acc(sy, lc, t64.w, t64.wq);
            reinterpret_cast<ulonglong2 *>(a.out)[o] = res;
        }
    } // namespace

    void bsgs_single(Context *c, const u64 *ext, const u64 *const *key, const uint32_t *const *perm, const u64 *const *pt,
                     int n_rot, int key_kl, u64 *out, const u64 *cP, long long batch, const KsShape &sh, bool accumulate)
    {
        MOAI_REQUIRE(n_rot >= 1 && n_rot <= BSGS_MAX_BABY && sh.digits == 1, "bad single-digit stage");
        BsgsSingleArgs a;
        a.ext = ext;
        a.out = out;
        a.cP = cP;
        a.ids = sh.ids;
        a.n_rot = n_rot;
        a.rns = sh.rns;
        a.n_data = sh.n_data;
        a.key_kl = key_kl;
        a.accumulate = accumulate ? 1 : 0;
        for (int r = 0; r < n_rot; r++)
        {
            a.key[r] = key[r];
            a.perm[r] = perm[r];
            a.pt[r] = pt[r];
        }
        const long long ctas = batch * 2 * sh.rns * (long long)((c->n / 2) / EW_THREADS);
        MOAI_REQUIRE((c->n / 2) % EW_THREADS == 0 && ctas < (1ll << 31), "unsupported shape for the fused inner sums");
        KernelTimer kt(c, "k_bsgs_single", 1);
        k_bsgs_single<<<(unsigned)ctas, EW_THREADS, 0, c->stream>>>(a, batch * 2, c->log_n - 1, c->d_limb, c->d_two64);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    // mod-down by the special prime alone (S/evaluator.cpp:2962-3018): in [P][limbs + 1][n] -> out [P][limbs][n]
    void moddown_special(Context *c, const u64 *in, long long P, int limbs, const u64 *addend, u64 *out,
                         bool addend_even_only)
    {
        divide_round_last(c, in, P, limbs + 1, c->kl - 1, addend, out, addend_even_only);
    }

    void bsgs_inner(Context *c, const u64 *const *rot, int n_baby, const u64 *const *pt, int n_giant, u64 *const *out,
                    long long batch, int limbs)
    {
        MOAI_REQUIRE(n_baby >= 1 && n_baby <= BSGS_MAX_BABY && n_giant >= 1 && n_giant <= BSGS_MAX_GIANT,
                     "BSGS plan exceeds the fused kernel's limits");
        BsgsArgs a;
        a.n_baby = n_baby;
        a.n_giant = n_giant;
        for (int j = 0; j < n_baby; j++)
        {
            a.rot[j] = rot[j];
        }
        for (int g = 0; g < n_giant; g++)
        {
            a.out[g] = out[g];
            for (int j = 0; j < n_baby; j++)
            {
                a.pt[g][j] = pt[g * n_baby + j];
            }
        }
        const long long ctas = batch * 2 * limbs * (long long)((c->n / 2) / EW_THREADS);
        MOAI_REQUIRE((c->n / 2) % EW_THREADS == 0 && ctas < (1ll << 31), "unsupported shape for the fused inner sums");
        KernelTimer kt9(c, "k_bsgs_inner", 1);
        k_bsgs_inner<<<(unsigned)ctas, EW_THREADS, 0, c->stream>>>(a, batch * 2, c->log_n - 1, limbs, c->d_limb,
                                                                   c->d_two64);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void rescale(Context *c, const u64 *in, u64 *out, long long batch, int polys, int limbs)
    {
        MOAI_REQUIRE(limbs >= 2, "end of modulus switching chain reached");
        divide_round_last(c, in, batch * polys, limbs, limbs - 1, nullptr, out);
    }

    void mod_switch_drop(Context *c, const u64 *in, u64 *out, long long batch, int polys, int limbs_in, int limbs_out)
    {
        MOAI_REQUIRE(limbs_out >= 1 && limbs_out <= limbs_in, "end of modulus switching chain reached");
        const size_t n = c->n;
        { KernelTimer ktm(c, "k_copy_mod_switch", 1); MOAI_CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)limbs_out * n * sizeof(u64), in,
                                          (size_t)limbs_in * n * sizeof(u64), (size_t)limbs_out * n * sizeof(u64),
                                          (size_t)(batch * polys), cudaMemcpyDeviceToDevice, c->stream)); }
    }

    void mod_raise(Context *c, const u64 *in, u64 *out, long long batch, int polys, int limbs_out)
    {
        const size_t n = c->n;
        const long long P = batch * polys;
        Scratch d(P * n * sizeof(u64), c->stream);
        { KernelTimer ktm(c, "k_copy_modraise", 1); MOAI_CUDA_CHECK(cudaMemcpyAsync(d.p, in, P * n * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream)); }
        ntt_inverse(c, d.as<u64>(), P, c->d_ids, 1);
        const long long total2 = P * limbs_out * (long long)(n / 2);
        k_modraise_expand<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
            d.as<ulonglong2>(), reinterpret_cast<ulonglong2 *>(out), total2, c->log_n - 1, limbs_out, c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
        ntt_forward(c, out, P * limbs_out, c->d_ids, limbs_out);
    }

    void apply_galois_ntt(Context *c, const u64 *in, u64 *out, long long count_polys_limbs, uint32_t elt)
    {
        const uint32_t *table = c->galois_table(elt);
        const long long total = count_polys_limbs * (long long)c->n;
        if (!total)
        {
            return;
        }
        KernelTimer kt10(c, "k_galois", 1);
        k_galois<<<grid_for(total), EW_THREADS, 0, c->stream>>>(in, out, total, c->log_n, table);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    size_t ks_ext_bytes_per_ct(Context *c, int limbs)
    {
        return (size_t)(limbs + 1) * limbs * c->n * sizeof(u64);
    }

    // Digit decomposition of a key switch (S/evaluator.cpp:2805-2856): d_J = INTT_J(target[J]);
    // ext[b][I][J] = NTT_I(d_J mod m_I) for I over {q_0..q_{l-1}, p}.
    // `target_stride` = elements between the target polynomials of consecutive batch items
    // (0 = contiguous; 2 * limbs * n picks c1 out of a batch of size-2 ciphertexts when target = ct + limbs * n).
    void ks_decompose(Context *c, const u64 *target, long long batch, int limbs, u64 *ext, long long target_stride)
    {
        MOAI_REQUIRE(limbs >= 1 && limbs <= c->kl - 1, "limb count out of range");
        const size_t n = c->n;
        const int rns = limbs + 1;
        const int *ids_ks = c->d_ids_ks + (size_t)limbs * (c->kl + 1);
        bool fuse_expand = true; // the FP64 path ingests digits below 2^52 only
        for (int l = 0; l < limbs; l++)
        {
            fuse_expand = fuse_expand && (c->q[l] >> 52) == 0;
        }
        Scratch d((size_t)batch * limbs * n * sizeof(u64), c->stream);
        MOAI_REQUIRE(target_stride % (long long)n == 0, "target stride must be a whole number of limbs");
        ntt_inverse_from(c, target, limbs, target_stride ? target_stride / (long long)n : limbs, d.as<u64>(), batch * limbs,
                         c->d_ids, limbs);
        // NTT_I(d_I mod q_I) reproduces the target limb itself, so the I == J digits need no
        // special case (S/evaluator.cpp:2831-2836 takes the NTT-form input there: same residues)
        if (fuse_expand)
        {
            // digit extension (d_J mod m_I) fused into the NTT's first pass: ext is written once
            NttPrologue pro;
            pro.src = d.as<u64>();
            pro.mode = 1;
            ntt_forward(c, ext, batch * rns * limbs, ids_ks, rns, limbs, &pro);
        }
        else
        {
            const long long total2 = batch * rns * limbs * (long long)(n / 2);
            KernelTimer kt11(c, "k_ks_expand", 1);
            k_ks_expand<<<grid_for(total2), EW_THREADS, 0, c->stream>>>(
                d.as<ulonglong2>(), reinterpret_cast<ulonglong2 *>(ext), total2, c->log_n - 1, limbs, ids_ks, c->d_limb);
            c->launches += 1;
            ntt_forward(c, ext, batch * rns * limbs, ids_ks, rns, limbs);
        }
    }

    // Inner product with the key and mod-down by the special prime (S/evaluator.cpp:2859-3018):
    // out[b][k] = addend[b][k] + round(sum_J ext[b][.][J] (.) key[J][k] / p).  `key_kl` = limbs stored
    // per key polynomial (c->kl for a SEAL-layout key, max_level + 1 for a truncated one).
    void ks_mac_moddown(Context *c, const u64 *ext, long long batch, int limbs, const u64 *ksk, int key_kl,
                        const u64 *addend, bool addend_c0_only, u64 *out)
    {
        MOAI_REQUIRE(key_kl >= limbs + 1 && key_kl <= c->kl, "key does not cover this level");
        const size_t n = c->n;
        const int rns = limbs + 1;
        const int *ids_ks = c->d_ids_ks + (size_t)limbs * (c->kl + 1);
        Scratch acc((size_t)batch * 2 * rns * n * sizeof(u64), c->stream);
        dim3 grid((unsigned)batch, (unsigned)rns, (unsigned)((n / 2) / EW_THREADS));
        KernelTimer kt12(c, "k_ks_mac", 1);
        k_ks_mac<<<grid, EW_THREADS, 0, c->stream>>>(reinterpret_cast<const ulonglong2 *>(ext),
                                                     reinterpret_cast<const ulonglong2 *>(ksk), acc.as<ulonglong2>(),
                                                     batch, c->log_n - 1, limbs, rns, limbs, key_kl, ids_ks, c->d_limb,
                                                     c->d_two64, 0);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
        divide_round_last(c, acc.as<u64>(), batch * 2, rns, c->kl - 1, addend, out, addend_c0_only);
    }

    // One complete key switch of a chunk with the fused kernel: INTT, pass A of the extended digits
    // (digit extension fused into its loads), then pass B + inner product in one kernel for the
    // FP64-path moduli; integer-path moduli (the 58-bit special prime) finish with the plain pass B
    // and the integer inner-product kernel.  Same residues as ks_decompose + ks_mac_moddown.
    static void ks_fused(Context *c, const u64 *target, long long batch, int limbs, u64 *ext, const u64 *ksk,
                         int key_kl, const u64 *addend, u64 *out, long long target_stride = 0,
                         bool addend_c0_only = false, int addend_group = 2)
    {
        const size_t n = c->n;
        const int rns = limbs + 1;
        const int *ids_ks = c->d_ids_ks + (size_t)limbs * (c->kl + 1);
        Scratch d((size_t)batch * limbs * n * sizeof(u64), c->stream);
        MOAI_REQUIRE(target_stride % (long long)n == 0, "target stride must be a whole number of limbs");
        ntt_inverse_from(c, target, limbs, target_stride ? target_stride / (long long)n : limbs, d.as<u64>(), batch * limbs,
                         c->d_ids, limbs);
        NttPrologue pro;
        pro.src = d.as<u64>();
        pro.mode = 1;
        ntt_forward(c, ext, batch * rns * limbs, ids_ks, rns, limbs, &pro, /*passes=*/1);
        Scratch acc((size_t)batch * 2 * rns * n * sizeof(u64), c->stream);
        ks_passb_mac(c, ext, batch, ks_shape_seal(c, limbs), ksk, key_kl, acc.as<u64>());
        for (int I = 0; I < rns; I++)
        {
            const int prime = I == limbs ? c->kl - 1 : I;
            if (c->h_limb[prime].fp_class != 0)
            {
                continue;
            }
            ntt_forward_pass_b_strided(c, ext + (size_t)I * limbs * n, batch, limbs, (long long)rns * limbs,
                                       c->d_ids + prime);
            dim3 grid((unsigned)batch, 1u, (unsigned)((n / 2) / EW_THREADS));
            KernelTimer kt13(c, "k_ks_mac", 1);
            k_ks_mac<<<grid, EW_THREADS, 0, c->stream>>>(reinterpret_cast<const ulonglong2 *>(ext),
                                                         reinterpret_cast<const ulonglong2 *>(ksk),
                                                         acc.as<ulonglong2>(), batch, c->log_n - 1, limbs, rns, limbs,
                                                         key_kl, ids_ks, c->d_limb, c->d_two64, I);
            c->launches += 1;
            MOAI_CUDA_CHECK(cudaGetLastError());
        }
        divide_round_last(c, acc.as<u64>(), batch * 2, rns, c->kl - 1, addend, out, addend_c0_only, addend_group);
    }

    static bool ks_can_fuse(Context *c, int limbs);

    // One rotation with a pre-permuted key and its own digit decomposition (giant steps, conjugation):
    // out = sigma((c0, 0) + keyswitch(c1; K')), through the fused key-switch kernel when the primes allow.
    void rotate_prepermuted(Context *c, const u64 *ct, long long batch, int limbs, uint32_t elt, const u64 *ksk_pre,
                            int key_kl, u64 *out, int k_extra)
    {
        const size_t n = c->n;
        const size_t per_ct = (size_t)2 * limbs * n;
        if (k_extra > 0)
        {
            Scratch tmpg((size_t)batch * per_ct * sizeof(u64), c->stream);
            ksg_switch(c, ct + (size_t)limbs * n, batch, limbs, k_extra, ksk_pre, key_kl, ct, tmpg.as<u64>(),
                       (long long)per_ct, true);
            apply_galois_ntt(c, tmpg.as<u64>(), out, batch * 2 * limbs, elt);
            return;
        }
        const long long chunk = ks_chunk(c, limbs, batch, ks_ext_budget());
        const bool fused = ks_can_fuse(c, limbs);
        Scratch ext((size_t)chunk * ks_ext_bytes_per_ct(c, limbs), c->stream);
        Scratch tmp((size_t)chunk * per_ct * sizeof(u64), c->stream);
        for (long long b0 = 0; b0 < batch; b0 += chunk)
        {
            const long long nb = (batch - b0) < chunk ? (batch - b0) : chunk;
            const u64 *src = ct + (size_t)b0 * per_ct;
            if (fused)
            {
                ks_fused(c, src + (size_t)limbs * n, nb, limbs, ext.as<u64>(), ksk_pre, key_kl, src, tmp.as<u64>(),
                         (long long)per_ct, true);
                apply_galois_ntt(c, tmp.as<u64>(), out + (size_t)b0 * per_ct, nb * 2 * limbs, elt);
            }
            else
            {
                ks_decompose(c, src + (size_t)limbs * n, nb, limbs, ext.as<u64>(), (long long)per_ct);
                rotate_hoisted(c, src, ext.as<u64>(), nb, limbs, elt, ksk_pre, key_kl, out + (size_t)b0 * per_ct);
            }
        }
    }

    static bool ks_can_fuse(Context *c, int limbs)
    {
        static const bool enabled = [] {
            const char *e = getenv("MOAI_KS_FUSED");
            return !e || atoi(e) != 0;
        }();
        if (!enabled)
        {
            return false;
        }
        bool any_fp = false;
        for (int l = 0; l < limbs; l++)
        {
            if ((c->q[l] >> 52) != 0)
            {
                return false; // digits must enter the FP64 prologue below 2^52
            }
            any_fp = any_fp || c->h_limb[l].fp_class != 0;
        }
        return any_fp;
    }

    // workspace budget for the extended digits of one key-switch chunk (MOAI_KS_EXT_GIB, default 16: a BSGS stage at
    // 34 limbs then runs 16 instead of 4 ciphertexts per pass — layer 46.1 -> 44.7 s; the board has 180 GB)
    size_t ks_ext_budget()
    {
        static const size_t budget = [] {
            const char *e = getenv("MOAI_KS_EXT_GIB");
            const double gib = e ? atof(e) : 16.0;
            return (size_t)((gib > 0.25 ? gib : 0.25) * 1073741824.0);
        }();
        return budget;
    }

    long long ks_chunk(Context *c, int limbs, long long batch, size_t budget_bytes)
    {
        long long chunk = (long long)(budget_bytes / ks_ext_bytes_per_ct(c, limbs));
        chunk = chunk < 1 ? 1 : (chunk > batch ? batch : chunk);
        const long long parts = (batch + chunk - 1) / chunk;
        return (batch + parts - 1) / parts; // equal chunks, no short tail
    }

    void switch_key(Context *c, u64 *ct, const u64 *target, long long batch, int limbs, const u64 *ksk, int key_kl,
                    int k_extra)
    {
        MOAI_REQUIRE(limbs >= 1 && limbs <= c->kl - 1, "limb count out of range");
        const size_t n = c->n;
        if (k_extra > 0)
        {
            ksg_switch(c, target, batch, limbs, k_extra, ksk, key_kl, ct, ct, 0, false);
            return;
        }
        if (key_kl <= 0)
        {
            key_kl = c->kl;
        }
        // bound the extended-digit workspace (batch chunking); the evk is streamed once per chunk
        const long long chunk = ks_chunk(c, limbs, batch, ks_ext_budget());
        const bool fused = ks_can_fuse(c, limbs);
        Scratch ext((size_t)chunk * ks_ext_bytes_per_ct(c, limbs), c->stream);
        for (long long b0 = 0; b0 < batch; b0 += chunk)
        {
            const long long nb = (batch - b0) < chunk ? (batch - b0) : chunk;
            u64 *ctb = ct + (size_t)b0 * 2 * limbs * n;
            if (fused)
            {
                ks_fused(c, target + (size_t)b0 * limbs * n, nb, limbs, ext.as<u64>(), ksk, key_kl, ctb, ctb);
            }
            else
            {
                ks_decompose(c, target + (size_t)b0 * limbs * n, nb, limbs, ext.as<u64>(), 0);
                ks_mac_moddown(c, ext.as<u64>(), nb, limbs, ksk, key_kl, ctb, false, ctb);
            }
        }
    }

    // Hoisted rotation (fast mode; not SEAL's residues, same plaintext up to rounding noise):
    //   out = sigma( (c0, 0) + ModDown( sum_J ext_J (.) K'_J ) ),  K' = sigma^-1(K) (key_prepare),
    // where ext is the digit decomposition of the UNROTATED c1 (shared by every rotation of the
    // same ciphertext).  sigma commutes with the NTT-domain products and, up to the sign of the
    // rounding, with the mod-down, so the automorphism is applied once at the end.
    void rotate_hoisted(Context *c, const u64 *ct, const u64 *ext, long long batch, int limbs, uint32_t elt,
                        const u64 *ksk_pre, int key_kl, u64 *out)
    {
        const size_t n = c->n;
        Scratch tmp((size_t)batch * 2 * limbs * n * sizeof(u64), c->stream);
        ks_mac_moddown(c, ext, batch, limbs, ksk_pre, key_kl, ct, true, tmp.as<u64>());
        apply_galois_ntt(c, tmp.as<u64>(), out, batch * 2 * limbs, elt);
    }

    // Up to KSM_R hoisted rotations of the same ciphertexts in one pass over the extended digits (FP64 inner
    // products, csrc/ntt.cu); each result then gets its own mod-down and permutation.
    void rotate_hoisted_multi(Context *c, const u64 *ct, const u64 *ext, long long batch, int limbs, int n_rot,
                              const uint32_t *elts, const u64 *const *ksk_pre, const int *key_kl, u64 *const *outs)
    {
        MOAI_REQUIRE(n_rot >= 1 && n_rot <= KSM_R, "too many rotations for one multi-key pass");
        const size_t n = c->n;
        const int rns = limbs + 1;
        const int *ids_ks = c->d_ids_ks + (size_t)limbs * (c->kl + 1);
        const size_t acc_words = (size_t)batch * 2 * rns * n;
        Scratch acc(acc_words * n_rot * sizeof(u64), c->stream);
        u64 *accp[KSM_R];
        for (int r = 0; r < n_rot; r++)
        {
            accp[r] = acc.as<u64>() + acc_words * r;
            MOAI_REQUIRE(key_kl[r] >= limbs + 1 && key_kl[r] <= c->kl, "key does not cover this level");
        }
        ks_mac_multi(c, ext, batch, ks_shape_seal(c, limbs), n_rot, ksk_pre, key_kl, accp);
        for (int I = 0; I < rns; I++)
        {
            const int prime = I == limbs ? c->kl - 1 : I;
            if (c->h_limb[prime].fp_class != 0)
            {
                continue;
            }
            for (int r = 0; r < n_rot; r++) // integer-path modulus (the special prime)
            {
                dim3 grid((unsigned)batch, 1u, (unsigned)((n / 2) / EW_THREADS));
                KernelTimer kt14(c, "k_ks_mac", 1);
                k_ks_mac<<<grid, EW_THREADS, 0, c->stream>>>(reinterpret_cast<const ulonglong2 *>(ext),
                                                             reinterpret_cast<const ulonglong2 *>(ksk_pre[r]),
                                                             reinterpret_cast<ulonglong2 *>(accp[r]), batch, c->log_n - 1,
                                                             limbs, rns, limbs, key_kl[r], ids_ks, c->d_limb, c->d_two64, I);
                c->launches += 1;
            }
            MOAI_CUDA_CHECK(cudaGetLastError());
        }
        Scratch tmp((size_t)batch * 2 * limbs * n * sizeof(u64), c->stream);
        for (int r = 0; r < n_rot; r++)
        {
            divide_round_last(c, accp[r], batch * 2, rns, c->kl - 1, ct, tmp.as<u64>(), true);
            apply_galois_ntt(c, tmp.as<u64>(), outs[r], batch * 2 * limbs, elts[r]);
        }
    }

    bool ks_multi_enabled(Context *c, int limbs)
    {
        static const bool multi = [] {
            const char *e = getenv("MOAI_KSM_MULTI");
            return !e || atoi(e) != 0;
        }();
        return multi && ks_can_fuse(c, limbs);
    }

    // K' = sigma_elt^-1(K) restricted to `max_limbs` digits / data limbs (+ the special prime):
    // in  [kl-1][2][kl][n] (SEAL layout, S/kswitchkeys.h:335-340), out [max_limbs][2][max_limbs+1][n].
    // pre_permute = false only truncates.
    void key_prepare(Context *c, const u64 *in, uint32_t elt, int max_limbs, bool pre_permute, u64 *out)
    {
        MOAI_REQUIRE(max_limbs >= 1 && max_limbs <= c->kl - 1, "max_limbs out of range");
        MOAI_REQUIRE(in != out, "key_prepare is out of place");
        const size_t n = c->n;
        const int okl = max_limbs + 1;
        uint32_t inv = 1;
        if (pre_permute)
        {
            // inverse of the odd element modulo 2N (Newton iteration doubles the correct bits)
            const uint64_t m = 2 * (uint64_t)n;
            uint64_t x = elt;
            for (int i = 0; i < 6; i++)
            {
                x = (x * (2 + m * 4 - (uint64_t)elt * x % m)) % m;
            }
            MOAI_REQUIRE((uint64_t)elt * x % m == 1, "Galois element is not invertible");
            inv = (uint32_t)x;
        }
        for (int J = 0; J < max_limbs; J++)
        {
            for (int k = 0; k < 2; k++)
            {
                const u64 *src = in + ((size_t)J * 2 + k) * c->kl * n;
                u64 *dst = out + ((size_t)J * 2 + k) * okl * n;
                if (pre_permute)
                {
                    apply_galois_ntt(c, src, dst, max_limbs, inv);
                    apply_galois_ntt(c, src + (size_t)(c->kl - 1) * n, dst + (size_t)max_limbs * n, 1, inv);
                }
                else
                {
                    { KernelTimer ktm(c, "k_copy_key_prepare", 1); MOAI_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)max_limbs * n * sizeof(u64),
                                                    cudaMemcpyDeviceToDevice, c->stream)); }
                    { KernelTimer ktm(c, "k_copy_key_prepare", 1); MOAI_CUDA_CHECK(cudaMemcpyAsync(dst + (size_t)max_limbs * n, src + (size_t)(c->kl - 1) * n,
                                                    n * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream)); }
                }
            }
        }
    }

    void relinearize_rescale(Context *c, const u64 *in3, u64 *out2, long long batch, int limbs, const u64 *ksk, int key_kl,
                             int k_extra)
    {
        // fast mode, grouped keys only: out2[batch][2][limbs - 1][n] = rescale(relinearize(in3)) with the rescale taken
        // inside the key switch's mod-down (ksg_moddown_rescale)
        MOAI_REQUIRE(k_extra > 0 && limbs >= 2, "merged relinearize + rescale needs a grouped-digit key");
        const size_t poly = (size_t)limbs * c->n;
        ksg_switch(c, in3 + 2 * poly, batch, limbs, k_extra, ksk, key_kl, in3, out2, (long long)(3 * poly), false, 3,
                   /*rescale=*/true);
    }

    void relinearize(Context *c, const u64 *in3, u64 *out2, long long batch, int limbs, const u64 *ksk, int key_kl,
                     int k_extra)
    {
        const size_t n = c->n;
        const size_t poly = (size_t)limbs * n;
        if (key_kl <= 0)
        {
            key_kl = c->kl;
        }
        // out2 = (c0, c1) + keyswitch(c2): the fused / grouped paths read c2 and add (c0, c1) straight out of the
        // size-3 input (strided target, addend_group = 3) instead of splitting it into two copies first
        if (k_extra > 0)
        {
            ksg_switch(c, in3 + 2 * poly, batch, limbs, k_extra, ksk, key_kl, in3, out2, (long long)(3 * poly), false, 3);
            return;
        }
        if (ks_can_fuse(c, limbs))
        {
            const long long chunk = ks_chunk(c, limbs, batch, ks_ext_budget());
            Scratch ext((size_t)chunk * ks_ext_bytes_per_ct(c, limbs), c->stream);
            for (long long b0 = 0; b0 < batch; b0 += chunk)
            {
                const long long nb = (batch - b0) < chunk ? (batch - b0) : chunk;
                ks_fused(c, in3 + (size_t)b0 * 3 * poly + 2 * poly, nb, limbs, ext.as<u64>(), ksk, key_kl,
                         in3 + (size_t)b0 * 3 * poly, out2 + (size_t)b0 * 2 * poly, (long long)(3 * poly), false, 3);
            }
            return;
        }
        // out2 <- (c0, c1) ; target <- c2
        { KernelTimer ktm(c, "k_copy_relin_c0c1", 1); MOAI_CUDA_CHECK(cudaMemcpy2DAsync(out2, 2 * poly * sizeof(u64), in3, 3 * poly * sizeof(u64),
                                          2 * poly * sizeof(u64), (size_t)batch, cudaMemcpyDeviceToDevice, c->stream)); }
        Scratch tg((size_t)batch * poly * sizeof(u64), c->stream);
        { KernelTimer ktm(c, "k_copy_relin_c2", 1); MOAI_CUDA_CHECK(cudaMemcpy2DAsync(tg.p, poly * sizeof(u64), in3 + 2 * poly, 3 * poly * sizeof(u64),
                                          poly * sizeof(u64), (size_t)batch, cudaMemcpyDeviceToDevice, c->stream)); }
        switch_key(c, out2, tg.as<u64>(), batch, limbs, ksk, key_kl, k_extra);
    }

    void apply_galois(Context *c, const u64 *in, u64 *out, long long batch, int limbs, uint32_t elt, const u64 *ksk,
                      int key_kl)
    {
        const size_t n = c->n;
        const size_t poly = (size_t)limbs * n;
        MOAI_REQUIRE(in != out, "apply_galois is out of place");
        // permute both polys into a scratch pair, then out = (sigma(c0), 0) + keyswitch(sigma(c1))
        Scratch perm((size_t)batch * 2 * poly * sizeof(u64), c->stream);
        apply_galois_ntt(c, in, perm.as<u64>(), batch * 2 * limbs, elt);
        Scratch tg((size_t)batch * poly * sizeof(u64), c->stream);
        { KernelTimer ktm(c, "k_copy_galois", 1); MOAI_CUDA_CHECK(cudaMemcpy2DAsync(tg.p, poly * sizeof(u64), perm.as<u64>() + poly, 2 * poly * sizeof(u64),
                                          poly * sizeof(u64), (size_t)batch, cudaMemcpyDeviceToDevice, c->stream)); }
        { KernelTimer ktm(c, "k_copy_galois", 1); MOAI_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)batch * 2 * poly * sizeof(u64), c->stream)); }
        { KernelTimer ktm(c, "k_copy_galois", 1); MOAI_CUDA_CHECK(cudaMemcpy2DAsync(out, 2 * poly * sizeof(u64), perm.p, 2 * poly * sizeof(u64),
                                          poly * sizeof(u64), (size_t)batch, cudaMemcpyDeviceToDevice, c->stream)); }
        switch_key(c, out, tg.as<u64>(), batch, limbs, ksk, key_kl);
    }
} // namespace moai
