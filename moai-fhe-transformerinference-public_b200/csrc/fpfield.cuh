// Exact modular arithmetic on the FP64 pipe for primes below 2^51 (shared by the NTT kernels, csrc/ntt.cu, and the
// FP64 inner sums of csrc/ops.cu).  B200 has no 64-bit integer multiplier; a DFMA-based modular product retires ~2.7x
// faster than the 128-bit integer one (tools/microbench.cu).
#pragma once
#include "ntt.cuh"

namespace moai
{
    // ---- exact FP64 path for primes p < 2^51.
    // Residues are integer-valued doubles.  With M = 1.5 * 2^52, rnd(y) = (y + M) - M is
    // round-to-nearest-integer for |y| < 2^51; quot(x) = fma(x, 1/p, M) - M = rnd(x/p).
    //   red(x)      = x - rnd(x/p) p                 : |x| < 2^53  ->  |red| <= p/2 + 1
    //   mul(a, w)   : h = fl(a w), l = fma(a, w, -h) (so a w = h + l exactly),
    //                 r = fma(-rnd(h/p), p, h) + l   : needs |a| < 2^52, |w| <= p/2;
    //                 |r| <= 1.125 p for p < 2^51 and <= 0.52 p for p < 2^48 (error terms: rounding of
    //                 h/p, of 1/p, and |l| <= ulp(h)/2); r is an exact integer because h - t p is an
    //                 integer below 2^53.
    // WIDE  (2^48 <= p, 2p + 64 < 2^52).  Forward: products are left unreduced (|v| <= 1.125 p) and
    //        the 16 registers are reduced after every second stage: 0.5p -> 1.625p -> 2.75p (< 2^53),
    //        multiplier inputs <= 1.625p < 2^52.  Inverse: mul() is followed by red() (|v| <= p/2 + 1)
    //        and every sum is reduced.
    // NARROW (p < 2^48): 32p of headroom below 2^53; no intermediate reductions in the forward
    //        transform (|x| <= 2p + 8 * 0.52p per pass); in the inverse, sums double for at most 4
    //        stages between the phase reductions.
    template <bool WIDE>
    struct FpField
    {
        typedef double elem;
        typedef double tw_t;
        double p, pinv, inv_n, inv_n_w, pshift;
        u64 pi;
        const double *__restrict__ tab;

        __device__ FpField(const NttArgs &a, int limb, const LimbConst &lc)
            : p(lc.pd), pinv(lc.pinv), inv_n(lc.inv_n_d), inv_n_w(lc.inv_n_w_d),
              pshift(lc.pd + 4503599627370496.0), pi(lc.q), tab(a.tw_fp + ((size_t)limb << a.log_n))
        {}
        // element-wise use (no twiddle table)
        __device__ explicit FpField(const LimbConst &lc)
            : p(lc.pd), pinv(lc.pinv), inv_n(lc.inv_n_d), inv_n_w(lc.inv_n_w_d),
              pshift(lc.pd + 4503599627370496.0), pi(lc.q), tab(nullptr)
        {}
        __device__ __forceinline__ tw_t tw(size_t idx) const { return __ldg(tab + idx); }
        __device__ __forceinline__ elem pro_reduce(elem x) const { return red(x); }
        // nearest integer to x * pinv: the product is folded into the magic-constant addition (one
        // FMA, one rounding fewer than mul + add), valid for |x * pinv| < 2^51
        __device__ __forceinline__ double quot(double x) const
        {
            const double M = 6755399441055744.0;
            return __dadd_rn(__fma_rn(x, pinv, M), -M);
        }
        __device__ __forceinline__ double red(double x) const
        {
            return __fma_rn(-quot(x), p, x);
        }
        // a w mod p without the final reduction: |result| <= 1.125 p (WIDE) / 0.65 p (NARROW)
        __device__ __forceinline__ double mul_lazy(double a, double w) const
        {
            const double h = __dmul_rn(a, w);
            const double l = __fma_rn(a, w, -h);
            return __dadd_rn(__fma_rn(-quot(h), p, h), l);
        }
        __device__ __forceinline__ double mul(double a, double w) const
        {
            const double r = mul_lazy(a, w);
            return WIDE ? red(r) : r;
        }
        // canonical / lazy uint64 below 2^52 -> double, exactly (bit trick, no I2F)
        __device__ __forceinline__ elem in_outer(u64 v) const
        {
            return __dadd_rn(__longlong_as_double((long long)(v | 0x4330000000000000ull)), -4503599627370496.0);
        }
        __device__ __forceinline__ elem in_mid(u64 v) const { return __longlong_as_double((long long)v); }
        __device__ __forceinline__ u64 out_mid(elem x) const { return (u64)__double_as_longlong(x); }
        // |red(x)| <= p/2 + 1: shift into the positive range in FP64 (one exact add), then finish
        // with integer compares on the otherwise idle ALU pipe.  2^52 + p + r is an integer in
        // [2^52, 2^53), so its mantissa field is p + r exactly.
        __device__ __forceinline__ u64 canon(elem x) const
        {
            const double r = __dadd_rn(red(x), pshift);
            const u64 v = (u64)__double_as_longlong(r) & 0x000FFFFFFFFFFFFFull; // p + red(x) in (p/2 - 2, 3p/2 + 2)
            return v >= pi ? v - pi : v;
        }
        __device__ __forceinline__ u64 out_fwd(elem x) const { return canon(x); }
        __device__ __forceinline__ u64 out_inv(elem x) const { return canon(x); }
        // the residue as a CENTRED double (|r| <= p/2 + 1), bit pattern: what the base-conversion prologue multiplies
        __device__ __forceinline__ u64 out_fp(elem x) const { return (u64)__double_as_longlong(red(x)); }
        // forward: only the WIDE class needs the per-phase reduction; inverse: only NARROW does
        // (WIDE reduces every sum inside gs()).
        __device__ __forceinline__ void phase_begin_fwd(elem (&x)[16]) const
        {
            if (WIDE)
            {
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = red(x[k]);
                }
            }
        }
        __device__ __forceinline__ void phase_begin_inv(elem (&x)[16]) const
        {
            if (!WIDE)
            {
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = red(x[k]);
                }
            }
        }
        // forward only: the WIDE class no longer reduces every product; values grow by <= 1.125 p per
        // stage from <= p/2 + 1 and all 16 registers are reduced after every second stage
        // (phase_begin_fwd / phase_mid_fwd), so multiplier inputs stay <= 1.625 p < 2^52 and sums
        // <= 2.75 p < 2^53: 2 x 16 reductions per four stages instead of 16 + 32.
        __device__ __forceinline__ void phase_mid_fwd(elem (&x)[16]) const
        {
            phase_begin_fwd(x);
        }
        __device__ __forceinline__ void ct(elem &x, elem &y, const tw_t &w) const
        {
            const double v = mul_lazy(y, w);
            const double u = x;
            x = __dadd_rn(u, v);
            y = __dadd_rn(u, -v);
        }
        __device__ __forceinline__ void gs(elem &x, elem &y, const tw_t &w) const
        {
            const double u = x, v = y;
            const double s = __dadd_rn(u, v);
            x = WIDE ? red(s) : s;
            y = mul(__dadd_rn(u, -v), w);
        }
        __device__ __forceinline__ void gs_last(elem &x, elem &y) const
        {
            const double u = x, v = y;
            x = mul(__dadd_rn(u, v), inv_n);
            y = mul(__dadd_rn(u, -v), inv_n_w);
        }
    };
} // namespace moai
