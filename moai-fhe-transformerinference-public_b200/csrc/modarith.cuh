// Device-side modular arithmetic for RNS limbs (unsigned 64-bit residues, primes < 2^61).
//
// Every public operation of the reference returns canonical residues in [0, q)
// (SURVEY Appendix D), so the internal reduction strategy is free as long as the final value
// is the same residue.  We use
//   * Shoup multiplication by a precomputed (operand, quotient) pair for twiddles and other
//     per-limb constants   — same function as S/util/uintarithsmallmod.h:313-326;
//   * a single-word Barrett reduction of 128-bit products for data x data products
//     (replaces S/util/uintarithsmallmod.h:166-204, same result after canonicalisation).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace moai
{
    typedef unsigned long long u64;
    typedef unsigned int u32;

    // Per-prime constants kept in device memory (one entry per key-level limb).
    struct LimbConst
    {
        u64 q;         // the prime
        u64 two_q;     // 2q
        u64 bar_m;     // floor(2^(64+shift) / q), shift = bits(q) - 1
        u32 bar_shift; // bits(q) - 1
        u32 pad;
        u64 inv_n;     // N^-1 mod q (Shoup operand)
        u64 inv_n_quo; // Shoup quotient of inv_n
        u64 inv_n_w;   // (last inverse-NTT root * N^-1) mod q
        u64 inv_n_w_quo;
        u64 q0_mod;    // q_0 mod q  (ModRaise correction)
        // FP64 NTT path (csrc/ntt.cu): 0 = integer path, 1 = p < 2^48, 2 = p < 2^51
        double pd, pinv;
        double inv_n_d, inv_n_w_d; // symmetric representatives of inv_n, inv_n_w
        int fp_class;
        int pad2;
    };

    struct u128
    {
        u64 lo, hi;
    };

    __device__ __forceinline__ u128 mul_wide(u64 a, u64 b)
    {
        u128 r;
        r.lo = a * b;
        r.hi = __umul64hi(a, b);
        return r;
    }

    // acc += a * b   (128-bit accumulate; caller guarantees no overflow)
    __device__ __forceinline__ void mac_wide(u128 &acc, u64 a, u64 b)
    {
        u64 lo = a * b;
        u64 hi = __umul64hi(a, b);
        asm("add.cc.u64 %0, %0, %2;\n\t"
            "addc.u64 %1, %1, %3;"
            : "+l"(acc.lo), "+l"(acc.hi)
            : "l"(lo), "l"(hi));
    }

    // x * w mod q in [0, 2q): w < q with Shoup quotient wq = floor(w * 2^64 / q); any 64-bit x.
    __device__ __forceinline__ u64 mul_shoup_lazy(u64 x, u64 w, u64 wq, u64 q)
    {
        u64 hi = __umul64hi(x, wq);
        return w * x - hi * q;
    }

    __device__ __forceinline__ u64 mul_shoup(u64 x, u64 w, u64 wq, u64 q)
    {
        u64 r = mul_shoup_lazy(x, w, wq, q);
        return r >= q ? r - q : r;
    }

    __device__ __forceinline__ u64 csub(u64 x, u64 q)
    {
        return x >= q ? x - q : x;
    }

    // Reduce a 128-bit value z < q * 2^64-ish (z < 2^(2*bits(q)) suffices, and more generally
    // z >> shift must fit 64 bits) to [0, q).  t = floor((z >> s) * M / 2^64) underestimates
    // floor(z / q) by at most 2, so two conditional subtractions finish the job.
    __device__ __forceinline__ u64 barrett_reduce_wide(u128 z, const LimbConst &c)
    {
        u64 zs = (z.lo >> c.bar_shift) | (z.hi << (64 - c.bar_shift));
        u64 t = __umul64hi(zs, c.bar_m);
        u64 r = z.lo - t * c.q;
        r = csub(r, c.two_q);
        return csub(r, c.q);
    }

    // General 128-bit reduction for lazily accumulated sums (z arbitrary < 2^128): first fold the
    // high word, then finish with barrett_reduce_wide.  Used by the multiply-accumulate kernels.
    __device__ __forceinline__ u64 barrett_reduce_acc(u128 z, const LimbConst &c, u64 two64_mod_q, u64 two64_mod_q_quo)
    {
        // z = hi * 2^64 + lo  ->  hi' = hi * (2^64 mod q) (lazy, < 2q) ; then lo + hi' < 2^64 + 2q
        u64 h = mul_shoup_lazy(z.hi, two64_mod_q, two64_mod_q_quo, c.q); // [0, 2q)
        u128 s;
        s.lo = z.lo + h;
        s.hi = s.lo < h ? 1 : 0;
        return barrett_reduce_wide(s, c);
    }

    __device__ __forceinline__ u64 mulmod(u64 a, u64 b, const LimbConst &c)
    {
        return barrett_reduce_wide(mul_wide(a, b), c);
    }

    // x mod q for any 64-bit x
    __device__ __forceinline__ u64 reduce64(u64 x, const LimbConst &c)
    {
        u128 z;
        z.lo = x;
        z.hi = 0;
        return barrett_reduce_wide(z, c);
    }

    __device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q)
    {
        u64 s = a + b;
        return s >= q ? s - q : s;
    }

    __device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q)
    {
        return a >= b ? a - b : a + q - b;
    }

    __device__ __forceinline__ u64 negmod(u64 a, u64 q)
    {
        return a ? q - a : 0;
    }
} // namespace moai
