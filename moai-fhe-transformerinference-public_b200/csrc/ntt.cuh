// Negacyclic NTT / INTT over RNS limbs for sm_100a — two-pass, register-blocked radix-16.
//
// Computes exactly the function of the reference's Harvey transforms
// (forward: DWTHandler::transform_to_rev, S/util/dwthandler.h:94-191, natural -> bit-reversed;
//  inverse: transform_from_rev, S/util/dwthandler.h:202-356, bit-reversed -> natural with N^-1
//  folded into the last stage), with the tables of NTTTables::initialize (S/util/ntt.cpp:241-300).
//
// B200 mapping (N = 2^logN, viewed as R x 256 with R = N/256; i = a*256 + b):
//   pass A : the logN-8 stages that pair rows (gap >= 256).  One CTA owns all R rows of 16
//            adjacent columns (R x 128 B tile, every global access a full 128-byte line); each
//            thread keeps 16 residues in registers and runs 4 stages without touching memory,
//            one shared-memory transpose re-blocks the tile for the remaining stages.
//   pass B : the 8 stages inside a 256-element row.  One CTA owns 16 rows; 4 + 4 register
//            stages with one padded (conflict-free) shared-memory transpose.
// A limb (512 KiB at N = 65536) exceeds one SM's shared memory, hence two passes: 2 MiB of
// traffic per limb-transform against the 1 MiB algorithmic floor; the pass-A -> pass-B
// intermediate of a batch stays L2-resident when the batch fits in the 126 MB L2.
// Butterflies use lazy Harvey arithmetic in [0, 4q) with Shoup twiddles (one 16-byte load per
// twiddle); the ALU cost (64-bit mul-hi emulated by IMAD.WIDE) is what bounds this kernel.
#pragma once
#include "context.hpp"

namespace moai
{
    struct NttArgs
    {
        u64 *data;             // [count][n]
        const Twiddle *tw;     // forward or inverse table, [kl][n]
        const LimbConst *limb; // [kl]
        const int *limb_ids;   // [period]; poly p uses prime limb_ids[(p / div) % period]
        int period;
        int div;
        int log_n;
        long long count;
    };

    // ---- register butterfly stages over 16 residues ------------------------------------------
    // Cooley-Tukey stage pairing k and k+GAP; tw[j] is the twiddle of the j-th block of 2*GAP.
    template <int GAP>
    __device__ __forceinline__ void ct_stage(u64 (&x)[16], const Twiddle (&tw)[8], u64 q, u64 two_q)
    {
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            if (!(k & GAP))
            {
                const Twiddle w = tw[k / (2 * GAP)];
                u64 u = csub(x[k], two_q);
                u64 v = mul_shoup_lazy(x[k + GAP], w.w, w.wq, q);
                x[k] = u + v;
                x[k + GAP] = u + two_q - v;
            }
        }
    }

    // Gentleman-Sande stage pairing k and k+GAP.
    template <int GAP>
    __device__ __forceinline__ void gs_stage(u64 (&x)[16], const Twiddle (&tw)[8], u64 q, u64 two_q)
    {
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            if (!(k & GAP))
            {
                const Twiddle w = tw[k / (2 * GAP)];
                u64 u = x[k], v = x[k + GAP];
                x[k] = csub(u + v, two_q);
                x[k + GAP] = mul_shoup_lazy(u + two_q - v, w.w, w.wq, q);
            }
        }
    }

    // Last inverse stage (GAP = 8 of the final phase): output scaled by N^-1.
    __device__ __forceinline__ void gs_stage_last(u64 (&x)[16], const LimbConst &lc)
    {
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            u64 u = x[k], v = x[k + 8];
            x[k] = mul_shoup_lazy(csub(u + v, lc.two_q), lc.inv_n, lc.inv_n_quo, lc.q);
            x[k + 8] = mul_shoup_lazy(u + lc.two_q - v, lc.inv_n_w, lc.inv_n_w_quo, lc.q);
        }
    }

    // Transforms `count` consecutive polynomials in place; polynomial p lives at data + p*n and
    // uses the prime with index d_limb_ids[(p / div) % period].
    void ntt_forward(Context *c, u64 *data, long long count, const int *d_limb_ids, int period, int div = 1);
    void ntt_inverse(Context *c, u64 *data, long long count, const int *d_limb_ids, int period, int div = 1);
} // namespace moai
