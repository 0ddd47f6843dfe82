// Negacyclic NTT / INTT over RNS limbs for sm_100a — two-pass, register-blocked radix-16.
//
// Computes exactly the function of the reference's Harvey transforms
// (forward: DWTHandler::transform_to_rev, S/util/dwthandler.h:94-191, natural -> bit-reversed;
//  inverse: transform_from_rev, S/util/dwthandler.h:202-356, bit-reversed -> natural with N^-1
//  folded into the last stage), with the tables of NTTTables::initialize (S/util/ntt.cpp:241-300).
// All public results are canonical residues, so the internal arithmetic is free.
//
// B200 mapping (N = 2^logN, viewed as R x 256 with R = N/256; i = a*256 + b):
//   pass A : the logN-8 stages that pair rows (gap >= 256).  One CTA owns all R rows of 16
//            adjacent columns (R x 128 B tile, every global access a full 128-byte line); each
//            thread keeps 16 residues in registers and runs 4 stages without touching memory,
//            one shared-memory transpose re-blocks the tile for the remaining stages.
//   pass B : the 8 stages inside a 256-element row.  One CTA owns 16 rows; 4 + 4 register
//            stages with one padded (conflict-free) shared-memory transpose.
// A limb (512 KiB at N = 65536) exceeds one SM's shared memory, hence two passes: 2 MiB of
// traffic per limb-transform against the 1 MiB algorithmic floor.
//
// Arithmetic.  B200 has no 64-bit integer multiplier and IMAD.WIDE (32x32->64) issues at a
// quarter of the FP64 FMA rate (tools/microbench.cu, profiles/), so the integer Shoup butterfly is
// multiplier-bound at ~0.9 us per limb-transform.  Two exact paths are therefore compiled and
// selected per limb (CTA-uniform):
//   * FP64 path for primes below 2^51 (all 35 data primes of the repo's chain): residues are
//     integer-valued doubles in a symmetric lazy range; x*w mod p = (h + l) - rint(h/p) p with
//     h = fl(x w), l = fma(x, w, -h) — every step is exact because all intermediate integers stay
//     below 2^53 (bounds in csrc/ntt.cu).  No precomputed Shoup quotient: 8-byte twiddles.
//   * integer Harvey/Shoup path (lazy [0, 4q)) for wider primes (the 58-bit special prime).
#pragma once
#include "context.hpp"

namespace moai
{
    // Fast base conversion fused into the forward NTT's first pass (NttPrologue mode 3).  Output polynomial
    // (b, I, g) of an [batch][rns][digits][n] buffer is the conversion of source limbs s0[g] .. s0[g] + cnt[g] - 1
    // of batch item b (coefficient form, already multiplied by (Q_g / q_j)^-1 mod q_j) into target modulus m_I:
    //     sum_j y_j * [(Q_g / q_j) mod m_I]  -  v * [Q_g mod m_I],   v = rint(sum_j y_j / q_j)
    // i.e. the centred CRT lift of the digit, |D| <= Q_g / 2 (the float v is the same for every target modulus:
    // identical operations on identical inputs).  Used for the grouped digits of the fast-mode key switch and
    // for its mod-down by several primes (csrc/ksgroup.cu).
    constexpr int CONV_MAX = 16; // source limbs per digit
    struct ConvTab
    {
        const int *s0 = nullptr;             // [digits]
        const int *cnt = nullptr;            // [digits]
        const double *invq = nullptr;        // [src_limbs] 1 / q_src
        const unsigned char *wide = nullptr; // [src_limbs] 1: residues may reach 2^52 (split 2^26 hi + lo on the FP64 path)
        const u64 *B = nullptr;              // [digits][rns][CONV_MAX]  (Q_g / q_j) mod m_I
        const u64 *B26 = nullptr;            // [digits][rns][CONV_MAX]  2^26 (Q_g / q_j) mod m_I   (wide sources)
        const double *Bd = nullptr;          // the same two tables as centred doubles (FP64-path targets only)
        const double *B26d = nullptr;
        const u64 *negQ = nullptr;           // [digits][rns]  (-Q_g) mod m_I
        const double *negQd = nullptr;       // centred double
        // [rns] or nullptr: digit whose group CONTAINS target I (its residue modulo q_I is the input limb itself,
        // times a constant): the fused key-switch kernel takes that product directly and the polynomial is skipped
        const int *own = nullptr;
        int src_limbs = 0;
        // Tensor-core form of the conversion (FP64-path targets, N >= 2^13; nullptr: FP64 products as above).
        // The sum over sources is a u8 x u8 GEMM with a constant right-hand side: K = the 8 bytes of every source
        // residue, N = the 8 bytes of  C(j, a, I) = 2^(8a) (Q_g / q_j) mod m_I,  M = coefficients;
        //     sum_j y_j B_jI  ==  sum_b 2^(8b) sum_(j,a) y_(j,a) C(j, a, I)_b    (mod m_I)
        // so the FP64 pipe only recombines four 32-bit partial sums per coefficient (one exact product) instead of
        // doing 9 operations per source.  v = rint(sum_j y_j / q_j) is computed ONCE per coefficient by conv_quotient()
        // into byte 7 of the group's first source (free: residues are below 2^56) and rides in the GEMM with the
        // constant (-Q_g) mod m_I.  BT: mma.sync m16n8k32 B fragments, [digits][rns][CONV_KSTEPS][32 lanes][2].
        const uint32_t *BT = nullptr;
        const double *c32d = nullptr; // [rns] centred 2^32 mod m_I
        // [src_limbs] or nullptr: 1 = this source limb arrives as CENTRED DOUBLES (the inverse transform wrote them,
        // NttArgs::fp_out): the prologue multiplies them as they are — no conversion per (source, target) — and the
        // quotient v comes precomputed from conv_quotient_fp (NttPrologue::conv_v) instead of one FMA per (source, target)
        const unsigned char *fpsrc = nullptr;
        const u64 *srcq = nullptr; // [src_limbs] the source primes (integer-path targets canonicalise the doubles)
    };
    constexpr int CONV_KSTEPS = CONV_MAX / 4; // four sources (32 bytes) per mma k-step
    // writes v into byte 7 of every group's first source limb of src [batch][src_limbs][n] (once, before the transform)
    void conv_quotient(Context *c, u64 *src, long long batch, const ConvTab &tab, int digits);
    // v[batch][digits][n] (doubles) = rint(sum_j y_j / q_j) from sources that are centred doubles (fpsrc) or residues
    void conv_quotient_fp(Context *c, const u64 *src, long long batch, const ConvTab &tab, int digits, double *v);

    // Optional constant folded into the inverse transform: the last stage multiplies by N^-1 (and by the last root);
    // with scale[slot] it multiplies by c * N^-1 instead, so "INTT then scale by a per-limb constant" is one pass
    // (slot = (p / div) % period, as for the limb ids).
    struct NttScale
    {
        u64 inv_n, inv_n_quo, inv_n_w, inv_n_w_quo; // c N^-1 and c w N^-1 mod q with Shoup quotients
        double inv_n_d, inv_n_w_d;                  // the same two as centred doubles
    };
    NttScale ntt_scale_make(Context *c, int prime, u64 constant);

    // Optional epilogue of the forward transform's grouped pass B: the tail of a divide-and-round
    // (S/util/rns.cpp:881-901, S/evaluator.cpp:2990-3018).  Polynomial p = (P, j) (period = targets, div = 1) is not
    // stored; instead  out[P][j] = (in[P][j] - NTT(u)[P][j]) * inv[j] (+ addend)  is written, so the transformed
    // correction never travels through HBM.
    struct FinishEpi
    {
        const u64 *in = nullptr;     // [P][limbs_in][n]
        const u64 *addend = nullptr; // polynomial P <-> (P / 2) * addend_group + P % 2 of [..][targets][n]
        u64 *out = nullptr;          // [P][targets][n]
        const Twiddle *inv = nullptr; // [targets]
        int limbs_in = 0, addend_even_only = 0, addend_group = 2;
        // merged mod-down + rescale (csrc/ksgroup.cu): the addend is multiplied by addend_mul[j] (q_last^-1 mod q_j)
        // and its polynomials hold addend_limbs limbs (0: as many as there are targets)
        const Twiddle *addend_mul = nullptr;
        int addend_limbs = 0;
    };

    struct NttArgs
    {
        u64 *data;             // [count][n]
        const Twiddle *tw;     // integer path: forward or inverse Shoup table, [kl][n]
        const double *tw_fp;   // FP64 path: the same roots as symmetric doubles, [kl][n]
        const LimbConst *limb; // [kl]
        const int *limb_ids;   // [period]; poly p uses prime limb_ids[(p / div) % period]
        int period;
        int div;
        int log_n;
        long long count;
        // Optional fused prologue of the forward transform (pass A reads `src` instead of `data`):
        //   src_mode 1 (key-switch digit extension, S/evaluator.cpp:2831-2856): polynomial p reads the
        //              coefficient-form digit src[(p / (period*div)) * div + p % div] and reduces it
        //              modulo its own prime;
        //   src_mode 2 (divide-and-round expansion, S/util/rns.cpp:851-880 / S/evaluator.cpp:2966-2990):
        //              polynomial p reads t = src[p / period] (coefficients modulo prime `last_id`) and
        //              forms ((t + half) mod q_last) mod q_i + (q_i - half mod q_i).
        //   src_mode 3 (fast base conversion, ConvTab above): polynomial p = (b, I, g) with
        //              b = p / (period*div), I = (p / div) % period, g = p % div.
        // pass B only: logical polynomial p lives at data + ((p / grp_size) * grp_stride + p % grp_size) * n
        // (grp_size = 0: contiguous)
        long long grp_size = 0, grp_stride = 0;
        const u64 *src = nullptr;
        int src_mode = 0;
        int last_id = 0;
        int kl = 0;
        const u64 *half_mod = nullptr; // [kl][kl]
        ConvTab conv;
        long long skipped = 0;
        const NttScale *scale = nullptr; // inverse transform only
        long long p_base = 0;            // inverse transform only: first polynomial of this launch (L2-sized chunks)
        int fp_out = 0;                  // inverse transform only: FP64-path limbs are stored as centred doubles
        const double *conv_v = nullptr;  // src_mode 3: precomputed quotients [batch][div][n] (ConvTab::fpsrc)
        FinishEpi fin;                   // forward transform, grouped pass B only
    };

    struct NttPrologue
    {
        const u64 *src = nullptr;
        int mode = 0;
        int last_id = 0;
        const ConvTab *conv = nullptr; // mode 3
        long long skipped = 0;         // polynomials the pass-A kernel skips (ConvTab::own): not counted as work units
        const double *conv_v = nullptr; // mode 3 with ConvTab::fpsrc
    };

    // Shape of a key-switch inner product.  SEAL's per-prime digits (S/evaluator.cpp:2805-2909): digits = limbs,
    // rns = limbs + 1, n_data = limbs.  Grouped digits (csrc/ksgroup.cu): digits = groups in use, rns = limbs + k + 1.
    // Key layout [digit][2][key_kl][n]; target I reads key limb I (I < n_data) or I + key_kl - rns (the trailing
    // extra / special primes sit at the end of a key polynomial).
    struct KsShape
    {
        int digits = 0;
        int rns = 0;
        int n_data = 0;
        const int *ids = nullptr; // device [rns]: prime index of target modulus I
    };
    KsShape ks_shape_seal(Context *c, int limbs);


    // Transforms `count` consecutive polynomials in place; polynomial p lives at data + p*n and
    // uses the prime with index d_limb_ids[(p / div) % period].
    // passes: bit 0 = pass A (row-pairing stages, + prologue), bit 1 = pass B (in-row stages)
    // fin: fuse the divide-and-round tail into pass B when the launch allows it (grouped pass B, FP64-path primes);
    // returns true when it was applied (the caller then skips its finish kernel)
    bool ntt_forward(Context *c, u64 *data, long long count, const int *d_limb_ids, int period, int div = 1,
                     const NttPrologue *pro = nullptr, int passes = 3, const FinishEpi *fin = nullptr);
    // pass B of `groups` runs of `grp_size` consecutive polynomials, run g starting at
    // data + g * grp_stride * n, all modulo the single prime *d_limb_id
    void ntt_forward_pass_b_strided(Context *c, u64 *data, long long groups, long long grp_size, long long grp_stride,
                                    const int *d_limb_id);
    // Hoisted rotations: inner products of one set of extended digits (NTT form) with up to KSM_R keys
    // in one pass on the FP64 pipe; acc[r][batch][2][limbs+1][n] receives canonical residues for every
    // FP64-path modulus (integer-path moduli are left untouched).
    constexpr int KSM_R = 4;
    void ks_mac_multi(Context *c, const u64 *ext, long long batch, const KsShape &sh, int n_keys, const u64 *const *ksk,
                      const int *key_kl, u64 *const *acc);
    // Fused key-switch kernel (csrc/ntt.cu): pass B of the digit-extension NTT + inner product with
    // the evk for every FP64-path modulus.  mid = pass-A output [batch][rns][digits][n];
    // acc[batch][2][rns][n] receives canonical residues for those moduli (integer-path moduli
    // are left untouched for the un-fused kernels).
    // direct / own (optional, grouped digits): own[I] = digit whose NTT-form residue modulo target I is
    // direct[batch][n_data][n] (no transform needed; mid holds nothing for it)
    void ks_passb_mac(Context *c, const u64 *mid, long long batch, const KsShape &sh, const u64 *ksk, int key_kl,
                      u64 *acc, const u64 *direct = nullptr, const int *own = nullptr);
    void ntt_inverse(Context *c, u64 *data, long long count, const int *d_limb_ids, int period, int div = 1);
    // out of place: polynomial p is read from src + ((p / grp_size) * grp_stride + p % grp_size) * n (runs of grp_size
    // polynomials, grp_stride polynomials apart: one limb range of every ciphertext of a batch) and the coefficients
    // land contiguously in data — saves the gather copy in front of an in-place transform
    // fp_out: FP64-path limbs are written as centred doubles (bit patterns) for the base-conversion prologue
    void ntt_inverse_from(Context *c, const u64 *src, long long grp_size, long long grp_stride, u64 *data, long long count,
                          const int *d_limb_ids, int period, int div = 1, const NttScale *d_scale = nullptr,
                          bool fp_out = false);
} // namespace moai
