// Fast-mode key switching with GROUPED digits ("level-aware hybrid key switching").
//
// SEAL's key switch (S/evaluator.cpp:2724-3021) has one digit per RNS prime and one special prime p: at l limbs it
// costs l (l + 1) forward NTTs, which is 60 % of the whole encoder layer.  The keys it uses, however, contain more
// than that: K_J = (b_J, a_J) with  b_J + a_J s = e_J + F_J s'  (mod Q_L p),  F_J = p mod q_J, 0 mod every other
// prime (S/keygenerator.cpp:316-371).  Two facts follow, and neither needs the secret key:
//   * keys add:  K_G = sum_{J in G} K_J  encrypts  F_G = sum F_J,  the CRT element that is p modulo every prime of
//     the group G and 0 modulo every other prime;
//   * a ciphertext at l < L limbs does not use the primes q_l .. q_{L-1}, but the key still holds their limbs.  With
//     E = the top k data primes and P' = p * prod(E),  F_G = P' * t_G  where t_G = prod(E)^-1 mod q_J (J in G), 0 mod
//     the other primes below l:  K_G restricted to the basis {q_0..q_{l-1}} + E + {p} is a hybrid key-switching key
//     (Han-Ki) for the special modulus P' and the digit  D_G = CRT lift over G of [prod(E) c]_{q_J}.
// So:   sum_G D_G (.) K_G = P' c s' + sum_G D_G e_G   (mod Q_l P'),   and dividing by P' leaves c s' plus a noise of
// sum_G D_G e_G / P' — small as long as Q_G <= P' (the groups are sized that way).  A key switch at l limbs then
// needs  digits * (l + k + 1)  forward NTTs with digits = ceil(l / ~(k + 1)) instead of l (l + 1): 288 instead of 1056
// at l = 32 (k = 3), 84 instead of 462 at l = 21 (k = 6), plus two fast base conversions (digit extension, mod-down by
// the k + 1 primes of P'), both fused into the first pass of the NTT that follows them (ConvTab, ntt.cuh).
//
// Same plaintext as SEAL's key switch, different (still negligible: <= 2^11 against scales >= 2^46) noise, hence
// fast mode only; the SEAL-exact path is untouched.  The grouped keys are derived on the device from the keys a
// stock SEAL client ships (ksg_key_prepare); a group partition depends on k only, so one grouped key serves every
// level l <= L - k (groups that reach beyond l are used partially).
#pragma once
#include "ntt.cuh"
#include "ops.cuh"
#include <vector>

namespace moai
{
    // highest level a key with k extra primes can serve: the chain's data primes minus the k borrowed ones
    int ksg_max_limbs(Context *c, int k);
    // digit groups of such a key truncated to lmax data limbs, and its layout [digits][2][lmax + k + 1][n]
    int ksg_digits(Context *c, int k, int lmax);
    inline int ksg_key_kl(int k, int lmax)
    {
        return lmax + k + 1;
    }
    size_t ksg_key_words(Context *c, int k, int lmax);
    // modelled cost (microseconds per ciphertext) of one key switch at `limbs` with k extra primes (k = 0: SEAL's
    // digits) and the k that minimises it
    double ksg_cost(Context *c, int limbs, int k);
    int ksg_best_k(Context *c, int limbs);

    // K_G = sum_{J in G} sigma_elt^-1(K_J) over the basis {q_0..q_{lmax-1}} + E + {p}.
    // in: SEAL layout [kl-1][2][kl][n]; out: [digits][2][lmax + k + 1][n].  pre_permute = false (relinearisation keys)
    // skips the automorphism.
    void ksg_key_prepare(Context *c, const u64 *in, uint32_t elt, int k, int lmax, bool pre_permute, u64 *out);

    size_t ksg_ext_bytes_per_ct(Context *c, int limbs, int k);
    // ext[b][I][g][n] = NTT_I(D_g mod m_I) (passes & 2) or its pass-A half (passes == 1)
    void ksg_decompose(Context *c, const u64 *target, long long batch, int limbs, int k, u64 *ext,
                       long long target_stride, int passes);
    // out[P][limbs][n] = round(acc[P][limbs + k + 1][n] / P')  (+ addend; addend_even_only: even polynomials only)
    // addend_group: polynomials per ciphertext in the addend's layout (2; 3 when relinearize adds (c0, c1) of its input)
    // in_stride: input polynomial P sits at acc + P * in_stride * (limbs + k + 1) * n (2 with acc + rns * n: the c1's only)
    void ksg_moddown(Context *c, const u64 *acc, long long polys, int limbs, int k, const u64 *addend,
                     bool addend_even_only, u64 *out, int addend_group = 2, int in_stride = 1);
    // key switch WITHOUT the mod-down: acc[batch][2][limbs + k + 1][n] = sum_G D_G (.) K_G; workspaces
    // ext (ksg_ext_bytes_per_ct per item) and direct ([batch][limbs][n])
    void ksg_switch_acc(Context *c, const u64 *target, long long batch, int limbs, int k, const u64 *ksk, int key_kl,
                        u64 *ext, u64 *direct, u64 *acc, long long target_stride);
    // total = sum_g sigma_g(acc_g + (extra_g.c0, 0)) in the key-switch basis (giant steps of the lazy BSGS)
    void ksg_giants_sum(Context *c, int n_giants, const u64 *const *acc, const u64 *const *extra,
                        const uint32_t *const *perm, u64 *total, long long batch, int limbs, int k);
    // complete key switch: out[b] = addend[b] + keyswitch(target[b]); out / addend are size-2 ciphertexts
    void ksg_switch(Context *c, const u64 *target, long long batch, int limbs, int k, const u64 *ksk, int key_kl,
                    const u64 *addend, u64 *out, long long target_stride, bool addend_c0_only, int addend_group = 2,
                    bool rescale = false);
    // out[P][limbs - 1][n] = round((acc + P' addend) / (P' q_{limbs-1})): mod-down and the rescale after it in one
    // division (fast mode; acc's limb limbs - 1 is modified in place)
    void ksg_moddown_rescale(Context *c, u64 *acc, long long polys, int limbs, int k, const u64 *addend,
                             bool addend_even_only, u64 *out, int addend_group = 2);
    // hoisted rotations from one decomposition (up to KSM_R keys per pass)
    void ksg_rotate_hoisted_multi(Context *c, const u64 *ct, const u64 *ext, long long batch, int limbs, int k, int n_rot,
                                  const uint32_t *elts, const u64 *const *ksk_pre, const int *key_kl, u64 *const *outs);
    // ---- pieces for the lazy (extended-basis) BSGS of the bootstrapping's linear stages -----------------------------
    constexpr int KS_SINGLE = -1; // digit layout "one digit": the polynomial itself is small (right after ModRaise)
    struct KsExtInfo
    {
        KsShape shape;            // digits / targets of the key-switch basis at this level
        int k = 0;                // extra primes (0 = SEAL's digits, KS_SINGLE)
        std::vector<int> h_ids;   // prime index of every target
        std::vector<u64> h_pmod;  // P' mod q_j for the data limbs
    };
    KsExtInfo ks_ext_info(Context *c, int k, int limbs);       // k >= 0
    KsExtInfo ks_ext_info_single(Context *c, int limbs);
    void ks_int_targets(Context *c, const KsShape &sh, const std::vector<int> &h_ids, u64 *ext, long long batch,
                        const u64 *ksk, int key_kl, u64 *acc, bool need_pass_b);
    void ks_moddown(Context *c, const u64 *acc, long long polys, int limbs, int k, const u64 *addend, bool addend_even_only,
                    u64 *out); // k == 0: by the special prime alone
    // Single-digit key switch of a MOD-RAISED ciphertext.  After ModRaise (Bootstrapper.cpp:2938-2992) c1 is the
    // centred lift of coefficients modulo q_0: as an integer polynomial it is below q_0 / 2 < p, i.e. it IS one
    // digit, and K_all = sum over ALL J of K_J encrypts p * s' (F = p modulo every prime).  The "decomposition" is the
    // ciphertext's own limbs plus one NTT for the special prime: ext [batch][limbs + 1][1][n]; key [1][2][kl][n].
    size_t ks_single_ext_bytes_per_ct(Context *c, int limbs);
    void ks_hoist_modraised(Context *c, const u64 *c1, long long batch, int limbs, u64 *ext, long long c1_stride);
    void ks_key_prepare_single(Context *c, const u64 *in, uint32_t elt, bool pre_permute, u64 *out);
    void ksg_release(Context *c);
} // namespace moai
