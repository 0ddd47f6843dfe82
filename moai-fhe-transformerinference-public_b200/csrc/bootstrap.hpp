// Full-slot CKKS bootstrapping for the batched Evaluator.
//
// Same pipeline and level budget as the reference's Bootstrapper::bootstrap_full_3
// (M/source/bootstrapping/Bootstrapper.cpp:3231-3251): ModRaise (1 -> L limbs) -> CoeffToSlot
// (3 levels) -> EvalMod on the real and imaginary halves (cosine polynomial of degree 59 in 6
// levels + 2 double-angle steps) -> SlotToCoeff (3 levels); 14 levels in total, output scale =
// final_scale.  Results depend on FP64 polynomial coefficients, so parity with the reference is
// by tolerance, not bit-exact (BASELINE.json north_star).
//
// The implementation is our own (not a transcription of Bootstrapper.cpp):
//  * CoeffToSlot / SlotToCoeff matrices are built from the radix-2 factors of the CKKS "special
//    FFT" (slot j <-> root zeta^{5^j}), log2(slots) butterfly levels merged into 3 sparse-diagonal
//    matrices; the bit-reversal permutations of the two directions cancel and are never applied.
//  * Every diagonal is pre-rotated for baby-step/giant-step evaluation and ENCODED ONCE per
//    (stage, level) on the device at set-up (the reference re-encodes ~318 vectors per call,
//    Bootstrapper.cpp:2026-2027); plaintext scales are chosen so that every rescale is exact.
//  * EvalMod: least-squares Chebyshev fit of cos(2 pi (x - 1/4) / 2^r) on the union of the
//    intervals [i - 2^-w, i + 2^-w], |i| < K (the domain the reference's multi-interval Remez
//    uses, M/source/bootstrapping/RemezCos.h:11-16), evaluated by a depth-optimal recursive
//    Chebyshev division (59 = 32 + 16 + 8 + 3 -> 6 levels) with exact target scales.
//  * Everything is batched over independent ciphertexts.
#pragma once
#include "evaluator.hpp"

namespace moai
{
    typedef std::complex<double> cd;

    struct BootParams
    {
        int boundary_K = 25;  // |t / q0| < K            (test_full_scheme.hpp:346)
        int deg = 59;         // cosine polynomial degree (:347)
        int double_angles = 2; // scale_factor r          (:348)
        int log_width = 10;   // loge                     (:352)
        int total_limbs = 35; // data limbs after ModRaise (:368)
        double final_scale = 70368744177664.0; // 2^46
        // 1: plan the BSGS stages for hoisted baby steps (more, cheaper baby steps);
        // 2: additionally keep the baby-step rotations in the key-switch basis (lazy mod-down: one mod-down per giant
        //    step, csrc/ops.cu k_bsgs_ext) and run the first CoeffToSlot stage — whose input is the mod-raised
        //    ciphertext, a single small digit — with single-digit keys and baby steps only
        int hoisting = 0;
        // levels the recursive Chebyshev division of the cosine spends: ceil(log2(deg + 1)) (59 -> 6, 31 -> 5).
        // The whole EvalMod spends poly_levels() + double_angles, which must be 8 (output at total_limbs - 14):
        // the reference's (59, 2); (31, 3) also fits the budget but not the precision (fit error 3e-4).
        int poly_levels() const
        {
            int l = 0;
            while ((1 << l) < deg + 1)
            {
                l++;
            }
            return l;
        }
    };

    // one sparse-diagonal matrix, prepared for BSGS evaluation at a fixed level
    struct LinearStage
    {
        int limbs = 0;                    // level (limb count) at which the stage is applied
        std::map<int, std::vector<cd>> diags; // signed offset -> diagonal (length n slots)
        int stride = 1, giant = 1;        // offsets are stride * (giant * i + j)
        std::vector<int> baby;            // j values present (0 first)
        std::vector<int> giants;          // i values present
        // encoded, pre-rotated diagonals on the device: key (i, j)
        std::map<std::pair<int, int>, Pt> pts;
        double pt_scale = 0;              // scale the plaintexts are encoded with
        double diag_factor = 1.0;         // constant folded into every diagonal at encode time
        // the same plaintexts over the key-switch basis of the lazy path (data limbs + extra primes), encoded on first
        // use for the digit layout the registered keys select (ext_layout: extra primes, or KS_SINGLE)
        std::map<std::pair<int, int>, Pt> pts_ext;
        int ext_layout = -99;
        bool first = false;               // applied to the mod-raised ciphertext (first CoeffToSlot stage)
    };

    class Bootstrapper
    {
    public:
        Bootstrapper(Context *ctx, const BootParams &p);
        // re-plan the linear stages for hoisted (pre-permuted-key) rotations; call before required_steps()
        void set_hoisting(int mode);
        // steps of the first CoeffToSlot stage's baby steps when they run on single-digit keys (hoisting mode 2)
        std::vector<int> single_digit_steps() const;
        std::vector<int> required_steps() const; // rotation steps (normalised to [0, slots)) the BSGS plans use
        // (step, limbs) for every rotation key and level it is used at; step 0 = the complex conjugation
        std::vector<std::pair<int, int>> required_step_levels() const;
        // in: batch of size-2 ciphertexts at 1 limb (chain_index 0); returns them at
        // total_limbs - 14 limbs with scale final_scale
        // stop_after (diagnostics, moai_bootstrap_phase_debug): 0 = the whole bootstrapping; 1 = return after ModRaise
        // (scale q0); 2 = after CoeffToSlot (2 x batch ciphertexts: real halves then imaginary halves); 3 = after
        // EvalMod (same order)
        Ct bootstrap(const Evaluator &ev, const Ct &in, const Keys &keys, int stop_after = 0);

        // Bootstrapping of REAL-slot messages, two per bootstrapping.  A full-slot bootstrapping refreshes N
        // independent real coefficients; a real-slot message only uses N/2 of them (its polynomial is fixed by
        // the conjugation), so z = a + i b carries two messages through ONE bootstrapping:
        //   z <- a + X^(N/2) b  (exact monomial product at one limb),  declared at twice the input scale
        //   r  = bootstrap(z) ~ (a + i b) / 2,   c = conj(r)   (one key switch at total_limbs - 14 limbs)
        //   a' = r + c,  b' = -i (r - c)
        // MOAI's activations are real, so the 4 x 768 bootstrappings of an encoder layer
        // (M/test/test_full_scheme.hpp:654-660, 758-764, 991-995, 1081-1085) become 4 x 384.
        // in: [B][2][1][N]; pairs (j, j + ceil(B/2)); an unpaired ciphertext travels alone.  `into`: optional
        // storage for the [B][2][total_limbs - 14][N] result; chunk_pairs bounds the workspace.
        // [pair_first, pair_last): the pairs this call processes (default: all ceil(B/2)); the others' slots of the
        // result are left untouched — a rank of a multi-GPU run computes its share and gathers the rest (comm.hpp)
        Ct bootstrap_real_pairs(const Evaluator &ev, const Ct &in, const Keys &keys, long long chunk_pairs,
                                const Ct *into = nullptr, long long pair_first = 0, long long pair_last = -1);

        // host-side artefacts, exposed for the CPU test-suite
        const std::vector<double> &cheb_coeffs() const
        {
            return cheb_;
        }
        const LinearStage &stage(int dir, int idx) const
        {
            return dir == 0 ? cts_[idx] : stc_[idx];
        }
        int slots() const
        {
            return (int)(c_->n / 2);
        }

        BootParams prm;
        // softmax_boot's masked plaintext batches, encoded once per (mask, shift constant, level, scale) (csrc/modules.cu)
        std::map<std::string, Pt> mask_pts;

    private:
        Context *c_;
        LinearStage cts_[3], stc_[3];
        std::vector<double> cheb_;
        double stc_encoded_for_scale_ = 0; // input scale the SlotToCoeff constants were folded for

        void build_matrices();
        void plan_bsgs(LinearStage &st) const;
        void prepare_stage(LinearStage &st, double pt_scale, double diag_factor);
        // ids == nullptr: limbs 0 .. st.limbs - 1 into st.pts; else the listed primes into st.pts_ext
        void encode_stage(const Evaluator &ev, LinearStage &st, const std::vector<int> *ids);
        Ct linear_transform_lazy(const Evaluator &ev, const Ct &ct, LinearStage &st, const Keys &keys) const;
        Ct finish_giants(const Evaluator &ev, const Ct &ct, LinearStage &st, const Keys &keys, const std::vector<int> &gi,
                         std::vector<Ct> &inner) const;
        void fit_cosine();
        Ct linear_transform(const Evaluator &ev, const Ct &ct, LinearStage &st, const Keys &keys);
        Ct eval_mod(const Evaluator &ev, const Ct &y, const Keys &keys) const;
        Ct eval_cheb(const Evaluator &ev, const std::vector<double> &coef, int target_limbs, double target_scale,
                     const std::map<int, Ct> &T, const Keys &keys) const;
        // p = rescale(relinearize(acc3)) + rest: the products q * T_g of one remainder chain share a level and a
        // scale, so they are summed as size-3 ciphertexts and relinearized ONCE (lazy relinearization)
        void eval_cheb_parts(const Evaluator &ev, const std::vector<double> &coef, int target_limbs, double target_scale,
                             const std::map<int, Ct> &T, const Keys &keys, Ct &acc3, Ct &rest) const;
    };
} // namespace moai
