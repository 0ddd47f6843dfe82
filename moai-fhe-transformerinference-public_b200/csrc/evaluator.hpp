// Host-side C++ evaluator over batches of device ciphertexts.
//
// Mirrors the call surface and the metadata rules of the reference's seal::Evaluator /
// seal::CKKSEncoder (S/evaluator.h:93-1386, S/ckks.h:148-432) — same method names, same scale
// arithmetic in doubles, same argument checks and exception kinds — but every object is a BATCH
// of independent ciphertexts resident in HBM, because that is the only parallelism the
// reference has (`#pragma omp parallel for` over ciphertexts).  The MOAI module functions
// (csrc/modules.cu) are written against this class the way the reference's are written against
// seal::Evaluator.
#pragma once
#include "ntt.cuh"
#include "ops.cuh"
#include "ksgroup.hpp"
#include <cmath>
#include <complex>
#include <map>
#include <memory>
#include <vector>

namespace moai
{
    struct DevBuf
    {
        void *p = nullptr;
        cudaStream_t s;
        DevBuf(size_t bytes, cudaStream_t stream) : s(stream)
        {
            p = device_alloc(bytes, stream);
        }
        ~DevBuf()
        {
            device_free(p, s);
        }
        DevBuf(const DevBuf &) = delete;
        DevBuf &operator=(const DevBuf &) = delete;
    };

    // `batch` ciphertexts of identical shape/level/scale: [batch][size][limbs][n].
    // Copies are shallow handles (shared buffer); clone() deep-copies like seal::Ciphertext's
    // copy-assignment does.
    struct Ct
    {
        std::shared_ptr<DevBuf> buf;
        u64 *d = nullptr;
        long long batch = 0;
        int size = 2;
        int limbs = 0;
        double scale = 1.0;
        bool empty() const
        {
            return d == nullptr;
        }
    };

    // Plaintext: either a polynomial per batch item / broadcast ([count][limbs][n], count = 1 or
    // batch) or a scalar encoding = one constant per limb (S/ckks.cpp:131-153).
    struct Pt
    {
        std::shared_ptr<DevBuf> buf;
        u64 *d = nullptr;
        long long count = 0;
        int limbs = 0;
        double scale = 1.0;
        bool is_scalar = false;
        std::vector<u64> consts;
    };

    // One key-switching key on the device.  key_kl = limbs stored per key polynomial: the
    // context's kl for SEAL's layout [kl-1][2][kl][n] (S/kswitchkeys.h:335-340), L + 1 for a key
    // truncated to its first L digits / data limbs (+ the special prime) by key_prepare.
    // k_extra > 0: a grouped-digit key of ksg_key_prepare (csrc/ksgroup.hpp), [digits][2][max_limbs + k_extra + 1][n].
    struct KeyRef
    {
        const u64 *p = nullptr;
        int key_kl = 0;
        int k_extra = 0;
        int max_limbs() const
        {
            return key_kl - 1 - k_extra;
        }
    };

    // Device key material: SEAL's RelinKeys / GaloisKeys as raw device pointers.
    //  * galois      : keys exactly as SEAL generates them -> SEAL-bit-exact rotations.
    //  * galois_fast : pre-permuted keys K' = sigma^-1(K) (key_prepare), possibly several level
    //                  truncations per element -> hoisted rotations (same plaintext, different noise).
    //  * relin_fast  : grouped-digit variants of the relinearisation key (fast mode).
    struct Keys
    {
        KeyRef relin;
        std::vector<KeyRef> relin_fast;
        std::map<uint32_t, KeyRef> galois;
        std::map<uint32_t, std::vector<KeyRef>> galois_fast;
        // single-digit keys [1][2][kl][n] = sigma^-1(sum of all digits) for rotations of a MOD-RAISED ciphertext
        // (ksgroup.hpp): used by the first CoeffToSlot stage of the bootstrapping only
        std::map<uint32_t, KeyRef> galois_single;
        // cheapest (ksg_cost) pre-permuted key of `elt` that covers `limbs` levels (nullptr when none);
        // only_k >= 0 restricts the choice to keys with that many extra primes
        const KeyRef *fast(Context *c, uint32_t elt, int limbs, int only_k = -1) const
        {
            auto it = galois_fast.find(elt);
            return it == galois_fast.end() ? nullptr : pick(c, it->second, limbs, only_k);
        }
        // cheapest relinearisation key at this level: a grouped variant, else SEAL's
        const KeyRef *relin_at(Context *c, int limbs) const
        {
            const KeyRef *g = pick(c, relin_fast, limbs, -1);
            if (g && (!relin.p || ksg_cost(c, limbs, g->k_extra) < ksg_cost(c, limbs, 0)))
            {
                return g;
            }
            return relin.p ? &relin : nullptr;
        }
        static const KeyRef *pick(Context *c, const std::vector<KeyRef> &cands, int limbs, int only_k)
        {
            const KeyRef *best = nullptr;
            double best_cost = 0;
            for (auto &k : cands)
            {
                if (k.max_limbs() < limbs || (only_k >= 0 && k.k_extra != only_k) ||
                    (k.k_extra > 0 && limbs + k.k_extra > c->kl - 1))
                {
                    continue;
                }
                const double cost = ksg_cost(c, limbs, k.k_extra);
                if (!best || cost < best_cost || (cost == best_cost && k.key_kl < best->key_kl))
                {
                    best = &k;
                    best_cost = cost;
                }
            }
            return best;
        }
    };

    // Digit decomposition of c1 of a batch of ciphertexts, shared by every hoisted rotation of them.
    struct Hoisted
    {
        Ct src;                      // the unrotated ciphertexts (shared storage)
        std::shared_ptr<DevBuf> ext; // [batch][limbs + 1][limbs][n]; grouped digits: [batch][limbs + k + 1][digits][n]
        int k_extra = 0;
    };

    class Evaluator
    {
    public:
        Context *c;
        explicit Evaluator(Context *ctx) : c(ctx)
        {}

        size_t n() const
        {
            return c->n;
        }

        // ---- storage -------------------------------------------------------------------------
        Ct alloc(long long batch, int size, int limbs, double scale) const;
        Ct wrap(u64 *d, long long batch, int size, int limbs, double scale) const; // caller-owned memory
        Ct view(const Ct &a, long long b0, long long count) const;                 // sub-batch, shared storage
        // sum_j coef[j] * terms[j], each term mod-switched to `limbs` on the fly and its constant encoded at
        // out_scale / terms[j].scale, in one pass (fused multiply_const + mod_switch_to + add of a polynomial leaf)
        Ct lincomb_scalar(const std::vector<Ct> &terms, const std::vector<double> &coefs, int limbs, double out_scale) const;
        Ct clone(const Ct &a) const;
        void copy_into(const Ct &src, Ct &dst, long long dst_b0) const;            // dst[dst_b0 ...] = src
        Ct concat(const std::vector<Ct> &parts) const;
        Ct repeat(const Ct &a, long long times) const;                            // batch-1 -> batch-`times`

        // ---- S/evaluator.h surface ------------------------------------------------------------
        Ct add(const Ct &a, const Ct &b) const;          // b.batch == 1 broadcasts over a's batch
        Ct sub(const Ct &a, const Ct &b) const;
        void add_inplace(Ct &a, const Ct &b) const;
        void sub_inplace(Ct &a, const Ct &b) const;
        Ct negate(const Ct &a) const;
        Ct add_plain(const Ct &a, const Pt &p) const;
        void add_plain_inplace(Ct &a, const Pt &p) const;
        Ct sub_plain(const Ct &a, const Pt &p) const;
        Ct multiply_plain(const Ct &a, const Pt &p) const;
        Ct multiply(const Ct &a, const Ct &b) const;     // size 2 x size 2 -> size 3
        // multiply(mod_switch_to(a, l), mod_switch_to(b, l)) with l = min(a.limbs, b.limbs), without the copies
        Ct multiply_lowered(const Ct &a, const Ct &b) const;
        Ct square(const Ct &a) const;
        void multiply_accumulate(Ct &acc3, const Ct &a, const Ct &b) const; // acc3 += a x b (size 3)
        Ct relinearize(const Ct &a3, const Keys &k) const;
        Ct rescale_to_next(const Ct &a) const;
        // rescale_to_next(relinearize(a3)); one merged division with grouped-digit keys (MOAI_MERGE_RESCALE=0 disables)
        Ct relin_rescale(const Ct &a3, const Keys &k) const;
        Ct mod_switch_to(const Ct &a, int limbs) const;
        Ct mod_switch_to_next(const Ct &a) const
        {
            return mod_switch_to(a, a.limbs - 1);
        }
        Ct rotate_vector(const Ct &a, int steps, const Keys &k) const;
        Ct complex_conjugate(const Ct &a, const Keys &k) const;
        // fast mode: one decomposition, many rotations.  Needs pre-permuted keys.
        bool has_fast_key(int steps, int limbs, const Keys &k) const;
        Hoisted hoist(const Ct &a, int k_extra = 0) const;
        Ct rotate_hoisted(const Hoisted &h, int steps, const Keys &k) const;
        Ct rotate_fast(const Ct &a, uint32_t elt, const KeyRef &key) const; // chunked hoist + rotate
        // rotations of one batch by several steps: hoisted when every key is pre-permuted, else one by one
        std::vector<Ct> rotate_many(const Ct &a, const std::vector<int> &steps, const Keys &k) const;
        Ct sum_batch(const Ct &a) const; // one ciphertext = sum over the batch (modular)
        Ct inner_product(const Ct &a, const Ct &b) const;      // size-3 sum_j a[j] x b[j]
        Ct sum_sub_square(const Ct &a, const Ct &m) const;     // size-3 sum_j (a[j] - m)^2, m one ciphertext

        // ---- S/ckks.h surface -----------------------------------------------------------------
        Pt encode(double value, int limbs, double scale) const;
        Pt encode(const std::vector<std::complex<double>> &values, int limbs, double scale) const;
        Pt encode(const std::vector<double> &values, int limbs, double scale) const;
        // `count` plaintexts from one host array [count][n_vals] of complex values
        Pt encode_batch(const std::complex<double> *values, long long count, int n_vals, int limbs,
                        double scale) const;

        // ---- fork-only ops (S/evaluator.cpp:395-594, S/evaluator.h:1300-1386) ------------------
        Ct multiply_const(const Ct &a, double value) const;
        Ct add_const(const Ct &a, double value) const;
        Ct multiply_vector_reduced_error(const Ct &a, const std::vector<std::complex<double>> &v) const;
        Ct add_reduced_error(const Ct &a, const Ct &b) const;
        Ct sub_reduced_error(const Ct &a, const Ct &b) const;
        Ct multiply_reduced_error(const Ct &a, const Ct &b, const Keys &k) const;
        void double_inplace(Ct &a) const;
        // a = 2 a + value (double_inplace + add_const in one pass), and acc3 (size 3) += x2 (size 2)
        void double_add_const_inplace(Ct &a, double value) const;
        void add_into3(Ct &acc3, const Ct &x2) const;

        // ---- Bootstrapper::modraise_inplace -----------------------------------------------------
        Ct mod_raise(const Ct &a, int limbs_out) const;

        double last_prime(int limbs) const
        {
            return static_cast<double>(c->q[limbs - 1]);
        }

    private:
        void check_same(const Ct &a, const Ct &b, bool need_scale) const;
        Ct reduced_error_adjust(const Ct &hi, const Ct &lo) const;
    };

    // SEAL's util::are_close (S/util/common.h:569-573)
    inline bool are_close(double v1, double v2)
    {
        double sf = std::max(std::max(std::fabs(v1), std::fabs(v2)), 1.0);
        return std::fabs(v1 - v2) < std::numeric_limits<double>::epsilon() * sf;
    }
} // namespace moai
