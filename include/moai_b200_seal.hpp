// moai_b200_seal.hpp — header-only C++ facade over libmoai_b200.so's C ABI that supplies the `seal::`
// class surface MOAI's module code is written against, so that the reference's module headers
// (M/source/matrix_mul/*.hpp, M/source/non_linear_func/*.hpp, M/source/att_block/*.hpp) recompile
// UNCHANGED and run on the B200 backend.  (S/ = thirdparty/SEAL-4.1-bs/native/src/seal,
// M/ = include/ of petitioner/MOAI-FHE-TransformerInference-Public)
//
// How it is used: put `include/facade` of this repository BEFORE SEAL on the include path.
// `include/facade/seal/seal.h` includes this file, which defines the classes in
// `moai_b200::sealapi` and makes `seal` a namespace alias of it — `using namespace seal;` and
// `seal::Ciphertext` in the module code resolve here, while the mangled names stay distinct from the
// real library's (a client-side translation unit may keep using stock SEAL in the same process).
//
// What lives where:
//   * Every arithmetic operation is ONE call into the C ABI (include/moai_b200.h,
//     include/moai_b200_modules.h) on device memory; this header only keeps the metadata SEAL keeps
//     on the host (parms_id, scale, size, is_ntt_form) and applies SEAL's own checks and exception
//     types (std::invalid_argument / std::logic_error, S/evaluator.cpp:133-177, 1593-1596, 2594-2597).
//   * There is no CPU arithmetic path: without a CUDA device SEALContext's constructor throws.
//     The one host computation is CKKSEncoder::decode (CRT composition + FFT of a downloaded
//     plaintext) — in the reference that call only feeds debug prints (layernorm.hpp:279-310).
//   * A deployment generates its keys on the client with stock SEAL; KeyGenerator exists
//     here so that the reference's driver and test programs, which create keys in the function that evaluates,
//     compile unchanged (with the same PRNG seed it yields SEAL's keys bit for bit).
//     Keys and ciphertexts cross the boundary either as raw residues (Ciphertext::upload / download,
//     RelinKeys::upload, GaloisKeys::upload, PublicKey::upload, SecretKey::upload; SEAL's own layouts,
//     S/ciphertext.h:339-370, S/kswitchkeys.h:335-340) or in SEAL's wire format: Ciphertext::save / load,
//     PublicKey::load, RelinKeys::load, GaloisKeys::load read what the stock library's save() writes
//     (compr_mode_type::none), INCLUDING the seeded form a client normally ships — the uniform half of every
//     key component is a 64-byte seed, expanded here with SEAL's PRNG and rejection sampling
//     (moai_b200_seal_prng.hpp) — which halves the key upload.
//   * Encryptor (public-key encryption, S/encryptor.cpp:88-318, S/util/rlwe.cpp:224-309) IS provided, for
//     M/source/matrix_mul/Batch_encode_encrypt.hpp: u and the errors are sampled on the host from SEAL's PRNG
//     stream, the arithmetic runs on the device; with the same seed the ciphertext is SEAL's bit for bit.
//   * Ciphertext::data() / Plaintext::data() return DEVICE pointers.
//   * parms_id values are SEAL's own (BLAKE2b-256 of the level's parameters, S/encryptionparams.cpp:124-158),
//     so serialized objects carry ids a stock SEAL context recognises and vice versa.
//
// Threading: the reference calls one shared Evaluator from all OpenMP threads
// (M/test/test_full_scheme.hpp:654-660).  All entry points here are re-entrant; calls into one context
// are serialised onto that context's CUDA stream by a mutex (the GPU's parallelism is inside the
// kernels; for throughput use the batched module entry points of moai_b200_modules.h).
#ifndef MOAI_B200_SEAL_HPP
#define MOAI_B200_SEAL_HPP

#include "moai_b200.h"
#include "moai_b200_modules.h"
#include "moai_b200_seal_prng.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <istream>
#include <limits>
#include <map>
#include <memory>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace moai_b200
{
namespace sealapi
{
    // ------------------------------------------------------------------------------------------
    // small value types (S/encryptionparams.h, S/modulus.h, S/context.h)
    // ------------------------------------------------------------------------------------------
    using parms_id_type = std::array<std::uint64_t, 4>;
    static constexpr parms_id_type parms_id_zero = { { 0, 0, 0, 0 } };

    enum class scheme_type : std::uint8_t
    {
        none = 0x0,
        bfv = 0x1,
        ckks = 0x2,
        bgv = 0x3
    };

    enum class sec_level_type : int
    {
        none = 0,
        tc128 = 128,
        tc192 = 192,
        tc256 = 256
    };

    namespace util
    {
        // S/util/common.h:569-573
        inline bool are_close(double v1, double v2)
        {
            double sf = std::max(std::max(std::fabs(v1), std::fabs(v2)), 1.0);
            return std::fabs(v1 - v2) < std::numeric_limits<double>::epsilon() * sf;
        }

        inline std::uint64_t mulmod(std::uint64_t a, std::uint64_t b, std::uint64_t q)
        {
            return static_cast<std::uint64_t>((static_cast<unsigned __int128>(a) * b) % q);
        }

        inline std::uint64_t powmod(std::uint64_t a, std::uint64_t e, std::uint64_t q)
        {
            std::uint64_t r = 1 % q;
            a %= q;
            while (e)
            {
                if (e & 1)
                {
                    r = mulmod(r, a, q);
                }
                a = mulmod(a, a, q);
                e >>= 1;
            }
            return r;
        }

        // deterministic Miller-Rabin for 64-bit inputs (role of S/util/numth.cpp:295-361)
        inline bool is_prime(std::uint64_t n)
        {
            static const std::uint64_t bases[12] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37 };
            if (n < 2)
            {
                return false;
            }
            for (std::uint64_t p : bases)
            {
                if (n % p == 0)
                {
                    return n == p;
                }
            }
            std::uint64_t d = n - 1;
            int r = 0;
            while (!(d & 1))
            {
                d >>= 1;
                r++;
            }
            for (std::uint64_t a : bases)
            {
                std::uint64_t x = powmod(a, d, n);
                if (x == 1 || x == n - 1)
                {
                    continue;
                }
                bool composite = true;
                for (int i = 1; i < r; i++)
                {
                    x = mulmod(x, x, n);
                    if (x == n - 1)
                    {
                        composite = false;
                        break;
                    }
                }
                if (composite)
                {
                    return false;
                }
            }
            return true;
        }
    } // namespace util

    class Modulus
    {
    public:
        Modulus(std::uint64_t value = 0) : value_(value)
        {}
        std::uint64_t value() const noexcept
        {
            return value_;
        }
        int bit_count() const noexcept
        {
            int b = 0;
            for (std::uint64_t v = value_; v; v >>= 1)
            {
                b++;
            }
            return b;
        }
        bool is_zero() const noexcept
        {
            return value_ == 0;
        }
        bool operator==(const Modulus &o) const noexcept
        {
            return value_ == o.value_;
        }

    private:
        std::uint64_t value_;
    };

    class CoeffModulus
    {
    public:
        // CoeffModulus::Create(poly_modulus_degree, bit_sizes) (S/modulus.cpp:143-184): for every bit
        // size the largest primes below 2^bits congruent to 1 mod 2N, found downwards; a request takes
        // the smallest unused prime of its bit size (get_primes returns them in decreasing order and the
        // loop pops from the back).
        static std::vector<Modulus> Create(std::size_t poly_modulus_degree, const std::vector<int> &bit_sizes)
        {
            if (poly_modulus_degree < 2 || (poly_modulus_degree & (poly_modulus_degree - 1)))
            {
                throw std::invalid_argument("poly_modulus_degree is invalid");
            }
            std::map<int, std::vector<std::uint64_t>> found;
            const std::uint64_t factor = 2 * static_cast<std::uint64_t>(poly_modulus_degree);
            for (int b : bit_sizes)
            {
                if (b < 2 || b > 61)
                {
                    throw std::invalid_argument("bit_sizes is invalid");
                }
                found[b];
            }
            for (auto &kv : found)
            {
                const std::size_t count =
                    static_cast<std::size_t>(std::count(bit_sizes.begin(), bit_sizes.end(), kv.first));
                std::uint64_t value = ((std::uint64_t(1) << kv.first) - 1) / factor * factor + 1;
                const std::uint64_t lower = std::uint64_t(1) << (kv.first - 1);
                while (kv.second.size() < count && value > lower)
                {
                    if (util::is_prime(value))
                    {
                        kv.second.push_back(value);
                    }
                    value -= factor;
                }
                if (kv.second.size() < count)
                {
                    throw std::logic_error("failed to find enough qualifying primes");
                }
            }
            std::vector<Modulus> out;
            for (int b : bit_sizes)
            {
                out.emplace_back(found[b].back());
                found[b].pop_back();
            }
            return out;
        }
    };

    class EncryptionParameters
    {
    public:
        EncryptionParameters(scheme_type scheme = scheme_type::ckks) : scheme_(scheme)
        {
            if (scheme != scheme_type::ckks)
            {
                throw std::invalid_argument("the B200 backend evaluates CKKS only (MOAI never reaches BFV/BGV)");
            }
        }
        void set_poly_modulus_degree(std::size_t n)
        {
            poly_modulus_degree_ = n;
        }
        void set_coeff_modulus(const std::vector<Modulus> &m)
        {
            coeff_modulus_ = m;
        }
        // fork addition (S/encryptionparams.h): sparse ternary secrets; a client-side notion, recorded only
        void set_secret_key_hamming_weight(std::size_t h)
        {
            hamming_weight_ = h;
        }
        std::size_t secret_key_hamming_weight() const noexcept
        {
            return hamming_weight_;
        }
        scheme_type scheme() const noexcept
        {
            return scheme_;
        }
        std::size_t poly_modulus_degree() const noexcept
        {
            return poly_modulus_degree_;
        }
        const std::vector<Modulus> &coeff_modulus() const noexcept
        {
            return coeff_modulus_;
        }
        // CKKS has no plaintext modulus (S/encryptionparams.h:240-262); present for code that prints parameters
        Modulus plain_modulus() const noexcept
        {
            return Modulus(0);
        }
        const parms_id_type &parms_id() const noexcept
        {
            return parms_id_;
        }
        // S/encryptionparams.h:268-285: the source of randomness of client-side operations (Encryptor)
        void set_random_generator(std::shared_ptr<UniformRandomGeneratorFactory> f) noexcept
        {
            random_generator_ = std::move(f);
        }
        std::shared_ptr<UniformRandomGeneratorFactory> random_generator() const noexcept
        {
            return random_generator_ ? random_generator_ : UniformRandomGeneratorFactory::DefaultFactory();
        }

    private:
        friend class SEALContext;
        std::shared_ptr<UniformRandomGeneratorFactory> random_generator_;
        scheme_type scheme_;
        std::size_t poly_modulus_degree_ = 0;
        std::vector<Modulus> coeff_modulus_;
        std::size_t hamming_weight_ = 0;
        parms_id_type parms_id_ = parms_id_zero;
    };

    namespace detail
    {
        // SEAL throws std::invalid_argument / std::logic_error; the C ABI returns codes and keeps the text
        inline void chk(std::int32_t rc)
        {
            if (rc == MOAI_OK)
            {
                return;
            }
            const char *msg = moai_last_error();
            std::string text = (msg && *msg) ? msg : "libmoai_b200 call failed";
            if (rc == MOAI_INVALID_ARGUMENT)
            {
                throw std::invalid_argument(text);
            }
            if (rc == MOAI_LOGIC_ERROR)
            {
                throw std::logic_error(text);
            }
            if (rc == MOAI_OUT_OF_MEMORY)
            {
                throw std::bad_alloc();
            }
            throw std::runtime_error(text);
        }

        // one per SEALContext; shared by every object that owns device memory of that context
        struct ContextImpl;
        struct CtxMutex : std::recursive_mutex
        {
            const ContextImpl *owner = nullptr;
        };
        struct ContextImpl
        {
            moai_context *h = nullptr; // the root context (its stream serves the batched entry points)
            int log_n = 0;
            std::size_t n = 0;
            std::vector<std::uint64_t> primes; // key level: data primes then the special prime
            mutable CtxMutex mu;
            // Thread lanes (one CUDA stream per calling host thread).  Off: every call is serialised by `mu`
            // onto the root stream.  On (SEALContext::set_thread_lanes(true) or MOAI_FACADE_LANES=1): each host thread
            // gets its own lane of the context (moai_context_fork: shared tables, own stream and arena), calls are
            // issued WITHOUT the mutex and every call synchronises its lane before it returns — SEAL's calls are
            // synchronous, so an object produced by one thread is complete before another thread can touch it.
            mutable std::atomic<bool> lanes{ false };
            mutable std::mutex lanes_mu;
            mutable std::vector<moai_context *> lane_list;
            const std::uint64_t uid = next_uid();

            ContextImpl()
            {
                mu.owner = this;
                const char *e = std::getenv("MOAI_FACADE_LANES");
                lanes = e && e[0] == '1';
            }
            static std::uint64_t next_uid()
            {
                static std::atomic<std::uint64_t> n{ 1 };
                return n++;
            }
            // the context handle the CALLING THREAD issues work on
            moai_context *cur() const
            {
                if (!lanes.load(std::memory_order_relaxed))
                {
                    return h;
                }
                thread_local std::map<std::uint64_t, moai_context *> mine; // by context uid: survives address reuse
                auto it = mine.find(uid);
                if (it != mine.end())
                {
                    return it->second;
                }
                moai_context *lane = nullptr;
                chk(moai_context_fork(h, &lane));
                {
                    std::lock_guard<std::mutex> lk(lanes_mu);
                    lane_list.push_back(lane);
                }
                mine[uid] = lane;
                return lane;
            }

            ~ContextImpl()
            {
                for (moai_context *l : lane_list)
                {
                    moai_context_destroy(l);
                }
                if (h)
                {
                    moai_context_destroy(h);
                }
            }
            std::size_t key_limbs() const
            {
                return primes.size();
            }
            std::size_t first_limbs() const
            {
                return primes.size() > 1 ? primes.size() - 1 : 1;
            }
        };
        using ContextPtr = std::shared_ptr<ContextImpl>;
        // Scope of one facade call.  Lanes off: holds the context's mutex.  Lanes on: holds nothing and, on the way out,
        // waits for the calling thread's lane (the call's results are complete when it returns, like SEAL's).
        class Lock
        {
        public:
            explicit Lock(CtxMutex &m) : m_(m), locked_(!m.owner->lanes.load(std::memory_order_relaxed))
            {
                if (locked_)
                {
                    m_.lock();
                }
            }
            ~Lock()
            {
                if (locked_)
                {
                    m_.unlock();
                }
                else
                {
                    try
                    {
                        moai_synchronize(m_.owner->cur()); // the lane exists already unless the call never reached the device
                    }
                    catch (...)
                    {
                    }
                }
            }
            Lock(const Lock &) = delete;
            Lock &operator=(const Lock &) = delete;

        private:
            CtxMutex &m_;
            bool locked_;
        };
        // Scope of a call that runs on the ROOT stream whatever the lane mode (the batched module / bootstrapping entry
        // points, whose handles keep per-context caches): always the mutex; with lanes on it also drains the root stream
        // before returning so that other threads' lanes see complete results.
        class RootLock
        {
        public:
            explicit RootLock(CtxMutex &m) : m_(m)
            {
                m_.lock();
            }
            ~RootLock()
            {
                if (m_.owner->lanes.load(std::memory_order_relaxed))
                {
                    moai_synchronize(m_.owner->h);
                }
                m_.unlock();
            }
            RootLock(const RootLock &) = delete;
            RootLock &operator=(const RootLock &) = delete;

        private:
            CtxMutex &m_;
        };

        // stream-ordered device block from the library's arena (the role of S/util/mempool.h)
        class DeviceBlock
        {
        public:
            DeviceBlock() = default;
            DeviceBlock(const DeviceBlock &) = delete;
            DeviceBlock &operator=(const DeviceBlock &) = delete;
            DeviceBlock(DeviceBlock &&o) noexcept
            {
                swap(o);
            }
            DeviceBlock &operator=(DeviceBlock &&o) noexcept
            {
                if (this != &o)
                {
                    release();
                    swap(o);
                }
                return *this;
            }
            ~DeviceBlock()
            {
                release();
            }
            void swap(DeviceBlock &o) noexcept
            {
                std::swap(c_, o.c_);
                std::swap(p_, o.p_);
                std::swap(words_, o.words_);
            }
            void release() noexcept
            {
                if (p_ && c_)
                {
                    Lock lk(c_->mu);
                    moai_free(c_->cur(), p_);
                }
                p_ = nullptr;
                words_ = 0;
                c_.reset();
            }
            // capacity >= words afterwards; contents are NOT preserved when the block grows
            void ensure(const ContextPtr &c, std::size_t words)
            {
                if (p_ && c_ == c && words_ >= words)
                {
                    return;
                }
                release();
                if (!c)
                {
                    throw std::invalid_argument("encryption parameters are not set correctly");
                }
                if (words)
                {
                    Lock lk(c->mu);
                    void *p = nullptr;
                    chk(moai_malloc(c->cur(), words * sizeof(std::uint64_t), &p));
                    p_ = static_cast<std::uint64_t *>(p);
                }
                words_ = words;
                c_ = c;
            }
            std::uint64_t *ptr() const noexcept
            {
                return p_;
            }
            const ContextPtr &context() const noexcept
            {
                return c_;
            }

        private:
            ContextPtr c_;
            std::uint64_t *p_ = nullptr;
            std::size_t words_ = 0;
        };

        // SEAL's parms_id: BLAKE2b-256 over (scheme, N, the level's primes, plain_modulus = 0) as u64 words
        // (S/encryptionparams.cpp:124-158, S/util/hash.h:30-37)
        inline parms_id_type level_id(const ContextImpl &c, std::size_t limbs, bool key_level)
        {
            (void)key_level;
            std::vector<std::uint64_t> words;
            words.push_back(static_cast<std::uint64_t>(scheme_type::ckks));
            words.push_back(static_cast<std::uint64_t>(c.n));
            for (std::size_t i = 0; i < limbs; i++)
            {
                words.push_back(c.primes[i]);
            }
            words.push_back(0);
            parms_id_type id;
            util::blake2b(id.data(), 32, words.data(), words.size() * sizeof(std::uint64_t));
            return id;
        }
    } // namespace detail

    // ------------------------------------------------------------------------------------------
    // SEALContext and its chain of ContextData (S/context.h:93-560): key level (all primes), then one
    // level per dropped prime down to a single prime; chain_index counts down to 0.
    // ------------------------------------------------------------------------------------------
    class SEALContext
    {
    public:
        class ContextData
        {
        public:
            const EncryptionParameters &parms() const noexcept
            {
                return parms_;
            }
            const parms_id_type &parms_id() const noexcept
            {
                return parms_.parms_id();
            }
            std::size_t chain_index() const noexcept
            {
                return chain_index_;
            }
            int total_coeff_modulus_bit_count() const noexcept
            {
                return total_bits_;
            }
            std::shared_ptr<const ContextData> next_context_data() const noexcept
            {
                return next_;
            }
            std::shared_ptr<const ContextData> prev_context_data() const noexcept
            {
                return prev_.lock();
            }

        private:
            friend class SEALContext;
            EncryptionParameters parms_;
            std::size_t chain_index_ = 0;
            int total_bits_ = 0;
            std::shared_ptr<const ContextData> next_;
            std::weak_ptr<const ContextData> prev_;
        };

        // SEALContext(parms, expand_mod_chain, sec_level) (S/context.h:385-391) + the CUDA device to use.
        // Throws std::runtime_error when no CUDA device / library context can be created: there is no
        // CPU fallback.
        SEALContext(const EncryptionParameters &parms, bool expand_mod_chain = true,
                    sec_level_type sec_level = sec_level_type::tc128, int device = 0)
        {
            (void)sec_level;
            const std::size_t n = parms.poly_modulus_degree();
            int log_n = 0;
            while ((std::size_t(1) << log_n) < n)
            {
                log_n++;
            }
            if (n < 2 || (std::size_t(1) << log_n) != n || parms.coeff_modulus().empty())
            {
                throw std::invalid_argument("encryption parameters are not set correctly");
            }
            impl_ = std::make_shared<detail::ContextImpl>();
            impl_->log_n = log_n;
            impl_->n = n;
            for (auto &m : parms.coeff_modulus())
            {
                impl_->primes.push_back(m.value());
            }
            detail::chk(moai_context_create(log_n, impl_->primes.data(), static_cast<std::int32_t>(impl_->primes.size()),
                                            device, &impl_->h));
            // chain
            const std::size_t kl = impl_->primes.size();
            const std::size_t first = impl_->first_limbs();
            std::vector<std::shared_ptr<ContextData>> levels;
            auto make = [&](std::size_t limbs, bool key) {
                auto cd = std::make_shared<ContextData>();
                cd->parms_ = parms;
                cd->parms_.coeff_modulus_.assign(parms.coeff_modulus().begin(), parms.coeff_modulus().begin() + limbs);
                cd->parms_.parms_id_ = detail::level_id(*impl_, limbs, key);
                for (std::size_t i = 0; i < limbs; i++)
                {
                    cd->total_bits_ += cd->parms_.coeff_modulus_[i].bit_count();
                }
                return cd;
            };
            if (kl > 1)
            {
                levels.push_back(make(kl, true));
            }
            const std::size_t last = expand_mod_chain ? 1 : first;
            for (std::size_t limbs = first; limbs >= last; limbs--)
            {
                levels.push_back(make(limbs, false));
                if (limbs == 1)
                {
                    break;
                }
            }
            for (std::size_t i = 0; i < levels.size(); i++)
            {
                levels[i]->chain_index_ = levels.size() - 1 - i;
                if (i + 1 < levels.size())
                {
                    levels[i]->next_ = levels[i + 1];
                    levels[i + 1]->prev_ = levels[i];
                }
                by_id_[levels[i]->parms_id()] = levels[i];
            }
            key_ = levels.front();
            first_ = kl > 1 ? levels[1] : levels[0];
            last_ = levels.back();
        }

        std::shared_ptr<const ContextData> get_context_data(const parms_id_type &id) const
        {
            auto it = by_id_.find(id);
            return it == by_id_.end() ? nullptr : it->second;
        }
        std::shared_ptr<const ContextData> key_context_data() const
        {
            return key_;
        }
        std::shared_ptr<const ContextData> first_context_data() const
        {
            return first_;
        }
        std::shared_ptr<const ContextData> last_context_data() const
        {
            return last_;
        }
        const parms_id_type &key_parms_id() const
        {
            return key_->parms_id();
        }
        const parms_id_type &first_parms_id() const
        {
            return first_->parms_id();
        }
        const parms_id_type &last_parms_id() const
        {
            return last_->parms_id();
        }
        bool parameters_set() const
        {
            return impl_ && impl_->h;
        }
        bool using_keyswitching() const
        {
            return impl_->primes.size() > 1;
        }
        const char *parameter_error_message() const
        {
            return parameters_set() ? "valid" : "invalid";
        }

        // ---- extensions (not in SEAL) ----
        moai_context *handle() const
        {
            return impl_->h;
        }
        const detail::ContextPtr &impl() const
        {
            return impl_;
        }
        // the level with `limbs` data primes (what a raw ciphertext from the client is tagged with)
        const parms_id_type &parms_id_for_limbs(std::size_t limbs) const
        {
            for (auto cd = first_; cd; cd = cd->next_context_data())
            {
                if (cd->parms().coeff_modulus().size() == limbs)
                {
                    return cd->parms_id();
                }
            }
            throw std::invalid_argument("no level with that many limbs");
        }
        void synchronize() const
        {
            detail::Lock lk(impl_->mu);
            detail::chk(moai_synchronize(impl_->cur()));
        }
        // one CUDA stream (lane) per calling host thread instead of one mutex for all of them; switch it while no other
        // thread is inside a facade call (e.g. right after constructing the context)
        void set_thread_lanes(bool on) const
        {
            detail::chk(moai_synchronize(impl_->h));
            impl_->lanes = on;
        }
        bool thread_lanes() const
        {
            return impl_->lanes;
        }

    private:
        detail::ContextPtr impl_;
        std::map<parms_id_type, std::shared_ptr<const ContextData>> by_id_;
        std::shared_ptr<const ContextData> key_, first_, last_;
    };

    namespace detail
    {
        inline std::size_t limbs_of(const SEALContext &ctx, const parms_id_type &id, const char *what)
        {
            auto cd = ctx.get_context_data(id);
            if (!cd)
            {
                throw std::invalid_argument(std::string(what) + " is not valid for encryption parameters");
            }
            return cd->parms().coeff_modulus().size();
        }
    } // namespace detail

    namespace detail
    {
        // SEAL's stream framing (S/serialization.h:76-93): 16-byte header, little endian
        constexpr std::uint16_t seal_magic = 0xA15E;
        constexpr std::uint8_t seal_header_size = 0x10;

        template <typename T>
        inline void put(std::ostream &o, const T &v)
        {
            o.write(reinterpret_cast<const char *>(&v), sizeof(T));
        }
        template <typename T>
        inline T get(std::istream &in)
        {
            T v;
            in.read(reinterpret_cast<char *>(&v), sizeof(T));
            if (!in)
            {
                throw std::runtime_error("I/O error");
            }
            return v;
        }
        inline void put_header(std::ostream &o, std::uint64_t total_size)
        {
            put<std::uint16_t>(o, seal_magic);
            put<std::uint8_t>(o, seal_header_size);
            put<std::uint8_t>(o, 4); // version 4.1
            put<std::uint8_t>(o, 1);
            put<std::uint8_t>(o, 0); // compr_mode_type::none
            put<std::uint16_t>(o, 0);
            put<std::uint64_t>(o, total_size);
        }
        // returns the total size (header included) the header announces
        inline std::uint64_t get_header(std::istream &in)
        {
            const auto magic = get<std::uint16_t>(in);
            const auto hsize = get<std::uint8_t>(in);
            const auto vmaj = get<std::uint8_t>(in);
            (void)get<std::uint8_t>(in);
            const auto compr = get<std::uint8_t>(in);
            (void)get<std::uint16_t>(in);
            const auto total = get<std::uint64_t>(in);
            if (magic != seal_magic || hsize != seal_header_size)
            {
                throw std::logic_error("loaded SEALHeader is invalid");
            }
            if (vmaj != 4) // version 3 streams have no correction_factor field; the parser below reads the v4 layout only
            {
                throw std::logic_error("incompatible version");
            }
            if (compr != 0)
            {
                throw std::logic_error("unsupported compression mode (save with compr_mode_type::none)");
            }
            return total;
        }

        // One serialized Ciphertext / PublicKey (S/ciphertext.cpp:190-330) -> host residues [size][limbs][N].
        // A seeded object stores c0 and a PRNG seed; c1 is expanded at the object's own level
        // (Ciphertext::expand_seed, S/ciphertext.cpp:118-136).
        // Seeded key digits are expanded ON THE DEVICE (moai_expand_seeds, csrc/seedexpand.cu) when the bound library is
        // the CUDA backend: the loader then leaves c1 zero, hands the seeds to the uploader, and the uploader expands
        // them straight into the key's device buffer.  (The CPU test double of the C ABI, moai_version() < 0, and
        // MOAI_FACADE_HOST_SEEDS=1 keep SEAL's host-side expansion.)
        inline bool device_seed_expansion()
        {
            static const bool on = [] {
                const char *e = std::getenv("MOAI_FACADE_HOST_SEEDS");
                return moai_version() >= 100 && !(e && e[0] == '1');
            }();
            return on;
        }
        struct PendingSeeds
        {
            std::vector<std::uint64_t> seeds;  // 8 words per seeded digit
            std::vector<std::size_t> digits;   // which digits of the key they belong to
            // dev: the key's device buffer [digits][2][kl][N]
            void expand(const ContextPtr &c, std::uint64_t *dev) const
            {
                const std::size_t kl = c->key_limbs(), per = 2 * kl * c->n;
                Lock lk(c->mu);
                std::size_t i = 0;
                while (i < digits.size())
                {
                    std::size_t run = 1; // consecutive digits go in one call
                    while (i + run < digits.size() && digits[i + run] == digits[i] + run)
                    {
                        run++;
                    }
                    chk(moai_expand_seeds(c->cur(), seeds.data() + 8 * i, std::int64_t(run), std::int32_t(kl),
                                          dev + digits[i] * per + kl * c->n, std::int64_t(per)));
                    i += run;
                }
            }
        };

        struct LoadedCiphertext
        {
            parms_id_type parms_id;
            bool is_ntt_form = true, was_seeded = false, seed_deferred = false;
            prng_seed_type seed{};
            std::size_t size = 0, limbs = 0;
            double scale = 1.0;
            std::vector<std::uint64_t> data;
        };
        inline LoadedCiphertext load_ciphertext_stream(const ContextImpl &c, std::istream &in, bool defer_seed = false)
        {
            LoadedCiphertext r;
            (void)get_header(in);
            in.read(reinterpret_cast<char *>(r.parms_id.data()), 32);
            r.is_ntt_form = get<std::uint8_t>(in) != 0;
            r.size = static_cast<std::size_t>(get<std::uint64_t>(in));
            const auto n = get<std::uint64_t>(in);
            r.limbs = static_cast<std::size_t>(get<std::uint64_t>(in));
            r.scale = get<double>(in);
            (void)get<std::uint64_t>(in); // correction_factor (BGV only)
            if (n != c.n || r.limbs < 1 || r.limbs > c.key_limbs() || r.size < 2 || r.size > 6 ||
                r.parms_id != level_id(c, r.limbs, false))
            {
                throw std::logic_error("ciphertext data is invalid");
            }
            (void)get_header(in); // the DynArray's own header
            const auto count = get<std::uint64_t>(in);
            const std::uint64_t total = std::uint64_t(r.size) * r.limbs * c.n;
            if (count != total && !(r.size == 2 && count == total / 2))
            {
                throw std::logic_error("ciphertext data is invalid");
            }
            r.data.resize(total);
            in.read(reinterpret_cast<char *>(r.data.data()), std::streamsize(count * sizeof(std::uint64_t)));
            if (!in)
            {
                throw std::runtime_error("I/O error");
            }
            // is_data_valid_for (S/valcheck.cpp:201-260): every residue below its prime — the kernels assume canonical
            // input (x + q - y, FP64 NTT inputs below 2^52), so a malformed stream must be refused here
            for (std::uint64_t w = 0; w < count; w++)
            {
                if (r.data[w] >= c.primes[(w / c.n) % r.limbs])
                {
                    throw std::logic_error("ciphertext data is invalid");
                }
            }
            if (count != total)
            {
                // UniformRandomGeneratorInfo: header, prng_type, 64-byte seed (S/randomgen.cpp:99-121)
                (void)get_header(in);
                const auto type = static_cast<prng_type>(get<std::uint8_t>(in));
                prng_seed_type seed;
                in.read(reinterpret_cast<char *>(seed.data()), prng_seed_byte_count);
                if (!in)
                {
                    throw std::runtime_error("I/O error");
                }
                if (defer_seed && type == prng_type::blake2xb && device_seed_expansion())
                {
                    r.seed = seed; // c1 stays zero: the uploader expands the seed on the device
                    r.seed_deferred = true;
                }
                else
                {
                    std::vector<std::uint64_t> primes(c.primes.begin(), c.primes.begin() + r.limbs);
                    util::sample_poly_uniform(UniformRandomGeneratorInfo(type, seed).make_prng(), primes, c.n,
                                              r.data.data() + total / 2);
                }
                r.was_seeded = true;
            }
            return r;
        }
    } // namespace detail

    // ------------------------------------------------------------------------------------------
    // Plaintext (S/plaintext.h): CKKS plaintexts are NTT-form residues [limbs][N] on the device.  A
    // scalar encoding (CKKSEncoder::encode(double, ...), S/ckks.cpp:77-216) is the same constant in every
    // NTT slot of a limb, so it is kept as `limbs` host constants and applied by the scalar kernels.
    // ------------------------------------------------------------------------------------------
    class Plaintext
    {
    public:
        Plaintext() = default;
        Plaintext(const Plaintext &o)
        {
            *this = o;
        }
        Plaintext(Plaintext &&) noexcept = default;
        Plaintext &operator=(Plaintext &&) noexcept = default;
        Plaintext &operator=(const Plaintext &o)
        {
            if (this == &o)
            {
                return *this;
            }
            limbs_ = o.limbs_;
            scale_ = o.scale_;
            pid_ = o.pid_;
            scalar_ = o.scalar_;
            consts_ = o.consts_;
            if (!o.scalar_ && o.limbs_ && o.mem_.context())
            {
                auto c = o.mem_.context();
                mem_.ensure(c, limbs_ * c->n);
                detail::Lock lk(c->mu);
                detail::chk(moai_memcpy_d2d(c->cur(), mem_.ptr(), o.mem_.ptr(), limbs_ * c->n * sizeof(std::uint64_t)));
            }
            else
            {
                mem_.release();
            }
            return *this;
        }

        double &scale() noexcept
        {
            return scale_;
        }
        const double &scale() const noexcept
        {
            return scale_;
        }
        parms_id_type &parms_id() noexcept
        {
            return pid_;
        }
        const parms_id_type &parms_id() const noexcept
        {
            return pid_;
        }
        bool is_ntt_form() const noexcept
        {
            return pid_ != parms_id_zero;
        }
        std::size_t coeff_modulus_size() const noexcept
        {
            return limbs_;
        }
        std::size_t coeff_count() const noexcept
        {
            return mem_.context() ? limbs_ * mem_.context()->n : 0;
        }
        // DEVICE pointer ([limbs][N]); nullptr for a scalar encoding
        std::uint64_t *data() noexcept
        {
            return scalar_ ? nullptr : mem_.ptr();
        }
        const std::uint64_t *data() const noexcept
        {
            return scalar_ ? nullptr : mem_.ptr();
        }
        bool is_scalar() const noexcept
        {
            return scalar_;
        }
        const std::vector<std::uint64_t> &scalar_consts() const noexcept
        {
            return consts_;
        }

        // ---- extensions: raw residues [limbs][N] (NTT form) from / to the host ----
        void upload(const SEALContext &ctx, const std::uint64_t *host, std::size_t limbs, double scale)
        {
            set_vector(ctx, limbs, scale);
            auto &c = ctx.impl();
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_h2d(c->cur(), mem_.ptr(), host, limbs * c->n * sizeof(std::uint64_t)));
            detail::chk(moai_synchronize(c->cur()));
        }
        void download(std::uint64_t *host) const
        {
            auto c = mem_.context();
            if (scalar_ || !c)
            {
                throw std::logic_error("plaintext holds no residue array");
            }
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_d2h(c->cur(), host, mem_.ptr(), limbs_ * c->n * sizeof(std::uint64_t)));
        }

        // used by CKKSEncoder / Evaluator / Decryptor
        void set_vector(const SEALContext &ctx, std::size_t limbs, double scale)
        {
            mem_.ensure(ctx.impl(), limbs * ctx.impl()->n);
            limbs_ = limbs;
            scale_ = scale;
            pid_ = ctx.parms_id_for_limbs(limbs);
            scalar_ = false;
            consts_.clear();
        }
        void set_scalar(const SEALContext &ctx, std::vector<std::uint64_t> consts, double scale)
        {
            mem_.release();
            limbs_ = consts.size();
            scale_ = scale;
            pid_ = ctx.parms_id_for_limbs(limbs_);
            scalar_ = true;
            consts_ = std::move(consts);
        }
        // mod_switch_to for an NTT-form plaintext drops trailing limbs (S/evaluator.cpp:1566-1581)
        void drop_to(const SEALContext &ctx, std::size_t limbs)
        {
            limbs_ = limbs;
            pid_ = ctx.parms_id_for_limbs(limbs);
            if (scalar_)
            {
                consts_.resize(limbs);
            }
        }

    private:
        detail::DeviceBlock mem_;
        std::size_t limbs_ = 0;
        double scale_ = 1.0;
        parms_id_type pid_ = parms_id_zero;
        bool scalar_ = false;
        std::vector<std::uint64_t> consts_;
    };

    // ------------------------------------------------------------------------------------------
    // Ciphertext (S/ciphertext.h): value semantics, [size][limbs][N] residues on the device
    // ------------------------------------------------------------------------------------------
    class Ciphertext
    {
    public:
        Ciphertext() = default;
        explicit Ciphertext(const SEALContext &ctx)
        {
            resize(ctx, ctx.first_parms_id(), 2);
        }
        Ciphertext(const SEALContext &ctx, const parms_id_type &id, std::size_t size = 2)
        {
            resize(ctx, id, size);
        }
        Ciphertext(const Ciphertext &o)
        {
            *this = o;
        }
        Ciphertext(Ciphertext &&) noexcept = default;
        Ciphertext &operator=(Ciphertext &&) noexcept = default;
        // deep copy (S/ciphertext.h:701-715): `nx[i] = x[i]`, `vector<Ciphertext> c_g(g, enc_W[i])`
        Ciphertext &operator=(const Ciphertext &o)
        {
            if (this == &o)
            {
                return *this;
            }
            size_ = o.size_;
            limbs_ = o.limbs_;
            scale_ = o.scale_;
            pid_ = o.pid_;
            ntt_ = o.ntt_;
            auto c = o.mem_.context();
            if (c && o.words())
            {
                mem_.ensure(c, o.words());
                detail::Lock lk(c->mu);
                detail::chk(moai_memcpy_d2d(c->cur(), mem_.ptr(), o.mem_.ptr(), o.words() * sizeof(std::uint64_t)));
            }
            else
            {
                mem_.release();
            }
            return *this;
        }

        void resize(const SEALContext &ctx, const parms_id_type &id, std::size_t size)
        {
            const std::size_t limbs = detail::limbs_of(ctx, id, "parms_id");
            if (size != 0 && (size < 2 || size > 6))
            {
                throw std::invalid_argument("invalid size");
            }
            shape(ctx.impl(), size, limbs);
            pid_ = id;
        }
        void release()
        {
            mem_.release();
            size_ = limbs_ = 0;
            pid_ = parms_id_zero;
            scale_ = 1.0;
        }

        std::size_t size() const noexcept
        {
            return size_;
        }
        std::size_t coeff_modulus_size() const noexcept
        {
            return limbs_;
        }
        std::size_t poly_modulus_degree() const noexcept
        {
            return mem_.context() ? mem_.context()->n : 0;
        }
        bool &is_ntt_form() noexcept
        {
            return ntt_;
        }
        bool is_ntt_form() const noexcept
        {
            return ntt_;
        }
        parms_id_type &parms_id() noexcept
        {
            return pid_;
        }
        const parms_id_type &parms_id() const noexcept
        {
            return pid_;
        }
        double &scale() noexcept
        {
            return scale_;
        } // callers write it: Ct_pt_matrix_mul.hpp:41
        const double &scale() const noexcept
        {
            return scale_;
        }
        // DEVICE pointers
        std::uint64_t *data() noexcept
        {
            return mem_.ptr();
        }
        const std::uint64_t *data() const noexcept
        {
            return mem_.ptr();
        }
        std::uint64_t *data(std::size_t poly_index)
        {
            return mem_.ptr() + poly_index * limbs_ * poly_modulus_degree();
        }
        const std::uint64_t *data(std::size_t poly_index) const
        {
            return mem_.ptr() + poly_index * limbs_ * poly_modulus_degree();
        }

        // ---- extensions: raw residues in SEAL's layout [size][limbs][N] from / to the host
        //      (Ciphertext::data() of stock SEAL on the client side)
        void upload(const SEALContext &ctx, const std::uint64_t *host, std::size_t size, std::size_t limbs, double scale)
        {
            resize(ctx, ctx.parms_id_for_limbs(limbs), size);
            scale_ = scale;
            ntt_ = true;
            auto &c = ctx.impl();
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_h2d(c->cur(), mem_.ptr(), host, words() * sizeof(std::uint64_t)));
            detail::chk(moai_synchronize(c->cur()));
        }
        void download(std::uint64_t *host) const
        {
            auto c = mem_.context();
            if (!c || !words())
            {
                throw std::logic_error("ciphertext is empty");
            }
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_d2h(c->cur(), host, mem_.ptr(), words() * sizeof(std::uint64_t)));
        }

        // ---- SEAL's wire format (Ciphertext::save / load with compr_mode_type::none, S/ciphertext.cpp:153-330):
        //      byte-identical to the stock library, so encrypted inputs / outputs / checkpoints interoperate
        std::streamoff save_size() const
        {
            return std::streamoff(16 + 32 + 1 + 24 + 8 + 8 + 16 + 8 + words() * sizeof(std::uint64_t));
        }
        std::streamoff save(std::ostream &stream) const
        {
            auto c = mem_.context();
            if (!c || !words())
            {
                throw std::logic_error("ciphertext is empty");
            }
            std::vector<std::uint64_t> host(words());
            download(host.data());
            const std::uint64_t dyn = 16 + 8 + host.size() * sizeof(std::uint64_t);
            detail::put_header(stream, std::uint64_t(save_size()));
            stream.write(reinterpret_cast<const char *>(pid_.data()), 32);
            detail::put<std::uint8_t>(stream, ntt_ ? 1 : 0);
            detail::put<std::uint64_t>(stream, size_);
            detail::put<std::uint64_t>(stream, c->n);
            detail::put<std::uint64_t>(stream, limbs_);
            detail::put<double>(stream, scale_);
            detail::put<std::uint64_t>(stream, 1); // correction_factor
            detail::put_header(stream, dyn);
            detail::put<std::uint64_t>(stream, host.size());
            stream.write(reinterpret_cast<const char *>(host.data()), std::streamsize(host.size() * sizeof(std::uint64_t)));
            if (!stream)
            {
                throw std::runtime_error("I/O error");
            }
            return save_size();
        }
        // accepts the seeded form as well (Encryptor::encrypt_symmetric(...).save() of a stock client)
        void load(const SEALContext &ctx, std::istream &stream)
        {
            detail::LoadedCiphertext r = detail::load_ciphertext_stream(*ctx.impl(), stream);
            if (!ctx.get_context_data(r.parms_id))
            {
                throw std::logic_error("ciphertext data is invalid");
            }
            upload(ctx, r.data.data(), r.size, r.limbs, r.scale);
            ntt_ = r.is_ntt_form;
        }

        // internal: (re)shape without preserving contents
        void shape(const detail::ContextPtr &c, std::size_t size, std::size_t limbs)
        {
            mem_.ensure(c, size * limbs * c->n);
            size_ = size;
            limbs_ = limbs;
        }
        std::size_t words() const noexcept
        {
            return mem_.context() ? size_ * limbs_ * mem_.context()->n : 0;
        }
        const detail::ContextPtr &context() const noexcept
        {
            return mem_.context();
        }
        void swap_storage(Ciphertext &o) noexcept
        {
            mem_.swap(o.mem_);
            std::swap(size_, o.size_);
            std::swap(limbs_, o.limbs_);
        }

    private:
        detail::DeviceBlock mem_;
        std::size_t size_ = 0, limbs_ = 0;
        double scale_ = 1.0;
        parms_id_type pid_ = parms_id_zero;
        bool ntt_ = true; // CKKS ciphertexts live in NTT form (S/encryptor.cpp:262-275)
    };

    // ------------------------------------------------------------------------------------------
    // keys (S/secretkey.h, S/relinkeys.h, S/galoiskeys.h, S/kswitchkeys.h): generated by stock SEAL on
    // the client, uploaded as raw residues
    // ------------------------------------------------------------------------------------------
    // PublicKey (S/publickey.h): one size-2 "ciphertext" at the key level, [2][key limbs][N]
    class PublicKey
    {
    public:
        PublicKey() = default;
        // PublicKey::data().data() of SEAL
        void upload(const SEALContext &ctx, const std::uint64_t *host)
        {
            auto &c = ctx.impl();
            mem_ = std::make_shared<detail::DeviceBlock>();
            mem_->ensure(c, 2 * c->key_limbs() * c->n);
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_h2d(c->cur(), mem_->ptr(), host, 2 * c->key_limbs() * c->n * sizeof(std::uint64_t)));
            detail::chk(moai_synchronize(c->cur()));
        }
        // what PublicKey::save writes (a Ciphertext stream; seeded form accepted)
        void load(const SEALContext &ctx, std::istream &stream)
        {
            detail::LoadedCiphertext r = detail::load_ciphertext_stream(*ctx.impl(), stream);
            if (r.size != 2 || r.limbs != ctx.impl()->key_limbs())
            {
                throw std::logic_error("PublicKey data is invalid");
            }
            upload(ctx, r.data.data());
        }
        const std::uint64_t *data() const noexcept
        {
            return mem_ ? mem_->ptr() : nullptr;
        }
        const detail::ContextPtr &context() const
        {
            static const detail::ContextPtr none;
            return mem_ ? mem_->context() : none;
        }

    private:
        std::shared_ptr<detail::DeviceBlock> mem_;
    };

    class SecretKey
    {
    public:
        SecretKey() = default;
        // SecretKey::data() of SEAL: [key limbs][N], NTT form
        void upload(const SEALContext &ctx, const std::uint64_t *host)
        {
            auto &c = ctx.impl();
            mem_ = std::make_shared<detail::DeviceBlock>();
            mem_->ensure(c, c->key_limbs() * c->n);
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_h2d(c->cur(), mem_->ptr(), host, c->key_limbs() * c->n * sizeof(std::uint64_t)));
            detail::chk(moai_synchronize(c->cur()));
        }
        const std::uint64_t *data() const noexcept
        {
            return mem_ ? mem_->ptr() : nullptr;
        }

    private:
        std::shared_ptr<detail::DeviceBlock> mem_;
    };

    namespace detail
    {
        struct KeySet
        {
            ContextPtr c;
            moai_keys *h = nullptr;
            std::vector<std::unique_ptr<DeviceBlock>> blocks;
            std::map<std::uint32_t, const std::uint64_t *> galois; // SEAL-layout keys by Galois element
            const std::uint64_t *relin = nullptr;
            struct FastKey
            {
                std::uint32_t elt;
                const std::uint64_t *p;
                int key_limbs;
            };
            std::vector<FastKey> fast; // level-truncated, pre-permuted keys (fast mode)

            explicit KeySet(const ContextPtr &ctx) : c(ctx)
            {
                Lock lk(c->mu);
                chk(moai_keys_create(c->cur(), &h));
            }
            ~KeySet()
            {
                if (h)
                {
                    Lock lk(c->mu);
                    moai_keys_destroy(h);
                }
            }
            // one KSwitchKeys entry: [key_limbs-1 digits][2][key_limbs][N]
            std::uint64_t *store(const std::uint64_t *host)
            {
                const std::size_t kl = c->key_limbs();
                const std::size_t words = (kl - 1) * 2 * kl * c->n;
                blocks.emplace_back(new DeviceBlock());
                blocks.back()->ensure(c, words);
                Lock lk(c->mu);
                chk(moai_memcpy_h2d(c->cur(), blocks.back()->ptr(), host, words * sizeof(std::uint64_t)));
                chk(moai_synchronize(c->cur()));
                return blocks.back()->ptr();
            }
        };
    } // namespace detail

    class KSwitchKeys
    {
    public:
        moai_keys *handle() const
        {
            return set_ ? set_->h : nullptr;
        }

    protected:
        void need(const SEALContext &ctx)
        {
            if (!set_ || set_->c != ctx.impl())
            {
                set_ = std::make_shared<detail::KeySet>(ctx.impl());
            }
        }
        // seeds of the key being loaded whose uniform halves are still to be expanded on the device (load_entries)
        static detail::PendingSeeds &pending_ref()
        {
            static thread_local detail::PendingSeeds p;
            return p;
        }
        // KSwitchKeys::load_members (S/kswitchkeys.cpp:86-150): parms_id, then a vector (by key index) of vectors
        // (one PublicKey per decomposition digit).  Calls `sink(index, host [digits][2][kl][N])` per present key.
        // Seeded entries — what `keygen.create_relin_keys()` / `create_galois_keys(...)` return for shipping —
        // are expanded on the fly; returns how many digits were seeded.
        template <typename Sink>
        static std::size_t load_entries(const SEALContext &ctx, std::istream &stream, Sink &&sink)
        {
            const detail::ContextImpl &c = *ctx.impl();
            (void)detail::get_header(stream);
            parms_id_type id;
            stream.read(reinterpret_cast<char *>(id.data()), 32);
            if (!stream || id != ctx.key_parms_id())
            {
                throw std::logic_error("KSwitchKeys data is invalid");
            }
            const auto dim1 = detail::get<std::uint64_t>(stream);
            if (dim1 > c.n) // at most one key per Galois element (S/galoiskeys.h:52-56: index = (elt - 1) / 2 < N)
            {
                throw std::logic_error("KSwitchKeys data is invalid");
            }
            const std::size_t kl = c.key_limbs(), per = 2 * kl * c.n;
            std::size_t seeded = 0;
            for (std::uint64_t index = 0; index < dim1; index++)
            {
                const auto dim2 = detail::get<std::uint64_t>(stream);
                if (dim2 == 0)
                {
                    continue;
                }
                if (dim2 != kl - 1)
                {
                    throw std::logic_error("KSwitchKeys data is invalid");
                }
                std::vector<std::uint64_t> key(std::size_t(dim2) * per);
                pending_ref() = detail::PendingSeeds();
                for (std::uint64_t j = 0; j < dim2; j++)
                {
                    detail::LoadedCiphertext r = detail::load_ciphertext_stream(c, stream, /*defer_seed=*/true);
                    if (r.size != 2 || r.limbs != kl || !r.is_ntt_form)
                    {
                        throw std::logic_error("KSwitchKeys data is invalid");
                    }
                    seeded += r.was_seeded;
                    if (r.seed_deferred)
                    {
                        pending_ref().seeds.insert(pending_ref().seeds.end(), r.seed.begin(), r.seed.end());
                        pending_ref().digits.push_back(std::size_t(j));
                    }
                    std::copy(r.data.begin(), r.data.end(), key.begin() + std::size_t(j) * per);
                }
                sink(static_cast<std::size_t>(index), key.data()); // the uploaders below consume pending_ref()
                pending_ref() = detail::PendingSeeds();
            }
            return seeded;
        }
        std::shared_ptr<detail::KeySet> set_;
        void expand_pending(std::uint64_t *dev)
        {
            if (!pending_ref().digits.empty())
            {
                pending_ref().expand(set_->c, dev);
                detail::Lock lk(set_->c->mu);
                detail::chk(moai_synchronize(set_->c->cur()));
            }
        }
    };

    class RelinKeys : public KSwitchKeys
    {
    public:
        // RelinKeys::key(2) of SEAL, digit after digit: [digits][2][key_limbs][N]
        void upload(const SEALContext &ctx, const std::uint64_t *host)
        {
            need(ctx);
            set_->relin = set_->store(host);
            expand_pending(const_cast<std::uint64_t *>(set_->relin));
            detail::Lock lk(set_->c->mu);
            detail::chk(moai_keys_set_relin(set_->h, set_->relin));
        }
        // RelinKeys::load (S/relinkeys.h, S/kswitchkeys.cpp:86-150); index = key_power - 2.  Returns the number of
        // seeded digits that were expanded.
        std::size_t load(const SEALContext &ctx, std::istream &stream)
        {
            return load_entries(ctx, stream, [&](std::size_t index, const std::uint64_t *host) {
                if (index != 0)
                {
                    throw std::logic_error("only the key for s^2 is supported (MOAI relinearises after every product)");
                }
                upload(ctx, host);
            });
        }
        bool has_key(std::size_t key_power) const
        {
            return key_power == 2 && set_ && set_->relin;
        }
        const std::uint64_t *device_key() const
        {
            return set_ ? set_->relin : nullptr;
        }
    };

    class GaloisKeys : public KSwitchKeys
    {
    public:
        // GaloisKeys::key(galois_elt) of SEAL, digit after digit
        void upload(const SEALContext &ctx, std::uint32_t galois_elt, const std::uint64_t *host)
        {
            need(ctx);
            const std::uint64_t *p = set_->store(host);
            expand_pending(const_cast<std::uint64_t *>(p));
            set_->galois[galois_elt] = p;
            detail::Lock lk(set_->c->mu);
            detail::chk(moai_keys_add_galois(set_->h, galois_elt, p));
        }
        // fast mode: re-lay an uploaded key out level-truncated and pre-permuted;
        // rotations by that element then use the hoisting-capable path (not SEAL's residues)
        void upload_fast(const SEALContext &ctx, std::uint32_t galois_elt, const std::uint64_t *host, int max_limbs)
        {
            need(ctx);
            auto &c = set_->c;
            detail::DeviceBlock full;
            const std::size_t kl = c->key_limbs();
            const std::size_t words = (kl - 1) * 2 * kl * c->n;
            full.ensure(c, words);
            set_->blocks.emplace_back(new detail::DeviceBlock());
            set_->blocks.back()->ensure(c, std::size_t(max_limbs) * 2 * (max_limbs + 1) * c->n);
            {
                detail::Lock lk0(c->mu);
                detail::chk(moai_memcpy_h2d(c->cur(), full.ptr(), host, words * sizeof(std::uint64_t)));
            }
            expand_pending(full.ptr());
            detail::Lock lk(c->mu);
            detail::chk(moai_key_prepare(c->cur(), full.ptr(), galois_elt, max_limbs, 1, set_->blocks.back()->ptr()));
            detail::chk(moai_keys_add_galois_fast(set_->h, galois_elt, set_->blocks.back()->ptr(), max_limbs + 1));
            detail::chk(moai_synchronize(c->cur()));
            set_->fast.push_back({ galois_elt, set_->blocks.back()->ptr(), max_limbs + 1 });
        }
        const std::shared_ptr<detail::KeySet> &key_set() const
        {
            return set_;
        }
        // GaloisKeys::load; index = (galois_elt - 1) / 2 (S/galoiskeys.h:52-56).  With max_limbs > 0 every key is
        // re-laid out level-truncated and pre-permuted for the fast mode (upload_fast) instead of SEAL's layout.
        std::size_t load(const SEALContext &ctx, std::istream &stream, int fast_max_limbs = 0)
        {
            return load_entries(ctx, stream, [&](std::size_t index, const std::uint64_t *host) {
                const std::uint32_t elt = static_cast<std::uint32_t>(2 * index + 1);
                if (fast_max_limbs > 0)
                {
                    upload_fast(ctx, elt, host, fast_max_limbs);
                }
                else
                {
                    upload(ctx, elt, host);
                }
            });
        }
        bool has_key(std::uint32_t galois_elt) const
        {
            return set_ && set_->galois.count(galois_elt) != 0;
        }
        const std::uint64_t *device_key(std::uint32_t galois_elt) const
        {
            if (!set_)
            {
                return nullptr;
            }
            auto it = set_->galois.find(galois_elt);
            return it != set_->galois.end() ? it->second : nullptr;
        }
    };

    // ------------------------------------------------------------------------------------------
    // CKKSEncoder (S/ckks.h:148-432)
    // ------------------------------------------------------------------------------------------
    class CKKSEncoder
    {
    public:
        explicit CKKSEncoder(const SEALContext &ctx) : ctx_(ctx)
        {}

        std::size_t slot_count() const noexcept
        {
            return ctx_.impl()->n >> 1;
        }

        // encode(vector<double> | vector<complex<double>>, parms_id, scale, destination) (S/ckks.h:457-638):
        // the device FFT reproduces SEAL's operation order, the residues are SEAL's bit for bit
        void encode(const std::vector<std::complex<double>> &values, const parms_id_type &id, double scale,
                    Plaintext &destination) const
        {
            const std::size_t limbs = detail::limbs_of(ctx_, id, "parms_id");
            if (values.size() > slot_count())
            {
                throw std::invalid_argument("values_size is too large");
            }
            check_scale(scale, limbs);
            auto &c = ctx_.impl();
            destination.set_vector(ctx_, limbs, scale);
            detail::DeviceBlock dv;
            const std::size_t count = std::max<std::size_t>(values.size(), 1);
            dv.ensure(c, 2 * count);
            std::vector<std::complex<double>> one(1);
            const std::complex<double> *src = values.empty() ? one.data() : values.data();
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_h2d(c->cur(), dv.ptr(), src, 2 * count * sizeof(double)));
            detail::chk(moai_encode_vector(c->cur(), reinterpret_cast<const double *>(dv.ptr()), 1,
                                           static_cast<std::int32_t>(count), scale, static_cast<std::int32_t>(limbs),
                                           destination.data()));
            detail::chk(moai_synchronize(c->cur())); // `src` may be a temporary
        }
        void encode(const std::vector<double> &values, const parms_id_type &id, double scale, Plaintext &destination) const
        {
            std::vector<std::complex<double>> v(values.begin(), values.end());
            encode(v, id, scale, destination);
        }
        template <typename T>
        void encode(const std::vector<T> &values, double scale, Plaintext &destination) const
        {
            encode(values, ctx_.first_parms_id(), scale, destination);
        }
        // encode(double, parms_id, scale, destination) (S/ckks.cpp:77-216)
        void encode(double value, const parms_id_type &id, double scale, Plaintext &destination) const
        {
            const std::size_t limbs = detail::limbs_of(ctx_, id, "parms_id");
            check_scale(scale, limbs);
            std::vector<std::uint64_t> consts(limbs);
            auto &c = ctx_.impl();
            {
                detail::Lock lk(c->mu);
                detail::chk(moai_encode_scalar_consts(c->cur(), value, scale, static_cast<std::int32_t>(limbs), consts.data()));
            }
            destination.set_scalar(ctx_, std::move(consts), scale);
        }
        void encode(double value, double scale, Plaintext &destination) const
        {
            encode(value, ctx_.first_parms_id(), scale, destination);
        }
        // encode(int64, parms_id, destination) (S/ckks.cpp:218-289): exact integer, scale 1
        void encode(std::int64_t value, const parms_id_type &id, Plaintext &destination) const
        {
            encode(static_cast<double>(value), id, 1.0, destination);
        }

        // decode (S/ckks.h:644-760).  Host computation on a downloaded plaintext: inverse NTT on the
        // device, mixed-radix CRT composition to a centred integer, division by the scale, and the
        // canonical embedding evaluated by one FFT.  Agrees with SEAL's decode to floating-point rounding
        // (not bit for bit: different operation order); the reference only prints the result.
        void decode(const Plaintext &plain, std::vector<std::complex<double>> &destination) const
        {
            auto &c = ctx_.impl();
            const std::size_t n = c->n, limbs = plain.coeff_modulus_size();
            if (!limbs || !plain.is_ntt_form())
            {
                throw std::invalid_argument("plain is not in NTT form");
            }
            std::vector<std::uint64_t> res(limbs * n);
            if (plain.is_scalar())
            {
                // a constant polynomial: coefficient 0 only
                std::fill(res.begin(), res.end(), 0);
                for (std::size_t l = 0; l < limbs; l++)
                {
                    res[l * n] = plain.scalar_consts()[l];
                }
            }
            else
            {
                detail::DeviceBlock tmp;
                tmp.ensure(c, limbs * n);
                detail::Lock lk(c->mu);
                detail::chk(moai_memcpy_d2d(c->cur(), tmp.ptr(), plain.data(), limbs * n * sizeof(std::uint64_t)));
                detail::chk(moai_ntt_inverse(c->cur(), tmp.ptr(), 1, 1, static_cast<std::int32_t>(limbs)));
                detail::chk(moai_memcpy_d2h(c->cur(), res.data(), tmp.ptr(), limbs * n * sizeof(std::uint64_t)));
            }
            std::vector<std::complex<double>> a(n);
            compose_centered(*c, res, limbs, plain.scale(), a);
            // slot j = p(zeta^(5^j mod 2N)), zeta = exp(i*pi/N): B_t = p(zeta^(2t+1)) = sum_k (c_k zeta^k) w^(tk)
            const double pi = 3.14159265358979323846264338327950288;
            for (std::size_t k = 0; k < n; k++)
            {
                a[k] *= std::polar(1.0, pi * double(k) / double(n));
            }
            fft_positive(a);
            destination.resize(n >> 1);
            const std::uint64_t m = 2 * std::uint64_t(n);
            std::uint64_t e = 1;
            for (std::size_t j = 0; j < (n >> 1); j++)
            {
                destination[j] = a[(e - 1) >> 1];
                e = (e * 5) % m; // SEAL's generator in this fork (S/util/galois.h:169)
            }
        }
        void decode(const Plaintext &plain, std::vector<double> &destination) const
        {
            std::vector<std::complex<double>> v;
            decode(plain, v);
            destination.resize(v.size());
            for (std::size_t i = 0; i < v.size(); i++)
            {
                destination[i] = v[i].real();
            }
        }

    private:
        void check_scale(double scale, std::size_t limbs) const
        {
            // S/ckks.h:480-484: scale must be positive and fit the level's total modulus
            auto cd = ctx_.get_context_data(ctx_.parms_id_for_limbs(limbs));
            if (!(scale > 0) || static_cast<int>(std::log2(scale)) + 1 >= cd->total_coeff_modulus_bit_count())
            {
                throw std::invalid_argument("scale out of bounds");
            }
        }

        // x mod q_0..q_{l-1}  ->  centred x / scale, by Garner's mixed-radix digits
        // x = v_0 + v_1 q_0 + v_2 q_0 q_1 + ...  (0 <= v_i < q_i)
        static void compose_centered(const detail::ContextImpl &c, const std::vector<std::uint64_t> &res,
                                     std::size_t limbs, double scale, std::vector<std::complex<double>> &out)
        {
            const std::size_t n = c.n;
            const auto &q = c.primes;
            // inv[i][j] = q_j^-1 mod q_i  (j < i)
            std::vector<std::vector<std::uint64_t>> inv(limbs);
            for (std::size_t i = 0; i < limbs; i++)
            {
                inv[i].resize(i);
                for (std::size_t j = 0; j < i; j++)
                {
                    inv[i][j] = util::powmod(q[j] % q[i], q[i] - 2, q[i]);
                }
            }
            std::vector<double> qd(limbs);
            for (std::size_t i = 0; i < limbs; i++)
            {
                qd[i] = static_cast<double>(q[i]);
            }
            std::vector<std::uint64_t> v(limbs);
            for (std::size_t k = 0; k < n; k++)
            {
                for (std::size_t i = 0; i < limbs; i++)
                {
                    std::uint64_t t = res[i * n + k];
                    for (std::size_t j = 0; j < i; j++)
                    {
                        // t = (t - v_j) * q_j^-1 mod q_i
                        const std::uint64_t vj = v[j] % q[i];
                        t = util::mulmod(t >= vj ? t - vj : t + q[i] - vj, inv[i][j], q[i]);
                    }
                    v[i] = t;
                }
                // negative iff x > (Q-1)/2: compare digit by digit from the top with the digits of (Q-1)/2,
                // which are (q_i - 1)/2 (Q - 1 = sum (q_i - 1) P_i and every q_i is odd)
                bool neg = false;
                for (std::size_t i = limbs; i-- > 0;)
                {
                    const std::uint64_t half = (q[i] - 1) >> 1;
                    if (v[i] != half)
                    {
                        neg = v[i] > half;
                        break;
                    }
                }
                // |x| by Horner from the top digit; for a negative value the digits of Q - x are
                // (q_i - 1 - v_i) with one added at the bottom
                double mag = 0.0;
                for (std::size_t i = limbs; i-- > 0;)
                {
                    const double d = neg ? static_cast<double>(q[i] - 1 - v[i]) : static_cast<double>(v[i]);
                    mag = mag * qd[i] + d;
                }
                if (neg)
                {
                    mag += 1.0;
                }
                out[k] = std::complex<double>((neg ? -mag : mag) / scale, 0.0);
            }
        }

        // in-place radix-2 DFT with the positive exponent: A_t = sum_k a_k exp(+2 pi i t k / n)
        static void fft_positive(std::vector<std::complex<double>> &a)
        {
            const std::size_t n = a.size();
            for (std::size_t i = 1, j = 0; i < n; i++)
            {
                std::size_t bit = n >> 1;
                for (; j & bit; bit >>= 1)
                {
                    j ^= bit;
                }
                j ^= bit;
                if (i < j)
                {
                    std::swap(a[i], a[j]);
                }
            }
            const double pi = 3.14159265358979323846264338327950288;
            for (std::size_t len = 2; len <= n; len <<= 1)
            {
                std::vector<std::complex<double>> w(len >> 1);
                for (std::size_t k = 0; k < (len >> 1); k++)
                {
                    w[k] = std::polar(1.0, 2.0 * pi * double(k) / double(len));
                }
                for (std::size_t i = 0; i < n; i += len)
                {
                    for (std::size_t k = 0; k < (len >> 1); k++)
                    {
                        const std::complex<double> u = a[i + k], t = a[i + k + (len >> 1)] * w[k];
                        a[i + k] = u + t;
                        a[i + k + (len >> 1)] = u - t;
                    }
                }
            }
        }

        const SEALContext ctx_; // a copy, like SEAL's own classes keep (shared state inside)
    };

    // ------------------------------------------------------------------------------------------
    // Evaluator (S/evaluator.h:93-1386).  Constructor signature of the fork: Evaluator(context, encoder)
    // (S/evaluator.h:84).
    // ------------------------------------------------------------------------------------------
    class Evaluator
    {
    public:
        Evaluator(const SEALContext &ctx, const CKKSEncoder &encoder) : ctx_(ctx), encoder_(encoder)
        {}
        explicit Evaluator(const SEALContext &ctx) : ctx_(ctx), own_encoder_(new CKKSEncoder(ctx)), encoder_(*own_encoder_)
        {}

        // ---- negate / add / sub (S/evaluator.cpp:130-350) ----
        void negate_inplace(Ciphertext &a) const
        {
            valid(a, "encrypted");
            call([&](moai_context *h) { return moai_negate(h, a.data(), a.data(), 1, i32(a.size()), i32(a.coeff_modulus_size())); });
        }
        void negate(const Ciphertext &a, Ciphertext &dst) const
        {
            dst = a;
            negate_inplace(dst);
        }
        void add_inplace(Ciphertext &a, const Ciphertext &b) const
        {
            addsub(a, b, false);
        }
        void add(const Ciphertext &a, const Ciphertext &b, Ciphertext &dst) const
        {
            if (&b == &dst)
            {
                add_inplace(dst, a);
            }
            else
            {
                dst = a;
                add_inplace(dst, b);
            }
        }
        void add_many(const std::vector<Ciphertext> &v, Ciphertext &dst) const
        {
            if (v.empty())
            {
                throw std::invalid_argument("encrypteds cannot be empty");
            }
            Ciphertext acc = v[0];
            for (std::size_t i = 1; i < v.size(); i++)
            {
                add_inplace(acc, v[i]);
            }
            dst = std::move(acc);
        }
        void sub_inplace(Ciphertext &a, const Ciphertext &b) const
        {
            addsub(a, b, true);
        }
        void sub(const Ciphertext &a, const Ciphertext &b, Ciphertext &dst) const
        {
            if (&b == &dst)
            {
                // dst = a - dst
                Ciphertext t = a;
                sub_inplace(t, b);
                dst = std::move(t);
            }
            else
            {
                dst = a;
                sub_inplace(dst, b);
            }
        }

        // ---- ct x ct (S/evaluator.cpp:770-909, 1223-1282): size-2 operands, size-3 result ----
        void multiply_inplace(Ciphertext &a, const Ciphertext &b) const
        {
            valid(a, "encrypted1");
            valid(b, "encrypted2");
            if (a.parms_id() != b.parms_id())
            {
                throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
            }
            if (a.size() != 2 || b.size() != 2)
            {
                throw std::logic_error("the B200 backend multiplies size-2 ciphertexts (MOAI relinearises after every product)");
            }
            check_new_scale(a.scale() * b.scale(), a.coeff_modulus_size());
            Ciphertext r;
            r.shape(ctx_.impl(), 3, a.coeff_modulus_size());
            call([&](moai_context *h) { return moai_multiply(h, a.data(), b.data(), r.data(), 1, i32(a.coeff_modulus_size()), 0); });
            a.swap_storage(r);
            a.scale() = a.scale() * b.scale();
        }
        void multiply(const Ciphertext &a, const Ciphertext &b, Ciphertext &dst) const
        {
            if (&b == &dst)
            {
                multiply_inplace(dst, a);
            }
            else
            {
                dst = a;
                multiply_inplace(dst, b);
            }
        }
        void square_inplace(Ciphertext &a) const
        {
            valid(a, "encrypted");
            if (a.size() != 2)
            {
                throw std::logic_error("the B200 backend squares size-2 ciphertexts");
            }
            check_new_scale(a.scale() * a.scale(), a.coeff_modulus_size());
            Ciphertext r;
            r.shape(ctx_.impl(), 3, a.coeff_modulus_size());
            call([&](moai_context *h) { return moai_square(h, a.data(), r.data(), 1, i32(a.coeff_modulus_size())); });
            a.swap_storage(r);
            a.scale() = a.scale() * a.scale();
        }
        void square(const Ciphertext &a, Ciphertext &dst) const
        {
            dst = a;
            square_inplace(dst);
        }

        // ---- relinearize (S/evaluator.cpp:1345-1400) ----
        void relinearize_inplace(Ciphertext &a, const RelinKeys &rk) const
        {
            valid(a, "encrypted");
            if (a.size() == 2)
            {
                return; // destination_size == encrypted_size
            }
            if (a.size() != 3)
            {
                throw std::logic_error("the B200 backend relinearises size-3 ciphertexts");
            }
            if (!rk.has_key(2))
            {
                throw std::invalid_argument("not enough relinearization keys");
            }
            Ciphertext r;
            r.shape(ctx_.impl(), 2, a.coeff_modulus_size());
            call([&](moai_context *h) {
                return moai_relinearize(h, a.data(), r.data(), 1, i32(a.coeff_modulus_size()), rk.device_key());
            });
            a.swap_storage(r);
        }
        void relinearize(const Ciphertext &a, const RelinKeys &rk, Ciphertext &dst) const
        {
            dst = a;
            relinearize_inplace(dst, rk);
        }

        // ---- level changes (S/evaluator.cpp:1402-1720) ----
        void rescale_to_next_inplace(Ciphertext &a) const
        {
            valid(a, "encrypted");
            const std::size_t limbs = a.coeff_modulus_size();
            if (limbs < 2)
            {
                throw std::invalid_argument("end of modulus switching chain reached");
            }
            Ciphertext r;
            r.shape(ctx_.impl(), a.size(), limbs - 1);
            call([&](moai_context *h) { return moai_rescale_to_next(h, a.data(), r.data(), 1, i32(a.size()), i32(limbs)); });
            a.swap_storage(r);
            a.scale() = a.scale() / static_cast<double>(ctx_.impl()->primes[limbs - 1]);
            a.parms_id() = ctx_.parms_id_for_limbs(limbs - 1);
        }
        void rescale_to_next(const Ciphertext &a, Ciphertext &dst) const
        {
            dst = a;
            rescale_to_next_inplace(dst);
        }
        void rescale_to_inplace(Ciphertext &a, const parms_id_type &id) const
        {
            const std::size_t target = detail::limbs_of(ctx_, id, "parms_id");
            if (target > a.coeff_modulus_size())
            {
                throw std::invalid_argument("cannot switch to higher level modulus");
            }
            while (a.coeff_modulus_size() > target)
            {
                rescale_to_next_inplace(a);
            }
        }
        void mod_switch_to_inplace(Ciphertext &a, const parms_id_type &id) const
        {
            valid(a, "encrypted");
            const std::size_t target = detail::limbs_of(ctx_, id, "parms_id");
            const std::size_t limbs = a.coeff_modulus_size();
            if (target > limbs)
            {
                throw std::invalid_argument("cannot switch to higher level modulus");
            }
            if (target == limbs)
            {
                return;
            }
            Ciphertext r;
            r.shape(ctx_.impl(), a.size(), target);
            call([&](moai_context *h) {
                return moai_mod_switch_to(h, a.data(), r.data(), 1, i32(a.size()), i32(limbs), i32(target));
            });
            a.swap_storage(r);
            a.parms_id() = id;
        }
        void mod_switch_to(const Ciphertext &a, const parms_id_type &id, Ciphertext &dst) const
        {
            dst = a;
            mod_switch_to_inplace(dst, id);
        }
        void mod_switch_to_next_inplace(Ciphertext &a) const
        {
            if (a.coeff_modulus_size() < 2)
            {
                throw std::invalid_argument("end of modulus switching chain reached");
            }
            mod_switch_to_inplace(a, ctx_.parms_id_for_limbs(a.coeff_modulus_size() - 1));
        }
        void mod_switch_to_next(const Ciphertext &a, Ciphertext &dst) const
        {
            dst = a;
            mod_switch_to_next_inplace(dst);
        }
        void mod_switch_to_inplace(Plaintext &p, const parms_id_type &id) const
        {
            const std::size_t target = detail::limbs_of(ctx_, id, "parms_id");
            if (!p.is_ntt_form())
            {
                throw std::invalid_argument("plain is not in NTT form");
            }
            if (target > p.coeff_modulus_size())
            {
                throw std::invalid_argument("cannot switch to higher level modulus");
            }
            p.drop_to(ctx_, target);
        }
        void mod_switch_to_next_inplace(Plaintext &p) const
        {
            if (p.coeff_modulus_size() < 2)
            {
                throw std::invalid_argument("end of modulus switching chain reached");
            }
            p.drop_to(ctx_, p.coeff_modulus_size() - 1);
        }

        // ---- plaintext ops (S/evaluator.cpp:1938-2373) ----
        void add_plain_inplace(Ciphertext &a, const Plaintext &p) const
        {
            plain_op(a, p, 0);
        }
        void add_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &dst) const
        {
            dst = a;
            add_plain_inplace(dst, p);
        }
        void sub_plain_inplace(Ciphertext &a, const Plaintext &p) const
        {
            plain_op(a, p, 1);
        }
        void sub_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &dst) const
        {
            dst = a;
            sub_plain_inplace(dst, p);
        }
        void multiply_plain_inplace(Ciphertext &a, const Plaintext &p) const
        {
            plain_op(a, p, 2);
        }
        void multiply_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &dst) const
        {
            // one pass: read a, write dst (no copy first)
            valid(a, "encrypted");
            if (&a == &dst)
            {
                multiply_plain_inplace(dst, p);
                return;
            }
            check_plain(a, p, false);
            check_new_scale(a.scale() * p.scale(), a.coeff_modulus_size());
            dst.shape(ctx_.impl(), a.size(), a.coeff_modulus_size());
            dst.parms_id() = a.parms_id();
            dst.is_ntt_form() = a.is_ntt_form();
            dst.scale() = a.scale() * p.scale(); // S/evaluator.cpp:2370
            const std::int32_t size = i32(a.size()), limbs = i32(a.coeff_modulus_size());
            if (p.is_scalar())
            {
                call([&](moai_context *h) { return moai_multiply_scalar(h, a.data(), p.scalar_consts().data(), dst.data(), 1, size, limbs); });
            }
            else
            {
                call([&](moai_context *h) { return moai_multiply_plain(h, a.data(), p.data(), dst.data(), 1, size, limbs, 0); });
            }
        }

        // ---- rotations (S/evaluator.cpp:2563-2722), incl. SEAL's NAF fallback for missing keys ----
        void rotate_vector_inplace(Ciphertext &a, int steps, const GaloisKeys &gk) const
        {
            valid(a, "encrypted");
            if (a.size() != 2)
            {
                throw std::invalid_argument("encrypted size must be 2");
            }
            if (!gk.handle())
            {
                throw std::invalid_argument("Galois key not present");
            }
            if (steps == 0)
            {
                return;
            }
            Ciphertext r;
            r.shape(ctx_.impl(), 2, a.coeff_modulus_size());
            call([&](moai_context *h) {
                return moai_rotate_vector(h, gk.handle(), a.data(), r.data(), 1, i32(a.coeff_modulus_size()), steps);
            });
            a.swap_storage(r);
        }
        void rotate_vector(const Ciphertext &a, int steps, const GaloisKeys &gk, Ciphertext &dst) const
        {
            dst = a;
            rotate_vector_inplace(dst, steps, gk);
        }
        void complex_conjugate_inplace(Ciphertext &a, const GaloisKeys &gk) const
        {
            valid(a, "encrypted");
            if (a.size() != 2)
            {
                throw std::invalid_argument("encrypted size must be 2");
            }
            std::uint32_t elt = 0;
            call([&](moai_context *h) { return moai_galois_elt_from_step(h, 0, &elt); });
            const std::uint64_t *key = gk.device_key(elt);
            if (!key)
            {
                throw std::invalid_argument("Galois key not present");
            }
            Ciphertext r;
            r.shape(ctx_.impl(), 2, a.coeff_modulus_size());
            call([&](moai_context *h) {
                return moai_apply_galois(h, a.data(), r.data(), 1, i32(a.coeff_modulus_size()), elt, key);
            });
            a.swap_storage(r);
        }
        void complex_conjugate(const Ciphertext &a, const GaloisKeys &gk, Ciphertext &dst) const
        {
            dst = a;
            complex_conjugate_inplace(dst, gk);
        }

        // ---- NTT form (S/evaluator.cpp:2200-2334); CKKS data normally never leaves NTT form ----
        void transform_from_ntt_inplace(Ciphertext &a) const
        {
            valid(a, "encrypted");
            if (!a.is_ntt_form())
            {
                throw std::invalid_argument("encrypted_ntt is not in NTT form");
            }
            call([&](moai_context *h) { return moai_ntt_inverse(h, a.data(), 1, i32(a.size()), i32(a.coeff_modulus_size())); });
            a.is_ntt_form() = false;
        }
        void transform_to_ntt_inplace(Ciphertext &a) const
        {
            valid(a, "encrypted");
            if (a.is_ntt_form())
            {
                throw std::invalid_argument("encrypted is already in NTT form");
            }
            call([&](moai_context *h) { return moai_ntt_forward(h, a.data(), 1, i32(a.size()), i32(a.coeff_modulus_size())); });
            a.is_ntt_form() = true;
        }

        // ---- operations the fork adds (S/evaluator.cpp:395-594, S/evaluator.h:1300-1386) ----
        void add_const_inplace(Ciphertext &a, double value) const
        {
            Plaintext p;
            encoder_.encode(value, a.parms_id(), a.scale(), p); // encode at the first level + mod switch == encode here
            add_plain_inplace(a, p);
        }
        void add_const(const Ciphertext &a, double value, Ciphertext &dst) const
        {
            dst = a;
            add_const_inplace(dst, value);
        }
        void multiply_const_inplace(Ciphertext &a, double value) const
        {
            Plaintext p;
            encoder_.encode(value, a.parms_id(), a.scale(), p);
            multiply_plain_inplace(a, p);
        }
        void multiply_const(const Ciphertext &a, double value, Ciphertext &dst) const
        {
            Plaintext p;
            encoder_.encode(value, a.parms_id(), a.scale(), p);
            multiply_plain(a, p, dst);
        }
        template <typename T>
        void multiply_vector_inplace(Ciphertext &a, const std::vector<T> &value) const
        {
            Plaintext p;
            encoder_.encode(value, a.parms_id(), a.scale(), p);
            multiply_plain_inplace(a, p);
        }
        template <typename T>
        void multiply_vector(Ciphertext &a, const std::vector<T> &value, Ciphertext &dst) const
        {
            dst = a;
            multiply_vector_inplace(dst, value);
        }
        template <typename T>
        void multiply_vector_inplace_reduced_error(Ciphertext &a, const std::vector<T> &value) const
        {
            multiply_vector_inplace(a, value);
        }
        template <typename T>
        void multiply_vector_reduced_error(Ciphertext &a, const std::vector<T> &value, Ciphertext &dst) const
        {
            dst = a;
            multiply_vector_inplace_reduced_error(dst, value);
        }
        void double_inplace(Ciphertext &a) const
        {
            valid(a, "encrypted");
            call([&](moai_context *h) { return moai_add(h, a.data(), a.data(), a.data(), 1, i32(a.size()), i32(a.coeff_modulus_size())); });
        }
        // add / sub / multiply of operands at different levels and slightly different scales: the higher
        // one is multiplied by a correcting constant, rescaled and switched down (S/evaluator.cpp:420-594)
        void add_inplace_reduced_error(Ciphertext &a, const Ciphertext &b) const
        {
            reduced_error(a, b, 0, nullptr);
        }
        void add_reduced_error(const Ciphertext &a, const Ciphertext &b, Ciphertext &dst) const
        {
            Ciphertext t = a;
            reduced_error(t, b, 0, nullptr);
            dst = std::move(t);
        }
        void sub_inplace_reduced_error(Ciphertext &a, const Ciphertext &b) const
        {
            reduced_error(a, b, 1, nullptr);
        }
        void sub_reduced_error(const Ciphertext &a, const Ciphertext &b, Ciphertext &dst) const
        {
            Ciphertext t = a;
            reduced_error(t, b, 1, nullptr);
            dst = std::move(t);
        }
        void multiply_inplace_reduced_error(Ciphertext &a, const Ciphertext &b, const RelinKeys &rk) const
        {
            reduced_error(a, b, 2, &rk);
        }
        void multiply_reduced_error(const Ciphertext &a, const Ciphertext &b, const RelinKeys &rk, Ciphertext &dst) const
        {
            Ciphertext t = a;
            reduced_error(t, b, 2, &rk);
            dst = std::move(t);
        }

        const SEALContext &context() const
        {
            return ctx_;
        }

    private:
        static std::int32_t i32(std::size_t v)
        {
            return static_cast<std::int32_t>(v);
        }
        template <typename F>
        void call(F &&f) const
        {
            auto &c = ctx_.impl();
            detail::Lock lk(c->mu);
            detail::chk(f(c->cur()));
        }
        void valid(const Ciphertext &a, const char *name) const
        {
            // is_metadata_valid_for / is_buffer_valid (S/valcheck.cpp): right context, known level, has data
            if (!a.data() || a.context() != ctx_.impl() || !ctx_.get_context_data(a.parms_id()) ||
                detail::limbs_of(ctx_, a.parms_id(), name) != a.coeff_modulus_size())
            {
                throw std::invalid_argument(std::string(name) + " is not valid for encryption parameters");
            }
        }
        void check_new_scale(double scale, std::size_t limbs) const
        {
            // S/evaluator.cpp:106-116, 807-812: the product scale must fit the level's modulus
            auto cd = ctx_.get_context_data(ctx_.parms_id_for_limbs(limbs));
            if (!(scale > 0) || static_cast<int>(std::log2(scale)) >= cd->total_coeff_modulus_bit_count())
            {
                throw std::invalid_argument("scale out of bounds");
            }
        }
        void check_plain(const Ciphertext &a, const Plaintext &p, bool need_scale) const
        {
            if (!p.is_ntt_form() || (!p.is_scalar() && !p.data()))
            {
                throw std::invalid_argument("plain is not valid for encryption parameters");
            }
            if (a.parms_id() != p.parms_id())
            {
                throw std::invalid_argument("encrypted and plain parameter mismatch");
            }
            if (need_scale && !util::are_close(a.scale(), p.scale()))
            {
                throw std::invalid_argument("scale mismatch");
            }
        }
        void addsub(Ciphertext &a, const Ciphertext &b, bool sub) const
        {
            // S/evaluator.cpp:155-240 / 250-350: same level, matching scales; the result has the larger size
            valid(a, "encrypted1");
            valid(b, "encrypted2");
            if (a.parms_id() != b.parms_id())
            {
                throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
            }
            if (a.is_ntt_form() != b.is_ntt_form())
            {
                throw std::invalid_argument("NTT form mismatch");
            }
            if (!util::are_close(a.scale(), b.scale()))
            {
                throw std::invalid_argument("scale mismatch");
            }
            const std::int32_t limbs = i32(a.coeff_modulus_size());
            const std::size_t common = std::min(a.size(), b.size());
            if (a.size() >= b.size())
            {
                call([&](moai_context *h) {
                    return sub ? moai_sub(h, a.data(), b.data(), a.data(), 1, i32(common), limbs)
                               : moai_add(h, a.data(), b.data(), a.data(), 1, i32(common), limbs);
                });
                return;
            }
            // b is larger: its extra polynomials are copied (negated for sub)
            Ciphertext r;
            r.shape(ctx_.impl(), b.size(), a.coeff_modulus_size());
            const std::size_t poly = a.coeff_modulus_size() * ctx_.impl()->n;
            call([&](moai_context *h) {
                std::int32_t rc = sub ? moai_sub(h, a.data(), b.data(), r.data(), 1, i32(common), limbs)
                                      : moai_add(h, a.data(), b.data(), r.data(), 1, i32(common), limbs);
                if (rc != MOAI_OK)
                {
                    return rc;
                }
                const std::size_t extra = b.size() - common;
                if (sub)
                {
                    return moai_negate(h, b.data() + common * poly, r.data() + common * poly, 1, i32(extra), limbs);
                }
                return moai_memcpy_d2d(h, r.data() + common * poly, b.data() + common * poly,
                                       extra * poly * sizeof(std::uint64_t));
            });
            a.swap_storage(r);
        }
        void plain_op(Ciphertext &a, const Plaintext &p, int op) const
        {
            valid(a, "encrypted");
            check_plain(a, p, op != 2);
            const std::int32_t limbs = i32(a.coeff_modulus_size());
            if (op == 2)
            {
                // out of place + swap: the multiply kernels read and write through distinct pointers
                Ciphertext r;
                multiply_plain(a, p, r);
                a.swap_storage(r);
                a.scale() = r.scale();
                return;
            }
            // add_plain / sub_plain touch c0 only (S/evaluator.cpp:2022-2040); the C ABI takes the whole ciphertext
            if (p.is_scalar())
            {
                std::vector<std::uint64_t> k = p.scalar_consts();
                if (op == 1)
                {
                    for (std::size_t l = 0; l < k.size(); l++)
                    {
                        k[l] = k[l] ? ctx_.impl()->primes[l] - k[l] : 0;
                    }
                }
                call([&](moai_context *h) { return moai_add_scalar(h, a.data(), k.data(), a.data(), 1, i32(a.size()), limbs); });
            }
            else if (op == 0)
            {
                call([&](moai_context *h) { return moai_add_plain(h, a.data(), p.data(), a.data(), 1, i32(a.size()), limbs, 0); });
            }
            else
            {
                call([&](moai_context *h) { return moai_sub_plain(h, a.data(), p.data(), a.data(), 1, i32(a.size()), limbs, 0); });
            }
        }
        // the common prefix of S/evaluator.cpp:430-452, 488-510, 547-569
        Ciphertext adjusted(const Ciphertext &hi, const Ciphertext &lo) const
        {
            const double ql = static_cast<double>(ctx_.impl()->primes[hi.coeff_modulus_size() - 1]);
            const double scale_adjust = lo.scale() * ql / (hi.scale() * hi.scale());
            Ciphertext adj;
            multiply_const(hi, scale_adjust, adj);
            adj.scale() = lo.scale() * ql;
            rescale_to_next_inplace(adj);
            mod_switch_to_inplace(adj, lo.parms_id());
            return adj;
        }
        void reduced_error(Ciphertext &a, const Ciphertext &b, int op, const RelinKeys *rk) const
        {
            auto apply = [&](Ciphertext &x, const Ciphertext &y) {
                if (op == 0)
                {
                    add_inplace(x, y);
                }
                else if (op == 1)
                {
                    sub_inplace(x, y);
                }
                else
                {
                    multiply_inplace(x, y);
                    relinearize_inplace(x, *rk);
                }
            };
            const std::size_t la = a.coeff_modulus_size(), lb = b.coeff_modulus_size();
            if (la == lb)
            {
                a.scale() = b.scale();
                apply(a, b);
            }
            else if (la < lb)
            {
                Ciphertext adj = adjusted(b, a);
                a.scale() = adj.scale();
                apply(a, adj);
            }
            else
            {
                Ciphertext adj = adjusted(a, b);
                adj.scale() = b.scale();
                apply(adj, b);
                a = std::move(adj);
            }
        }

        const SEALContext ctx_; // a copy, like SEAL's own classes keep (shared state inside)
        std::unique_ptr<CKKSEncoder> own_encoder_;
        const CKKSEncoder &encoder_;
    };

    // ------------------------------------------------------------------------------------------
    // KeyGenerator (S/keygenerator.cpp:56-336, S/util/rlwe.cpp:311-430) — so that the reference's driver and test
    // programs (M/test/**), which create their keys in the same function that evaluates, compile unchanged.
    // A deployment generates keys on the client with stock SEAL; with the same PRNG seed this class produces SEAL's
    // keys bit for bit (dense ternary secrets; for the fork's sparse secrets see sample_poly_sparse_ternary).
    // Randomness on the host in SEAL's order, arithmetic on the device:
    //   pk / every key-switching digit:  seed <- prng (64 bytes);  a <- uniform(Blake2xb(seed));  e <- chi(prng);
    //                                    (c0, c1) = (-(a s + e), a);   digit i:  c0[limb i] += (P mod q_i) * key'[limb i]
    // ------------------------------------------------------------------------------------------
    class KeyGenerator
    {
    public:
        explicit KeyGenerator(const SEALContext &ctx) : ctx_(ctx)
        {
            auto &c = ctx_.impl();
            const std::size_t kl = c->key_limbs(), n = c->n;
            auto &parms = ctx_.key_context_data()->parms();
            std::vector<std::uint64_t> s(kl * n);
            auto prng = parms.random_generator()->create();
            if (!parms.secret_key_hamming_weight())
            {
                util::sample_poly_ternary(prng, c->primes, n, s.data());
            }
            else
            {
                util::sample_poly_sparse_ternary(prng, c->primes, n, parms.secret_key_hamming_weight(), s.data());
            }
            detail::DeviceBlock d;
            d.ensure(c, kl * n);
            sk_host_.resize(kl * n);
            {
                detail::Lock lk(c->mu);
                detail::chk(moai_memcpy_h2d(c->cur(), d.ptr(), s.data(), s.size() * sizeof(std::uint64_t)));
                detail::chk(moai_ntt_forward(c->cur(), d.ptr(), 1, 1, static_cast<std::int32_t>(kl)));
                detail::chk(moai_memcpy_d2h(c->cur(), sk_host_.data(), d.ptr(), s.size() * sizeof(std::uint64_t)));
            }
            sk_.upload(ctx_, sk_host_.data());
        }

        const SecretKey &secret_key() const
        {
            return sk_;
        }
        void create_public_key(PublicKey &destination) const
        {
            auto &c = ctx_.impl();
            detail::DeviceBlock pk;
            pk.ensure(c, 2 * c->key_limbs() * c->n);
            encrypt_zero_symmetric(pk.ptr());
            std::vector<std::uint64_t> host(2 * c->key_limbs() * c->n);
            {
                detail::Lock lk(c->mu);
                detail::chk(moai_memcpy_d2h(c->cur(), host.data(), pk.ptr(), host.size() * sizeof(std::uint64_t)));
            }
            destination.upload(ctx_, host.data());
        }
        void create_relin_keys(RelinKeys &destination) const
        {
            auto &c = ctx_.impl();
            const std::int32_t kl = static_cast<std::int32_t>(c->key_limbs());
            detail::DeviceBlock s2;
            s2.ensure(c, c->key_limbs() * c->n);
            {
                detail::Lock lk(c->mu); // s^2 (compute_secret_key_array, S/keygenerator.cpp:232-290)
                detail::chk(moai_multiply_plain(c->cur(), sk_.data(), sk_.data(), s2.ptr(), 1, 1, kl, 0));
            }
            std::vector<std::uint64_t> key = one_kswitch_key(s2.ptr());
            destination.upload(ctx_, key.data());
        }
        void create_galois_keys(const std::vector<std::uint32_t> &galois_elts, GaloisKeys &destination) const
        {
            auto &c = ctx_.impl();
            const std::size_t kl = c->key_limbs(), n = c->n;
            GaloisKeys fresh;
            for (std::uint32_t elt : galois_elts)
            {
                if (!(elt & 1) || elt >= 2 * n)
                {
                    throw std::invalid_argument("Galois element is not valid");
                }
                if (fresh.has_key(elt))
                {
                    continue;
                }
                // apply_galois_ntt on the secret (S/util/galois.cpp:18-55, 192-218): result[i] = s[table[i]]
                std::vector<std::uint64_t> rot(kl * n);
                for (std::size_t i = 0; i < n; i++)
                {
                    const std::uint64_t reversed = 2 * reverse_bits(i, c->log_n) + 1;
                    const std::uint64_t index_raw = ((std::uint64_t(elt) * reversed) >> 1) & (n - 1);
                    const std::size_t src = reverse_bits(static_cast<std::size_t>(index_raw), c->log_n);
                    for (std::size_t l = 0; l < kl; l++)
                    {
                        rot[l * n + i] = sk_host_[l * n + src];
                    }
                }
                detail::DeviceBlock d;
                d.ensure(c, kl * n);
                {
                    detail::Lock lk(c->mu);
                    detail::chk(moai_memcpy_h2d(c->cur(), d.ptr(), rot.data(), rot.size() * sizeof(std::uint64_t)));
                    detail::chk(moai_synchronize(c->cur()));
                }
                std::vector<std::uint64_t> key = one_kswitch_key(d.ptr());
                fresh.upload(ctx_, elt, key.data());
            }
            destination = std::move(fresh);
        }
        void create_galois_keys(const std::vector<int> &steps, GaloisKeys &destination) const
        {
            std::vector<std::uint32_t> elts;
            auto &c = ctx_.impl();
            detail::Lock lk(c->mu);
            for (int st : steps)
            {
                std::uint32_t e = 0;
                detail::chk(moai_galois_elt_from_step(c->cur(), st, &e));
                elts.push_back(e);
            }
            create_galois_keys(elts, destination);
        }
        // all power-of-two rotations and the conjugation (GaloisTool::get_elts_all, S/util/galois.cpp:106-131)
        void create_galois_keys(GaloisKeys &destination) const
        {
            const std::uint64_t m = 2 * std::uint64_t(ctx_.impl()->n);
            std::vector<std::uint32_t> elts{ static_cast<std::uint32_t>(m - 1) };
            std::uint64_t pos = 5, neg = 1;
            while ((neg * 5) % m != 1)
            {
                neg += 2; // 5^-1 mod 2N (found once; m is a power of two, the inverse is odd)
            }
            for (int i = 0; i < ctx_.impl()->log_n - 1; i++)
            {
                elts.push_back(static_cast<std::uint32_t>(pos));
                pos = (pos * pos) & (m - 1);
                elts.push_back(static_cast<std::uint32_t>(neg));
                neg = (neg * neg) & (m - 1);
            }
            create_galois_keys(elts, destination);
        }

    private:
        static std::size_t reverse_bits(std::size_t v, int bits)
        {
            std::size_t r = 0;
            for (int i = 0; i < bits; i++)
            {
                r = (r << 1) | ((v >> i) & 1);
            }
            return r;
        }
        // encrypt_zero_symmetric at the key level, NTT form, into dst [2][kl][N] (device)
        void encrypt_zero_symmetric(std::uint64_t *dst) const
        {
            auto &c = ctx_.impl();
            const std::size_t kl = c->key_limbs(), n = c->n;
            const std::int32_t l32 = static_cast<std::int32_t>(kl);
            auto bootstrap_prng = ctx_.key_context_data()->parms().random_generator()->create();
            prng_seed_type public_seed;
            bootstrap_prng->generate(prng_seed_byte_count, reinterpret_cast<seal_byte *>(public_seed.data()));
            std::vector<std::uint64_t> a(kl * n), e(kl * n);
            util::sample_poly_uniform(UniformRandomGeneratorFactory::DefaultFactory()->create(public_seed), c->primes, n, a.data());
            util::sample_poly_cbd(bootstrap_prng, c->primes, n, e.data());
            detail::DeviceBlock de;
            de.ensure(c, kl * n);
            std::uint64_t *c0 = dst, *c1 = dst + kl * n;
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_h2d(c->cur(), c1, a.data(), a.size() * sizeof(std::uint64_t)));
            detail::chk(moai_memcpy_h2d(c->cur(), de.ptr(), e.data(), e.size() * sizeof(std::uint64_t)));
            detail::chk(moai_ntt_forward(c->cur(), de.ptr(), 1, 1, l32));
            detail::chk(moai_multiply_plain(c->cur(), c1, sk_.data(), c0, 1, 1, l32, 0)); // a s
            detail::chk(moai_add(c->cur(), de.ptr(), c0, c0, 1, 1, l32));                  // + e
            detail::chk(moai_negate(c->cur(), c0, c0, 1, 1, l32));                         // -(a s + e)
            detail::chk(moai_synchronize(c->cur()));
        }
        // generate_one_kswitch_key (S/keygenerator.cpp:292-336): host copy [digits][2][kl][N]
        std::vector<std::uint64_t> one_kswitch_key(const std::uint64_t *new_key_dev) const
        {
            auto &c = ctx_.impl();
            if (!ctx_.using_keyswitching())
            {
                throw std::logic_error("keyswitching is not supported by the context");
            }
            const std::size_t kl = c->key_limbs(), n = c->n, digits = kl - 1, per = 2 * kl * n;
            const std::int32_t l32 = static_cast<std::int32_t>(kl);
            detail::DeviceBlock all, tmp;
            all.ensure(c, digits * per);
            tmp.ensure(c, kl * n);
            for (std::size_t i = 0; i < digits; i++)
            {
                std::uint64_t *dst = all.ptr() + i * per;
                encrypt_zero_symmetric(dst);
                std::vector<std::uint64_t> consts(kl, 0);
                consts[i] = c->primes[kl - 1] % c->primes[i]; // P mod q_i
                detail::Lock lk(c->mu);
                detail::chk(moai_multiply_scalar(c->cur(), new_key_dev, consts.data(), tmp.ptr(), 1, 1, l32));
                detail::chk(moai_add(c->cur(), dst, tmp.ptr(), dst, 1, 1, l32));
            }
            std::vector<std::uint64_t> host(digits * per);
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_d2h(c->cur(), host.data(), all.ptr(), host.size() * sizeof(std::uint64_t)));
            return host;
        }

        const SEALContext ctx_; // a copy, like SEAL's own classes keep (shared state inside)
        SecretKey sk_;
        std::vector<std::uint64_t> sk_host_;
    };

    // ------------------------------------------------------------------------------------------
    // Encryptor, public-key mode (S/encryptor.cpp:88-318; S/util/rlwe.cpp:224-309):
    //   (c0, c1) = round_down_one_level( pk * u + (e0, e1) ),  u <- R_3, e <- chi;  c0 += plain
    // computed one level above the target (the key level for a fresh ciphertext) and divided by that level's
    // last prime with rounding, exactly like SEAL.  u, e0, e1 come from parms.random_generator()->create() in
    // SEAL's order; NTTs, products, sums and the rounding run on the device.
    // ------------------------------------------------------------------------------------------
    class Encryptor
    {
    public:
        Encryptor(const SEALContext &ctx, const PublicKey &pk) : ctx_(ctx), pk_(pk)
        {
            if (!pk.data() || pk.context() != ctx.impl())
            {
                throw std::invalid_argument("public key is not valid for encryption parameters");
            }
        }

        void encrypt_zero(const parms_id_type &id, Ciphertext &destination) const
        {
            auto cd = ctx_.get_context_data(id);
            if (!cd || id == ctx_.key_parms_id())
            {
                throw std::invalid_argument("parms_id is not valid for encryption parameters");
            }
            auto &c = ctx_.impl();
            const std::size_t limbs = cd->parms().coeff_modulus().size(), n = c->n;
            // the level above: one more prime of the chain, or ALL primes (key level) above the first data level
            const bool has_prev = ctx_.using_keyswitching();
            const std::size_t le = has_prev ? (limbs == c->first_limbs() ? c->key_limbs() : limbs + 1) : limbs;
            // the key level is (q_0 .. q_{L-1}, p); a lower "level above" is (q_0 .. q_limbs): a prefix either way
            const std::vector<std::uint64_t> primes(c->primes.begin(), c->primes.begin() + le);
            auto prng = cd->parms().random_generator()->create();
            std::vector<std::uint64_t> u(le * n), e(2 * le * n);
            util::sample_poly_ternary(prng, primes, n, u.data());
            util::sample_poly_cbd(prng, primes, n, e.data());
            util::sample_poly_cbd(prng, primes, n, e.data() + le * n);

            detail::DeviceBlock du, de, pk, wide;
            du.ensure(c, le * n);
            de.ensure(c, 2 * le * n);
            pk.ensure(c, 2 * le * n);
            wide.ensure(c, 2 * le * n);
            const std::int32_t l32 = static_cast<std::int32_t>(le);
            detail::Lock lk(c->mu);
            detail::chk(moai_memcpy_h2d(c->cur(), du.ptr(), u.data(), u.size() * sizeof(std::uint64_t)));
            detail::chk(moai_memcpy_h2d(c->cur(), de.ptr(), e.data(), e.size() * sizeof(std::uint64_t)));
            // the first `le` limbs of each public-key polynomial (all of them at the key level)
            const std::size_t kl = c->key_limbs();
            for (int j = 0; j < 2; j++)
            {
                detail::chk(moai_memcpy_d2d(c->cur(), pk.ptr() + j * le * n, pk_.data() + j * kl * n,
                                            le * n * sizeof(std::uint64_t)));
            }
            detail::chk(moai_ntt_forward(c->cur(), du.ptr(), 1, 1, l32));
            detail::chk(moai_multiply_plain(c->cur(), pk.ptr(), du.ptr(), wide.ptr(), 1, 2, l32, 0));
            detail::chk(moai_ntt_forward(c->cur(), de.ptr(), 1, 2, l32));
            detail::chk(moai_add(c->cur(), wide.ptr(), de.ptr(), wide.ptr(), 1, 2, l32));
            destination.resize(ctx_, id, 2);
            if (le == limbs)
            {
                detail::chk(moai_memcpy_d2d(c->cur(), destination.data(), wide.ptr(), 2 * le * n * sizeof(std::uint64_t)));
            }
            else
            {
                // divide_and_round_q_last_ntt_inplace (S/util/rns.cpp:830-901) == the rescaling kernel
                detail::chk(moai_rescale_to_next(c->cur(), wide.ptr(), destination.data(), 1, 2, l32));
            }
            detail::chk(moai_synchronize(c->cur())); // the host vectors go out of scope
            destination.is_ntt_form() = true;
            destination.scale() = 1.0;
        }
        void encrypt_zero(Ciphertext &destination) const
        {
            encrypt_zero(ctx_.first_parms_id(), destination);
        }
        void encrypt(const Plaintext &plain, Ciphertext &destination) const
        {
            if (!plain.is_ntt_form())
            {
                throw std::invalid_argument("plain must be in NTT form");
            }
            if (!ctx_.get_context_data(plain.parms_id()))
            {
                throw std::invalid_argument("plain is not valid for encryption parameters");
            }
            encrypt_zero(plain.parms_id(), destination);
            auto &c = ctx_.impl();
            const std::int32_t limbs = static_cast<std::int32_t>(destination.coeff_modulus_size());
            detail::Lock lk(c->mu);
            if (plain.is_scalar())
            {
                detail::chk(moai_add_scalar(c->cur(), destination.data(), plain.scalar_consts().data(), destination.data(), 1, 2, limbs));
            }
            else
            {
                detail::chk(moai_add_plain(c->cur(), destination.data(), plain.data(), destination.data(), 1, 2, limbs, 0));
            }
            destination.scale() = plain.scale();
        }

    private:
        const SEALContext ctx_; // a copy, like SEAL's own classes keep (shared state inside)
        const PublicKey pk_;
    };

    // ------------------------------------------------------------------------------------------
    // Decryptor (S/decryptor.cpp:154-187, 299-381): c0 + c1 s (+ c2 s^2) on the device.  The reference's
    // server-side modules take the secret key only to print intermediate values (layernorm.hpp:164,
    // gelu_others.hpp:9); production deployments pass an empty SecretKey and never call decrypt.
    // ------------------------------------------------------------------------------------------
    class Decryptor
    {
    public:
        Decryptor(const SEALContext &ctx, const SecretKey &sk) : ctx_(ctx), sk_(sk)
        {}

        void decrypt(const Ciphertext &a, Plaintext &destination) const
        {
            if (!sk_.data())
            {
                throw std::invalid_argument("secret key is not set (the server side does not hold it)");
            }
            if (!a.data() || a.size() < 2 || a.size() > 3)
            {
                throw std::invalid_argument("encrypted is not valid for encryption parameters");
            }
            auto &c = ctx_.impl();
            const std::size_t limbs = a.coeff_modulus_size(), poly = limbs * c->n;
            const std::int32_t l = static_cast<std::int32_t>(limbs);
            destination.set_vector(ctx_, limbs, a.scale());
            detail::DeviceBlock tmp;
            tmp.ensure(c, poly);
            detail::Lock lk(c->mu);
            // the first `limbs` limbs of the key-level secret are the secret at this level
            if (a.size() == 3)
            {
                // (c2 s + c1) s + c0
                detail::chk(moai_multiply_plain(c->cur(), a.data(2), sk_.data(), tmp.ptr(), 1, 1, l, 0));
                detail::chk(moai_add(c->cur(), tmp.ptr(), a.data(1), tmp.ptr(), 1, 1, l));
                detail::chk(moai_multiply_plain(c->cur(), tmp.ptr(), sk_.data(), tmp.ptr(), 1, 1, l, 0));
            }
            else
            {
                detail::chk(moai_multiply_plain(c->cur(), a.data(1), sk_.data(), tmp.ptr(), 1, 1, l, 0));
            }
            detail::chk(moai_add(c->cur(), tmp.ptr(), a.data(0), destination.data(), 1, 1, l));
            detail::chk(moai_synchronize(c->cur()));
        }

    private:
        const SEALContext ctx_; // a copy, like SEAL's own classes keep (shared state inside)
        const SecretKey sk_;
    };
} // namespace sealapi
} // namespace moai_b200

#ifndef MOAI_B200_NO_SEAL_ALIAS
namespace seal = moai_b200::sealapi;
#endif

#endif // MOAI_B200_SEAL_HPP
