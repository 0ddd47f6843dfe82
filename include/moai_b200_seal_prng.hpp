// moai_b200_seal_prng.hpp — the randomness SEAL's client-side objects consume, restated for the facade
// (include/moai_b200_seal.hpp): BLAKE2b / BLAKE2Xb, the Blake2xb PRNG with SEAL's buffering, the PRNG
// factory, UniformRandomGeneratorInfo, and the three samplers the CKKS path uses.  Host code only.
//
// Why it exists: "identical keys, randomness and inputs" (BASELINE.json north_star).  A seeded key or
// ciphertext of a stock SEAL client stores a 64-byte seed in place of its uniform polynomial
// (S/keygenerator.cpp:164-232, S/ciphertext.cpp:205-226); expanding it needs exactly SEAL's PRNG stream and
// SEAL's rejection sampling.  Encryptor::encrypt draws u <- R_3 and e_0, e_1 <- chi from the same stream
// (S/util/rlwe.cpp:224-309).  With the same seed these functions return SEAL's values bit for bit
// (tests/test_facade.py compares against the reference's real library).
//
// References: S/util/blake2b.c, S/util/blake2xb.c (RFC 7693 + the BLAKE2X counter construction),
// S/randomgen.{h,cpp}, S/randomtostd.h, S/util/rlwe.cpp:20-166, S/util/hash.h:30-37.
#ifndef MOAI_B200_SEAL_PRNG_HPP
#define MOAI_B200_SEAL_PRNG_HPP

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <random>
#include <stdexcept>
#include <vector>

namespace moai_b200
{
namespace sealapi
{
    using seal_byte = unsigned char;
    constexpr std::size_t prng_seed_uint64_count = 8;
    constexpr std::size_t prng_seed_byte_count = prng_seed_uint64_count * 8;
    using prng_seed_type = std::array<std::uint64_t, prng_seed_uint64_count>;

    enum class prng_type : std::uint8_t
    {
        unknown = 0,
        blake2xb = 1,
        shake256 = 2
    };

    namespace util
    {
        // ---- BLAKE2b (RFC 7693; S/util/blake2b.c) ----
        struct Blake2bParam
        {
            std::uint8_t digest_length = 64, key_length = 0, fanout = 1, depth = 1;
            std::uint32_t leaf_length = 0, node_offset = 0, xof_length = 0;
            std::uint8_t node_depth = 0, inner_length = 0;
        };

        class Blake2b
        {
        public:
            explicit Blake2b(const Blake2bParam &p) : outlen_(p.digest_length)
            {
                static const std::uint64_t iv[8] = { 0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                                     0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                                     0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL };
                std::uint8_t block[64] = { 0 };
                block[0] = p.digest_length;
                block[1] = p.key_length;
                block[2] = p.fanout;
                block[3] = p.depth;
                store32(block + 4, p.leaf_length);
                store32(block + 8, p.node_offset);
                store32(block + 12, p.xof_length);
                block[16] = p.node_depth;
                block[17] = p.inner_length;
                for (int i = 0; i < 8; i++)
                {
                    h_[i] = iv[i] ^ load64(block + 8 * i);
                }
            }

            void update(const void *in, std::size_t inlen)
            {
                const std::uint8_t *p = static_cast<const std::uint8_t *>(in);
                while (inlen)
                {
                    if (buflen_ == 128)
                    {
                        // a full buffer is only compressed once more input is known to follow
                        t_ += 128;
                        compress(buf_, false);
                        buflen_ = 0;
                    }
                    const std::size_t take = std::min<std::size_t>(128 - buflen_, inlen);
                    std::memcpy(buf_ + buflen_, p, take);
                    buflen_ += take;
                    p += take;
                    inlen -= take;
                }
            }

            void final(void *out)
            {
                t_ += buflen_;
                std::memset(buf_ + buflen_, 0, 128 - buflen_);
                compress(buf_, true);
                std::uint8_t full[64];
                for (int i = 0; i < 8; i++)
                {
                    store64(full + 8 * i, h_[i]);
                }
                std::memcpy(out, full, outlen_);
            }

        private:
            static std::uint64_t load64(const std::uint8_t *p)
            {
                std::uint64_t v = 0;
                for (int i = 7; i >= 0; i--)
                {
                    v = (v << 8) | p[i];
                }
                return v;
            }
            static void store64(std::uint8_t *p, std::uint64_t v)
            {
                for (int i = 0; i < 8; i++)
                {
                    p[i] = std::uint8_t(v >> (8 * i));
                }
            }
            static void store32(std::uint8_t *p, std::uint32_t v)
            {
                for (int i = 0; i < 4; i++)
                {
                    p[i] = std::uint8_t(v >> (8 * i));
                }
            }
            static std::uint64_t rotr(std::uint64_t x, int n)
            {
                return (x >> n) | (x << (64 - n));
            }
            void compress(const std::uint8_t *block, bool last)
            {
                static const std::uint64_t iv[8] = { 0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                                     0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                                     0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL };
                static const std::uint8_t sigma[12][16] = {
                    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 },
                    { 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4 }, { 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8 },
                    { 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13 }, { 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9 },
                    { 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11 }, { 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10 },
                    { 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5 }, { 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0 },
                    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 }
                };
                std::uint64_t m[16], v[16];
                for (int i = 0; i < 16; i++)
                {
                    m[i] = load64(block + 8 * i);
                }
                for (int i = 0; i < 8; i++)
                {
                    v[i] = h_[i];
                    v[i + 8] = iv[i];
                }
                v[12] ^= t_; // the high counter word stays zero below 2^64 bytes
                if (last)
                {
                    v[14] = ~v[14];
                }
#define MOAI_B2B_G(a, b, c, d, x, y)                                                                                   \
    v[a] = v[a] + v[b] + (x);                                                                                          \
    v[d] = rotr(v[d] ^ v[a], 32);                                                                                      \
    v[c] = v[c] + v[d];                                                                                                \
    v[b] = rotr(v[b] ^ v[c], 24);                                                                                      \
    v[a] = v[a] + v[b] + (y);                                                                                          \
    v[d] = rotr(v[d] ^ v[a], 16);                                                                                      \
    v[c] = v[c] + v[d];                                                                                                \
    v[b] = rotr(v[b] ^ v[c], 63);
                for (int r = 0; r < 12; r++)
                {
                    const std::uint8_t *s = sigma[r];
                    MOAI_B2B_G(0, 4, 8, 12, m[s[0]], m[s[1]])
                    MOAI_B2B_G(1, 5, 9, 13, m[s[2]], m[s[3]])
                    MOAI_B2B_G(2, 6, 10, 14, m[s[4]], m[s[5]])
                    MOAI_B2B_G(3, 7, 11, 15, m[s[6]], m[s[7]])
                    MOAI_B2B_G(0, 5, 10, 15, m[s[8]], m[s[9]])
                    MOAI_B2B_G(1, 6, 11, 12, m[s[10]], m[s[11]])
                    MOAI_B2B_G(2, 7, 8, 13, m[s[12]], m[s[13]])
                    MOAI_B2B_G(3, 4, 9, 14, m[s[14]], m[s[15]])
                }
#undef MOAI_B2B_G
                for (int i = 0; i < 8; i++)
                {
                    h_[i] ^= v[i] ^ v[i + 8];
                }
            }

            std::uint64_t h_[8];
            std::uint64_t t_ = 0;
            std::uint8_t buf_[128];
            std::size_t buflen_ = 0;
            std::size_t outlen_;
        };

        // unkeyed BLAKE2b with a `outlen`-byte digest (HashFunction::hash uses 32 bytes, S/util/hash.h:30-37)
        inline void blake2b(void *out, std::size_t outlen, const void *in, std::size_t inlen)
        {
            Blake2bParam p;
            p.digest_length = static_cast<std::uint8_t>(outlen);
            Blake2b h(p);
            h.update(in, inlen);
            h.final(out);
        }

        // BLAKE2Xb (S/util/blake2xb.c:33-150): root hash of (key block, input) with xof_length in the parameter
        // block, then output block i = BLAKE2b(root; node_offset = i, fanout = depth = 0, leaf = inner = 64)
        inline void blake2xb(void *out, std::size_t outlen, const void *in, std::size_t inlen, const void *key,
                             std::size_t keylen)
        {
            if (!out || outlen == 0 || outlen > 0xFFFFFFFFull || keylen > 64)
            {
                throw std::invalid_argument("blake2xb: bad arguments");
            }
            Blake2bParam rootp;
            rootp.digest_length = 64;
            rootp.key_length = static_cast<std::uint8_t>(keylen);
            rootp.xof_length = static_cast<std::uint32_t>(outlen);
            Blake2b root(rootp);
            if (keylen)
            {
                std::uint8_t block[128] = { 0 };
                std::memcpy(block, key, keylen);
                root.update(block, 128);
            }
            root.update(in, inlen);
            std::uint8_t h0[64];
            root.final(h0);
            std::uint8_t *dst = static_cast<std::uint8_t *>(out);
            for (std::uint32_t i = 0; outlen > 0; i++)
            {
                const std::size_t n = std::min<std::size_t>(64, outlen);
                Blake2bParam p;
                p.digest_length = static_cast<std::uint8_t>(n);
                p.key_length = 0;
                p.fanout = 0;
                p.depth = 0;
                p.leaf_length = 64;
                p.node_offset = i;
                p.xof_length = rootp.xof_length;
                p.node_depth = 0;
                p.inner_length = 64;
                Blake2b c(p);
                c.update(h0, 64);
                c.final(dst);
                dst += n;
                outlen -= n;
            }
        }
    } // namespace util

    // ------------------------------------------------------------------------------------------
    // UniformRandomGenerator with SEAL's 4096-byte buffer (S/randomgen.h:301-404, randomgen.cpp:176-211):
    // the byte stream is buffer(counter 0) || buffer(counter 1) || ..., requests of any size continue it
    // ------------------------------------------------------------------------------------------
    class UniformRandomGenerator
    {
    public:
        explicit UniformRandomGenerator(const prng_seed_type &seed) : seed_(seed), buffer_(4096), head_(4096)
        {}
        virtual ~UniformRandomGenerator() = default;

        void generate(std::size_t byte_count, seal_byte *destination)
        {
            std::lock_guard<std::mutex> lock(mutex_);
            while (byte_count)
            {
                if (head_ == buffer_.size())
                {
                    refill_buffer();
                    head_ = 0;
                }
                const std::size_t n = std::min(byte_count, buffer_.size() - head_);
                std::memcpy(destination, buffer_.data() + head_, n);
                head_ += n;
                destination += n;
                byte_count -= n;
            }
        }
        std::uint32_t generate()
        {
            std::uint32_t r;
            generate(sizeof(r), reinterpret_cast<seal_byte *>(&r));
            return r;
        }
        const prng_seed_type &seed() const noexcept
        {
            return seed_;
        }
        virtual prng_type type() const noexcept = 0;

    protected:
        virtual void refill_buffer() = 0;
        prng_seed_type seed_;
        std::vector<seal_byte> buffer_;
        std::uint64_t counter_ = 0;

    private:
        std::size_t head_;
        std::mutex mutex_;
    };

    class Blake2xbPRNG : public UniformRandomGenerator
    {
    public:
        explicit Blake2xbPRNG(const prng_seed_type &seed) : UniformRandomGenerator(seed)
        {}
        prng_type type() const noexcept override
        {
            return prng_type::blake2xb;
        }

    protected:
        void refill_buffer() override
        {
            // blake2xb(buffer, 4096, &counter, 8, seed, 64); counter++  (S/randomgen.cpp:201-211)
            util::blake2xb(buffer_.data(), buffer_.size(), &counter_, sizeof(counter_), seed_.data(), prng_seed_byte_count);
            counter_++;
        }
    };

    inline prng_seed_type random_seed()
    {
        std::random_device rd;
        prng_seed_type s;
        for (auto &w : s)
        {
            w = (std::uint64_t(rd()) << 32) | rd();
        }
        return s;
    }

    class UniformRandomGeneratorFactory
    {
    public:
        UniformRandomGeneratorFactory() : use_random_seed_(true)
        {}
        explicit UniformRandomGeneratorFactory(const prng_seed_type &default_seed)
            : default_seed_(default_seed), use_random_seed_(false)
        {}
        virtual ~UniformRandomGeneratorFactory() = default;
        // with a default seed EVERY create() returns the same stream (S/randomgen.h:420-436) — what the
        // deterministic tests rely on, on both sides
        std::shared_ptr<UniformRandomGenerator> create() const
        {
            return use_random_seed_ ? create_impl(random_seed()) : create_impl(default_seed_);
        }
        std::shared_ptr<UniformRandomGenerator> create(const prng_seed_type &seed) const
        {
            return create_impl(seed);
        }
        bool use_random_seed() const noexcept
        {
            return use_random_seed_;
        }
        static std::shared_ptr<UniformRandomGeneratorFactory> DefaultFactory();

    protected:
        virtual std::shared_ptr<UniformRandomGenerator> create_impl(const prng_seed_type &seed) const = 0;

    private:
        prng_seed_type default_seed_ = {};
        bool use_random_seed_;
    };

    class Blake2xbPRNGFactory : public UniformRandomGeneratorFactory
    {
    public:
        Blake2xbPRNGFactory() = default;
        explicit Blake2xbPRNGFactory(const prng_seed_type &default_seed) : UniformRandomGeneratorFactory(default_seed)
        {}

    protected:
        std::shared_ptr<UniformRandomGenerator> create_impl(const prng_seed_type &seed) const override
        {
            return std::make_shared<Blake2xbPRNG>(seed);
        }
    };

    inline std::shared_ptr<UniformRandomGeneratorFactory> UniformRandomGeneratorFactory::DefaultFactory()
    {
        static std::shared_ptr<UniformRandomGeneratorFactory> f{ new Blake2xbPRNGFactory() };
        return f;
    }

    // what a seeded object stores after its 0xFFFF... marker (S/randomgen.h:55-160)
    class UniformRandomGeneratorInfo
    {
    public:
        UniformRandomGeneratorInfo() = default;
        UniformRandomGeneratorInfo(prng_type type, const prng_seed_type &seed) : type_(type), seed_(seed)
        {}
        std::shared_ptr<UniformRandomGenerator> make_prng() const
        {
            if (type_ == prng_type::blake2xb)
            {
                return std::make_shared<Blake2xbPRNG>(seed_);
            }
            // the reference builds with SEAL_DEFAULT_PRNG = Blake2xb; a Shake256 stream is a different client build
            throw std::invalid_argument("unsupported prng_type in seeded object (only blake2xb)");
        }
        prng_type type() const noexcept
        {
            return type_;
        }
        const prng_seed_type &seed() const noexcept
        {
            return seed_;
        }

    private:
        prng_type type_ = prng_type::unknown;
        prng_seed_type seed_ = {};
    };

    // std::uniform_int_distribution / ClippedNormalDistribution see the PRNG through this adapter
    // (S/randomtostd.h:21-75): 32-bit outputs
    class RandomToStandardAdapter
    {
    public:
        using result_type = std::uint32_t;
        explicit RandomToStandardAdapter(std::shared_ptr<UniformRandomGenerator> g) : g_(std::move(g))
        {
            if (!g_)
            {
                throw std::invalid_argument("generator cannot be null");
            }
        }
        result_type operator()()
        {
            return g_->generate();
        }
        static constexpr result_type min() noexcept
        {
            return std::numeric_limits<result_type>::min();
        }
        static constexpr result_type max() noexcept
        {
            return std::numeric_limits<result_type>::max();
        }

    private:
        std::shared_ptr<UniformRandomGenerator> g_;
    };

    namespace util
    {
        // destination: [limbs][n] residues of ONE polynomial; primes: the level's moduli
        // u <- R_3 (S/util/rlwe.cpp:20-38): same std::uniform_int_distribution over the same 32-bit adapter
        inline void sample_poly_ternary(const std::shared_ptr<UniformRandomGenerator> &prng,
                                        const std::vector<std::uint64_t> &primes, std::size_t n, std::uint64_t *destination)
        {
            RandomToStandardAdapter engine(prng);
            std::uniform_int_distribution<std::uint64_t> dist(0, 2);
            for (std::size_t i = 0; i < n; i++)
            {
                const std::uint64_t rand = dist(engine);
                const std::uint64_t flag = static_cast<std::uint64_t>(-static_cast<std::int64_t>(rand == 0));
                for (std::size_t j = 0; j < primes.size(); j++)
                {
                    destination[j * n + i] = rand + (flag & primes[j]) - 1;
                }
            }
        }

        // the fork's sparse ternary secret of a given Hamming weight (S/util/rlwe.cpp:40-72).  The reference draws
        // positions from [0, n] INCLUSIVE and indexes destination[n] when n comes up (one element past the first
        // limb: undefined behaviour, probability weight/(n+1) per key); that draw is skipped here, every other
        // draw is consumed exactly as the reference does.
        inline void sample_poly_sparse_ternary(const std::shared_ptr<UniformRandomGenerator> &prng,
                                               const std::vector<std::uint64_t> &primes, std::size_t n,
                                               std::size_t hamming_weight, std::uint64_t *destination)
        {
            RandomToStandardAdapter engine(prng);
            std::uniform_int_distribution<std::uint64_t> dist(0, 1), dist_non_zero_position(0, n);
            std::fill(destination, destination + primes.size() * n, std::uint64_t(0));
            std::size_t current_weight = 0;
            while (current_weight < hamming_weight)
            {
                const std::size_t index = static_cast<std::size_t>(dist_non_zero_position(engine));
                if (index >= n || destination[index] != 0)
                {
                    continue;
                }
                const std::uint64_t nonzero_rand = 2 * dist(engine);
                const std::uint64_t flag = static_cast<std::uint64_t>(-static_cast<std::int64_t>(nonzero_rand == 0));
                for (std::size_t j = 0; j < primes.size(); j++)
                {
                    destination[j * n + index] = nonzero_rand + (flag & primes[j]) - 1;
                }
                current_weight++;
            }
        }

        // e <- centred binomial with sigma 3.2 (S/util/rlwe.cpp:104-135), SEAL's default noise
        inline void sample_poly_cbd(const std::shared_ptr<UniformRandomGenerator> &prng,
                                    const std::vector<std::uint64_t> &primes, std::size_t n, std::uint64_t *destination)
        {
            auto hw = [](unsigned char v) {
                int c = 0;
                for (; v; v &= static_cast<unsigned char>(v - 1))
                {
                    c++;
                }
                return c;
            };
            for (std::size_t i = 0; i < n; i++)
            {
                unsigned char x[6];
                prng->generate(6, x);
                x[2] &= 0x1F;
                x[5] &= 0x1F;
                const std::int32_t noise = hw(x[0]) + hw(x[1]) + hw(x[2]) - hw(x[3]) - hw(x[4]) - hw(x[5]);
                const std::uint64_t flag = static_cast<std::uint64_t>(-static_cast<std::int64_t>(noise < 0));
                for (std::size_t j = 0; j < primes.size(); j++)
                {
                    destination[j * n + i] = static_cast<std::uint64_t>(static_cast<std::int64_t>(noise)) + (flag & primes[j]);
                }
            }
        }

        // a <- uniform mod every prime (S/util/rlwe.cpp:137-166): the whole destination is filled from the stream
        // first; a rejected word (>= the largest multiple of q below 2^64) is replaced from the words that FOLLOW
        inline void sample_poly_uniform(const std::shared_ptr<UniformRandomGenerator> &prng,
                                        const std::vector<std::uint64_t> &primes, std::size_t n, std::uint64_t *destination)
        {
            prng->generate(primes.size() * n * sizeof(std::uint64_t), reinterpret_cast<seal_byte *>(destination));
            constexpr std::uint64_t max_random = 0xFFFFFFFFFFFFFFFFull;
            for (std::size_t j = 0; j < primes.size(); j++)
            {
                const std::uint64_t q = primes[j];
                const std::uint64_t max_multiple = max_random - (max_random % q) - 1;
                std::uint64_t *d = destination + j * n;
                for (std::size_t i = 0; i < n; i++)
                {
                    std::uint64_t rand = d[i];
                    while (rand >= max_multiple)
                    {
                        prng->generate(sizeof(rand), reinterpret_cast<seal_byte *>(&rand));
                    }
                    d[i] = rand % q;
                }
            }
        }
    } // namespace util
} // namespace sealapi
} // namespace moai_b200

#endif // MOAI_B200_SEAL_PRNG_HPP
