// Expansion of SEEDED key / ciphertext components on the device.
//
// A stock SEAL client ships the uniform half `a` of every key-switching-key digit (and of a symmetric ciphertext)
// as a 64-byte PRNG seed instead of kl x N residues (Serializable<...>, S/keygenerator.cpp:164-232,
// S/util/rlwe.cpp:137-166, 328-368): half the bytes on the wire.  The receiver regenerates
//     a <- sample_poly_uniform(Blake2xbPRNG(seed))
// i.e. (S/randomgen.cpp:176-211, S/util/blake2xb.c:33-150) the byte stream
//     buffer(0) || buffer(1) || ...,   buffer(c) = BLAKE2Xb(out = 4096 bytes, in = c as 8 LE bytes, key = seed)
//     BLAKE2Xb: h0 = BLAKE2b(key block, in; xof_length = 4096);  64-byte block i = BLAKE2b(h0; node_offset = i, ...)
// read as 64-bit words: word j*N + i fills coefficient i of limb j; a word >= the largest multiple of q_j below 2^64
// is REJECTED and replaced by the next unused word AFTER the bulk (again subject to rejection), in scan order.
// Every buffer and every child block is independent, so the bulk is one embarrassingly parallel kernel; only the
// few rejected positions (~2^-6 of the 58-bit special prime's limb, ~2^-13 of a 51-bit limb) need the sequential
// fix-up, done by one thread per polynomial over the sorted list of rejected positions.  Same residues as SEAL's
// loader bit for bit (tests/test_gpu_zz_facade.py::test_seed_expansion_on_device against the host restatement in
// include/moai_b200_seal_prng.hpp, itself pinned to the real library).
#include "ops.cuh"
#include <vector>

namespace moai
{
    namespace
    {
        constexpr int SE_TAIL_WORDS = 8192; // replacement words kept per polynomial (16 buffers)
        constexpr int SE_MAX_REJ = 8192;    // rejected positions handled per polynomial

        // message schedule; indices resolve at compile time in the fully unrolled rounds (m stays in registers)
        __device__ constexpr unsigned char se_sigma[12][16] = {
            { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 },
            { 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4 }, { 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8 },
            { 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13 }, { 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9 },
            { 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11 }, { 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10 },
            { 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5 }, { 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0 },
            { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 }
        };

        __device__ __forceinline__ u64 rotr64(u64 x, int n)
        {
            return (x >> n) | (x << (64 - n));
        }

        // BLAKE2b compression F (RFC 7693 section 3.2): h <- F(h, m, t, last)
        __device__ void b2b_compress(u64 (&h)[8], const u64 (&m)[16], u64 t, bool last)
        {
            const u64 iv[8] = { 0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL };
            u64 v[16];
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                v[i] = h[i];
                v[i + 8] = iv[i];
            }
            v[12] ^= t;
            if (last)
            {
                v[14] = ~v[14];
            }
#define MOAI_G(a, b, c, d, x, y)                                                                                       \
    v[a] = v[a] + v[b] + (x);                                                                                          \
    v[d] = rotr64(v[d] ^ v[a], 32);                                                                                    \
    v[c] = v[c] + v[d];                                                                                                \
    v[b] = rotr64(v[b] ^ v[c], 24);                                                                                    \
    v[a] = v[a] + v[b] + (y);                                                                                          \
    v[d] = rotr64(v[d] ^ v[a], 16);                                                                                    \
    v[c] = v[c] + v[d];                                                                                                \
    v[b] = rotr64(v[b] ^ v[c], 63);
#pragma unroll
            for (int r = 0; r < 12; r++)
            {
                const unsigned char *s = se_sigma[r];
                MOAI_G(0, 4, 8, 12, m[s[0]], m[s[1]])
                MOAI_G(1, 5, 9, 13, m[s[2]], m[s[3]])
                MOAI_G(2, 6, 10, 14, m[s[4]], m[s[5]])
                MOAI_G(3, 7, 11, 15, m[s[6]], m[s[7]])
                MOAI_G(0, 5, 10, 15, m[s[8]], m[s[9]])
                MOAI_G(1, 6, 11, 12, m[s[10]], m[s[11]])
                MOAI_G(2, 7, 8, 13, m[s[12]], m[s[13]])
                MOAI_G(3, 4, 9, 14, m[s[14]], m[s[15]])
            }
#undef MOAI_G
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                h[i] ^= v[i] ^ v[i + 8];
            }
        }

        __device__ __forceinline__ void b2b_init(u64 (&h)[8], u64 p0, u64 p1, u64 p2)
        {
            h[0] = 0x6a09e667f3bcc908ULL ^ p0;
            h[1] = 0xbb67ae8584caa73bULL ^ p1;
            h[2] = 0x3c6ef372fe94f82bULL ^ p2;
            h[3] = 0xa54ff53a5f1d36f1ULL;
            h[4] = 0x510e527fade682d1ULL;
            h[5] = 0x9b05688c2b3e6c1fULL;
            h[6] = 0x1f83d9abfb41bd6bULL;
            h[7] = 0x5be0cd19137e2179ULL;
        }

        // One CTA = one 4096-byte PRNG buffer (64 threads, one 64-byte child block each) of polynomial blockIdx.y.
        // Buffers [0, bulk_bufs) fill the polynomial (residue or rejection mark), the following ones its tail.
        __global__ void __launch_bounds__(64) k_seed_expand(const u64 *__restrict__ seeds, u64 *out, long long out_stride,
                                                            u64 *__restrict__ tail, unsigned *__restrict__ rej_count,
                                                            unsigned *__restrict__ rej_pos, int log_n, int limbs,
                                                            long long bulk_bufs, const LimbConst *__restrict__ lcs)
        {
            __shared__ u64 h0[8];
            const long long poly = blockIdx.y;
            const long long buf = blockIdx.x;
            const u64 *seed = seeds + poly * 8;
            if (threadIdx.x == 0)
            {
                // root: digest 64, key 64, fanout 1, depth 1, xof_length 4096
                u64 h[8], m[16];
                b2b_init(h, 64ull | (64ull << 8) | (1ull << 16) | (1ull << 24), 4096ull << 32, 0);
#pragma unroll
                for (int i = 0; i < 16; i++)
                {
                    m[i] = i < 8 ? seed[i] : 0;
                }
                b2b_compress(h, m, 128, false); // the key block
#pragma unroll
                for (int i = 0; i < 16; i++)
                {
                    m[i] = 0;
                }
                m[0] = (u64)buf;                 // the buffer counter, 8 little-endian bytes
                b2b_compress(h, m, 136, true);
#pragma unroll
                for (int i = 0; i < 8; i++)
                {
                    h0[i] = h[i];
                }
            }
            __syncthreads();
            // child block i: digest 64, key 0, fanout 0, depth 0, leaf_length 64, node_offset i, xof 4096, inner_length 64
            u64 h[8], m[16];
            const u64 i = threadIdx.x;
            b2b_init(h, 64ull | (64ull << 32), i | (4096ull << 32), 64ull << 8);
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                m[k] = k < 8 ? h0[k] : 0;
            }
            b2b_compress(h, m, 64, true);
            const long long word0 = buf * 512 + (long long)i * 8;
            const long long n = 1ll << log_n, total = (long long)limbs * n;
            if (buf >= bulk_bufs)
            {
                const long long t0 = word0 - bulk_bufs * 512;
#pragma unroll
                for (int k = 0; k < 8; k++)
                {
                    if (t0 + k < SE_TAIL_WORDS)
                    {
                        tail[poly * SE_TAIL_WORDS + t0 + k] = h[k];
                    }
                }
                return;
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
            {
                const long long pos = word0 + k;
                if (pos >= total)
                {
                    // the bulk ends inside this buffer: the rest already is tail
                    const long long t0 = pos - total;
                    if (t0 < SE_TAIL_WORDS)
                    {
                        tail[poly * SE_TAIL_WORDS + t0] = h[k];
                    }
                    continue;
                }
                const u64 q = lcs[pos >> log_n].q;
                const u64 max_multiple = 0xFFFFFFFFFFFFFFFFull - (0xFFFFFFFFFFFFFFFFull % q) - 1;
                if (h[k] >= max_multiple)
                {
                    const unsigned at = atomicAdd(rej_count + poly, 1u);
                    if (at < SE_MAX_REJ)
                    {
                        rej_pos[poly * SE_MAX_REJ + at] = (unsigned)pos;
                    }
                }
                else
                {
                    out[poly * out_stride + pos] = h[k] % q;
                }
            }
        }

        // One CTA per polynomial: sort the rejected positions, then replace them in scan order from the tail words.
        __global__ void __launch_bounds__(1024) k_seed_fixup(u64 *out, long long out_stride, const u64 *__restrict__ tail,
                                                             const unsigned *__restrict__ rej_count,
                                                             const unsigned *__restrict__ rej_pos, int log_n,
                                                             long long tail_skip, const LimbConst *__restrict__ lcs,
                                                             int *__restrict__ error)
        {
            __shared__ unsigned pos[SE_MAX_REJ];
            const long long poly = blockIdx.x;
            const unsigned cnt = rej_count[poly];
            if (cnt > SE_MAX_REJ)
            {
                if (threadIdx.x == 0)
                {
                    *error = 1;
                }
                return;
            }
            for (unsigned i = threadIdx.x; i < SE_MAX_REJ; i += blockDim.x)
            {
                pos[i] = i < cnt ? rej_pos[poly * SE_MAX_REJ + i] : 0xFFFFFFFFu;
            }
            __syncthreads();
            // bitonic sort, ascending
            for (unsigned k = 2; k <= SE_MAX_REJ; k <<= 1)
            {
                for (unsigned j = k >> 1; j > 0; j >>= 1)
                {
                    for (unsigned i = threadIdx.x; i < SE_MAX_REJ; i += blockDim.x)
                    {
                        const unsigned l = i ^ j;
                        if (l > i)
                        {
                            const unsigned a = pos[i], b = pos[l];
                            const bool up = (i & k) == 0;
                            if ((a > b) == up)
                            {
                                pos[i] = b;
                                pos[l] = a;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
            if (threadIdx.x == 0)
            {
                long long t = tail_skip; // tail words before this index belong to the bulk's last buffer... none: see host
                for (unsigned r = 0; r < cnt; r++)
                {
                    const unsigned p = pos[r];
                    const u64 q = lcs[p >> log_n].q;
                    const u64 max_multiple = 0xFFFFFFFFFFFFFFFFull - (0xFFFFFFFFFFFFFFFFull % q) - 1;
                    u64 v;
                    do
                    {
                        if (t >= SE_TAIL_WORDS)
                        {
                            *error = 1;
                            return;
                        }
                        v = tail[poly * SE_TAIL_WORDS + t++];
                    } while (v >= max_multiple);
                    out[poly * out_stride + p] = v % q;
                }
            }
        }
    } // namespace

    // out[i][limbs][n] (consecutive polynomials `out_stride` words apart) <- sample_poly_uniform(Blake2xbPRNG(seeds[i]))
    // over the first `limbs` primes of the context's key-level list.  h_seeds: count x 8 words (host).
    void expand_seeds(Context *c, const u64 *h_seeds, long long count, int limbs, u64 *d_out, long long out_stride)
    {
        MOAI_REQUIRE(count >= 1 && limbs >= 1 && limbs <= c->kl, "bad seed expansion shape");
        MOAI_REQUIRE((long long)limbs * (long long)c->n < (1ll << 32), "polynomial too large for the rejection list");
        const long long total = (long long)limbs * (long long)c->n;
        const long long bulk_bufs = total / 512; // N >= 4096: the bulk is a whole number of 512-word buffers
        MOAI_REQUIRE(total % 512 == 0, "unsupported ring degree");
        const long long tail_bufs = SE_TAIL_WORDS / 512;
        Scratch seeds((size_t)count * 8 * sizeof(u64), c->stream);
        Scratch tail((size_t)count * SE_TAIL_WORDS * sizeof(u64), c->stream);
        Scratch rej_count((size_t)count * sizeof(unsigned) + sizeof(int), c->stream);
        Scratch rej_pos((size_t)count * SE_MAX_REJ * sizeof(unsigned), c->stream);
        MOAI_CUDA_CHECK(cudaMemcpyAsync(seeds.p, h_seeds, (size_t)count * 8 * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        MOAI_CUDA_CHECK(cudaMemsetAsync(rej_count.p, 0, (size_t)count * sizeof(unsigned) + sizeof(int), c->stream));
        int *d_err = reinterpret_cast<int *>(rej_count.as<unsigned>() + count);
        {
            KernelTimer kt(c, "k_seed_expand", count);
            dim3 grid((unsigned)(bulk_bufs + tail_bufs), (unsigned)count);
            k_seed_expand<<<grid, 64, 0, c->stream>>>(seeds.as<u64>(), d_out, out_stride, tail.as<u64>(),
                                                      rej_count.as<unsigned>(), rej_pos.as<unsigned>(), c->log_n, limbs,
                                                      bulk_bufs, c->d_limb);
            c->launches += 1;
        }
        {
            KernelTimer kt(c, "k_seed_fixup", count);
            k_seed_fixup<<<(unsigned)count, 1024, 0, c->stream>>>(d_out, out_stride, tail.as<u64>(), rej_count.as<unsigned>(),
                                                                  rej_pos.as<unsigned>(), c->log_n, 0, c->d_limb, d_err);
            c->launches += 1;
        }
        MOAI_CUDA_CHECK(cudaGetLastError());
        int h_err = 0;
        MOAI_CUDA_CHECK(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (h_err)
        {
            throw StatusError{ LOGIC_ERROR, "seed expansion ran out of replacement words (more rejections than provisioned)" };
        }
    }
} // namespace moai
