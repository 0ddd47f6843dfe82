// Grouped-digit key switching (design and derivation: ksgroup.hpp).
#include "ksgroup.hpp"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace moai
{
    namespace
    {
        typedef unsigned __int128 u128h;

        u64 h_mulmod(u64 a, u64 b, u64 q)
        {
            return (u64)((u128h)a * b % q);
        }
        u64 h_powmod(u64 a, u64 e, u64 q)
        {
            u64 r = 1 % q;
            a %= q;
            while (e)
            {
                if (e & 1)
                {
                    r = h_mulmod(r, a, q);
                }
                a = h_mulmod(a, a, q);
                e >>= 1;
            }
            return r;
        }
        u64 h_invmod(u64 a, u64 q) // q prime
        {
            return h_powmod(a % q, q - 2, q);
        }
        Twiddle h_shoup(u64 w, u64 q)
        {
            Twiddle t;
            t.w = w;
            t.wq = (u64)(((u128h)w << 64) / q);
            return t;
        }
        double h_centred(u64 v, u64 q)
        {
            return v > q / 2 ? -(double)(q - v) : (double)v;
        }

        // consecutive groups of the data primes 0 .. L-k-1, each with prod(G) <= p * prod(E) (bit-length sums)
        std::vector<int> group_starts(const Context *c, int k)
        {
            const int L = c->kl - 1;
            double cap = std::log2((double)c->q[c->kl - 1]);
            for (int i = L - k; i < L; i++)
            {
                cap += std::log2((double)c->q[i]);
            }
            std::vector<int> st{ 0 };
            double bits = 0;
            int cnt = 0;
            for (int J = 0; J < L - k; J++)
            {
                const double b = std::log2((double)c->q[J]);
                if (cnt > 0 && (bits + b > cap || cnt >= CONV_MAX))
                {
                    st.push_back(J);
                    bits = 0;
                    cnt = 0;
                }
                bits += b;
                cnt++;
            }
            st.push_back(L - k);
            return st;
        }

        struct KsgTables
        {
            int k = 0, limbs = 0, rns = 0, digits = 0;
            KsShape shape;
            ConvTab dec, md;
            const Twiddle *d_yconst = nullptr; // [limbs]  prod(E) (Q_g / q_J)^-1 mod q_J
            const Twiddle *d_zconst = nullptr; // [k + 1]  (P' / p_i)^-1 mod p_i
            const Twiddle *d_pinv = nullptr;   // [limbs]  P'^-1 mod q_j
            // merged mod-down + rescale (divide by P'' = P' q_{limbs-1} in one go; limbs >= 2)
            ConvTab mdr;                         // sources q_{limbs-1}, E, p  ->  targets q_0 .. q_{limbs-2}
            const NttScale *d_zscale_r = nullptr; // [k + 2]  (P'' / p_i)^-1 mod p_i folded into the inverse transform
            const Twiddle *d_pinv_r = nullptr;   // [limbs - 1]  P''^-1 mod q_j
            const Twiddle *d_qlinv = nullptr;    // [limbs - 1]  q_{limbs-1}^-1 mod q_j
            Twiddle pmod_last;                   // P' mod q_{limbs-1}
            const Twiddle *d_qe = nullptr;     // [limbs]  prod(E) mod q_j
            const NttScale *d_yscale = nullptr; // [limbs]  yconst folded into the inverse transform's N^-1
            const NttScale *d_zscale = nullptr; // [k + 1]  zconst likewise
            const int *d_own = nullptr;        // [rns]    digit whose group holds target I (FP64-path data primes), else -1
            int n_own = 0;
            const int *d_ids = nullptr;        // [rns]
            std::vector<u64> h_pmod;           // [limbs]  P' mod q_j
            std::vector<int> h_ids;            // [rns]
            void *blob = nullptr;
        };

        struct Blob
        {
            std::vector<unsigned char> bytes;
            size_t put(const void *p, size_t n)
            {
                const size_t off = (bytes.size() + 15) & ~(size_t)15;
                bytes.resize(off + n);
                std::memcpy(bytes.data() + off, p, n);
                return off;
            }
            template <class T>
            size_t put(const std::vector<T> &v)
            {
                return put(v.data(), v.size() * sizeof(T));
            }
        };

        // conversion tables of ONE source set (`src` prime indices, grouped by s0 / cnt) into `tgt` prime indices
        struct ConvOffsets
        {
            size_t s0, cnt, invq, wide, B, B26, Bd, B26d, negQ, negQd, BT, c32d, fpsrc, srcq;
            int src_limbs;
            bool mma, fp;
        };
        // the conversion's tensor-core form (ConvTab::BT) needs whole warps in the pass-A CTA (N >= 2^13) and a free
        // top byte in every group's first source.  OPT-IN (MOAI_CONV_MMA=1, read when a level's tables are first built):
        // it takes the FP64 pipe from 61-68 % to 18-22 % busy but is issue- / latency-bound in all three variants
        // measured (relinearize + rescale, 64 ciphertexts at 28 limbs: FP64 products 7.14 ms; fragments straight from
        // global memory 2x slower; cp.async staging 8.22 ms; bulk copies + mbarriers 10.06 ms — profiles/conv_mma_r2.md)
        bool conv_mma_enabled(const Context *c)
        {
            const char *e = std::getenv("MOAI_CONV_MMA");
            return e && e[0] == '1' && c->log_n >= 13;
        }
        ConvOffsets build_conv(const Context *c, Blob &bl, const std::vector<int> &src, const std::vector<int> &s0,
                               const std::vector<int> &cnt, const std::vector<int> &tgt, bool want_mma = true)
        {
            const int digits = (int)s0.size(), rns = (int)tgt.size();
            std::vector<double> invq(src.size());
            std::vector<unsigned char> wide(src.size());
            for (size_t j = 0; j < src.size(); j++)
            {
                invq[j] = 1.0 / (double)c->q[src[j]];
                wide[j] = (c->q[src[j]] >> 52) != 0;
            }
            std::vector<u64> B((size_t)digits * rns * CONV_MAX, 0), B26(B.size(), 0), negQ((size_t)digits * rns, 0);
            std::vector<double> Bd(B.size(), 0.0), B26d(B.size(), 0.0), negQd(negQ.size(), 0.0);
            for (int g = 0; g < digits; g++)
            {
                for (int I = 0; I < rns; I++)
                {
                    const u64 m = c->q[tgt[I]];
                    u64 Qg = 1 % m;
                    for (int j = 0; j < cnt[g]; j++)
                    {
                        Qg = h_mulmod(Qg, c->q[src[s0[g] + j]] % m, m);
                        u64 b = 1 % m; // (Q_g / q_j) mod m as a product (no big integers needed)
                        for (int j2 = 0; j2 < cnt[g]; j2++)
                        {
                            if (j2 != j)
                            {
                                b = h_mulmod(b, c->q[src[s0[g] + j2]] % m, m);
                            }
                        }
                        const size_t at = ((size_t)g * rns + I) * CONV_MAX + j;
                        B[at] = b;
                        B26[at] = h_mulmod(b, ((u64)1 << 26) % m, m);
                        Bd[at] = h_centred(B[at], m);
                        B26d[at] = h_centred(B26[at], m);
                    }
                    negQ[(size_t)g * rns + I] = Qg ? m - Qg : 0;
                    negQd[(size_t)g * rns + I] = h_centred(negQ[(size_t)g * rns + I], m);
                }
            }
            // B fragments of mma.sync m16n8k32 (u8): k = 8 (source within the k-step) + byte of the residue, n = byte of the
            // constant C(j, a, I) = 2^(8a) (Q_g / q_j) mod m_I; the v slot (first source, byte 7) holds (-Q_g) mod m_I
            bool mma = want_mma && conv_mma_enabled(c);
            for (int g = 0; g < digits; g++)
            {
                mma = mma && (c->q[src[s0[g]]] >> 56) == 0;
            }
            std::vector<uint32_t> BT(mma ? (size_t)digits * rns * CONV_KSTEPS * 64 : 0, 0u);
            std::vector<double> c32d(rns, 0.0);
            for (int I = 0; I < rns; I++)
            {
                c32d[I] = h_centred(((u64)1 << 32) % c->q[tgt[I]], c->q[tgt[I]]);
            }
            for (int g = 0; mma && g < digits; g++)
            {
                for (int I = 0; I < rns; I++)
                {
                    const u64 m = c->q[tgt[I]];
                    u64 cst_tab[CONV_MAX][8];
                    for (int j = 0; j < cnt[g]; j++)
                    {
                        u64 v = B[((size_t)g * rns + I) * CONV_MAX + j];
                        for (int ab = 0; ab < 8; ab++)
                        {
                            cst_tab[j][ab] = v;
                            v = h_mulmod(v, 256 % m, m);
                        }
                    }
                    cst_tab[0][7] = negQ[(size_t)g * rns + I];
                    for (int sI = 0; sI < CONV_KSTEPS; sI++)
                    {
                        for (int lane = 0; lane < 32; lane++)
                        {
                            const int g8 = lane >> 2, c4 = lane & 3;
                            for (int r = 0; r < 2; r++)
                            {
                                uint32_t word = 0;
                                for (int e = 0; e < 4; e++)
                                {
                                    // lane c4 of a quad holds source c4 of the k-step: bytes 0-3 in register 0, 4-7 in 1
                                    const int j = 4 * sI + c4, ab = 4 * r + e;
                                    const u64 cst = j < cnt[g] ? cst_tab[j][ab] : 0;
                                    word |= (uint32_t)((cst >> (8 * g8)) & 0xFF) << (8 * e);
                                }
                                BT[((((size_t)g * rns + I) * CONV_KSTEPS + sI) * 32 + lane) * 2 + r] = word;
                            }
                        }
                    }
                }
            }
            // FP64-path sources as centred doubles with precomputed quotients: OPT-IN (MOAI_CONV_FPSRC=1).  It removes 2 of
            // the 9 FP64 operations per (source, target), but the conversion pass A only gets 4 % faster (7.19 -> 6.90 ms
            // at 28 limbs, 64 ciphertexts) and the quotient pre-pass costs 0.34 ms: no net gain (relin_rescale 16.03 vs
            // 16.05 ms; 33 limbs 27.5 vs 27.9; 22 limbs 11.14 vs 11.21)
            const char *efp = std::getenv("MOAI_CONV_FPSRC");
            const bool fp = !mma && efp && efp[0] == '1';
            std::vector<unsigned char> fpsrc(src.size(), 0);
            std::vector<u64> srcq(src.size());
            for (size_t j = 0; j < src.size(); j++)
            {
                fpsrc[j] = fp && c->h_limb[src[j]].fp_class != 0;
                srcq[j] = c->q[src[j]];
            }
            ConvOffsets o;
            o.fp = fp;
            o.fpsrc = bl.put(fpsrc);
            o.srcq = bl.put(srcq);
            o.mma = mma;
            o.BT = bl.put(BT);
            o.c32d = bl.put(c32d);
            o.s0 = bl.put(s0);
            o.cnt = bl.put(cnt);
            o.invq = bl.put(invq);
            o.wide = bl.put(wide);
            o.B = bl.put(B);
            o.B26 = bl.put(B26);
            o.Bd = bl.put(Bd);
            o.B26d = bl.put(B26d);
            o.negQ = bl.put(negQ);
            o.negQd = bl.put(negQd);
            o.src_limbs = (int)src.size();
            return o;
        }
        ConvTab bind_conv(const ConvOffsets &o, const unsigned char *base)
        {
            ConvTab t;
            t.s0 = reinterpret_cast<const int *>(base + o.s0);
            t.cnt = reinterpret_cast<const int *>(base + o.cnt);
            t.invq = reinterpret_cast<const double *>(base + o.invq);
            t.wide = base + o.wide;
            t.B = reinterpret_cast<const u64 *>(base + o.B);
            t.B26 = reinterpret_cast<const u64 *>(base + o.B26);
            t.Bd = reinterpret_cast<const double *>(base + o.Bd);
            t.B26d = reinterpret_cast<const double *>(base + o.B26d);
            t.negQ = reinterpret_cast<const u64 *>(base + o.negQ);
            t.negQd = reinterpret_cast<const double *>(base + o.negQd);
            t.src_limbs = o.src_limbs;
            t.BT = o.mma ? reinterpret_cast<const uint32_t *>(base + o.BT) : nullptr;
            t.c32d = reinterpret_cast<const double *>(base + o.c32d);
            t.fpsrc = o.fp ? base + o.fpsrc : nullptr;
            t.srcq = reinterpret_cast<const u64 *>(base + o.srcq);
            return t;
        }

        const KsgTables &tables(Context *lane, int k, int limbs)
        {
            Context *c = lane->root(); // the tables are immutable: every lane of a context shares the root's cache
            std::lock_guard<std::mutex> lk(c->ksg_mu);
            auto it = c->ksg_cache.find({ k, limbs });
            if (it != c->ksg_cache.end())
            {
                return *static_cast<KsgTables *>(it->second);
            }
            const int L = c->kl - 1;
            MOAI_REQUIRE(k >= 0 && limbs >= 1 && limbs + k <= L, "no spare primes for grouped digits at this level");
            KsgTables *t = new KsgTables();
            t->k = k;
            t->limbs = limbs;
            t->rns = limbs + k + 1;
            // targets: own data primes, the k borrowed top primes, the special prime
            std::vector<int> ids;
            for (int i = 0; i < limbs; i++)
            {
                ids.push_back(i);
            }
            for (int i = L - k; i < L; i++)
            {
                ids.push_back(i);
            }
            ids.push_back(L);
            // digits: the key's fixed groups cut at `limbs`
            std::vector<int> s0, cnt;
            if (k == 0) // SEAL's digits: one per prime, extended by a plain reduction (no conversion tables)
            {
                for (int i = 0; i < limbs; i++)
                {
                    s0.push_back(i);
                    cnt.push_back(1);
                }
            }
            else
            {
                const std::vector<int> st = group_starts(c, k);
                for (size_t g = 0; g + 1 < st.size() && st[g] < limbs; g++)
                {
                    s0.push_back(st[g]);
                    cnt.push_back(std::min(st[g + 1], limbs) - st[g]);
                }
            }
            t->digits = (int)s0.size();
            std::vector<int> src_dec(limbs);
            for (int i = 0; i < limbs; i++)
            {
                src_dec[i] = i;
            }
            Blob bl;
            const size_t o_ids = bl.put(ids);
            const ConvOffsets o_dec = build_conv(c, bl, src_dec, s0, cnt, ids, /*want_mma=*/k > 0);
            // y_J = c_J * prod(E) * (Q_g / q_J)^-1 mod q_J
            std::vector<Twiddle> yconst(limbs);
            for (int g = 0; g < t->digits; g++)
            {
                for (int j = 0; j < cnt[g]; j++)
                {
                    const int J = s0[g] + j;
                    const u64 q = c->q[J];
                    u64 v = 1;
                    for (int j2 = 0; j2 < cnt[g]; j2++)
                    {
                        if (j2 != j)
                        {
                            v = h_mulmod(v, c->q[s0[g] + j2] % q, q);
                        }
                    }
                    v = h_invmod(v, q);
                    for (int i = L - k; i < L; i++)
                    {
                        v = h_mulmod(v, c->q[i] % q, q);
                    }
                    yconst[J] = h_shoup(v, q);
                }
            }
            const size_t o_y = bl.put(yconst);
            // mod-down by P' = prod(E) * p: one group of k + 1 source limbs into the data primes
            std::vector<int> src_md(ids.begin() + limbs, ids.end()), tgt_md(ids.begin(), ids.begin() + limbs);
            const ConvOffsets o_md = build_conv(c, bl, src_md, std::vector<int>{ 0 }, std::vector<int>{ k + 1 }, tgt_md);
            std::vector<Twiddle> zconst(k + 1), pinv(limbs);
            for (int i = 0; i <= k; i++)
            {
                const u64 q = c->q[src_md[i]];
                u64 v = 1;
                for (int i2 = 0; i2 <= k; i2++)
                {
                    if (i2 != i)
                    {
                        v = h_mulmod(v, c->q[src_md[i2]] % q, q);
                    }
                }
                zconst[i] = h_shoup(h_invmod(v, q), q);
            }
            for (int j = 0; j < limbs; j++)
            {
                const u64 q = c->q[j];
                u64 v = 1;
                for (int i = 0; i <= k; i++)
                {
                    v = h_mulmod(v, c->q[src_md[i]] % q, q);
                }
                pinv[j] = h_shoup(h_invmod(v, q), q);
            }
            // P' mod q_j: the factor that lifts an un-switched polynomial into the key-switch basis
            t->h_pmod.resize(limbs);
            for (int j = 0; j < limbs; j++)
            {
                const u64 q = c->q[j];
                u64 v = 1;
                for (int i = 0; i <= k; i++)
                {
                    v = h_mulmod(v, c->q[src_md[i]] % q, q);
                }
                t->h_pmod[j] = v;
            }
            t->h_ids = ids;
            std::vector<Twiddle> qe(limbs);
            for (int j = 0; j < limbs; j++)
            {
                const u64 q = c->q[j];
                u64 v = 1;
                for (int i = L - k; i < L; i++)
                {
                    v = h_mulmod(v, c->q[i] % q, q);
                }
                qe[j] = h_shoup(v, q);
            }
            std::vector<int> own(t->rns, -1);
            for (int g = 0; g < t->digits; g++)
            {
                for (int j = 0; j < cnt[g]; j++)
                {
                    if (c->h_limb[s0[g] + j].fp_class != 0)
                    {
                        own[s0[g] + j] = g;
                    }
                }
            }
            for (int v : own)
            {
                t->n_own += v >= 0;
            }
            std::vector<NttScale> yscale(limbs), zscale(k + 1);
            for (int j = 0; j < limbs; j++)
            {
                yscale[j] = ntt_scale_make(c, j, yconst[j].w);
            }
            for (int i = 0; i <= k; i++)
            {
                zscale[i] = ntt_scale_make(c, src_md[i], zconst[i].w);
            }
            const size_t o_ys = bl.put(yscale), o_zs = bl.put(zscale);
            // merged mod-down + rescale: one more source limb (the level's last data prime), one target fewer
            ConvOffsets o_mdr{};
            size_t o_zsr = 0, o_pr = 0, o_ql = 0;
            if (limbs >= 2)
            {
                std::vector<int> src_r(ids.begin() + limbs - 1, ids.end()), tgt_r(ids.begin(), ids.begin() + limbs - 1);
                o_mdr = build_conv(c, bl, src_r, std::vector<int>{ 0 }, std::vector<int>{ k + 2 }, tgt_r);
                std::vector<NttScale> zscale_r(k + 2);
                for (int i = 0; i < k + 2; i++)
                {
                    const u64 q = c->q[src_r[i]];
                    u64 v = 1;
                    for (int i2 = 0; i2 < k + 2; i2++)
                    {
                        if (i2 != i)
                        {
                            v = h_mulmod(v, c->q[src_r[i2]] % q, q);
                        }
                    }
                    zscale_r[i] = ntt_scale_make(c, src_r[i], h_invmod(v, q));
                }
                std::vector<Twiddle> pinv_r(limbs - 1), qlinv(limbs - 1);
                for (int j = 0; j < limbs - 1; j++)
                {
                    const u64 q = c->q[j];
                    u64 v = 1;
                    for (int i = 0; i < k + 2; i++)
                    {
                        v = h_mulmod(v, c->q[src_r[i]] % q, q);
                    }
                    pinv_r[j] = h_shoup(h_invmod(v, q), q);
                    qlinv[j] = h_shoup(h_invmod(c->q[limbs - 1] % q, q), q);
                }
                o_zsr = bl.put(zscale_r);
                o_pr = bl.put(pinv_r);
                o_ql = bl.put(qlinv);
                t->pmod_last = h_shoup(t->h_pmod[limbs - 1], c->q[limbs - 1]);
            }
            const size_t o_qe = bl.put(qe), o_own = bl.put(own);
            const size_t o_z = bl.put(zconst), o_p = bl.put(pinv);
            MOAI_CUDA_CHECK(cudaMalloc(&t->blob, bl.bytes.size()));
            MOAI_CUDA_CHECK(cudaMemcpy(t->blob, bl.bytes.data(), bl.bytes.size(), cudaMemcpyHostToDevice));
            const unsigned char *base = static_cast<const unsigned char *>(t->blob);
            t->d_ids = reinterpret_cast<const int *>(base + o_ids);
            t->dec = bind_conv(o_dec, base);
            t->md = bind_conv(o_md, base);
            t->d_yconst = reinterpret_cast<const Twiddle *>(base + o_y);
            t->d_zconst = reinterpret_cast<const Twiddle *>(base + o_z);
            t->d_pinv = reinterpret_cast<const Twiddle *>(base + o_p);
            t->d_qe = reinterpret_cast<const Twiddle *>(base + o_qe);
            t->d_yscale = reinterpret_cast<const NttScale *>(base + o_ys);
            t->d_zscale = reinterpret_cast<const NttScale *>(base + o_zs);
            if (limbs >= 2)
            {
                t->mdr = bind_conv(o_mdr, base);
                t->d_zscale_r = reinterpret_cast<const NttScale *>(base + o_zsr);
                t->d_pinv_r = reinterpret_cast<const Twiddle *>(base + o_pr);
                t->d_qlinv = reinterpret_cast<const Twiddle *>(base + o_ql);
            }
            t->d_own = reinterpret_cast<const int *>(base + o_own);
            t->shape.digits = t->digits;
            t->shape.rns = t->rns;
            t->shape.n_data = limbs;
            t->shape.ids = t->d_ids;
            c->ksg_cache[{ k, limbs }] = t;
            return *t;
        }

        // out[b][l][n] = src[b * stride + l * n ...] * consts[l]  (mod q_l): prod(E) * c1 in NTT form
        __global__ void k_scale_from(const ulonglong2 *__restrict__ src, long long stride2, ulonglong2 *__restrict__ out,
                                     long long total2, int log_n2, int limbs, const Twiddle *__restrict__ consts,
                                     const LimbConst *__restrict__ lcs)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total2)
            {
                return;
            }
            const long long per = (long long)limbs << log_n2;
            const long long b = i / per, rest = i % per;
            const int l = (int)(rest >> log_n2);
            const u64 q = lcs[l].q;
            const Twiddle w = consts[l];
            ulonglong2 v = src[b * stride2 + rest];
            v.x = mul_shoup(v.x, w.w, w.wq, q);
            v.y = mul_shoup(v.y, w.w, w.wq, q);
            out[i] = v;
        }

        // out[G][k][I] = sum_{J in G} in[J][k][prime(I)]   (16-byte lanes; <= CONV_MAX terms below 2^61 each)
        struct KeySumArgs
        {
            const ulonglong2 *src[CONV_MAX];
            int cnt;
        };
        __global__ void k_key_group_sum(KeySumArgs a, ulonglong2 *out, long long n2, u64 q)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n2)
            {
                return;
            }
            ulonglong2 s = a.src[0][i];
            for (int j = 1; j < a.cnt; j++)
            {
                const ulonglong2 v = a.src[j][i];
                s.x = addmod(s.x, v.x, q);
                s.y = addmod(s.y, v.y, q);
            }
            out[i] = s;
        }

        // merged mod-down + rescale: acc[P][rns][n] (limb `limb`) += (P' mod q_limb) * addend[(P/2)*group + P%2][limbs][n] (limb `limb`)
        __global__ void k_add_scaled_limb(ulonglong2 *acc, const ulonglong2 *__restrict__ addend, long long total2, int log_n2,
                                          int rns, int limbs, int limb, int addend_group, int even_only, Twiddle w, u64 q)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [P][n/2]
            if (i >= total2)
            {
                return;
            }
            const long long P = i >> log_n2, within = i & (((long long)1 << log_n2) - 1);
            if (even_only && (P & 1))
            {
                return;
            }
            ulonglong2 *a = acc + (((P * rns + limb) << log_n2) + within);
            const ulonglong2 z = addend[((((P >> 1) * addend_group + (P & 1)) * limbs + limb) << log_n2) + within];
            ulonglong2 v = *a;
            v.x = addmod(v.x, mul_shoup(z.x, w.w, w.wq, q), q);
            v.y = addmod(v.y, mul_shoup(z.y, w.w, w.wq, q), q);
            *a = v;
        }

        size_t ksg_ext_budget_bytes()
        {
            return ks_ext_budget();
        }
    } // namespace

    void ksg_release(Context *c)
    {
        for (auto &kv : c->ksg_cache)
        {
            KsgTables *t = static_cast<KsgTables *>(kv.second);
            cudaFree(t->blob);
            delete t;
        }
        c->ksg_cache.clear();
    }

    int ksg_max_limbs(Context *c, int k)
    {
        return c->kl - 1 - k;
    }

    int ksg_digits(Context *c, int k, int lmax)
    {
        MOAI_REQUIRE(k >= 1 && lmax >= 1 && lmax + k <= c->kl - 1, "no spare primes for grouped digits at this level");
        const std::vector<int> st = group_starts(c, k);
        int d = 0;
        for (size_t g = 0; g + 1 < st.size() && st[g] < lmax; g++)
        {
            d++;
        }
        return d;
    }

    size_t ksg_key_words(Context *c, int k, int lmax)
    {
        return (size_t)ksg_digits(c, k, lmax) * 2 * ksg_key_kl(k, lmax) * c->n;
    }

    // Cost model in microseconds per ciphertext at N = 65536 (limb-transform 0.5, inverse 0.6, one source limb of a
    // base conversion into one target limb 0.035, one (digit, target) pair of the evk inner product 0.06, one
    // element-wise pass over a limb 0.08); only the ORDER of the candidates matters.
    double ksg_cost(Context *c, int limbs, int k)
    {
        const double NTT = 0.5, INTT = 0.6, CONV = 0.035, MAC = 0.06, EW = 0.08;
        if (k == 0)
        {
            return limbs * INTT + (double)limbs * (limbs + 1) * (NTT + MAC) + 2 * (INTT + limbs * NTT + limbs * EW);
        }
        const int digits = ksg_digits(c, k, limbs), m = limbs + k + 1;
        const double dec = limbs * (INTT + EW) + (double)digits * m * NTT + (double)limbs * m * CONV;
        const double mac = (double)digits * m * MAC;
        const double md = 2 * ((k + 1) * (INTT + EW) + limbs * NTT + (double)(k + 1) * limbs * CONV + limbs * EW);
        return dec + mac + md;
    }

    int ksg_best_k(Context *c, int limbs)
    {
        int best = 0;
        double bc = ksg_cost(c, limbs, 0);
        for (int k = 1; limbs + k <= c->kl - 1 && k + 1 <= CONV_MAX; k++)
        {
            const double v = ksg_cost(c, limbs, k);
            if (v < bc)
            {
                bc = v;
                best = k;
            }
        }
        return best;
    }

    void ksg_key_prepare(Context *c, const u64 *in, uint32_t elt, int k, int lmax, bool pre_permute, u64 *out)
    {
        const int L = c->kl - 1;
        MOAI_REQUIRE(k >= 1 && lmax >= 1 && lmax + k <= L, "no spare primes for grouped digits at this level");
        MOAI_REQUIRE(k + 1 <= CONV_MAX, "too many extra primes");
        MOAI_REQUIRE(in != out, "key preparation is out of place");
        const size_t n = c->n;
        const int okl = ksg_key_kl(k, lmax);
        const std::vector<int> st = group_starts(c, k);
        const int digits = ksg_digits(c, k, lmax);
        uint32_t inv = 1;
        if (pre_permute)
        {
            const uint64_t m = 2 * (uint64_t)n;
            uint64_t x = elt;
            for (int i = 0; i < 6; i++)
            {
                x = (x * (2 + m * 4 - (uint64_t)elt * x % m)) % m;
            }
            MOAI_REQUIRE((uint64_t)elt * x % m == 1, "Galois element is not invertible");
            inv = (uint32_t)x;
        }
        Scratch tmp(n * sizeof(u64), c->stream);
        const long long n2 = (long long)(n / 2);
        for (int g = 0; g < digits; g++)
        {
            // the WHOLE group of the key (also the primes at or above lmax: their F_J vanish on the basis in use)
            const int j0 = st[g], j1 = st[g + 1];
            for (int kk = 0; kk < 2; kk++)
            {
                for (int ol = 0; ol < okl; ol++)
                {
                    const int prime = ol < lmax ? ol : (ol < lmax + k ? L - k + (ol - lmax) : L);
                    KeySumArgs a;
                    a.cnt = j1 - j0;
                    for (int j = 0; j < a.cnt; j++)
                    {
                        a.src[j] = reinterpret_cast<const ulonglong2 *>(in + (((size_t)(j0 + j) * 2 + kk) * c->kl + prime) * n);
                    }
                    u64 *dst = out + (((size_t)g * 2 + kk) * okl + ol) * n;
                    u64 *sum = pre_permute ? tmp.as<u64>() : dst;
                    {
                        KernelTimer kt(c, "k_key_group_sum", 1);
                        k_key_group_sum<<<(unsigned)((n2 + 255) / 256), 256, 0, c->stream>>>(
                            a, reinterpret_cast<ulonglong2 *>(sum), n2, c->q[prime]);
                        c->launches += 1;
                    }
                    if (pre_permute)
                    {
                        apply_galois_ntt(c, sum, dst, 1, inv);
                    }
                }
            }
        }
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    size_t ksg_ext_bytes_per_ct(Context *c, int limbs, int k)
    {
        return (size_t)(limbs + k + 1) * ksg_digits(c, k, limbs) * c->n * sizeof(u64);
    }

    // prod(E) * target in NTT form, [batch][limbs][n]: the residues of every digit modulo its OWN primes
    static void ksg_direct(Context *c, const KsgTables &t, const u64 *target, long long batch, long long target_stride,
                           u64 *direct)
    {
        const long long total2 = batch * t.limbs * (long long)(c->n / 2);
        const long long stride2 = (target_stride ? target_stride : (long long)t.limbs * (long long)c->n) / 2;
        KernelTimer kt(c, "k_scale_slots", 1);
        k_scale_from<<<(unsigned)((total2 + 255) / 256), 256, 0, c->stream>>>(
            reinterpret_cast<const ulonglong2 *>(target), stride2, reinterpret_cast<ulonglong2 *>(direct), total2,
            c->log_n - 1, t.limbs, t.d_qe, c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ksg_decompose(Context *c, const u64 *target, long long batch, int limbs, int k, u64 *ext,
                       long long target_stride, int passes)
    {
        const KsgTables &t = tables(c, k, limbs);
        const size_t n = c->n;
        Scratch d((size_t)batch * limbs * n * sizeof(u64), c->stream);
        MOAI_REQUIRE(target_stride % (long long)n == 0, "target stride must be a whole number of limbs");
        // y_J = c_J * prod(E) (Q_g / q_J)^-1: the constant rides on the inverse transform's N^-1
        ntt_inverse_from(c, target, limbs, target_stride ? target_stride / (long long)n : limbs, d.as<u64>(), batch * limbs,
                         c->d_ids, limbs, 1, t.d_yscale, t.dec.fpsrc != nullptr);
        if (t.dec.BT)
        {
            conv_quotient(c, d.as<u64>(), batch, t.dec, t.digits);
        }
        Scratch vq(t.dec.fpsrc ? (size_t)batch * t.digits * n * sizeof(double) : 0, c->stream);
        NttPrologue pro;
        if (t.dec.fpsrc)
        {
            conv_quotient_fp(c, d.as<u64>(), batch, t.dec, t.digits, vq.as<double>());
            pro.conv_v = vq.as<double>();
        }
        pro.src = d.as<u64>();
        pro.mode = 3;
        ConvTab dec = t.dec;
        // pass A alone feeds the fused key-switch kernel, which takes every digit's residue modulo its own primes
        // directly (ksg_direct): those polynomials are skipped
        dec.own = passes == 1 ? t.d_own : nullptr;
        pro.skipped = passes == 1 ? batch * t.n_own : 0;
        pro.conv = &dec;
        ntt_forward(c, ext, batch * t.rns * t.digits, t.d_ids, t.rns, t.digits, &pro, passes);
    }

    // out[P][limbs - 1][n] = round((acc + P' addend) / (P' q_{limbs-1})): the rescale that follows a key switch, taken
    // in the same division (the level's last data prime is one more prime of the special modulus): 2 (limbs - 1) + 2
    // transforms fewer than mod-down + rescale_to_next (S/evaluator.cpp:1402-1481 after :2910-3018), one rounding
    // instead of two.  acc's limb limbs - 1 is modified in place.
    void ksg_moddown_rescale(Context *c, u64 *acc, long long polys, int limbs, int k, const u64 *addend,
                             bool addend_even_only, u64 *out, int addend_group)
    {
        const KsgTables &t = tables(c, k, limbs);
        MOAI_REQUIRE(limbs >= 2, "end of modulus switching chain reached");
        const size_t n = c->n;
        const int np = k + 2, targets = limbs - 1;
        if (addend)
        {
            const long long total2 = polys * (long long)(n / 2);
            KernelTimer kt(c, "k_addsub", 1);
            k_add_scaled_limb<<<(unsigned)((total2 + 255) / 256), 256, 0, c->stream>>>(
                reinterpret_cast<ulonglong2 *>(acc), reinterpret_cast<const ulonglong2 *>(addend), total2, c->log_n - 1,
                t.rns, limbs, limbs - 1, addend_group, addend_even_only ? 1 : 0, t.pmod_last, c->q[limbs - 1]);
            c->launches += 1;
            MOAI_CUDA_CHECK(cudaGetLastError());
        }
        Scratch r((size_t)polys * np * n * sizeof(u64), c->stream);
        ntt_inverse_from(c, acc + (size_t)targets * n, np, t.rns, r.as<u64>(), polys * np, t.d_ids + targets, np, 1,
                         t.d_zscale_r, t.mdr.fpsrc != nullptr);
        if (t.mdr.BT)
        {
            conv_quotient(c, r.as<u64>(), polys, t.mdr, 1);
        }
        Scratch u((size_t)polys * targets * n * sizeof(u64), c->stream);
        Scratch vq(t.mdr.fpsrc ? (size_t)polys * n * sizeof(double) : 0, c->stream);
        NttPrologue pro;
        if (t.mdr.fpsrc)
        {
            conv_quotient_fp(c, r.as<u64>(), polys, t.mdr, 1, vq.as<double>());
            pro.conv_v = vq.as<double>();
        }
        pro.src = r.as<u64>();
        pro.mode = 3;
        pro.conv = &t.mdr;
        FinishEpi fin;
        fin.in = acc;
        fin.addend = addend;
        fin.out = out;
        fin.inv = t.d_pinv_r;
        fin.limbs_in = t.rns;
        fin.addend_even_only = addend_even_only ? 1 : 0;
        fin.addend_group = addend_group;
        fin.addend_mul = t.d_qlinv;
        fin.addend_limbs = limbs;
        if (ntt_forward(c, u.as<u64>(), polys * targets, c->d_ids, targets, 1, &pro, 3, &fin))
        {
            return;
        }
        divround_finish(c, acc, u.as<u64>(), addend, out, polys, targets, t.rns, t.d_pinv_r, addend_even_only, addend_group,
                        t.d_qlinv, limbs);
    }

    void ksg_moddown(Context *c, const u64 *acc, long long polys, int limbs, int k, const u64 *addend,
                     bool addend_even_only, u64 *out, int addend_group, int in_stride)
    {
        const KsgTables &t = tables(c, k, limbs);
        const size_t n = c->n;
        const int np = k + 1;
        const int rns_in = t.rns * in_stride; // limbs between consecutive input polynomials
        Scratch r((size_t)polys * np * n * sizeof(u64), c->stream);
        ntt_inverse_from(c, acc + (size_t)limbs * n, np, rns_in, r.as<u64>(), polys * np, t.d_ids + limbs, np, 1,
                         t.d_zscale, t.md.fpsrc != nullptr);
        if (t.md.BT)
        {
            conv_quotient(c, r.as<u64>(), polys, t.md, 1);
        }
        Scratch u((size_t)polys * limbs * n * sizeof(u64), c->stream);
        Scratch vq(t.md.fpsrc ? (size_t)polys * n * sizeof(double) : 0, c->stream);
        NttPrologue pro;
        if (t.md.fpsrc)
        {
            conv_quotient_fp(c, r.as<u64>(), polys, t.md, 1, vq.as<double>());
            pro.conv_v = vq.as<double>();
        }
        pro.src = r.as<u64>();
        pro.mode = 3;
        pro.conv = &t.md;
        FinishEpi fin;
        fin.in = acc;
        fin.addend = addend;
        fin.out = out;
        fin.inv = t.d_pinv;
        fin.limbs_in = rns_in;
        fin.addend_even_only = addend_even_only ? 1 : 0;
        fin.addend_group = addend_group;
        if (ntt_forward(c, u.as<u64>(), polys * limbs, c->d_ids, limbs, 1, &pro, 3, &fin))
        {
            return;
        }
        divround_finish(c, acc, u.as<u64>(), addend, out, polys, limbs, rns_in, t.d_pinv, addend_even_only, addend_group);
    }

    // inner products of the integer-path target moduli (the special prime): plain pass B + 128-bit MAC
    void ks_int_targets(Context *c, const KsShape &sh, const std::vector<int> &h_ids, u64 *ext, long long batch,
                        const u64 *ksk, int key_kl, u64 *acc, bool need_pass_b)
    {
        const size_t n = c->n;
        for (int I = 0; I < sh.rns; I++)
        {
            const int prime = h_ids[I];
            if (c->h_limb[prime].fp_class != 0)
            {
                continue;
            }
            if (need_pass_b)
            {
                ntt_forward_pass_b_strided(c, ext + (size_t)I * sh.digits * n, batch, sh.digits,
                                           (long long)sh.rns * sh.digits, c->d_ids + prime);
            }
            ks_mac_int(c, ext, ksk, acc, batch, sh, key_kl, I);
        }
    }
    static void ksg_int_targets(Context *c, const KsgTables &t, u64 *ext, long long batch, const u64 *ksk, int key_kl,
                                u64 *acc, bool need_pass_b)
    {
        ks_int_targets(c, t.shape, t.h_ids, ext, batch, ksk, key_kl, acc, need_pass_b);
    }

    KsExtInfo ks_ext_info(Context *c, int k, int limbs)
    {
        const KsgTables &t = tables(c, k, limbs);
        KsExtInfo e;
        e.shape = t.shape;
        e.k = k;
        e.h_ids = t.h_ids;
        e.h_pmod = t.h_pmod;
        return e;
    }

    KsExtInfo ks_ext_info_single(Context *c, int limbs)
    {
        KsExtInfo e = ks_ext_info(c, 0, limbs);
        e.shape.digits = 1;
        e.k = KS_SINGLE;
        return e;
    }

    void ks_moddown(Context *c, const u64 *acc, long long polys, int limbs, int k, const u64 *addend, bool addend_even_only,
                    u64 *out)
    {
        if (k > 0)
        {
            ksg_moddown(c, acc, polys, limbs, k, addend, addend_even_only, out);
        }
        else
        {
            moddown_special(c, acc, polys, limbs, addend, out, addend_even_only);
        }
    }

    // centred lift of coefficients modulo q_0 into the prime `target` (one limb of ModRaise)
    namespace
    {
        __global__ void k_lift_q0(const ulonglong2 *__restrict__ d, ulonglong2 *__restrict__ out, long long total2, u64 q0,
                                  LimbConst lc)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total2)
            {
                return;
            }
            const u64 half = q0 >> 1, corr = lc.q - reduce64(q0, lc);
            const ulonglong2 v = d[i];
            ulonglong2 r;
            r.x = reduce64(v.x, lc);
            r.y = reduce64(v.y, lc);
            if (v.x > half)
            {
                r.x = addmod(r.x, corr, lc.q);
            }
            if (v.y > half)
            {
                r.y = addmod(r.y, corr, lc.q);
            }
            out[i] = r;
        }
    } // namespace

    size_t ks_single_ext_bytes_per_ct(Context *c, int limbs)
    {
        return (size_t)(limbs + 1) * c->n * sizeof(u64);
    }

    void ks_hoist_modraised(Context *c, const u64 *c1, long long batch, int limbs, u64 *ext, long long c1_stride)
    {
        const size_t n = c->n;
        const size_t row = (size_t)limbs * n * sizeof(u64);
        {
            KernelTimer ktm(c, "k_copy_ks_target", 1);
            MOAI_CUDA_CHECK(cudaMemcpy2DAsync(ext, row + n * sizeof(u64), c1, c1_stride ? (size_t)c1_stride * sizeof(u64) : row,
                                              row, (size_t)batch, cudaMemcpyDeviceToDevice, c->stream));
        }
        // the special-prime limb: INTT of limb 0 (coefficients modulo q_0), centred lift, NTT
        Scratch d((size_t)batch * n * sizeof(u64), c->stream);
        {
            KernelTimer ktm(c, "k_copy_ks_target", 1);
            MOAI_CUDA_CHECK(cudaMemcpy2DAsync(d.p, n * sizeof(u64), c1, c1_stride ? (size_t)c1_stride * sizeof(u64) : row,
                                              n * sizeof(u64), (size_t)batch, cudaMemcpyDeviceToDevice, c->stream));
        }
        ntt_inverse(c, d.as<u64>(), batch, c->d_ids, 1);
        Scratch sp((size_t)batch * n * sizeof(u64), c->stream);
        const long long total2 = batch * (long long)(n / 2);
        {
            KernelTimer kt(c, "k_lift_q0", 1);
            k_lift_q0<<<(unsigned)((total2 + 255) / 256), 256, 0, c->stream>>>(d.as<ulonglong2>(), sp.as<ulonglong2>(), total2,
                                                                               c->q[0], c->h_limb[c->kl - 1]);
            c->launches += 1;
        }
        ntt_forward(c, sp.as<u64>(), batch, c->d_ids + (c->kl - 1), 1);
        {
            KernelTimer ktm(c, "k_copy_ks_target", 1);
            MOAI_CUDA_CHECK(cudaMemcpy2DAsync(ext + (size_t)limbs * n, row + n * sizeof(u64), sp.p, n * sizeof(u64),
                                              n * sizeof(u64), (size_t)batch, cudaMemcpyDeviceToDevice, c->stream));
        }
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ks_key_prepare_single(Context *c, const u64 *in, uint32_t elt, bool pre_permute, u64 *out)
    {
        MOAI_REQUIRE(in != out, "key preparation is out of place");
        const size_t n = c->n;
        const int L = c->kl - 1;
        uint32_t inv = 1;
        if (pre_permute)
        {
            const uint64_t m = 2 * (uint64_t)n;
            uint64_t x = elt;
            for (int i = 0; i < 6; i++)
            {
                x = (x * (2 + m * 4 - (uint64_t)elt * x % m)) % m;
            }
            MOAI_REQUIRE((uint64_t)elt * x % m == 1, "Galois element is not invertible");
            inv = (uint32_t)x;
        }
        Scratch tmp(n * sizeof(u64), c->stream);
        const long long n2 = (long long)(n / 2);
        for (int kk = 0; kk < 2; kk++)
        {
            for (int ol = 0; ol < c->kl; ol++)
            {
                u64 *dst = out + ((size_t)kk * c->kl + ol) * n;
                u64 *sum = pre_permute ? tmp.as<u64>() : dst;
                for (int j0 = 0; j0 < L; j0 += CONV_MAX - 1)
                {
                    KeySumArgs a;
                    a.cnt = 0;
                    if (j0 > 0)
                    {
                        a.src[a.cnt++] = reinterpret_cast<const ulonglong2 *>(sum); // running sum (same-index read/write)
                    }
                    for (int j = j0; j < std::min(L, j0 + CONV_MAX - 1); j++)
                    {
                        a.src[a.cnt++] = reinterpret_cast<const ulonglong2 *>(in + (((size_t)j * 2 + kk) * c->kl + ol) * n);
                    }
                    KernelTimer kt(c, "k_key_group_sum", 1);
                    k_key_group_sum<<<(unsigned)((n2 + 255) / 256), 256, 0, c->stream>>>(
                        a, reinterpret_cast<ulonglong2 *>(sum), n2, c->q[ol]);
                    c->launches += 1;
                }
                if (pre_permute)
                {
                    apply_galois_ntt(c, sum, dst, 1, inv);
                }
            }
        }
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ksg_switch_acc(Context *c, const u64 *target, long long batch, int limbs, int k, const u64 *ksk, int key_kl,
                        u64 *ext, u64 *direct, u64 *acc, long long target_stride)
    {
        const KsgTables &t = tables(c, k, limbs);
        MOAI_REQUIRE(key_kl >= t.rns, "grouped key does not cover this level");
        ksg_decompose(c, target, batch, limbs, k, ext, target_stride, /*passes=*/1);
        ksg_direct(c, t, target, batch, target_stride, direct);
        ks_passb_mac(c, ext, batch, t.shape, ksk, key_kl, acc, direct, t.d_own);
        ksg_int_targets(c, t, ext, batch, ksk, key_kl, acc, true);
    }

    namespace
    {
        struct GiantSumArgs
        {
            const u64 *acc[BSGS_MAX_GIANT];      // [batch][2][rns][n]: key-switch sums of a giant step (or its inner sums)
            const u64 *extra[BSGS_MAX_GIANT];    // [batch][2][rns][n] or nullptr: polynomial 0 is added before the rotation
            const uint32_t *perm[BSGS_MAX_GIANT]; // Galois table of the giant step (nullptr = identity)
            int n;
        };
        // total[b][p][I][.] = sum_g sigma_g(acc_g[b][p][I][.] + (p == 0 ? extra_g[b][0][I][.] : 0))   (mod m_I)
        __global__ void k_giants_sum(GiantSumArgs a, ulonglong2 *total, long long total2, int log_n, int rns,
                                     const int *__restrict__ ids, const LimbConst *__restrict__ lcs)
        {
            const long long i2 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i2 >= total2)
            {
                return;
            }
            const long long i = i2 * 2;
            const long long pl = i >> log_n; // (b * 2 + p) * rns + I
            const int I = (int)(pl % rns);
            const bool even = ((pl / rns) & 1) == 0;
            const u64 q = lcs[ids[I]].q;
            const long long base = pl << log_n;
            const uint32_t within = (uint32_t)(i & (((long long)1 << log_n) - 1));
            u64 sx = 0, sy = 0;
            for (int g = 0; g < a.n; g++)
            {
                const uint32_t *tb = a.perm[g];
                const long long jx = base + (tb ? tb[within] : within), jy = base + (tb ? tb[within + 1] : within + 1);
                u64 vx = a.acc[g][jx], vy = a.acc[g][jy];
                if (even && a.extra[g])
                {
                    vx = addmod(vx, a.extra[g][jx], q);
                    vy = addmod(vy, a.extra[g][jy], q);
                }
                sx = addmod(sx, vx, q);
                sy = addmod(sy, vy, q);
            }
            total[i2] = make_ulonglong2(sx, sy);
        }
    } // namespace

    void ksg_giants_sum(Context *c, int n_giants, const u64 *const *acc, const u64 *const *extra,
                        const uint32_t *const *perm, u64 *total, long long batch, int limbs, int k)
    {
        const KsgTables &t = tables(c, k, limbs);
        MOAI_REQUIRE(n_giants >= 1 && n_giants <= BSGS_MAX_GIANT, "too many giant steps for one pass");
        GiantSumArgs a;
        a.n = n_giants;
        for (int g = 0; g < n_giants; g++)
        {
            a.acc[g] = acc[g];
            a.extra[g] = extra[g];
            a.perm[g] = perm[g];
        }
        const long long total2 = batch * 2 * t.rns * (long long)(c->n / 2);
        KernelTimer kt(c, "k_giants_sum", 1);
        k_giants_sum<<<(unsigned)((total2 + 255) / 256), 256, 0, c->stream>>>(a, reinterpret_cast<ulonglong2 *>(total), total2,
                                                                              c->log_n, t.rns, t.d_ids, c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ksg_switch(Context *c, const u64 *target, long long batch, int limbs, int k, const u64 *ksk, int key_kl,
                    const u64 *addend, u64 *out, long long target_stride, bool addend_c0_only, int addend_group,
                    bool rescale)
    {
        const KsgTables &t = tables(c, k, limbs);
        MOAI_REQUIRE(key_kl >= t.rns, "grouped key does not cover this level");
        const size_t n = c->n;
        const size_t per_ext = ksg_ext_bytes_per_ct(c, limbs, k);
        long long chunk = (long long)(ksg_ext_budget_bytes() / per_ext);
        chunk = chunk < 1 ? 1 : (chunk > batch ? batch : chunk);
        chunk = (batch + (batch + chunk - 1) / chunk - 1) / ((batch + chunk - 1) / chunk); // equal chunks, no short tail
        Scratch ext((size_t)chunk * per_ext, c->stream);
        Scratch acc((size_t)chunk * 2 * t.rns * n * sizeof(u64), c->stream);
        Scratch direct((size_t)chunk * limbs * n * sizeof(u64), c->stream);
        const size_t tstride = target_stride ? (size_t)target_stride : (size_t)limbs * n;
        for (long long b0 = 0; b0 < batch; b0 += chunk)
        {
            const long long nb = std::min(chunk, batch - b0);
            ksg_decompose(c, target + (size_t)b0 * tstride, nb, limbs, k, ext.as<u64>(), target_stride, /*passes=*/1);
            ksg_direct(c, t, target + (size_t)b0 * tstride, nb, target_stride, direct.as<u64>());
            ks_passb_mac(c, ext.as<u64>(), nb, t.shape, ksk, key_kl, acc.as<u64>(), direct.as<u64>(), t.d_own);
            ksg_int_targets(c, t, ext.as<u64>(), nb, ksk, key_kl, acc.as<u64>(), true);
            const u64 *ad = addend ? addend + (size_t)b0 * addend_group * limbs * n : nullptr;
            if (rescale)
            {
                ksg_moddown_rescale(c, acc.as<u64>(), nb * 2, limbs, k, ad, addend_c0_only,
                                    out + (size_t)b0 * 2 * (limbs - 1) * n, addend_group);
                continue;
            }
            const size_t off = (size_t)b0 * 2 * limbs * n;
            ksg_moddown(c, acc.as<u64>(), nb * 2, limbs, k, ad, addend_c0_only, out + off, addend_group);
        }
    }

    void ksg_rotate_hoisted_multi(Context *c, const u64 *ct, const u64 *ext, long long batch, int limbs, int k, int n_rot,
                                  const uint32_t *elts, const u64 *const *ksk_pre, const int *key_kl, u64 *const *outs)
    {
        MOAI_REQUIRE(n_rot >= 1 && n_rot <= KSM_R, "too many rotations for one multi-key pass");
        const KsgTables &t = tables(c, k, limbs);
        const size_t n = c->n;
        const size_t acc_words = (size_t)batch * 2 * t.rns * n;
        Scratch acc(acc_words * n_rot * sizeof(u64), c->stream);
        u64 *accp[KSM_R];
        for (int r = 0; r < n_rot; r++)
        {
            accp[r] = acc.as<u64>() + acc_words * r;
            MOAI_REQUIRE(key_kl[r] >= t.rns, "grouped key does not cover this level");
        }
        ks_mac_multi(c, ext, batch, t.shape, n_rot, ksk_pre, key_kl, accp);
        for (int r = 0; r < n_rot; r++)
        {
            ksg_int_targets(c, t, const_cast<u64 *>(ext), batch, ksk_pre[r], key_kl[r], accp[r], false);
        }
        Scratch tmp((size_t)batch * 2 * limbs * n * sizeof(u64), c->stream);
        for (int r = 0; r < n_rot; r++)
        {
            ksg_moddown(c, accp[r], batch * 2, limbs, k, ct, true, tmp.as<u64>());
            apply_galois_ntt(c, tmp.as<u64>(), outs[r], batch * 2 * limbs, elts[r]);
        }
    }
} // namespace moai
