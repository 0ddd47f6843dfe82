// Context construction: derives every table the kernels need from (log N, prime chain).
// The caller passes the primes of its SEALContext key level (data primes then the special
// prime); the tables below are the same mathematical objects SEAL builds in
// NTTTables::initialize (S/util/ntt.cpp:241-300), RNSTool::initialize (S/util/rns.cpp:578-787),
// GaloisTool (S/util/galois.cpp:18-95) and CKKSEncoder::CKKSEncoder (S/ckks.cpp:20-75).
#include <memory>
#include "context.hpp"
#include <chrono>
#include "comm.hpp"
#include <set>
#include <cstdio>
#include <cmath>
#include <complex>
#include <cstring>
#include <cstdlib>

namespace moai
{
    namespace
    {
        typedef unsigned __int128 u128h;

        thread_local std::string g_last_error;

        u64 h_mulmod(u64 a, u64 b, u64 q)
        {
            return (u64)(((u128h)a * b) % q);
        }

        u64 h_powmod(u64 a, u64 e, u64 q)
        {
            u64 r = 1;
            a %= q;
            for (; e; e >>= 1)
            {
                if (e & 1)
                {
                    r = h_mulmod(r, a, q);
                }
                a = h_mulmod(a, a, q);
            }
            return r;
        }

        u64 h_invmod(u64 a, u64 q)
        {
            return h_powmod(a % q, q - 2, q);
        }

        u64 h_shoup(u64 w, u64 q)
        {
            return (u64)((((u128h)w) << 64) / q);
        }

        u64 bitrev(u64 x, int bits)
        {
            u64 r = 0;
            for (int i = 0; i < bits; i++)
            {
                r = (r << 1) | ((x >> i) & 1);
            }
            return r;
        }

        // Smallest primitive 2N-th root of unity mod q (what SEAL's try_minimal_primitive_root
        // selects, S/util/numth.cpp:385-413): take any primitive root and scan its odd powers.
        u64 minimal_primitive_root(u64 two_n, u64 q)
        {
            u64 cofactor = (q - 1) / two_n;
            u64 g = 0;
            for (u64 a = 2;; a++)
            {
                g = h_powmod(a, cofactor, q);
                if (h_powmod(g, two_n >> 1, q) == q - 1)
                {
                    break;
                }
            }
            u64 g2 = h_mulmod(g, g, q), cur = g, best = g;
            for (u64 i = 0; i < two_n; i += 2)
            {
                best = cur < best ? cur : best;
                cur = h_mulmod(cur, g2, q);
            }
            return best;
        }

        template <class T>
        T *to_device(const std::vector<T> &h)
        {
            T *d = nullptr;
            MOAI_CUDA_CHECK(cudaMalloc(&d, h.size() * sizeof(T)));
            MOAI_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
            return d;
        }
    } // namespace

    void set_last_error(const std::string &msg)
    {
        g_last_error = msg;
    }

    const std::string &last_error()
    {
        return g_last_error;
    }

    // ---------------------------------------------------------------------------------------------
    // Device memory: a best-fit, splitting / coalescing block cache over cudaMalloc'ed segments, one
    // arena pair per stream.
    //
    // Every temporary of the pipeline (ciphertext batches, extended digits, ...) is allocated and freed
    // in stream order on the context's stream, so a freed block may be handed out again immediately:
    // the work that last touched it precedes the new owner's work on the same stream.  That makes
    // allocation a host-side map lookup.  (cudaMallocAsync, used before, spent 64 s of host time per
    // encoder layer growing / defragmenting its pool under GiB-sized, ever-changing requests and
    // stalled the launch queue: profiles/layer_r1_fast_b.json "alloc_host".)  When cudaMalloc fails,
    // all cached blocks are returned to the driver and the request is retried once.
    // ---------------------------------------------------------------------------------------------
    static thread_local const char *g_phase = "";
    void set_phase(const char *name)
    {
        g_phase = name ? name : "";
    }

    namespace
    {
        // One arena = a set of cudaMalloc'ed segments carved into blocks: best-fit with splitting on
        // allocation, coalescing with the neighbours on release (the scheme of caching tensor
        // allocators).  Unlike one free list per size, a freed 17 GiB batch can serve the next
        // level's slightly smaller requests, so the memory held stays close to the working set.
        struct Block
        {
            char *ptr;
            size_t size;
            bool is_free;
            Block *prev, *next; // neighbours inside the same segment
        };

        struct Arena
        {
            size_t granule, min_segment;
            std::set<std::pair<size_t, Block *>> free_set; // (size, block), best fit = lower_bound
            std::map<void *, Block *> used;
            size_t owned = 0, in_use = 0;

            Arena(size_t g, size_t seg) : granule(g), min_segment(seg)
            {}

            void *take(Block *b, size_t want)
            {
                if (b->size - want >= granule)
                {
                    Block *rest = new Block{ b->ptr + want, b->size - want, true, b, b->next };
                    if (b->next)
                    {
                        b->next->prev = rest;
                    }
                    b->next = rest;
                    b->size = want;
                    free_set.insert({ rest->size, rest });
                }
                b->is_free = false;
                used[b->ptr] = b;
                in_use += b->size;
                return b->ptr;
            }

            // nullptr when a new segment is needed
            void *alloc_cached(size_t want)
            {
                auto it = free_set.lower_bound({ want, nullptr });
                if (it == free_set.end())
                {
                    return nullptr;
                }
                Block *b = it->second;
                free_set.erase(it);
                return take(b, want);
            }

            cudaError_t grow(size_t want, void **out)
            {
                const size_t seg = want > min_segment ? want : min_segment;
                void *p = nullptr;
                cudaError_t e = cudaMalloc(&p, seg);
                if (e != cudaSuccess && seg > want)
                {
                    cudaGetLastError();
                    e = cudaMalloc(&p, want); // memory is tight: ask for exactly what is needed
                    if (e == cudaSuccess)
                    {
                        owned += want;
                        *out = take(new Block{ (char *)p, want, true, nullptr, nullptr }, want);
                        return e;
                    }
                }
                if (e != cudaSuccess)
                {
                    return e;
                }
                owned += seg;
                *out = take(new Block{ (char *)p, seg, true, nullptr, nullptr }, want);
                return e;
            }

            bool release(void *p)
            {
                auto it = used.find(p);
                if (it == used.end())
                {
                    return false;
                }
                Block *b = it->second;
                used.erase(it);
                in_use -= b->size;
                b->is_free = true;
                if (b->next && b->next->is_free)
                {
                    Block *n = b->next;
                    free_set.erase({ n->size, n });
                    b->size += n->size;
                    b->next = n->next;
                    if (n->next)
                    {
                        n->next->prev = b;
                    }
                    delete n;
                }
                if (b->prev && b->prev->is_free)
                {
                    Block *pv = b->prev;
                    free_set.erase({ pv->size, pv });
                    pv->size += b->size;
                    pv->next = b->next;
                    if (b->next)
                    {
                        b->next->prev = pv;
                    }
                    delete b;
                    b = pv;
                }
                free_set.insert({ b->size, b });
                return true;
            }

            // return every completely free segment to the driver
            void trim()
            {
                for (auto it = free_set.begin(); it != free_set.end();)
                {
                    Block *b = it->second;
                    if (!b->prev && !b->next)
                    {
                        cudaFree(b->ptr);
                        owned -= b->size;
                        it = free_set.erase(it);
                        delete b;
                    }
                    else
                    {
                        ++it;
                    }
                }
            }
        };

        struct DeviceCache
        {
            std::mutex mu;
            struct PerStream
            {
                Arena small{ 512, (size_t)8 << 20 };            // requests below 1 MiB
                Arena large{ (size_t)1 << 20, (size_t)1 << 30 }; // everything else, 1 GiB segments at least
            };
            // keyed by (device, stream): stream 0 / torch's default stream is the same handle on every device, and a
            // block cudaMalloc'ed on one device must never serve a context on another
            typedef std::pair<int, cudaStream_t> ArenaKey;
            std::map<ArenaKey, PerStream> arenas;
            AllocStats stats;

            static int current_device()
            {
                int dev = 0;
                cudaGetDevice(&dev);
                return dev;
            }

            // every device this cache holds memory on (trim returns blocks that in-flight work may still use)
            void sync_all_devices_locked()
            {
                const int cur = current_device();
                std::set<int> devs{ cur };
                for (auto &kv : arenas)
                {
                    devs.insert(kv.first.first);
                }
                for (int d : devs)
                {
                    cudaSetDevice(d);
                    cudaDeviceSynchronize();
                }
                cudaSetDevice(cur);
            }

            void trim_all_locked()
            {
                for (auto &kv : arenas)
                {
                    kv.second.small.trim();
                    kv.second.large.trim();
                }
            }

            size_t owned_locked() const
            {
                size_t t = 0;
                for (auto &kv : arenas)
                {
                    t += kv.second.small.owned + kv.second.large.owned;
                }
                return t;
            }

            size_t in_use_locked() const
            {
                size_t t = 0;
                for (auto &kv : arenas)
                {
                    t += kv.second.small.in_use + kv.second.large.in_use;
                }
                return t;
            }

            void *alloc(size_t bytes, cudaStream_t stream)
            {
                bytes = bytes ? bytes : 8;
                std::lock_guard<std::mutex> lock(mu);
                stats.calls += 1;
                PerStream &ps = arenas[ArenaKey(current_device(), stream)];
                Arena &a = bytes < ((size_t)1 << 20) ? ps.small : ps.large;
                const size_t want = (bytes + a.granule - 1) / a.granule * a.granule;
                if (void *p = a.alloc_cached(want))
                {
                    return p;
                }
                const auto t0 = std::chrono::steady_clock::now();
                void *p = nullptr;
                cudaError_t e = a.grow(want, &p);
                if (e != cudaSuccess)
                {
                    cudaGetLastError();
                    stats.retries += 1;
                    sync_all_devices_locked();
                    trim_all_locked();
                    e = a.grow(want, &p);
                }
                stats.host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                if (e != cudaSuccess)
                {
                    cudaGetLastError();
                    size_t free_b = 0, total_b = 0;
                    cudaMemGetInfo(&free_b, &total_b);
                    char buf[320];
                    snprintf(buf, sizeof buf,
                             "device allocation of %.2f GiB failed in phase '%s': %s (device free %.1f of %.1f GiB; "
                             "this library holds %.1f GiB, %.1f GiB in use)",
                             want / 1073741824.0, g_phase, cudaGetErrorString(e), free_b / 1073741824.0,
                             total_b / 1073741824.0, owned_locked() / 1073741824.0, in_use_locked() / 1073741824.0);
                    throw StatusError{ e == cudaErrorMemoryAllocation ? OUT_OF_MEMORY : CUDA_ERROR, buf };
                }
                return p;
            }

            void release(void *p, cudaStream_t stream)
            {
                std::lock_guard<std::mutex> lock(mu);
                auto it = arenas.find(ArenaKey(current_device(), stream));
                if (it != arenas.end() && (it->second.large.release(p) || it->second.small.release(p)))
                {
                    return;
                }
                for (auto &kv : arenas) // freed on another stream (or with another current device) than it was allocated on
                {
                    if (kv.second.large.release(p) || kv.second.small.release(p))
                    {
                        return;
                    }
                }
            }
        };

        DeviceCache &cache()
        {
            static DeviceCache *c = new DeviceCache(); // leaked on purpose: outlives every static destructor
            return *c;
        }
    } // namespace

    AllocStats alloc_stats()
    {
        std::lock_guard<std::mutex> lock(cache().mu);
        AllocStats s = cache().stats;
        s.owned_bytes = cache().owned_locked();
        s.cached_bytes = s.owned_bytes - cache().in_use_locked();
        return s;
    }

    void *device_alloc(size_t bytes, cudaStream_t stream)
    {
        return cache().alloc(bytes, stream);
    }

    void device_free(void *p, cudaStream_t stream)
    {
        if (p)
        {
            cache().release(p, stream);
        }
    }

    void device_release_cached()
    {
        std::lock_guard<std::mutex> lock(cache().mu);
        cache().sync_all_devices_locked();
        cache().trim_all_locked();
    }

    Scratch::Scratch(size_t bytes, cudaStream_t stream) : s(stream)
    {
        if (bytes)
        {
            p = device_alloc(bytes, stream);
        }
    }

    Scratch::~Scratch()
    {
        device_free(p, s);
    }

    void Context::kernel_timers_collect(bool wait)
    {
        if (wait && !kpending.empty())
        {
            cudaEventSynchronize(kpending.back().e1);
        }
        while (!kpending.empty())
        {
            PendingKernel &k = kpending.front();
            if (cudaEventQuery(k.e1) != cudaSuccess)
            {
                cudaGetLastError();
                break;
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, k.e0, k.e1);
            auto &slot = prof[k.name];
            slot.first += ms;
            slot.second += k.units;
            kpool.push_back(k.e0);
            kpool.push_back(k.e1);
            kpending.pop_front();
        }
    }

    Context::~Context()
    {
        comm_destroy(this);
        kernel_timers_collect(true);
        for (cudaEvent_t e : kpool)
        {
            cudaEventDestroy(e);
        }
        if (owns_stream && stream)
        {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
        }
        if (parent)
        {
            return; // a lane: every table belongs to the parent
        }
        cudaFree(d_limb);
        cudaFree(d_fwd);
        cudaFree(d_inv);
        cudaFree(d_fwd_fp);
        cudaFree(d_inv_fp);
        cudaFree(d_inv_last);
        cudaFree(d_half_mod);
        cudaFree(d_two64);
        cudaFree(d_ids);
        cudaFree(d_ids_ks);
        cudaFree(d_fft_inv_roots);
        cudaFree(d_index_map);
        for (auto &kv : galois_tables)
        {
            cudaFree(kv.second);
        }
        ksg_release(this);
    }

    uint32_t Context::elt_from_step(int step) const
    {
        // generator 5 (fork: S/util/galois.h:169); step 0 = complex conjugation (2N - 1)
        uint64_t m = (uint64_t)n << 1;
        if (step == 0)
        {
            return (uint32_t)(m - 1);
        }
        uint32_t pos = (uint32_t)(step < 0 ? -(int64_t)step : step);
        if (pos >= (n >> 1))
        {
            throw StatusError{ INVALID_ARGUMENT, "step count too large" };
        }
        uint32_t e = step < 0 ? (uint32_t)(n >> 1) - pos : pos;
        uint64_t g = 1;
        while (e--)
        {
            g = (g * 5) & (m - 1);
        }
        return (uint32_t)g;
    }

    const uint32_t *Context::galois_table(uint32_t elt)
    {
        if (parent)
        {
            return parent->galois_table(elt);
        }
        std::lock_guard<std::mutex> lk(galois_mu);
        auto it = galois_tables.find(elt);
        if (it != galois_tables.end())
        {
            return it->second;
        }
        if (!(elt & 1) || elt >= 2 * n)
        {
            throw StatusError{ INVALID_ARGUMENT, "Galois element is not valid" };
        }
        // NTT-domain automorphism as an index permutation (S/util/galois.cpp:36-43)
        std::vector<uint32_t> h(n);
        for (size_t i = 0; i < n; i++)
        {
            u64 rev = bitrev(n + i, log_n + 1);
            u64 idx = (((u64)elt * rev) >> 1) & (n - 1);
            h[i] = (uint32_t)bitrev(idx, log_n);
        }
        uint32_t *d = to_device(h);
        galois_tables[elt] = d;
        return d;
    }

    Context *context_fork(Context *parent)
    {
        MOAI_REQUIRE(parent != nullptr, "null context");
        Context *r = parent->root();
        MOAI_CUDA_CHECK(cudaSetDevice(r->device));
        Context *c = new Context();
        c->parent = r;
        c->device = r->device;
        c->log_n = r->log_n;
        c->n = r->n;
        c->kl = r->kl;
        c->q = r->q;
        c->sm_count = r->sm_count;
        c->d_limb = r->d_limb;
        c->d_fwd = r->d_fwd;
        c->d_inv = r->d_inv;
        c->d_fwd_fp = r->d_fwd_fp;
        c->d_inv_fp = r->d_inv_fp;
        c->d_inv_last = r->d_inv_last;
        c->d_half_mod = r->d_half_mod;
        c->d_two64 = r->d_two64;
        c->d_ids = r->d_ids;
        c->d_ids_ks = r->d_ids_ks;
        c->d_fft_inv_roots = r->d_fft_inv_roots;
        c->d_index_map = r->d_index_map;
        c->h_fft_inv_roots = r->h_fft_inv_roots;
        c->h_index_map = r->h_index_map;
        c->h_limb = r->h_limb;
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess)
        {
            c->parent = r; // nothing owned yet
            delete c;
            MOAI_CUDA_CHECK(e);
        }
        c->owns_stream = true;
        return c;
    }

    Context *context_create(int log_n, const u64 *primes, int kl, int device)
    {
        MOAI_REQUIRE(log_n >= 12 && log_n <= 16, "log_n must be in [12, 16]");
        MOAI_REQUIRE(kl >= 2 && kl <= 64, "need 2..64 primes (data primes + special prime)");
        MOAI_CUDA_CHECK(cudaSetDevice(device));
        std::unique_ptr<Context> owner(new Context()); // freed if a later check throws
        Context *c = owner.get();
        c->device = device;
        c->log_n = log_n;
        c->n = (size_t)1 << log_n;
        c->kl = kl;
        c->q.assign(primes, primes + kl);
        size_t n = c->n;
        cudaDeviceProp prop;
        MOAI_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        c->sm_count = prop.multiProcessorCount;

        std::vector<Twiddle> fwd((size_t)kl * n), inv((size_t)kl * n);
        std::vector<double> fwd_fp((size_t)kl * n), inv_fp((size_t)kl * n);
        c->h_limb.resize(kl);
        // NTT arithmetic selection (see csrc/ntt.cuh).  MOAI_NTT_FP=0 forces the integer path;
        // MOAI_NTT_WIDE_FP_EVERY=k sends every k-th 48..51-bit prime through the FP64 path (default 1 =
        // all of them; 0 = none).  Measured on B200 (profiles/ntt_arith_r1.md): all-FP64 is fastest; leaving
        // part of the wide primes on the integer pipe does not overlap the two pipes enough to pay.
        const bool use_fp = !(getenv("MOAI_NTT_FP") && atoi(getenv("MOAI_NTT_FP")) == 0);
        const int wide_every = getenv("MOAI_NTT_WIDE_FP_EVERY") ? atoi(getenv("MOAI_NTT_WIDE_FP_EVERY")) : 1;
        int wide_seen = 0;
        for (int l = 0; l < kl; l++)
        {
            u64 q = primes[l];
            MOAI_REQUIRE(q > 2 && (q >> 61) == 0 && (q - 1) % (2 * n) == 0, "prime must be < 2^61 and = 1 mod 2N");
            u64 psi = minimal_primitive_root(2 * n, q);
            u64 ipsi = h_invmod(psi, q);
            Twiddle *f = fwd.data() + (size_t)l * n, *g = inv.data() + (size_t)l * n;
            u64 pw = psi, ipw = ipsi;
            f[0] = Twiddle{ 1, h_shoup(1, q) };
            g[0] = Twiddle{ 1, h_shoup(1, q) };
            for (size_t i = 1; i < n; i++)
            {
                f[bitrev(i, log_n)] = Twiddle{ pw, h_shoup(pw, q) };
                g[bitrev(i - 1, log_n) + 1] = Twiddle{ ipw, h_shoup(ipw, q) };
                pw = h_mulmod(pw, psi, q);
                ipw = h_mulmod(ipw, ipsi, q);
            }
            LimbConst &lc = c->h_limb[l];
            lc.q = q;
            lc.two_q = q << 1;
            int bits = 64 - __builtin_clzll(q);
            lc.bar_shift = (u32)(bits - 1);
            lc.bar_m = (u64)((((u128h)1) << (64 + bits - 1)) / q);
            lc.pad = 0;
            lc.inv_n = h_invmod((u64)n % q, q);
            lc.inv_n_quo = h_shoup(lc.inv_n, q);
            lc.inv_n_w = h_mulmod(g[n - 1].w, lc.inv_n, q); // last GS stage root times N^-1
            lc.inv_n_w_quo = h_shoup(lc.inv_n_w, q);
            lc.q0_mod = primes[0] % q;
            auto sym = [&](u64 v) { return v > q / 2 ? -(double)(q - v) : (double)v; };
            lc.pd = (double)q;
            lc.pinv = 1.0 / (double)q;
            lc.inv_n_d = sym(lc.inv_n);
            lc.inv_n_w_d = sym(lc.inv_n_w);
            lc.fp_class = 0;
            lc.pad2 = 0;
            if (use_fp && (q >> 48) == 0)
            {
                lc.fp_class = 1;
            }
            else if (use_fp && 2 * q + 64 < ((u64)1 << 52))
            {
                wide_seen++;
                if (wide_every > 0 && wide_seen % wide_every == 0)
                {
                    lc.fp_class = 2;
                }
            }
            for (size_t i = 0; i < n; i++)
            {
                fwd_fp[(size_t)l * n + i] = sym(f[i].w);
                inv_fp[(size_t)l * n + i] = sym(g[i].w);
            }
        }
        c->d_fwd_fp = to_device(fwd_fp);
        c->d_inv_fp = to_device(inv_fp);
        c->d_fwd = to_device(fwd);
        c->d_inv = to_device(inv);
        c->d_limb = to_device(c->h_limb);

        std::vector<Twiddle> inv_last((size_t)kl * kl), two64(kl);
        std::vector<u64> half_mod((size_t)kl * kl);
        for (int last = 0; last < kl; last++)
        {
            for (int i = 0; i < kl; i++)
            {
                Twiddle t{ 0, 0 };
                if (i != last)
                {
                    t.w = h_invmod(primes[last] % primes[i], primes[i]);
                    t.wq = h_shoup(t.w, primes[i]);
                }
                inv_last[(size_t)last * kl + i] = t;
                half_mod[(size_t)last * kl + i] = (primes[last] >> 1) % primes[i];
            }
        }
        for (int i = 0; i < kl; i++)
        {
            u64 r = (u64)((((u128h)1) << 64) % primes[i]);
            two64[i] = Twiddle{ r, h_shoup(r, primes[i]) };
        }
        c->d_inv_last = to_device(inv_last);
        c->d_half_mod = to_device(half_mod);
        c->d_two64 = to_device(two64);

        std::vector<int> ids(kl), ids_ks((size_t)kl * (kl + 1), 0);
        for (int i = 0; i < kl; i++)
        {
            ids[i] = i;
        }
        for (int l = 1; l < kl; l++)
        {
            for (int i = 0; i < l; i++)
            {
                ids_ks[(size_t)l * (kl + 1) + i] = i;
            }
            ids_ks[(size_t)l * (kl + 1) + l] = kl - 1;
        }
        c->d_ids = to_device(ids);
        c->d_ids_ks = to_device(ids_ks);

        // CKKS encoder tables: index map (generator 5 orbit, S/ckks.cpp:36-52) and the inverse
        // FFT roots conj(zeta^{bitrev(i-1)+1}) (S/ckks.cpp:56-66, S/util/croots.cpp:17-72).
        size_t slots = n >> 1;
        u64 m = (u64)n << 1;
        c->h_index_map.resize(n);
        u64 pos = 1;
        for (size_t i = 0; i < slots; i++)
        {
            c->h_index_map[i] = (uint32_t)bitrev((pos - 1) >> 1, log_n);
            c->h_index_map[slots | i] = (uint32_t)bitrev((m - pos - 1) >> 1, log_n);
            pos = (pos * 5) & (m - 1);
        }
        // Octant table + symmetries, evaluated exactly like the reference so that the doubles
        // agree bit for bit with std::polar on the same libm.
        std::vector<std::complex<double>> oct(m / 8 + 1);
        const double PI = 3.1415926535897932384626433832795028842;
        for (size_t i = 0; i <= m / 8; i++)
        {
            oct[i] = std::polar<double>(1.0, 2 * PI * static_cast<double>(i) / static_cast<double>(m));
        }
        std::function<std::complex<double>(size_t)> root = [&](size_t index) -> std::complex<double> {
            index &= m - 1;
            if (index <= m / 8)
            {
                return oct[index];
            }
            else if (index <= m / 4)
            {
                auto a = oct[m / 4 - index];
                return { a.imag(), a.real() };
            }
            else if (index <= m / 2)
            {
                return -std::conj(root(m / 2 - index));
            }
            else if (index <= 3 * m / 4)
            {
                return -root(index - m / 2);
            }
            return std::conj(root(m - index));
        };
        c->h_fft_inv_roots.assign(n, double2{ 0, 0 });
        for (size_t i = 1; i < n; i++)
        {
            auto z = std::conj(root(bitrev(i - 1, log_n) + 1));
            c->h_fft_inv_roots[i] = double2{ z.real(), z.imag() };
        }
        c->d_fft_inv_roots = to_device(c->h_fft_inv_roots);
        c->d_index_map = to_device(c->h_index_map);

        // keep freed scratch memory cached in the stream-ordered pool
        cudaMemPool_t pool;
        MOAI_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thresh = UINT64_MAX;
        MOAI_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        return owner.release();
    }
} // namespace moai
