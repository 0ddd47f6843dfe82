// Host-side context of the B200 CKKS evaluation backend.
//
// Mirrors what the reference keeps per SEALContext::ContextData (S/context.h) for the hot
// path: the prime chain, per-prime NTT tables (S/util/ntt.cpp:241-300), the rescale / mod-down
// constants inv_q_last_mod_q (S/util/rns.cpp:578-787), Galois permutation tables
// (S/util/galois.cpp:18-51) and the CKKS encoder tables (S/ckks.cpp:20-75) — all resident in
// HBM and built eagerly at creation.
#pragma once
#include "modarith.cuh"
#include <cuda_runtime.h>
#include <functional>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace moai
{
    enum Status : int
    {
        OK = 0,
        INVALID_ARGUMENT = 1, // maps to std::invalid_argument in the C++ facade
        LOGIC_ERROR = 2,      // maps to std::logic_error
        CUDA_ERROR = 3,
        OUT_OF_MEMORY = 4,
    };

    struct StatusError
    {
        int code;
        std::string msg;
    };

    void set_last_error(const std::string &msg);
    const std::string &last_error();
    // stream-ordered allocation from a best-fit block cache (context.cu); freed blocks are reusable at
    // once by later work on the same stream.  device_release_cached() returns the cache to the driver.
    void *device_alloc(size_t bytes, cudaStream_t stream);
    void device_free(void *p, cudaStream_t stream);
    void device_release_cached();
    void set_phase(const char *name); // last pipeline phase entered (for allocation-failure reports)
    struct AllocStats
    {
        unsigned long long calls = 0, retries = 0; // retries = allocations that had to flush the cache
        double host_ms = 0;                        // host time spent inside cudaMalloc (cache misses)
        size_t cached_bytes = 0, owned_bytes = 0;
    };
    AllocStats alloc_stats();

#define MOAI_CUDA_CHECK(expr)                                                                                          \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess)                                                                                         \
        {                                                                                                              \
            throw ::moai::StatusError{ _e == cudaErrorMemoryAllocation ? ::moai::OUT_OF_MEMORY : ::moai::CUDA_ERROR,   \
                                       std::string(#expr) + ": " + cudaGetErrorString(_e) };                           \
        }                                                                                                              \
    } while (0)

#define MOAI_REQUIRE(cond, msg)                                                                                        \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            throw ::moai::StatusError{ ::moai::INVALID_ARGUMENT, std::string(msg) };                                   \
        }                                                                                                              \
    } while (0)

    // Twiddle entry: Shoup pair, one 16-byte load in the kernels.
    struct __align__(16) Twiddle
    {
        u64 w;
        u64 wq;
    };

    struct Comm; // comm.hpp: NCCL communicator of a context that shares one packed batch with other GPUs

    struct Context
    {
        int device = 0;
        Comm *comm = nullptr;
        int log_n = 0;
        size_t n = 0;
        int kl = 0; // key-level limb count (data primes + special prime)
        std::vector<u64> q;
        cudaStream_t stream = nullptr;
        int sm_count = 148;

        // device tables
        LimbConst *d_limb = nullptr;   // [kl]
        Twiddle *d_fwd = nullptr;      // [kl][n]  psi powers, bit-reversed order
        Twiddle *d_inv = nullptr;      // [kl][n]  psi^-1 powers, scrambled order
        double *d_fwd_fp = nullptr;    // [kl][n]  the same roots as symmetric doubles (FP64 NTT path)
        double *d_inv_fp = nullptr;    // [kl][n]
        Twiddle *d_inv_last = nullptr; // [kl][kl] inv_last[last][i] = q_last^-1 mod q_i (i != last)
        u64 *d_half_mod = nullptr;     // [kl][kl] half_mod[last][i] = (q_last >> 1) mod q_i
        Twiddle *d_two64 = nullptr;    // [kl] 2^64 mod q_i
        int *d_ids = nullptr;          // [kl] identity limb ids 0..kl-1
        int *d_ids_ks = nullptr;       // [kl][kl+1]: row l = {0..l-1, kl-1} target moduli of a key switch at l limbs

        // CKKS encoder tables
        double2 *d_fft_inv_roots = nullptr; // [n]
        uint32_t *d_index_map = nullptr;    // [n]
        std::vector<double2> h_fft_inv_roots;
        std::vector<uint32_t> h_index_map;

        std::vector<LimbConst> h_limb;

        // bookkeeping for bench.py: kernels launched, and (when profiling) per-phase device time
        unsigned long long launches = 0;
        bool profiling = false;
        std::map<std::string, std::pair<double, long long>> prof; // name -> (ms, count)
        // per-kernel timers (KernelTimer): event pairs recorded on the launching stream and resolved later, so that
        // timing a kernel never stalls the launch queue; prof[name] accumulates (ms, units)
        struct PendingKernel
        {
            const char *name;
            cudaEvent_t e0, e1;
            long long units;
        };
        std::deque<PendingKernel> kpending;
        std::vector<cudaEvent_t> kpool;
        void kernel_timers_collect(bool wait); // fold finished pairs into prof (wait: synchronise first)

        std::mutex galois_mu;
        std::map<uint32_t, uint32_t *> galois_tables; // elt -> device [n]
        // grouped-digit key switching (csrc/ksgroup.cu): conversion tables per (extra primes, level)
        std::mutex ksg_mu;
        std::map<std::pair<int, int>, void *> ksg_cache;

        // A LANE of another context (context_fork): shares every immutable table of `parent` (twiddles, limb constants,
        // encoder tables, Galois permutations, grouped-digit conversion tables) and owns only what is per caller — its
        // CUDA stream (hence its arena: the allocator is keyed by (device, stream)), timers and counters.  One lane per
        // host thread lets threads issue work concurrently (one stream per calling thread).
        Context *parent = nullptr;
        bool owns_stream = false;
        Context *root()
        {
            return parent ? parent : this;
        }

        ~Context();
        const uint32_t *galois_table(uint32_t elt);
        uint32_t elt_from_step(int step) const;
    };

    Context *context_create(int log_n, const u64 *primes, int kl, int device);
    Context *context_fork(Context *parent); // a lane with its own non-blocking stream; must not outlive `parent`
    void ksg_release(Context *c); // csrc/ksgroup.cu

    // RAII device timer around a kernel (only active when Context::profiling is set): records
    // CUDA events on the launching stream and accumulates the elapsed time under `name`.
    struct PhaseTimer
    {
        Context *c;
        const char *name;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        PhaseTimer(Context *ctx, const char *nm) : c(ctx), name(nm)
        {
            set_phase(nm);
            if (c->profiling)
            {
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                cudaEventRecord(e0, c->stream);
            }
        }
        ~PhaseTimer()
        {
            if (e0)
            {
                cudaEventRecord(e1, c->stream);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                auto &slot = c->prof[name];
                slot.first += ms;
                slot.second += 1;
                cudaEventDestroy(e0);
                cudaEventDestroy(e1);
            }
        }
    };


    // RAII device timer around ONE kernel launch (only when Context::profiling): CUDA events on the launching stream,
    // never synchronises; `units` = the kernel's work units (e.g. limb-transforms) for roofline accounting.
    struct KernelTimer
    {
        Context *c;
        Context::PendingKernel p{};
        KernelTimer(Context *ctx, const char *nm, long long units) : c(ctx)
        {
            if (!c->profiling)
            {
                return;
            }
            auto take = [&]() {
                cudaEvent_t e;
                if (c->kpool.empty())
                {
                    cudaEventCreate(&e);
                }
                else
                {
                    e = c->kpool.back();
                    c->kpool.pop_back();
                }
                return e;
            };
            p.name = nm;
            p.units = units;
            p.e0 = take();
            p.e1 = take();
            cudaEventRecord(p.e0, c->stream);
        }
        ~KernelTimer()
        {
            if (p.e0)
            {
                cudaEventRecord(p.e1, c->stream);
                c->kpending.push_back(p);
                if (c->kpending.size() > 65536)
                {
                    c->kernel_timers_collect(false);
                }
            }
        }
    };

    // stream-ordered scratch allocation
    struct Scratch
    {
        void *p = nullptr;
        cudaStream_t s;
        Scratch(size_t bytes, cudaStream_t stream);
        ~Scratch();
        template <class T>
        T *as()
        {
            return reinterpret_cast<T *>(p);
        }
        Scratch(const Scratch &) = delete;
        Scratch &operator=(const Scratch &) = delete;
    };
} // namespace moai
