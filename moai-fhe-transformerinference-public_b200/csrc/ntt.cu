#include "ntt.cuh"

namespace moai
{
    namespace
    {
        constexpr int TB = 16;       // columns per pass-A tile (16 x 8 B = one 128-byte line)
        constexpr int ROWS = 16;     // rows per pass-B CTA
        constexpr int ROW_PAD = 272; // 256 + 16: padded row, conflict-free stride-17 access

        template <int COUNT>
        __device__ __forceinline__ void load_tw(Twiddle (&tw)[8], const Twiddle *__restrict__ src)
        {
#pragma unroll
            for (int j = 0; j < COUNT; j++)
            {
                const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(src + j));
                tw[j].w = v.x;
                tw[j].wq = v.y;
            }
        }

        // ------------------------------------------------------------------ forward, pass A
        template <int LOGR>
        __global__ void __launch_bounds__((1 << LOGR) / 16 * TB) ntt_fwd_pass_a(NttArgs a)
        {
            constexpr int R = 1 << LOGR, T1 = R / 16;
            __shared__ u64 sm[R * TB];
            const int tb = threadIdx.x & (TB - 1), t = threadIdx.x >> 4;
            const long long poly = blockIdx.x / (256 / TB);
            const int tile = blockIdx.x % (256 / TB);
            const int limb = a.limb_ids[(poly / a.div) % a.period];
            const LimbConst lc = a.limb[limb];
            const Twiddle *__restrict__ tw_tab = a.tw + ((size_t)limb << a.log_n);
            u64 *base = a.data + ((size_t)poly << a.log_n) + tile * TB + tb;
            const u64 q = lc.q, two_q = lc.two_q;

            u64 x[16];
            Twiddle tw[8];
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = base[(size_t)(t + T1 * k) * 256];
            }
            // stages 0..3 pair the top four bits of a (k): root index 2^s + block
            load_tw<1>(tw, tw_tab + 1);
            ct_stage<8>(x, tw, q, two_q);
            load_tw<2>(tw, tw_tab + 2);
            ct_stage<4>(x, tw, q, two_q);
            load_tw<4>(tw, tw_tab + 4);
            ct_stage<2>(x, tw, q, two_q);
            load_tw<8>(tw, tw_tab + 8);
            ct_stage<1>(x, tw, q, two_q);
            if constexpr (LOGR > 4)
            {
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    sm[(t + T1 * k) * TB + tb] = x[k];
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = sm[(16 * t + k) * TB + tb];
                }
                // remaining stages: a = 16 t + k, row gap G in {R/32 .. 1}; root index R/(2G) + a/(2G)
                if constexpr (R / 32 >= 8)
                {
                    load_tw<1>(tw, tw_tab + R / 16 + t);
                    ct_stage<8>(x, tw, q, two_q);
                }
                if constexpr (R / 32 >= 4)
                {
                    load_tw<2>(tw, tw_tab + R / 8 + 2 * t);
                    ct_stage<4>(x, tw, q, two_q);
                }
                if constexpr (R / 32 >= 2)
                {
                    load_tw<4>(tw, tw_tab + R / 4 + 4 * t);
                    ct_stage<2>(x, tw, q, two_q);
                }
                load_tw<8>(tw, tw_tab + R / 2 + 8 * t);
                ct_stage<1>(x, tw, q, two_q);
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    base[(size_t)(16 * t + k) * 256] = x[k];
                }
            }
            else
            {
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    base[(size_t)(t + T1 * k) * 256] = x[k];
                }
            }
        }

        // ------------------------------------------------------------------ forward, pass B
        __global__ void __launch_bounds__(ROWS * 16) ntt_fwd_pass_b(NttArgs a)
        {
            __shared__ u64 sm[ROWS * ROW_PAD];
            const int t = threadIdx.x & 15, r = threadIdx.x >> 4;
            const int R = 1 << (a.log_n - 8);
            const int ctas_per_poly = R / ROWS;
            const long long poly = blockIdx.x / ctas_per_poly;
            const int row = (blockIdx.x % ctas_per_poly) * ROWS + r;
            const int limb = a.limb_ids[(poly / a.div) % a.period];
            const LimbConst lc = a.limb[limb];
            const Twiddle *__restrict__ tw_tab = a.tw + ((size_t)limb << a.log_n);
            u64 *base = a.data + ((size_t)poly << a.log_n) + (size_t)row * 256;
            const u64 q = lc.q, two_q = lc.two_q;
            const size_t ra = (size_t)R + row; // root index of stage t' is 2^t' (R + a) + b / (2 gap)

            u64 x[16];
            Twiddle tw[8];
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = base[t + 16 * k];
            }
            load_tw<1>(tw, tw_tab + ra);
            ct_stage<8>(x, tw, q, two_q);
            load_tw<2>(tw, tw_tab + 2 * ra);
            ct_stage<4>(x, tw, q, two_q);
            load_tw<4>(tw, tw_tab + 4 * ra);
            ct_stage<2>(x, tw, q, two_q);
            load_tw<8>(tw, tw_tab + 8 * ra);
            ct_stage<1>(x, tw, q, two_q);
            u64 *srow = sm + r * ROW_PAD;
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                srow[t + 17 * k] = x[k];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = srow[17 * t + k];
            }
            load_tw<1>(tw, tw_tab + 16 * ra + t);
            ct_stage<8>(x, tw, q, two_q);
            load_tw<2>(tw, tw_tab + 32 * ra + 2 * t);
            ct_stage<4>(x, tw, q, two_q);
            load_tw<4>(tw, tw_tab + 64 * ra + 4 * t);
            ct_stage<2>(x, tw, q, two_q);
            load_tw<8>(tw, tw_tab + 128 * ra + 8 * t);
            ct_stage<1>(x, tw, q, two_q);
            // canonical residues out: [0, 4q) -> [0, q)   (S/util/ntt.cpp:425-434)
            ulonglong2 *out = reinterpret_cast<ulonglong2 *>(base + 16 * t);
#pragma unroll
            for (int k = 0; k < 16; k += 2)
            {
                ulonglong2 v;
                v.x = csub(csub(x[k], two_q), q);
                v.y = csub(csub(x[k + 1], two_q), q);
                out[k >> 1] = v;
            }
        }

        // ------------------------------------------------------------------ inverse, pass B'
        __global__ void __launch_bounds__(ROWS * 16) ntt_inv_pass_b(NttArgs a)
        {
            __shared__ u64 sm[ROWS * ROW_PAD];
            const int t = threadIdx.x & 15, r = threadIdx.x >> 4;
            const int R = 1 << (a.log_n - 8);
            const size_t n = (size_t)1 << a.log_n;
            const int ctas_per_poly = R / ROWS;
            const long long poly = blockIdx.x / ctas_per_poly;
            const int row = (blockIdx.x % ctas_per_poly) * ROWS + r;
            const int limb = a.limb_ids[(poly / a.div) % a.period];
            const LimbConst lc = a.limb[limb];
            const Twiddle *__restrict__ tw_tab = a.tw + ((size_t)limb << a.log_n);
            u64 *base = a.data + ((size_t)poly << a.log_n) + (size_t)row * 256;
            const u64 q = lc.q, two_q = lc.two_q;

            u64 x[16];
            Twiddle tw[8];
            const ulonglong2 *in = reinterpret_cast<const ulonglong2 *>(base + 16 * t);
#pragma unroll
            for (int k = 0; k < 16; k += 2)
            {
                ulonglong2 v = in[k >> 1];
                x[k] = v.x;
                x[k + 1] = v.y;
            }
            // stage with gap g: root index n - n/g + 1 + row*(128/g) + b/(2g); here b = 16 t + k
            load_tw<8>(tw, tw_tab + (n - n + 1) + (size_t)row * 128 + 8 * t);
            gs_stage<1>(x, tw, q, two_q);
            load_tw<4>(tw, tw_tab + (n - n / 2 + 1) + (size_t)row * 64 + 4 * t);
            gs_stage<2>(x, tw, q, two_q);
            load_tw<2>(tw, tw_tab + (n - n / 4 + 1) + (size_t)row * 32 + 2 * t);
            gs_stage<4>(x, tw, q, two_q);
            load_tw<1>(tw, tw_tab + (n - n / 8 + 1) + (size_t)row * 16 + t);
            gs_stage<8>(x, tw, q, two_q);
            u64 *srow = sm + r * ROW_PAD;
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                srow[17 * t + k] = x[k];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = srow[t + 17 * k];
            }
            // b = t + 16 k ; gaps 16, 32, 64, 128
            load_tw<8>(tw, tw_tab + (n - n / 16 + 1) + (size_t)row * 8);
            gs_stage<1>(x, tw, q, two_q);
            load_tw<4>(tw, tw_tab + (n - n / 32 + 1) + (size_t)row * 4);
            gs_stage<2>(x, tw, q, two_q);
            load_tw<2>(tw, tw_tab + (n - n / 64 + 1) + (size_t)row * 2);
            gs_stage<4>(x, tw, q, two_q);
            load_tw<1>(tw, tw_tab + (n - n / 128 + 1) + (size_t)row);
            gs_stage<8>(x, tw, q, two_q);
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                base[t + 16 * k] = x[k];
            }
        }

        // ------------------------------------------------------------------ inverse, pass A'
        template <int LOGR>
        __global__ void __launch_bounds__((1 << LOGR) / 16 * TB) ntt_inv_pass_a(NttArgs a)
        {
            constexpr int R = 1 << LOGR, T1 = R / 16;
            __shared__ u64 sm[R * TB];
            const int tb = threadIdx.x & (TB - 1), t = threadIdx.x >> 4;
            const size_t n = (size_t)1 << a.log_n;
            const long long poly = blockIdx.x / (256 / TB);
            const int tile = blockIdx.x % (256 / TB);
            const int limb = a.limb_ids[(poly / a.div) % a.period];
            const LimbConst lc = a.limb[limb];
            const Twiddle *__restrict__ tw_tab = a.tw + ((size_t)limb << a.log_n);
            u64 *base = a.data + ((size_t)poly << a.log_n) + tile * TB + tb;
            const u64 q = lc.q, two_q = lc.two_q;

            u64 x[16];
            Twiddle tw[8];
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = base[(size_t)(16 * t + k) * 256];
            }
            // row gap G: root index n - R/G + 1 + a/(2G), a = 16 t + k
            load_tw<8>(tw, tw_tab + (n - R + 1) + 8 * t);
            gs_stage<1>(x, tw, q, two_q);
            load_tw<4>(tw, tw_tab + (n - R / 2 + 1) + 4 * t);
            gs_stage<2>(x, tw, q, two_q);
            load_tw<2>(tw, tw_tab + (n - R / 4 + 1) + 2 * t);
            gs_stage<4>(x, tw, q, two_q);
            if constexpr (LOGR == 4)
            {
                gs_stage_last(x, lc);
            }
            else
            {
                load_tw<1>(tw, tw_tab + (n - R / 8 + 1) + t);
                gs_stage<8>(x, tw, q, two_q);
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    sm[(16 * t + k) * TB + tb] = x[k];
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = sm[(t + T1 * k) * TB + tb];
                }
                // a = t + T1 k ; row gaps G = T1 * kgap for kgap >= 256 / R
                if constexpr (256 / R <= 1)
                {
                    load_tw<8>(tw, tw_tab + (n - R / T1 + 1));
                    gs_stage<1>(x, tw, q, two_q);
                }
                if constexpr (256 / R <= 2)
                {
                    load_tw<4>(tw, tw_tab + (n - R / (2 * T1) + 1));
                    gs_stage<2>(x, tw, q, two_q);
                }
                if constexpr (256 / R <= 4)
                {
                    load_tw<2>(tw, tw_tab + (n - R / (4 * T1) + 1));
                    gs_stage<4>(x, tw, q, two_q);
                }
                gs_stage_last(x, lc);
            }
            // canonical residues out: [0, 2q) -> [0, q)   (S/util/ntt.cpp:466-472)
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                const int row = (LOGR == 4) ? (16 * t + k) : (t + T1 * k);
                base[(size_t)row * 256] = csub(x[k], q);
            }
        }

        template <int LOGR>
        void launch_fwd(const NttArgs &a, cudaStream_t s)
        {
            const long long ctas_a = a.count * (256 / TB);
            ntt_fwd_pass_a<LOGR><<<(unsigned)ctas_a, (1 << LOGR) / 16 * TB, 0, s>>>(a);
            const long long ctas_b = a.count * ((1 << LOGR) / ROWS);
            ntt_fwd_pass_b<<<(unsigned)ctas_b, ROWS * 16, 0, s>>>(a);
        }

        template <int LOGR>
        void launch_inv(const NttArgs &a, cudaStream_t s)
        {
            const long long ctas_b = a.count * ((1 << LOGR) / ROWS);
            ntt_inv_pass_b<<<(unsigned)ctas_b, ROWS * 16, 0, s>>>(a);
            const long long ctas_a = a.count * (256 / TB);
            ntt_inv_pass_a<LOGR><<<(unsigned)ctas_a, (1 << LOGR) / 16 * TB, 0, s>>>(a);
        }
    } // namespace

    void ntt_forward(Context *c, u64 *data, long long count, const int *d_limb_ids, int period, int div)
    {
        if (count <= 0)
        {
            return;
        }
        NttArgs a{ data, c->d_fwd, c->d_limb, d_limb_ids, period, div, c->log_n, count };
        switch (c->log_n)
        {
        case 12: launch_fwd<4>(a, c->stream); break;
        case 13: launch_fwd<5>(a, c->stream); break;
        case 14: launch_fwd<6>(a, c->stream); break;
        case 15: launch_fwd<7>(a, c->stream); break;
        case 16: launch_fwd<8>(a, c->stream); break;
        default: throw StatusError{ INVALID_ARGUMENT, "unsupported log_n" };
        }
        c->launches += 2;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ntt_inverse(Context *c, u64 *data, long long count, const int *d_limb_ids, int period, int div)
    {
        if (count <= 0)
        {
            return;
        }
        NttArgs a{ data, c->d_inv, c->d_limb, d_limb_ids, period, div, c->log_n, count };
        switch (c->log_n)
        {
        case 12: launch_inv<4>(a, c->stream); break;
        case 13: launch_inv<5>(a, c->stream); break;
        case 14: launch_inv<6>(a, c->stream); break;
        case 15: launch_inv<7>(a, c->stream); break;
        case 16: launch_inv<8>(a, c->stream); break;
        default: throw StatusError{ INVALID_ARGUMENT, "unsupported log_n" };
        }
        c->launches += 2;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }
} // namespace moai
