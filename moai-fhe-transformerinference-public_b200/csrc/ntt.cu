#include <algorithm>
#include "ntt.cuh"
#include "fpfield.cuh"
#include <cstdlib>

namespace moai
{
    namespace
    {
        constexpr int TB = 16;       // columns per pass-A tile (16 x 8 B = one 128-byte line)
        constexpr int ROWS = 16;     // rows per pass-B CTA
        constexpr int ROW_PAD = 272; // 256 + 16: padded row, conflict-free stride-17 access

        // =====================================================================================
        // Arithmetic policies.  A policy owns the per-limb constants and defines the element type
        // kept in registers, how elements enter/leave global memory at the outer boundary
        // (canonical uint64) and between the two passes (policy-private), and the butterflies.
        // =====================================================================================

        // ---- integer Harvey / Shoup (any prime < 2^61): values lazy in [0, 4q) forward, [0, 2q) inverse
        struct IntField
        {
            typedef u64 elem;
            typedef Twiddle tw_t;
            u64 q, two_q;
            u64 inv_n, inv_n_quo, inv_n_w, inv_n_w_quo;
            u64 bar_m;
            u32 bar_shift;
            const Twiddle *__restrict__ tab;

            __device__ IntField(const NttArgs &a, int limb, const LimbConst &lc)
                : q(lc.q), two_q(lc.two_q), inv_n(lc.inv_n), inv_n_quo(lc.inv_n_quo), inv_n_w(lc.inv_n_w),
                  inv_n_w_quo(lc.inv_n_w_quo), bar_m(lc.bar_m), bar_shift(lc.bar_shift),
                  tab(a.tw + ((size_t)limb << a.log_n))
            {}
            __device__ __forceinline__ elem pro_reduce(elem v) const
            {
                const u64 t = __umul64hi(v >> bar_shift, bar_m);
                return csub(csub(v - t * q, two_q), q);
            }
            __device__ __forceinline__ tw_t tw(size_t idx) const
            {
                const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(tab + idx));
                return Twiddle{ v.x, v.y };
            }
            __device__ __forceinline__ elem in_outer(u64 v) const { return v; }
            __device__ __forceinline__ elem in_mid(u64 v) const { return v; }
            __device__ __forceinline__ u64 out_mid(elem x) const { return x; }
            __device__ __forceinline__ u64 out_fwd(elem x) const { return csub(csub(x, two_q), q); } // [0,4q)->[0,q)
            __device__ __forceinline__ u64 out_inv(elem x) const { return csub(x, q); }              // [0,2q)->[0,q)
            __device__ __forceinline__ u64 out_fp(elem x) const { return csub(x, q); } // integer-path limbs stay residues
            __device__ __forceinline__ void phase_begin_fwd(elem (&)[16]) const {}
            __device__ __forceinline__ void phase_mid_fwd(elem (&)[16]) const {}
            __device__ __forceinline__ void phase_begin_inv(elem (&)[16]) const {}
            __device__ __forceinline__ void ct(elem &x, elem &y, const tw_t &w) const
            {
                const u64 u = csub(x, two_q);
                const u64 v = mul_shoup_lazy(y, w.w, w.wq, q);
                x = u + v;
                y = u + two_q - v;
            }
            __device__ __forceinline__ void gs(elem &x, elem &y, const tw_t &w) const
            {
                const u64 u = x, v = y;
                x = csub(u + v, two_q);
                y = mul_shoup_lazy(u + two_q - v, w.w, w.wq, q);
            }
            __device__ __forceinline__ void gs_last(elem &x, elem &y) const
            {
                const u64 u = x, v = y;
                x = mul_shoup_lazy(csub(u + v, two_q), inv_n, inv_n_quo, q);
                y = mul_shoup_lazy(u + two_q - v, inv_n_w, inv_n_w_quo, q);
            }
        };

        // ---- register stages over 16 elements; tw index of the j-th block of 2*GAP is base + j
        template <int GAP, class F>
        __device__ __forceinline__ void ct_stage_tw(const F &f, typename F::elem (&x)[16],
                                                    const typename F::tw_t (&tw)[8 / GAP]);
        template <int GAP, class F>
        __device__ __forceinline__ void gs_stage_tw(const F &f, typename F::elem (&x)[16],
                                                    const typename F::tw_t (&tw)[8 / GAP]);

        template <int GAP, class F>
        __device__ __forceinline__ void ct_stage(const F &f, typename F::elem (&x)[16], size_t tw_base)
        {
            typename F::tw_t tw[8 / GAP];
#pragma unroll
            for (int j = 0; j < 8 / GAP; j++)
            {
                tw[j] = f.tw(tw_base + j);
            }
            ct_stage_tw<GAP>(f, x, tw);
        }

        // same stage with the twiddles read through a pointer (shared-memory staging in pass B)
        template <int GAP, class F>
        __device__ __forceinline__ void ct_stage_p(const F &f, typename F::elem (&x)[16], const typename F::tw_t *ptr)
        {
            typename F::tw_t tw[8 / GAP];
#pragma unroll
            for (int j = 0; j < 8 / GAP; j++)
            {
                tw[j] = ptr[j];
            }
            ct_stage_tw<GAP>(f, x, tw);
        }

        template <int GAP, class F>
        __device__ __forceinline__ void ct_stage_tw(const F &f, typename F::elem (&x)[16],
                                                    const typename F::tw_t (&tw)[8 / GAP])
        {
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                if (!(k & GAP))
                {
                    f.ct(x[k], x[k + GAP], tw[k / (2 * GAP)]);
                }
            }
        }

        template <int GAP, class F>
        __device__ __forceinline__ void gs_stage(const F &f, typename F::elem (&x)[16], size_t tw_base)
        {
            typename F::tw_t tw[8 / GAP];
#pragma unroll
            for (int j = 0; j < 8 / GAP; j++)
            {
                tw[j] = f.tw(tw_base + j);
            }
            gs_stage_tw<GAP>(f, x, tw);
        }

        template <int GAP, class F>
        __device__ __forceinline__ void gs_stage_p(const F &f, typename F::elem (&x)[16], const typename F::tw_t *ptr)
        {
            typename F::tw_t tw[8 / GAP];
#pragma unroll
            for (int j = 0; j < 8 / GAP; j++)
            {
                tw[j] = ptr[j];
            }
            gs_stage_tw<GAP>(f, x, tw);
        }

        template <int GAP, class F>
        __device__ __forceinline__ void gs_stage_tw(const F &f, typename F::elem (&x)[16],
                                                    const typename F::tw_t (&tw)[8 / GAP])
        {
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                if (!(k & GAP))
                {
                    f.gs(x[k], x[k + GAP], tw[k / (2 * GAP)]);
                }
            }
        }

        template <class F>
        __device__ __forceinline__ void gs_stage_last(const F &f, typename F::elem (&x)[16])
        {
#pragma unroll
            for (int k = 0; k < 8; k++)
            {
                f.gs_last(x[k], x[k + 8]);
            }
        }

        // =====================================================================================
        // Pass bodies (policy-generic).  `sm` is the CTA's shared buffer viewed as 8-byte words.
        // =====================================================================================
        struct ProArgs
        {
            const u64 *src; // source polynomial + tile offset (== base when there is no prologue)
            int mode;
            u64 q_last, half, fix;
            LimbConst lc;   // target limb (mode 2 integer reduction)
        };

        template <int LOGR, class F>
        __device__ __forceinline__ void fwd_pass_a_stages(const F &f, typename F::elem (&x)[16], u64 *base, u64 *sm, int t,
                                                          int tb);

        template <int LOGR, class F>
        __device__ __forceinline__ void fwd_pass_a_body(const F &f, u64 *base, u64 *sm, int t, int tb, const ProArgs &pa)
        {
            constexpr int R = 1 << LOGR, T1 = R / 16;
            typename F::elem x[16];
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                u64 v = pa.src[(size_t)(t + T1 * k) * 256];
                if (pa.mode == 2)
                {
                    v = reduce64(addmod(v, pa.half, pa.q_last), pa.lc) + pa.fix; // < 2 q_i
                }
                x[k] = f.in_outer(v);
                if (pa.mode == 1)
                {
                    x[k] = f.pro_reduce(x[k]);
                }
            }
            fwd_pass_a_stages<LOGR>(f, x, base, sm, t, tb);
        }

        template <int LOGR, class F>
        __device__ __forceinline__ void fwd_pass_a_stages(const F &f, typename F::elem (&x)[16], u64 *base, u64 *sm, int t,
                                                          int tb)
        {
            constexpr int R = 1 << LOGR, T1 = R / 16;
            typename F::elem *smf = reinterpret_cast<typename F::elem *>(sm);
            f.phase_begin_fwd(x);
            // stages 0..3 pair the top four bits of a (k): root index 2^s + block
            ct_stage<8>(f, x, 1);
            ct_stage<4>(f, x, 2);
            f.phase_mid_fwd(x);
            ct_stage<2>(f, x, 4);
            ct_stage<1>(f, x, 8);
            if constexpr (LOGR > 4)
            {
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    smf[(t + T1 * k) * TB + tb] = x[k];
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = smf[(16 * t + k) * TB + tb];
                }
                f.phase_begin_fwd(x);
                // remaining stages: a = 16 t + k, row gap G in {R/32 .. 1}; root index R/(2G) + a/(2G)
                if constexpr (R / 32 >= 8)
                {
                    ct_stage<8>(f, x, R / 16 + t);
                }
                if constexpr (R / 32 >= 4)
                {
                    ct_stage<4>(f, x, R / 8 + 2 * t);
                }
                if constexpr (R / 32 >= 8)
                {
                    f.phase_mid_fwd(x); // two stages done
                }
                if constexpr (R / 32 >= 2)
                {
                    ct_stage<2>(f, x, R / 4 + 4 * t);
                }
                if constexpr (R / 32 == 4)
                {
                    f.phase_mid_fwd(x); // stages <4>, <2> done
                }
                ct_stage<1>(f, x, R / 2 + 8 * t);
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    base[(size_t)(16 * t + k) * 256] = f.out_mid(x[k]);
                }
            }
            else
            {
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    base[(size_t)(t + T1 * k) * 256] = f.out_mid(x[k]);
                }
            }
        }

        template <class F>
        __device__ __forceinline__ void fwd_pass_b_body(const F &f, u64 *base, u64 *srow64, int t, size_t ra)
        {
            typename F::elem x[16];
            typename F::elem *srow = reinterpret_cast<typename F::elem *>(srow64);
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = f.in_mid(base[t + 16 * k]);
            }
            f.phase_begin_fwd(x);
            // root index of stage t' is 2^t' (R + a) + b / (2 gap)
            ct_stage<8>(f, x, ra);
            ct_stage<4>(f, x, 2 * ra);
            f.phase_mid_fwd(x);
            ct_stage<2>(f, x, 4 * ra);
            ct_stage<1>(f, x, 8 * ra);
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
            f.phase_begin_fwd(x);
            ct_stage<8>(f, x, 16 * ra + t);
            ct_stage<4>(f, x, 32 * ra + 2 * t);
            f.phase_mid_fwd(x);
            ct_stage<2>(f, x, 64 * ra + 4 * t);
            ct_stage<1>(f, x, 128 * ra + 8 * t);
            ulonglong2 *out = reinterpret_cast<ulonglong2 *>(base + 16 * t);
#pragma unroll
            for (int k = 0; k < 16; k += 2)
            {
                ulonglong2 v;
                v.x = f.out_fwd(x[k]);
                v.y = f.out_fwd(x[k + 1]);
                out[k >> 1] = v;
            }
        }

        template <class F>
        __device__ __forceinline__ void inv_pass_b_body(const F &f, u64 *base, const u64 *in_base, u64 *srow64, int t,
                                                        size_t row, size_t n)
        {
            typename F::elem x[16];
            typename F::elem *srow = reinterpret_cast<typename F::elem *>(srow64);
            const ulonglong2 *in = reinterpret_cast<const ulonglong2 *>(in_base + 16 * t);
#pragma unroll
            for (int k = 0; k < 16; k += 2)
            {
                const ulonglong2 v = in[k >> 1];
                x[k] = f.in_outer(v.x);
                x[k + 1] = f.in_outer(v.y);
            }
            // stage with gap g: root index n - n/g + 1 + row*(128/g) + b/(2g); here b = 16 t + k
            gs_stage<1>(f, x, (n - n + 1) + row * 128 + 8 * t);
            gs_stage<2>(f, x, (n - n / 2 + 1) + row * 64 + 4 * t);
            gs_stage<4>(f, x, (n - n / 4 + 1) + row * 32 + 2 * t);
            gs_stage<8>(f, x, (n - n / 8 + 1) + row * 16 + t);
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
            f.phase_begin_inv(x);
            // b = t + 16 k ; gaps 16, 32, 64, 128
            gs_stage<1>(f, x, (n - n / 16 + 1) + row * 8);
            gs_stage<2>(f, x, (n - n / 32 + 1) + row * 4);
            gs_stage<4>(f, x, (n - n / 64 + 1) + row * 2);
            gs_stage<8>(f, x, (n - n / 128 + 1) + row);
            f.phase_begin_inv(x);
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                base[t + 16 * k] = f.out_mid(x[k]);
            }
        }

        template <int LOGR, class F>
        __device__ __forceinline__ void inv_pass_a_body(const F &f, u64 *base, u64 *sm, int t, int tb, size_t n,
                                                        bool fp_out = false)
        {
            constexpr int R = 1 << LOGR, T1 = R / 16;
            typename F::elem x[16];
            typename F::elem *smf = reinterpret_cast<typename F::elem *>(sm);
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = f.in_mid(base[(size_t)(16 * t + k) * 256]);
            }
            // row gap G: root index n - R/G + 1 + a/(2G), a = 16 t + k
            gs_stage<1>(f, x, (n - R + 1) + 8 * t);
            gs_stage<2>(f, x, (n - R / 2 + 1) + 4 * t);
            gs_stage<4>(f, x, (n - R / 4 + 1) + 2 * t);
            if constexpr (LOGR == 4)
            {
                gs_stage_last(f, x);
            }
            else
            {
                gs_stage<8>(f, x, (n - R / 8 + 1) + t);
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    smf[(16 * t + k) * TB + tb] = x[k];
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = smf[(t + T1 * k) * TB + tb];
                }
                f.phase_begin_inv(x);
                // a = t + T1 k ; row gaps G = T1 * kgap for kgap >= 256 / R
                if constexpr (256 / R <= 1)
                {
                    gs_stage<1>(f, x, n - R / T1 + 1);
                }
                if constexpr (256 / R <= 2)
                {
                    gs_stage<2>(f, x, n - R / (2 * T1) + 1);
                }
                if constexpr (256 / R <= 4)
                {
                    gs_stage<4>(f, x, n - R / (4 * T1) + 1);
                }
                gs_stage_last(f, x);
            }
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                const int row = (LOGR == 4) ? (16 * t + k) : (t + T1 * k);
                base[(size_t)row * 256] = fp_out ? f.out_fp(x[k]) : f.out_inv(x[k]);
            }
        }

        // =====================================================================================
        // Kernels: resolve (poly, tile) -> limb, pick the arithmetic path of that limb.
        // =====================================================================================
#define MOAI_DISPATCH_FIELD(lc, CALL)                                                                                  \
    if ((lc).fp_class == 1)                                                                                            \
    {                                                                                                                  \
        const FpField<false> f(a, limb, lc);                                                                           \
        CALL;                                                                                                          \
    }                                                                                                                  \
    else if ((lc).fp_class == 2)                                                                                       \
    {                                                                                                                  \
        const FpField<true> f(a, limb, lc);                                                                            \
        CALL;                                                                                                          \
    }                                                                                                                  \
    else                                                                                                               \
    {                                                                                                                  \
        const IntField f(a, limb, lc);                                                                                 \
        CALL;                                                                                                          \
    }

        template <int LOGR>
        __global__ void __launch_bounds__((1 << LOGR) / 16 * TB) ntt_fwd_pass_a(NttArgs a)
        {
            __shared__ u64 sm[(1 << LOGR) * TB];
            const int tb = threadIdx.x & (TB - 1), t = threadIdx.x >> 4;
            const long long poly = blockIdx.x / (256 / TB);
            const int tile = blockIdx.x % (256 / TB);
            const int limb = a.limb_ids[(poly / a.div) % a.period];
            const LimbConst lc = a.limb[limb];
            u64 *base = a.data + ((size_t)poly << a.log_n) + tile * TB + tb;
            ProArgs pa;
            pa.src = base;
            pa.mode = a.src_mode;
            if (a.src_mode == 1)
            {
                const long long sp = (poly / ((long long)a.period * a.div)) * a.div + poly % a.div;
                pa.src = a.src + ((size_t)sp << a.log_n) + tile * TB + tb;
            }
            else if (a.src_mode == 2)
            {
                const long long sp = poly / a.period;
                pa.src = a.src + ((size_t)sp << a.log_n) + tile * TB + tb;
                pa.q_last = a.limb[a.last_id].q;
                pa.half = pa.q_last >> 1;
                pa.fix = lc.q - a.half_mod[(size_t)a.last_id * a.kl + limb];
                pa.lc = lc;
            }
            MOAI_DISPATCH_FIELD(lc, (fwd_pass_a_body<LOGR>(f, base, sm, t, tb, pa)))
        }

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
            const long long phys = a.grp_size ? (poly / a.grp_size) * a.grp_stride + poly % a.grp_size : poly;
            u64 *base = a.data + ((size_t)phys << a.log_n) + (size_t)row * 256;
            const size_t ra = (size_t)R + row;
            MOAI_DISPATCH_FIELD(lc, (fwd_pass_b_body(f, base, sm + r * ROW_PAD, t, ra)))
        }

        __global__ void __launch_bounds__(ROWS * 16) ntt_inv_pass_b(NttArgs a)
        {
            __shared__ u64 sm[ROWS * ROW_PAD];
            const int t = threadIdx.x & 15, r = threadIdx.x >> 4;
            const int R = 1 << (a.log_n - 8);
            const size_t n = (size_t)1 << a.log_n;
            const int ctas_per_poly = R / ROWS;
            const long long poly = a.p_base + blockIdx.x / ctas_per_poly;
            const int row = (blockIdx.x % ctas_per_poly) * ROWS + r;
            const int limb = a.limb_ids[(poly / a.div) % a.period];
            const LimbConst lc = a.limb[limb];
            u64 *base = a.data + ((size_t)poly << a.log_n) + (size_t)row * 256;
            // optional out-of-place first pass: polynomial p is read from runs of grp_size polynomials grp_stride apart
            const u64 *in_base = base;
            if (a.src)
            {
                const long long phys = (poly / a.grp_size) * a.grp_stride + poly % a.grp_size;
                in_base = a.src + ((size_t)phys << a.log_n) + (size_t)row * 256;
            }
            MOAI_DISPATCH_FIELD(lc, (inv_pass_b_body(f, base, in_base, sm + r * ROW_PAD, t, (size_t)row, n)))
        }

        template <int LOGR>
        __global__ void __launch_bounds__((1 << LOGR) / 16 * TB) ntt_inv_pass_a(NttArgs a)
        {
            __shared__ u64 sm[(1 << LOGR) * TB];
            const int tb = threadIdx.x & (TB - 1), t = threadIdx.x >> 4;
            const size_t n = (size_t)1 << a.log_n;
            const long long poly = a.p_base + blockIdx.x / (256 / TB);
            const int tile = blockIdx.x % (256 / TB);
            const int slot = (int)((poly / a.div) % a.period);
            const int limb = a.limb_ids[slot];
            LimbConst lc = a.limb[limb];
            if (a.scale) // a per-limb constant folded into the N^-1 of the last stage
            {
                const NttScale sc = a.scale[slot];
                lc.inv_n = sc.inv_n;
                lc.inv_n_quo = sc.inv_n_quo;
                lc.inv_n_w = sc.inv_n_w;
                lc.inv_n_w_quo = sc.inv_n_w_quo;
                lc.inv_n_d = sc.inv_n_d;
                lc.inv_n_w_d = sc.inv_n_w_d;
            }
            u64 *base = a.data + ((size_t)poly << a.log_n) + tile * TB + tb;
            MOAI_DISPATCH_FIELD(lc, (inv_pass_a_body<LOGR>(f, base, sm, t, tb, n, a.fp_out != 0)))
        }


        // =====================================================================================
        // Pass A with the fast base conversion of ConvTab (ntt.cuh) as its prologue: the 16 residues a thread
        // starts from are computed from up to CONV_MAX source limbs instead of being loaded.  FP64-path targets
        // use the exact FP64 products of the NTT (terms <= 1.125 p, reduced every 2 (51-bit class) / 8 (46-bit
        // class) terms, so sums stay below 2^53); integer-path targets accumulate 128-bit sums (<= 16 products of
        // 61-bit factors) and reduce once.  v = rint(sum y_j / q_j) is computed with the same FP64 instruction
        // sequence in both versions, so every target modulus sees the same integer digit.
        // =====================================================================================
        // residue -> double for the quotient estimate: exact (and the same value in every instantiation) — the
        // mantissa trick below 2^52 (one DADD instead of an I2F), the conversion instruction for wider sources
        __device__ __forceinline__ double conv_y_double(u64 y, bool wide)
        {
            return wide ? __ull2double_rn(y)
                        : __dadd_rn(__longlong_as_double((long long)(y | 0x4330000000000000ull)), -4503599627370496.0);
        }
        __device__ __forceinline__ double conv_rint(double v)
        {
            const double M = 6755399441055744.0;
            return __dadd_rn(__dadd_rn(v, M), -M);
        }

        template <bool WIDE>
        __device__ __forceinline__ void conv_prologue(const FpField<WIDE> &f, const NttArgs &a, double (&x)[16],
                                                      const u64 *src0, size_t row_stride, int g, int I, int rns)
        {
            const ConvTab &cv = a.conv;
            const int s0 = cv.s0[g], cnt = cv.cnt[g];
            const size_t trow = ((size_t)g * rns + I) * CONV_MAX;
            double vs[16];
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = 0.0;
                vs[k] = 0.0;
            }
            const int red_every = WIDE ? 2 : 8;
            int since = 0;
            const bool pre_v = a.conv_v != nullptr; // quotients precomputed (conv_quotient_fp)
            for (int j = 0; j < cnt; j++)
            {
                const double iq = cv.invq[s0 + j];
                const double bd = cv.Bd[trow + j];
                const u64 *sj = src0 + ((size_t)j << a.log_n);
                u64 y[16];
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    y[k] = sj[(size_t)k * row_stride];
                }
                if (cv.fpsrc && cv.fpsrc[s0 + j])
                {
                    // centred doubles straight from the inverse transform: 7 FP64 operations per (source, target)
#pragma unroll
                    for (int k = 0; k < 16; k++)
                    {
                        const double yd = __longlong_as_double((long long)y[k]);
                        if (!pre_v)
                        {
                            vs[k] = __fma_rn(yd, iq, vs[k]);
                        }
                        x[k] = __dadd_rn(x[k], f.mul_lazy(yd, bd));
                    }
                    since += 1;
                }
                else if (cv.wide[s0 + j])
                {
                    const double b26 = cv.B26d[trow + j];
#pragma unroll
                    for (int k = 0; k < 16; k++)
                    {
                        if (!pre_v)
                        {
                            vs[k] = __fma_rn(conv_y_double(y[k], true), iq, vs[k]);
                        }
                        const double hi = f.in_outer(y[k] >> 26), lo = f.in_outer(y[k] & 0x3FFFFFFull);
                        x[k] = __dadd_rn(x[k], __dadd_rn(f.mul_lazy(hi, b26), f.mul_lazy(lo, bd)));
                    }
                    since += 2;
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < 16; k++)
                    {
                        const double yd = conv_y_double(y[k], false); // == f.in_outer(y[k])
                        vs[k] = __fma_rn(yd, iq, vs[k]);
                        x[k] = __dadd_rn(x[k], f.mul_lazy(yd, bd));
                    }
                    since += 1;
                }
                if (since >= red_every)
                {
#pragma unroll
                    for (int k = 0; k < 16; k++)
                    {
                        x[k] = f.red(x[k]);
                    }
                    since = 0;
                }
            }
            const double nq = cv.negQd[(size_t)g * rns + I];
            if (pre_v)
            {
                // v[b][digits][n]: item b and the element offset follow from src0 = a.src + ((b src_limbs + s0) << log_n) + off
                const size_t diff = (size_t)(src0 - a.src);
                const size_t bq = (diff >> a.log_n) / (size_t)cv.src_limbs, eoff = diff & (((size_t)1 << a.log_n) - 1);
                const double *vp = a.conv_v + (((bq * (size_t)a.div + (size_t)g) << a.log_n) + eoff);
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    vs[k] = vp[(size_t)k * row_stride];
                }
            }
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = f.red(__dadd_rn(x[k], f.mul_lazy(pre_v ? vs[k] : conv_rint(vs[k]), nq)));
            }
        }

        __device__ __forceinline__ void conv_prologue(const IntField &f, const NttArgs &a, u64 (&x)[16], const u64 *src0,
                                                      size_t row_stride, int g, int I, int rns, const LimbConst &lc,
                                                      const Twiddle &t64)
        {
            const ConvTab &cv = a.conv;
            const int s0 = cv.s0[g], cnt = cv.cnt[g];
            const size_t trow = ((size_t)g * rns + I) * CONV_MAX;
            double vs[16];
            u128 acc[16];
            const bool pre_v = a.conv_v != nullptr;
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                acc[k].lo = 0;
                acc[k].hi = 0;
                vs[k] = 0.0;
            }
            for (int j = 0; j < cnt; j++)
            {
                const double iq = cv.invq[s0 + j];
                const u64 bj = cv.B[trow + j];
                const bool wide = cv.wide[s0 + j] != 0;
                const u64 *sj = src0 + ((size_t)j << a.log_n);
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    u64 y = sj[(size_t)k * row_stride];
                    if (cv.BT) // conv_quotient() left v in byte 7 of the first source
                    {
                        if (j == 0)
                        {
                            vs[k] = (double)(y >> 56);
                            y &= 0x00FFFFFFFFFFFFFFull;
                        }
                    }
                    else if (cv.fpsrc && cv.fpsrc[s0 + j])
                    {
                        // a centred double: its canonical residue is y + q for negative y, and every such wrap adds
                        // one Q_g to the sum, i.e. one to the quotient (counted in vs)
                        const double yd = __longlong_as_double((long long)y);
                        long long yi = __double2ll_rn(yd);
                        if (!pre_v)
                        {
                            vs[k] = __fma_rn(yd, iq, vs[k]);
                        }
                        if (yi < 0)
                        {
                            yi += (long long)cv.srcq[s0 + j];
                            vs[k] = __dadd_rn(vs[k], 1.0);
                        }
                        y = (u64)yi;
                    }
                    else if (!pre_v)
                    {
                        vs[k] = __fma_rn(conv_y_double(y, wide), iq, vs[k]);
                    }
                    mac_wide(acc[k], y, bj);
                }
            }
            const u64 nq = cv.negQ[(size_t)g * rns + I];
            if (pre_v)
            {
                const size_t diff = (size_t)(src0 - a.src);
                const size_t bq = (diff >> a.log_n) / (size_t)cv.src_limbs, eoff = diff & (((size_t)1 << a.log_n) - 1);
                const double *vp = a.conv_v + (((bq * (size_t)a.div + (size_t)g) << a.log_n) + eoff);
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    vs[k] = __dadd_rn(vs[k], vp[(size_t)k * row_stride]);
                }
            }
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                // canonical sources: the quotient of the canonical sum = the centred one + the number of wrapped sources
                // (vs holds both)
                const u64 v = (u64)__double2ll_rn(conv_rint(vs[k]));
                mac_wide(acc[k], v, nq);
                x[k] = barrett_reduce_acc(acc[k], lc, t64.w, t64.wq);
            }
            (void)f;
        }

        __device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
        {
            const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
        }
        __device__ __forceinline__ void cp_async_commit()
        {
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        __device__ __forceinline__ void cp_async_wait_1()
        {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        }

        // ---- the conversion on the tensor cores (ConvTab::BT, ntt.cuh) ----------------------------------------------
        __device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
        {
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
        }

        // CTA tile = R rows x 16 columns; one m16 tile = the 16 columns of one row (mma row g <-> column 2g, row g + 8 <->
        // column 2g + 1, and lane c of a quad holds source c of the k-step: one 16-byte shared-memory read is a whole A
        // fragment).  Four phases of R/4 rows; per phase and k-step (four sources) every thread issues ONE 128-byte bulk
        // copy (cp.async.bulk -> mbarrier; a row of a source) into a ring of CONV_NBUF staging buffers, so CONV_NBUF - 1
        // stages of loads are always in flight.  The warps leave four 32-bit partial sums per coefficient
        // (P_c = S_2c + 2^8 S_2c+1, weight 2^16c) in four swizzled planes and every thread recombines its own 4
        // coefficients of the phase:   x = red( (P0 + 2^16 P1) + [(P2 + 2^16 P3) * (2^32 mod m)] ).
        // srcb = first source limb of the group, at this tile's first column.
        // Shared memory (dynamic): staging ring | planes R x 64 B | mbarriers; the ring's first R x 128 B are the
        // transpose buffer of the stages that follow.  Source jj of a buffer is skewed by 32 jj bytes (bank spread).
        constexpr int CONV_NBUF = 3;
        template <int LOGR>
        __host__ __device__ constexpr int conv_stage_bytes()
        {
            return 4 * ((1 << LOGR) / 4 * 128 + 32);
        }
        template <int LOGR>
        __host__ __device__ constexpr int conv_mma_smem()
        {
            return CONV_NBUF * conv_stage_bytes<LOGR>() + (1 << LOGR) * 64 + 8 * CONV_NBUF;
        }
        __device__ __forceinline__ unsigned smem_u32(const void *p)
        {
            return (unsigned)__cvta_generic_to_shared(p);
        }
        __device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
        {
            unsigned done = 0;
            while (!done)
            {
                asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                             "selp.b32 %0, 1, 0, P1;\n\t}"
                             : "=r"(done)
                             : "r"(bar), "r"(parity)
                             : "memory");
            }
        }
        template <int LOGR, bool WIDE>
        __device__ __forceinline__ void conv_prologue_mma(const FpField<WIDE> &f, const NttArgs &a, double (&x)[16],
                                                          const u64 *srcb, int g, int I, int rns, unsigned char *smem, int t,
                                                          int tb)
        {
            constexpr int R = 1 << LOGR, T1 = R / 16, QR = R / 4, PLANE = QR * 16, SRC_B = QR * 128 + 32,
                          STAGE_B = conv_stage_bytes<LOGR>();
            const ConvTab &cv = a.conv;
            const int cnt = cv.cnt[g];
            const int ksteps = (cnt + 3) >> 2;
            const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            const int g8 = lane >> 2, c4 = lane & 3;
            uint32_t *planes = reinterpret_cast<uint32_t *>(smem + CONV_NBUF * STAGE_B);
            const unsigned bar0 = smem_u32(smem + CONV_NBUF * STAGE_B + R * 64);
            if (tid == 0)
            {
#pragma unroll
                for (int i = 0; i < CONV_NBUF; i++)
                {
                    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
                }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            uint32_t bf[CONV_KSTEPS][2];
            {
                const uint32_t *bt = cv.BT + ((((size_t)g * rns + I) * CONV_KSTEPS) * 32 + lane) * 2;
#pragma unroll
                for (int s = 0; s < CONV_KSTEPS; s++)
                {
                    bf[s][0] = s < ksteps ? __ldg(bt + s * 64) : 0u;
                    bf[s][1] = s < ksteps ? __ldg(bt + s * 64 + 1) : 0u;
                }
            }
            const double c32 = cv.c32d[I];
            const int n_stages = 4 * ksteps; // stage u = (phase u / ksteps, k-step u % ksteps), buffer u % CONV_NBUF
            // this thread's row of a stage: source tid / QR of the k-step, row tid % QR of the phase
            const int my_jj = tid / QR, my_rr = tid % QR;
            const u64 *my_src = srcb + (size_t)my_rr * 256;
            const unsigned my_dst = smem_u32(smem) + my_jj * SRC_B + my_rr * 128;
            int iu = 0, ip = 0, is = 0, ib = 0; // next stage to issue: index, phase, k-step, buffer
            auto issue = [&]() {
                const int j = 4 * is + my_jj;
                const unsigned bar = bar0 + 8 * ib;
                if (tid == 0)
                {
                    const int nvalid = cnt - 4 * is < 4 ? cnt - 4 * is : 4;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nvalid * QR * 128)
                                 : "memory");
                }
                if (j < cnt)
                {
                    const u64 *src = my_src + ((size_t)j << a.log_n) + (size_t)ip * (QR * 256);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(my_dst + ib * STAGE_B), "l"(src), "r"(128), "r"(bar)
                                 : "memory");
                }
                iu++;
                if (++is == ksteps)
                {
                    is = 0;
                    ip++;
                }
                if (++ib == CONV_NBUF)
                {
                    ib = 0;
                }
            };
            for (int i = 0; i < CONV_NBUF - 1 && iu < n_stages; i++)
            {
                issue();
            }
            int acc[8][4];
            int p = 0, sI = 0, cb = 0;
            unsigned cpar = 0; // parity of the current buffer's mbarrier phase
            const unsigned char *frag = smem + c4 * SRC_B + (warp * 8) * 128 + g8 * 16;
            for (int u = 0; u < n_stages; u++)
            {
                if (iu < n_stages)
                {
                    issue(); // into the buffer every thread finished reading before the last __syncthreads
                }
                mbar_wait(bar0 + 8 * cb, cpar);
                if (sI == 0)
                {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                    {
                        acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0;
                    }
                }
                uint32_t bb0 = bf[0][0], bb1 = bf[0][1];
#pragma unroll
                for (int s2 = 1; s2 < CONV_KSTEPS; s2++)
                {
                    if (sI == s2)
                    {
                        bb0 = bf[s2][0];
                        bb1 = bf[s2][1];
                    }
                }
                // sources beyond cnt: their B rows are zero, whatever the staging buffer still holds there
                const unsigned char *fb = frag + cb * STAGE_B;
#pragma unroll
                for (int i = 0; i < 8; i++)
                {
                    const uint4 w = *reinterpret_cast<const uint4 *>(fb + i * 128);
                    const uint32_t af[4] = { w.x, w.z, w.y, w.w };
                    mma_u8(acc[i], af, bb0, bb1);
                }
                if (sI == ksteps - 1)
                {
                    uint32_t *pl = planes + c4 * PLANE;
#pragma unroll
                    for (int i = 0; i < 8; i++)
                    {
                        const int rr = warp * 8 + i;
                        uint2 v;
                        v.x = (uint32_t)acc[i][0] + ((uint32_t)acc[i][1] << 8);
                        v.y = (uint32_t)acc[i][2] + ((uint32_t)acc[i][3] << 8);
                        *reinterpret_cast<uint2 *>(pl + ((rr * 16 + 2 * g8) ^ (8 * c4))) = v;
                    }
                    __syncthreads();
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        const int idx = (t + T1 * k) * 16 + tb; // row t + T1 (4 p + k) - p QR
                        const u64 p0 = planes[idx], p1 = planes[PLANE + (idx ^ 8)], p2 = planes[2 * PLANE + (idx ^ 16)],
                                  p3 = planes[3 * PLANE + (idx ^ 24)];
                        const double lo = f.in_outer(p0 + (p1 << 16)), hi = f.in_outer(p2 + (p3 << 16));
                        const double v = f.red(__dadd_rn(lo, f.mul_lazy(hi, c32)));
#pragma unroll
                        for (int pp = 0; pp < 4; pp++)
                        {
                            if (pp == p)
                            {
                                x[4 * pp + k] = v;
                            }
                        }
                    }
                }
                __syncthreads();
                if (++sI == ksteps)
                {
                    sI = 0;
                    p++;
                }
                if (++cb == CONV_NBUF)
                {
                    cb = 0;
                    cpar ^= 1;
                }
            }
        }

        template <int LOGR>
        __global__ void __launch_bounds__((1 << LOGR) / 16 * TB, 2) ntt_fwd_pass_a_conv(NttArgs a, const Twiddle *two64)
        {
            extern __shared__ __align__(16) unsigned char conv_smem[]; // R x 128 B, or conv_mma_smem (ConvTab::BT)
            u64 *sm = reinterpret_cast<u64 *>(conv_smem);
            constexpr int T1 = (1 << LOGR) / 16;
            const int tb = threadIdx.x & (TB - 1), t = threadIdx.x >> 4;
            const long long poly = blockIdx.x / (256 / TB);
            const int tile = blockIdx.x % (256 / TB);
            const int I = (int)((poly / a.div) % a.period);
            const int g = (int)(poly % a.div);
            const long long b = poly / ((long long)a.period * a.div);
            if (a.conv.own && a.conv.own[I] == g)
            {
                return; // the digit's residue modulo its own primes is taken directly by the fused key-switch kernel
            }
            const int limb = a.limb_ids[I];
            const LimbConst lc = a.limb[limb];
            u64 *base = a.data + ((size_t)poly << a.log_n) + tile * TB + tb;
            // element k of this thread: row t + T1 k, column tile * TB + tb, of source limb s0[g] + j of item b
            const u64 *src0 = a.src + (((size_t)b * a.conv.src_limbs + a.conv.s0[g]) << a.log_n) + (size_t)t * 256 +
                              tile * TB + tb;
            const size_t row_stride = (size_t)T1 * 256;
            const u64 *srcb = a.src + (((size_t)b * a.conv.src_limbs + a.conv.s0[g]) << a.log_n) + tile * TB;
            if (lc.fp_class == 1)
            {
                const FpField<false> f(a, limb, lc);
                double x[16];
                if constexpr (LOGR >= 5)
                {
                    if (a.conv.BT)
                    {
                        conv_prologue_mma<LOGR>(f, a, x, srcb, g, I, a.period, conv_smem, t, tb);
                    }
                    else
                    {
                        conv_prologue(f, a, x, src0, row_stride, g, I, a.period);
                    }
                }
                else
                {
                    conv_prologue(f, a, x, src0, row_stride, g, I, a.period);
                }
                fwd_pass_a_stages<LOGR>(f, x, base, sm, t, tb);
            }
            else if (lc.fp_class == 2)
            {
                const FpField<true> f(a, limb, lc);
                double x[16];
                if constexpr (LOGR >= 5)
                {
                    if (a.conv.BT)
                    {
                        conv_prologue_mma<LOGR>(f, a, x, srcb, g, I, a.period, conv_smem, t, tb);
                    }
                    else
                    {
                        conv_prologue(f, a, x, src0, row_stride, g, I, a.period);
                    }
                }
                else
                {
                    conv_prologue(f, a, x, src0, row_stride, g, I, a.period);
                }
                fwd_pass_a_stages<LOGR>(f, x, base, sm, t, tb);
            }
            else
            {
                const IntField f(a, limb, lc);
                u64 x[16];
                conv_prologue(f, a, x, src0, row_stride, g, I, a.period, lc, two64[limb]);
                fwd_pass_a_stages<LOGR>(f, x, base, sm, t, tb);
            }
        }

        // =====================================================================================
        // Fused key-switch kernel: pass B of the digit-extension NTT + inner product with the evk.
        //
        // One CTA owns FR rows (FR x 256 output coefficients) of ONE target modulus I of ONE
        // ciphertext and loops over the digits J: it finishes NTT_I(d_J) for its rows in registers
        // (the 8 in-row stages, twiddles loaded once for all J) and multiplies the result straight
        // into the two accumulators  acc_k += NTT_I(d_J) (.) key[J][k][I]  (S/evaluator.cpp:2859-2883),
        // so the extended digits never reach HBM in NTT form and the separate MAC pass disappears.
        // Arithmetic: the exact FP64 products of the NTT (FpField::mul); the accumulators are integer-
        // valued doubles reduced often enough to stay below 2^53; the final canonical residues are
        // the same numbers the 128-bit integer inner product yields.  Data tiles and key tiles of
        // the next digit are prefetched with cp.async into per-thread shared-memory slots.
        // =====================================================================================
        constexpr int FR = 8;          // rows per CTA
        constexpr int FT = FR * 16;    // threads per CTA
        // shared memory: row tiles (landing zone of the next digit + transpose), per-thread twiddle
        // slots of the last four stages, row-uniform twiddles of the first four, key staging
        constexpr int KS_SM_ROWS = FR * ROW_PAD * 8, KS_SM_TW2 = FT * 8 * 16, KS_SM_TW1 = FR * 16 * 8,
                      KS_SM_KEYS = FT * 16 * 16;
        constexpr int KS_FUSED_SMEM = KS_SM_ROWS + KS_SM_TW2 + KS_SM_TW1 + KS_SM_KEYS; // 67584 B: 3 CTAs / SM

        struct KsFusedArgs
        {
            const u64 *mid;         // [batch][rns][digits][n]: pass-A output of the extended digits
            const u64 *ksk;         // [digits..][2][key_kl][n]
            u64 *acc;               // [batch][2][rns][n] canonical
            const double *tw_fp;    // [kl][n]
            const LimbConst *limb;  // [kl]
            const int *ids_ks;      // [rns]: prime index of target modulus I
            const u64 *direct;      // [batch][n_data][n] or nullptr (see ks_passb_mac)
            const int *own;         // [rns] or nullptr
            int limbs, rns, n_data, key_kl, log_n; // limbs = digits (KsShape, ntt.cuh)
            long long batch;
            int items; // ciphertexts per CTA
        };

        template <bool WIDE>
        __device__ __forceinline__ void ks_fused_body(const FpField<WIDE> &f, const KsFusedArgs &a, int I, long long b,
                                                      int n_items, int rb, unsigned char *smem)
        {
            const int tid = threadIdx.x, t = tid & 15, r = tid >> 4;
            const int R = 1 << (a.log_n - 8);
            const int row = rb * FR + r;
            const size_t ra = (size_t)R + row;
            double *srow = reinterpret_cast<double *>(smem) + r * ROW_PAD;
            double2 *tws = reinterpret_cast<double2 *>(smem + KS_SM_ROWS) + tid; // unit u at tws[u * FT]
            double2 *tw1 = reinterpret_cast<double2 *>(smem + KS_SM_ROWS + KS_SM_TW2) + r * 8;
            ulonglong2 *kst = reinterpret_cast<ulonglong2 *>(smem + KS_SM_ROWS + KS_SM_TW2 + KS_SM_TW1) + tid;

            const int key_limb = I < a.n_data ? I : I + a.key_kl - a.rns;
            const u64 *key0 = a.ksk + ((size_t)key_limb << a.log_n) + (size_t)row * 256 + 16 * t;
            const size_t key_poly = (size_t)a.key_kl << a.log_n; // stride between key[J][0] and key[J][1]
            const int own = a.own ? a.own[I] : -1; // digit taken from a.direct instead of being transformed
            const int digits = a.limbs;
            // digit order within an item: the direct one first, the others in order
            auto digit_at = [&](int i) { return own < 0 ? i : (i == 0 ? own : (i <= own ? i - 1 : i)); };

            // the row's 256 residues of digit J of item bb, natural order, 16 bytes per copy (coalesced).  The digit a
            // target owns comes from a.direct (prod(E) * c_I in NTT form: no transform) through the same tile.
            auto issue_data = [&](long long bb, int J) {
                const u64 *src = J == own ? a.direct + (((size_t)bb * a.n_data + I) << a.log_n) + (size_t)row * 256
                                          : a.mid + ((((size_t)bb * a.rns + I) * digits + J) << a.log_n) + (size_t)row * 256;
#pragma unroll
                for (int j = 0; j < 8; j++)
                {
                    cp_async16(srow + 2 * (t + 16 * j), src + 2 * (t + 16 * j));
                }
            };
            auto issue_keys = [&](int J) {
                const u64 *src = key0 + (size_t)J * 2 * key_poly;
#pragma unroll
                for (int u = 0; u < 8; u++)
                {
                    cp_async16(kst + u * FT, src + 2 * u);
                    cp_async16(kst + (8 + u) * FT, src + key_poly + 2 * u);
                }
            };
            issue_data(b, digit_at(0));
            cp_async_commit();
            issue_keys(digit_at(0));
            cp_async_commit();

            // twiddles, loaded once for all digits AND all items of this CTA: the 15 row-uniform ones of the first four
            // stages (tw1: [0] = 2^0 block, [1..2], [3..6], [7..14]) and this thread's 15 of the last four
            if (t < 15)
            {
                const int lvl = t == 0 ? 0 : (t < 3 ? 1 : (t < 7 ? 2 : 3));
                const size_t idx = (ra << lvl) + (t - ((1 << lvl) - 1));
                reinterpret_cast<double *>(tw1)[t] = f.tw(idx);
            }
            {
                const double *g1 = f.tab + 128 * ra + 8 * t, *g2 = f.tab + 64 * ra + 4 * t, *g4 = f.tab + 32 * ra + 2 * t;
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    tws[u * FT] = make_double2(__ldg(g1 + 2 * u), __ldg(g1 + 2 * u + 1));
                }
                tws[4 * FT] = make_double2(__ldg(g2), __ldg(g2 + 1));
                tws[5 * FT] = make_double2(__ldg(g2 + 2), __ldg(g2 + 3));
                tws[6 * FT] = make_double2(__ldg(g4), __ldg(g4 + 1));
                tws[7 * FT] = make_double2(__ldg(f.tab + 16 * ra + t), 0.0);
            }

            double acc0[16], acc1[16];
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                acc0[k] = 0.0;
                acc1[k] = 0.0;
            }
            // canonical key residue (< 2^52) -> double.  The key is NOT centred: for a multiplier input
            // |x| <= 6.2 p (narrow) / p/2 + 1 (wide) the product stays below 2^99 / 2^101, its low part
            // below p/8, so FpField::mul's result bound grows only from 0.52 p to 0.65 p (narrow).
            auto key_d = [&](u64 v) { return f.in_outer(v); };
            const int red_every = WIDE ? 2 : 8; // |acc| <= 0.5 p + 8 * 0.65 p  /  0.5 p + 2 * 1.125 p

            // (item, digit) steps of this CTA flattened into one software pipeline: the next step's row tile and key tiles
            // are in flight while the current one computes, across item boundaries too
            long long bb = b;
            int it = 0;
            const long long total = (long long)n_items * digits;
            for (long long sq = 0; sq < total; sq++)
            {
                const int J = digit_at(it);
                const bool last_digit = it + 1 == digits;
                const long long bn = last_digit ? bb + 1 : bb;
                const int Jn = digit_at(last_digit ? 0 : it + 1);
                const bool more = sq + 1 < total;
                double x[16];
                cp_async_wait_1(); // data(J) has landed (keys(J) may still be in flight)
                __syncwarp();      // a row is half a warp: its 16 threads' copies are now visible to each other
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = srow[t + 16 * k];
                }
                if (J != own)
                {
                    f.phase_begin_fwd(x);
                    const double2 w0 = tw1[0], w1 = tw1[1], w2 = tw1[2], w3 = tw1[3], w4 = tw1[4], w5 = tw1[5],
                                  w6 = tw1[6], w7 = tw1[7];
                    const double t8[1] = { w0.x };
                    const double t4[2] = { w0.y, w1.x };
                    const double t2[4] = { w1.y, w2.x, w2.y, w3.x };
                    const double t1[8] = { w3.y, w4.x, w4.y, w5.x, w5.y, w6.x, w6.y, w7.x };
                    ct_stage_tw<8>(f, x, t8);
                    ct_stage_tw<4>(f, x, t4);
                    f.phase_mid_fwd(x);
                    ct_stage_tw<2>(f, x, t2);
                    ct_stage_tw<1>(f, x, t1);
                }
                __syncwarp(); // every thread of the row has read its inputs: the tile can be overwritten
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    srow[t + 17 * k] = x[k];
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = srow[17 * t + k];
                }
                __syncwarp();
                if (more)
                {
                    issue_data(bn, Jn); // lands in the row tile while the last stages and the MAC run
                }
                cp_async_commit();
                if (J != own)
                {
                    f.phase_begin_fwd(x);
                    const double2 a0 = tws[0 * FT], a1 = tws[1 * FT], a2 = tws[2 * FT], a3 = tws[3 * FT];
                    const double2 b0 = tws[4 * FT], b1 = tws[5 * FT], c0 = tws[6 * FT], d0 = tws[7 * FT];
                    const double t8[1] = { d0.x };
                    const double t4[2] = { c0.x, c0.y };
                    const double t2[4] = { b0.x, b0.y, b1.x, b1.y };
                    const double t1[8] = { a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y };
                    ct_stage_tw<8>(f, x, t8);
                    ct_stage_tw<4>(f, x, t4);
                    f.phase_mid_fwd(x);
                    ct_stage_tw<2>(f, x, t2);
                    ct_stage_tw<1>(f, x, t1);
                    if (WIDE)
                    {
#pragma unroll
                        for (int k = 0; k < 16; k++)
                        {
                            x[k] = f.red(x[k]); // multiplier input below 2^52
                        }
                    }
                }
                else
                {
                    // prod(E) * c_I arrives as canonical residues (the bit pattern in_mid / in_outer expect differ)
#pragma unroll
                    for (int k = 0; k < 16; k++)
                    {
                        x[k] = f.red(f.in_outer((u64)__double_as_longlong(x[k])));
                    }
                }
                cp_async_wait_1(); // keys(J) have landed
#pragma unroll
                for (int u = 0; u < 8; u++)
                {
                    const ulonglong2 k0 = kst[u * FT], k1 = kst[(8 + u) * FT];
                    acc0[2 * u] = __dadd_rn(acc0[2 * u], f.mul_lazy(x[2 * u], key_d(k0.x)));
                    acc0[2 * u + 1] = __dadd_rn(acc0[2 * u + 1], f.mul_lazy(x[2 * u + 1], key_d(k0.y)));
                    acc1[2 * u] = __dadd_rn(acc1[2 * u], f.mul_lazy(x[2 * u], key_d(k1.x)));
                    acc1[2 * u + 1] = __dadd_rn(acc1[2 * u + 1], f.mul_lazy(x[2 * u + 1], key_d(k1.y)));
                }
                if (more)
                {
                    issue_keys(Jn);
                }
                cp_async_commit();
                if (last_digit)
                {
                    u64 *o0 = a.acc + ((((size_t)bb * 2 + 0) * a.rns + I) << a.log_n) + (size_t)row * 256 + 16 * t;
                    u64 *o1 = a.acc + ((((size_t)bb * 2 + 1) * a.rns + I) << a.log_n) + (size_t)row * 256 + 16 * t;
#pragma unroll
                    for (int k = 0; k < 16; k += 2)
                    {
                        ulonglong2 v;
                        v.x = f.canon(acc0[k]);
                        v.y = f.canon(acc0[k + 1]);
                        reinterpret_cast<ulonglong2 *>(o0)[k >> 1] = v;
                        v.x = f.canon(acc1[k]);
                        v.y = f.canon(acc1[k + 1]);
                        reinterpret_cast<ulonglong2 *>(o1)[k >> 1] = v;
                        acc0[k] = acc0[k + 1] = acc1[k] = acc1[k + 1] = 0.0;
                    }
                    bb++;
                    it = 0;
                }
                else
                {
                    if ((it + 1) % red_every == 0)
                    {
#pragma unroll
                        for (int k = 0; k < 16; k++)
                        {
                            acc0[k] = f.red(acc0[k]);
                            acc1[k] = f.red(acc1[k]);
                        }
                    }
                    it++;
                }
            }
        }

        __global__ void __launch_bounds__(FT, 3) ks_passb_mac_kernel(KsFusedArgs a, NttArgs na)
        {
            extern __shared__ __align__(16) unsigned char ks_smem[];
            // a CTA owns up to a.items consecutive ciphertexts of one (row block, target modulus): twiddles loaded once,
            // one software pipeline across them (the start-up latency of a 3-digit item is a third of its run time)
            const long long b = (long long)blockIdx.x * a.items;
            const int n_items = (int)(a.batch - b < a.items ? a.batch - b : a.items);
            const int rb = blockIdx.y;
            const int I = blockIdx.z;
            const int limb = a.ids_ks[I];
            const LimbConst lc = a.limb[limb];
            if (lc.fp_class == 1)
            {
                const FpField<false> f(na, limb, lc);
                ks_fused_body<false>(f, a, I, b, n_items, rb, ks_smem);
            }
            else if (lc.fp_class == 2)
            {
                const FpField<true> f(na, limb, lc);
                ks_fused_body<true>(f, a, I, b, n_items, rb, ks_smem);
            }
            // integer-path moduli (the 58-bit special prime) are handled by the un-fused kernels
        }

        // =====================================================================================
        // Grouped pass B: the same 8 in-row stages as ntt_fwd_pass_b, but one CTA transforms its 8 rows of
        // UP TO GS POLYNOMIALS THAT SHARE A PRIME one after the other (in a ciphertext batch [P][limbs]
        // those are the polynomials `limbs` apart; in the extended digits of a key switch the `limbs`
        // consecutive ones).  The 15 + 15 twiddles per thread are loaded once per CTA instead of once
        // per polynomial and the next polynomial's rows arrive by cp.async while the current one
        // computes, which removes the twiddle-latency stalls that held ntt_fwd_pass_b at 53 % FP64
        // utilisation (profiles/ncu_full_ntt_r1.csv).  Integer-path primes run the classic body.
        // =====================================================================================
        constexpr int GS = 16; // polynomials per CTA
        constexpr int GROUPED_SMEM = KS_SM_ROWS + KS_SM_TW2 + KS_SM_TW1;

        template <bool WIDE>
        __device__ __forceinline__ void fwd_pass_b_grouped_fp(const FpField<WIDE> &f, const NttArgs &a, long long slot,
                                                              long long q0, long long q1, int rb, unsigned char *smem)
        {
            const int tid = threadIdx.x, t = tid & 15, r = tid >> 4;
            const int R = 1 << (a.log_n - 8);
            const int row = rb * FR + r;
            const size_t ra = (size_t)R + row;
            double *srow = reinterpret_cast<double *>(smem) + r * ROW_PAD;
            double2 *tws = reinterpret_cast<double2 *>(smem + KS_SM_ROWS) + tid;
            double2 *tw1 = reinterpret_cast<double2 *>(smem + KS_SM_ROWS + KS_SM_TW2) + r * 8;
            // q-th polynomial of this prime: p = ((q / div) * period + slot) * div + q % div
            auto poly_ptr = [&](long long q) {
                const long long p = ((q / a.div) * a.period + slot) * a.div + q % a.div;
                return a.data + ((size_t)p << a.log_n) + (size_t)row * 256;
            };
            auto issue_data = [&](long long q) {
                const u64 *src = poly_ptr(q);
#pragma unroll
                for (int j = 0; j < 8; j++)
                {
                    cp_async16(srow + 2 * (t + 16 * j), src + 2 * (t + 16 * j));
                }
            };
            issue_data(q0);
            cp_async_commit();
            if (t < 15)
            {
                const int lvl = t == 0 ? 0 : (t < 3 ? 1 : (t < 7 ? 2 : 3));
                const size_t idx = (ra << lvl) + (t - ((1 << lvl) - 1));
                reinterpret_cast<double *>(tw1)[t] = f.tw(idx);
            }
            {
                const double *g1 = f.tab + 128 * ra + 8 * t, *g2 = f.tab + 64 * ra + 4 * t, *g4 = f.tab + 32 * ra + 2 * t;
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    tws[u * FT] = make_double2(__ldg(g1 + 2 * u), __ldg(g1 + 2 * u + 1));
                }
                tws[4 * FT] = make_double2(__ldg(g2), __ldg(g2 + 1));
                tws[5 * FT] = make_double2(__ldg(g2 + 2), __ldg(g2 + 3));
                tws[6 * FT] = make_double2(__ldg(g4), __ldg(g4 + 1));
                tws[7 * FT] = make_double2(__ldg(f.tab + 16 * ra + t), 0.0);
            }
            for (long long q = q0; q < q1; q++)
            {
                double x[16];
                if (a.fin.out)
                {
                    // the epilogue's operands (this thread's 128-byte line of the input limb and of the addend) are
                    // pulled into L2 now: by the time the 8 stages are done the loads below are L2 hits
                    const FinishEpi &e = a.fin;
                    const size_t off = (size_t)row * 256 + 16 * t;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(e.in + (((size_t)q * e.limbs_in + slot) << a.log_n) + off));
                    if (e.addend && !(e.addend_even_only && (q & 1)))
                    {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(
                            e.addend + (((size_t)((q >> 1) * e.addend_group + (q & 1)) *
                                             (e.addend_limbs ? e.addend_limbs : a.period) + slot) << a.log_n) + off));
                    }
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp(); // a row is half a warp
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = srow[t + 16 * k];
                }
                f.phase_begin_fwd(x);
                {
                    const double2 w0 = tw1[0], w1 = tw1[1], w2 = tw1[2], w3 = tw1[3], w4 = tw1[4], w5 = tw1[5],
                                  w6 = tw1[6], w7 = tw1[7];
                    const double t8[1] = { w0.x };
                    const double t4[2] = { w0.y, w1.x };
                    const double t2[4] = { w1.y, w2.x, w2.y, w3.x };
                    const double t1[8] = { w3.y, w4.x, w4.y, w5.x, w5.y, w6.x, w6.y, w7.x };
                    ct_stage_tw<8>(f, x, t8);
                    ct_stage_tw<4>(f, x, t4);
                    f.phase_mid_fwd(x);
                    ct_stage_tw<2>(f, x, t2);
                    ct_stage_tw<1>(f, x, t1);
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    srow[t + 17 * k] = x[k];
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 16; k++)
                {
                    x[k] = srow[17 * t + k];
                }
                __syncwarp();
                if (q + 1 < q1)
                {
                    issue_data(q + 1);
                }
                cp_async_commit();
                f.phase_begin_fwd(x);
                {
                    const double2 a0 = tws[0 * FT], a1 = tws[1 * FT], a2 = tws[2 * FT], a3 = tws[3 * FT];
                    const double2 b0 = tws[4 * FT], b1 = tws[5 * FT], c0 = tws[6 * FT], d0 = tws[7 * FT];
                    const double t8[1] = { d0.x };
                    const double t4[2] = { c0.x, c0.y };
                    const double t2[4] = { b0.x, b0.y, b1.x, b1.y };
                    const double t1[8] = { a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y };
                    ct_stage_tw<8>(f, x, t8);
                    ct_stage_tw<4>(f, x, t4);
                    f.phase_mid_fwd(x);
                    ct_stage_tw<2>(f, x, t2);
                    ct_stage_tw<1>(f, x, t1);
                }
                if (a.fin.out)
                {
                    // divide-and-round tail (FinishEpi): q is the polynomial index P, slot the target limb
                    const FinishEpi &e = a.fin;
                    const size_t off = (size_t)row * 256 + 16 * t;
                    const ulonglong2 *in = reinterpret_cast<const ulonglong2 *>(
                        e.in + (((size_t)q * e.limbs_in + slot) << a.log_n) + off);
                    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(e.out + (((size_t)q * a.period + slot) << a.log_n) + off);
                    const ulonglong2 *ad = nullptr;
                    if (e.addend && !(e.addend_even_only && (q & 1)))
                    {
                        ad = reinterpret_cast<const ulonglong2 *>(
                            e.addend + (((size_t)((q >> 1) * e.addend_group + (q & 1)) *
                                             (e.addend_limbs ? e.addend_limbs : a.period) + slot) << a.log_n) + off);
                    }
                    const Twiddle inv = e.inv[slot];
                    const Twiddle am = e.addend_mul ? e.addend_mul[slot] : Twiddle{ 0, 0 };
                    const u64 qq = f.pi;
#pragma unroll
                    for (int k = 0; k < 16; k += 2)
                    {
                        const ulonglong2 c0 = in[k >> 1];
                        ulonglong2 v;
                        v.x = mul_shoup(submod(c0.x, f.out_fwd(x[k]), qq), inv.w, inv.wq, qq);
                        v.y = mul_shoup(submod(c0.y, f.out_fwd(x[k + 1]), qq), inv.w, inv.wq, qq);
                        if (ad)
                        {
                            ulonglong2 z = ad[k >> 1];
                            if (e.addend_mul)
                            {
                                z.x = mul_shoup(z.x, am.w, am.wq, qq);
                                z.y = mul_shoup(z.y, am.w, am.wq, qq);
                            }
                            v.x = addmod(v.x, z.x, qq);
                            v.y = addmod(v.y, z.y, qq);
                        }
                        o[k >> 1] = v;
                    }
                    continue;
                }
                ulonglong2 *out = reinterpret_cast<ulonglong2 *>(poly_ptr(q) + 16 * t);
#pragma unroll
                for (int k = 0; k < 16; k += 2)
                {
                    ulonglong2 v;
                    v.x = f.out_fwd(x[k]);
                    v.y = f.out_fwd(x[k + 1]);
                    out[k >> 1] = v;
                }
            }
        }

        __global__ void __launch_bounds__(FT) ntt_fwd_pass_b_grouped(NttArgs a, long long seq_len)
        {
            extern __shared__ __align__(16) unsigned char grp_smem[];
            const long long q0 = (long long)blockIdx.x * GS;
            const long long q1 = q0 + GS < seq_len ? q0 + GS : seq_len;
            const int rb = blockIdx.y;
            const long long slot = blockIdx.z;
            const int limb = a.limb_ids[slot];
            const LimbConst lc = a.limb[limb];
            if (lc.fp_class == 1)
            {
                const FpField<false> f(a, limb, lc);
                fwd_pass_b_grouped_fp<false>(f, a, slot, q0, q1, rb, grp_smem);
            }
            else if (lc.fp_class == 2)
            {
                const FpField<true> f(a, limb, lc);
                fwd_pass_b_grouped_fp<true>(f, a, slot, q0, q1, rb, grp_smem);
            }
            else
            {
                const IntField f(a, limb, lc);
                const int t = threadIdx.x & 15, r = threadIdx.x >> 4;
                const int R = 1 << (a.log_n - 8);
                const int row = rb * FR + r;
                for (long long q = q0; q < q1; q++)
                {
                    const long long p = ((q / a.div) * a.period + slot) * a.div + q % a.div;
                    u64 *base = a.data + ((size_t)p << a.log_n) + (size_t)row * 256;
                    __syncthreads(); // the transpose rows are reused by the next polynomial
                    fwd_pass_b_body(f, base, reinterpret_cast<u64 *>(grp_smem) + r * ROW_PAD, t, (size_t)R + row);
                }
            }
        }

        // =====================================================================================
        // Hoisted rotations: inner products of ONE set of extended digits (NTT form) with KSM_R keys in
        // one pass: the big operand (limbs (limbs+1) x 512 KiB per ciphertext) is read once for KSM_R
        // rotations, and the products run on the FP64 pipe (exact, FpField::mul_lazy), which retires a
        // modular MAC ~2.7x faster than the 128-bit integer path.  Digit loop unrolled by two with all
        // loads of a pair issued first.  Integer-path moduli are left to the integer kernel.
        // =====================================================================================
        struct KsMacMultiArgs
        {
            const u64 *ext;          // [batch][rns][digits][n] NTT form, canonical
            const u64 *ksk[KSM_R];
            u64 *acc[KSM_R];         // each [batch][2][rns][n]
            int key_kl[KSM_R];
            const LimbConst *limb;
            const int *ids_ks;
            int limbs, rns, n_data, log_n; // limbs = digits (KsShape, ntt.cuh)
        };

        // Streams of one digit J for a CTA: the extended digit and key[J][0], key[J][1] of each of the R keys,
        // 256 x 16 bytes each.  They travel through a KSM_STAGES-deep cp.async ring of thread-private
        // 16-byte slots (a thread consumes exactly what it copied: no CTA barrier), so the loads of the next
        // digits are in flight while the current one is multiplied: the kernel was latency-bound on its
        // loads (ncu: long_scoreboard 13 warps per issue, FP64 pipe 49 %) with direct global loads.
        constexpr int KSM_STAGES = 3;
        __host__ __device__ constexpr int ksm_smem_bytes(int r)
        {
            return KSM_STAGES * (1 + 2 * r) * 256 * 16;
        }

        template <bool WIDE, int R>
        __device__ __forceinline__ void ks_mac_multi_body(const FpField<WIDE> &f, const KsMacMultiArgs &a, int I, long long b,
                                                          long long within, unsigned char *smem)
        {
            const int log_n2 = a.log_n - 1;
            const long long n2 = (long long)1 << log_n2;
            const ulonglong2 *e = reinterpret_cast<const ulonglong2 *>(a.ext) + (((b * a.rns + I) * a.limbs) << log_n2) + within;
            const ulonglong2 *kp[R];
            long long kstep[R], kpoly[R];
#pragma unroll
            for (int r = 0; r < R; r++)
            {
                const int key_limb = I < a.n_data ? I : I + a.key_kl[r] - a.rns;
                kp[r] = reinterpret_cast<const ulonglong2 *>(a.ksk[r]) + (long long)key_limb * n2 + within;
                kpoly[r] = (long long)a.key_kl[r] * n2;  // key[J][0] -> key[J][1]
                kstep[r] = 2 * kpoly[r];                 // key[J] -> key[J + 1]
            }
            constexpr int KSM_STREAMS = 1 + 2 * R;
            ulonglong2 *slot = reinterpret_cast<ulonglong2 *>(smem) + threadIdx.x; // stream q of stage s: slot[(s * STREAMS + q) * 256]
            auto issue = [&](int J) {
                ulonglong2 *dst = slot + (size_t)(J % KSM_STAGES) * KSM_STREAMS * 256;
                cp_async16(dst, e + ((long long)J << log_n2));
#pragma unroll
                for (int r = 0; r < R; r++)
                {
                    const ulonglong2 *p = kp[r] + (long long)J * kstep[r];
                    cp_async16(dst + (1 + 2 * r) * 256, p);
                    cp_async16(dst + (2 + 2 * r) * 256, p + kpoly[r]);
                }
            };
#pragma unroll
            for (int J = 0; J < KSM_STAGES - 1; J++)
            {
                if (J < a.limbs)
                {
                    issue(J);
                }
                cp_async_commit();
            }
            double acc[R][4];
#pragma unroll
            for (int r = 0; r < R; r++)
            {
                acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0;
            }
            const int red_every = WIDE ? 2 : 8; // |term| <= 1.125 p (51-bit class) / 0.65 p (46-bit class)
            for (int J = 0; J < a.limbs; J++)
            {
                asm volatile("cp.async.wait_group %0;" ::"n"(KSM_STAGES - 2) : "memory"); // digit J has landed
                const ulonglong2 *src = slot + (size_t)(J % KSM_STAGES) * KSM_STREAMS * 256;
                const ulonglong2 v = src[0];
                const double ex = f.in_outer(v.x), ey = f.in_outer(v.y);
#pragma unroll
                for (int r = 0; r < R; r++)
                {
                    const ulonglong2 k0 = src[(1 + 2 * r) * 256], k1 = src[(2 + 2 * r) * 256];
                    acc[r][0] = __dadd_rn(acc[r][0], f.mul_lazy(ex, f.in_outer(k0.x)));
                    acc[r][1] = __dadd_rn(acc[r][1], f.mul_lazy(ey, f.in_outer(k0.y)));
                    acc[r][2] = __dadd_rn(acc[r][2], f.mul_lazy(ex, f.in_outer(k1.x)));
                    acc[r][3] = __dadd_rn(acc[r][3], f.mul_lazy(ey, f.in_outer(k1.y)));
                }
                // digit J + STAGES - 1 lands in the slots digit J - 1 was read from (same thread, program order)
                if (J + KSM_STAGES - 1 < a.limbs)
                {
                    issue(J + KSM_STAGES - 1);
                }
                cp_async_commit();
                if ((J + 1) % red_every == 0)
                {
#pragma unroll
                    for (int r = 0; r < R; r++)
                    {
#pragma unroll
                        for (int q = 0; q < 4; q++)
                        {
                            acc[r][q] = f.red(acc[r][q]);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++)
            {
                ulonglong2 *o = reinterpret_cast<ulonglong2 *>(a.acc[r]);
                ulonglong2 r0, r1;
                r0.x = f.canon(acc[r][0]);
                r0.y = f.canon(acc[r][1]);
                r1.x = f.canon(acc[r][2]);
                r1.y = f.canon(acc[r][3]);
                o[(((b * 2 + 0) * a.rns + I) << log_n2) + within] = r0;
                o[(((b * 2 + 1) * a.rns + I) << log_n2) + within] = r1;
            }
        }

        template <int R>
        __global__ void __launch_bounds__(256) ks_mac_multi_kernel(KsMacMultiArgs a, NttArgs na)
        {
            extern __shared__ __align__(16) unsigned char ksm_smem[];
            // grid: x = ciphertext (fastest: CTAs sharing a key tile run together), y = I, z = coefficient block
            const long long b = blockIdx.x;
            const int I = blockIdx.y;
            const long long within = (long long)blockIdx.z * blockDim.x + threadIdx.x;
            const int limb = a.ids_ks[I];
            const LimbConst lc = a.limb[limb];
            if (lc.fp_class == 1)
            {
                const FpField<false> f(na, limb, lc);
                ks_mac_multi_body<false, R>(f, a, I, b, within, ksm_smem);
            }
            else if (lc.fp_class == 2)
            {
                const FpField<true> f(na, limb, lc);
                ks_mac_multi_body<true, R>(f, a, I, b, within, ksm_smem);
            }
        }

        // KernelTimer names: "k_ntt_fwd_pass_a" / "k_ntt_fwd_pass_b" with units = limb-transforms (bench.py's roofline)
        // returns true when the FinishEpi epilogue was applied
        template <int LOGR>
        bool launch_fwd(Context *c, NttArgs a, cudaStream_t s, bool do_a = true, bool do_b = true)
        {
            bool fin_applied = false;
            if (do_a)
            {
                KernelTimer kt(c, a.src_mode == 3 ? "k_ntt_fwd_pass_a_conv" : "k_ntt_fwd_pass_a", a.count - a.skipped);
                const long long ctas_a = a.count * (256 / TB);
                if (a.src_mode == 3)
                {
                    constexpr int smem_max = LOGR >= 5 ? conv_mma_smem<LOGR>() : (1 << LOGR) * TB * 8;
                    static const bool attr_set = [] {
                        MOAI_CUDA_CHECK(cudaFuncSetAttribute(ntt_fwd_pass_a_conv<LOGR>,
                                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
                        return true;
                    }();
                    (void)attr_set;
                    const int smem = a.conv.BT ? smem_max : (1 << LOGR) * TB * 8;
                    ntt_fwd_pass_a_conv<LOGR><<<(unsigned)ctas_a, (1 << LOGR) / 16 * TB, smem, s>>>(a, c->d_two64);
                }
                else
                {
                    ntt_fwd_pass_a<LOGR><<<(unsigned)ctas_a, (1 << LOGR) / 16 * TB, 0, s>>>(a);
                }
            }
            if (do_b)
            {
                // polynomials sharing a prime: count / period of them per limb slot
                const long long seq_len = a.count / a.period;
                static const bool grouped_on = [] {
                    const char *e = getenv("MOAI_NTT_GROUPED");
                    return !e || atoi(e) != 0;
                }();
                const bool grouped = grouped_on && !a.grp_size && a.count % ((long long)a.period * a.div) == 0 &&
                                     seq_len >= 4 && a.period <= 65535;
                if (!grouped)
                {
                    a.fin = FinishEpi();
                }
                fin_applied = a.fin.out != nullptr;
                // (with the divide-and-round tail as its epilogue the kernel moves 1.5-2 MiB per unit, not 1: own timer)
                KernelTimer kt(c, fin_applied ? "k_ntt_fwd_pass_b_finish" : "k_ntt_fwd_pass_b", a.count);
                if (grouped)
                {
                    dim3 grid((unsigned)((seq_len + GS - 1) / GS), (unsigned)((1 << LOGR) / FR), (unsigned)a.period);
                    ntt_fwd_pass_b_grouped<<<grid, FT, GROUPED_SMEM, s>>>(a, seq_len);
                }
                else
                {
                    const long long ctas_b = a.count * ((1 << LOGR) / ROWS);
                    ntt_fwd_pass_b<<<(unsigned)ctas_b, ROWS * 16, 0, s>>>(a);
                }
            }
            return fin_applied;
        }

        // The two passes run over chunks small enough for pass B's output to still be in L2 when pass A reads it
        // (MOAI_NTT_L2_LIMBS polynomials per chunk; default 0 = one launch pair: the serialised tails of the small launches
        // cost more than the L2 hits save — relin_rescale at 28 limbs, 64 ciphertexts)
        template <int LOGR>
        void launch_inv(const NttArgs &a0, cudaStream_t s)
        {
            static const long long l2_limbs = [] {
                const char *e = getenv("MOAI_NTT_L2_LIMBS");
                return e ? atoll(e) : 0ll; // measured: chunks of 48 / 96 / 192 limbs cost 2.70 / 2.21 / 2.03 ms against 1.76 ms in one launch pair
            }();
            long long chunk = l2_limbs > 0 ? std::max<long long>(1, (l2_limbs << 16) >> a0.log_n) : a0.count;
            if (chunk * 2 > a0.count)
            {
                chunk = a0.count; // not worth splitting
            }
            for (long long p0 = 0; p0 < a0.count; p0 += chunk)
            {
                NttArgs a = a0;
                a.p_base = p0;
                const long long cnt = std::min(chunk, a0.count - p0);
                const long long ctas_b = cnt * ((1 << LOGR) / ROWS);
                ntt_inv_pass_b<<<(unsigned)ctas_b, ROWS * 16, 0, s>>>(a);
                const long long ctas_a = cnt * (256 / TB);
                ntt_inv_pass_a<LOGR><<<(unsigned)ctas_a, (1 << LOGR) / 16 * TB, 0, s>>>(a);
            }
        }
    } // namespace

    namespace
    {
        // v = rint(sum_j y_j / q_j) of every (item, digit, coefficient), written into byte 7 of the digit's first source
        __global__ void k_conv_quot(ulonglong2 *src, long long total2, int log_n2, int digits, int src_limbs, ConvTab cv)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [batch][digits][n/2]
            if (i >= total2)
            {
                return;
            }
            const long long within = i & (((long long)1 << log_n2) - 1);
            const long long bg = i >> log_n2;
            const int g = (int)(bg % digits);
            const long long b = bg / digits;
            const int s0 = cv.s0[g], cnt = cv.cnt[g];
            ulonglong2 *p0 = src + (((b * src_limbs + s0) << log_n2) + within);
            double vx = 0.0, vy = 0.0;
            ulonglong2 first = *p0;
            for (int j = 0; j < cnt; j++)
            {
                const ulonglong2 y = j == 0 ? first : p0[(long long)j << log_n2];
                const double iq = cv.invq[s0 + j];
                const bool wide = cv.wide[s0 + j] != 0;
                vx = __fma_rn(conv_y_double(y.x, wide), iq, vx);
                vy = __fma_rn(conv_y_double(y.y, wide), iq, vy);
            }
            first.x |= (u64)__double2ll_rn(conv_rint(vx)) << 56;
            first.y |= (u64)__double2ll_rn(conv_rint(vy)) << 56;
            *p0 = first;
        }
    } // namespace

    namespace
    {
        __global__ void k_conv_quot_fp(const ulonglong2 *__restrict__ src, double2 *__restrict__ v, long long total2, int log_n2,
                                       int digits, int src_limbs, ConvTab cv)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [batch][digits][n/2]
            if (i >= total2)
            {
                return;
            }
            const long long within = i & (((long long)1 << log_n2) - 1);
            const long long bg = i >> log_n2;
            const int g = (int)(bg % digits);
            const long long b = bg / digits;
            const int s0 = cv.s0[g], cnt = cv.cnt[g];
            const ulonglong2 *p0 = src + (((b * src_limbs + s0) << log_n2) + within);
            double vx = 0.0, vy = 0.0;
            for (int j = 0; j < cnt; j++)
            {
                const ulonglong2 y = p0[(long long)j << log_n2];
                const double iq = cv.invq[s0 + j];
                if (cv.fpsrc[s0 + j])
                {
                    vx = __fma_rn(__longlong_as_double((long long)y.x), iq, vx);
                    vy = __fma_rn(__longlong_as_double((long long)y.y), iq, vy);
                }
                else
                {
                    const bool wide = cv.wide[s0 + j] != 0;
                    vx = __fma_rn(conv_y_double(y.x, wide), iq, vx);
                    vy = __fma_rn(conv_y_double(y.y, wide), iq, vy);
                }
            }
            v[i] = make_double2(conv_rint(vx), conv_rint(vy));
        }
    } // namespace

    void conv_quotient_fp(Context *c, const u64 *src, long long batch, const ConvTab &tab, int digits, double *v)
    {
        const long long total2 = batch * digits * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        KernelTimer kt(c, "k_conv_quot_fp", 1);
        k_conv_quot_fp<<<(unsigned)((total2 + 255) / 256), 256, 0, c->stream>>>(
            reinterpret_cast<const ulonglong2 *>(src), reinterpret_cast<double2 *>(v), total2, c->log_n - 1, digits,
            tab.src_limbs, tab);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void conv_quotient(Context *c, u64 *src, long long batch, const ConvTab &tab, int digits)
    {
        const long long total2 = batch * digits * (long long)(c->n / 2);
        if (!total2)
        {
            return;
        }
        KernelTimer kt(c, "k_conv_quot", 1);
        k_conv_quot<<<(unsigned)((total2 + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<ulonglong2 *>(src), total2,
                                                                             c->log_n - 1, digits, tab.src_limbs, tab);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    bool ntt_forward(Context *c, u64 *data, long long count, const int *d_limb_ids, int period, int div,
                     const NttPrologue *pro, int passes, const FinishEpi *fin)
    {
        if (count <= 0)
        {
            return false;
        }
        const bool do_a = (passes & 1) != 0, do_b = (passes & 2) != 0;
        NttArgs a{ data, c->d_fwd, c->d_fwd_fp, c->d_limb, d_limb_ids, period, div, c->log_n, count };
        if (pro)
        {
            a.src = pro->src;
            a.src_mode = pro->mode;
            a.last_id = pro->last_id;
            a.kl = c->kl;
            a.half_mod = c->d_half_mod;
            if (pro->mode == 3)
            {
                MOAI_REQUIRE(pro->conv != nullptr, "base-conversion prologue without tables");
                a.conv = *pro->conv;
                a.skipped = pro->skipped;
                a.conv_v = pro->conv_v;
            }
        }
        if (fin && do_b && div == 1)
        {
            // only FP64-path target primes take the grouped kernel's FP64 body (the integer body stores in place)
            bool all_fp = d_limb_ids == c->d_ids;
            for (int l = 0; all_fp && l < period; l++)
            {
                all_fp = c->h_limb[l].fp_class != 0;
            }
            static const bool fuse_on = [] {
                const char *e = getenv("MOAI_FUSE_FINISH");
                return !e || atoi(e) != 0;
            }();
            if (all_fp && fuse_on)
            {
                a.fin = *fin;
            }
        }
        bool applied = false;
        switch (c->log_n)
        {
        case 12: applied = launch_fwd<4>(c, a, c->stream, do_a, do_b); break;
        case 13: applied = launch_fwd<5>(c, a, c->stream, do_a, do_b); break;
        case 14: applied = launch_fwd<6>(c, a, c->stream, do_a, do_b); break;
        case 15: applied = launch_fwd<7>(c, a, c->stream, do_a, do_b); break;
        case 16: applied = launch_fwd<8>(c, a, c->stream, do_a, do_b); break;
        default: throw StatusError{ INVALID_ARGUMENT, "unsupported log_n" };
        }
        c->launches += (do_a ? 1 : 0) + (do_b ? 1 : 0);
        MOAI_CUDA_CHECK(cudaGetLastError());
        return applied;
    }

    namespace
    {
        template <int R>
        void launch_ks_mac_multi(const KsMacMultiArgs &a, const NttArgs &na, dim3 grid, cudaStream_t s)
        {
            static const cudaError_t attr =
                cudaFuncSetAttribute(ks_mac_multi_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, ksm_smem_bytes(R));
            (void)attr;
            ks_mac_multi_kernel<R><<<grid, 256, ksm_smem_bytes(R), s>>>(a, na);
        }
    } // namespace

    KsShape ks_shape_seal(Context *c, int limbs)
    {
        KsShape sh;
        sh.digits = limbs;
        sh.rns = limbs + 1;
        sh.n_data = limbs;
        sh.ids = c->d_ids_ks + (size_t)limbs * (c->kl + 1);
        return sh;
    }

    void ks_mac_multi(Context *c, const u64 *ext, long long batch, const KsShape &sh, int n_keys, const u64 *const *ksk,
                      const int *key_kl, u64 *const *acc)
    {
        MOAI_REQUIRE(n_keys >= 1 && n_keys <= KSM_R, "too many keys for one multi-key inner product");
        KsMacMultiArgs a;
        a.ext = ext;
        for (int r = 0; r < KSM_R; r++)
        {
            a.ksk[r] = r < n_keys ? ksk[r] : nullptr;
            a.acc[r] = r < n_keys ? acc[r] : nullptr;
            a.key_kl[r] = r < n_keys ? key_kl[r] : 0;
        }
        a.limb = c->d_limb;
        a.ids_ks = sh.ids;
        a.limbs = sh.digits;
        a.rns = sh.rns;
        a.n_data = sh.n_data;
        a.log_n = c->log_n;
        NttArgs na{ nullptr, c->d_fwd, c->d_fwd_fp, c->d_limb, nullptr, 1, 1, c->log_n, 0 };
        dim3 grid((unsigned)batch, (unsigned)sh.rns, (unsigned)((c->n / 2) / 256));
        KernelTimer kt(c, "k_ks_mac_multi", batch * (long long)n_keys);
        switch (n_keys)
        {
        case 1: launch_ks_mac_multi<1>(a, na, grid, c->stream); break;
        case 2: launch_ks_mac_multi<2>(a, na, grid, c->stream); break;
        case 3: launch_ks_mac_multi<3>(a, na, grid, c->stream); break;
        default: launch_ks_mac_multi<4>(a, na, grid, c->stream); break;
        }
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ntt_forward_pass_b_strided(Context *c, u64 *data, long long groups, long long grp_size, long long grp_stride,
                                    const int *d_limb_id)
    {
        if (groups <= 0 || grp_size <= 0)
        {
            return;
        }
        NttArgs a{ data, c->d_fwd, c->d_fwd_fp, c->d_limb, d_limb_id, 1, 1, c->log_n, groups * grp_size };
        a.grp_size = grp_size;
        a.grp_stride = grp_stride;
        const long long ctas_b = a.count * ((1 << (c->log_n - 8)) / ROWS);
        KernelTimer kt(c, "k_ntt_fwd_pass_b_strided", a.count);
        ntt_fwd_pass_b<<<(unsigned)ctas_b, ROWS * 16, 0, c->stream>>>(a);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ks_passb_mac(Context *c, const u64 *mid, long long batch, const KsShape &sh, const u64 *ksk, int key_kl,
                      u64 *acc, const u64 *direct, const int *own)
    {
        const int limbs = sh.digits;
        MOAI_REQUIRE(c->log_n >= 12, "unsupported log_n");
        static const cudaError_t attr = cudaFuncSetAttribute(
            ks_passb_mac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KS_FUSED_SMEM);
        (void)attr;
        const int rns = sh.rns;
        KsFusedArgs a;
        a.mid = mid;
        a.ksk = ksk;
        a.acc = acc;
        a.tw_fp = c->d_fwd_fp;
        a.limb = c->d_limb;
        a.ids_ks = sh.ids;
        a.limbs = limbs;
        a.rns = rns;
        a.n_data = sh.n_data;
        a.direct = direct;
        a.own = direct ? own : nullptr;
        a.key_kl = key_kl;
        a.log_n = c->log_n;
        NttArgs na{ nullptr, c->d_fwd, c->d_fwd_fp, c->d_limb, nullptr, 1, 1, c->log_n, 0 };
        const int R = 1 << (c->log_n - 8);
        // ciphertexts per CTA: as many as keep >= 8 waves of CTAs on the GPU (3 CTAs per SM), at most 8
        static const int items_max = [] {
            const char *e = getenv("MOAI_KS_ITEMS");
            return e ? atoi(e) : 8;
        }();
        long long items = batch * (R / FR) * rns / ((long long)c->sm_count * 3 * 8);
        items = items < 1 ? 1 : (items > items_max ? items_max : items);
        if (limbs >= 12) // long digit loops amortise the start-up by themselves (measured: 17 digits, 9.6 vs 9.95 ms)
        {
            items = 1;
        }
        a.batch = batch;
        a.items = (int)items;
        dim3 grid((unsigned)((batch + items - 1) / items), (unsigned)(R / FR), (unsigned)rns);
        // units: limb-transforms finished inside the kernel (one per (ciphertext, digit, target modulus))
        // (with `direct`, one digit per FP64-path data target needs no transform: n_data of them at most)
        KernelTimer kt(c, "k_ks_passb_mac", batch * ((long long)limbs * rns - (direct ? sh.n_data : 0)));
        ks_passb_mac_kernel<<<grid, FT, KS_FUSED_SMEM, c->stream>>>(a, na);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    void ntt_inverse(Context *c, u64 *data, long long count, const int *d_limb_ids, int period, int div)
    {
        ntt_inverse_from(c, nullptr, 0, 0, data, count, d_limb_ids, period, div, nullptr);
    }

    NttScale ntt_scale_make(Context *c, int prime, u64 constant)
    {
        typedef unsigned __int128 u128h;
        const LimbConst &lc = c->h_limb[prime];
        const u64 q = lc.q;
        auto mk = [&](u64 base, u64 &w, u64 &wq, double &d) {
            w = (u64)((u128h)base * (constant % q) % q);
            wq = (u64)(((u128h)w << 64) / q);
            d = w > q / 2 ? -(double)(q - w) : (double)w;
        };
        NttScale s;
        mk(lc.inv_n, s.inv_n, s.inv_n_quo, s.inv_n_d);
        mk(lc.inv_n_w, s.inv_n_w, s.inv_n_w_quo, s.inv_n_w_d);
        return s;
    }

    void ntt_inverse_from(Context *c, const u64 *src, long long grp_size, long long grp_stride, u64 *data, long long count,
                          const int *d_limb_ids, int period, int div, const NttScale *d_scale, bool fp_out)
    {
        if (count <= 0)
        {
            return;
        }
        NttArgs a{ data, c->d_inv, c->d_inv_fp, c->d_limb, d_limb_ids, period, div, c->log_n, count };
        a.scale = d_scale;
        a.fp_out = fp_out ? 1 : 0;
        if (src)
        {
            MOAI_REQUIRE(grp_size >= 1 && grp_stride >= grp_size, "bad source layout");
            a.src = src;
            a.grp_size = grp_size;
            a.grp_stride = grp_stride;
        }
        KernelTimer kt(c, "k_ntt_inv", count);
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
