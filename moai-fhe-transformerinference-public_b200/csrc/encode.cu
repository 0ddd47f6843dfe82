// CKKS vector encoder on the device: CKKSEncoder::encode_internal(vector)
// S/ckks.h:457-638 — slots -> conjugate-symmetric array through the index map -> in-place complex
// inverse DWT with fix = scale / N -> round(real part) -> residue per limb -> forward NTT.
//
// The FP64 butterflies reproduce the reference's operation order exactly (FFTHandler::
// transform_from_rev with std::complex<double>, S/util/dwthandler.h:202-356): separate IEEE
// multiply / add / subtract, never fused (the __d*_rn intrinsics are never contracted to FMA),
// complex product = (ac - bd, ad + bc), same root table (S/ckks.cpp:56-66, built on the host with
// the same libm).  The plaintexts are therefore bit-identical to SEAL's.
// Kernel structure = the inverse NTT's: pass B' (8 stages inside 256-element rows, 4+4 register
// stages, padded smem transpose) then pass A' (row-pairing stages, 16-column tiles).
#include "ntt.cuh"
#include "ops.cuh"

namespace moai
{
    namespace
    {
        constexpr int TB = 8;        // columns per pass-A tile (8 x 16 B = one 128-byte line)
        constexpr int ROWS = 8;      // rows per pass-B CTA (16 B elements: half the NTT's row count)
        constexpr int ROW_PAD = 272;

        __device__ __forceinline__ double2 cadd(double2 a, double2 b)
        {
            return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y));
        }
        __device__ __forceinline__ double2 csubc(double2 a, double2 b)
        {
            return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y));
        }
        __device__ __forceinline__ double2 cmul(double2 a, double2 b)
        {
            return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
                                __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
        }
        __device__ __forceinline__ double2 cscale(double2 a, double s)
        {
            return make_double2(__dmul_rn(a.x, s), __dmul_rn(a.y, s));
        }

        template <int GAP>
        __device__ __forceinline__ void fgs_stage(double2 (&x)[16], const double2 *__restrict__ tw)
        {
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                if (!(k & GAP))
                {
                    const double2 r = __ldg(tw + k / (2 * GAP));
                    const double2 u = x[k], v = x[k + GAP];
                    x[k] = cadd(u, v);
                    x[k + GAP] = cmul(csubc(u, v), r);
                }
            }
        }

        __device__ __forceinline__ void fgs_stage_last(double2 (&x)[16], double2 r, double fix)
        {
            const double2 sr = cscale(r, fix);
#pragma unroll
            for (int k = 0; k < 8; k++)
            {
                const double2 u = x[k], v = x[k + 8];
                x[k] = cscale(cadd(u, v), fix);
                x[k + 8] = cmul(csubc(u, v), sr);
            }
        }

        // data: [count][n] double2, in place
        __global__ void __launch_bounds__(ROWS * 16)
            fft_inv_pass_b(double2 *__restrict__ data, const double2 *__restrict__ roots, int log_n)
        {
            __shared__ double2 sm[ROWS * ROW_PAD];
            const int t = threadIdx.x & 15, r = threadIdx.x >> 4;
            const int R = 1 << (log_n - 8);
            const size_t n = (size_t)1 << log_n;
            const int ctas_per_poly = R / ROWS;
            const long long poly = blockIdx.x / ctas_per_poly;
            const int row = (blockIdx.x % ctas_per_poly) * ROWS + r;
            double2 *base = data + ((size_t)poly << log_n) + (size_t)row * 256;
            double2 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = base[16 * t + k];
            }
            fgs_stage<1>(x, roots + (n - n + 1) + (size_t)row * 128 + 8 * t);
            fgs_stage<2>(x, roots + (n - n / 2 + 1) + (size_t)row * 64 + 4 * t);
            fgs_stage<4>(x, roots + (n - n / 4 + 1) + (size_t)row * 32 + 2 * t);
            fgs_stage<8>(x, roots + (n - n / 8 + 1) + (size_t)row * 16 + t);
            double2 *srow = sm + r * ROW_PAD;
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
            fgs_stage<1>(x, roots + (n - n / 16 + 1) + (size_t)row * 8);
            fgs_stage<2>(x, roots + (n - n / 32 + 1) + (size_t)row * 4);
            fgs_stage<4>(x, roots + (n - n / 64 + 1) + (size_t)row * 2);
            fgs_stage<8>(x, roots + (n - n / 128 + 1) + (size_t)row);
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                base[t + 16 * k] = x[k];
            }
        }

        // Pass A' + epilogue: the real parts are rounded and reduced into every limb.
        //   out: [count][limbs][n] uint64 (coefficient form; NTT follows)
        template <int LOGR>
        __global__ void __launch_bounds__((1 << LOGR) / 16 * TB)
            fft_inv_pass_a(const double2 *__restrict__ data, const double2 *__restrict__ roots, double fix, int limbs,
                           const LimbConst *__restrict__ lcs, const Twiddle *__restrict__ two64, u64 *__restrict__ out,
                           int *__restrict__ overflow)
        {
            constexpr int R = 1 << LOGR, T1 = R / 16;
            constexpr int log_n = LOGR + 8;
            __shared__ double2 sm[R * TB];
            const int tb = threadIdx.x & (TB - 1), t = threadIdx.x >> 3;
            const size_t n = (size_t)1 << log_n;
            const long long poly = blockIdx.x / (256 / TB);
            const int tile = blockIdx.x % (256 / TB);
            const double2 *base = data + ((size_t)poly << log_n) + tile * TB + tb;
            double2 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                x[k] = base[(size_t)(16 * t + k) * 256];
            }
            fgs_stage<1>(x, roots + (n - R + 1) + 8 * t);
            fgs_stage<2>(x, roots + (n - R / 2 + 1) + 4 * t);
            fgs_stage<4>(x, roots + (n - R / 4 + 1) + 2 * t);
            const double2 last_root = __ldg(roots + n - 1);
            if constexpr (LOGR == 4)
            {
                fgs_stage_last(x, last_root, fix);
            }
            else
            {
                fgs_stage<8>(x, roots + (n - R / 8 + 1) + t);
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
                if constexpr (256 / R <= 1)
                {
                    fgs_stage<1>(x, roots + (n - R / T1 + 1));
                }
                if constexpr (256 / R <= 2)
                {
                    fgs_stage<2>(x, roots + (n - R / (2 * T1) + 1));
                }
                if constexpr (256 / R <= 4)
                {
                    fgs_stage<4>(x, roots + (n - R / (4 * T1) + 1));
                }
                fgs_stage_last(x, last_root, fix);
            }
            // round, sign-magnitude, residues (S/ckks.h:536-600; <= 64-bit and <= 128-bit paths agree)
            const double two64d = 18446744073709551616.0;
#pragma unroll
            for (int k = 0; k < 16; k++)
            {
                const int row = (LOGR == 4) ? (16 * t + k) : (t + T1 * k);
                const double c = round(x[k].x);
                const bool neg = signbit(c);
                const double mag = fabs(c);
                u64 lo, hi;
                if (mag < two64d)
                {
                    lo = (u64)mag;
                    hi = 0;
                }
                else
                {
                    if (!(mag < 3.4028236692093846e38))
                    {
                        *overflow = 1; // > 128 bits: SEAL switches to multi-precision decompose
                    }
                    hi = (u64)(mag / two64d);
                    lo = (u64)fmod(mag, two64d);
                }
                u64 *dst = out + ((size_t)poly * limbs << log_n) + (size_t)row * 256 + tile * TB + tb;
                for (int l = 0; l < limbs; l++)
                {
                    const LimbConst lc = lcs[l];
                    u128 z;
                    z.lo = lo;
                    z.hi = hi;
                    const Twiddle t64 = two64[l];
                    u64 res = hi ? barrett_reduce_acc(z, lc, t64.w, t64.wq) : reduce64(lo, lc);
                    dst[(size_t)l << log_n] = neg ? negmod(res, lc.q) : res;
                }
            }
        }

        // cv[p][index_map[i]] = v_i, cv[p][index_map[i + slots]] = conj(v_i)   (S/ckks.h:503-508)
        // values: [count][n_vals] complex (device); slots beyond n_vals stay zero
        __global__ void k_fill_conj(const double2 *__restrict__ values, int n_vals, long long vstride,
                                    double2 *__restrict__ cv, long long total, int log_n,
                                    const uint32_t *__restrict__ index_map)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // over [count][slots]
            if (i >= total)
            {
                return;
            }
            const int slots = 1 << (log_n - 1);
            const long long p = i >> (log_n - 1);
            const int s = (int)(i & (slots - 1));
            double2 v = make_double2(0.0, 0.0);
            if (s < n_vals)
            {
                v = values[p * vstride + s];
            }
            double2 *dst = cv + ((size_t)p << log_n);
            dst[index_map[s]] = v;
            dst[index_map[s + slots]] = make_double2(v.x, -v.y);
        }

        // cv for the masked weights of ct_pt_matrix_mul_wo_pre_w_mask: value(s) = mask[s] == 1 ? w_p : 0
        __global__ void k_fill_conj_masked(const double *__restrict__ w, const int *__restrict__ mask,
                                           double2 *__restrict__ cv, long long total, int log_n,
                                           const uint32_t *__restrict__ index_map)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total)
            {
                return;
            }
            const int slots = 1 << (log_n - 1);
            const long long p = i >> (log_n - 1);
            const int s = (int)(i & (slots - 1));
            const double val = mask[s] == 1 ? w[p] : 0.0;
            double2 *dst = cv + ((size_t)p << log_n);
            dst[index_map[s]] = make_double2(val, 0.0);
            dst[index_map[s + slots]] = make_double2(val, -0.0);
        }

        template <int LOGR>
        void launch_fft(Context *c, double2 *cv, long long count, double fix, int limbs, u64 *out, int *d_overflow)
        {
            fft_inv_pass_b<<<(unsigned)(count * ((1 << LOGR) / ROWS)), ROWS * 16, 0, c->stream>>>(cv, c->d_fft_inv_roots,
                                                                                                 c->log_n);
            fft_inv_pass_a<LOGR><<<(unsigned)(count * (256 / TB)), (1 << LOGR) / 16 * TB, 0, c->stream>>>(
                cv, c->d_fft_inv_roots, fix, limbs, c->d_limb, c->d_two64, out, d_overflow);
            c->launches += 2;
        }

        // cv: [count][n] filled conj arrays (destroyed) -> out [count][limbs][n] NTT form
        void encode_from_conj(Context *c, double2 *cv, long long count, double scale, int limbs, u64 *out)
        {
            MOAI_REQUIRE(scale > 0, "scale out of bounds");
            const double fix = scale / (double)c->n;
            Scratch flag(sizeof(int), c->stream);
            MOAI_CUDA_CHECK(cudaMemsetAsync(flag.p, 0, sizeof(int), c->stream));
            switch (c->log_n)
            {
            case 12: launch_fft<4>(c, cv, count, fix, limbs, out, flag.as<int>()); break;
            case 13: launch_fft<5>(c, cv, count, fix, limbs, out, flag.as<int>()); break;
            case 14: launch_fft<6>(c, cv, count, fix, limbs, out, flag.as<int>()); break;
            case 15: launch_fft<7>(c, cv, count, fix, limbs, out, flag.as<int>()); break;
            case 16: launch_fft<8>(c, cv, count, fix, limbs, out, flag.as<int>()); break;
            default: throw StatusError{ INVALID_ARGUMENT, "unsupported log_n" };
            }
            MOAI_CUDA_CHECK(cudaGetLastError());
            ntt_forward(c, out, count * limbs, c->d_ids, limbs);
            int h = 0;
            MOAI_CUDA_CHECK(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            MOAI_REQUIRE(h == 0, "encoded values are too large");
        }
    } // namespace

    // values: DEVICE [count][n_vals] complex (interleaved re, im); out: [count][limbs][n]
    void encode_vector(Context *c, const double *d_values, long long count, int n_vals, double scale, int limbs,
                       u64 *out)
    {
        MOAI_REQUIRE(n_vals >= 0 && (size_t)n_vals <= c->n / 2, "values_size is too large");
        if (!count)
        {
            return;
        }
        Scratch cv((size_t)count * c->n * sizeof(double2), c->stream);
        const long long total = count * (long long)(c->n / 2);
        k_fill_conj<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(
            reinterpret_cast<const double2 *>(d_values), n_vals, n_vals, cv.as<double2>(), total, c->log_n,
            c->d_index_map);
        c->launches += 1;
        encode_from_conj(c, cv.as<double2>(), count, scale, limbs, out);
    }

    // plaintexts of the masked ct-pt matmul: pt[p] = encode(w[p] * mask), p < count
    void encode_masked_weights(Context *c, const double *d_w, const int *d_mask, long long count, double scale,
                               int limbs, u64 *out)
    {
        if (!count)
        {
            return;
        }
        Scratch cv((size_t)count * c->n * sizeof(double2), c->stream);
        const long long total = count * (long long)(c->n / 2);
        k_fill_conj_masked<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(d_w, d_mask, cv.as<double2>(), total,
                                                                                  c->log_n, c->d_index_map);
        c->launches += 1;
        encode_from_conj(c, cv.as<double2>(), count, scale, limbs, out);
    }
} // namespace moai
