// Rotation-free column-packed ciphertext-plaintext matmul (north_star (3)), fused.
//
// Reference: ct_pt_matrix_mul_wo_pre / _wo_pre_large (M/source/matrix_mul/Ct_pt_matrix_mul.hpp:
// 4-101):  out[i] = rescale( sum_j X[j] (*) encode_scalar(W[j][i]) ).  A scalar-encoded plaintext
// is the constant round(W[j][i] * scale) mod q_l in every NTT slot (S/ckks.cpp:110-153), so per
// limb l and polynomial p this is a modular GEMM
//      Y[i][p][l][t] = sum_j X[j][p][l][t] * Wc[l][j][i]  (mod q_l),   t < N, j < K, i < C
// followed by the rescale of every output ciphertext.  The reference spends K*C encode +
// multiply_plain + add calls on it; here the K-sum is kept in 128-bit registers (products of
// canonical residues < 2^116, K <= 4096 of them fit) and reduced once per output element.
//
// B200 mapping (CUDA-core version): CTA = 256 threads x 2 coefficients (16-byte loads of X) x
// TN = 8 output columns; the weight tile lives in shared memory and is read as broadcasts; grid
// ordered so that CTAs sharing an X tile (same p, l, t-range, different i) are adjacent and the
// re-reads of X hit the 126 MB L2 instead of HBM.  The kernel is bound by the integer pipe
// (4 IMAD.WIDE + carry adds per 64x64->128 MAC), not by HBM.
#include "ntt.cuh"
#include "ops.cuh"
#include <cmath>

namespace moai
{
    namespace
    {
        constexpr int MM_THREADS = 256;
        constexpr int MM_TN = 8;   // output columns per CTA
        constexpr int MM_KC = 256; // K chunk staged in shared memory

        // Wc[l][j][i] = residue of round(W[j][i] * scale) mod q_l (sign-magnitude like SEAL)
        __global__ void k_encode_weights(const double *__restrict__ W, u64 *__restrict__ Wc, long long kc, int limbs,
                                         double scale, const LimbConst *__restrict__ lcs)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= kc)
            {
                return;
            }
            const double v = round(W[i] * scale);
            const bool neg = signbit(v);
            const u64 mag = (u64)fabs(v);
            for (int l = 0; l < limbs; l++)
            {
                const LimbConst lc = lcs[l];
                u64 r = reduce64(mag, lc);
                Wc[(long long)l * kc + i] = neg ? negmod(r, lc.q) : r;
            }
        }

        // Y[i][p][l][t] = sum_j X[j][p][l][t] * Wc[l][j][i]
        __global__ void __launch_bounds__(MM_THREADS)
            k_ctpt_gemm(const ulonglong2 *__restrict__ X, const u64 *__restrict__ Wc, ulonglong2 *__restrict__ Y, int K,
                        int C, int limbs, int log_n2, const LimbConst *__restrict__ lcs,
                        const Twiddle *__restrict__ two64)
        {
            __shared__ u64 ws[MM_KC * MM_TN];
            // blockIdx.x = ((pl * tiles_t) + tile_t) * tiles_i + tile_i   (tile_i fastest)
            const int tiles_i = (C + MM_TN - 1) / MM_TN;
            const int tiles_t = (1 << log_n2) / MM_THREADS;
            const int tile_i = blockIdx.x % tiles_i;
            const int rest = blockIdx.x / tiles_i;
            const int tile_t = rest % tiles_t;
            const int pl = rest / tiles_t; // p * limbs + l
            const int l = pl % limbs;
            const LimbConst lc = lcs[l];
            const Twiddle t64 = two64[l];
            const long long ct_stride2 = (long long)2 * limbs << log_n2; // one ciphertext, in 16-byte units
            const long long off = ((long long)pl << log_n2) + (long long)tile_t * MM_THREADS + threadIdx.x;
            const int i0 = tile_i * MM_TN;

            u128 acc[2][MM_TN];
#pragma unroll
            for (int c = 0; c < MM_TN; c++)
            {
                acc[0][c] = u128{ 0, 0 };
                acc[1][c] = u128{ 0, 0 };
            }
            const u64 *wl = Wc + (long long)l * K * C;
            for (int j0 = 0; j0 < K; j0 += MM_KC)
            {
                const int jn = min(MM_KC, K - j0);
                __syncthreads();
                for (int e = threadIdx.x; e < jn * MM_TN; e += MM_THREADS)
                {
                    const int jj = e / MM_TN, cc = e % MM_TN;
                    ws[e] = (i0 + cc < C) ? wl[(long long)(j0 + jj) * C + i0 + cc] : 0;
                }
                __syncthreads();
                const ulonglong2 *xp = X + (long long)j0 * ct_stride2 + off;
#pragma unroll 2
                for (int jj = 0; jj < jn; jj++)
                {
                    const ulonglong2 x = xp[(long long)jj * ct_stride2];
                    const ulonglong2 *wrow = reinterpret_cast<const ulonglong2 *>(ws + jj * MM_TN);
#pragma unroll
                    for (int c = 0; c < MM_TN; c += 2)
                    {
                        const ulonglong2 w = wrow[c >> 1];
                        mac_wide(acc[0][c], x.x, w.x);
                        mac_wide(acc[1][c], x.y, w.x);
                        mac_wide(acc[0][c + 1], x.x, w.y);
                        mac_wide(acc[1][c + 1], x.y, w.y);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < MM_TN; c++)
            {
                if (i0 + c < C)
                {
                    ulonglong2 r;
                    r.x = barrett_reduce_acc(acc[0][c], lc, t64.w, t64.wq);
                    r.y = barrett_reduce_acc(acc[1][c], lc, t64.w, t64.wq);
                    Y[(long long)(i0 + c) * ct_stride2 + off] = r;
                }
            }
        }
    } // namespace

    // X: [K][2][limbs][n] device; W: host row-major K x C doubles; out: [C][2][limbs-1][n] device
    void ct_pt_matmul_scalar(Context *c, const u64 *X, const double *h_W, int K, int C, int limbs, double scale,
                             u64 *out)
    {
        MOAI_REQUIRE(K >= 1 && C >= 1, "bad dimensions of X or W");
        MOAI_REQUIRE(limbs >= 2 && limbs <= c->kl - 1, "end of modulus switching chain reached");
        MOAI_REQUIRE(K <= 4096, "K too large for the 128-bit lazy accumulator");
        const size_t n = c->n;
        const long long kc = (long long)K * C;
        double wmax = 0;
        for (long long i = 0; i < kc; i++)
        {
            wmax = std::fmax(wmax, std::fabs(h_W[i]));
        }
        MOAI_REQUIRE(wmax * scale < 9.0e18, "encoded value is too large");
        Scratch dW(kc * sizeof(double), c->stream);
        Scratch dWc((size_t)limbs * kc * sizeof(u64), c->stream);
        Scratch Y((size_t)C * 2 * limbs * n * sizeof(u64), c->stream);
        MOAI_CUDA_CHECK(cudaMemcpyAsync(dW.p, h_W, kc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        k_encode_weights<<<(unsigned)((kc + 255) / 256), 256, 0, c->stream>>>(dW.as<double>(), dWc.as<u64>(), kc, limbs,
                                                                            scale, c->d_limb);
        c->launches += 2;
        const int tiles_i = (C + MM_TN - 1) / MM_TN;
        const int tiles_t = (int)((n / 2) / MM_THREADS);
        const long long ctas = (long long)2 * limbs * tiles_t * tiles_i;
        {
            PhaseTimer pt(c, "ctpt_gemm");
            k_ctpt_gemm<<<(unsigned)ctas, MM_THREADS, 0, c->stream>>>(reinterpret_cast<const ulonglong2 *>(X),
                                                                     dWc.as<u64>(), Y.as<ulonglong2>(), K, C, limbs,
                                                                     c->log_n - 1, c->d_limb, c->d_two64);
        }
        MOAI_CUDA_CHECK(cudaGetLastError());
        rescale(c, Y.as<u64>(), out, C, 2, limbs);
    }
} // namespace moai
