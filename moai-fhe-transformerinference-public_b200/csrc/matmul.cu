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
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace moai
{
    namespace
    {
        constexpr int MM_THREADS = 256;
        constexpr int MM_TN = 8;   // output columns per CTA
        constexpr int MM_KC = 256; // K chunk staged in shared memory

        // Wc[l][j][i] = residue of round(W[j][i] * scale) mod q_l (sign-magnitude like SEAL)
        __global__ void k_encode_weights(const double *__restrict__ W, u64 *__restrict__ Wc, long long kc, int limbs,
                                         double scale, const LimbConst *__restrict__ lcs, int split26)
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
                r = neg ? negmod(r, lc.q) : r;
                // split26: low 26 bits in the low word, the rest in the high word (see k_ctpt_gemm26)
                Wc[(long long)l * kc + i] = split26 ? (((r >> 26) << 32) | (r & 0x3FFFFFFull)) : r;
            }
        }

        // Y[i][p][l][t] = sum_j X[j][p][l][t] * Wc[l][j][i]
        __global__ void __launch_bounds__(MM_THREADS)
            k_ctpt_gemm(const ulonglong2 *__restrict__ X, const u64 *__restrict__ Wc, ulonglong2 *__restrict__ Y, int K,
                        int C, int ldw, int limbs, int log_n2, const LimbConst *__restrict__ lcs,
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
            const u64 *wl = Wc + (long long)l * K * ldw;
            for (int j0 = 0; j0 < K; j0 += MM_KC)
            {
                const int jn = min(MM_KC, K - j0);
                __syncthreads();
                for (int e = threadIdx.x; e < jn * MM_TN; e += MM_THREADS)
                {
                    const int jj = e / MM_TN, cc = e % MM_TN;
                    ws[e] = (i0 + cc < C) ? wl[(long long)(j0 + jj) * ldw + i0 + cc] : 0;
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

        // Same GEMM for primes below 2^52 (all data primes of the repo's chain are <= 51 bits).
        // Residues are split at 26 bits: x = x1 * 2^26 + x0, w = w1 * 2^26 + w0 (the weights arrive
        // pre-split, one half per 32-bit word).  The three columns
        //      c0 += x0 w0,   c1 += x0 w1 + x1 w0,   c2 += x1 w1
        // are plain 64-bit accumulators: every product is < 2^52, so up to 1024 terms fit without a
        // carry chain.  One MAC = exactly four IMAD.WIDE.U32 with accumulate and nothing else —
        // the integer pipe's minimum for a 52x52-bit product; the carries are paid once per 1024 j.
        constexpr int MM_FOLD = 1024;
        template <int TM>
        __global__ void __launch_bounds__(MM_THREADS, TM == 1 ? 2 : 1)
            k_ctpt_gemm26(const u64 *__restrict__ Xw, const u64 *__restrict__ Wc, u64 *__restrict__ Yw,
                          int K, int C, int ldw, int limbs, int log_n, const LimbConst *__restrict__ lcs,
                          const Twiddle *__restrict__ two64)
        {
            // TM consecutive coefficients per thread; strides below are in units of TM words
            typedef typename std::conditional<TM == 2, ulonglong2, u64>::type vec_t;
            const vec_t *X = reinterpret_cast<const vec_t *>(Xw);
            vec_t *Y = reinterpret_cast<vec_t *>(Yw);
            const int log_n2 = log_n - (TM == 2 ? 1 : 0);
            __shared__ u64 ws[MM_KC * MM_TN];
            const int tiles_i = (C + MM_TN - 1) / MM_TN;
            const int tiles_t = (1 << log_n2) / MM_THREADS;
            const int tile_i = blockIdx.x % tiles_i;
            const int rest = blockIdx.x / tiles_i;
            const int tile_t = rest % tiles_t;
            const int pl = rest / tiles_t; // p * limbs + l
            const int l = pl % limbs;
            const LimbConst lc = lcs[l];
            const Twiddle t64 = two64[l];
            const long long ct_stride2 = (long long)2 * limbs << log_n2;
            const long long off = ((long long)pl << log_n2) + (long long)tile_t * MM_THREADS + threadIdx.x;
            const int i0 = tile_i * MM_TN;

            u64 c0[TM][MM_TN], c1[TM][MM_TN], c2[TM][MM_TN];
            u64 part[TM][MM_TN]; // canonical partial sums of the folded chunks
#pragma unroll
            for (int c = 0; c < MM_TN; c++)
            {
#pragma unroll
                for (int e = 0; e < TM; e++)
                {
                    c0[e][c] = c1[e][c] = c2[e][c] = 0;
                    part[e][c] = 0;
                }
            }
            auto fold = [&]() {
#pragma unroll
                for (int c = 0; c < MM_TN; c++)
                {
#pragma unroll
                    for (int e = 0; e < TM; e++)
                    {
                        // part += (c0 + c1 * 2^26 + c2 * 2^52) mod q
                        u64 lo = c0[e][c], hi = 0;
                        u64 t = c1[e][c] << 26;
                        lo += t;
                        hi += (lo < t) + (c1[e][c] >> 38);
                        t = c2[e][c] << 52;
                        lo += t;
                        hi += (lo < t) + (c2[e][c] >> 12);
                        part[e][c] = addmod(part[e][c], barrett_reduce_acc(u128{ lo, hi }, lc, t64.w, t64.wq), lc.q);
                        c0[e][c] = c1[e][c] = c2[e][c] = 0;
                    }
                }
            };
            const u64 *wl = Wc + (long long)l * K * ldw;
            for (int j0 = 0; j0 < K; j0 += MM_KC)
            {
                const int jn = min(MM_KC, K - j0);
                __syncthreads();
                for (int e = threadIdx.x; e < jn * MM_TN; e += MM_THREADS)
                {
                    const int jj = e / MM_TN, cc = e % MM_TN;
                    ws[e] = (i0 + cc < C) ? wl[(long long)(j0 + jj) * ldw + i0 + cc] : 0;
                }
                __syncthreads();
                if (j0 && (j0 % MM_FOLD) == 0)
                {
                    fold();
                }
                const vec_t *xp = X + (long long)j0 * ct_stride2 + off;
#pragma unroll 2
                for (int jj = 0; jj < jn; jj++)
                {
                    const vec_t xv = xp[(long long)jj * ct_stride2];
                    u32 x0[TM], x1[TM];
                    if constexpr (TM == 2)
                    {
                        x0[0] = (u32)xv.x & 0x3FFFFFFu;
                        x1[0] = (u32)(xv.x >> 26);
                        x0[1] = (u32)xv.y & 0x3FFFFFFu;
                        x1[1] = (u32)(xv.y >> 26);
                    }
                    else
                    {
                        x0[0] = (u32)xv & 0x3FFFFFFu;
                        x1[0] = (u32)(xv >> 26);
                    }
                    const uint2 *wrow = reinterpret_cast<const uint2 *>(ws + jj * MM_TN);
#pragma unroll
                    for (int c = 0; c < MM_TN; c++)
                    {
                        const uint2 w = wrow[c]; // (w0, w1): low 26 bits, remaining bits
#pragma unroll
                        for (int e = 0; e < TM; e++)
                        {
                            c0[e][c] += (u64)x0[e] * w.x;
                            c1[e][c] += (u64)x0[e] * w.y;
                            c1[e][c] += (u64)x1[e] * w.x;
                            c2[e][c] += (u64)x1[e] * w.y;
                        }
                    }
                }
            }
            fold();
#pragma unroll
            for (int c = 0; c < MM_TN; c++)
            {
                if (i0 + c < C)
                {
                    if constexpr (TM == 2)
                    {
                        ulonglong2 r;
                        r.x = part[0][c];
                        r.y = part[1][c];
                        Y[(long long)(i0 + c) * ct_stride2 + off] = r;
                    }
                    else
                    {
                        Y[(long long)(i0 + c) * ct_stride2 + off] = part[0][c];
                    }
                }
            }
        }

        // ======================================================================================
        // Tensor-core version (north_star (3): "int8 limb-split").  Residues and weights are split into
        // byte planes, x = sum_a 2^(8a) x_a, w = sum_b 2^(8b) w_b, and the modular GEMM becomes NP^2
        // unsigned 8-bit GEMMs whose s32 results are summed per diagonal s = a + b:
        //      D_s[m][i] = sum_j sum_{a+b=s} x_a[j][m] * w_b[j][i]      (< NP * K * 255^2 < 2^31 for K <= 4096)
        //      Y[m][i]   = (sum_s 2^(8s) D_s[m][i]) mod q               (exact: same canonical residue)
        // One warp owns a 16 (coefficients) x 16 (output columns) tile: 2 n-tiles x (2 NP - 1) diagonals
        // x 4 = 104 accumulator registers for NP = 7, and issues NP^2 mma.sync.m16n8k32.u8.u8 per n-tile
        // and 32-deep k-step.  The A fragments are built on the fly from the ciphertexts (16 coalesced
        // 64-bit loads + two 4x4 byte transposes per fragment register set, PRMT on the ALU pipe, hidden
        // under the tensor pipe); the weights arrive pre-packed in fragment order (k_pack_weights).
        // Measured basis (tools/microbench_mma.cu on B200): mma.sync u8 571 T MAC/s -> 11.7 T modular
        // MAC/s at NP = 7, against 1.18 T for the IMAD kernel above.
        // ======================================================================================
        constexpr int IM_WARPS_M = 4, IM_WARPS_N = 2;               // CTA = 8 warps: 64 coefficients x 32 columns
        constexpr int IM_TILE_M = 16 * IM_WARPS_M, IM_TILE_N = 16 * IM_WARPS_N;

        // Wp[l][ks][nt][b][lane] (uint2): B fragments of byte plane b of Wc[l][32 ks .. +31][8 nt .. +7]
        __global__ void k_pack_weights(const double *__restrict__ W, uint2 *__restrict__ Wp, int K, int C, int Kp, int Cp,
                                       int limbs, int np, double scale, const LimbConst *__restrict__ lcs)
        {
            const long long total = (long long)limbs * (Kp / 32) * (Cp / 8) * 32;
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total)
            {
                return;
            }
            const int lane = (int)(i % 32);
            const long long r1 = i / 32;
            const int nt = (int)(r1 % (Cp / 8));
            const long long r2 = r1 / (Cp / 8);
            const int ks = (int)(r2 % (Kp / 32));
            const int l = (int)(r2 / (Kp / 32));
            const LimbConst lc = lcs[l];
            const int g = lane >> 2, t = lane & 3;
            const int col = nt * 8 + g;
            u64 w[8];
#pragma unroll
            for (int e = 0; e < 8; e++)
            {
                const int j = ks * 32 + (e >> 2) * 16 + 4 * t + (e & 3);
                u64 r = 0;
                if (j < K && col < C)
                {
                    // residue of round(W * scale) mod q_l, sign-magnitude like SEAL (S/ckks.cpp:110-153)
                    const double v = round(W[(long long)j * C + col] * scale);
                    const u64 mag = (u64)fabs(v);
                    r = reduce64(mag, lc);
                    r = signbit(v) ? negmod(r, lc.q) : r;
                }
                w[e] = r;
            }
            for (int b = 0; b < np; b++)
            {
                uint2 o;
                o.x = (u32)((w[0] >> (8 * b)) & 0xFF) | ((u32)((w[1] >> (8 * b)) & 0xFF) << 8) |
                      ((u32)((w[2] >> (8 * b)) & 0xFF) << 16) | ((u32)((w[3] >> (8 * b)) & 0xFF) << 24);
                o.y = (u32)((w[4] >> (8 * b)) & 0xFF) | ((u32)((w[5] >> (8 * b)) & 0xFF) << 8) |
                      ((u32)((w[6] >> (8 * b)) & 0xFF) << 16) | ((u32)((w[7] >> (8 * b)) & 0xFF) << 24);
                Wp[((((long long)l * (Kp / 32) + ks) * (Cp / 8) + nt) * np + b) * 32 + lane] = o;
            }
        }

        __device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel)
        {
            u32 r;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
            return r;
        }

        // r[e] = four bytes of element e  ->  p[i] = byte i of elements 0..3 (element 0 in the low byte)
        __device__ __forceinline__ void transpose4x4(const u32 (&r)[4], u32 (&p)[4])
        {
            const u32 t0 = prmt(r[0], r[1], 0x5140), t1 = prmt(r[0], r[1], 0x7362);
            const u32 t2 = prmt(r[2], r[3], 0x5140), t3 = prmt(r[2], r[3], 0x7362);
            p[0] = prmt(t0, t2, 0x5410);
            p[1] = prmt(t0, t2, 0x7632);
            p[2] = prmt(t1, t3, 0x5410);
            p[3] = prmt(t1, t3, 0x7632);
        }

        __device__ __forceinline__ void imma_u8(int (&d)[4], const u32 (&a)[4], const uint2 &b)
        {
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                         "{%0,%1,%2,%3};"
                         : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
        }

        __device__ __forceinline__ void mm_cp_async16(void *smem, const void *gmem, bool valid)
        {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
            const int bytes = valid ? 16 : 0; // src-size 0: the 16 bytes are zero-filled
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(bytes) : "memory");
        }

        constexpr int IM_STAGES = 3;
        constexpr int IM_A_BYTES = 32 * IM_TILE_M * 8; // one k-step of raw residues: [32 j][64 m] u64, swizzled
        __host__ __device__ constexpr int im_b_bytes(int np)
        {
            return (IM_TILE_N / 8) * np * 32 * 8; // B fragments of the CTA's n-tiles, all planes
        }
        __host__ __device__ constexpr int im_smem_bytes(int np)
        {
            return IM_STAGES * (IM_A_BYTES + im_b_bytes(np));
        }

        // NP = byte planes per residue (6 for primes below 2^48, 7 below 2^56).
        // X: [K][2][limbs][n], Y: [C][2][limbs][n] (pre-rescale), one launch per class of limbs (limb_mask).
        // Operands of the next two k-steps travel global -> shared memory with cp.async (3-stage ring):
        // the raw residues of the CTA's 64 coefficients x 32 ciphertexts (rows of 512 B, 32-byte chunks
        // XOR-swizzled with (j >> 2) & 3 so that the fragment gathers are bank-conflict free) and the
        // pre-packed weight fragments of its 4 n-tiles.
        template <int NP>
        __global__ void __launch_bounds__(256, 1)
            k_ctpt_gemm_imma(const u64 *__restrict__ X, const uint2 *__restrict__ Wp, u64 *__restrict__ Y, int K, int C,
                             int Kp, int Cp, int wnp, int tiles_n, int limbs, int log_n,
                             const LimbConst *__restrict__ lcs, const Twiddle *__restrict__ two64,
                             unsigned long long limb_mask, int pl_first)
        {
            // Cp, wnp: padded column count and planes per entry of the packed weight layout;
            // C, tiles_n: columns / 32-column tiles of this launch (a column chunk of the layout);
            // pl_first: first (polynomial, limb) slice of this launch (the grid spans consecutive ones)
            constexpr int ND = 2 * NP - 1;
            constexpr int B_BYTES = im_b_bytes(NP);
            extern __shared__ __align__(16) unsigned char im_smem[];
            const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            const int g = lane >> 2, t = lane & 3;
            const int wm = warp % IM_WARPS_M, wn = warp / IM_WARPS_M;
            // blockIdx.x = ((pl * tiles_m) + tile_m) * tiles_n + tile_n   (tile_n fastest: X tile stays in L2)
            const int tiles_m = (1 << log_n) / IM_TILE_M;
            const int tile_n = blockIdx.x % tiles_n;
            const int rest = blockIdx.x / tiles_n;
            const int tile_m = rest % tiles_m;
            const int pl = pl_first + rest / tiles_m; // p * limbs + l
            const int l = pl % limbs;
            if (!((limb_mask >> l) & 1))
            {
                return;
            }
            const size_t n = (size_t)1 << log_n;
            const size_t ct_stride = (size_t)2 * limbs * n;
            const size_t m_cta = (size_t)tile_m * IM_TILE_M;
            const size_t m0 = m_cta + 16 * wm + g; // this thread's first row (second = +8)
            const int nt_cta = tile_n * (IM_TILE_N / 8);
            const int nt0 = nt_cta + 2 * wn; // first of this warp's two n-tiles
            const u64 *xg = X + (size_t)pl * n + m_cta;
            const uint2 *wg = Wp + (((size_t)l * (Kp / 32)) * (Cp / 8) + nt_cta) * wnp * 32;
            const size_t wp_ks = (size_t)(Cp / 8) * wnp * 32;
            const int nks = Kp / 32;

            // stage loader: A rows j = tid / 8 (32 rows), four 16-byte chunks each; B: contiguous B_BYTES
            auto issue_stage = [&](int ks) {
                unsigned char *sa = im_smem + (ks % IM_STAGES) * (IM_A_BYTES + B_BYTES);
                unsigned char *sb = sa + IM_A_BYTES;
                const int jr = tid >> 3, c4 = tid & 7;
                const int j = ks * 32 + jr;
                const u64 *src = xg + (size_t)(j < K ? j : 0) * ct_stride;
                const int swz = ((jr >> 2) & 3) << 1; // XOR on the 32-byte chunk index = bits 1-2 of the 16-byte one
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    const int ch = c4 * 4 + u;
                    mm_cp_async16(sa + jr * 512 + ((ch ^ swz) << 4), src + 2 * ch, j < K);
                }
                const unsigned char *bsrc = reinterpret_cast<const unsigned char *>(wg + (size_t)ks * wp_ks);
                for (int off = tid * 16; off < B_BYTES; off += 256 * 16)
                {
                    // layout planes per entry may exceed NP (7 stored, 6 used): copy plane by plane
                    const int e = off / (NP * 256), rem = off % (NP * 256); // n-tile, byte inside its NP planes
                    mm_cp_async16(sb + off, bsrc + (size_t)e * wnp * 256 + rem, true);
                }
            };

            int acc[2][ND][4];
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
#pragma unroll
                for (int s = 0; s < ND; s++)
                {
                    acc[q][s][0] = acc[q][s][1] = acc[q][s][2] = acc[q][s][3] = 0;
                }
            }
            issue_stage(0);
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (nks > 1)
            {
                issue_stage(1);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            for (int ks = 0; ks < nks; ks++)
            {
                asm volatile("cp.async.wait_group 1;" ::: "memory"); // stage ks has landed
                __syncthreads();                                      // ... for every thread; stage ks - 1 is free
                if (ks + 2 < nks)
                {
                    issue_stage(ks + 2);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                const unsigned char *sa = im_smem + (ks % IM_STAGES) * (IM_A_BYTES + B_BYTES);
                const uint2 *sb = reinterpret_cast<const uint2 *>(sa + IM_A_BYTES);
                // A fragments: planes[a][kk * 2 + rr] for k-half kk and row g + 8 rr
                u32 A[NP][4];
#pragma unroll
                for (int kk = 0; kk < 2; kk++)
                {
#pragma unroll
                    for (int rr = 0; rr < 2; rr++)
                    {
                        u32 lo[4], hi[4], pl4[4], ph4[4];
                        const int mloc = 16 * wm + g + 8 * rr;
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            const int jr = kk * 16 + 4 * t + e; // (jr >> 2) & 3 == t
                            const u64 v = *reinterpret_cast<const u64 *>(sa + jr * 512 + (((mloc >> 2) ^ t) << 5) +
                                                                         ((mloc & 3) << 3));
                            lo[e] = (u32)v;
                            hi[e] = (u32)(v >> 32);
                        }
                        transpose4x4(lo, pl4);
                        transpose4x4(hi, ph4);
#pragma unroll
                        for (int a = 0; a < NP; a++)
                        {
                            A[a][kk * 2 + rr] = a < 4 ? pl4[a] : ph4[a - 4];
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 2; q++)
                {
                    uint2 B[NP];
#pragma unroll
                    for (int b = 0; b < NP; b++)
                    {
                        B[b] = sb[((2 * wn + q) * NP + b) * 32 + lane];
                    }
#pragma unroll
                    for (int a = 0; a < NP; a++)
                    {
#pragma unroll
                        for (int b = 0; b < NP; b++)
                        {
                            imma_u8(acc[q][a + b], A[a], B[b]);
                        }
                    }
                }
            }
            // epilogue: Y = (sum_s 2^(8s) D_s) mod q; c0,c1: row g, columns 2t, 2t+1; c2,c3: row g + 8
            const LimbConst lc = lcs[l];
            const Twiddle t64 = two64[l];
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    u64 lo = 0, hi = 0;
#pragma unroll
                    for (int s = 0; s < ND; s++)
                    {
                        const u64 v = (u64)(u32)acc[q][s][r];
                        const int sh = 8 * s;
                        if (sh == 0)
                        {
                            lo = v;
                        }
                        else if (sh < 64)
                        {
                            const u64 add = v << sh;
                            lo += add;
                            hi += (lo < add) + (v >> (64 - sh));
                        }
                        else
                        {
                            hi += v << (sh - 64);
                        }
                    }
                    const int col = (nt0 + q) * 8 + 2 * t + (r & 1);
                    if (col < C)
                    {
                        Y[(size_t)col * ct_stride + (size_t)pl * n + m0 + 8 * (r >> 1)] =
                            barrett_reduce_acc(u128{ lo, hi }, lc, t64.w, t64.wq);
                    }
                }
            }
        }
    } // namespace

    namespace
    {
        // byte-plane GEMM of the slices [pl_first, pl_first + pl_count) (pl = polynomial * limbs + limb):
        // one launch per plane-count class (7 planes for primes >= 2^48, 6 below) present among them
        void launch_imma(Context *c, const u64 *X, const uint2 *wp0, u64 *Y, int K, int cn, int Kp, int Cp, int np,
                         int limbs, int pl_first, int pl_count, cudaStream_t stream)
        {
            static const cudaError_t attr7 = cudaFuncSetAttribute(
                k_ctpt_gemm_imma<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, im_smem_bytes(7));
            static const cudaError_t attr6 = cudaFuncSetAttribute(
                k_ctpt_gemm_imma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, im_smem_bytes(6));
            (void)attr7;
            (void)attr6;
            unsigned long long mask6 = 0, mask7 = 0;
            for (int pl = pl_first; pl < pl_first + pl_count; pl++)
            {
                const int l = pl % limbs;
                ((c->q[l] >> 48) == 0 ? mask6 : mask7) |= 1ull << l;
            }
            const int cnp = (cn + IM_TILE_N - 1) / IM_TILE_N * IM_TILE_N;
            const long long ctas = (long long)pl_count * (c->n / IM_TILE_M) * (cnp / IM_TILE_N);
            if (mask7)
            {
                k_ctpt_gemm_imma<7><<<(unsigned)ctas, 256, im_smem_bytes(7), stream>>>(
                    X, wp0, Y, K, cn, Kp, Cp, np, cnp / IM_TILE_N, limbs, c->log_n, c->d_limb, c->d_two64, mask7, pl_first);
                c->launches += 1;
            }
            if (mask6)
            {
                k_ctpt_gemm_imma<6><<<(unsigned)ctas, 256, im_smem_bytes(6), stream>>>(
                    X, wp0, Y, K, cn, Kp, Cp, np, cnp / IM_TILE_N, limbs, c->log_n, c->d_limb, c->d_two64, mask6, pl_first);
                c->launches += 1;
            }
            MOAI_CUDA_CHECK(cudaGetLastError());
        }
    } // namespace

    // X: [K][2][limbs][n] device; W: host row-major K x C doubles; out: [C][2][limbs-1][n] device
    void ct_pt_matmul_scalar(Context *c, const u64 *X, const double *h_W, int K, int C, int limbs, double scale,
                             u64 *out, const u64 *post_pt)
    {
        // post_pt (optional, [limbs][n]): every pre-rescale output column is multiplied by this plaintext before the
        // rescale (the factorised masked matmul below)
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
        MOAI_CUDA_CHECK(cudaMemcpyAsync(dW.p, h_W, kc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        bool narrow = true; // every limb prime below 2^52 -> carry-free 26-bit split kernel
        for (int l = 0; l < limbs; l++)
        {
            narrow = narrow && (c->q[l] >> 52) == 0;
        }
        // 4 (default): int8 byte-plane GEMM, tcgen05.mma with TMEM accumulators (csrc/matmul_tc5.cu);
        // 3: the same GEMM on the legacy mma.sync path; 0: IMAD 128-bit; 1, 2: IMAD 26-bit split
        static const int variant_env = getenv("MOAI_GEMM_VARIANT") ? atoi(getenv("MOAI_GEMM_VARIANT")) : 4;
        const int variant = (variant_env == 4 && n < 128) ? 3 : variant_env;
        bool bytes7 = n >= (size_t)IM_TILE_M; // every limb prime below 2^56 -> at most 7 byte planes
        for (int l = 0; l < limbs; l++)
        {
            bytes7 = bytes7 && (c->q[l] >> 56) == 0;
        }
        if (variant == 4 && bytes7 && n >= 128)
        {
            // tcgen05.mma.kind::i8 with TMEM accumulators (csrc/matmul_tc5.cu)
            const int np = 7;
            Scratch dBp(tc5_packed_weight_bytes(K, C, limbs, np), c->stream);
            tc5_pack_weights(c, dW.as<double>(), dBp.as<unsigned char>(), K, C, limbs, np, scale);
            const int col_chunk = 768;
            Scratch Y((size_t)std::min(C, col_chunk) * 2 * limbs * n * sizeof(u64), c->stream);
            MOAI_REQUIRE(C <= col_chunk || C % 32 == 0, "column count must be a multiple of 32 beyond 768");
            for (int c0 = 0; c0 < C; c0 += col_chunk)
            {
                const int cn = std::min(col_chunk, C - c0);
                {
                    PhaseTimer pt(c, "ctpt_gemm");
                    tc5_gemm(c, X, dBp.as<unsigned char>(), Y.as<u64>(), K, C, c0, cn, np, limbs, 0, 2 * limbs, c->stream);
                }
                if (post_pt)
                {
                    ew_multiply_plain(c, Y.as<u64>(), post_pt, Y.as<u64>(), cn, 2, limbs, 0);
                }
                rescale(c, Y.as<u64>(), out + (size_t)c0 * 2 * (limbs - 1) * n, cn, 2, limbs);
            }
            return;
        }
        if (variant == 3 && bytes7)
        {
            const int Kp = (K + 31) / 32 * 32, Cp = (C + IM_TILE_N - 1) / IM_TILE_N * IM_TILE_N;
            const int np = 7;
            Scratch dWp((size_t)limbs * (Kp / 32) * (Cp / 8) * np * 32 * sizeof(uint2), c->stream);
            const long long pack_threads = (long long)limbs * (Kp / 32) * (Cp / 8) * 32;
            k_pack_weights<<<(unsigned)((pack_threads + 255) / 256), 256, 0, c->stream>>>(
                dW.as<double>(), dWp.as<uint2>(), K, C, Kp, Cp, limbs, np, scale, c->d_limb);
            c->launches += 1;
            const int col_chunk = 768;
            Scratch Y((size_t)std::min(C, col_chunk) * 2 * limbs * n * sizeof(u64), c->stream);
            MOAI_REQUIRE(C <= col_chunk || C % IM_TILE_N == 0, "column count must be a multiple of 32 beyond 768");
            for (int c0 = 0; c0 < C; c0 += col_chunk)
            {
                const int cn = std::min(col_chunk, C - c0);
                {
                    PhaseTimer pt(c, "ctpt_gemm");
                    // the packed weights of column chunk c0 start c0 / 8 n-tiles into every k-step row
                    launch_imma(c, X, dWp.as<uint2>() + (size_t)(c0 / 8) * np * 32, Y.as<u64>(), K, cn, Kp, Cp, np, limbs,
                                0, 2 * limbs, c->stream);
                }
                MOAI_CUDA_CHECK(cudaGetLastError());
                if (post_pt)
                {
                    ew_multiply_plain(c, Y.as<u64>(), post_pt, Y.as<u64>(), cn, 2, limbs, 0);
                }
                rescale(c, Y.as<u64>(), out + (size_t)c0 * 2 * (limbs - 1) * n, cn, 2, limbs);
            }
            return;
        }
        const bool split26 = narrow && variant != 0 && variant != 3;
        k_encode_weights<<<(unsigned)((kc + 255) / 256), 256, 0, c->stream>>>(dW.as<double>(), dWc.as<u64>(), kc, limbs,
                                                                            scale, c->d_limb, split26 ? 1 : 0);
        c->launches += 1;
        // output columns in chunks: bounds the pre-rescale buffer Y (2 * limbs * N words per column)
        const int col_chunk = 768;
        const int tiles_t = (int)((n / 2) / MM_THREADS);
        Scratch Y((size_t)std::min(C, col_chunk) * 2 * limbs * n * sizeof(u64), c->stream);
        for (int c0 = 0; c0 < C; c0 += col_chunk)
        {
            const int cn = std::min(col_chunk, C - c0);
            const int tiles_i = (cn + MM_TN - 1) / MM_TN;
            const long long ctas = (long long)2 * limbs * tiles_t * tiles_i;
            {
                PhaseTimer pt(c, "ctpt_gemm");
                if (split26 && variant == 2)
                {
                    k_ctpt_gemm26<1><<<(unsigned)(2 * ctas), MM_THREADS, 0, c->stream>>>(
                        X, dWc.as<u64>() + c0, Y.as<u64>(), K, cn, C, limbs, c->log_n, c->d_limb, c->d_two64);
                }
                else if (split26)
                {
                    k_ctpt_gemm26<2><<<(unsigned)ctas, MM_THREADS, 0, c->stream>>>(
                        X, dWc.as<u64>() + c0, Y.as<u64>(), K, cn, C, limbs, c->log_n, c->d_limb, c->d_two64);
                }
                else
                {
                    k_ctpt_gemm<<<(unsigned)ctas, MM_THREADS, 0, c->stream>>>(
                        reinterpret_cast<const ulonglong2 *>(X), dWc.as<u64>() + c0, Y.as<ulonglong2>(), K, cn, C, limbs,
                        c->log_n - 1, c->d_limb, c->d_two64);
                }
            }
            c->launches += 1;
            MOAI_CUDA_CHECK(cudaGetLastError());
            if (post_pt)
            {
                ew_multiply_plain(c, Y.as<u64>(), post_pt, Y.as<u64>(), cn, 2, limbs, 0);
            }
            rescale(c, Y.as<u64>(), out + (size_t)c0 * 2 * (limbs - 1) * n, cn, 2, limbs);
        }
    }


    // Same module with HOST buffers (what the reference's callers hold: vector<Ciphertext> in host memory):
    // h_X [K][2][limbs][n] -> h_out [C][2][limbs-1][n].  The upload is cut into the 2 * limbs (polynomial, limb)
    // slices of the batch — highest limb first, because the rescale needs it for every output limb — and the
    // byte-plane GEMM of a slice starts as soon as the slice has landed, so the tensor cores run under the
    // PCIe transfer; the rescaled columns are downloaded in chunks while the next chunk is rescaled.
    void ct_pt_matmul_scalar_host(Context *c, const u64 *h_X, const double *h_W, int K, int C, int limbs, double scale,
                                  u64 *h_out)
    {
        MOAI_REQUIRE(K >= 1 && C >= 1 && C <= 768, "bad dimensions of X or W");
        MOAI_REQUIRE(limbs >= 2 && limbs <= c->kl - 1, "end of modulus switching chain reached");
        MOAI_REQUIRE(K <= 4096, "K too large for the lazy accumulators");
        const size_t n = c->n;
        MOAI_REQUIRE(n >= (size_t)IM_TILE_M, "ring degree too small for the tensor-core kernel");
        for (int l = 0; l < limbs; l++)
        {
            MOAI_REQUIRE((c->q[l] >> 56) == 0, "primes of 56 bits or more are not supported by the byte-plane kernel");
        }
        const long long kc = (long long)K * C;
        double wmax = 0;
        for (long long i = 0; i < kc; i++)
        {
            wmax = std::fmax(wmax, std::fabs(h_W[i]));
        }
        MOAI_REQUIRE(wmax * scale < 9.0e18, "encoded value is too large");
        // side streams and events: destroyed on EVERY exit path, after the streams have drained (a check that throws
        // midway must neither leak them nor hand the device buffers back to the arena while a copy still uses them)
        struct SideResources
        {
            cudaStream_t s_in = nullptr, s_out = nullptr;
            std::vector<cudaEvent_t> events;
            cudaEvent_t event()
            {
                cudaEvent_t e = nullptr;
                MOAI_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                events.push_back(e);
                return e;
            }
            void drain() const
            {
                if (s_in)
                {
                    cudaStreamSynchronize(s_in);
                }
                if (s_out)
                {
                    cudaStreamSynchronize(s_out);
                }
            }
            ~SideResources()
            {
                drain();
                for (cudaEvent_t e : events)
                {
                    cudaEventDestroy(e);
                }
                if (s_in)
                {
                    cudaStreamDestroy(s_in);
                }
                if (s_out)
                {
                    cudaStreamDestroy(s_out);
                }
            }
        } side;
        MOAI_CUDA_CHECK(cudaStreamCreateWithFlags(&side.s_in, cudaStreamNonBlocking));
        MOAI_CUDA_CHECK(cudaStreamCreateWithFlags(&side.s_out, cudaStreamNonBlocking));
        cudaStream_t s_in = side.s_in, s_out = side.s_out;
        std::vector<cudaEvent_t> landed(2 * limbs);
        for (auto &e : landed)
        {
            e = side.event();
        }
        const int Kp = (K + 31) / 32 * 32, Cp = (C + IM_TILE_N - 1) / IM_TILE_N * IM_TILE_N, np = 7;
        const size_t ct_words = (size_t)2 * limbs * n;
        static const bool use_tc5 = !(getenv("MOAI_GEMM_VARIANT") && atoi(getenv("MOAI_GEMM_VARIANT")) != 4);
        const bool tc5 = use_tc5 && n >= 128;
        {
            Scratch dW(kc * sizeof(double), c->stream);
            Scratch dWp(tc5 ? tc5_packed_weight_bytes(K, C, limbs, np)
                            : (size_t)limbs * (Kp / 32) * (Cp / 8) * np * 32 * sizeof(uint2),
                        c->stream);
            Scratch dX((size_t)K * ct_words * sizeof(u64), c->stream);
            Scratch Y((size_t)C * ct_words * sizeof(u64), c->stream);
            Scratch dOut((size_t)C * 2 * (limbs - 1) * n * sizeof(u64), c->stream);
            struct DrainFirst // declared after the buffers: runs before they are released, also while unwinding
            {
                const SideResources &r;
                cudaStream_t main;
                ~DrainFirst()
                {
                    r.drain();
                    cudaStreamSynchronize(main);
                }
            } drain_first{ side, c->stream };
            MOAI_CUDA_CHECK(cudaMemcpyAsync(dW.p, h_W, kc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            if (tc5)
            {
                tc5_pack_weights(c, dW.as<double>(), dWp.as<unsigned char>(), K, C, limbs, np, scale);
            }
            else
            {
                const long long pack_threads = (long long)limbs * (Kp / 32) * (Cp / 8) * 32;
                k_pack_weights<<<(unsigned)((pack_threads + 255) / 256), 256, 0, c->stream>>>(
                    dW.as<double>(), dWp.as<uint2>(), K, C, Kp, Cp, limbs, np, scale, c->d_limb);
                c->launches += 1;
            }
            // the buffers come from the stream-ordered arena of c->stream: order the side streams after it
            cudaEvent_t ready = side.event();
            MOAI_CUDA_CHECK(cudaEventRecord(ready, c->stream));
            MOAI_CUDA_CHECK(cudaStreamWaitEvent(s_in, ready, 0));
            MOAI_CUDA_CHECK(cudaStreamWaitEvent(s_out, ready, 0));
            int order = 0;
            for (int l = limbs - 1; l >= 0; l--)
            {
                for (int p = 0; p < 2; p++, order++)
                {
                    const size_t off = ((size_t)p * limbs + l) * n;
                    MOAI_CUDA_CHECK(cudaMemcpy2DAsync(dX.as<u64>() + off, ct_words * sizeof(u64), h_X + off,
                                                      ct_words * sizeof(u64), n * sizeof(u64), (size_t)K,
                                                      cudaMemcpyHostToDevice, s_in));
                    MOAI_CUDA_CHECK(cudaEventRecord(landed[order], s_in));
                    MOAI_CUDA_CHECK(cudaStreamWaitEvent(c->stream, landed[order], 0));
                    PhaseTimer pt(c, "ctpt_gemm");
                    if (tc5)
                    {
                        tc5_gemm(c, dX.as<u64>(), dWp.as<unsigned char>(), Y.as<u64>(), K, C, 0, C, np, limbs,
                                 p * limbs + l, 1, c->stream);
                    }
                    else
                    {
                        launch_imma(c, dX.as<u64>(), dWp.as<uint2>(), Y.as<u64>(), K, C, Kp, Cp, np, limbs, p * limbs + l,
                                    1, c->stream);
                    }
                }
            }
            const int out_chunk = 96;
            const size_t out_ct = (size_t)2 * (limbs - 1) * n;
            std::vector<cudaEvent_t> done((C + out_chunk - 1) / out_chunk);
            for (int c0 = 0, k = 0; c0 < C; c0 += out_chunk, k++)
            {
                const int cn = std::min(out_chunk, C - c0);
                rescale(c, Y.as<u64>() + (size_t)c0 * ct_words, dOut.as<u64>() + (size_t)c0 * out_ct, cn, 2, limbs);
                done[k] = side.event();
                MOAI_CUDA_CHECK(cudaEventRecord(done[k], c->stream));
                MOAI_CUDA_CHECK(cudaStreamWaitEvent(s_out, done[k], 0));
                MOAI_CUDA_CHECK(cudaMemcpyAsync(h_out + (size_t)c0 * out_ct, dOut.as<u64>() + (size_t)c0 * out_ct,
                                                (size_t)cn * out_ct * sizeof(u64), cudaMemcpyDeviceToHost, s_out));
            }
            MOAI_CUDA_CHECK(cudaStreamSynchronize(s_out));
            MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        } // device buffers return to the arena after every stream has drained (DrainFirst); `side` frees the handles
    }

    namespace
    {
        // acc[i][p][l][t] (+)= X[j][p][l][t] * pt[(jj * CC + ii)][l][t]  for the j-chunk [j0, j0+KC), columns [i0, i0+CC)
        __global__ void k_masked_mac(const ulonglong2 *__restrict__ X, const ulonglong2 *__restrict__ pt,
                                     ulonglong2 *__restrict__ acc, int KCn, int CCn, int limbs, int log_n2, int first,
                                     const LimbConst *__restrict__ lcs, const Twiddle *__restrict__ two64)
        {
            // grid: x over [2*limbs][n/2 / 256], y = ii
            const int tiles_t = (1 << log_n2) / MM_THREADS;
            const int tile_t = blockIdx.x % tiles_t;
            const int pl = blockIdx.x / tiles_t;
            const int l = pl % limbs;
            const int ii = blockIdx.y;
            const LimbConst lc = lcs[l];
            const Twiddle t64 = two64[l];
            const long long within = (long long)tile_t * MM_THREADS + threadIdx.x;
            const long long ct_stride2 = (long long)2 * limbs << log_n2;
            const long long pt_stride2 = (long long)limbs << log_n2;
            u128 ax{ 0, 0 }, ay{ 0, 0 };
            for (int jj = 0; jj < KCn; jj++)
            {
                const ulonglong2 x = X[(long long)jj * ct_stride2 + ((long long)pl << log_n2) + within];
                const ulonglong2 m = pt[((long long)jj * CCn + ii) * pt_stride2 + ((long long)l << log_n2) + within];
                mac_wide(ax, x.x, m.x);
                mac_wide(ay, x.y, m.y);
            }
            ulonglong2 r;
            r.x = barrett_reduce_acc(ax, lc, t64.w, t64.wq);
            r.y = barrett_reduce_acc(ay, lc, t64.w, t64.wq);
            ulonglong2 *dst = acc + (long long)ii * ct_stride2 + ((long long)pl << log_n2) + within;
            if (!first)
            {
                const ulonglong2 o = *dst;
                r.x = addmod(r.x, o.x, lc.q);
                r.y = addmod(r.y, o.y, lc.q);
            }
            *dst = r;
        }
    } // namespace

    // ct_pt_matrix_mul_wo_pre_w_mask (M/source/matrix_mul/Ct_pt_matrix_mul.hpp:103-170), exact for any
    // 0/1 mask: every plaintext encode_vector(W[j][i] * mask) is produced on the device with the
    // reference's FFT (bit-identical), then multiplied and accumulated; K*C FFTs + K*C*limbs NTTs.
    void ct_pt_matmul_masked(Context *c, const u64 *X, const double *h_W, const int *h_mask, int K, int C, int limbs,
                             double scale, u64 *out)
    {
        MOAI_REQUIRE(K >= 1 && C >= 1, "bad dimensions of X or W");
        MOAI_REQUIRE(limbs >= 2 && limbs <= c->kl - 1, "end of modulus switching chain reached");
        const size_t n = c->n;
        // chunking: KC x CC plaintexts at a time (bounded workspace ~ 1.5 GiB incl. the FFT buffers)
        const size_t per_pt = n * sizeof(double2) + (size_t)limbs * n * sizeof(u64);
        long long budget = (long long)(((size_t)3 << 29) / per_pt);
        budget = budget < 1 ? 1 : budget;
        int CC = (int)(budget < C ? budget : C);
        int KC = (int)(budget / CC);
        KC = KC < 1 ? 1 : (KC > K ? K : KC);
        Scratch dmask((n / 2) * sizeof(int), c->stream);
        Scratch dw((size_t)KC * CC * sizeof(double), c->stream);
        Scratch pts((size_t)KC * CC * limbs * n * sizeof(u64), c->stream);
        Scratch Y((size_t)C * 2 * limbs * n * sizeof(u64), c->stream);
        MOAI_CUDA_CHECK(cudaMemcpyAsync(dmask.p, h_mask, (n / 2) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        std::vector<double> hw((size_t)KC * CC);
        const int tiles_t = (int)((n / 2) / MM_THREADS);
        for (int i0 = 0; i0 < C; i0 += CC)
        {
            const int cn = std::min(CC, C - i0);
            for (int j0 = 0; j0 < K; j0 += KC)
            {
                const int kn = std::min(KC, K - j0);
                for (int jj = 0; jj < kn; jj++)
                {
                    for (int ii = 0; ii < cn; ii++)
                    {
                        hw[(size_t)jj * cn + ii] = h_W[(size_t)(j0 + jj) * C + i0 + ii];
                    }
                }
                MOAI_CUDA_CHECK(cudaMemcpyAsync(dw.p, hw.data(), (size_t)kn * cn * sizeof(double),
                                                cudaMemcpyHostToDevice, c->stream));
                MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream)); // hw is reused by the next chunk
                encode_masked_weights(c, dw.as<double>(), dmask.as<int>(), (long long)kn * cn, scale, limbs,
                                      pts.as<u64>());
                dim3 grid((unsigned)(2 * limbs * tiles_t), (unsigned)cn);
                k_masked_mac<<<grid, MM_THREADS, 0, c->stream>>>(
                    reinterpret_cast<const ulonglong2 *>(X + (size_t)j0 * 2 * limbs * n), pts.as<ulonglong2>(),
                    Y.as<ulonglong2>() + (size_t)i0 * limbs * n, kn, cn, limbs, c->log_n - 1, j0 == 0 ? 1 : 0,
                    c->d_limb, c->d_two64);
                c->launches += 1;
                MOAI_CUDA_CHECK(cudaGetLastError());
            }
        }
        rescale(c, Y.as<u64>(), out, C, 2, limbs);
    }
    // Fast-mode ct_pt_matrix_mul_wo_pre_w_mask: the reference encodes one plaintext per weight,
    // encode(w_ji * mask) at `scale` (Ct_pt_matrix_mul.hpp:127-134) — K * C encodings and K * C * limbs transforms, which
    // is what ct_pt_matmul_masked reproduces bit for bit (2.8 s for 768 x 768 at N = 65536).  Encoding is linear up to
    // its rounding, so   sum_j X_j (.) encode(w_ji mask)  ~  [sum_j round(w_ji s_w) X_j] (.) encode(mask at 2^26),
    // s_w = scale / 2^26:  ONE scalar GEMM on the tensor cores with integer weights at scale s_w, ONE plaintext, the same
    // output scale (s_w 2^26 = scale exactly) and the same single rescale.  Error against the exact path: weights
    // quantised to 1 / s_w (2^-20 at scale 2^46) and the mask's coefficients rounded at 2^26 (slot error ~ sqrt(N) 0.29 /
    // 2^26 of the output value).  Measured on C1 (768 x 768, N = 65536, max |XW| = 25): 1.15e-4 max-abs, against 1.7e-5
    // for the exact path and a stated tolerance of 2.5e-3; 15 ms against 2.8 s (tests/test_gpu_fullsize.py).
    // Needs scale >= 2^44.
    // bits of the mask plaintext's scale: its slot error grows like sqrt(N) / 2^t, the weights' like sqrt(K) |x| / 2^(46 - t).
    // Measured on C1 at N = 65536 (max |XW| = 25): t = 26 -> 1.15e-4 max-abs, t = 28 -> 4.2e-4 (the weights dominate);
    // encoder layer at N = 4096, fast vs SEAL-exact keys: t = 26 -> 1.6e-5, t = 28 -> 5.6e-5.
    static double mask_pt_scale(const Context *)
    {
        return 67108864.0; // 2^26
    }
    bool ct_pt_matmul_masked_fast_ok(double scale)
    {
        return scale >= 17592186044416.0; // 2^44: at least 16 bits for the weights
    }
    void ct_pt_matmul_masked_fast(Context *c, const u64 *X, const double *h_W, const int *h_mask, int K, int C, int limbs,
                                  double scale, u64 *out)
    {
        MOAI_REQUIRE(ct_pt_matmul_masked_fast_ok(scale), "scale too small for the factorised masked matmul");
        const size_t n = c->n;
        Scratch dmask((n / 2) * sizeof(int), c->stream);
        Scratch done(sizeof(double), c->stream);
        Scratch V((size_t)limbs * n * sizeof(u64), c->stream);
        const double one = 1.0;
        MOAI_CUDA_CHECK(cudaMemcpyAsync(dmask.p, h_mask, (n / 2) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        MOAI_CUDA_CHECK(cudaMemcpyAsync(done.p, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        MOAI_CUDA_CHECK(cudaStreamSynchronize(c->stream)); // `one` is a stack variable
        const double vs = mask_pt_scale(c);
        encode_masked_weights(c, done.as<double>(), dmask.as<int>(), 1, vs, limbs, V.as<u64>());
        ct_pt_matmul_scalar(c, X, h_W, K, C, limbs, scale / vs, out, V.as<u64>());
    }
} // namespace moai
