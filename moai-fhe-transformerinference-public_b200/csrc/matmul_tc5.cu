// ct-pt matmul on the 5th-generation tensor cores: the int8 byte-plane GEMM of csrc/matmul.cu issued as
// tcgen05.mma.kind::i8 with the accumulators in tensor memory.
//
// Same arithmetic as k_ctpt_gemm_imma (Ct_pt_matrix_mul.hpp:20-42 as a modular GEMM per limb):
//      D_s[m][i] = sum_j sum_{a+b=s} x_a[j][m] * w_b[j][i],   Y[m][i] = (sum_s 2^(8s) D_s[m][i]) mod q
// B200 mapping: one CTA owns 128 coefficients (the 128 TMEM lanes) x 32 output columns; the 2 NP - 1
// diagonal accumulators D_s are 32 TMEM columns each (13 x 32 = 416 of the SM's 512 columns for NP = 7).
// Per k-step of 32 ciphertexts the 128 producer threads load their row's residues (coalesced along the
// coefficient index), peel the byte planes with PRMT and write them straight into tensor memory
// (tcgen05.st; A operand from TMEM); the weight planes arrive pre-packed in the canonical K-major,
// no-swizzle UMMA layout (8-row x 16-byte core matrices) by cp.async into two shared-memory stages.  A fifth
// warp issues the NP^2 MMAs (M = 128, N = 32, K = 32) of the step and commits them to an mbarrier, so the
// next step's loads and byte transposes run under the tensor cores.  The epilogue reads the diagonals with
// tcgen05.ld, recombines the 128-bit sum and reduces it modulo q — the exact canonical residue,
// bit-identical to the other GEMM kernels.
#include "ntt.cuh"
#include "ops.cuh"
#include <algorithm>

namespace moai
{
    namespace
    {
        constexpr int T5_M = 128, T5_N = 32, T5_K = 32;
        constexpr int T5_B_PLANE = T5_N * T5_K; // 1024 bytes
        constexpr bool T5_PER_PLANE = false; // hand-off granularity producers <-> MMA warp: per byte plane or per k-step
        constexpr int T5_LBO = 128, T5_SBO = 256; // K-major, no swizzle: k-chunk stride, 8-row-group stride

        // Bp[l][ks][nt][b][1024]: byte plane b of Wc[l][32 ks + k][32 nt + n] at
        // (n / 8) * SBO + (k / 16) * LBO + (n % 8) * 16 + k % 16
        __global__ void k_pack_weights_tc5(const double *__restrict__ W, unsigned char *__restrict__ Bp, int K, int C, int Kp,
                                           int Cp, int limbs, int np, double scale, const LimbConst *__restrict__ lcs)
        {
            const long long total = (long long)limbs * Kp * Cp;
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total)
            {
                return;
            }
            const int col = (int)(i % Cp);
            const int j = (int)((i / Cp) % Kp);
            const int l = (int)(i / ((long long)Cp * Kp));
            const LimbConst lc = lcs[l];
            u64 r = 0;
            if (j < K && col < C)
            {
                const double v = round(W[(long long)j * C + col] * scale);
                const u64 mag = (u64)fabs(v);
                r = reduce64(mag, lc);
                r = signbit(v) ? negmod(r, lc.q) : r;
            }
            const int ks = j / T5_K, k = j % T5_K, nt = col / T5_N, n = col % T5_N;
            unsigned char *dst = Bp + ((((size_t)l * (Kp / T5_K) + ks) * (Cp / T5_N) + nt) * np) * T5_B_PLANE +
                                 (n / 8) * T5_SBO + (k / 16) * T5_LBO + (n % 8) * 16 + k % 16;
            for (int b = 0; b < np; b++)
            {
                dst[(size_t)b * T5_B_PLANE] = (unsigned char)(r >> (8 * b));
            }
        }

        __device__ __forceinline__ u32 t5_prmt(u32 a, u32 b, u32 sel)
        {
            u32 r;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
            return r;
        }
        __device__ __forceinline__ void t5_transpose4x4(const u32 (&r)[4], u32 (&p)[4])
        {
            const u32 t0 = t5_prmt(r[0], r[1], 0x5140), t1 = t5_prmt(r[0], r[1], 0x7362);
            const u32 t2 = t5_prmt(r[2], r[3], 0x5140), t3 = t5_prmt(r[2], r[3], 0x7362);
            p[0] = t5_prmt(t0, t2, 0x5410);
            p[1] = t5_prmt(t0, t2, 0x7632);
            p[2] = t5_prmt(t1, t3, 0x5410);
            p[3] = t5_prmt(t1, t3, 0x7632);
        }

        __device__ __forceinline__ u64 t5_smem_desc(unsigned smem_addr)
        {
            // start address, LBO, SBO in 16-byte units; version 1 (Blackwell); no swizzle
            return (u64)((smem_addr >> 4) & 0x3FFF) | ((u64)(T5_LBO >> 4) << 16) | ((u64)(T5_SBO >> 4) << 32) | (1ull << 46);
        }

        __device__ __forceinline__ void t5_mbar_wait(unsigned bar, unsigned parity)
        {
            asm volatile("{\n\t"
                         ".reg .pred P1;\n\t"
                         "WAIT_%=:\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                         "@P1 bra DONE_%=;\n\t"
                         "bra WAIT_%=;\n\t"
                         "DONE_%=:\n\t"
                         "}\n" ::"r"(bar),
                         "r"(parity)
                         : "memory");
        }

        // A operand from tensor memory ("TS" form): D[tmem] += A[tmem] * B[smem]
        __device__ __forceinline__ void t5_mma_i8_ts(unsigned tmem_d, unsigned tmem_a, u64 desc_b, u32 idesc, u32 accumulate)
        {
            asm volatile("{\n\t"
                         ".reg .pred p;\n\t"
                         "setp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
                         "}\n" ::"r"(tmem_d),
                         "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
                         : "memory");
        }

        // Warp-specialised: warps 0-3 (one thread per coefficient = TMEM lane) prepare the operands, warp 4
        // issues the MMAs.  The byte planes of A go from registers straight into tensor memory
        // (tcgen05.st, 8 columns per plane): the tensor cores then fetch only the 1 KiB weight plane of each
        // MMA from shared memory instead of 5 KiB, which is what bounded the all-shared-memory version
        // (NP^2 MMAs per k-step re-reading the same NP A planes).  TMEM columns: (2 NP - 1) x 32 accumulators,
        // then NP x 8 for A.
        template <int NP>
        __global__ void __launch_bounds__(160, 1)
            k_ctpt_gemm_tc5(const u64 *__restrict__ X, const unsigned char *__restrict__ Bp, u64 *__restrict__ Y, int K, int C,
                            int Kp, int Cp, int wnp, int tiles_n, int limbs, int log_n, const LimbConst *__restrict__ lcs,
                            const Twiddle *__restrict__ two64, unsigned long long limb_mask, int pl_first)
        {
            constexpr int ND = 2 * NP - 1;
            constexpr int A_COL0 = ND * T5_N; // first TMEM column of the A planes
            constexpr int B_STAGE = NP * T5_B_PLANE;
            extern __shared__ __align__(1024) unsigned char t5_smem[]; // two stages of weight planes
            // per byte plane a: full[a] = plane a of the current step is in TMEM (128 arrivals),
            // done[a] = the NP MMAs that read it have completed (tcgen05.commit): the producers refill plane a
            // while the tensor cores work on the later planes of the step
            __shared__ __align__(8) unsigned long long bars[2 * NP];
            __shared__ u32 tmem_base_slot;
            const int tid = threadIdx.x, warp = tid >> 5;
            const int tiles_m = (1 << log_n) / T5_M;
            const int tile_n = blockIdx.x % tiles_n;
            const int rest = blockIdx.x / tiles_n;
            const int tile_m = rest % tiles_m;
            const int pl = pl_first + rest / tiles_m;
            const int l = pl % limbs;
            if (!((limb_mask >> l) & 1))
            {
                return;
            }
            const size_t n = (size_t)1 << log_n;
            const size_t ct_stride = (size_t)2 * limbs * n;
            const int nks = Kp / T5_K;
            const unsigned smem0 = (unsigned)__cvta_generic_to_shared(t5_smem);
            const unsigned bar_full = (unsigned)__cvta_generic_to_shared(&bars[0]);
            const unsigned bar_done = bar_full + 8 * NP;
            if (tid == 0)
            {
                for (int a = 0; a < NP; a++)
                {
                    asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(bar_full + 8 * a));
                    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_done + 8 * a));
                }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            if (warp == 0)
            {
                const unsigned slot = (unsigned)__cvta_generic_to_shared(&tmem_base_slot);
                asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
                asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const u32 tmem_base = tmem_base_slot;
            // instruction descriptor: D = s32, A = B = u8, both K-major, M = 128, N = 32 NP (all weight planes)
            const u32 idesc_wide = (2u << 4) | ((u32)((T5_N * NP) >> 3) << 17) | ((u32)(T5_M >> 4) << 24);

            if (warp == 4)
            {
                // ---- MMA issuer
                if ((tid & 31) == 0)
                {
                    for (int ks = 0; ks < nks; ks++)
                    {
                        const u64 bdesc0 = t5_smem_desc(smem0 + (ks & 1) * B_STAGE);
#pragma unroll
                        for (int a = 0; a < NP; a++)
                        {
                            if (T5_PER_PLANE || a == 0)
                            {
                                t5_mbar_wait(bar_full + 8 * (T5_PER_PLANE ? a : 0), (unsigned)(ks & 1));
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            }
                            // ONE MMA per A plane: the NP weight planes are stacked along N (they are contiguous
                            // in the stage and in the same core-matrix layout), so D columns [32 a, 32 (a + NP)) =
                            // the diagonals a .. a + NP - 1 receive x_a * (w_0 | w_1 | ... | w_{NP-1}) at once
                            t5_mma_i8_ts(tmem_base + (u32)a * T5_N, tmem_base + (u32)(A_COL0 + 8 * a), bdesc0, idesc_wide, 1u);
                            if (T5_PER_PLANE || a == NP - 1)
                            {
                                asm volatile(
                                    "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                        bar_done + 8 * (T5_PER_PLANE ? a : 0))
                                    : "memory");
                            }
                        }
                    }
                }
                __syncwarp();
            }
            else
            {
                // ---- operand producers: thread = coefficient = TMEM lane
                const size_t m = (size_t)tile_m * T5_M + tid;
                const u64 *xp = X + (size_t)pl * n + m;
                const unsigned char *bg = Bp + (((size_t)l * (Kp / T5_K)) * (Cp / T5_N) + tile_n) * wnp * T5_B_PLANE;
                const size_t bg_ks = (size_t)(Cp / T5_N) * wnp * T5_B_PLANE;
                const u32 lane_base = tmem_base + ((u32)(warp * 32) << 16);
                auto issue_b = [&](int ks) {
                    const unsigned char *src = bg + (size_t)ks * bg_ks;
                    unsigned char *dst = t5_smem + (ks & 1) * B_STAGE;
                    for (int off = tid * 16; off < B_STAGE; off += 128 * 16)
                    {
                        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + off);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + off) : "memory");
                    }
                };
                issue_b(0);
                asm volatile("cp.async.commit_group;" ::: "memory");
                // every MMA accumulates (one instruction spans NP diagonals): clear the accumulators first
#pragma unroll
                for (int s0 = 0; s0 < ND; s0++)
                {
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                                 "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
                                 "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_base +
                                                                                                     (u32)(s0 * T5_N)),
                                 "r"(0u)
                                 : "memory");
                }
                // the residues of the NEXT step are requested before the current step is processed (two
                // register sets), so a full step of work hides the global-load latency
                auto load_step = [&](int ks, u64 (&v)[32]) {
#pragma unroll
                    for (int e = 0; e < 32; e++)
                    {
                        const int j = ks * T5_K + e;
                        v[e] = j < K ? __ldg(xp + (size_t)j * ct_stride) : 0;
                    }
                };
                auto process_step = [&](int ks, const u64 (&v)[32]) {
                    // residues of this coefficient in the 32 ciphertexts of the step -> byte planes (registers)
                    u32 plane[8][NP]; // [group of 4 ciphertexts][plane]
#pragma unroll
                    for (int g4 = 0; g4 < 8; g4++)
                    {
                        u32 lo[4], hi[4], pl4[4], ph4[4];
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            lo[e] = (u32)v[4 * g4 + e];
                            hi[e] = (u32)(v[4 * g4 + e] >> 32);
                        }
                        t5_transpose4x4(lo, pl4);
                        t5_transpose4x4(hi, ph4);
#pragma unroll
                        for (int a = 0; a < NP; a++)
                        {
                            plane[g4][a] = a < 4 ? pl4[a] : ph4[a - 4];
                        }
                    }
                    // weight planes of this step (requested one step ago) must be visible before its first MMA
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
                    for (int a = 0; a < NP; a++)
                    {
                        if (ks >= 1 && (T5_PER_PLANE || a == 0))
                        {
                            // previous step's MMAs on plane a (or on all planes) have completed
                            t5_mbar_wait(bar_done + 8 * (T5_PER_PLANE ? a : 0), (unsigned)((ks - 1) & 1));
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                                         lane_base + (u32)(A_COL0 + 8 * a)),
                                     "r"(plane[0][a]), "r"(plane[1][a]), "r"(plane[2][a]), "r"(plane[3][a]),
                                     "r"(plane[4][a]), "r"(plane[5][a]), "r"(plane[6][a]), "r"(plane[7][a])
                                     : "memory");
                        if (T5_PER_PLANE || a == NP - 1)
                        {
                            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_full + 8 * (T5_PER_PLANE ? a : 0))
                                         : "memory");
                        }
                    }
                    // every MMA of the previous step is done: its weight stage can take the next step's planes
                    if (ks + 1 < nks)
                    {
                        issue_b(ks + 1);
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                u64 va[32], vb[32];
                load_step(0, va);
                for (int ks = 0; ks < nks; ks += 2)
                {
                    if (ks + 1 < nks)
                    {
                        load_step(ks + 1, vb);
                    }
                    process_step(ks, va);
                    if (ks + 1 < nks)
                    {
                        if (ks + 2 < nks)
                        {
                            load_step(ks + 2, va);
                        }
                        process_step(ks + 1, vb);
                    }
                }
                // every MMA has completed when the last plane's last commit has arrived
                t5_mbar_wait(bar_done + 8 * (T5_PER_PLANE ? NP - 1 : 0), (unsigned)((nks - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

                // epilogue: lane = coefficient; 16 columns at a time
                const LimbConst lc = lcs[l];
                const Twiddle t64 = two64[l];
#pragma unroll 1
                for (int half = 0; half < 2; half++)
                {
                    u64 lo[16], hi[16];
#pragma unroll
                    for (int c = 0; c < 16; c++)
                    {
                        lo[c] = 0;
                        hi[c] = 0;
                    }
#pragma unroll
                    for (int s = 0; s < ND; s++)
                    {
                        u32 d[16];
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]),
                                       "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]),
                                       "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                                     : "r"(lane_base + (u32)(s * T5_N + half * 16)));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        const int sh = 8 * s;
#pragma unroll
                        for (int c = 0; c < 16; c++)
                        {
                            const u64 v = (u64)d[c];
                            if (sh == 0)
                            {
                                lo[c] = v;
                            }
                            else if (sh < 64)
                            {
                                const u64 add = v << sh;
                                lo[c] += add;
                                hi[c] += (lo[c] < add) + (v >> (64 - sh));
                            }
                            else
                            {
                                hi[c] += v << (sh - 64);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 16; c++)
                    {
                        const int col = tile_n * T5_N + half * 16 + c;
                        if (col < C)
                        {
                            Y[(size_t)col * ct_stride + (size_t)pl * n + m] =
                                barrett_reduce_acc(u128{ lo[c], hi[c] }, lc, t64.w, t64.wq);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (warp == 0)
            {
                asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
            }
        }
    } // namespace

    // packed weights for the tcgen05 kernel: limbs x Kp x Cp x np bytes
    size_t tc5_packed_weight_bytes(int K, int C, int limbs, int np)
    {
        const int Kp = (K + T5_K - 1) / T5_K * T5_K, Cp = (C + T5_N - 1) / T5_N * T5_N;
        return (size_t)limbs * Kp * Cp * np;
    }

    void tc5_pack_weights(Context *c, const double *dW, unsigned char *Bp, int K, int C, int limbs, int np, double scale)
    {
        const int Kp = (K + T5_K - 1) / T5_K * T5_K, Cp = (C + T5_N - 1) / T5_N * T5_N;
        const long long total = (long long)limbs * Kp * Cp;
        k_pack_weights_tc5<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(dW, Bp, K, C, Kp, Cp, limbs, np, scale,
                                                                                  c->d_limb);
        c->launches += 1;
        MOAI_CUDA_CHECK(cudaGetLastError());
    }

    // GEMM of the slices [pl_first, pl_first + pl_count) for `cn` columns starting at column c0 of the packed layout
    void tc5_gemm(Context *c, const u64 *X, const unsigned char *Bp, u64 *Y, int K, int C_total, int c0, int cn, int np,
                  int limbs, int pl_first, int pl_count, cudaStream_t stream)
    {
        MOAI_REQUIRE(c0 % T5_N == 0, "column chunks must start on a 32-column tile");
        MOAI_REQUIRE(c->n >= (size_t)T5_M, "ring degree too small for the tcgen05 kernel");
        static const cudaError_t attr7 = cudaFuncSetAttribute(k_ctpt_gemm_tc5<7>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                              2 * 7 * T5_B_PLANE);
        static const cudaError_t attr6 = cudaFuncSetAttribute(k_ctpt_gemm_tc5<6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                              2 * 6 * T5_B_PLANE);
        (void)attr7;
        (void)attr6;
        const int Kp = (K + T5_K - 1) / T5_K * T5_K, Cp = (C_total + T5_N - 1) / T5_N * T5_N;
        unsigned long long mask6 = 0, mask7 = 0;
        for (int pl = pl_first; pl < pl_first + pl_count; pl++)
        {
            const int l = pl % limbs;
            ((c->q[l] >> 48) == 0 ? mask6 : mask7) |= 1ull << l;
        }
        const int tiles_n = (cn + T5_N - 1) / T5_N;
        const long long ctas = (long long)pl_count * (c->n / T5_M) * tiles_n;
        const unsigned char *bp0 = Bp + (size_t)(c0 / T5_N) * np * T5_B_PLANE;
        if (mask7)
        {
            k_ctpt_gemm_tc5<7><<<(unsigned)ctas, 160, 2 * 7 * T5_B_PLANE, stream>>>(
                X, bp0, Y, K, cn, Kp, Cp, np, tiles_n, limbs, c->log_n, c->d_limb, c->d_two64, mask7, pl_first);
            c->launches += 1;
        }
        if (mask6)
        {
            k_ctpt_gemm_tc5<6><<<(unsigned)ctas, 160, 2 * 6 * T5_B_PLANE, stream>>>(
                X, bp0, Y, K, cn, Kp, Cp, np, tiles_n, limbs, c->log_n, c->d_limb, c->d_two64, mask6, pl_first);
            c->launches += 1;
        }
        MOAI_CUDA_CHECK(cudaGetLastError());
    }
} // namespace moai
