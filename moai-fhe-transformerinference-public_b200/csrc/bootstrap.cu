// Full-slot CKKS bootstrapping (see bootstrap.hpp for the design and the reference pointers).
#include <cstdlib>
#include "bootstrap.hpp"
#include <algorithm>
#include <set>

namespace moai
{
    namespace
    {
        const double PI = 3.14159265358979323846264338327950288;

        typedef std::map<int, std::vector<cd>> DiagMat; // offset (mod n, in [0,n)) -> diagonal

        // (A * B)[p] : out = A (B in);  diag_{a+b}[p] += A_a[p] * B_b[(p + a) mod n]
        DiagMat mat_mul(const DiagMat &A, const DiagMat &B, int n)
        {
            DiagMat R;
            for (auto &ka : A)
            {
                for (auto &kb : B)
                {
                    const int a = ka.first, b = kb.first;
                    const int d = (a + b) % n;
                    auto &dst = R[d];
                    if (dst.empty())
                    {
                        dst.assign(n, cd(0, 0));
                    }
                    const auto &va = ka.second;
                    const auto &vb = kb.second;
                    for (int p = 0; p < n; p++)
                    {
                        dst[p] += va[p] * vb[(p + a) % n];
                    }
                }
            }
            // drop numerically empty diagonals
            for (auto it = R.begin(); it != R.end();)
            {
                double mx = 0;
                for (auto &z : it->second)
                {
                    mx = std::max(mx, std::abs(z));
                }
                it = mx == 0 ? R.erase(it) : std::next(it);
            }
            return R;
        }

        // One radix-2 level of the CKKS special FFT on n slots with block length len.
        //   forward (SlotToCoeff direction):  out[i+j] = u + ksi v, out[i+j+len/2] = u - ksi v
        //   inverse (CoeffToSlot direction):  out[i+j] = (u + v)/2, out[i+j+len/2] = (u - v) conj(ksi)/2
        // ksi_j = exp(2 pi i (5^j mod 4 len) / (4 len))   (slot j <-> root zeta^{5^j}, S/ckks.cpp:36-52)
        DiagMat fft_level(int n, int len, bool inverse)
        {
            const int lenh = len / 2, lenq = len * 4;
            std::vector<cd> ksi(lenh);
            long long pw = 1;
            for (int j = 0; j < lenh; j++)
            {
                const double ang = 2 * PI * (double)(pw % lenq) / (double)lenq;
                ksi[j] = cd(std::cos(ang), std::sin(ang));
                pw = (pw * 5) % lenq; // 5^j mod 4 len
            }
            DiagMat M;
            auto &d0 = M[0];
            d0.assign(n, cd(0, 0));
            const int up = lenh % n, dn = (n - lenh) % n;
            auto &dup = M[up];
            if (dup.empty())
            {
                dup.assign(n, cd(0, 0));
            }
            auto &ddn = M[dn];
            if (ddn.empty())
            {
                ddn.assign(n, cd(0, 0));
            }
            for (int p = 0; p < n; p++)
            {
                const int j = p % len;
                if (j < lenh)
                {
                    if (!inverse)
                    {
                        M[0][p] += cd(1, 0);
                        M[up][p] += ksi[j];
                    }
                    else
                    {
                        M[0][p] += cd(0.5, 0);
                        M[up][p] += cd(0.5, 0);
                    }
                }
                else
                {
                    const cd k = ksi[j - lenh];
                    if (!inverse)
                    {
                        M[dn][p] += cd(1, 0);
                        M[0][p] += -k;
                    }
                    else
                    {
                        M[dn][p] += std::conj(k) * 0.5;
                        M[0][p] += -std::conj(k) * 0.5;
                    }
                }
            }
            return M;
        }

        void scale_mat(DiagMat &M, double f)
        {
            for (auto &kv : M)
            {
                for (auto &z : kv.second)
                {
                    z *= f;
                }
            }
        }

        // Householder least squares (double precision is ample here: the fitted coefficients are O(1))
        std::vector<double> lstsq(std::vector<double> A, std::vector<double> b, int m, int ncol)
        {
            auto at = [&](int r, int cc) -> double & { return A[(size_t)r * ncol + cc]; };
            for (int k = 0; k < ncol; k++)
            {
                double nv = 0;
                for (int r = k; r < m; r++)
                {
                    nv += at(r, k) * at(r, k);
                }
                nv = std::sqrt(nv);
                if (nv == 0)
                {
                    continue;
                }
                std::vector<double> v(m - k);
                for (int r = k; r < m; r++)
                {
                    v[r - k] = at(r, k);
                }
                v[0] += v[0] >= 0 ? nv : -nv;
                double vn = 0;
                for (double t : v)
                {
                    vn += t * t;
                }
                vn = std::sqrt(vn);
                for (double &t : v)
                {
                    t /= vn;
                }
                for (int cc = k; cc < ncol; cc++)
                {
                    double dot = 0;
                    for (int r = k; r < m; r++)
                    {
                        dot += v[r - k] * at(r, cc);
                    }
                    for (int r = k; r < m; r++)
                    {
                        at(r, cc) -= 2 * v[r - k] * dot;
                    }
                }
                double dot = 0;
                for (int r = k; r < m; r++)
                {
                    dot += v[r - k] * b[r];
                }
                for (int r = k; r < m; r++)
                {
                    b[r] -= 2 * v[r - k] * dot;
                }
            }
            std::vector<double> x(ncol, 0.0);
            for (int k = ncol - 1; k >= 0; k--)
            {
                double s = b[k];
                for (int cc = k + 1; cc < ncol; cc++)
                {
                    s -= at(k, cc) * x[cc];
                }
                x[k] = s / at(k, k);
            }
            return x;
        }
    } // namespace

    // ------------------------------------------------------------------------------------ set-up
    Bootstrapper::Bootstrapper(Context *ctx, const BootParams &p) : prm(p), c_(ctx)
    {
        MOAI_REQUIRE(p.total_limbs >= 15 && p.total_limbs <= ctx->kl - 1, "bootstrapping needs at least 15 data limbs");
        MOAI_REQUIRE(p.deg >= 15 && p.deg <= 63, "cosine degree must be in [15, 63]");
        MOAI_REQUIRE(p.double_angles >= 0 && p.poly_levels() + p.double_angles == 8,
                     "EvalMod must spend 8 levels: ceil(log2(deg + 1)) + double_angles == 8, e.g. (59, 2) or (31, 3)");
        build_matrices();
        fit_cosine();
    }

    void Bootstrapper::build_matrices()
    {
        const int n = slots();
        int logn = 0;
        while ((1 << logn) < n)
        {
            logn++;
        }
        // levels per stage like the reference's 3-way split (Bootstrapper.cpp:90-92)
        const int d1 = logn / 3, d2 = (logn - d1) / 2, d3 = logn - d1 - d2;
        const int grp[3] = { d1, d2, d3 };
        // CoeffToSlot: inverse levels applied len = n, n/2, ..., 2 ; the first applied factor is rightmost
        {
            int len = n;
            for (int s = 0; s < 3; s++)
            {
                DiagMat M;
                for (int k = 0; k < grp[s]; k++, len >>= 1)
                {
                    DiagMat Lm = fft_level(n, len, true);
                    M = M.empty() ? Lm : mat_mul(Lm, M, n);
                }
                // fold 1/(2K): 1/2 for the real/imaginary extraction, 1/K to land in [-1, 1]
                scale_mat(M, std::cbrt(1.0 / (2.0 * prm.boundary_K)));
                cts_[s].diags.clear();
                for (auto &kv : M)
                {
                    const int off = kv.first > n / 2 ? kv.first - n : kv.first;
                    cts_[s].diags[off] = kv.second;
                }
                cts_[s].limbs = prm.total_limbs - s;
                cts_[s].first = s == 0;
                plan_bsgs(cts_[s]);
            }
        }
        // SlotToCoeff: forward levels applied len = 2, 4, ..., n (constants folded at encode time)
        {
            int len = 2;
            for (int s = 0; s < 3; s++)
            {
                DiagMat M;
                for (int k = 0; k < grp[2 - s]; k++, len <<= 1)
                {
                    DiagMat Lm = fft_level(n, len, false);
                    M = M.empty() ? Lm : mat_mul(Lm, M, n);
                }
                stc_[s].diags.clear();
                for (auto &kv : M)
                {
                    const int off = kv.first > n / 2 ? kv.first - n : kv.first;
                    stc_[s].diags[off] = kv.second;
                }
                stc_[s].limbs = prm.total_limbs - 3 - (prm.poly_levels() + prm.double_angles) - s;
                plan_bsgs(stc_[s]);
            }
        }
    }

    void Bootstrapper::plan_bsgs(LinearStage &st) const
    {
        // every offset is a multiple of the smallest non-zero |offset|
        int stride = 0;
        for (auto &kv : st.diags)
        {
            if (kv.first != 0)
            {
                const int a = std::abs(kv.first);
                stride = stride == 0 ? a : std::min(stride, a);
            }
        }
        stride = stride == 0 ? 1 : stride;
        for (auto &kv : st.diags)
        {
            MOAI_REQUIRE(kv.first % stride == 0, "diagonal offsets are not multiples of a common stride");
        }
        const int cnt = (int)st.diags.size();
        auto split = [&](int g, std::set<int> &bs, std::set<int> &gs) {
            bs.clear();
            gs.clear();
            for (auto &kv : st.diags)
            {
                const int e = kv.first / stride;
                const int i = g > 0 ? (int)std::floor((double)e / g) : 0; // g == 0: baby steps only (signed j)
                bs.insert(e - i * g);
                gs.insert(i);
            }
        };
        int g = 1;
        while (g * g < cnt)
        {
            g++;
        }
        std::set<int> bs, gs;
        if (prm.hoisting == 2)
        {
            // lazy mod-down: a baby step costs an inner product and one gather-multiply pass, a giant step a full
            // key switch plus the mod-down of its inner sum: ~1 : 8 (1 : 50 on the first stage, whose baby steps run
            // on single-digit keys while its giant steps would pay SEAL's 35-digit key switch: baby steps only)
            double best = 1e300;
            int best_g = g;
            const int max_baby = st.first ? 64 : BSGS_MAX_BABY;
            for (int cand = st.first ? 0 : 1; cand <= 64; cand++)
            {
                split(cand, bs, gs);
                if ((int)bs.size() > max_baby || (int)gs.size() > BSGS_MAX_GIANT)
                {
                    continue;
                }
                int nb = 0, ng = 0;
                for (int j : bs)
                {
                    nb += j != 0;
                }
                for (int i : gs)
                {
                    ng += i != 0;
                }
                const double cost = nb * 1.0 + ng * (st.first ? 50.0 : 8.0) + (int)gs.size() * (st.first ? 3.0 : 1.0);
                if (cost < best)
                {
                    best = cost;
                    best_g = cand;
                }
            }
            g = best_g;
        }
        else if (prm.hoisting)
        {
            // baby steps share one digit decomposition (hoisted: inner product + mod-down only), giant
            // steps pay a full key switch: weigh them ~1 : 3.5 (measured at 35 limbs) and stay inside
            // the fused inner-sum kernel's limits
            double best = 1e300;
            int best_g = g;
            for (int cand = 1; cand <= BSGS_MAX_BABY; cand++)
            {
                split(cand, bs, gs);
                if ((int)bs.size() > BSGS_MAX_BABY || (int)gs.size() > BSGS_MAX_GIANT)
                {
                    continue;
                }
                int nb = 0, ng = 0;
                for (int j : bs)
                {
                    nb += j != 0;
                }
                for (int i : gs)
                {
                    ng += i != 0;
                }
                const double cost = nb * 1.0 + ng * 3.5;
                if (cost < best)
                {
                    best = cost;
                    best_g = cand;
                }
            }
            g = best_g;
        }
        split(g, bs, gs);
        st.stride = stride;
        st.giant = g;
        st.baby.assign(bs.begin(), bs.end());
        st.giants.assign(gs.begin(), gs.end());
    }

    void Bootstrapper::set_hoisting(int mode)
    {
        MOAI_REQUIRE(mode >= 0 && mode <= 2, "hoisting mode must be 0, 1 or 2");
        prm.hoisting = mode;
        for (int s = 0; s < 3; s++)
        {
            cts_[s].first = s == 0;
            plan_bsgs(cts_[s]);
            plan_bsgs(stc_[s]);
            cts_[s].pts.clear(); // pre-rotation of the diagonals depends on the plan
            stc_[s].pts.clear();
            cts_[s].pts_ext.clear();
            stc_[s].pts_ext.clear();
            cts_[s].pt_scale = 0;
            stc_[s].pt_scale = 0;
        }
        stc_encoded_for_scale_ = 0;
    }

    std::vector<int> Bootstrapper::single_digit_steps() const
    {
        std::vector<int> out;
        if (prm.hoisting != 2)
        {
            return out;
        }
        const int n = slots();
        const LinearStage &st = cts_[0];
        for (int j : st.baby)
        {
            if (j)
            {
                out.push_back(((j * st.stride) % n + n) % n);
            }
        }
        return out;
    }

    std::vector<int> Bootstrapper::required_steps() const
    {
        const int n = slots();
        std::set<int> steps;
        for (int dir = 0; dir < 2; dir++)
        {
            for (int s = 0; s < 3; s++)
            {
                const LinearStage &st = dir == 0 ? cts_[s] : stc_[s];
                for (int j : st.baby)
                {
                    if (j)
                    {
                        steps.insert(((j * st.stride) % n + n) % n);
                    }
                }
                for (int i : st.giants)
                {
                    if (i)
                    {
                        steps.insert(((long long)i * st.giant * st.stride % n + n) % n);
                    }
                }
            }
        }
        return std::vector<int>(steps.begin(), steps.end());
    }

    std::vector<std::pair<int, int>> Bootstrapper::required_step_levels() const
    {
        const int n = slots();
        std::set<std::pair<int, int>> out;
        for (int dir = 0; dir < 2; dir++)
        {
            for (int s = 0; s < 3; s++)
            {
                const LinearStage &st = dir == 0 ? cts_[s] : stc_[s];
                const bool single = prm.hoisting == 2 && dir == 0 && s == 0; // level 0 = single-digit key
                for (int j : st.baby)
                {
                    if (j)
                    {
                        out.insert({ ((j * st.stride) % n + n) % n, single ? 0 : st.limbs });
                    }
                }
                for (int i : st.giants)
                {
                    if (i)
                    {
                        out.insert({ (int)(((long long)i * st.giant * st.stride % n + n) % n), st.limbs });
                    }
                }
            }
        }
        out.insert({ 0, cts_[2].limbs - 1 });       // conjugation after CoeffToSlot
        out.insert({ 0, prm.total_limbs - 14 });    // separation of the two real-slot messages
        return std::vector<std::pair<int, int>>(out.begin(), out.end());
    }

    void Bootstrapper::prepare_stage(LinearStage &st, double pt_scale, double diag_factor)
    {
        st.pt_scale = pt_scale;
        st.diag_factor = diag_factor;
        st.pts.clear();
        st.pts_ext.clear();
        st.ext_layout = -99;
    }

    void Bootstrapper::encode_stage(const Evaluator &ev, LinearStage &st, const std::vector<int> *ids)
    {
        const int n = slots();
        MOAI_REQUIRE(st.pt_scale > 0, "linear stage has no plaintext scale yet");
        auto &dst = ids ? st.pts_ext : st.pts;
        dst.clear();
        // all pre-rotated diagonals of the stage in one batched device encode
        std::vector<std::pair<int, int>> keys;
        std::vector<cd> vals;
        for (auto &kv : st.diags)
        {
            const int e = kv.first / st.stride;
            const int i = st.giant > 0 ? (int)std::floor((double)e / st.giant) : 0;
            const int j = e - i * st.giant;
            const long long G = (long long)i * st.giant * st.stride;
            keys.push_back({ i, j });
            // P[p] = diag[(p - G) mod n]
            for (int p = 0; p < n; p++)
            {
                vals.push_back(kv.second[(((long long)p - G) % n + n) % n] * st.diag_factor);
            }
        }
        Pt all;
        int per = st.limbs;
        if (!ids)
        {
            all = ev.encode_batch(vals.data(), (long long)keys.size(), n, st.limbs, st.pt_scale);
        }
        else
        {
            // encode over the whole chain, keep the listed primes: {0 .. l-1} and a trailing run {L-k .. L}
            const int kl = c_->kl;
            per = (int)ids->size();
            int lead = 0;
            while (lead < per && (*ids)[lead] == lead)
            {
                lead++;
            }
            for (int t = lead; t < per; t++)
            {
                MOAI_REQUIRE((*ids)[t] == kl - (per - t), "key-switch basis is not {0..l-1} + a trailing run of primes");
            }
            Pt full = ev.encode_batch(vals.data(), (long long)keys.size(), n, kl, st.pt_scale);
            all.limbs = per;
            all.scale = st.pt_scale;
            all.count = (long long)keys.size();
            all.buf = std::make_shared<DevBuf>(keys.size() * (size_t)per * c_->n * sizeof(u64), c_->stream);
            all.d = reinterpret_cast<u64 *>(all.buf->p);
            const size_t limb_b = c_->n * sizeof(u64);
            MOAI_CUDA_CHECK(cudaMemcpy2DAsync(all.d, per * limb_b, full.d, kl * limb_b, lead * limb_b, keys.size(),
                                              cudaMemcpyDeviceToDevice, c_->stream));
            MOAI_CUDA_CHECK(cudaMemcpy2DAsync(all.d + (size_t)lead * c_->n, per * limb_b,
                                              full.d + (size_t)(kl - (per - lead)) * c_->n, kl * limb_b,
                                              (per - lead) * limb_b, keys.size(), cudaMemcpyDeviceToDevice, c_->stream));
        }
        for (size_t k = 0; k < keys.size(); k++)
        {
            Pt one = all;
            one.d = all.d + k * (size_t)per * c_->n;
            one.count = 1;
            dst[keys[k]] = one;
        }
    }

    void Bootstrapper::fit_cosine()
    {
        // least squares on Chebyshev nodes of every interval [i - w, i + w], |i| < K, y = x / K
        const int K = prm.boundary_K, deg = prm.deg, pts_per = 8;
        const double w = std::ldexp(1.0, -prm.log_width);
        const int m = (2 * K - 1) * pts_per, ncol = deg + 1;
        std::vector<double> A((size_t)m * ncol), b(m);
        int row = 0;
        for (int i = -(K - 1); i <= K - 1; i++)
        {
            for (int t = 0; t < pts_per; t++, row++)
            {
                const double x = i + w * std::cos(PI * (t + 0.5) / pts_per);
                const double y = x / K;
                double t0 = 1, t1 = y;
                A[(size_t)row * ncol] = 1;
                A[(size_t)row * ncol + 1] = y;
                for (int k = 2; k <= deg; k++)
                {
                    const double t2 = 2 * y * t1 - t0;
                    A[(size_t)row * ncol + k] = t2;
                    t0 = t1;
                    t1 = t2;
                }
                b[row] = std::cos(2 * PI * (x - 0.25) / std::ldexp(1.0, prm.double_angles));
            }
        }
        cheb_ = lstsq(A, b, m, ncol);
    }

    // ------------------------------------------------------------------------------------ linear transforms
    Ct Bootstrapper::finish_giants(const Evaluator &ev, const Ct &ct, LinearStage &st, const Keys &keys,
                                   const std::vector<int> &gi, std::vector<Ct> &inner) const
    {
        const long long n = slots();
        auto norm = [&](long long step) { return (int)(((step % n) + n) % n); };
        Ct acc;
        for (size_t k = 0; k < gi.size(); k++)
        {
            const int gstep = norm((long long)gi[k] * st.giant * st.stride);
            Ct term = gstep != 0 ? ev.rotate_vector(inner[k], gstep, keys) : inner[k];
            inner[k] = Ct();
            if (acc.empty())
            {
                acc = term;
            }
            else
            {
                ev.add_inplace(acc, term);
            }
        }
        (void)ct;
        return ev.rescale_to_next(acc);
    }

    Ct Bootstrapper::linear_transform(const Evaluator &ev, const Ct &ct, LinearStage &st, const Keys &keys)
    {
        MOAI_REQUIRE(ct.limbs == st.limbs, "linear stage applied at the wrong level");
        if (prm.hoisting == 2)
        {
            Ct r = linear_transform_lazy(ev, ct, st, keys);
            if (!r.empty())
            {
                return r;
            }
        }
        if (st.pts.empty())
        {
            encode_stage(ev, st, nullptr);
        }
        const long long n = slots();
        auto norm = [&](long long step) { return (int)(((step % n) + n) % n); }; // left rotation in [0, slots)
        // baby steps: rotations of the same ciphertexts (hoisted when the keys are pre-permuted)
        std::vector<int> bsteps;
        for (int j : st.baby)
        {
            bsteps.push_back(norm((long long)j * st.stride));
        }
        MOAI_REQUIRE((int)st.baby.size() <= BSGS_MAX_BABY, "BSGS plan exceeds the fused kernel's limits");
        std::vector<Ct> rots = ev.rotate_many(ct, bsteps, keys);
        // every giant step's inner sum  sum_j P[i][j] (.) rot_j  in one fused pass
        const int nb = (int)st.baby.size();
        std::vector<int> gi;
        std::vector<const u64 *> pts;
        for (int i : st.giants)
        {
            bool any = false;
            for (int j : st.baby)
            {
                any = any || st.pts.count({ i, j });
            }
            if (!any)
            {
                continue;
            }
            gi.push_back(i);
            for (int j : st.baby)
            {
                auto it = st.pts.find({ i, j });
                pts.push_back(it == st.pts.end() ? nullptr : it->second.d);
            }
        }
        MOAI_REQUIRE(!gi.empty(), "empty linear stage");
        std::vector<Ct> inner(gi.size());
        std::vector<const u64 *> rp;
        std::vector<u64 *> op;
        for (auto &r : rots)
        {
            rp.push_back(r.d);
        }
        for (auto &x : inner)
        {
            x = ev.alloc(ct.batch, 2, ct.limbs, ct.scale * st.pt_scale);
            op.push_back(x.d);
        }
        bsgs_inner(c_, rp.data(), nb, pts.data(), (int)gi.size(), op.data(), ct.batch, ct.limbs);
        rots.clear();
        return finish_giants(ev, ct, st, keys, gi, inner);
    }

    // Lazy mod-down ("double hoisting"): rot_r(ct) P' = sigma_r(acc_r + (P' c0, 0)) stays in the key-switch basis, the
    // inner sums are formed there (k_bsgs_ext) and every giant step's inner sum is divided by P' ONCE.  The first
    // CoeffToSlot stage sees the mod-raised ciphertext, whose c1 is a single small digit (ksgroup.hpp): with single-digit
    // keys its "decomposition" is one NTT and its baby steps cost 2 x 36 limb products each, so the stage is planned
    // with baby steps only.  Returns an empty Ct when the registered keys do not allow the lazy path.
    Ct Bootstrapper::linear_transform_lazy(const Evaluator &ev, const Ct &ct, LinearStage &st, const Keys &keys) const
    {
        Context *c = c_;
        const size_t N = c->n;
        const long long n = slots();
        const int limbs = ct.limbs;
        auto norm = [&](long long step) { return (int)(((step % n) + n) % n); };
        std::vector<uint32_t> elts; // per baby step (0 = the identity)
        for (int j : st.baby)
        {
            const int step = norm((long long)j * st.stride);
            elts.push_back(step == 0 ? 0u : c->elt_from_step(step));
        }
        // digit layout: single-digit keys on the mod-raised input, else the cheapest layout every baby step has a key for
        int layout = -99;
        if (st.first)
        {
            bool ok = true;
            for (uint32_t e : elts)
            {
                ok = ok && (e == 0 || keys.galois_single.count(e));
            }
            if (ok)
            {
                layout = KS_SINGLE;
            }
        }
        if (layout == -99)
        {
            double best_cost = 0;
            for (int cand = 0; limbs + cand <= c->kl - 1; cand++)
            {
                bool ok = true;
                for (uint32_t e : elts)
                {
                    ok = ok && (e == 0 || keys.fast(c, e, limbs, cand) != nullptr);
                }
                const double cost = ok ? ksg_cost(c, limbs, cand) : 0;
                if (ok && (layout == -99 || cost < best_cost))
                {
                    layout = cand;
                    best_cost = cost;
                }
            }
        }
        if (layout == -99)
        {
            return Ct();
        }
        const KsExtInfo info = layout == KS_SINGLE ? ks_ext_info_single(c, limbs) : ks_ext_info(c, layout, limbs);
        const KsShape &sh = info.shape;
        auto *self = const_cast<Bootstrapper *>(this);
        if (st.pts_ext.empty() || st.ext_layout != layout)
        {
            self->encode_stage(ev, st, &info.h_ids);
            st.ext_layout = layout;
            st.pts.clear(); // the level-basis copies are not needed on this path
        }
        // giant steps present and their plaintext rows (one pointer per baby step)
        const int nbaby = (int)st.baby.size();
        std::vector<int> gi;
        std::vector<const u64 *> pts;
        for (int i : st.giants)
        {
            bool any = false;
            for (int j : st.baby)
            {
                any = any || st.pts_ext.count({ i, j });
            }
            if (!any)
            {
                continue;
            }
            gi.push_back(i);
            for (int j : st.baby)
            {
                auto it = st.pts_ext.find({ i, j });
                pts.push_back(it == st.pts_ext.end() ? nullptr : it->second.d);
            }
        }
        MOAI_REQUIRE(!gi.empty() && gi.size() <= (size_t)BSGS_MAX_GIANT, "BSGS plan exceeds the fused kernel's limits");
        const int G = (int)gi.size();
        // Lazy giant steps: the giants' key switches stay in the key-switch basis too.  Only the c1 of a giant's inner
        // sum is divided by P' (it has to be decomposed again); its c0 and the giant's key-switch sums are rotated and
        // added in the extended basis, and the stage's ONE remaining division is by P' q_last — the rescale that ends
        // the stage rides on it (ksg_moddown_rescale).  Per stage with 4 giants: 5 polynomial mod-downs instead of
        // 14 + 2 rescales.  Needs grouped keys of the stage's layout for every giant step (MOAI_LAZY_GIANTS=0 disables).
        static const bool lazy_giants_on = [] {
            const char *e = std::getenv("MOAI_LAZY_GIANTS");
            return !(e && e[0] == '0');
        }();
        std::vector<uint32_t> gelts(G, 0);
        std::vector<const KeyRef *> gkeys(G, nullptr);
        int n_gsw = 0; // giants that need a key switch
        bool lazy_giants = lazy_giants_on && limbs >= 2 && (layout > 0 || (layout == KS_SINGLE && G == 1));
        for (int g = 0; g < G && lazy_giants; g++)
        {
            const int gstep = norm((long long)gi[g] * st.giant * st.stride);
            if (gstep == 0)
            {
                continue;
            }
            gelts[g] = c->elt_from_step(gstep);
            gkeys[g] = layout > 0 ? keys.fast(c, gelts[g], limbs, layout) : nullptr;
            lazy_giants = gkeys[g] != nullptr;
            n_gsw++;
        }
        std::vector<Ct> inner(lazy_giants ? 0 : G);
        for (auto &x : inner)
        {
            x = ev.alloc(ct.batch, 2, limbs, ct.scale * st.pt_scale);
        }
        Ct lazy_out;
        if (lazy_giants)
        {
            lazy_out = ev.alloc(ct.batch, 2, limbs - 1, ct.scale * st.pt_scale / ev.last_prime(limbs));
        }
        const size_t per_out = (size_t)2 * (limbs - 1) * N;
        // keys of the non-identity baby steps
        std::vector<int> rot_idx; // baby index of every rotation that needs a key switch
        std::vector<const KeyRef *> rkeys;
        for (int j = 0; j < nbaby; j++)
        {
            if (elts[j] == 0)
            {
                continue;
            }
            rot_idx.push_back(j);
            rkeys.push_back(layout == KS_SINGLE ? &keys.galois_single.at(elts[j]) : keys.fast(c, elts[j], limbs, layout));
        }
        const int R = (int)rot_idx.size();
        if (layout == KS_SINGLE && G != 1)
        {
            return Ct(); // the single-digit stage is planned with baby steps only
        }
        // chunk: bound the workspace (digits + one inner product per rotation + the giants' sums) to the key-switch budget
        const size_t acc_words = (size_t)2 * sh.rns * N;
        const size_t ext_bytes = layout == KS_SINGLE ? ks_single_ext_bytes_per_ct(c, limbs)
                                 : (layout > 0 ? ksg_ext_bytes_per_ct(c, limbs, layout) : ks_ext_bytes_per_ct(c, limbs));
        const int acc_slots = layout == KS_SINGLE ? 0 : std::max(std::max(R, 1), lazy_giants ? n_gsw + 1 : 0);
        const size_t per_ct_ws = ext_bytes + (size_t)(acc_slots + G) * acc_words * sizeof(u64);
        long long chunk = std::max<long long>(1, (long long)(ks_ext_budget() / per_ct_ws));
        chunk = std::min<long long>(chunk, ct.batch);
        {
            const long long parts = (ct.batch + chunk - 1) / chunk;
            chunk = (ct.batch + parts - 1) / parts; // equal chunks, no short tail
        }
        Scratch ext((size_t)chunk * ext_bytes, c->stream);
        Scratch accs((size_t)(layout == KS_SINGLE ? 1 : acc_slots * chunk * acc_words) * sizeof(u64), c->stream);
        Scratch outs((size_t)G * chunk * acc_words * sizeof(u64), c->stream);
        Scratch cP((size_t)chunk * 2 * limbs * N * sizeof(u64), c->stream);
        const size_t per_ct = (size_t)2 * limbs * N;
        for (long long b0 = 0; b0 < ct.batch; b0 += chunk)
        {
            const long long nb = std::min(chunk, ct.batch - b0);
            const u64 *src = ct.d + (size_t)b0 * per_ct;
            if (layout == KS_SINGLE)
            {
                ks_hoist_modraised(c, src + (size_t)limbs * N, nb, limbs, ext.as<u64>(), (long long)per_ct);
            }
            else if (layout > 0)
            {
                ksg_decompose(c, src + (size_t)limbs * N, nb, limbs, layout, ext.as<u64>(), (long long)per_ct, 3);
            }
            else
            {
                ks_decompose(c, src + (size_t)limbs * N, nb, limbs, ext.as<u64>(), (long long)per_ct);
            }
            ew_multiply_scalar(c, src, info.h_pmod.data(), cP.as<u64>(), nb, 2, limbs);
            if (layout == KS_SINGLE)
            {
                // one fused pass per BSGS_MAX_BABY rotations: gather of the digit, product with the natural-order key,
                // product with the diagonal (k_bsgs_single)
                u64 *o = outs.as<u64>();
                int ri = 0;
                for (int j0 = 0; j0 < nbaby; j0 += BSGS_MAX_BABY)
                {
                    const int cnt = std::min(BSGS_MAX_BABY, nbaby - j0);
                    std::vector<const u64 *> kp(cnt), pp(cnt);
                    std::vector<const uint32_t *> perm(cnt);
                    for (int j = 0; j < cnt; j++)
                    {
                        const bool ident = elts[j0 + j] == 0;
                        kp[j] = ident ? nullptr : rkeys[ri]->p;
                        perm[j] = ident ? nullptr : c->galois_table(elts[j0 + j]);
                        ri += ident ? 0 : 1;
                        pp[j] = pts[j0 + j];
                    }
                    bsgs_single(c, ext.as<u64>(), kp.data(), perm.data(), pp.data(), cnt, c->kl, o, cP.as<u64>(), nb, sh, j0 > 0);
                }
                if (lazy_giants) // no giant rotation: mod-down and the stage's rescale in one division
                {
                    ksg_moddown_rescale(c, o, nb * 2, limbs, 0, nullptr, false, lazy_out.d + (size_t)b0 * per_out);
                    continue;
                }
                ks_moddown(c, o, nb * 2, limbs, 0, nullptr, false, inner[0].d + (size_t)b0 * per_ct);
                continue;
            }
            // inner products of every rotation with its key (no mod-down), KSM_R keys per pass over the digits
            for (int r0 = 0; r0 < R; r0 += KSM_R)
            {
                const int cnt = std::min(KSM_R, R - r0);
                const u64 *kp[KSM_R];
                int kkl[KSM_R];
                u64 *accp[KSM_R];
                for (int r = 0; r < cnt; r++)
                {
                    kp[r] = rkeys[r0 + r]->p;
                    kkl[r] = rkeys[r0 + r]->key_kl;
                    accp[r] = accs.as<u64>() + (size_t)(r0 + r) * nb * acc_words;
                    MOAI_REQUIRE(kkl[r] >= sh.rns, "key does not cover this level");
                }
                ks_mac_multi(c, ext.as<u64>(), nb, sh, cnt, kp, kkl, accp);
                for (int r = 0; r < cnt; r++)
                {
                    ks_int_targets(c, sh, info.h_ids, ext.as<u64>(), nb, kp[r], kkl[r], accp[r], false);
                }
            }
            // inner sums of every giant step in the key-switch basis, BSGS_MAX_BABY baby steps per pass
            std::vector<u64 *> op(G);
            for (int g = 0; g < G; g++)
            {
                op[g] = outs.as<u64>() + (size_t)g * nb * acc_words;
            }
            int ri = 0;
            for (int j0 = 0; j0 < nbaby; j0 += BSGS_MAX_BABY)
            {
                const int cnt = std::min(BSGS_MAX_BABY, nbaby - j0);
                std::vector<const u64 *> ap(cnt), pp((size_t)G * cnt);
                std::vector<const uint32_t *> perm(cnt);
                for (int j = 0; j < cnt; j++)
                {
                    if (elts[j0 + j] == 0)
                    {
                        ap[j] = nullptr;
                        perm[j] = nullptr;
                    }
                    else
                    {
                        ap[j] = accs.as<u64>() + (size_t)ri * nb * acc_words;
                        perm[j] = c->galois_table(elts[j0 + j]);
                        ri++;
                    }
                    for (int g = 0; g < G; g++)
                    {
                        pp[(size_t)g * cnt + j] = pts[(size_t)g * nbaby + j0 + j];
                    }
                }
                bsgs_ext(c, ap.data(), perm.data(), cnt, pp.data(), G, op.data(), cP.as<u64>(), nb, sh, j0 > 0);
            }
            if (lazy_giants)
            {
                // the baby steps' workspaces are free again: digits -> ext, (c1, direct) -> cP, sums -> accs
                u64 *c1buf = cP.as<u64>(), *direct = cP.as<u64>() + (size_t)nb * limbs * N;
                const u64 *accp[BSGS_MAX_GIANT], *extra[BSGS_MAX_GIANT];
                const uint32_t *perm[BSGS_MAX_GIANT];
                int slot = 0;
                for (int g = 0; g < G; g++)
                {
                    if (gelts[g] == 0)
                    {
                        accp[g] = op[g];
                        extra[g] = nullptr;
                        perm[g] = nullptr;
                        continue;
                    }
                    ksg_moddown(c, op[g] + (size_t)sh.rns * N, nb, limbs, layout, nullptr, false, c1buf, 2, 2);
                    u64 *acc_g = accs.as<u64>() + (size_t)slot++ * nb * acc_words;
                    ksg_switch_acc(c, c1buf, nb, limbs, layout, gkeys[g]->p, gkeys[g]->key_kl, ext.as<u64>(), direct, acc_g, 0);
                    accp[g] = acc_g;
                    extra[g] = op[g];
                    perm[g] = c->galois_table(gelts[g]);
                }
                u64 *total = accs.as<u64>() + (size_t)slot * nb * acc_words;
                ksg_giants_sum(c, G, accp, extra, perm, total, nb, limbs, layout);
                ksg_moddown_rescale(c, total, nb * 2, limbs, layout, nullptr, false, lazy_out.d + (size_t)b0 * per_out);
                continue;
            }
            for (int g = 0; g < G; g++)
            {
                ks_moddown(c, op[g], nb * 2, limbs, layout > 0 ? layout : 0, nullptr, false, inner[g].d + (size_t)b0 * per_ct);
            }
        }
        if (lazy_giants)
        {
            return lazy_out;
        }
        return finish_giants(ev, ct, const_cast<LinearStage &>(st), keys, gi, inner);
    }

    // MOAI_EVALMOD_FOLD=0: remainders and T_{a-b} take their own rescale (the pre-fold schedule)
    static bool fold_leaves()
    {
        static const bool on = [] {
            const char *e = std::getenv("MOAI_EVALMOD_FOLD");
            return !(e && e[0] == '0');
        }();
        return on;
    }

    // ------------------------------------------------------------------------------------ EvalMod
    // Chebyshev series sum coef[k] T_k evaluated to exactly (target_limbs, target_scale).
    // Recursive division p = q * T_g + r (g the largest giant <= deg p).  The remainder r is evaluated to the same
    // (level, scale) as p, so its own product q_r * T_{g/2} lands on the same level and the same scale as q * T_g:
    // the whole remainder chain q * T_g + q_r * T_{g/2} + ... is accumulated as size-3 ciphertexts and
    // relinearized + rescaled once (7 -> 4 key switches for the degree-59 cosine).
    void Bootstrapper::eval_cheb_parts(const Evaluator &ev, const std::vector<double> &coef, int target_limbs,
                                       double target_scale, const std::map<int, Ct> &T, const Keys &keys, Ct &acc3,
                                       Ct &rest) const
    {
        int d = (int)coef.size() - 1;
        while (d > 0 && coef[d] == 0.0)
        {
            d--;
        }
        const int kbaby = 8;
        if (d < kbaby)
        {
            // leaf: sum_j c_j T_j, every term brought to (target_limbs + 1) and scaled so that the
            // rescale lands exactly on target_scale
            const int lv = target_limbs + 1;
            const double ql = ev.last_prime(lv);
            // one fused pass: every T_j is read at its own level (the mod-switch to lv is free) and multiplied by its
            // constant encoded at target_scale * ql / T_j.scale, exactly what multiply_plain + add_inplace produced
            std::vector<Ct> terms;
            std::vector<double> cs;
            for (int j = 1; j <= d; j++)
            {
                if (coef[j] != 0.0)
                {
                    terms.push_back(T.at(j));
                    cs.push_back(coef[j]);
                }
            }
            Ct acc;
            if (!terms.empty())
            {
                acc = ev.lincomb_scalar(terms, cs, lv, target_scale * ql);
            }
            MOAI_REQUIRE(!acc.empty(), "degenerate polynomial leaf");
            if (!acc3.empty() && fold_leaves())
            {
                // the chain's products wait at (lv, target_scale * ql) for their one relinearize + rescale: the remainder
                // joins them there instead of paying a rescale of its own
                ev.add_plain_inplace(acc, ev.encode(coef[0], lv, target_scale * ql));
                ev.add_into3(acc3, acc);
                rest = Ct();
                return;
            }
            Ct r = ev.rescale_to_next(acc);
            r.scale = target_scale;
            rest = ev.add_plain(r, ev.encode(coef[0], r.limbs, target_scale));
            return;
        }
        // split at the largest giant g = 8 * 2^m <= d :  p = q * T_g + r
        int g = kbaby;
        while (g * 2 <= d)
        {
            g *= 2;
        }
        std::vector<double> q(d - g + 1, 0.0), r(g, 0.0);
        for (int i = 0; i < g; i++)
        {
            r[i] = coef[i];
        }
        q[0] = coef[g];
        for (int i = g + 1; i <= d; i++)
        {
            q[i - g] = 2 * coef[i];          // T_i = 2 T_g T_{i-g} - T_{2g-i}
            r[2 * g - i] -= coef[i];
        }
        const int lv = target_limbs + 1;
        const double ql = ev.last_prime(lv);
        const Ct &tgh = T.at(g); // read at level lv in place (no mod-switch copy)
        bool q_const = true;
        for (size_t i = 1; i < q.size(); i++)
        {
            q_const = q_const && q[i] == 0.0;
        }
        Ct prod2; // a constant quotient needs no ciphertext product
        if (q_const)
        {
            Ct tg = ev.mod_switch_to(tgh, lv);
            prod2 = ev.rescale_to_next(ev.multiply_plain(tg, ev.encode(q[0], lv, target_scale * ql / tg.scale)));
            prod2.scale = target_scale;
        }
        else
        {
            Ct qc = eval_cheb(ev, q, lv, target_scale * ql / tgh.scale, T, keys);
            MOAI_REQUIRE(qc.limbs == lv && tgh.limbs >= lv, "level bookkeeping of the Chebyshev division");
            Ct prod3 = ev.multiply_lowered(qc, tgh);
            prod3.scale = target_scale * ql; // qc was evaluated to exactly target_scale * ql / tg.scale
            if (acc3.empty())
            {
                acc3 = prod3;
            }
            else
            {
                ev.add_inplace(acc3, prod3);
            }
        }
        eval_cheb_parts(ev, r, target_limbs, target_scale, T, keys, acc3, rest);
        if (q_const)
        {
            rest = rest.empty() ? prod2 : ev.add(rest, prod2);
        }
    }

    Ct Bootstrapper::eval_cheb(const Evaluator &ev, const std::vector<double> &coef, int target_limbs,
                               double target_scale, const std::map<int, Ct> &T, const Keys &keys) const
    {
        Ct acc3, rest;
        eval_cheb_parts(ev, coef, target_limbs, target_scale, T, keys, acc3, rest);
        if (acc3.empty())
        {
            return rest;
        }
        Ct prod = ev.relin_rescale(acc3, keys);
        prod.scale = target_scale;
        return rest.empty() ? prod : ev.add(prod, rest);
    }

    Ct Bootstrapper::eval_mod(const Evaluator &ev, const Ct &y, const Keys &keys) const
    {
        // Chebyshev basis T_1..T_7 (baby) and T_8, T_16, T_32 (giant) of y = x / K
        std::map<int, Ct> T;
        T[1] = y;
        auto dbl_minus_one = [&](const Ct &sq) {
            Ct r = ev.relin_rescale(sq, keys);
            ev.double_add_const_inplace(r, -1.0); // one pass: 2 r - 1
            return r;
        };
        auto t_even = [&](int a) { return dbl_minus_one(ev.square(T.at(a))); }; // T_2a = 2 T_a^2 - 1
        auto t_sum = [&](int a, int b) {                                       // T_{a+b} = 2 T_a T_b - T_{a-b}, a > b
            Ct ta = T.at(a), tb = T.at(b);
            const int lv = std::min(ta.limbs, tb.limbs);
            Ct prod3 = ev.multiply_lowered(ta, tb);
            if (fold_leaves())
            {
                // 2 (T_a T_b - T_{a-b} / 2): the subtrahend joins the product BEFORE its relinearize + rescale, read at
                // its own (higher) level with the constant -1/2 encoded at prod.scale / T_{a-b}.scale — instead of the
                // multiply_const + rescale + mod-switch of sub_reduced_error (S/evaluator.cpp:478-534) afterwards
                const Ct &td = T.at(a - b);
                MOAI_REQUIRE(td.limbs >= lv, "level bookkeeping of the Chebyshev basis");
                Ct half = ev.lincomb_scalar({ td }, { -0.5 }, lv, prod3.scale);
                ev.add_into3(prod3, half);
                Ct p = ev.relin_rescale(prod3, keys);
                ev.double_inplace(p);
                return p;
            }
            Ct p = ev.relin_rescale(prod3, keys);
            ev.double_inplace(p);
            return ev.sub_reduced_error(p, T.at(a - b));
        };
        T[2] = t_even(1);
        T[3] = t_sum(2, 1);
        T[4] = t_even(2);
        T[5] = t_sum(3, 2);
        T[6] = t_even(3);
        T[7] = t_sum(4, 3);
        T[8] = t_even(4);
        if (prm.deg >= 16)
        {
            T[16] = t_even(8);
        }
        if (prm.deg >= 32)
        {
            T[32] = t_even(16);
        }
        Ct cosv = eval_cheb(ev, cheb_, y.limbs - prm.poly_levels(), y.scale, T, keys);
        for (int i = 0; i < prm.double_angles; i++)
        {
            cosv = dbl_minus_one(ev.square(cosv)); // cos(2a) = 2 cos(a)^2 - 1
        }
        return cosv;
    }

    // ------------------------------------------------------------------------------------ bootstrap
    Ct Bootstrapper::bootstrap(const Evaluator &ev, const Ct &in, const Keys &keys, int stop_after)
    {
        MOAI_REQUIRE(in.size == 2 && in.limbs == 1, "bootstrap expects size-2 ciphertexts at the last level");
        const int n = slots();
        const double q0 = (double)c_->q[0];
        // plaintexts of the linear stages: encoded once per (stage, input scale)
        if (cts_[0].pt_scale == 0)
        {
            for (int s = 0; s < 3; s++)
            {
                prepare_stage(cts_[s], ev.last_prime(cts_[s].limbs), 1.0);
            }
        }
        // 1. ModRaise; the plaintext is now t = m + q0 I, declared at scale q0
        Ct ct;
        {
            PhaseTimer t(c_, "boot_modraise");
            ct = ev.mod_raise(in, prm.total_limbs);
        }
        const double initial_scale = in.scale;
        ct.scale = q0;
        if (stop_after == 1)
        {
            return ct;
        }
        // 2. CoeffToSlot: slots <- (c_lo + i c_hi) / (2 K q0)   (bit-reversed order)
        for (int s = 0; s < 3; s++)
        {
            PhaseTimer t(c_, "boot_coeff_to_slot");
            ct = linear_transform(ev, ct, cts_[s], keys);
        }
        Ct conj = ev.complex_conjugate(ct, keys);
        Ct re = ev.add(ct, conj);                 // c_lo / (K q0)
        Ct im = ev.sub(ct, conj);                 // 2 i c_hi / (2 K q0)
        std::vector<cd> minus_i((size_t)n, cd(0, -1)), plus_i((size_t)n, cd(0, 1));
        im = ev.multiply_plain(im, ev.encode(minus_i, im.limbs, 1.0)); // exact monomial, no level
        if (stop_after == 2)
        {
            return ev.concat({ re, im });
        }
        // 3. EvalMod on both halves as one batch: sin(2 pi t / q0) ~ 2 pi m / q0
        Ct both;
        {
            PhaseTimer t(c_, "boot_eval_mod");
            both = eval_mod(ev, ev.concat({ re, im }), keys);
        }
        if (stop_after == 3)
        {
            return both;
        }
        re = ev.view(both, 0, in.batch);
        im = ev.view(both, in.batch, in.batch);
        // 4. SlotToCoeff on re + i im, constants q0 / (2 pi initial_scale) folded into the diagonals
        Ct w = ev.add(re, ev.multiply_plain(im, ev.encode(plus_i, im.limbs, 1.0)));
        if (stc_encoded_for_scale_ != initial_scale || stc_[0].pt_scale == 0)
        {
            const double f = std::cbrt(q0 / (2 * PI * initial_scale));
            for (int s = 0; s < 3; s++)
            {
                // the last stage's plaintext scale makes the final rescale land exactly on final_scale;
                // w.scale is invariant through the first two stages (plaintext scale = dropped prime)
                const double ps = s < 2 ? ev.last_prime(stc_[s].limbs)
                                        : prm.final_scale * ev.last_prime(stc_[s].limbs) / w.scale;
                prepare_stage(stc_[s], ps, f);
            }
            stc_encoded_for_scale_ = initial_scale;
        }
        MOAI_REQUIRE(w.limbs == stc_[0].limbs, "level budget mismatch before SlotToCoeff");
        for (int s = 0; s < 3; s++)
        {
            PhaseTimer t(c_, "boot_slot_to_coeff");
            w = linear_transform(ev, w, stc_[s], keys);
        }
        w.scale = prm.final_scale;
        return w;
    }
    Ct Bootstrapper::bootstrap_real_pairs(const Evaluator &ev, const Ct &in, const Keys &keys, long long chunk_pairs,
                                          const Ct *into, long long pair_first, long long pair_last)
    {
        MOAI_REQUIRE(in.size == 2 && in.limbs == 1, "bootstrap expects size-2 ciphertexts at the last level");
        MOAI_REQUIRE(chunk_pairs >= 1, "chunk must be positive");
        const long long B = in.batch, P = (B + 1) / 2;
        const int out_limbs = prm.total_limbs - 14;
        Ct out = into ? *into : ev.alloc(B, 2, out_limbs, prm.final_scale);
        MOAI_REQUIRE(out.batch == B && out.size == 2 && out.limbs == out_limbs, "output storage shape mismatch");
        out.scale = prm.final_scale;
        const size_t n = (size_t)slots();
        std::vector<cd> minus_i(n, cd(0, -1)), plus_i(n, cd(0, 1));
        const Pt pi_in = ev.encode(plus_i, 1, 1.0);           // X^(N/2): exact, no level
        const Pt mi_out = ev.encode(minus_i, out_limbs, 1.0);
        if (pair_last < 0)
        {
            pair_last = P;
        }
        MOAI_REQUIRE(pair_first >= 0 && pair_first <= pair_last && pair_last <= P, "bad pair range");
        for (long long p0 = pair_first; p0 < pair_last; p0 += chunk_pairs)
        {
            const long long np = std::min(chunk_pairs, pair_last - p0);
            const long long nb = std::max(0LL, std::min(np, B - P - p0)); // partners j + P < B
            Ct z = ev.clone(ev.view(in, p0, np));
            if (nb > 0)
            {
                Ct ib = ev.multiply_plain(ev.view(in, P + p0, nb), pi_in);
                Ct head = ev.view(z, 0, nb);
                ev.add_inplace(head, ib);
            }
            z.scale = 2.0 * in.scale; // the result then carries (a + i b) / 2 at final_scale
            Ct r = bootstrap(ev, z, keys);
            Ct cj = ev.complex_conjugate(r, keys);
            ev.copy_into(ev.add(r, cj), out, p0);
            if (nb > 0)
            {
                Ct d = ev.sub(ev.view(r, 0, nb), ev.view(cj, 0, nb));
                ev.copy_into(ev.multiply_plain(d, mi_out), out, P + p0);
            }
        }
        return out;
    }
} // namespace moai
