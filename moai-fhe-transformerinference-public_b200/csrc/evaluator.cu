// Implementation of the batched host-side Evaluator (see evaluator.hpp).  Every method keeps
// SEAL's metadata semantics; the cited lines are where the reference defines them.
#include <cstdlib>
#include "evaluator.hpp"
#include <algorithm>
#include <cstring>

namespace moai
{
#define EV_REQUIRE(cond, msg) MOAI_REQUIRE(cond, msg)

    // ---------------------------------------------------------------- storage
    Ct Evaluator::alloc(long long batch, int size, int limbs, double scale) const
    {
        Ct r;
        r.buf = std::make_shared<DevBuf>((size_t)batch * size * limbs * n() * sizeof(u64), c->stream);
        r.d = reinterpret_cast<u64 *>(r.buf->p);
        r.batch = batch;
        r.size = size;
        r.limbs = limbs;
        r.scale = scale;
        return r;
    }

    Ct Evaluator::wrap(u64 *d, long long batch, int size, int limbs, double scale) const
    {
        Ct r;
        r.d = d;
        r.batch = batch;
        r.size = size;
        r.limbs = limbs;
        r.scale = scale;
        return r;
    }

    Ct Evaluator::view(const Ct &a, long long b0, long long count) const
    {
        EV_REQUIRE(b0 >= 0 && count >= 0 && b0 + count <= a.batch, "view out of range");
        Ct r = a;
        r.d = a.d + (size_t)b0 * a.size * a.limbs * n();
        r.batch = count;
        return r;
    }

    Ct Evaluator::clone(const Ct &a) const
    {
        Ct r = alloc(a.batch, a.size, a.limbs, a.scale);
        { KernelTimer ktm(c, "k_copy_clone", 1); MOAI_CUDA_CHECK(cudaMemcpyAsync(r.d, a.d, (size_t)a.batch * a.size * a.limbs * n() * sizeof(u64),
                                        cudaMemcpyDeviceToDevice, c->stream)); }
        return r;
    }

    void Evaluator::copy_into(const Ct &src, Ct &dst, long long dst_b0) const
    {
        EV_REQUIRE(src.size == dst.size && src.limbs == dst.limbs && dst_b0 + src.batch <= dst.batch,
                   "copy_into shape mismatch");
        { KernelTimer ktm(c, "k_copy_into", 1); MOAI_CUDA_CHECK(cudaMemcpyAsync(dst.d + (size_t)dst_b0 * dst.size * dst.limbs * n(), src.d,
                                        (size_t)src.batch * src.size * src.limbs * n() * sizeof(u64),
                                        cudaMemcpyDeviceToDevice, c->stream)); }
    }

    Ct Evaluator::concat(const std::vector<Ct> &parts) const
    {
        EV_REQUIRE(!parts.empty(), "concat of nothing");
        long long total = 0;
        for (auto &p : parts)
        {
            total += p.batch;
        }
        Ct r = alloc(total, parts[0].size, parts[0].limbs, parts[0].scale);
        long long at = 0;
        for (auto &p : parts)
        {
            copy_into(p, r, at);
            at += p.batch;
        }
        return r;
    }

    Ct Evaluator::repeat(const Ct &a, long long times) const
    {
        EV_REQUIRE(a.batch == 1, "repeat expects a single ciphertext");
        Ct r = alloc(times, a.size, a.limbs, a.scale);
        for (long long i = 0; i < times; i++)
        {
            copy_into(a, r, i);
        }
        return r;
    }

    // ---------------------------------------------------------------- checks
    void Evaluator::check_same(const Ct &a, const Ct &b, bool need_scale) const
    {
        // S/evaluator.cpp:166-177: parms_id and (for add/sub) scale must match
        EV_REQUIRE(!a.empty() && !b.empty(), "encrypted is not valid for encryption parameters");
        EV_REQUIRE(a.limbs == b.limbs, "encrypted1 and encrypted2 parameter mismatch");
        EV_REQUIRE(a.batch == b.batch || b.batch == 1, "batch mismatch");
        if (need_scale)
        {
            EV_REQUIRE(are_close(a.scale, b.scale), "scale mismatch");
        }
    }

    // ---------------------------------------------------------------- add / sub / negate
    Ct Evaluator::add(const Ct &a, const Ct &b) const
    {
        check_same(a, b, true);
        EV_REQUIRE(a.size == b.size, "size mismatch"); // (SEAL pads the smaller one; MOAI never mixes sizes)
        Ct r = alloc(a.batch, a.size, a.limbs, a.scale);
        ew_addsub(c, EW_ADD, a.d, b.d, r.d, a.batch, a.size, a.limbs, b.batch == 1 && a.batch != 1);
        return r;
    }

    Ct Evaluator::sub(const Ct &a, const Ct &b) const
    {
        check_same(a, b, true);
        EV_REQUIRE(a.size == b.size, "size mismatch");
        Ct r = alloc(a.batch, a.size, a.limbs, a.scale);
        ew_addsub(c, EW_SUB, a.d, b.d, r.d, a.batch, a.size, a.limbs, b.batch == 1 && a.batch != 1);
        return r;
    }

    void Evaluator::add_inplace(Ct &a, const Ct &b) const
    {
        check_same(a, b, true);
        EV_REQUIRE(a.size == b.size, "size mismatch");
        ew_addsub(c, EW_ADD, a.d, b.d, a.d, a.batch, a.size, a.limbs, b.batch == 1 && a.batch != 1);
    }

    void Evaluator::sub_inplace(Ct &a, const Ct &b) const
    {
        check_same(a, b, true);
        EV_REQUIRE(a.size == b.size, "size mismatch");
        ew_addsub(c, EW_SUB, a.d, b.d, a.d, a.batch, a.size, a.limbs, b.batch == 1 && a.batch != 1);
    }

    Ct Evaluator::negate(const Ct &a) const
    {
        Ct r = alloc(a.batch, a.size, a.limbs, a.scale);
        ew_addsub(c, EW_NEG, a.d, a.d, r.d, a.batch, a.size, a.limbs, false);
        return r;
    }

    void Evaluator::double_inplace(Ct &a) const
    {
        // S/evaluator.h:1344: add_inplace(encrypted, encrypted)
        ew_addsub(c, EW_ADD, a.d, a.d, a.d, a.batch, a.size, a.limbs, false);
    }

    void Evaluator::double_add_const_inplace(Ct &a, double value) const
    {
        const Pt p = encode(value, a.limbs, a.scale);
        ew_double_add_scalar(c, a.d, p.consts.data(), a.d, a.batch, a.size, a.limbs);
    }

    void Evaluator::add_into3(Ct &acc3, const Ct &x2) const
    {
        check_same(acc3, x2, true);
        EV_REQUIRE(acc3.size == 3 && x2.size == 2 && acc3.batch == x2.batch, "add_into3 expects sizes 3 and 2");
        ew_add_into3(c, acc3.d, x2.d, acc3.batch, acc3.limbs);
    }

    Ct Evaluator::sum_batch(const Ct &a) const
    {
        Ct r = alloc(1, a.size, a.limbs, a.scale);
        moai::sum_batch(c, a.d, r.d, a.batch, a.size, a.limbs);
        return r;
    }

    Ct Evaluator::inner_product(const Ct &a, const Ct &b) const
    {
        check_same(a, b, false);
        EV_REQUIRE(a.size == 2 && b.size == 2 && a.batch == b.batch, "inner_product expects equal batches of size-2");
        Ct r = alloc(1, 3, a.limbs, a.scale * b.scale);
        moai::inner_product(c, a.d, b.d, r.d, a.batch, a.limbs, 0);
        return r;
    }

    Ct Evaluator::sum_sub_square(const Ct &a, const Ct &m) const
    {
        check_same(a, m, true);
        EV_REQUIRE(a.size == 2 && m.size == 2 && m.batch == 1, "sum_sub_square expects size-2 and a single subtrahend");
        Ct r = alloc(1, 3, a.limbs, a.scale * a.scale);
        moai::inner_product(c, a.d, m.d, r.d, a.batch, a.limbs, 1);
        return r;
    }

    // ---------------------------------------------------------------- plaintext ops
    Ct Evaluator::add_plain(const Ct &a, const Pt &p) const
    {
        // S/evaluator.cpp:1938-2044
        EV_REQUIRE(a.limbs == p.limbs, "encrypted and plain parameter mismatch");
        EV_REQUIRE(are_close(a.scale, p.scale), "scale mismatch");
        Ct r = alloc(a.batch, a.size, a.limbs, a.scale);
        if (p.is_scalar)
        {
            ew_add_scalar(c, a.d, p.consts.data(), r.d, a.batch, a.size, a.limbs);
        }
        else
        {
            EV_REQUIRE(p.count == 1 || p.count == a.batch, "plaintext batch mismatch");
            ew_addsub_plain(c, EW_ADD, a.d, p.d, r.d, a.batch, a.size, a.limbs,
                            p.count == 1 ? 0 : (long long)a.limbs * (long long)n());
        }
        return r;
    }

    void Evaluator::add_plain_inplace(Ct &a, const Pt &p) const
    {
        EV_REQUIRE(a.limbs == p.limbs, "encrypted and plain parameter mismatch");
        EV_REQUIRE(are_close(a.scale, p.scale), "scale mismatch");
        if (p.is_scalar)
        {
            ew_add_scalar(c, a.d, p.consts.data(), a.d, a.batch, a.size, a.limbs);
        }
        else
        {
            EV_REQUIRE(p.count == 1 || p.count == a.batch, "plaintext batch mismatch");
            ew_addsub_plain(c, EW_ADD, a.d, p.d, a.d, a.batch, a.size, a.limbs,
                            p.count == 1 ? 0 : (long long)a.limbs * (long long)n());
        }
    }

    Ct Evaluator::sub_plain(const Ct &a, const Pt &p) const
    {
        EV_REQUIRE(a.limbs == p.limbs, "encrypted and plain parameter mismatch");
        EV_REQUIRE(are_close(a.scale, p.scale), "scale mismatch");
        Ct r = alloc(a.batch, a.size, a.limbs, a.scale);
        if (p.is_scalar)
        {
            std::vector<u64> neg(p.limbs);
            for (int l = 0; l < p.limbs; l++)
            {
                neg[l] = p.consts[l] ? c->q[l] - p.consts[l] : 0;
            }
            ew_add_scalar(c, a.d, neg.data(), r.d, a.batch, a.size, a.limbs);
        }
        else
        {
            EV_REQUIRE(p.count == 1 || p.count == a.batch, "plaintext batch mismatch");
            ew_addsub_plain(c, EW_SUB, a.d, p.d, r.d, a.batch, a.size, a.limbs,
                            p.count == 1 ? 0 : (long long)a.limbs * (long long)n());
        }
        return r;
    }

    Ct Evaluator::multiply_plain(const Ct &a, const Pt &p) const
    {
        // S/evaluator.cpp:2154-2198, 2336-2373: new scale = ct.scale * plain.scale
        EV_REQUIRE(a.limbs == p.limbs, "encrypted and plain parameter mismatch");
        Ct r = alloc(a.batch, a.size, a.limbs, a.scale * p.scale);
        if (p.is_scalar)
        {
            ew_multiply_scalar(c, a.d, p.consts.data(), r.d, a.batch, a.size, a.limbs);
        }
        else
        {
            EV_REQUIRE(p.count == 1 || p.count == a.batch, "plaintext batch mismatch");
            ew_multiply_plain(c, a.d, p.d, r.d, a.batch, a.size, a.limbs,
                              p.count == 1 ? 0 : (long long)a.limbs * (long long)n());
        }
        return r;
    }

    Ct Evaluator::lincomb_scalar(const std::vector<Ct> &terms, const std::vector<double> &coefs, int limbs,
                                 double out_scale) const
    {
        EV_REQUIRE(!terms.empty() && terms.size() == coefs.size() && terms.size() <= 8, "1..8 terms expected");
        std::vector<const u64 *> in;
        std::vector<int> in_limbs;
        std::vector<u64> consts;
        for (size_t j = 0; j < terms.size(); j++)
        {
            const Ct &t = terms[j];
            EV_REQUIRE(t.size == terms[0].size && t.batch == terms[0].batch, "terms must share batch and size");
            EV_REQUIRE(t.limbs >= limbs, "a term is below the requested level");
            const Pt k = encode(coefs[j], limbs, out_scale / t.scale); // same constants multiply_plain would use
            in.push_back(t.d);
            in_limbs.push_back(t.limbs);
            consts.insert(consts.end(), k.consts.begin(), k.consts.end());
        }
        Ct r = alloc(terms[0].batch, terms[0].size, limbs, out_scale);
        ew_lincomb_scalar(c, (int)terms.size(), in.data(), in_limbs.data(), consts.data(), r.d, r.batch, r.size, limbs);
        return r;
    }

    // ---------------------------------------------------------------- ct x ct
    Ct Evaluator::multiply(const Ct &a, const Ct &b) const
    {
        // S/evaluator.cpp:770-909
        check_same(a, b, false);
        EV_REQUIRE(a.size == 2 && b.size == 2, "multiply expects size-2 ciphertexts");
        Ct r = alloc(a.batch, 3, a.limbs, a.scale * b.scale);
        ew_multiply(c, a.d, b.d, r.d, a.batch, a.limbs, false, b.batch == 1 && a.batch != 1);
        return r;
    }

    Ct Evaluator::multiply_lowered(const Ct &a, const Ct &b) const
    {
        EV_REQUIRE(a.size == 2 && b.size == 2, "multiply expects size-2 ciphertexts");
        EV_REQUIRE(a.batch == b.batch || b.batch == 1, "batch mismatch");
        const int limbs = std::min(a.limbs, b.limbs);
        Ct r = alloc(a.batch, 3, limbs, a.scale * b.scale);
        ew_multiply(c, a.d, b.d, r.d, a.batch, limbs, false, b.batch == 1 && a.batch != 1, a.limbs, b.limbs);
        return r;
    }

    Ct Evaluator::square(const Ct &a) const
    {
        EV_REQUIRE(a.size == 2, "square expects a size-2 ciphertext");
        Ct r = alloc(a.batch, 3, a.limbs, a.scale * a.scale);
        ew_square(c, a.d, r.d, a.batch, a.limbs);
        return r;
    }

    void Evaluator::multiply_accumulate(Ct &acc3, const Ct &a, const Ct &b) const
    {
        check_same(a, b, false);
        EV_REQUIRE(acc3.size == 3 && acc3.limbs == a.limbs && acc3.batch == a.batch, "accumulator shape mismatch");
        EV_REQUIRE(are_close(acc3.scale, a.scale * b.scale), "scale mismatch");
        ew_multiply(c, a.d, b.d, acc3.d, a.batch, a.limbs, true, b.batch == 1 && a.batch != 1);
    }

    Ct Evaluator::relinearize(const Ct &a3, const Keys &k) const
    {
        // S/evaluator.cpp:1345-1400
        EV_REQUIRE(a3.size == 3, "relinearize expects a size-3 ciphertext");
        const KeyRef *rk = k.relin_at(c, a3.limbs);
        EV_REQUIRE(rk != nullptr, "not enough relinearization keys");
        Ct r = alloc(a3.batch, 2, a3.limbs, a3.scale);
        moai::relinearize(c, a3.d, r.d, a3.batch, a3.limbs, rk->p, rk->key_kl, rk->k_extra);
        return r;
    }

    static bool merge_rescale_enabled()
    {
        static const bool on = [] {
            const char *e = std::getenv("MOAI_MERGE_RESCALE");
            return !(e && e[0] == '0');
        }();
        return on;
    }

    Ct Evaluator::relin_rescale(const Ct &a3, const Keys &k) const
    {
        // rescale_to_next(relinearize(a3)); with a grouped-digit key (fast mode) the two divisions are one
        // (ksg_moddown_rescale, csrc/ksgroup.cu): same plaintext, one rounding instead of two
        EV_REQUIRE(a3.size == 3, "relinearize expects a size-3 ciphertext");
        const KeyRef *rk = k.relin_at(c, a3.limbs);
        EV_REQUIRE(rk != nullptr, "not enough relinearization keys");
        if (rk->k_extra <= 0 || a3.limbs < 2 || !merge_rescale_enabled())
        {
            return rescale_to_next(relinearize(a3, k));
        }
        Ct r = alloc(a3.batch, 2, a3.limbs - 1, a3.scale / last_prime(a3.limbs));
        moai::relinearize_rescale(c, a3.d, r.d, a3.batch, a3.limbs, rk->p, rk->key_kl, rk->k_extra);
        return r;
    }

    Ct Evaluator::rescale_to_next(const Ct &a) const
    {
        // S/evaluator.cpp:1402-1481, 1682-1720: scale /= q_last
        EV_REQUIRE(a.limbs >= 2, "end of modulus switching chain reached");
        Ct r = alloc(a.batch, a.size, a.limbs - 1, a.scale / last_prime(a.limbs));
        rescale(c, a.d, r.d, a.batch, a.size, a.limbs);
        return r;
    }

    Ct Evaluator::mod_switch_to(const Ct &a, int limbs) const
    {
        // S/evaluator.cpp:1583-1652: cannot switch to a higher level
        EV_REQUIRE(limbs >= 1 && limbs <= a.limbs, "cannot switch to higher level modulus");
        if (limbs == a.limbs)
        {
            return a;
        }
        Ct r = alloc(a.batch, a.size, limbs, a.scale);
        mod_switch_drop(c, a.d, r.d, a.batch, a.size, a.limbs, limbs);
        return r;
    }

    Ct Evaluator::rotate_vector(const Ct &a, int steps, const Keys &k) const
    {
        // S/evaluator.cpp:2667-2722
        EV_REQUIRE(a.size == 2, "encrypted size must be 2");
        if (steps == 0)
        {
            return a;
        }
        uint32_t elt = c->elt_from_step(steps);
        auto it = k.galois.find(elt);
        if (it != k.galois.end())
        {
            EV_REQUIRE(it->second.max_limbs() >= a.limbs, "Galois key was truncated below this level");
            Ct r = alloc(a.batch, 2, a.limbs, a.scale);
            apply_galois(c, a.d, r.d, a.batch, a.limbs, elt, it->second.p, it->second.key_kl);
            return r;
        }
        if (const KeyRef *fk = k.fast(c, elt, a.limbs))
        {
            return rotate_fast(a, elt, *fk);
        }
        // NAF decomposition (S/util/numth.h:22-42)
        std::vector<int> naf;
        {
            const bool neg = steps < 0;
            long long v = neg ? -(long long)steps : steps;
            for (int i = 0; v; i++)
            {
                int zi = (v & 1) ? 2 - (int)(v & 3) : 0;
                v = (v - zi) >> 1;
                if (zi)
                {
                    naf.push_back((neg ? -zi : zi) * (1 << i));
                }
            }
        }
        EV_REQUIRE(naf.size() != 1, "Galois key not present");
        Ct cur = a;
        for (int s : naf)
        {
            if ((size_t)std::abs(s) != (n() >> 1))
            {
                cur = rotate_vector(cur, s, k);
            }
        }
        return cur;
    }

    Ct Evaluator::complex_conjugate(const Ct &a, const Keys &k) const
    {
        uint32_t elt = c->elt_from_step(0);
        auto it = k.galois.find(elt);
        if (it == k.galois.end())
        {
            const KeyRef *fk = k.fast(c, elt, a.limbs);
            EV_REQUIRE(fk != nullptr, "Galois key not present");
            return rotate_fast(a, elt, *fk);
        }
        Ct r = alloc(a.batch, 2, a.limbs, a.scale);
        apply_galois(c, a.d, r.d, a.batch, a.limbs, elt, it->second.p, it->second.key_kl);
        return r;
    }

    // ---------------------------------------------------------------- hoisted rotations (fast mode)
    bool Evaluator::has_fast_key(int steps, int limbs, const Keys &k) const
    {
        return steps == 0 || k.fast(c, c->elt_from_step(steps), limbs) != nullptr;
    }

    Hoisted Evaluator::hoist(const Ct &a, int k_extra) const
    {
        EV_REQUIRE(a.size == 2, "encrypted size must be 2");
        Hoisted h;
        h.src = a;
        h.k_extra = k_extra;
        if (k_extra > 0)
        {
            h.ext = std::make_shared<DevBuf>((size_t)a.batch * ksg_ext_bytes_per_ct(c, a.limbs, k_extra), c->stream);
            ksg_decompose(c, a.d + (size_t)a.limbs * n(), a.batch, a.limbs, k_extra, reinterpret_cast<u64 *>(h.ext->p),
                          2LL * a.limbs * (long long)n(), 3);
            return h;
        }
        h.ext = std::make_shared<DevBuf>((size_t)a.batch * ks_ext_bytes_per_ct(c, a.limbs), c->stream);
        ks_decompose(c, a.d + (size_t)a.limbs * n(), a.batch, a.limbs, reinterpret_cast<u64 *>(h.ext->p),
                     2LL * a.limbs * (long long)n());
        return h;
    }

    Ct Evaluator::rotate_hoisted(const Hoisted &h, int steps, const Keys &k) const
    {
        const Ct &a = h.src;
        if (steps == 0)
        {
            return a;
        }
        const uint32_t elt = c->elt_from_step(steps);
        const KeyRef *fk = k.fast(c, elt, a.limbs, h.k_extra);
        EV_REQUIRE(fk != nullptr, "pre-permuted Galois key not present");
        Ct r = alloc(a.batch, 2, a.limbs, a.scale);
        if (h.k_extra > 0)
        {
            u64 *outp = r.d;
            ksg_rotate_hoisted_multi(c, a.d, reinterpret_cast<const u64 *>(h.ext->p), a.batch, a.limbs, h.k_extra, 1, &elt,
                                     &fk->p, &fk->key_kl, &outp);
            return r;
        }
        moai::rotate_hoisted(c, a.d, reinterpret_cast<const u64 *>(h.ext->p), a.batch, a.limbs, elt, fk->p, fk->key_kl,
                             r.d);
        return r;
    }

    Ct Evaluator::rotate_fast(const Ct &a, uint32_t elt, const KeyRef &key) const
    {
        Ct r = alloc(a.batch, 2, a.limbs, a.scale);
        rotate_prepermuted(c, a.d, a.batch, a.limbs, elt, key.p, key.key_kl, r.d, key.k_extra);
        return r;
    }

    std::vector<Ct> Evaluator::rotate_many(const Ct &a, const std::vector<int> &steps, const Keys &k) const
    {
        std::vector<Ct> out(steps.size());
        bool all_fast = true;
        int nonzero = 0;
        for (int s : steps)
        {
            all_fast = all_fast && has_fast_key(s, a.limbs, k);
            nonzero += s != 0;
        }
        if (!all_fast || nonzero < 2)
        {
            for (size_t i = 0; i < steps.size(); i++)
            {
                out[i] = rotate_vector(a, steps[i], k);
            }
            return out;
        }
        // all rotations share one decomposition, hence one digit layout: the cheapest number of extra primes for
        // which EVERY step has a key (0 = SEAL's per-prime digits)
        int kx = -1;
        {
            double best_cost = 0;
            for (int cand = 0; a.limbs + cand <= c->kl - 1; cand++)
            {
                bool ok = true;
                for (int s : steps)
                {
                    ok = ok && (s == 0 || k.fast(c, c->elt_from_step(s), a.limbs, cand) != nullptr);
                }
                const double cost = ok ? ksg_cost(c, a.limbs, cand) : 0;
                if (ok && (kx < 0 || cost < best_cost))
                {
                    kx = cand;
                    best_cost = cost;
                }
            }
        }
        if (kx < 0) // keys exist for every step, but with different digit layouts
        {
            for (size_t i = 0; i < steps.size(); i++)
            {
                out[i] = rotate_vector(a, steps[i], k);
            }
            return out;
        }
        for (size_t i = 0; i < steps.size(); i++)
        {
            out[i] = steps[i] == 0 ? a : alloc(a.batch, 2, a.limbs, a.scale);
        }
        // the decomposition of a chunk is shared by all rotations; ~4 GiB of extended digits at a time
        long long chunk = ks_chunk(c, a.limbs, a.batch, ks_ext_budget());
        if (kx > 0)
        {
            chunk = std::max<long long>(1, std::min<long long>(a.batch, (long long)(ks_ext_budget() / ksg_ext_bytes_per_ct(c, a.limbs, kx))));
        }
        for (long long b0 = 0; b0 < a.batch; b0 += chunk)
        {
            const long long nb = std::min(chunk, a.batch - b0);
            Hoisted h = hoist(view(a, b0, nb), kx);
            const u64 *extp = reinterpret_cast<const u64 *>(h.ext->p);
            std::vector<size_t> todo;
            for (size_t i = 0; i < steps.size(); i++)
            {
                if (steps[i] != 0)
                {
                    todo.push_back(i);
                }
            }
            size_t g0 = 0;
            if (kx > 0 || ks_multi_enabled(c, a.limbs))
            {
                // up to KSM_R rotations share one pass over the extended digits
                for (; g0 < todo.size(); g0 += KSM_R)
                {
                    const int cnt = (int)std::min<size_t>(KSM_R, todo.size() - g0);
                    uint32_t elts[KSM_R];
                    const u64 *kp[KSM_R];
                    int kkl[KSM_R];
                    u64 *outs[KSM_R];
                    for (int r = 0; r < cnt; r++)
                    {
                        const size_t i = todo[g0 + r];
                        elts[r] = c->elt_from_step(steps[i]);
                        const KeyRef *fk = k.fast(c, elts[r], a.limbs, kx);
                        kp[r] = fk->p;
                        kkl[r] = fk->key_kl;
                        outs[r] = out[i].d + (size_t)b0 * 2 * a.limbs * n();
                    }
                    if (kx > 0)
                    {
                        ksg_rotate_hoisted_multi(c, h.src.d, extp, nb, a.limbs, kx, cnt, elts, kp, kkl, outs);
                    }
                    else
                    {
                        rotate_hoisted_multi(c, h.src.d, extp, nb, a.limbs, cnt, elts, kp, kkl, outs);
                    }
                }
            }
            for (; g0 < todo.size(); g0++)
            {
                const size_t i = todo[g0];
                const uint32_t elt = c->elt_from_step(steps[i]);
                const KeyRef *fk = k.fast(c, elt, a.limbs, kx);
                moai::rotate_hoisted(c, h.src.d, extp, nb, a.limbs, elt, fk->p, fk->key_kl,
                                     out[i].d + (size_t)b0 * 2 * a.limbs * n());
            }
        }
        return out;
    }

    // ---------------------------------------------------------------- encoder
    Pt Evaluator::encode(double value, int limbs, double scale) const
    {
        // S/ckks.cpp:77-216
        EV_REQUIRE(limbs >= 1 && limbs <= c->kl, "parms_id is not valid for encryption parameters");
        EV_REQUIRE(scale > 0, "scale out of bounds");
        Pt p;
        p.is_scalar = true;
        p.limbs = limbs;
        p.scale = scale;
        p.count = 1;
        p.consts.resize(limbs);
        double v = value * scale;
        int bit_count = (int)(std::log2(std::fabs(v))) + 2;
        double r = std::round(v);
        const bool neg = std::signbit(r);
        r = std::fabs(r);
        EV_REQUIRE(bit_count <= 128, "encoded value is too large");
        unsigned __int128 mag;
        if (bit_count <= 64)
        {
            mag = (u64)r;
        }
        else
        {
            const double two64 = std::pow(2.0, 64);
            mag = (((unsigned __int128)(u64)(r / two64)) << 64) | (u64)std::fmod(r, two64);
        }
        for (int l = 0; l < limbs; l++)
        {
            u64 res = (u64)(mag % c->q[l]);
            p.consts[l] = neg ? (res ? c->q[l] - res : 0) : res;
        }
        return p;
    }

    Pt Evaluator::encode_batch(const std::complex<double> *values, long long count, int n_vals, int limbs,
                               double scale) const
    {
        EV_REQUIRE(limbs >= 1 && limbs <= c->kl, "parms_id is not valid for encryption parameters");
        Pt p;
        p.limbs = limbs;
        p.scale = scale;
        p.count = count;
        p.buf = std::make_shared<DevBuf>((size_t)count * limbs * n() * sizeof(u64), c->stream);
        p.d = reinterpret_cast<u64 *>(p.buf->p);
        DevBuf dv((size_t)count * n_vals * sizeof(double) * 2, c->stream);
        MOAI_CUDA_CHECK(cudaMemcpyAsync(dv.p, values, (size_t)count * n_vals * sizeof(double) * 2,
                                        cudaMemcpyHostToDevice, c->stream));
        encode_vector(c, reinterpret_cast<const double *>(dv.p), count, n_vals, scale, limbs, p.d);
        return p;
    }

    Pt Evaluator::encode(const std::vector<std::complex<double>> &values, int limbs, double scale) const
    {
        return encode_batch(values.data(), 1, (int)values.size(), limbs, scale);
    }

    Pt Evaluator::encode(const std::vector<double> &values, int limbs, double scale) const
    {
        std::vector<std::complex<double>> v(values.begin(), values.end());
        return encode_batch(v.data(), 1, (int)v.size(), limbs, scale);
    }

    // ---------------------------------------------------------------- fork-only ops
    Ct Evaluator::multiply_const(const Ct &a, double value) const
    {
        // S/evaluator.cpp:402-409: encode(value, encrypted.scale()) then multiply_plain
        return multiply_plain(a, encode(value, a.limbs, a.scale));
    }

    Ct Evaluator::add_const(const Ct &a, double value) const
    {
        return add_plain(a, encode(value, a.limbs, a.scale));
    }

    Ct Evaluator::multiply_vector_reduced_error(const Ct &a, const std::vector<std::complex<double>> &v) const
    {
        // S/evaluator.h:1371-1386: encode(value, encrypted.scale()), mod-switch the plain, multiply_plain
        return multiply_plain(a, encode(v, a.limbs, a.scale));
    }

    // b' = rescale(hi * encode(s_lo * q_last(hi) / s_hi^2)) with its scale overwritten, switched down
    // to lo's level (the common prefix of S/evaluator.cpp:430-452, 488-510, 547-569)
    Ct Evaluator::reduced_error_adjust(const Ct &hi, const Ct &lo) const
    {
        const double ql = last_prime(hi.limbs);
        const double scale_adjust = lo.scale * ql / (hi.scale * hi.scale);
        Ct adj = multiply_const(hi, scale_adjust);
        adj.scale = lo.scale * ql;
        adj = rescale_to_next(adj);
        return mod_switch_to(adj, lo.limbs);
    }

    Ct Evaluator::add_reduced_error(const Ct &a, const Ct &b) const
    {
        // S/evaluator.cpp:420-476 (a = encrypted1, b = encrypted2)
        if (a.limbs == b.limbs)
        {
            Ct x = a;
            x.scale = b.scale;
            return add(x, b);
        }
        if (a.limbs < b.limbs)
        {
            Ct adj = reduced_error_adjust(b, a);
            Ct x = a;
            x.scale = adj.scale;
            return add(x, adj);
        }
        Ct adj = reduced_error_adjust(a, b);
        adj.scale = b.scale;
        return add(adj, b);
    }

    Ct Evaluator::sub_reduced_error(const Ct &a, const Ct &b) const
    {
        // S/evaluator.cpp:478-534
        if (a.limbs == b.limbs)
        {
            Ct x = a;
            x.scale = b.scale;
            return sub(x, b);
        }
        if (a.limbs < b.limbs)
        {
            Ct adj = reduced_error_adjust(b, a);
            Ct x = a;
            x.scale = adj.scale;
            return sub(x, adj);
        }
        Ct adj = reduced_error_adjust(a, b);
        adj.scale = b.scale;
        return sub(adj, b);
    }

    Ct Evaluator::multiply_reduced_error(const Ct &a, const Ct &b, const Keys &k) const
    {
        // S/evaluator.cpp:536-594
        if (a.limbs == b.limbs)
        {
            Ct x = a;
            x.scale = b.scale;
            return relinearize(multiply(x, b), k);
        }
        if (a.limbs < b.limbs)
        {
            Ct adj = reduced_error_adjust(b, a);
            Ct x = a;
            x.scale = adj.scale;
            return relinearize(multiply(x, adj), k);
        }
        Ct adj = reduced_error_adjust(a, b);
        adj.scale = b.scale;
        return relinearize(multiply(adj, b), k);
    }

    Ct Evaluator::mod_raise(const Ct &a, int limbs_out) const
    {
        EV_REQUIRE(a.limbs == 1, "mod_raise expects a ciphertext at the last level");
        Ct r = alloc(a.batch, a.size, limbs_out, a.scale);
        moai::mod_raise(c, a.d, r.d, a.batch, a.size, limbs_out);
        return r;
    }
} // namespace moai
