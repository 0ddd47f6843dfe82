// Bootstrapper.h for the facade include path: the class M/source/bootstrapping/Bootstrapper.h:15-221
// declares, with the constructor signature and the call sequence of the reference's driver
// (M/test/test_full_scheme.hpp:413-448, 654-660; M/source/non_linear_func/softmax.hpp:536):
//
//     Bootstrapper b(loge, logn, logNh, L, final_scale, boundary_K, deg, scale_factor, inverse_deg,
//                    context, keygen, encoder, encryptor, decryptor, evaluator, relin_keys, gal_keys);
//     b.prepare_mod_polynomial();
//     b.addLeftRotKeys_Linear_to_vector_3(gal_steps_vector);   // the client generates these keys
//     b.slot_vec.push_back(logn);
//     b.generate_LT_coefficient_3();
//     b.bootstrap_3(rtn, ct);
//
// bound to libmoai_b200.so's bootstrapper (moai_bootstrapper_create / moai_bootstrap,
// include/moai_b200_modules.h): same level budget (CoeffToSlot 3, cosine 6, double angles 2,
// SlotToCoeff 3; output L - 14 data limbs at final_scale) but our own plan — decrypted tolerance,
// not SEAL's residues.  Only full-slot bootstrapping (logn == logNh), the one
// MOAI uses, is provided.  keygen / encryptor are client-side objects the reference's class stores but
// never uses on this path; they are accepted as any type and ignored.
#pragma once
#include "seal/seal.h"
#include <chrono>
#include <condition_variable>
#include <exception>
#include <mutex>
#include <vector>

class Bootstrapper
{
    struct Request
    {
        const seal::Ciphertext *in = nullptr;
        seal::Ciphertext *out = nullptr;
        bool done = false;
        std::exception_ptr error;
    };

public:
    long loge, logn, n, logNh, Nh, L;
    double initial_scale = 0.0, final_scale;
    long boundary_K, sin_cos_deg, scale_factor, inverse_deg;
    std::vector<long> slot_vec;
    long slot_index = 0;

    template <class KeyGeneratorT, class EncryptorT>
    Bootstrapper(long _loge, long _logn, long _logNh, long _L, double _final_scale, long _boundary_K, long _sin_cos_deg,
                 long _scale_factor, long _inverse_deg, seal::SEALContext &_context, KeyGeneratorT &, seal::CKKSEncoder &,
                 EncryptorT &, seal::Decryptor &, seal::Evaluator &, seal::RelinKeys &_relin_keys,
                 seal::GaloisKeys &_gal_keys)
        : loge(_loge), logn(_logn), n(1L << _logn), logNh(_logNh), Nh(1L << _logNh), L(_L), final_scale(_final_scale),
          boundary_K(_boundary_K), sin_cos_deg(_sin_cos_deg), scale_factor(_scale_factor), inverse_deg(_inverse_deg),
          context(_context), relin_keys(_relin_keys), gal_keys(_gal_keys)
    {
        if (logn != logNh)
        {
            throw std::invalid_argument("the B200 backend provides full-slot bootstrapping (logn == logNh) only");
        }
    }
    // server-side construction without the client-side objects
    Bootstrapper(long _loge, long _logn, long _logNh, long _L, double _final_scale, long _boundary_K, long _sin_cos_deg,
                 long _scale_factor, long _inverse_deg, seal::SEALContext &_context, seal::RelinKeys &_relin_keys,
                 seal::GaloisKeys &_gal_keys)
        : loge(_loge), logn(_logn), n(1L << _logn), logNh(_logNh), Nh(1L << _logNh), L(_L), final_scale(_final_scale),
          boundary_K(_boundary_K), sin_cos_deg(_sin_cos_deg), scale_factor(_scale_factor), inverse_deg(_inverse_deg),
          context(_context), relin_keys(_relin_keys), gal_keys(_gal_keys)
    {
        if (logn != logNh)
        {
            throw std::invalid_argument("the B200 backend provides full-slot bootstrapping (logn == logNh) only");
        }
    }
    Bootstrapper(const Bootstrapper &) = delete;
    Bootstrapper &operator=(const Bootstrapper &) = delete;
    ~Bootstrapper()
    {
        seal::detail::RootLock lk(context.impl()->mu);
        if (keys_)
        {
            moai_keys_destroy(keys_);
        }
        if (h_)
        {
            moai_bootstrapper_destroy(h_);
        }
    }

    void set_final_scale(double s)
    {
        if (h_ && s != final_scale)
        {
            throw std::logic_error("set_final_scale must precede generate_LT_coefficient_3");
        }
        final_scale = s;
    }
    // fast mode: plan the linear transforms for hoisted baby steps (call before asking for the steps)
    void set_hoisting(bool on)
    {
        hoisting_ = on;
        if (h_)
        {
            seal::detail::RootLock lk(context.impl()->mu);
            seal::detail::chk(moai_bootstrapper_set_hoisting(h_, on ? 1 : 0));
        }
    }

    // Bootstrapper::prepare_mod_polynomial (Bootstrapper.cpp:3-26 + ModularReducer): the EvalMod
    // polynomial is fitted when the device-side plan is created
    void prepare_mod_polynomial()
    {
        create();
    }
    // Bootstrapper.cpp:89-185: appends the rotation steps whose Galois keys the linear transforms use
    void addLeftRotKeys_Linear_to_vector_3(std::vector<int> &gal_steps_vector)
    {
        create();
        std::vector<std::int32_t> steps(4096);
        std::int32_t count = 0;
        {
            seal::detail::RootLock lk(context.impl()->mu);
            seal::detail::chk(moai_bootstrapper_required_steps(h_, steps.data(), static_cast<std::int32_t>(steps.size()), &count));
        }
        for (std::int32_t i = 0; i < count; i++)
        {
            if (std::find(gal_steps_vector.begin(), gal_steps_vector.end(), int(steps[i])) == gal_steps_vector.end())
            {
                gal_steps_vector.push_back(int(steps[i]));
            }
        }
    }
    // the pre-rotated diagonals are encoded on first use inside the library
    void generate_LT_coefficient_3()
    {
        create();
    }

    // Opt-in request combining.  The reference's driver bootstraps 768 ciphertexts with one bootstrap_3 call each inside
    // `#pragma omp parallel for` (M/test/test_full_scheme.hpp:654-660): with combining on, calls that arrive
    // concurrently from different threads are collected by the first caller (the "leader") and go to the device as ONE
    // batched call — and, for real-slot messages (all of MOAI's activations), two ciphertexts per bootstrapping
    // (moai_bootstrap_real) — while the other callers wait for their result.  Each caller still gets exactly its own
    // ciphertext back; nothing changes for a single-threaded caller except `linger_us` of latency.
    //   real_slots: the messages are real (required for the two-per-bootstrapping packing); false = batching only
    //   max_batch : most requests per device call;  linger_us: how long the leader waits for more requests
    void set_combining(bool on, bool real_slots = true, int max_batch = 64, int linger_us = 300)
    {
        std::lock_guard<std::mutex> lk(qmu_);
        combining_ = on;
        combine_real_ = real_slots;
        combine_max_ = max_batch < 1 ? 1 : max_batch;
        combine_linger_us_ = linger_us < 0 ? 0 : linger_us;
    }
    // device calls made on behalf of combined requests so far (for tests / tuning)
    std::size_t combined_device_calls() const
    {
        return combined_calls_;
    }

    // Bootstrapper.cpp:3496-3502 (+ modraise_inplace :2938-2945 for the argument checks)
    void bootstrap_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher)
    {
        if (cipher.size() != 2)
        {
            throw std::invalid_argument("Ciphertexts of size 2 are supported only!");
        }
        if (cipher.coeff_modulus_size() != 1)
        {
            throw std::invalid_argument("Ciphertexts in the lowest level are supported only!");
        }
        if (combining_)
        {
            bootstrap_combined(rtncipher, cipher);
            return;
        }
        create();
        bind_keys();
        initial_scale = cipher.scale();
        const std::size_t out_limbs = static_cast<std::size_t>(L + 1 - 14);
        seal::Ciphertext out(context, context.parms_id_for_limbs(out_limbs), 2);
        std::int32_t got_limbs = 0;
        double got_scale = 0.0;
        {
            seal::detail::RootLock lk(context.impl()->mu);
            seal::detail::chk(moai_bootstrap(context.handle(), h_, keys_, cipher.data(), 1, cipher.scale(), out.data(),
                                             &got_limbs, &got_scale));
        }
        if (static_cast<std::size_t>(got_limbs) != out_limbs)
        {
            throw std::logic_error("unexpected level after bootstrapping");
        }
        out.scale() = got_scale;
        out.is_ntt_form() = true;
        rtncipher = std::move(out);
    }
    // explicit batch form (not in the reference): all ciphertexts in one device call, two real-slot ciphertexts per
    // bootstrapping when real_slots is true — what the driver's 768-iteration loop amounts to
    void bootstrap_3(std::vector<seal::Ciphertext> &rtnciphers, const std::vector<seal::Ciphertext> &ciphers,
                     bool real_slots = true)
    {
        std::vector<Request> rq(ciphers.size());
        std::vector<Request *> batch;
        rtnciphers.resize(ciphers.size());
        for (std::size_t i = 0; i < ciphers.size(); i++)
        {
            if (ciphers[i].size() != 2)
            {
                throw std::invalid_argument("Ciphertexts of size 2 are supported only!");
            }
            if (ciphers[i].coeff_modulus_size() != 1)
            {
                throw std::invalid_argument("Ciphertexts in the lowest level are supported only!");
            }
            rq[i].in = &ciphers[i];
            rq[i].out = &rtnciphers[i];
            batch.push_back(&rq[i]);
        }
        if (batch.empty())
        {
            return;
        }
        const bool saved = combine_real_;
        combine_real_ = real_slots;
        serve(batch);
        combine_real_ = saved;
        if (rq[0].error)
        {
            std::rethrow_exception(rq[0].error);
        }
    }
    void bootstrap_inplace_3(seal::Ciphertext &cipher)
    {
        seal::Ciphertext r;
        bootstrap_3(r, cipher);
        cipher = std::move(r);
    }

    moai_bootstrapper *handle()
    {
        create();
        return h_;
    }
    // the single key set (relinearisation + every registered Galois key) the fused modules run with
    moai_keys *bound_keys()
    {
        bind_keys();
        return keys_;
    }

private:
    void bootstrap_combined(seal::Ciphertext &rtncipher, const seal::Ciphertext &cipher)
    {
        Request rq;
        rq.in = &cipher;
        rq.out = &rtncipher;
        std::unique_lock<std::mutex> lk(qmu_);
        pending_.push_back(&rq);
        if (leader_active_)
        {
            qcv_.wait(lk, [&] { return rq.done; }); // a leader is collecting: it will serve this request
        }
        else
        {
            // first caller: become the leader.  pending_ held nothing else (a leader only leaves with it empty), so
            // this request is served by the first batch below.
            leader_active_ = true;
            while (!pending_.empty())
            {
                const std::size_t want = static_cast<std::size_t>(combine_max_);
                qcv_.wait_for(lk, std::chrono::microseconds(combine_linger_us_), [&] { return pending_.size() >= want; });
                const std::size_t take = std::min(pending_.size(), want);
                std::vector<Request *> batch(pending_.begin(), pending_.begin() + take);
                pending_.erase(pending_.begin(), pending_.begin() + take);
                lk.unlock();
                serve(batch);
                lk.lock();
                for (Request *r : batch)
                {
                    r->done = true;
                }
                qcv_.notify_all();
            }
            leader_active_ = false;
        }
        lk.unlock();
        if (rq.error)
        {
            std::rethrow_exception(rq.error);
        }
    }

    // one device call for a batch of requests that share their scale (others are served one scale at a time)
    void serve(std::vector<Request *> &batch)
    {
        try
        {
            create();
            bind_keys();
            auto &c = context.impl();
            const std::size_t n = c->n, out_limbs = static_cast<std::size_t>(L + 1 - 14);
            std::vector<bool> taken(batch.size(), false);
            for (std::size_t first = 0; first < batch.size(); first++)
            {
                if (taken[first])
                {
                    continue;
                }
                std::vector<std::size_t> group;
                for (std::size_t i = first; i < batch.size(); i++)
                {
                    if (!taken[i] && batch[i]->in->scale() == batch[first]->in->scale())
                    {
                        group.push_back(i);
                        taken[i] = true;
                    }
                }
                const std::size_t B = group.size();
                seal::detail::DeviceBlock in, out;
                in.ensure(c, B * 2 * n);
                out.ensure(c, B * 2 * out_limbs * n);
                std::int32_t got_limbs = 0;
                double got_scale = 0.0;
                {
                    seal::detail::RootLock lk(c->mu);
                    for (std::size_t k = 0; k < B; k++)
                    {
                        seal::detail::chk(moai_memcpy_d2d(c->h, in.ptr() + k * 2 * n, batch[group[k]]->in->data(),
                                                          2 * n * sizeof(std::uint64_t)));
                    }
                    const double scale = batch[first]->in->scale();
                    if (combine_real_)
                    {
                        seal::detail::chk(moai_bootstrap_real(c->h, h_, keys_, in.ptr(), static_cast<std::int64_t>(B), scale, 32,
                                                              out.ptr(), &got_limbs, &got_scale));
                    }
                    else
                    {
                        seal::detail::chk(moai_bootstrap(c->h, h_, keys_, in.ptr(), static_cast<std::int64_t>(B), scale, out.ptr(),
                                                         &got_limbs, &got_scale));
                    }
                    combined_calls_++;
                }
                if (static_cast<std::size_t>(got_limbs) != out_limbs)
                {
                    throw std::logic_error("unexpected level after bootstrapping");
                }
                const seal::parms_id_type &id = context.parms_id_for_limbs(out_limbs);
                for (std::size_t k = 0; k < B; k++)
                {
                    seal::Ciphertext r(context, id, 2);
                    r.scale() = got_scale;
                    r.is_ntt_form() = true;
                    {
                        seal::detail::RootLock lk(c->mu);
                        seal::detail::chk(moai_memcpy_d2d(c->h, r.data(), out.ptr() + k * 2 * out_limbs * n,
                                                          2 * out_limbs * n * sizeof(std::uint64_t)));
                    }
                    *batch[group[k]]->out = std::move(r);
                }
            }
            initial_scale = batch.back()->in->scale();
        }
        catch (...)
        {
            for (Request *r : batch)
            {
                r->error = std::current_exception();
            }
        }
    }

    void create()
    {
        if (h_)
        {
            return;
        }
        seal::detail::RootLock lk(context.impl()->mu);
        seal::detail::chk(moai_bootstrapper_create(context.handle(), static_cast<std::int32_t>(L + 1), final_scale,
                                                   static_cast<std::int32_t>(boundary_K), static_cast<std::int32_t>(sin_cos_deg),
                                                   static_cast<std::int32_t>(scale_factor), static_cast<std::int32_t>(loge), &h_));
        if (hoisting_)
        {
            seal::detail::chk(moai_bootstrapper_set_hoisting(h_, 1));
        }
    }
    // moai_bootstrap takes ONE key set: the relinearisation key and every Galois key registered so far
    void bind_keys()
    {
        auto gs = gal_keys.key_set();
        const std::size_t have = gs ? gs->galois.size() + gs->fast.size() : 0;
        if (keys_ && have == bound_)
        {
            return;
        }
        seal::detail::RootLock lk(context.impl()->mu);
        if (keys_)
        {
            moai_keys_destroy(keys_);
            keys_ = nullptr;
        }
        seal::detail::chk(moai_keys_create(context.handle(), &keys_));
        if (!relin_keys.has_key(2))
        {
            throw std::invalid_argument("not enough relinearization keys");
        }
        seal::detail::chk(moai_keys_set_relin(keys_, relin_keys.device_key()));
        if (gs)
        {
            for (auto &kv : gs->galois)
            {
                seal::detail::chk(moai_keys_add_galois(keys_, kv.first, kv.second));
            }
            for (auto &f : gs->fast)
            {
                seal::detail::chk(moai_keys_add_galois_fast(keys_, f.elt, f.p, f.key_limbs));
            }
        }
        bound_ = have;
    }

    seal::SEALContext &context;
    seal::RelinKeys &relin_keys;
    seal::GaloisKeys &gal_keys;
    moai_bootstrapper *h_ = nullptr;
    moai_keys *keys_ = nullptr;
    std::size_t bound_ = 0;
    bool hoisting_ = false;
    // request combining
    std::mutex qmu_;
    std::condition_variable qcv_;
    std::vector<Request *> pending_;
    bool leader_active_ = false, combining_ = false, combine_real_ = true;
    int combine_max_ = 64, combine_linger_us_ = 300;
    std::size_t combined_calls_ = 0;
};
