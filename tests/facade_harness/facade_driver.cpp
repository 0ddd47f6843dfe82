// facade_driver.cpp — TEST INFRASTRUCTURE.  The reference's module headers, UNMODIFIED, compiled against
// the B200 facade (include/facade/seal/seal.h -> include/moai_b200_seal.hpp) instead of stock SEAL, behind
// the same flat C interface oracle/refbuild/ref_wrap.cpp gives the real library — so one Python test feeds
// identical keys / ciphertexts to both and compares the residues bit for bit.
//
// Built twice by tests/facade_harness/build.py (needs /root/reference for the module headers; outputs in
// oracle/_ref/, git-ignored, travel to the GPU box):
//   libfacade_driver.so       linked against libmoai_b200.so                (the `-m gpu` tests)
//   libfacade_driver_mock.so  linked against the CPU test double mock_cabi.c (host-logic tests, no GPU)
#include "seal/seal.h"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <omp.h>
#include <set>
#include <sstream>
#include <string>
#include <sys/time.h>
#include <vector>

// The reference's own code, all of it: M/include.hpp pulls "seal/seal.h" (the facade), every module header under
// M/source/ and every test / driver header under M/test/ (all_layer_test, SEAL_ckks_test, ...) — unmodified.
// (include path: include/facade [+ include/facade_fused] before /root/reference/include)
#include "include.hpp"

using namespace seal;
using namespace std;

namespace
{
    struct Drv
    {
        unique_ptr<SEALContext> ctx;
        unique_ptr<CKKSEncoder> encoder;
        unique_ptr<Evaluator> evaluator;
        RelinKeys rlk;
        GaloisKeys glk;
        SecretKey sk;
        PublicKey pk;
        EncryptionParameters parms;
        int device = 0;
        unique_ptr<Bootstrapper> boot;
        size_t n = 0;
        string err;
    };

    void load_ct(const Drv &d, const uint64_t *raw, size_t size, size_t limbs, double scale, Ciphertext &ct)
    {
        ct.upload(*d.ctx, raw, size, limbs, scale);
    }

    struct CoutMute
    {
        std::streambuf *old;
        std::ostringstream sink;
        CoutMute() : old(std::cout.rdbuf(sink.rdbuf()))
        {}
        ~CoutMute()
        {
            std::cout.rdbuf(old);
        }
    };

    int store_all(const Drv &d, const vector<Ciphertext> &res, uint64_t *out, int *out_limbs, double *out_scale)
    {
        for (size_t i = 0; i < res.size(); i++)
        {
            *out_limbs = int(res[i].coeff_modulus_size());
            *out_scale = res[i].scale();
            res[i].download(out + i * res[i].size() * res[i].coeff_modulus_size() * d.n);
        }
        return int(res.size());
    }
} // namespace

#define FD_TRY try {
#define FD_CATCH(d)                                                                                                    \
    }                                                                                                                  \
    catch (const invalid_argument &e)                                                                                  \
    {                                                                                                                  \
        (d)->err = string("invalid_argument: ") + e.what();                                                            \
        return -1;                                                                                                     \
    }                                                                                                                  \
    catch (const logic_error &e)                                                                                       \
    {                                                                                                                  \
        (d)->err = string("logic_error: ") + e.what();                                                                 \
        return -2;                                                                                                     \
    }                                                                                                                  \
    catch (const exception &e)                                                                                         \
    {                                                                                                                  \
        (d)->err = string("exception: ") + e.what();                                                                   \
        return -3;                                                                                                     \
    }                                                                                                                  \
    return 0;

extern "C"
{
    // bits != nullptr: CoeffModulus::Create(N, bits) as M/test/test_full_scheme.hpp:381-389 does;
    // otherwise the given primes.  Never returns nullptr; fd_ok() tells whether the context exists.
    void *fd_create(int log_n, const int *bits, const uint64_t *primes, int count, int device)
    {
        auto d = new Drv();
        try
        {
            EncryptionParameters parms(scheme_type::ckks);
            size_t n = size_t(1) << log_n;
            parms.set_poly_modulus_degree(n);
            if (bits)
            {
                parms.set_coeff_modulus(CoeffModulus::Create(n, vector<int>(bits, bits + count)));
            }
            else
            {
                vector<Modulus> m;
                for (int i = 0; i < count; i++)
                {
                    m.emplace_back(primes[i]);
                }
                parms.set_coeff_modulus(m);
            }
            d->parms = parms;
            d->device = device;
            d->ctx = make_unique<SEALContext>(parms, true, sec_level_type::none, device);
            d->encoder = make_unique<CKKSEncoder>(*d->ctx);
            d->evaluator = make_unique<Evaluator>(*d->ctx, *d->encoder);
            d->n = n;
        }
        catch (const exception &e)
        {
            d->err = e.what();
            d->ctx.reset();
        }
        return d;
    }
    void fd_destroy(void *h)
    {
        delete static_cast<Drv *>(h);
    }
    const char *fd_error(void *h)
    {
        return static_cast<Drv *>(h)->err.c_str();
    }
    int fd_ok(void *h)
    {
        return static_cast<Drv *>(h)->ctx ? 1 : 0;
    }
    // moai_version() of whatever library this driver is bound to: >= 100 for libmoai_b200.so, -1 for the test double
    int fd_backend_version()
    {
        return int(moai_version());
    }
    int fd_n_key_limbs(void *h)
    {
        return int(static_cast<Drv *>(h)->ctx->key_context_data()->parms().coeff_modulus().size());
    }
    void fd_primes(void *h, uint64_t *out)
    {
        auto &cm = static_cast<Drv *>(h)->ctx->key_context_data()->parms().coeff_modulus();
        for (size_t i = 0; i < cm.size(); i++)
        {
            out[i] = cm[i].value();
        }
    }
    // chain_index of the level with `limbs` limbs, through get_context_data (what the modules print)
    int fd_chain_index(void *h, int limbs)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        return int(d->ctx->get_context_data(d->ctx->parms_id_for_limbs(limbs))->chain_index());
        FD_CATCH(d)
    }

    int fd_set_relin(void *h, const uint64_t *key)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        d->rlk.upload(*d->ctx, key);
        FD_CATCH(d)
    }
    int fd_add_galois(void *h, uint32_t elt, const uint64_t *key)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        d->glk.upload(*d->ctx, elt, key);
        FD_CATCH(d)
    }
    int fd_add_galois_fast(void *h, uint32_t elt, const uint64_t *key, int max_limbs)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        d->glk.upload_fast(*d->ctx, elt, key, max_limbs);
        FD_CATCH(d)
    }
    int fd_set_secret(void *h, const uint64_t *sk)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        d->sk.upload(*d->ctx, sk);
        FD_CATCH(d)
    }

    // ---- Evaluator ops on raw ciphertexts: same op codes and argument meaning as ref_eval ----
    enum
    {
        OP_ADD = 0,
        OP_SUB = 1,
        OP_MULTIPLY = 2,
        OP_SQUARE = 3,
        OP_RELINEARIZE = 4,
        OP_RESCALE = 5,
        OP_MOD_SWITCH = 6,
        OP_ROTATE = 7,
        OP_CONJUGATE = 8,
        OP_MULTIPLY_PLAIN = 9,
        OP_ADD_PLAIN = 10,
        OP_SUB_PLAIN = 11,
        OP_NEGATE = 12,
        OP_MULTIPLY_CONST = 13,
        OP_ADD_CONST = 14,
        OP_DOUBLE = 15,
        OP_ADD_REDUCED_ERROR = 16,
        OP_SUB_REDUCED_ERROR = 17,
        OP_MULTIPLY_REDUCED_ERROR = 18,
        OP_MULTIPLY_VECTOR_REDUCED_ERROR = 19,
    };

    int fd_eval(void *h, int op, const uint64_t *a, int size_a, int limbs_a, double scale_a, const uint64_t *b,
                int size_b, int limbs_b, double scale_b, int iarg, double darg, const double *varg, uint64_t *out,
                int *out_size, int *out_limbs, double *out_scale)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        Ciphertext ca, cb, res;
        Plaintext pb;
        load_ct(*d, a, size_a, limbs_a, scale_a, ca);
        auto &ev = *d->evaluator;
        auto need_ct = [&]() { load_ct(*d, b, size_b, limbs_b, scale_b, cb); };
        auto need_pt = [&]() { pb.upload(*d->ctx, b, limbs_b, scale_b); };
        switch (op)
        {
        case OP_ADD:
            need_ct();
            ev.add(ca, cb, res);
            break;
        case OP_SUB:
            need_ct();
            ev.sub(ca, cb, res);
            break;
        case OP_MULTIPLY:
            need_ct();
            ev.multiply(ca, cb, res);
            break;
        case OP_SQUARE:
            ev.square(ca, res);
            break;
        case OP_RELINEARIZE:
            ev.relinearize(ca, d->rlk, res);
            break;
        case OP_RESCALE:
            ev.rescale_to_next(ca, res);
            break;
        case OP_MOD_SWITCH:
            ev.mod_switch_to_next(ca, res);
            break;
        case OP_ROTATE:
            ev.rotate_vector(ca, iarg, d->glk, res);
            break;
        case OP_CONJUGATE:
            ev.complex_conjugate(ca, d->glk, res);
            break;
        case OP_MULTIPLY_PLAIN:
            need_pt();
            ev.multiply_plain(ca, pb, res);
            break;
        case OP_ADD_PLAIN:
            need_pt();
            ev.add_plain(ca, pb, res);
            break;
        case OP_SUB_PLAIN:
            need_pt();
            ev.sub_plain(ca, pb, res);
            break;
        case OP_NEGATE:
            ev.negate(ca, res);
            break;
        case OP_MULTIPLY_CONST:
            ev.multiply_const(ca, darg, res);
            break;
        case OP_ADD_CONST:
            ev.add_const(ca, darg, res);
            break;
        case OP_DOUBLE:
            res = ca;
            ev.double_inplace(res);
            break;
        case OP_ADD_REDUCED_ERROR:
            need_ct();
            ev.add_reduced_error(ca, cb, res);
            break;
        case OP_SUB_REDUCED_ERROR:
            need_ct();
            ev.sub_reduced_error(ca, cb, res);
            break;
        case OP_MULTIPLY_REDUCED_ERROR:
            need_ct();
            ev.multiply_reduced_error(ca, cb, d->rlk, res);
            break;
        case OP_MULTIPLY_VECTOR_REDUCED_ERROR:
        {
            vector<complex<double>> v(d->n / 2);
            for (size_t i = 0; i < v.size(); i++)
            {
                v[i] = complex<double>(varg[2 * i], varg[2 * i + 1]);
            }
            res = ca;
            ev.multiply_vector_inplace_reduced_error(res, v);
            break;
        }
        default:
            throw invalid_argument("unknown op");
        }
        res.download(out);
        *out_size = int(res.size());
        *out_limbs = int(res.coeff_modulus_size());
        *out_scale = res.scale();
        FD_CATCH(d)
    }

    // ---- client-side pieces: PRNG, seeded keys, wire format, Encryptor --------------------------------------
    int fd_prng_bytes(const uint64_t *seed8, int64_t n, uint8_t *out)
    {
        prng_seed_type seed;
        copy_n(seed8, seed.size(), seed.begin());
        Blake2xbPRNG prng(seed);
        int64_t done = 0, piece = 5; // other piece sizes than the reference wrapper: the stream must not care
        while (done < n)
        {
            int64_t take = min<int64_t>(piece, n - done);
            prng.generate(size_t(take), out + done);
            done += take;
            piece = piece * 2 + 3;
        }
        return 0;
    }
    int fd_sample_uniform(void *h, const uint64_t *seed8, uint64_t *out)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        prng_seed_type seed;
        copy_n(seed8, seed.size(), seed.begin());
        auto &cm = d->ctx->key_context_data()->parms().coeff_modulus();
        vector<uint64_t> primes;
        for (auto &m : cm)
        {
            primes.push_back(m.value());
        }
        seal::util::sample_poly_uniform(make_shared<Blake2xbPRNG>(seed), primes, d->n, out);
        FD_CATCH(d)
    }
    // same seeding rule as ref_create: every PRNG the context hands out starts from this seed
    int fd_set_prng_seed(void *h, uint64_t seed, int hamming_weight)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        d->parms.set_secret_key_hamming_weight(size_t(hamming_weight));
        prng_seed_type s;
        for (size_t i = 0; i < s.size(); i++)
        {
            s[i] = seed + 0x9E3779B97F4A7C15ULL * (i + 1);
        }
        d->parms.set_random_generator(make_shared<Blake2xbPRNGFactory>(s));
        // keys already uploaded stay valid only for the old context: this is called right after fd_create
        d->ctx = make_unique<SEALContext>(d->parms, true, sec_level_type::none, d->device);
        d->encoder = make_unique<CKKSEncoder>(*d->ctx);
        d->evaluator = make_unique<Evaluator>(*d->ctx, *d->encoder);
        FD_CATCH(d)
    }
    int fd_set_public_key(void *h, const uint64_t *pk)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        d->pk.upload(*d->ctx, pk);
        FD_CATCH(d)
    }
    // kind 0 = RelinKeys, 1 = GaloisKeys, 2 = PublicKey, in SEAL's wire format (seeded or not)
    int fd_load_keys(void *h, int kind, const uint8_t *buf, int64_t n_bytes, int *seeded_digits)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        stringstream ss(string(reinterpret_cast<const char *>(buf), size_t(n_bytes)));
        size_t seeded = 0;
        if (kind == 0)
        {
            seeded = d->rlk.load(*d->ctx, ss);
        }
        else if (kind == 1)
        {
            seeded = d->glk.load(*d->ctx, ss);
        }
        else
        {
            d->pk.load(*d->ctx, ss);
        }
        if (seeded_digits)
        {
            *seeded_digits = int(seeded);
        }
        FD_CATCH(d)
    }
    // the device copy of a key back as raw residues: kind 0 relin [digits][2][kl][N], 1 Galois key of `elt`,
    // 2 public key [2][kl][N]
    int fd_export_key(void *h, int kind, uint32_t elt, uint64_t *out)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        const size_t kl = size_t(fd_n_key_limbs(h));
        const uint64_t *src = kind == 0 ? d->rlk.device_key() : kind == 1 ? d->glk.device_key(elt) : d->pk.data();
        if (!src)
        {
            throw invalid_argument("key not present");
        }
        const size_t words = (kind == 2 ? 1 : kl - 1) * 2 * kl * d->n;
        seal::detail::chk(moai_memcpy_d2h(d->ctx->handle(), out, src, words * sizeof(uint64_t)));
        FD_CATCH(d)
    }
    // KeyGenerator of the facade: secret key, public key, relinearisation key, Galois keys for `steps`
    // (+ conjugation), or — steps == nullptr — the default set of create_galois_keys(GaloisKeys&)
    int fd_keygen(void *h, const int *steps, int n_steps, int with_conjugate)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        KeyGenerator keygen(*d->ctx);
        d->sk = keygen.secret_key();
        keygen.create_public_key(d->pk);
        keygen.create_relin_keys(d->rlk);
        if (steps)
        {
            vector<uint32_t> elts;
            for (int i = 0; i < n_steps; i++)
            {
                uint32_t e = 0;
                seal::detail::chk(moai_galois_elt_from_step(d->ctx->handle(), steps[i], &e));
                elts.push_back(e);
            }
            if (with_conjugate)
            {
                elts.push_back(uint32_t(2 * d->n - 1));
            }
            keygen.create_galois_keys(elts, d->glk);
        }
        else
        {
            keygen.create_galois_keys(d->glk);
        }
        FD_CATCH(d)
    }
    int fd_export_secret(void *h, uint64_t *out)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        if (!d->sk.data())
        {
            throw invalid_argument("secret key not present");
        }
        seal::detail::chk(moai_memcpy_d2h(d->ctx->handle(), out, d->sk.data(), size_t(fd_n_key_limbs(h)) * d->n * sizeof(uint64_t)));
        FD_CATCH(d)
    }
    int fd_has_galois(void *h, uint32_t elt)
    {
        return static_cast<Drv *>(h)->glk.has_key(elt) ? 1 : 0;
    }

    // Encryptor::encrypt of a raw plaintext -> [2][limbs][N]
    int fd_encrypt(void *h, const uint64_t *pt_raw, int limbs, double scale, uint64_t *out)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        Plaintext pt;
        pt.upload(*d->ctx, pt_raw, limbs, scale);
        Encryptor enc(*d->ctx, d->pk);
        Ciphertext ct;
        enc.encrypt(pt, ct);
        if (ct.size() != 2 || int(ct.coeff_modulus_size()) != limbs || ct.scale() != scale)
        {
            throw logic_error("unexpected ciphertext shape");
        }
        ct.download(out);
        FD_CATCH(d)
    }
    int fd_save_ciphertext(void *h, const uint64_t *ct_raw, int size, int limbs, double scale, uint8_t *buf, int64_t *n_bytes)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        Ciphertext ct;
        load_ct(*d, ct_raw, size, limbs, scale, ct);
        stringstream ss;
        const auto written = ct.save(ss);
        string b = ss.str();
        if (int64_t(b.size()) != int64_t(written) || int64_t(b.size()) > *n_bytes)
        {
            throw logic_error("save size mismatch");
        }
        memcpy(buf, b.data(), b.size());
        *n_bytes = int64_t(b.size());
        FD_CATCH(d)
    }
    int fd_load_ciphertext(void *h, const uint8_t *buf, int64_t n_bytes, uint64_t *ct_raw, int64_t cap_words, int *size,
                           int *limbs, double *scale)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        stringstream ss(string(reinterpret_cast<const char *>(buf), size_t(n_bytes)));
        Ciphertext ct;
        ct.load(*d->ctx, ss);
        if (int64_t(ct.size() * ct.coeff_modulus_size() * d->n) > cap_words)
        {
            throw invalid_argument("buffer too small");
        }
        ct.download(ct_raw);
        *size = int(ct.size());
        *limbs = int(ct.coeff_modulus_size());
        *scale = ct.scale();
        FD_CATCH(d)
    }
    // the reference's batch_input (Batch_encode_encrypt.hpp:8-38), unchanged, through the facade
    int fd_batch_input(void *h, const double *X, int num_X, int num_row, int num_col, double scale, uint64_t *out)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        vector<vector<vector<double>>> Xv(num_X, vector<vector<double>>(num_row, vector<double>(num_col)));
        for (int j = 0; j < num_X; j++)
        {
            for (int k = 0; k < num_row; k++)
            {
                for (int i = 0; i < num_col; i++)
                {
                    Xv[j][k][i] = X[(size_t(j) * num_row + k) * num_col + i];
                }
            }
        }
        vector<Ciphertext> res = batch_input(Xv, num_X, num_row, num_col, scale, *d->ctx, d->pk);
        for (size_t i = 0; i < res.size(); i++)
        {
            res[i].download(out + i * 2 * res[i].coeff_modulus_size() * d->n);
        }
        FD_CATCH(d)
    }

    // The reference's own test program SEAL_ckks_test (M/test/test_SEAL_ckks.hpp:106-250: key generation, encode,
    // encrypt, PI*x^3 + 0.4x + 1 with relinearisation and rescaling, decrypt, decode at N = 8192), UNMODIFIED, on this
    // backend; returns what it printed.
    int fd_reference_seal_ckks_test(char *printed, int cap)
    {
        try
        {
            CoutMute capture;
            SEAL_ckks_test();
            string s = capture.sink.str();
            strncpy(printed, s.c_str(), size_t(cap) - 1);
            printed[cap - 1] = 0;
        }
        catch (const exception &e)
        {
            strncpy(printed, e.what(), size_t(cap) - 1);
            printed[cap - 1] = 0;
            return -1;
        }
        return 0;
    }

    // Value semantics and aliasing (S/ciphertext.h:701-715, S/evaluator.cpp:155-240): destination == operand
    // must give the same residues as distinct objects; a copy must not follow later changes of its source.
    // Returns the number of failed checks.
    int fd_alias_checks(void *h, const uint64_t *a_raw, const uint64_t *b_raw, int limbs, double scale, int *failed)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        auto &ev = *d->evaluator;
        const size_t words2 = size_t(2) * limbs * d->n, words3 = size_t(3) * limbs * d->n;
        auto host = [&](const Ciphertext &c) {
            vector<uint64_t> v(c.size() * c.coeff_modulus_size() * d->n);
            c.download(v.data());
            return v;
        };
        int bad = 0;
        Ciphertext a, b;
        load_ct(*d, a_raw, 2, limbs, scale, a);
        load_ct(*d, b_raw, 2, limbs, scale, b);
        Ciphertext ref, x;
        // add(a, b, a) and add(a, b, b)
        ev.add(a, b, ref);
        x = a;
        ev.add(x, b, x);
        bad += host(x) != host(ref);
        x = b;
        ev.add(a, x, x);
        bad += host(x) != host(ref);
        // sub(a, b, b) = a - b
        ev.sub(a, b, ref);
        x = b;
        ev.sub(a, x, x);
        bad += host(x) != host(ref);
        // multiply(a, b, b), square(a, a)
        ev.multiply(a, b, ref);
        x = b;
        ev.multiply(a, x, x);
        bad += host(x) != host(ref) || x.size() != 3 || x.scale() != ref.scale();
        ev.square(a, ref);
        x = a;
        ev.square(x, x);
        bad += host(x) != host(ref);
        // multiply_plain(a, p, a) with a vector plaintext and with a scalar plaintext
        Plaintext pv, ps;
        vector<double> vals(d->n / 2, 0.25);
        vals[1] = -1.5;
        d->encoder->encode(vals, a.parms_id(), scale, pv);
        d->encoder->encode(0.75, a.parms_id(), scale, ps);
        for (Plaintext *p : { &pv, &ps })
        {
            ev.multiply_plain(a, *p, ref);
            x = a;
            ev.multiply_plain(x, *p, x);
            bad += host(x) != host(ref) || x.scale() != ref.scale();
            x = a;
            ev.multiply_plain_inplace(x, *p);
            bad += host(x) != host(ref);
        }
        // a scalar encoding and the vector encoding of the same constant are the same plaintext
        vector<double> cst(d->n / 2, 0.75);
        Plaintext pc;
        d->encoder->encode(cst, a.parms_id(), scale, pc);
        Ciphertext y;
        ev.multiply_plain(a, pc, y);
        ev.multiply_plain(a, ps, ref);
        bad += host(y) != host(ref);
        ev.add_plain(a, pc, y);
        ev.add_plain(a, ps, ref);
        bad += host(y) != host(ref);
        ev.sub_plain(a, pc, y);
        ev.sub_plain(a, ps, ref);
        bad += host(y) != host(ref);
        // deep copies: the copy keeps its value when the source changes; vector(n, ct) copies n times
        Ciphertext keep = a;
        vector<uint64_t> before = host(keep);
        ev.add_inplace(a, b);
        bad += host(keep) != before;
        vector<Ciphertext> many(3, b);
        ev.negate_inplace(many[1]);
        bad += host(many[0]) != host(b) || host(many[2]) != host(b) || host(many[1]) == host(b);
        // a moved-from / default ciphertext is rejected like an invalid one
        Ciphertext empty;
        try
        {
            ev.add_inplace(empty, b);
            bad++;
        }
        catch (const invalid_argument &)
        {}
        // Plaintext mod_switch_to_inplace drops limbs: equal to encoding at the lower level
        if (limbs >= 2)
        {
            Plaintext lo;
            d->encoder->encode(vals, d->ctx->parms_id_for_limbs(limbs - 1), scale, lo);
            Plaintext hi = pv;
            ev.mod_switch_to_inplace(hi, lo.parms_id());
            vector<uint64_t> h1((limbs - 1) * d->n), h2((limbs - 1) * d->n);
            hi.download(h1.data());
            lo.download(h2.data());
            bad += h1 != h2 || hi.parms_id() != lo.parms_id();
        }
        (void)words2;
        (void)words3;
        *failed = bad;
        FD_CATCH(d)
    }

    // encode(vector<complex>) / encode(vector<double>) / encode(double) -> raw residues [limbs][N]
    int fd_encode_complex(void *h, const double *values, int n_vals, int limbs, double scale, uint64_t *out)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        vector<complex<double>> v(n_vals);
        for (int i = 0; i < n_vals; i++)
        {
            v[i] = complex<double>(values[2 * i], values[2 * i + 1]);
        }
        Plaintext pt;
        d->encoder->encode(v, d->ctx->parms_id_for_limbs(limbs), scale, pt);
        pt.download(out);
        FD_CATCH(d)
    }
    int fd_encode_real(void *h, const double *values, int n_vals, int limbs, double scale, uint64_t *out)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        vector<double> v(values, values + n_vals);
        Plaintext pt;
        d->encoder->encode(v, d->ctx->parms_id_for_limbs(limbs), scale, pt);
        pt.download(out);
        FD_CATCH(d)
    }
    // decode of a raw plaintext -> n/2 complex values (interleaved)
    int fd_decode(void *h, const uint64_t *raw, int limbs, double scale, double *out)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        Plaintext pt;
        pt.upload(*d->ctx, raw, limbs, scale);
        vector<complex<double>> v;
        d->encoder->decode(pt, v);
        for (size_t i = 0; i < v.size(); i++)
        {
            out[2 * i] = v[i].real();
            out[2 * i + 1] = v[i].imag();
        }
        FD_CATCH(d)
    }
    // Decryptor::decrypt -> raw plaintext residues [limbs][N]
    int fd_decrypt(void *h, const uint64_t *ct_raw, int size, int limbs, double scale, uint64_t *out)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        Ciphertext ct;
        load_ct(*d, ct_raw, size, limbs, scale, ct);
        Decryptor dec(*d->ctx, d->sk);
        Plaintext pt;
        dec.decrypt(ct, pt);
        pt.download(out);
        FD_CATCH(d)
    }

    // ---- the reference's modules, same signatures as ref_ct_pt_matmul / ref_gelu_v2 / ref_layernorm /
    //      ref_ct_ct_matmul in oracle/refbuild/ref_wrap.cpp ----
    int fd_ct_pt_matmul(void *h, int variant, const uint64_t *X, const double *W, const int *mask, int K, int C, int limbs,
                        double scale, uint64_t *out, double *seconds)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        size_t ctsz = size_t(2) * limbs * d->n;
        vector<Ciphertext> enc_X(K);
        for (int j = 0; j < K; j++)
        {
            load_ct(*d, X + size_t(j) * ctsz, 2, limbs, scale, enc_X[j]);
        }
        vector<vector<double>> Wm(K, vector<double>(C));
        for (int j = 0; j < K; j++)
        {
            for (int i = 0; i < C; i++)
            {
                Wm[j][i] = W[size_t(j) * C + i];
            }
        }
        vector<int> bias_vec;
        if (mask)
        {
            bias_vec.assign(mask, mask + d->n / 2);
        }
        vector<Ciphertext> res;
        auto t0 = chrono::steady_clock::now();
        if (variant == 0)
        {
            res = ct_pt_matrix_mul_wo_pre(enc_X, Wm, K, C, K, *d->ctx);
        }
        else if (variant == 1)
        {
            res = ct_pt_matrix_mul_wo_pre_large(enc_X, Wm, K, C, K, *d->ctx);
        }
        else
        {
            res = ct_pt_matrix_mul_wo_pre_w_mask(enc_X, Wm, bias_vec, K, C, K, *d->ctx);
        }
        d->ctx->synchronize();
        if (seconds)
        {
            *seconds = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
        }
        size_t outsz = size_t(2) * (limbs - 1) * d->n;
        for (int i = 0; i < C; i++)
        {
            if (res[i].size() != 2 || res[i].coeff_modulus_size() != size_t(limbs - 1))
            {
                throw logic_error("unexpected output shape");
            }
            res[i].download(out + size_t(i) * outsz);
        }
        FD_CATCH(d)
    }

    int fd_gelu_v2(void *h, const uint64_t *x, int count, int limbs, double scale, uint64_t *out, int *out_limbs,
                   double *out_scale)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * d->n;
        for (int i = 0; i < count; i++)
        {
            Ciphertext ct;
            load_ct(*d, x + size_t(i) * ctsz, 2, limbs, scale, ct);
            Ciphertext res = gelu_v2(ct, *d->ctx, d->rlk, d->sk);
            *out_limbs = int(res.coeff_modulus_size());
            *out_scale = res.scale();
            res.download(out + size_t(i) * 2 * res.coeff_modulus_size() * d->n);
        }
        FD_CATCH(d)
    }

    // `printed` (optional, capacity cap): what the module wrote to std::cout (its debug decryptions)
    int fd_layernorm(void *h, int variant, const uint64_t *x, int num_ct, int limbs, double scale, const double *gamma,
                     const double *beta, const int *bias_vec, uint64_t *out, int *out_limbs, double *out_scale,
                     char *printed, int cap)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * d->n;
        vector<Ciphertext> xs(num_ct);
        for (int i = 0; i < num_ct; i++)
        {
            load_ct(*d, x + size_t(i) * ctsz, 2, limbs, scale, xs[i]);
        }
        vector<double> g(gamma, gamma + num_ct), b(beta, beta + num_ct);
        vector<int> bv(bias_vec, bias_vec + d->n / 2);
        vector<Ciphertext> res = variant == 1 ? layernorm(xs, g, b, bv, *d->ctx, d->rlk, d->sk)
                                              : layernorm2(xs, g, b, bv, *d->ctx, d->rlk, d->sk);
        store_all(*d, res, out, out_limbs, out_scale);
        if (printed && cap > 0)
        {
            string s = mute.sink.str();
            strncpy(printed, s.c_str(), size_t(cap) - 1);
            printed[cap - 1] = 0;
        }
        FD_CATCH(d)
    }

    int fd_ct_ct_matmul(void *h, int which, const uint64_t *X, int nX, const uint64_t *W, int nW, int limbs, double scale_X,
                        double scale_W, int col_X, int row_X, int col_W, int row_W, int num_batch, uint64_t *out,
                        int *out_count, int *out_limbs, double *out_scale)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * d->n;
        vector<Ciphertext> xs(nX), ws(nW);
        for (int i = 0; i < nX; i++)
        {
            load_ct(*d, X + size_t(i) * ctsz, 2, limbs, scale_X, xs[i]);
        }
        for (int i = 0; i < nW; i++)
        {
            load_ct(*d, W + size_t(i) * ctsz, 2, limbs, scale_W, ws[i]);
        }
        vector<Ciphertext> res =
            which == 0 ? ct_ct_matrix_mul_colpacking(xs, ws, d->glk, d->rlk, *d->ctx, col_X, row_X, col_W, row_W, num_batch)
                       : ct_ct_matrix_mul_diagpacking(xs, ws, d->glk, d->rlk, *d->ctx, col_X, row_X, col_W, row_W,
                                                      num_batch);
        *out_count = store_all(*d, res, out, out_limbs, out_scale);
        FD_CATCH(d)
    }

    // exp / inverse of softmax.hpp (:9-82): no bootstrapping, bit-exact against the same headers on real SEAL
    // only where NTL exists; here they are compared with the op-for-op C-ABI modules (moai_exp / moai_inverse)
    int fd_exp(void *h, const uint64_t *x, int limbs, double scale, uint64_t *out, int *out_limbs, double *out_scale)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        CoutMute mute;
        Ciphertext ct;
        load_ct(*d, x, 2, limbs, scale, ct);
        Ciphertext res = exp(ct, *d->ctx, d->rlk);
        *out_limbs = int(res.coeff_modulus_size());
        *out_scale = res.scale();
        res.download(out);
        FD_CATCH(d)
    }
    int fd_inverse(void *h, const uint64_t *x, int limbs, double scale, int iter, uint64_t *out, int *out_limbs,
                   double *out_scale)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        CoutMute mute;
        Ciphertext ct;
        load_ct(*d, x, 2, limbs, scale, ct);
        Ciphertext res = inverse(ct, *d->ctx, d->rlk, iter);
        *out_limbs = int(res.coeff_modulus_size());
        *out_scale = res.scale();
        res.download(out);
        FD_CATCH(d)
    }

    // ---- bootstrapping through the facade's Bootstrapper (GPU only) ----
    // constructor arguments as M/test/test_full_scheme.hpp:413-431; returns the rotation steps it needs
    int fd_boot_create(void *h, int loge, int logn, int total_level, double final_scale, int boundary_K, int deg,
                       int scale_factor, int hoisting, int *steps, int cap, int *n_steps)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        d->boot = make_unique<Bootstrapper>(loge, logn, logn, total_level, final_scale, boundary_K, deg, scale_factor, 1,
                                            *d->ctx, d->rlk, d->glk);
        d->boot->set_hoisting(hoisting != 0);
        d->boot->prepare_mod_polynomial();
        vector<int> v;
        d->boot->addLeftRotKeys_Linear_to_vector_3(v);
        d->boot->slot_vec.push_back(logn);
        d->boot->generate_LT_coefficient_3();
        *n_steps = int(v.size());
        for (int i = 0; i < int(v.size()) && i < cap; i++)
        {
            steps[i] = v[i];
        }
        FD_CATCH(d)
    }
    int fd_bootstrap_3(void *h, const uint64_t *x, double scale, uint64_t *out, int *out_limbs, double *out_scale)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        if (!d->boot)
        {
            throw logic_error("fd_boot_create first");
        }
        Ciphertext ct, rtn;
        load_ct(*d, x, 2, 1, scale, ct);
        d->boot->bootstrap_3(rtn, ct);
        *out_limbs = int(rtn.coeff_modulus_size());
        *out_scale = rtn.scale();
        rtn.download(out);
        FD_CATCH(d)
    }
    // argument checks of bootstrap_3 on a ciphertext at `limbs` limbs (expected to throw for limbs != 1)
    int fd_bootstrap_limbs(void *h, const uint64_t *x, int limbs, double scale)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        if (!d->boot)
        {
            throw logic_error("fd_boot_create first");
        }
        Ciphertext ct, rtn;
        load_ct(*d, x, 2, limbs, scale, ct);
        d->boot->bootstrap_3(rtn, ct);
        FD_CATCH(d)
    }
    // Request combining of the facade Bootstrapper: n ciphertexts, one bootstrap_3 call each from an OpenMP loop like
    // the reference's driver (test_full_scheme.hpp:654-660).  x: [n][2][1][N] -> out: [n][2][L-13][N]; *device_calls =
    // device calls the combiner made.
    int fd_boot_combined(void *h, const uint64_t *x, int n_cts, double scale, int real_slots, int max_batch, int linger_us,
                         uint64_t *out, int *out_limbs, int *device_calls)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        if (!d->boot)
        {
            throw logic_error("fd_boot_create first");
        }
        d->boot->set_combining(true, real_slots != 0, max_batch, linger_us);
        const size_t before = d->boot->combined_device_calls();
        vector<Ciphertext> in(n_cts), rtn(n_cts);
        for (int i = 0; i < n_cts; i++)
        {
            load_ct(*d, x + size_t(i) * 2 * d->n, 2, 1, scale, in[i]);
        }
        vector<string> errors(n_cts);
#pragma omp parallel for
        for (int i = 0; i < (max_batch < 0 ? 0 : n_cts); i++)
        {
            try
            {
                d->boot->bootstrap_3(rtn[i], in[i]);
            }
            catch (const exception &e)
            {
                errors[i] = e.what();
            }
        }
        d->boot->set_combining(false);
        if (max_batch < 0)
        {
            // the explicit batch overload instead of the combiner
            rtn.clear();
            d->boot->bootstrap_3(rtn, in, real_slots != 0);
        }
        for (int i = 0; i < n_cts; i++)
        {
            if (!errors[i].empty())
            {
                throw logic_error("request " + to_string(i) + ": " + errors[i]);
            }
            *out_limbs = int(rtn[i].coeff_modulus_size());
            rtn[i].download(out + size_t(i) * 2 * rtn[i].coeff_modulus_size() * d->n);
        }
        *device_calls = int(d->boot->combined_device_calls() - before);
        FD_CATCH(d)
    }

    // Thread lanes (SEALContext::set_thread_lanes): n ciphertexts processed one per iteration of an OpenMP loop, the way
    // every module header of the reference calls its shared Evaluator (`#pragma omp parallel for`), here with a chain
    // that produces, hands over and frees objects across calls:  out[i] = rescale(relin(x[i]^2) + x[i] * x[(i+1) % n]
    // relinearized) rotated by one step.  x: [n][2][limbs][N] -> out: [n][2][limbs-1][N].  *threads = OpenMP threads
    // that took part.
    int fd_parallel_chain(void *h, const uint64_t *x, int n_cts, int limbs, double scale, int lanes, int with_rotation,
                          uint64_t *out, int *threads, double *loop_ms)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        d->ctx->set_thread_lanes(lanes != 0);
        vector<Ciphertext> in(n_cts), rtn(n_cts);
        const size_t per = size_t(2) * limbs * d->n;
        for (int i = 0; i < n_cts; i++)
        {
            load_ct(*d, x + size_t(i) * per, 2, limbs, scale, in[i]);
        }
        vector<string> errors(n_cts);
        vector<int> tids(n_cts, 0);
        d->ctx->synchronize();
        const auto t0 = chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1)
        for (int i = 0; i < n_cts; i++)
        {
            try
            {
                tids[i] = omp_get_thread_num();
                Ciphertext sq, pr;
                d->evaluator->square(in[i], sq);
                d->evaluator->relinearize_inplace(sq, d->rlk);
                d->evaluator->multiply(in[i], in[(i + 1) % n_cts], pr);
                d->evaluator->relinearize_inplace(pr, d->rlk);
                d->evaluator->add_inplace(sq, pr);
                d->evaluator->rescale_to_next_inplace(sq);
                if (with_rotation)
                {
                    d->evaluator->rotate_vector(sq, 1, d->glk, rtn[i]);
                }
                else
                {
                    rtn[i] = sq;
                }
            }
            catch (const exception &e)
            {
                errors[i] = e.what();
            }
        }
        *loop_ms = chrono::duration<double, milli>(chrono::steady_clock::now() - t0).count(); // every call has drained
        d->ctx->set_thread_lanes(false);
        set<int> used(tids.begin(), tids.end());
        *threads = int(used.size());
        for (int i = 0; i < n_cts; i++)
        {
            if (!errors[i].empty())
            {
                throw logic_error("item " + to_string(i) + ": " + errors[i]);
            }
            rtn[i].download(out + size_t(i) * 2 * (limbs - 1) * d->n);
        }
        FD_CATCH(d)
    }

    // softmax_boot of the reference (softmax.hpp:308-581) on `num` ciphertexts
    int fd_softmax_boot(void *h, const uint64_t *x, int num, int limbs, double scale, const int *bias_vec, int input_num,
                        int iter, int layer_id, uint64_t *out, int *out_limbs, double *out_scale)
    {
        auto d = static_cast<Drv *>(h);
        FD_TRY
        if (!d->boot)
        {
            throw logic_error("fd_boot_create first");
        }
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * d->n;
        vector<Ciphertext> xs(num);
        for (int i = 0; i < num; i++)
        {
            load_ct(*d, x + size_t(i) * ctsz, 2, limbs, scale, xs[i]);
        }
        vector<int> bv(bias_vec, bias_vec + d->n / 2);
        vector<Ciphertext> res = softmax_boot(xs, bv, input_num, *d->ctx, d->rlk, iter, d->sk, *d->boot, layer_id);
        store_all(*d, res, out, out_limbs, out_scale);
        FD_CATCH(d)
    }
}
