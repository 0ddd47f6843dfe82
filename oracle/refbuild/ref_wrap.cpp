// oracle/_ref wrapper — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// A flat C interface over the *unmodified* vendored SEAL-4.1-bs of the reference
// (/root/reference/thirdparty/SEAL-4.1-bs/native/src/seal, compiled where it lies by
// oracle/refbuild/Makefile) so that Python tests can
//   * pin oracle/ckks_oracle.c (the C restatement) against the real implementation, and
//   * compare the CUDA path bit-for-bit on identical keys / randomness / inputs.
// Every buffer is raw uint64 residues in SEAL's own layout [poly][limb][coeff]
// (S/ciphertext.h:339-370).  Nothing here re-implements arithmetic: each entry point loads the
// raw residues into seal:: value types and calls the reference's public API.
#include "seal/seal.h"
#include "seal/util/rlwe.h"
#include <cstdint>
#include <cstring>
#include <complex>
#include <memory>
#include <vector>
#include <map>
#include <string>
#include <chrono>
#include <omp.h>
#include <iostream>
#include <sys/time.h> // gettimeofday: the reference headers rely on M/include.hpp:17 for it

// The reference's own module code, unmodified (M/source/matrix_mul/Ct_pt_matrix_mul.hpp).
#include "source/matrix_mul/Batch_encode_encrypt.hpp"
#include "source/matrix_mul/Ct_pt_matrix_mul.hpp"
#include "source/matrix_mul/Ct_ct_matrix_mul.hpp"
#include "source/non_linear_func/layernorm.hpp"
#include "source/non_linear_func/gelu_others.hpp"
// The reference's bootstrapper and the modules that contain a bootstrapping, unmodified.  Their NTL dependency
// (Polynomial.h:3, Remez.h:3, func.h:3-4) is met by oracle/refbuild/ntl_shim (NTL::RR over libmpfr.so.6).
#include "source/bootstrapping/Bootstrapper.h"
#include "source/non_linear_func/softmax.hpp"
#include "source/att_block/single_att_block.hpp"
#include <sstream>

using namespace seal;
using namespace std;

namespace
{
    struct Ref
    {
        unique_ptr<SEALContext> ctx;
        unique_ptr<KeyGenerator> keygen;
        unique_ptr<CKKSEncoder> encoder;
        unique_ptr<Evaluator> evaluator;
        unique_ptr<Encryptor> encryptor;
        unique_ptr<Decryptor> decryptor;
        SecretKey sk;
        PublicKey pk;
        RelinKeys rlk;
        GaloisKeys glk;
        bool have_rlk = false;
        unique_ptr<Bootstrapper> boot; // the reference's Bootstrapper (M/source/bootstrapping/Bootstrapper.h:14-221)
        size_t n = 0;
        size_t n_data_limbs = 0; // limbs at the first (fresh-ciphertext) level
        string err;
    };

    parms_id_type parms_for_limbs(const Ref &r, size_t limbs)
    {
        auto cd = r.ctx->first_context_data();
        while (cd && cd->parms().coeff_modulus().size() != limbs)
        {
            cd = cd->next_context_data();
        }
        if (!cd)
        {
            throw invalid_argument("no level with that many limbs");
        }
        return cd->parms_id();
    }

    void load_ct(const Ref &r, const uint64_t *raw, size_t size, size_t limbs, double scale, Ciphertext &ct)
    {
        ct.resize(*r.ctx, parms_for_limbs(r, limbs), size);
        ct.is_ntt_form() = true;
        ct.scale() = scale;
        memcpy(ct.data(), raw, size * limbs * r.n * sizeof(uint64_t));
    }

    void store_ct(const Ref &r, const Ciphertext &ct, uint64_t *raw)
    {
        memcpy(raw, ct.data(), ct.size() * ct.coeff_modulus_size() * r.n * sizeof(uint64_t));
    }

    void load_pt(const Ref &r, const uint64_t *raw, size_t limbs, double scale, Plaintext &pt)
    {
        pt.parms_id() = parms_id_zero;
        pt.resize(limbs * r.n);
        memcpy(pt.data(), raw, limbs * r.n * sizeof(uint64_t));
        pt.parms_id() = parms_for_limbs(r, limbs);
        pt.scale() = scale;
    }
} // namespace

#define REF_TRY(r) try {
#define REF_CATCH(r)                                                                                                   \
    }                                                                                                                  \
    catch (const exception &e)                                                                                         \
    {                                                                                                                  \
        (r)->err = e.what();                                                                                           \
        return -1;                                                                                                     \
    }                                                                                                                  \
    return 0;

extern "C"
{
    // Build the CKKS context the way M/test/test_full_scheme.hpp:381-389 does:
    // CoeffModulus::Create(N, bits), sparse ternary secret of the given Hamming weight (0 = dense),
    // SEALContext(parms, true, sec_level_type::none).  `seed` makes all SEAL randomness deterministic.
    void *ref_create(int log_n, const int *bits, int n_bits, int hamming_weight, uint64_t seed)
    {
        auto r = new Ref();
        try
        {
            EncryptionParameters parms(scheme_type::ckks);
            size_t n = size_t(1) << log_n;
            parms.set_poly_modulus_degree(n);
            vector<int> bv(bits, bits + n_bits);
            parms.set_coeff_modulus(CoeffModulus::Create(n, bv));
            parms.set_secret_key_hamming_weight(size_t(hamming_weight));
            prng_seed_type s;
            for (size_t i = 0; i < s.size(); i++)
            {
                s[i] = seed + 0x9E3779B97F4A7C15ULL * (i + 1);
            }
            parms.set_random_generator(make_shared<Blake2xbPRNGFactory>(s));
            r->ctx = make_unique<SEALContext>(parms, true, sec_level_type::none);
            if (!r->ctx->parameters_set())
            {
                throw invalid_argument(string("bad parameters: ") + r->ctx->parameter_error_message());
            }
            r->n = n;
            r->n_data_limbs = r->ctx->first_context_data()->parms().coeff_modulus().size();
            r->keygen = make_unique<KeyGenerator>(*r->ctx);
            r->sk = r->keygen->secret_key();
            r->keygen->create_public_key(r->pk);
            r->encoder = make_unique<CKKSEncoder>(*r->ctx);
            r->evaluator = make_unique<Evaluator>(*r->ctx, *r->encoder);
            r->encryptor = make_unique<Encryptor>(*r->ctx, r->pk);
            r->decryptor = make_unique<Decryptor>(*r->ctx, r->sk);
        }
        catch (const exception &e)
        {
            r->err = e.what();
        }
        return r;
    }

    void ref_destroy(void *h)
    {
        delete static_cast<Ref *>(h);
    }

    const char *ref_error(void *h)
    {
        return static_cast<Ref *>(h)->err.c_str();
    }

    int ref_ok(void *h)
    {
        return static_cast<Ref *>(h)->ctx && static_cast<Ref *>(h)->err.empty() ? 1 : 0;
    }

    // All primes of the key level (data primes then the special prime).
    int ref_n_key_limbs(void *h)
    {
        auto r = static_cast<Ref *>(h);
        return int(r->ctx->key_context_data()->parms().coeff_modulus().size());
    }

    void ref_primes(void *h, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        auto &cm = r->ctx->key_context_data()->parms().coeff_modulus();
        for (size_t i = 0; i < cm.size(); i++)
        {
            out[i] = cm[i].value();
        }
    }

    // NTT tables of key-level limb `limb`: root powers (operand, quotient) in SEAL's stored order,
    // S/util/ntt.cpp:254-296.
    void ref_ntt_tables(void *h, int limb, uint64_t *root_op, uint64_t *root_quo, uint64_t *inv_root_op,
                        uint64_t *inv_root_quo, uint64_t *inv_n_op_quo)
    {
        auto r = static_cast<Ref *>(h);
        auto &t = r->ctx->key_context_data()->small_ntt_tables()[limb];
        for (size_t i = 0; i < r->n; i++)
        {
            root_op[i] = t.get_from_root_powers(i).operand;
            root_quo[i] = t.get_from_root_powers(i).quotient;
            inv_root_op[i] = t.get_from_inv_root_powers(i).operand;
            inv_root_quo[i] = t.get_from_inv_root_powers(i).quotient;
        }
        inv_n_op_quo[0] = t.inv_degree_modulo().operand;
        inv_n_op_quo[1] = t.inv_degree_modulo().quotient;
    }

    // In-place forward / inverse negacyclic NTT of `count` consecutive polynomials for limb `limb`.
    void ref_ntt(void *h, int limb, uint64_t *data, int count)
    {
        auto r = static_cast<Ref *>(h);
        auto &t = r->ctx->key_context_data()->small_ntt_tables()[limb];
        for (int c = 0; c < count; c++)
        {
            util::ntt_negacyclic_harvey(data + size_t(c) * r->n, t);
        }
    }

    void ref_intt(void *h, int limb, uint64_t *data, int count)
    {
        auto r = static_cast<Ref *>(h);
        auto &t = r->ctx->key_context_data()->small_ntt_tables()[limb];
        for (int c = 0; c < count; c++)
        {
            util::inverse_ntt_negacyclic_harvey(data + size_t(c) * r->n, t);
        }
    }

    // ---- keys -------------------------------------------------------------------------------
    // Secret key in NTT form at the key level: [limb][N].
    void ref_secret_key(void *h, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        memcpy(out, r->sk.data().data(), size_t(ref_n_key_limbs(h)) * r->n * sizeof(uint64_t));
    }

    // Public key: [2][key limbs][N].
    void ref_public_key(void *h, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        memcpy(out, r->pk.data().data(), 2 * size_t(ref_n_key_limbs(h)) * r->n * sizeof(uint64_t));
    }

    int ref_make_relin_key(void *h)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        r->keygen->create_relin_keys(r->rlk);
        r->have_rlk = true;
        REF_CATCH(r)
    }

    // Adds Galois keys for the given rotation steps (0 is ignored by SEAL; conjugation is requested
    // with `with_conjugate`).
    int ref_make_galois_keys(void *h, const int *steps, int n_steps, int with_conjugate)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        vector<uint32_t> elts;
        auto gt = r->ctx->key_context_data()->galois_tool();
        for (int i = 0; i < n_steps; i++)
        {
            if (steps[i] != 0)
            {
                elts.push_back(gt->get_elt_from_step(steps[i]));
            }
        }
        if (with_conjugate)
        {
            elts.push_back(gt->get_elt_from_step(0));
        }
        GaloisKeys add;
        r->keygen->create_galois_keys(elts, add);
        // merge into r->glk
        auto &dst = r->glk.data();
        auto &src = add.data();
        if (dst.size() < src.size())
        {
            dst.resize(src.size());
        }
        for (size_t i = 0; i < src.size(); i++)
        {
            if (!src[i].empty())
            {
                dst[i] = src[i];
            }
        }
        r->glk.parms_id() = add.parms_id();
        REF_CATCH(r)
    }

    uint32_t ref_galois_elt_from_step(void *h, int step)
    {
        auto r = static_cast<Ref *>(h);
        return r->ctx->key_context_data()->galois_tool()->get_elt_from_step(step);
    }

    int ref_has_galois_key(void *h, uint32_t elt)
    {
        auto r = static_cast<Ref *>(h);
        return r->glk.has_key(elt) ? 1 : 0;
    }

    // One key-switching key as [digit][poly(2)][key limb][N]  (S/kswitchkeys.h:335-340).
    // kind 0 = relin key, kind 1 = Galois key for element `elt`.
    int ref_export_kswitch_key(void *h, int kind, uint32_t elt, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        const vector<PublicKey> *k = nullptr;
        if (kind == 0)
        {
            k = &r->rlk.key(2);
        }
        else
        {
            k = &r->glk.key(elt);
        }
        size_t kl = size_t(ref_n_key_limbs(h));
        size_t per = 2 * kl * r->n;
        for (size_t j = 0; j < k->size(); j++)
        {
            memcpy(out + j * per, (*k)[j].data().data(), per * sizeof(uint64_t));
        }
        REF_CATCH(r)
    }

    // ---- client side -------------------------------------------------------------------------
    // encode(vector<complex>) at `limbs` limbs.  values: interleaved re,im, `n_vals` complex numbers.
    int ref_encode_complex(void *h, const double *values, int n_vals, int limbs, double scale, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        vector<complex<double>> v(n_vals);
        for (int i = 0; i < n_vals; i++)
        {
            v[i] = complex<double>(values[2 * i], values[2 * i + 1]);
        }
        Plaintext pt;
        r->encoder->encode(v, parms_for_limbs(*r, limbs), scale, pt);
        memcpy(out, pt.data(), size_t(limbs) * r->n * sizeof(uint64_t));
        REF_CATCH(r)
    }

    int ref_encode_real(void *h, const double *values, int n_vals, int limbs, double scale, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        vector<double> v(values, values + n_vals);
        Plaintext pt;
        r->encoder->encode(v, parms_for_limbs(*r, limbs), scale, pt);
        memcpy(out, pt.data(), size_t(limbs) * r->n * sizeof(uint64_t));
        REF_CATCH(r)
    }

    int ref_encode_scalar(void *h, double value, int limbs, double scale, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        Plaintext pt;
        r->encoder->encode(value, parms_for_limbs(*r, limbs), scale, pt);
        memcpy(out, pt.data(), size_t(limbs) * r->n * sizeof(uint64_t));
        REF_CATCH(r)
    }

    // decode a plaintext given as raw residues -> n/2 complex values (interleaved)
    int ref_decode(void *h, const uint64_t *raw, int limbs, double scale, double *out)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        Plaintext pt;
        load_pt(*r, raw, limbs, scale, pt);
        vector<complex<double>> v;
        r->encoder->decode(pt, v);
        for (size_t i = 0; i < v.size(); i++)
        {
            out[2 * i] = v[i].real();
            out[2 * i + 1] = v[i].imag();
        }
        REF_CATCH(r)
    }

    // Encrypt a raw plaintext (public-key encryption, S/encryptor.cpp) -> size-2 ct at `limbs`.
    int ref_encrypt(void *h, const uint64_t *pt_raw, int limbs, double scale, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        Plaintext pt;
        load_pt(*r, pt_raw, limbs, scale, pt);
        Ciphertext ct;
        r->encryptor->encrypt(pt, ct);
        store_ct(*r, ct, out);
        REF_CATCH(r)
    }

    // Decrypt raw ct -> raw plaintext residues [limbs][N] (NTT form).
    int ref_decrypt(void *h, const uint64_t *ct_raw, int size, int limbs, double scale, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        Ciphertext ct;
        load_ct(*r, ct_raw, size, limbs, scale, ct);
        Plaintext pt;
        r->decryptor->decrypt(ct, pt);
        memcpy(out, pt.data(), size_t(limbs) * r->n * sizeof(uint64_t));
        REF_CATCH(r)
    }

    // ---- Evaluator ops on raw ciphertexts ----------------------------------------------------
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

    // Generic entry: a (size_a, limbs_a, scale_a), optional b (ct or plain raw, limbs_b, scale_b),
    // integer arg `iarg` (rotation steps), double arg `darg` (constant), vector arg for
    // multiply_vector (complex interleaved, n/2 values).  Output written to out; *out_size,
    // *out_limbs, *out_scale filled.
    int ref_eval(void *h, int op, const uint64_t *a, int size_a, int limbs_a, double scale_a, const uint64_t *b,
                 int size_b, int limbs_b, double scale_b, int iarg, double darg, const double *varg, uint64_t *out,
                 int *out_size, int *out_limbs, double *out_scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        Ciphertext ca, cb, res;
        Plaintext pb;
        load_ct(*r, a, size_a, limbs_a, scale_a, ca);
        auto &ev = *r->evaluator;
        switch (op)
        {
        case OP_ADD:
            load_ct(*r, b, size_b, limbs_b, scale_b, cb);
            ev.add(ca, cb, res);
            break;
        case OP_SUB:
            load_ct(*r, b, size_b, limbs_b, scale_b, cb);
            ev.sub(ca, cb, res);
            break;
        case OP_MULTIPLY:
            load_ct(*r, b, size_b, limbs_b, scale_b, cb);
            ev.multiply(ca, cb, res);
            break;
        case OP_SQUARE:
            ev.square(ca, res);
            break;
        case OP_RELINEARIZE:
            ev.relinearize(ca, r->rlk, res);
            break;
        case OP_RESCALE:
            ev.rescale_to_next(ca, res);
            break;
        case OP_MOD_SWITCH:
            ev.mod_switch_to_next(ca, res);
            break;
        case OP_ROTATE:
            ev.rotate_vector(ca, iarg, r->glk, res);
            break;
        case OP_CONJUGATE:
            ev.complex_conjugate(ca, r->glk, res);
            break;
        case OP_MULTIPLY_PLAIN:
            load_pt(*r, b, limbs_b, scale_b, pb);
            ev.multiply_plain(ca, pb, res);
            break;
        case OP_ADD_PLAIN:
            load_pt(*r, b, limbs_b, scale_b, pb);
            ev.add_plain(ca, pb, res);
            break;
        case OP_SUB_PLAIN:
            load_pt(*r, b, limbs_b, scale_b, pb);
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
            load_ct(*r, b, size_b, limbs_b, scale_b, cb);
            ev.add_reduced_error(ca, cb, res);
            break;
        case OP_SUB_REDUCED_ERROR:
            load_ct(*r, b, size_b, limbs_b, scale_b, cb);
            ev.sub_reduced_error(ca, cb, res);
            break;
        case OP_MULTIPLY_REDUCED_ERROR:
            load_ct(*r, b, size_b, limbs_b, scale_b, cb);
            ev.multiply_reduced_error(ca, cb, r->rlk, res);
            break;
        case OP_MULTIPLY_VECTOR_REDUCED_ERROR:
        {
            vector<complex<double>> v(r->n / 2);
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
        store_ct(*r, res, out);
        *out_size = int(res.size());
        *out_limbs = int(res.coeff_modulus_size());
        *out_scale = res.scale();
        REF_CATCH(r)
    }

    // ---- MOAI module level: the reference's ct-pt matmuls ----------------
    // variant 0: ct_pt_matrix_mul_wo_pre          (Ct_pt_matrix_mul.hpp:4-49)
    // variant 1: ct_pt_matrix_mul_wo_pre_large    (:51-101, col_W must be a multiple of 128)
    // variant 2: ct_pt_matrix_mul_wo_pre_w_mask   (:103-170, col_W must be a multiple of 128)
    // variant 3: the loop body of variant 2 (:124-165) over an arbitrary number of columns, one
    //            column per OpenMP iteration — used to time a bounded sample on small hosts.
    // X: [K][2][limbs][n]; W row-major K x C; mask: n/2 ints (variants 2, 3); out [C][2][limbs-1][n].
    int ref_ct_pt_matmul(void *h, int variant, const uint64_t *X, const double *W, const int *mask, int K, int C,
                         int limbs, double scale, uint64_t *out, double *seconds)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        size_t ctsz = size_t(2) * limbs * r->n;
        vector<Ciphertext> enc_X(K);
        for (int j = 0; j < K; j++)
        {
            load_ct(*r, X + size_t(j) * ctsz, 2, limbs, scale, enc_X[j]);
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
            bias_vec.assign(mask, mask + r->n / 2);
        }
        vector<Ciphertext> res;
        auto t0 = chrono::steady_clock::now();
        if (variant == 0)
        {
            res = ct_pt_matrix_mul_wo_pre(enc_X, Wm, K, C, K, *r->ctx);
        }
        else if (variant == 1)
        {
            res = ct_pt_matrix_mul_wo_pre_large(enc_X, Wm, K, C, K, *r->ctx);
        }
        else if (variant == 2)
        {
            res = ct_pt_matrix_mul_wo_pre_w_mask(enc_X, Wm, bias_vec, K, C, K, *r->ctx);
        }
        else
        {
            res.resize(C);
            size_t slot_count = r->encoder->slot_count();
#pragma omp parallel for schedule(dynamic)
            for (int i = 0; i < C; i++)
            {
                for (int j = 0; j < K; j++)
                {
                    vector<double> tempw(slot_count, 0);
                    for (size_t kk = 0; kk < slot_count; ++kk)
                    {
                        if (bias_vec[kk] == 1)
                        {
                            tempw[kk] = Wm[j][i];
                        }
                    }
                    Plaintext ecd;
                    r->encoder->encode(tempw, enc_X[j].parms_id(), enc_X[j].scale(), ecd);
                    if (j == 0)
                    {
                        r->evaluator->multiply_plain(enc_X[0], ecd, res[i]);
                    }
                    else
                    {
                        Ciphertext tempx;
                        r->evaluator->multiply_plain(enc_X[j], ecd, tempx);
                        r->evaluator->add_inplace(res[i], tempx);
                    }
                }
                r->evaluator->rescale_to_next_inplace(res[i]);
                res[i].scale() = scale;
            }
        }
        auto t1 = chrono::steady_clock::now();
        if (seconds)
        {
            *seconds = chrono::duration<double>(t1 - t0).count();
        }
        size_t outsz = size_t(2) * (limbs - 1) * r->n;
        for (int i = 0; i < C; i++)
        {
            if (res[i].size() != 2 || res[i].coeff_modulus_size() != size_t(limbs - 1))
            {
                throw logic_error("unexpected output shape");
            }
            memcpy(out + size_t(i) * outsz, res[i].data(), outsz * sizeof(uint64_t));
        }
        REF_CATCH(r)
    }

    // The reference's modules print debug decryptions to std::cout; silence them while they run.
    struct CoutMute
    {
        std::streambuf *old;
        std::ostringstream sink;
        CoutMute() : old(std::cout.rdbuf())
        {
            if (!getenv("REF_VERBOSE")) // REF_VERBOSE=1 lets the reference's debug decryptions through
            {
                std::cout.rdbuf(sink.rdbuf());
            }
        }
        ~CoutMute()
        {
            std::cout.rdbuf(old);
        }
    };

    // gelu_v2 (M/source/non_linear_func/gelu_others.hpp:4-154) on `count` ciphertexts [count][2][limbs][n]
    int ref_gelu_v2(void *h, const uint64_t *x, int count, int limbs, double scale, uint64_t *out, int *out_limbs,
                    double *out_scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * r->n;
        for (int i = 0; i < count; i++)
        {
            Ciphertext ct;
            load_ct(*r, x + size_t(i) * ctsz, 2, limbs, scale, ct);
            Ciphertext res = gelu_v2(ct, *r->ctx, r->rlk, r->sk);
            *out_limbs = int(res.coeff_modulus_size());
            *out_scale = res.scale();
            memcpy(out + size_t(i) * 2 * res.coeff_modulus_size() * r->n, res.data(),
                   2 * res.coeff_modulus_size() * r->n * sizeof(uint64_t));
        }
        REF_CATCH(r)
    }

    // layernorm / layernorm2 (M/source/non_linear_func/layernorm.hpp:157-547); variant 1 or 2
    int ref_layernorm(void *h, int variant, const uint64_t *x, int num_ct, int limbs, double scale, const double *gamma,
                      const double *beta, const int *bias_vec, uint64_t *out, int *out_limbs, double *out_scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * r->n;
        vector<Ciphertext> xs(num_ct);
        for (int i = 0; i < num_ct; i++)
        {
            load_ct(*r, x + size_t(i) * ctsz, 2, limbs, scale, xs[i]);
        }
        vector<double> g(gamma, gamma + num_ct), b(beta, beta + num_ct);
        vector<int> bv(bias_vec, bias_vec + r->n / 2);
        vector<Ciphertext> res = variant == 1 ? layernorm(xs, g, b, bv, *r->ctx, r->rlk, r->sk)
                                              : layernorm2(xs, g, b, bv, *r->ctx, r->rlk, r->sk);
        for (int i = 0; i < num_ct; i++)
        {
            *out_limbs = int(res[i].coeff_modulus_size());
            *out_scale = res[i].scale();
            memcpy(out + size_t(i) * 2 * res[i].coeff_modulus_size() * r->n, res[i].data(),
                   2 * res[i].coeff_modulus_size() * r->n * sizeof(uint64_t));
        }
        REF_CATCH(r)
    }

    // ct_ct_matrix_mul_colpacking (which = 0) / _diagpacking (which = 1)
    // (M/source/matrix_mul/Ct_ct_matrix_mul.hpp:5-156)
    int ref_ct_ct_matmul(void *h, int which, const uint64_t *X, int nX, const uint64_t *W, int nW, int limbs,
                         double scale_X, double scale_W, int col_X, int row_X, int col_W, int row_W, int num_batch,
                         uint64_t *out, int *out_count, int *out_limbs, double *out_scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * r->n;
        vector<Ciphertext> xs(nX), ws(nW);
        for (int i = 0; i < nX; i++)
        {
            load_ct(*r, X + size_t(i) * ctsz, 2, limbs, scale_X, xs[i]);
        }
        for (int i = 0; i < nW; i++)
        {
            load_ct(*r, W + size_t(i) * ctsz, 2, limbs, scale_W, ws[i]);
        }
        vector<Ciphertext> res =
            which == 0 ? ct_ct_matrix_mul_colpacking(xs, ws, r->glk, r->rlk, *r->ctx, col_X, row_X, col_W, row_W, num_batch)
                       : ct_ct_matrix_mul_diagpacking(xs, ws, r->glk, r->rlk, *r->ctx, col_X, row_X, col_W, row_W,
                                                      num_batch);
        *out_count = int(res.size());
        for (size_t i = 0; i < res.size(); i++)
        {
            *out_limbs = int(res[i].coeff_modulus_size());
            *out_scale = res[i].scale();
            memcpy(out + i * 2 * res[i].coeff_modulus_size() * r->n, res[i].data(),
                   2 * res[i].coeff_modulus_size() * r->n * sizeof(uint64_t));
        }
        REF_CATCH(r)
    }

    // all Galois elements currently present (so the same key set can be uploaded to the GPU)
    int ref_galois_elts(void *h, uint32_t *out, int cap)
    {
        auto r = static_cast<Ref *>(h);
        int k = 0;
        auto &d = r->glk.data();
        for (size_t i = 0; i < d.size(); i++)
        {
            if (!d[i].empty() && k < cap)
            {
                out[k++] = uint32_t(2 * i + 1);
            }
        }
        return k;
    }

    // ---- wire format: Ciphertext::save / load with compr_mode_type::none (S/ciphertext.cpp:153-330) ----
    // *n_bytes in: capacity of buf, out: bytes written
    int ref_save_ciphertext(void *h, const uint64_t *ct_raw, int size, int limbs, double scale, uint8_t *buf,
                            int64_t *n_bytes)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        Ciphertext ct;
        load_ct(*r, ct_raw, size, limbs, scale, ct);
        const auto need = ct.save_size(compr_mode_type::none);
        if (need > *n_bytes)
        {
            throw invalid_argument("buffer too small");
        }
        *n_bytes = ct.save(reinterpret_cast<seal_byte *>(buf), size_t(need), compr_mode_type::none);
        REF_CATCH(r)
    }

    // loads with full validity checks (Ciphertext::load); fills the raw residues and the metadata
    int ref_load_ciphertext(void *h, const uint8_t *buf, int64_t n_bytes, uint64_t *ct_raw, int64_t cap_words, int *size,
                            int *limbs, double *scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        Ciphertext ct;
        ct.load(*r->ctx, reinterpret_cast<const seal_byte *>(buf), size_t(n_bytes));
        const int64_t words = int64_t(ct.size() * ct.coeff_modulus_size() * r->n);
        if (words > cap_words)
        {
            throw invalid_argument("buffer too small");
        }
        store_ct(*r, ct, ct_raw);
        *size = int(ct.size());
        *limbs = int(ct.coeff_modulus_size());
        *scale = ct.scale();
        REF_CATCH(r)
    }

    int ref_omp_threads()
    {
        return omp_get_max_threads();
    }

    // torch.distributed.run exports OMP_NUM_THREADS=1 to its children; the CPU baseline must still use every host core
    void ref_set_omp_threads(int n)
    {
        if (n > 0)
        {
            omp_set_num_threads(n);
        }
    }

    // ---- randomness and key wire format (for the facade's client-side pieces) --------------------------------
    // the first n bytes of SEAL's Blake2xb PRNG stream for a seed (S/randomgen.cpp:176-211)
    int ref_prng_bytes(const uint64_t *seed8, int64_t n, uint8_t *out)
    {
        prng_seed_type seed;
        copy_n(seed8, seed.size(), seed.begin());
        Blake2xbPRNG prng(seed);
        // in odd-sized pieces, to exercise the buffering
        int64_t done = 0, piece = 7;
        while (done < n)
        {
            int64_t take = min<int64_t>(piece, n - done);
            prng.generate(size_t(take), reinterpret_cast<seal_byte *>(out + done));
            done += take;
            piece = piece * 3 + 1;
        }
        return 0;
    }

    // util::sample_poly_uniform at the key level (S/util/rlwe.cpp:137-166): out [key limbs][N]
    int ref_sample_uniform(void *h, const uint64_t *seed8, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        prng_seed_type seed;
        copy_n(seed8, seed.size(), seed.begin());
        util::sample_poly_uniform(make_shared<Blake2xbPRNG>(seed), r->ctx->key_context_data()->parms(), out);
        REF_CATCH(r)
    }

    // what a client ships: kind 0 = RelinKeys, 1 = GaloisKeys for `steps` (+ conjugation), 2 = PublicKey;
    // seeded != 0: the Serializable<> form (the uniform halves replaced by PRNG seeds), compr_mode none.
    // The seeded keys are freshly generated (a Serializable cannot be kept); load them back with ref_load_keys.
    int ref_save_keys(void *h, int kind, int seeded, const int *steps, int n_steps, int with_conjugate, uint8_t *buf,
                      int64_t *n_bytes)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        stringstream ss;
        vector<uint32_t> elts;
        auto gt = r->ctx->key_context_data()->galois_tool();
        for (int i = 0; i < n_steps; i++)
        {
            elts.push_back(gt->get_elt_from_step(steps[i]));
        }
        if (with_conjugate)
        {
            elts.push_back(gt->get_elt_from_step(0));
        }
        if (kind == 0)
        {
            if (seeded)
            {
                r->keygen->create_relin_keys().save(ss, compr_mode_type::none);
            }
            else
            {
                r->rlk.save(ss, compr_mode_type::none);
            }
        }
        else if (kind == 1)
        {
            if (seeded)
            {
                r->keygen->create_galois_keys(elts).save(ss, compr_mode_type::none);
            }
            else
            {
                r->glk.save(ss, compr_mode_type::none);
            }
        }
        else
        {
            if (seeded)
            {
                r->keygen->create_public_key().save(ss, compr_mode_type::none);
            }
            else
            {
                r->pk.save(ss, compr_mode_type::none);
            }
        }
        string b = ss.str();
        if (int64_t(b.size()) > *n_bytes)
        {
            throw invalid_argument("buffer too small");
        }
        memcpy(buf, b.data(), b.size());
        *n_bytes = int64_t(b.size());
        REF_CATCH(r)
    }

    // SEAL's own loader (expands seeds); replaces the handle's relin / Galois / public key
    int ref_load_keys(void *h, int kind, const uint8_t *buf, int64_t n_bytes)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        if (kind == 0)
        {
            r->rlk.load(*r->ctx, reinterpret_cast<const seal_byte *>(buf), size_t(n_bytes));
            r->have_rlk = true;
        }
        else if (kind == 1)
        {
            r->glk.load(*r->ctx, reinterpret_cast<const seal_byte *>(buf), size_t(n_bytes));
        }
        else
        {
            r->pk.load(*r->ctx, reinterpret_cast<const seal_byte *>(buf), size_t(n_bytes));
            r->encryptor = make_unique<Encryptor>(*r->ctx, r->pk);
        }
        REF_CATCH(r)
    }

    // a seeded symmetric-key ciphertext stream (Encryptor::encrypt_symmetric(...).save), as a client ships inputs
    int ref_save_ciphertext_seeded(void *h, const uint64_t *pt_raw, int limbs, double scale, uint8_t *buf, int64_t *n_bytes)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        Plaintext pt;
        load_pt(*r, pt_raw, limbs, scale, pt);
        Encryptor enc(*r->ctx, r->sk);
        stringstream ss;
        enc.encrypt_symmetric(pt).save(ss, compr_mode_type::none);
        string b = ss.str();
        if (int64_t(b.size()) > *n_bytes)
        {
            throw invalid_argument("buffer too small");
        }
        memcpy(buf, b.data(), b.size());
        *n_bytes = int64_t(b.size());
        REF_CATCH(r)
    }

    // the reference's batch_input (M/source/matrix_mul/Batch_encode_encrypt.hpp:8-38): X [num_X][num_row][num_col]
    // -> num_col fresh ciphertexts [2][first limbs][n]
    int ref_batch_input(void *h, const double *X, int num_X, int num_row, int num_col, double scale, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
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
        vector<Ciphertext> res = batch_input(Xv, num_X, num_row, num_col, scale, *r->ctx, r->pk);
        for (size_t i = 0; i < res.size(); i++)
        {
            store_ct(*r, res[i], out + i * 2 * r->n_data_limbs * r->n);
        }
        REF_CATCH(r)
    }

    // ---- the reference's Bootstrapper (M/source/bootstrapping/Bootstrapper.cpp), compiled unmodified ---------------
    // Construction as in M/test/test_full_scheme.hpp:413-433: Bootstrapper(loge, logn, logNh, L, final_scale, K, deg,
    // scale_factor, inverse_deg, ...) then prepare_mod_polynomial() — the reference's own multi-interval Remez
    // (common/Remez.cpp:557-586) at 1000+ bits.  Needs the relinearization key (ref_make_relin_key) first.
    int ref_boot_create(void *h, int loge, int logn, int logNh, int L, double final_scale, int boundary_K, int deg,
                        int scale_factor, int inverse_deg)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        if (!r->have_rlk)
        {
            throw logic_error("ref_boot_create: make the relinearization key first");
        }
        CoutMute mute;
        r->boot = make_unique<Bootstrapper>(loge, logn, logNh, L, final_scale, boundary_K, deg, scale_factor, inverse_deg,
                                            *r->ctx, *r->keygen, *r->encoder, *r->encryptor, *r->decryptor, *r->evaluator,
                                            r->rlk, r->glk);
        r->boot->prepare_mod_polynomial();
        REF_CATCH(r)
    }

    // Rotation steps the driver asks keys for (test_full_scheme.hpp:436-443): 0, 2^i for i < logN - 1, then
    // addLeftRotKeys_Linear_to_vector_3.  Returns the count.
    int ref_boot_steps(void *h, int *out, int cap)
    {
        auto r = static_cast<Ref *>(h);
        if (!r->boot)
        {
            return -1;
        }
        vector<int> steps;
        steps.push_back(0);
        int log_n = 0;
        while ((size_t(1) << log_n) < r->n)
        {
            log_n++;
        }
        for (int i = 0; i < log_n - 1; i++)
        {
            steps.push_back(1 << i);
        }
        r->boot->addLeftRotKeys_Linear_to_vector_3(steps);
        for (size_t i = 0; i < steps.size() && int(i) < cap; i++)
        {
            out[i] = steps[i];
        }
        return int(steps.size());
    }

    // slot_vec.push_back(logn); generate_LT_coefficient_3()   (test_full_scheme.hpp:445-448)
    int ref_boot_prepare(void *h)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        if (!r->boot)
        {
            throw logic_error("no bootstrapper");
        }
        CoutMute mute;
        r->boot->slot_vec.push_back(r->boot->logn);
        r->boot->generate_LT_coefficient_3();
        REF_CATCH(r)
    }

    // The EvalMod polynomial the reference generated: Chebyshev coefficients (variable x / K) of the cosine after the
    // inverse-sine constant has been folded in (ModularReducer.cpp:34-48), and that constant.
    int ref_boot_polynomial(void *h, double *cheb, int cap, double *scale_inverse_coeff)
    {
        auto r = static_cast<Ref *>(h);
        if (!r->boot)
        {
            return -1;
        }
        auto &p = r->boot->mod_reducer->sin_cos_polynomial;
        for (long i = 0; i <= p.deg && i < cap; i++)
        {
            cheb[i] = to_double(p.chebcoeff[i]);
        }
        *scale_inverse_coeff = r->boot->mod_reducer->scale_inverse_coeff;
        return int(p.deg);
    }

    // Bootstrapper::bootstrap_3 (Bootstrapper.cpp:3496-3502) on one ciphertext at the lowest level.
    int ref_boot_bootstrap_3(void *h, const uint64_t *in, double scale, uint64_t *out, int *out_limbs, double *out_scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        if (!r->boot)
        {
            throw logic_error("no bootstrapper");
        }
        CoutMute mute;
        Ciphertext ct, res;
        load_ct(*r, in, 2, 1, scale, ct);
        r->boot->bootstrap_3(res, ct);
        *out_limbs = int(res.coeff_modulus_size());
        *out_scale = res.scale();
        store_ct(*r, res, out);
        REF_CATCH(r)
    }

    // The phases of bootstrap_full_3 (Bootstrapper.cpp:3231-3251) one by one, for per-phase comparisons:
    // phase 0: modraise_inplace + scale = q0      in: 1 ct @1 limb           out: 1 ct @ all data limbs
    // phase 1: coefftoslot_full_3                  in: 1 ct                   out: 2 cts (real, imaginary halves)
    // phase 2: modular_reduction                   in: 1 ct                   out: 1 ct
    // phase 3: slottocoeff_full_3 + final scale    in: 2 cts                  out: 1 ct
    int ref_boot_phase(void *h, int phase, const uint64_t *in, int limbs, double scale, uint64_t *out, int *out_limbs,
                       double *out_scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        if (!r->boot)
        {
            throw logic_error("no bootstrapper");
        }
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * r->n;
        Ciphertext a, b, o1, o2;
        load_ct(*r, in, 2, limbs, scale, a);
        if (phase == 0)
        {
            r->boot->modraise_inplace(a);
            const auto &modulus = r->ctx->first_context_data()->parms().coeff_modulus();
            a.scale() = double(modulus[0].value());
            o1 = a;
        }
        else if (phase == 1)
        {
            r->boot->coefftoslot_full_3(o1, o2, a);
        }
        else if (phase == 2)
        {
            r->boot->mod_reducer->modular_reduction(o1, a);
        }
        else if (phase == 3)
        {
            load_ct(*r, in + ctsz, 2, limbs, scale, b);
            r->boot->slottocoeff_full_3(o1, a, b);
            o1.scale() = r->boot->final_scale;
        }
        else
        {
            throw invalid_argument("phase");
        }
        *out_limbs = int(o1.coeff_modulus_size());
        *out_scale = o1.scale();
        store_ct(*r, o1, out);
        if (phase == 1)
        {
            store_ct(*r, o2, out + size_t(2) * o1.coeff_modulus_size() * r->n);
        }
        REF_CATCH(r)
    }

    // softmax_boot (M/source/non_linear_func/softmax.hpp:308-581) on `num` ciphertexts [num][2][limbs][n]
    int ref_softmax_boot(void *h, const uint64_t *x, int num, int limbs, double scale, const int *bias_vec, int input_num,
                         int iter, int layer_id, uint64_t *out, int *out_limbs, double *out_scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        if (!r->boot)
        {
            throw logic_error("no bootstrapper");
        }
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * r->n;
        vector<Ciphertext> xs(num);
        for (int i = 0; i < num; i++)
        {
            load_ct(*r, x + size_t(i) * ctsz, 2, limbs, scale, xs[i]);
        }
        vector<int> bv(bias_vec, bias_vec + r->n / 2);
        vector<Ciphertext> res = softmax_boot(xs, bv, input_num, *r->ctx, r->rlk, iter, r->sk, *r->boot, layer_id);
        for (size_t i = 0; i < res.size(); i++)
        {
            *out_limbs = int(res[i].coeff_modulus_size());
            *out_scale = res[i].scale();
            memcpy(out + i * 2 * res[i].coeff_modulus_size() * r->n, res[i].data(),
                   2 * res[i].coeff_modulus_size() * r->n * sizeof(uint64_t));
        }
        REF_CATCH(r)
    }

    // single_att_block (M/source/att_block/single_att_block.hpp:10-207): one attention head.
    // X [num_col][2][limbs][n]; WQ/WK/WV row-major [num_col][col_W]; bQ/bK/bV [col_W].
    int ref_single_att_block(void *h, const uint64_t *X, int num_col, int limbs, double scale, const double *WQ,
                             const double *WK, const double *WV, int col_W, const double *bQ, const double *bK,
                             const double *bV, const int *bias_vec, int input_num, int num_batch, int iter, int layer_id,
                             uint64_t *out, int *out_count, int *out_limbs, double *out_scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        if (!r->boot)
        {
            throw logic_error("no bootstrapper");
        }
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * r->n;
        vector<Ciphertext> xs(num_col);
        for (int i = 0; i < num_col; i++)
        {
            load_ct(*r, X + size_t(i) * ctsz, 2, limbs, scale, xs[i]);
        }
        auto mat = [&](const double *W) {
            vector<vector<double>> m(num_col, vector<double>(col_W));
            for (int i = 0; i < num_col; i++)
            {
                for (int j = 0; j < col_W; j++)
                {
                    m[i][j] = W[size_t(i) * col_W + j];
                }
            }
            return m;
        };
        vector<double> vq(bQ, bQ + col_W), vk(bK, bK + col_W), vv(bV, bV + col_W);
        vector<int> bv(bias_vec, bias_vec + r->n / 2);
        vector<Ciphertext> res = single_att_block(xs, mat(WQ), mat(WK), mat(WV), vq, vk, vv, bv, input_num, *r->ctx, r->rlk,
                                                  r->glk, *r->boot, num_batch, r->sk, iter, layer_id);
        *out_count = int(res.size());
        for (size_t i = 0; i < res.size(); i++)
        {
            *out_limbs = int(res[i].coeff_modulus_size());
            *out_scale = res[i].scale();
            memcpy(out + i * 2 * res[i].coeff_modulus_size() * r->n, res[i].data(),
                   2 * res[i].coeff_modulus_size() * r->n * sizeof(uint64_t));
        }
        REF_CATCH(r)
    }

    // exp (which = 0, softmax.hpp:9-47) / inverse (which = 1, softmax.hpp:49-82) of the reference header on `count`
    // ciphertexts [count][2][limbs][n]
    int ref_exp_inverse(void *h, int which, const uint64_t *x, int count, int limbs, double scale, int iter, uint64_t *out,
                        int *out_limbs, double *out_scale)
    {
        auto r = static_cast<Ref *>(h);
        REF_TRY(r)
        CoutMute mute;
        size_t ctsz = size_t(2) * limbs * r->n;
        for (int i = 0; i < count; i++)
        {
            Ciphertext ct;
            load_ct(*r, x + size_t(i) * ctsz, 2, limbs, scale, ct);
            Ciphertext res = which == 0 ? exp(ct, *r->ctx, r->rlk) : inverse(ct, *r->ctx, r->rlk, iter);
            *out_limbs = int(res.coeff_modulus_size());
            *out_scale = res.scale();
            memcpy(out + size_t(i) * 2 * res.coeff_modulus_size() * r->n, res.data(),
                   2 * res.coeff_modulus_size() * r->n * sizeof(uint64_t));
        }
        REF_CATCH(r)
    }
}
