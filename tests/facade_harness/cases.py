"""Parity cases shared by tests/test_facade.py (facade over the CPU test double) and tests/test_gpu_zz_facade.py
(facade over libmoai_b200.so): the reference's UNMODIFIED module headers run twice — against the reference's
real SEAL (oracle.SealRef) and against the facade (facade_harness.FacadeDriver) — on identical SEAL-generated
keys and encryptions; every residue of every output must match."""
import numpy as np

from oracle import (OP_ADD, OP_SUB, OP_MULTIPLY, OP_SQUARE, OP_RELINEARIZE, OP_RESCALE, OP_MOD_SWITCH, OP_ROTATE,
                    OP_CONJUGATE, OP_MULTIPLY_PLAIN, OP_ADD_PLAIN, OP_SUB_PLAIN, OP_NEGATE, OP_MULTIPLY_CONST,
                    OP_ADD_CONST, OP_DOUBLE, OP_ADD_REDUCED_ERROR, OP_SUB_REDUCED_ERROR, OP_MULTIPLY_REDUCED_ERROR,
                    OP_MULTIPLY_VECTOR_REDUCED_ERROR)

SCALE = 2.0 ** 30


def encrypt_batch(r, rng, count, limbs, sigma=0.5, mask=None, scale=SCALE):
    top = r.kl - 1
    cts, vals = [], []
    for _ in range(count):
        v = rng.normal(size=r.n // 2) * sigma
        if mask is not None:
            v = v * mask
        ct = r.encrypt(r.encode_real(v, scale, top), top, scale).reshape(2, top, r.n)[:, :limbs, :]
        cts.append(np.ascontiguousarray(ct))
        vals.append(v)
    return np.stack(cts), np.stack(vals)


def same(a, b):
    """(raw, size?, limbs, scale) tuples from both sides: metadata equal, residues bit-identical."""
    assert a[1:] == b[1:], (a[1:], b[1:])
    assert a[0].shape == b[0].shape and (a[0] == b[0]).all()


def case_context(r, d):
    """CoeffModulus::Create and the modulus chain (S/modulus.cpp:143-184, S/context.cpp:422-524)."""
    assert d.kl == r.kl and (d.q == r.q).all()
    for limbs in range(1, r.kl):
        assert d.chain_index(limbs) == limbs - 1
    assert d.chain_index(r.kl) < 0      # only the key level has all primes: not a data level


def case_evaluator_ops(r, d, rng, limbs=3):
    """One call of every Evaluator method family the MOAI sources use (SURVEY section 8(b) census)."""
    x, _ = encrypt_batch(r, rng, 2, limbs)
    a, b = x[0].reshape(-1), x[1].reshape(-1)
    n = r.n
    for op in (OP_ADD, OP_SUB, OP_MULTIPLY):
        same(d.eval(op, a, 2, limbs, SCALE, b=b, size_b=2, limbs_b=limbs, scale_b=SCALE),
             r.eval(op, a, 2, limbs, SCALE, b=b, size_b=2, limbs_b=limbs, scale_b=SCALE))
    for op in (OP_SQUARE, OP_RESCALE, OP_MOD_SWITCH, OP_NEGATE, OP_DOUBLE, OP_CONJUGATE):
        same(d.eval(op, a, 2, limbs, SCALE), r.eval(op, a, 2, limbs, SCALE))
    # size-3: add of mixed sizes, relinearize, rescale of a size-3 ciphertext
    sq = r.eval(OP_SQUARE, a, 2, limbs, SCALE)
    same(d.eval(OP_RELINEARIZE, sq[0], 3, limbs, sq[3]), r.eval(OP_RELINEARIZE, sq[0], 3, limbs, sq[3]))
    same(d.eval(OP_RESCALE, sq[0], 3, limbs, sq[3]), r.eval(OP_RESCALE, sq[0], 3, limbs, sq[3]))
    sq2 = r.eval(OP_MULTIPLY, a, 2, limbs, SCALE, b=b, size_b=2, limbs_b=limbs, scale_b=SCALE)
    same(d.eval(OP_ADD, sq[0], 3, limbs, sq[3], b=sq2[0], size_b=3, limbs_b=limbs, scale_b=sq2[3]),
         r.eval(OP_ADD, sq[0], 3, limbs, sq[3], b=sq2[0], size_b=3, limbs_b=limbs, scale_b=sq2[3]))
    # rotations: keys present (1, -1, 256), NAF fallback (3 = 4 - 1, 7 = 8 - 1 needs 8: only where present)
    for steps in (1, -1, 256, 3, 5):
        same(d.eval(OP_ROTATE, a, 2, limbs, SCALE, iarg=steps), r.eval(OP_ROTATE, a, 2, limbs, SCALE, iarg=steps))
    # plaintext operands: vector encodings produced by the facade's encoder must be SEAL's residues
    v = rng.normal(size=n // 2)
    pt = d.encode_real(v, SCALE, limbs)
    assert (pt == r.encode_real(v, SCALE, limbs)).all()
    z = rng.normal(size=n // 4) + 1j * rng.normal(size=n // 4)          # fewer values than slots
    assert (d.encode(z, SCALE, limbs) == r.encode(z, SCALE, limbs)).all()
    for op in (OP_MULTIPLY_PLAIN, OP_ADD_PLAIN, OP_SUB_PLAIN):
        same(d.eval(op, a, 2, limbs, SCALE, b=pt, size_b=1, limbs_b=limbs, scale_b=SCALE),
             r.eval(op, a, 2, limbs, SCALE, b=pt, size_b=1, limbs_b=limbs, scale_b=SCALE))
    # the fork's additions (S/evaluator.cpp:395-594)
    for op, val in ((OP_MULTIPLY_CONST, -0.37), (OP_ADD_CONST, 2.5)):
        same(d.eval(op, a, 2, limbs, SCALE, darg=val), r.eval(op, a, 2, limbs, SCALE, darg=val))
    vv = rng.normal(size=n // 2) + 1j * rng.normal(size=n // 2)
    same(d.eval(OP_MULTIPLY_VECTOR_REDUCED_ERROR, a, 2, limbs, SCALE, varg=vv),
         r.eval(OP_MULTIPLY_VECTOR_REDUCED_ERROR, a, 2, limbs, SCALE, varg=vv))
    lo = np.ascontiguousarray(x[1][:, : limbs - 1, :]).reshape(-1)       # one level lower, slightly different scale
    for op in (OP_ADD_REDUCED_ERROR, OP_SUB_REDUCED_ERROR, OP_MULTIPLY_REDUCED_ERROR):
        for (p, lp, sp, q_, lq, sq_) in ((a, limbs, SCALE, lo, limbs - 1, SCALE * 1.0001),
                                          (lo, limbs - 1, SCALE * 1.0001, a, limbs, SCALE),
                                          (a, limbs, SCALE, b, limbs, SCALE * 1.0001)):
            same(d.eval(op, p, 2, lp, sp, b=q_, size_b=2, limbs_b=lq, scale_b=sq_),
                 r.eval(op, p, 2, lp, sp, b=q_, size_b=2, limbs_b=lq, scale_b=sq_))


def case_errors(r, d, rng, limbs=3):
    """SEAL's exception rules survive the C ABI: same exception class for the same misuse."""
    import pytest
    from . import FacadeError
    x, _ = encrypt_batch(r, rng, 2, limbs)
    a = x[0].reshape(-1)
    lo = np.ascontiguousarray(x[1][:, : limbs - 1, :]).reshape(-1)
    cases = [
        (dict(op=OP_ADD, a=a, size_a=2, limbs_a=limbs, scale_a=SCALE, b=lo, size_b=2, limbs_b=limbs - 1, scale_b=SCALE),
         "invalid_argument"),                                             # parms mismatch, S/evaluator.cpp:166-170
        (dict(op=OP_ADD, a=a, size_a=2, limbs_a=limbs, scale_a=SCALE, b=a, size_b=2, limbs_b=limbs, scale_b=SCALE * 2),
         "invalid_argument"),                                             # scale mismatch, :171-174
        (dict(op=OP_ROTATE, a=a, size_a=2, limbs_a=limbs, scale_a=SCALE, iarg=1024), "invalid_argument"),   # no key
    ]
    last = np.ascontiguousarray(x[0][:, :1, :]).reshape(-1)
    cases.append((dict(op=OP_RESCALE, a=last, size_a=2, limbs_a=1, scale_a=SCALE), "invalid_argument"))   # end of chain
    for kw, kind in cases:
        with pytest.raises(RuntimeError):
            r.eval(kw["op"], kw["a"], kw["size_a"], kw["limbs_a"], kw["scale_a"],
                   **{k: v for k, v in kw.items() if k not in ("op", "a", "size_a", "limbs_a", "scale_a")})
        with pytest.raises(FacadeError, match=kind):
            d.eval(kw["op"], kw["a"], kw["size_a"], kw["limbs_a"], kw["scale_a"],
                   **{k: v for k, v in kw.items() if k not in ("op", "a", "size_a", "limbs_a", "scale_a")})


def case_decrypt_decode(r, d, rng, limbs=3):
    """Decryptor::decrypt bit-exact; CKKSEncoder::decode (host CRT + FFT) within 1e-9 of SEAL's decode."""
    x, vals = encrypt_batch(r, rng, 1, limbs)
    a = x[0].reshape(-1)
    pt = d.decrypt(a, 2, limbs, SCALE)
    assert (pt == r.decrypt(a, 2, limbs, SCALE)).all()
    sq = r.eval(OP_SQUARE, a, 2, limbs, SCALE)
    assert (d.decrypt(sq[0], 3, limbs, sq[3]) == r.decrypt(sq[0], 3, limbs, sq[3])).all()
    got, exp = d.decode(pt, limbs, SCALE), r.decode(pt, limbs, SCALE)
    assert np.abs(got - exp).max() < 1e-9
    assert np.abs(got.real - vals[0]).max() < 1e-4


def case_ct_pt(r, d, rng, variant, limbs=2):
    """ct_pt_matrix_mul_wo_pre / _large / _w_mask (M/source/matrix_mul/Ct_pt_matrix_mul.hpp:4-170)."""
    K = 5
    Cc = 4 if variant == 0 else 128         # the _large / _w_mask variants loop over 128 column groups
    X, _ = encrypt_batch(r, rng, K, limbs)
    W = rng.normal(size=(K, Cc)) * 0.3
    mask = None
    if variant == 2:
        mask = np.zeros(r.n // 2, dtype=np.int32)
        mask[::16][:5] = 1
    exp, _ = r.ct_pt_matmul(variant, X.reshape(-1), W, mask, K, Cc, limbs, SCALE)
    got, _ = d.ct_pt_matmul(variant, X.reshape(-1), W, mask, K, Cc, limbs, SCALE)
    assert (got == exp).all()


def case_gelu(r, d, rng, limbs=9):
    x, _ = encrypt_batch(r, rng, 2, limbs, sigma=1.0)
    same(d.gelu_v2(x.reshape(-1), 2, limbs, SCALE), r.gelu_v2(x.reshape(-1), 2, limbs, SCALE))


def case_layernorm(r, d, rng, variant, limbs=21):
    num_ct = 768                            # the reference hard-codes 768 = 48 x 16 (layernorm.hpp:242-262)
    mask = np.zeros(r.n // 2, dtype=np.int32)
    mask[::16][:5] = 1
    x16, _ = encrypt_batch(r, rng, 16, limbs, sigma=0.3, mask=mask)
    x = np.ascontiguousarray(np.tile(x16, (num_ct // 16, 1, 1, 1)))
    gamma, beta = rng.normal(size=num_ct), rng.normal(size=num_ct) * 0.1
    exp = r.layernorm(variant, x.reshape(-1), num_ct, limbs, SCALE, gamma, beta, mask)
    got = d.layernorm(variant, x.reshape(-1), num_ct, limbs, SCALE, gamma, beta, mask, want_printed=True)
    same(got[:3], exp)
    # the module's debug prints (decrypt + decode through the facade) carry real numbers
    assert "decrypt of var" in got[3] and "nan" not in got[3].lower()


def case_ct_ct(r, d, rng, which):
    if which == 0:
        limbs, col_X, row_X, nb = 4, 3, 8, 16
        X, _ = encrypt_batch(r, rng, col_X, limbs)
        W, _ = encrypt_batch(r, rng, col_X, limbs)
        args = (0, X.reshape(-1), col_X, W.reshape(-1), col_X, limbs, SCALE, SCALE, col_X, row_X, col_X, row_X, nb)
    else:
        limbs, col_X, col_W, nb = 3, 8, 3, 16
        X, _ = encrypt_batch(r, rng, col_X, limbs)
        W, _ = encrypt_batch(r, rng, col_W, limbs)
        args = (1, X.reshape(-1), col_X, W.reshape(-1), col_W, limbs, SCALE, SCALE, col_X, col_X, col_W, col_X, nb)
    same(d.ct_ct_matmul(*args), r.ct_ct_matmul(*args))


def replay_exp(r, x, limbs):
    """softmax.hpp:9-47 as a sequence of real-SEAL Evaluator calls (softmax.hpp itself needs NTL to compile
    against stock SEAL; against the facade it compiles as it is)."""
    c, _, l, s = r.eval(OP_MULTIPLY_PLAIN, x, 2, limbs, SCALE, b=r.encode_scalar(0.0078125, SCALE, limbs), size_b=1,
                        limbs_b=limbs, scale_b=SCALE)
    c, _, l, s = r.eval(OP_RESCALE, c, 2, l, s)
    c, _, l, s = r.eval(OP_ADD_PLAIN, c, 2, l, s, b=r.encode_scalar(1.0, s, l), size_b=1, limbs_b=l, scale_b=s)
    for _ in range(7):
        c, _, l, s = r.eval(OP_SQUARE, c, 2, l, s)
        c, _, l, s = r.eval(OP_RELINEARIZE, c, 3, l, s)
        c, _, l, s = r.eval(OP_RESCALE, c, 2, l, s)
    return c, l, s


def replay_inverse(r, x, limbs, iters):
    """softmax.hpp:49-82 on real SEAL."""
    one = r.encode_scalar(1.0, SCALE, limbs)
    y, _, l, s = r.eval(OP_SUB_PLAIN, x, 2, limbs, SCALE, b=one, size_b=1, limbs_b=limbs, scale_b=SCALE)
    y, _, l, s = r.eval(OP_NEGATE, y, 2, l, s)
    res, _, lr, sr = r.eval(OP_ADD_PLAIN, y, 2, l, s, b=one, size_b=1, limbs_b=l, scale_b=s)
    for _ in range(iters):
        y, _, l, s = r.eval(OP_SQUARE, y, 2, l, s)
        y, _, l, s = r.eval(OP_RELINEARIZE, y, 3, l, s)
        y, _, l, s = r.eval(OP_RESCALE, y, 2, l, s)
        tmp, _, lt, st = r.eval(OP_ADD_PLAIN, y, 2, l, s, b=r.encode_scalar(1.0, s, l), size_b=1, limbs_b=l, scale_b=s)
        while lr > lt:
            res, _, lr, sr = r.eval(OP_MOD_SWITCH, res, 2, lr, sr)
        res, _, lr, sr = r.eval(OP_MULTIPLY, res, 2, lr, sr, b=tmp, size_b=2, limbs_b=lt, scale_b=st)
        res, _, lr, sr = r.eval(OP_RELINEARIZE, res, 3, lr, sr)
        res, _, lr, sr = r.eval(OP_RESCALE, res, 2, lr, sr)
    return res, lr, sr


def case_exp_inverse(r, d, rng, limbs=12):
    """exp / inverse of the reference's softmax.hpp, compiled unchanged against the facade."""
    x, _ = encrypt_batch(r, rng, 1, limbs, sigma=1.0)
    a = x[0].reshape(-1)
    same(d.exp(a, limbs, SCALE), replay_exp(r, a, limbs))
    same(d.inverse(a, limbs, SCALE, 3), replay_inverse(r, a, limbs, 3))


def case_value_semantics(r, d, rng, limbs=3):
    """destination == operand, deep copies, scalar vs vector encodings of a constant, plaintext mod switch."""
    x, _ = encrypt_batch(r, rng, 2, limbs)
    assert d.alias_checks(x[0].reshape(-1), x[1].reshape(-1), limbs, SCALE) == 0


# ---- client-side pieces: PRNG, seeded keys, wire format, Encryptor ----
def case_prng(make_ref, make_drv):
    """SEAL's Blake2xb PRNG stream and its uniform sampler with rejection, byte for byte."""
    r, d = make_ref(5), make_drv(5)
    rng = np.random.default_rng(100)
    for _ in range(3):
        seed = rng.integers(0, 2 ** 63, 8, dtype=np.uint64)
        assert (d.prng_bytes(seed, 10007) == r.prng_bytes(seed, 10007)).all()      # crosses two 4096-byte refills
        assert (d.sample_uniform(seed) == r.sample_uniform(seed)).all()


def case_key_wire_format(make_ref, make_drv):
    """RelinKeys / GaloisKeys / PublicKey streams of the stock library, plain and SEEDED (the form a client ships:
    half the bytes), loaded by the facade == loaded by SEAL itself."""
    seed = 9
    r, d = make_ref(seed), make_drv(seed)
    r.make_relin_key()
    r.make_galois_keys([1, -2], conjugate=True)
    digits = r.kl - 1
    for seeded in (False, True):
        blob = r.save_keys(0, seeded)
        r.load_keys(0, blob)                       # SEAL's own loader (expands the seeds)
        assert d.load_keys(0, blob) == (digits if seeded else 0)
        assert (d.export_key(0) == r.export_relin_key()).all()
        blob = r.save_keys(1, seeded, steps=[1, -2], conjugate=True)
        r.load_keys(1, blob)
        elts = r.galois_elts()
        assert len(elts) == 3
        assert d.load_keys(1, blob) == (3 * digits if seeded else 0)
        for e in elts:
            assert (d.export_key(1, e) == r.export_galois_key(e)).all()
        blob = r.save_keys(2, seeded)
        r.load_keys(2, blob)
        d.load_keys(2, blob)
        assert (d.export_key(2) == r.public_key()).all()
    # a seeded stream is about half the size of the plain one
    assert len(r.save_keys(0, True)) < 0.51 * len(r.save_keys(0, False)) + 4096


def case_encrypt(make_ref, make_drv):
    """Encryptor::encrypt with the same PRNG seed: SEAL's ciphertext bit for bit (u, e0, e1 from the same stream,
    arithmetic one level up, division by the last prime with rounding)."""
    seed = 13
    r, d = make_ref(seed), make_drv(seed)
    d.set_public_key(r.public_key())
    rng = np.random.default_rng(101)
    top = r.kl - 1
    for limbs in (top, top - 1, 1):
        v = rng.normal(size=r.n // 2)
        pt = r.encode_real(v, SCALE, limbs)
        got, exp = d.encrypt(pt, limbs, SCALE), r.encrypt(pt, limbs, SCALE)
        assert (got == exp).all()
    # and it decrypts
    assert np.abs(r.decode(r.decrypt(got, 2, 1, SCALE), 1, SCALE).real - v).max() < 1e-4


def case_ciphertext_wire_format(make_ref, make_drv):
    seed = 17
    r, d = make_ref(seed), make_drv(seed)
    rng = np.random.default_rng(102)
    top = r.kl - 1
    pt = r.encode_real(rng.normal(size=r.n // 2), SCALE, top)
    ct = r.encrypt(pt, top, SCALE)
    blob = r.save_ciphertext(ct, 2, top, SCALE)
    assert d.save_ciphertext(ct, 2, top, SCALE) == blob                      # byte-identical to Ciphertext::save
    got = d.load_ciphertext(blob, 3 * r.kl * r.n)
    assert (got[0] == ct).all() and got[1:] == (2, top, SCALE)
    sq = r.eval(OP_SQUARE, ct, 2, top, SCALE)                                # size 3, squared scale, lower level
    sq = r.eval(OP_MOD_SWITCH, sq[0], 3, top, sq[3])
    blob = r.save_ciphertext(sq[0], 3, sq[2], sq[3])
    assert d.save_ciphertext(sq[0], 3, sq[2], sq[3]) == blob
    # the seeded form of a symmetric-key encryption (what a client ships as input): c1 expanded from the seed
    blob = r.save_ciphertext_seeded(pt, top, SCALE)
    exp = r.load_ciphertext(blob)
    got = d.load_ciphertext(blob, 3 * r.kl * r.n)
    assert (got[0] == exp[0]).all() and got[1:] == tuple(exp[1:])
    assert len(blob) < 0.51 * len(r.save_ciphertext(ct, 2, top, SCALE)) + 256


def case_batch_input(make_ref, make_drv):
    """The reference's batch_input (M/source/matrix_mul/Batch_encode_encrypt.hpp:8-38), unchanged, through the
    facade's CKKSEncoder + Encryptor == the same header on real SEAL."""
    seed = 19
    r, d = make_ref(seed), make_drv(seed)
    d.set_public_key(r.public_key())
    rng = np.random.default_rng(103)
    X = rng.normal(size=(4, 8, 3))
    assert (d.batch_input(X, SCALE) == r.batch_input(X, SCALE)).all()


def case_keygen(make_ref, make_drv):
    """KeyGenerator with the same PRNG seed: SEAL's secret key, public key, relinearisation key and Galois keys bit
    for bit; the keys work (a rotation with the generated key equals SEAL's)."""
    seed = 23
    r, d = make_ref(seed), make_drv(seed)
    r.make_relin_key()
    r.make_galois_keys([1, -2, 16], conjugate=True)
    d.keygen(steps=[1, -2, 16], conjugate=True)
    assert (d.export_secret() == r.secret_key()).all()
    assert (d.export_key(2) == r.public_key()).all()
    assert (d.export_key(0) == r.export_relin_key()).all()
    elts = r.galois_elts()
    assert len(elts) == 4
    for e in elts:
        assert (d.export_key(1, e) == r.export_galois_key(e)).all()
    rng = np.random.default_rng(104)
    x, _ = encrypt_batch(r, rng, 1, 2)
    a = x[0].reshape(-1)
    same(d.eval(OP_ROTATE, a, 2, 2, SCALE, iarg=16), r.eval(OP_ROTATE, a, 2, 2, SCALE, iarg=16))
    sq = r.eval(OP_SQUARE, a, 2, 2, SCALE)
    same(d.eval(OP_RELINEARIZE, sq[0], 3, 2, sq[3]), r.eval(OP_RELINEARIZE, sq[0], 3, 2, sq[3]))
    # the default key set of create_galois_keys(GaloisKeys&): conjugation and +-2^k (S/util/galois.cpp:106-131)
    d2 = make_drv(seed)
    d2.keygen()
    n2 = 2 * r.n
    want = {n2 - 1}
    pos, neg = 5, pow(5, -1, n2)
    for _ in range(r.log_n - 1):
        want |= {pos, neg}
        pos, neg = pos * pos % n2, neg * neg % n2
    assert all(d2.has_galois(e) for e in want) and not d2.has_galois(3 if 3 not in want else 7)


def case_keygen_sparse(make_ref_sparse, make_drv_sparse):
    """The fork's sparse ternary secret (Hamming weight h, M/test/test_full_scheme.hpp:366): same key as SEAL."""
    seed = 29
    r, d = make_ref_sparse(seed), make_drv_sparse(seed)
    d.keygen(steps=[1])
    sk = d.export_secret()
    assert (sk == r.secret_key()).all()
