"""GPU parity of the MOAI module layer against the reference's UNMODIFIED
module headers (gelu_others.hpp, layernorm.hpp, Ct_ct_matrix_mul.hpp) compiled against the
reference's real SEAL (oracle/_ref), on SEAL-generated keys and SEAL encryptions.  Bit-exact:
every ciphertext residue of every output must match."""
import math

import numpy as np
import pytest

from oracle import (OP_ADD_PLAIN, OP_MULTIPLY, OP_MULTIPLY_PLAIN, OP_NEGATE, OP_RELINEARIZE, OP_RESCALE, OP_SQUARE,
                    OP_SUB_PLAIN, OP_MOD_SWITCH)

pytestmark = pytest.mark.gpu
SCALE = 2.0 ** 30


def encrypt_batch(r, rng, count, limbs, sigma=0.5, mask=None):
    top = r.kl - 1
    cts, vals = [], []
    for _ in range(count):
        v = rng.normal(size=r.n // 2) * sigma
        if mask is not None:
            v = v * mask
        ct = r.encrypt(r.encode_real(v, SCALE, top), top, SCALE).reshape(2, top, r.n)[:, :limbs, :]
        cts.append(np.ascontiguousarray(ct))
        vals.append(v)
    return np.stack(cts), np.stack(vals)


def test_gelu_v2_vs_reference(pkg, backend_deep, sealref_deep, keys_deep):
    r, be = sealref_deep, backend_deep
    rng = np.random.default_rng(1)
    limbs = 9
    x, vals = encrypt_batch(r, rng, 3, limbs, sigma=1.0)
    exp, el, es = r.gelu_v2(x.reshape(-1), 3, limbs, SCALE)
    got, gs = be.gelu_v2(keys_deep, pkg.to_device(x), SCALE)
    assert got.shape[2] == el == 2 and gs == es
    assert (pkg.to_host(got).reshape(-1) == exp).all()


@pytest.mark.parametrize("variant", [1, 2])
def test_layernorm_vs_reference(pkg, backend_deep, sealref_deep, keys_deep, variant):
    r, be = sealref_deep, backend_deep
    rng = np.random.default_rng(10 + variant)
    limbs, num_ct = 21, 768                 # the reference hard-codes 768 = 48 x 16 (layernorm.hpp:242-262)
    mask = np.zeros(r.n // 2, dtype=np.int32)
    mask[::16][:5] = 1                      # 5 valid token slots, like the reference run
    x16, _ = encrypt_batch(r, rng, 16, limbs, sigma=0.3, mask=mask)
    x = np.ascontiguousarray(np.tile(x16, (num_ct // 16, 1, 1, 1)))   # 16 distinct encryptions, tiled
    gamma, beta = rng.normal(size=num_ct), rng.normal(size=num_ct) * 0.1
    exp, el, es = r.layernorm(variant, x.reshape(-1), num_ct, limbs, SCALE, gamma, beta, mask)
    got, gs = be.layernorm(keys_deep, pkg.to_device(x), SCALE, gamma, beta, mask, variant=variant)
    assert got.shape[2] == el == 1 and gs == es
    assert (pkg.to_host(got).reshape(-1) == exp).all()


def test_ct_ct_colpacking_vs_reference(pkg, backend_deep, sealref_deep, keys_deep):
    r, be = sealref_deep, backend_deep
    rng = np.random.default_rng(20)
    limbs, col_X, row_X, nb = 4, 3, 8, 16
    X, _ = encrypt_batch(r, rng, col_X, limbs)
    W, _ = encrypt_batch(r, rng, col_X, limbs)
    exp, cnt, el, es = r.ct_ct_matmul(0, X.reshape(-1), col_X, W.reshape(-1), col_X, limbs, SCALE, SCALE, col_X, row_X,
                                      col_X, row_X, nb)
    got, gs = be.ct_ct_matrix_mul_colpacking(keys_deep, pkg.to_device(X), pkg.to_device(W), SCALE, SCALE, col_X, row_X,
                                             col_X, row_X, nb)
    assert cnt == row_X and got.shape[0] == row_X and got.shape[2] == el and gs == es
    assert (pkg.to_host(got).reshape(-1) == exp).all()


def test_ct_ct_diagpacking_vs_reference(pkg, backend_deep, sealref_deep, keys_deep):
    r, be = sealref_deep, backend_deep
    rng = np.random.default_rng(30)
    limbs, col_X, col_W, nb = 3, 8, 3, 16
    X, _ = encrypt_batch(r, rng, col_X, limbs)     # row_X = col_X diagonals
    W, _ = encrypt_batch(r, rng, col_W, limbs)
    exp, cnt, el, es = r.ct_ct_matmul(1, X.reshape(-1), col_X, W.reshape(-1), col_W, limbs, SCALE, SCALE, col_X, col_X,
                                      col_W, col_X, nb)
    got, gs = be.ct_ct_matrix_mul_diagpacking(keys_deep, pkg.to_device(X), pkg.to_device(W), SCALE, SCALE, col_X, col_X,
                                              col_W, col_X, nb)
    assert cnt == col_W and got.shape[2] == el and gs == es
    assert (pkg.to_host(got).reshape(-1) == exp).all()


def test_exp_and_inverse_vs_seal_op_sequence(pkg, backend_deep, sealref_deep, keys_deep):
    """exp / inverse live in softmax.hpp, which needs NTL to compile; the same op sequence
    (softmax.hpp:9-47, 49-82) is replayed on real SEAL through ref_eval."""
    r, be = sealref_deep, backend_deep
    rng = np.random.default_rng(40)
    limbs = 12
    x, _ = encrypt_batch(r, rng, 2, limbs, sigma=1.0)

    def ev(op, a, sa, la, sc, **kw):
        return r.eval(op, a, sa, la, sc, **kw)

    # exp
    got, gs = be.exp(keys_deep, pkg.to_device(x), SCALE)
    for i in range(2):
        c, _, l, s = ev(OP_MULTIPLY_PLAIN, x[i].reshape(-1), 2, limbs, SCALE, b=r.encode_scalar(0.0078125, SCALE, limbs),
                        size_b=1, limbs_b=limbs, scale_b=SCALE)
        c, _, l, s = ev(OP_RESCALE, c, 2, l, s)
        c, _, l, s = ev(OP_ADD_PLAIN, c, 2, l, s, b=r.encode_scalar(1.0, s, l), size_b=1, limbs_b=l, scale_b=s)
        for _ in range(7):
            c, _, l, s = ev(OP_SQUARE, c, 2, l, s)
            c, _, l, s = ev(OP_RELINEARIZE, c, 3, l, s)
            c, _, l, s = ev(OP_RESCALE, c, 2, l, s)
        assert got.shape[2] == l and gs == s
        assert (pkg.to_host(got[i]).reshape(-1) == c).all()
    # inverse, 3 iterations
    got, gs = be.inverse(keys_deep, pkg.to_device(x), SCALE, 3)
    for i in range(2):
        one = r.encode_scalar(1.0, SCALE, limbs)
        y, _, l, s = ev(OP_SUB_PLAIN, x[i].reshape(-1), 2, limbs, SCALE, b=one, size_b=1, limbs_b=limbs, scale_b=SCALE)
        y, _, l, s = ev(OP_NEGATE, y, 2, l, s)
        res, _, lr, sr = ev(OP_ADD_PLAIN, y, 2, l, s, b=one, size_b=1, limbs_b=l, scale_b=s)
        for _ in range(3):
            y, _, l, s = ev(OP_SQUARE, y, 2, l, s)
            y, _, l, s = ev(OP_RELINEARIZE, y, 3, l, s)
            y, _, l, s = ev(OP_RESCALE, y, 2, l, s)
            tmp, _, lt, st = ev(OP_ADD_PLAIN, y, 2, l, s, b=r.encode_scalar(1.0, s, l), size_b=1, limbs_b=l, scale_b=s)
            while lr > lt:
                res, _, lr, sr = ev(OP_MOD_SWITCH, res, 2, lr, sr)
            res, _, lr, sr = ev(OP_MULTIPLY, res, 2, lr, sr, b=tmp, size_b=2, limbs_b=lt, scale_b=st)
            res, _, lr, sr = ev(OP_RELINEARIZE, res, 3, lr, sr)
            res, _, lr, sr = ev(OP_RESCALE, res, 2, lr, sr)
        assert got.shape[2] == lr and gs == sr
        assert (pkg.to_host(got[i]).reshape(-1) == res).all()


def test_gelu_decrypts_to_gelu(pkg, backend_deep, sealref_deep, keys_deep):
    """Meaning check (tolerance 0.1): decrypted gelu_v2 output ~ x * Phi(x) on [-3, 3].  The
    reference's degree-24 polynomial itself deviates from GELU by up to ~0.05 on this range, and
    the chain here has 30-bit primes (scale 2^30), far below the repo's 46 bits."""
    r, be = sealref_deep, backend_deep
    rng = np.random.default_rng(50)
    limbs, top = 9, sealref_deep.kl - 1
    v = rng.uniform(-3, 3, size=r.n // 2)
    ct = r.encrypt(r.encode_real(v, SCALE, top), top, SCALE).reshape(1, 2, top, r.n)[:, :, :limbs, :]
    got, gs = be.gelu_v2(keys_deep, pkg.to_device(np.ascontiguousarray(ct)), SCALE)
    dec = r.decode(r.decrypt(pkg.to_host(got).reshape(-1), 2, got.shape[2], gs), got.shape[2], gs).real
    ref = np.array([0.5 * t * (1 + math.erf(t / math.sqrt(2))) for t in v])
    assert np.abs(dec - ref).max() < 0.1


def test_gelu_fast_mode_bsgs_matches_reference_evaluation(pkg, backend_deep, sealref_deep, keys_deep):
    """With grouped relinearisation keys registered (fast mode) gelu_v2 evaluates the SAME degree-24 polynomial
    baby-step / giant-step (9 relinearizations instead of 23, csrc/modules.cu gelu_bsgs), with the reference's scale
    drift folded into the coefficients.  On this 30-bit chain (scale 2^30) the polynomial's coefficients of ~1e4 amplify
    the encryption noise of either evaluation order to ~1e-2 (test_gelu_decrypts_to_gelu allows 0.1 for the same reason):
    stated tolerance 5e-2 between the two orders (measured 3.5e-2).  The repo's 46-bit chain is gated in
    tests/test_gpu_fullsize.py against the reference's decrypted output: 1.1e-4 even on the 17 activations beyond the
    polynomial's domain."""
    r, be = sealref_deep, backend_deep
    rng = np.random.default_rng(51)
    limbs, top = 9, sealref_deep.kl - 1
    v = rng.uniform(-3, 3, size=(2, r.n // 2))
    ct = np.stack([r.encrypt(r.encode_real(v[i], SCALE, top), top, SCALE).reshape(2, top, r.n)[:, :limbs, :] for i in range(2)])
    d = pkg.to_device(np.ascontiguousarray(ct))
    relin = pkg.to_device(r.export_relin_key())
    relin4 = relin.reshape(r.kl - 1, 2, r.kl, r.n)
    variants = [be.key_prepare_grouped(relin4, 0, lv, k_extra=k, pre_permute=False)
                for k, lv in sorted(be.ksg_plan(range(1, 10)).items())]
    assert variants
    k_fast = be.make_keys(relin=relin, grouped={0: variants})
    ref, rs = be.gelu_v2(keys_deep, d, SCALE)
    got, gs = be.gelu_v2(k_fast, d, SCALE)
    assert got.shape == ref.shape and gs == rs
    assert not (got == ref).all()
    dr = np.stack([r.decode(r.decrypt(pkg.to_host(ref)[i].reshape(-1), 2, ref.shape[2], rs), ref.shape[2], rs).real for i in range(2)])
    dg = np.stack([r.decode(r.decrypt(pkg.to_host(got)[i].reshape(-1), 2, got.shape[2], gs), got.shape[2], gs).real for i in range(2)])
    err = np.abs(dg - dr).max()
    print("gelu: BSGS vs all-powers evaluation, max-abs difference %.3g" % err)
    assert err < 5e-2, err
