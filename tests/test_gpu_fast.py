"""GPU checks of the fast mode: level-truncated, pre-permuted Galois keys
and hoisted rotations.  A hoisted rotation is NOT SEAL's rotate_vector bit for bit — SEAL applies the
automorphism before the digit decomposition (S/evaluator.cpp:2635-2657), hoisting decomposes the
unrotated polynomial — so parity is by TOLERANCE on the decrypted slots against the SEAL-exact
rotation of the same ciphertext with the same key: stated tolerance 1e-6 max-abs at scale 2^46
(both results carry key-switching noise of ~2^-25 relative to the scale).  What IS bit-exact, and
checked so: a hoisted rotation equals the un-hoisted fast rotation (same decomposition), and a
truncated key gives the same residues as the full pre-permuted key."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BITS = [51] + [46] * 2 + [51] * 14 + [58]        # 17 data limbs, same shape as the repo's chain
SCALE = 2.0 ** 46


@pytest.fixture(scope="module")
def env(pkg):
    from oracle import Oracle
    o = Oracle(12, BITS)
    be = pkg.Backend(12, o.q)
    sk = o.gen_secret(3, hamming_weight=64)
    return o, be, sk


def encrypt_batch(o, sk, rng, B, limbs):
    zs = (rng.normal(size=(B, o.n // 2)) + 1j * rng.normal(size=(B, o.n // 2))) * 0.5
    cts = np.stack([o.encrypt_sym(sk, 70 + i, o.encode(zs[i], SCALE, limbs), limbs) for i in range(B)])
    return zs, cts.reshape(B, 2, limbs, o.n)


def decrypt_batch(o, sk, pkg, ct):
    ct = pkg.to_host(ct)
    limbs = ct.shape[2]
    return np.stack([o.decode(o.decrypt(sk, ct[i].reshape(-1), 2, limbs), limbs, SCALE) for i in range(ct.shape[0])])


@pytest.mark.parametrize("limbs", [17, 9, 2])
def test_fast_rotation_matches_exact_rotation(pkg, env, limbs):
    o, be, sk = env
    rng = np.random.default_rng(limbs)
    steps = [1, 5, 64, o.n // 2 - 3, 7, 100]   # 6 rotations: one pass with four keys, one with two
    zs, cts = encrypt_batch(o, sk, rng, 3, limbs)
    d = pkg.to_device(cts)
    exact, fast, fast_trunc = {}, {}, {}
    for i, st in enumerate(steps):
        e = o.elt_from_step(st)
        k = pkg.to_device(o.gen_galois_key(sk, 400 + i, e).reshape(o.kl - 1, 2, o.kl, o.n))
        exact[e] = k
        fast[e] = be.key_prepare(k, e)                       # full size, pre-permuted
        fast_trunc[e] = be.key_prepare(k, e, max_limbs=limbs)  # truncated to this level
    k_exact = be.make_keys(galois=exact)
    k_fast = be.make_keys(galois_fast=fast)
    k_trunc = be.make_keys(galois_fast=fast_trunc)
    many = be.rotate_many(k_fast, d, steps)                  # hoisted: one decomposition
    many_t = be.rotate_many(k_trunc, d, steps)
    assert (many == many_t).all(), "truncated key must give the same residues"
    for i, st in enumerate(steps):
        single = be.rotate_vector_keys(k_fast, d, st)        # un-hoisted fast rotation
        assert (single == many[i]).all(), "hoisted and un-hoisted fast rotations must agree bit for bit"
        ref = decrypt_batch(o, sk, pkg, be.rotate_vector_keys(k_exact, d, st))
        got = decrypt_batch(o, sk, pkg, many[i])
        want = np.roll(zs, -st, axis=1)
        assert np.abs(ref - want).max() < 1e-6
        assert np.abs(got - want).max() < 1e-6, (st, np.abs(got - want).max())
        assert np.abs(got - ref).max() < 1e-6


def test_key_too_short_is_rejected(pkg, env):
    o, be, sk = env
    rng = np.random.default_rng(6)
    _, cts = encrypt_batch(o, sk, rng, 1, 9)
    e = o.elt_from_step(1)
    k = pkg.to_device(o.gen_galois_key(sk, 78, e).reshape(o.kl - 1, 2, o.kl, o.n))
    keys = be.make_keys(galois_fast={e: be.key_prepare(k, e, max_limbs=4)})
    with pytest.raises(pkg.MoaiError):
        be.rotate_vector_keys(keys, pkg.to_device(cts), 1)   # 9 limbs > 4: no key, NAF has a single term


def test_bootstrap_hoisted(pkg, env):
    """Bootstrapping planned for hoisted baby steps, all Galois keys pre-permuted: same tolerance as
    the exact-key pipeline (tests/test_gpu_bootstrap.py): max-abs slot error < 2e-3."""
    o, be, sk = env
    boot = pkg.Bootstrapper(be, total_limbs=17)
    plain_steps = boot.required_steps()
    boot.set_hoisting(True)
    steps = boot.required_steps()
    assert set(steps) != set(plain_steps)     # re-planned: more (cheap) baby steps, fewer giant steps
    fast = {}
    for i, st in enumerate(steps + [0]):
        e = o.elt_from_step(st)
        k = pkg.to_device(o.gen_galois_key(sk, 1000 + i, e).reshape(o.kl - 1, 2, o.kl, o.n))
        fast[e] = be.key_prepare(k, e)
    keys = be.make_keys(relin=pkg.to_device(o.gen_relin_key(sk, 5)), galois_fast=fast)
    rng = np.random.default_rng(1)
    B = 3
    zs = (rng.normal(size=(B, o.n // 2)) + 1j * rng.normal(size=(B, o.n // 2))) * 0.1
    zs[2] = 0.0
    cts = np.stack([o.encrypt_sym(sk, 50 + i, o.encode(zs[i], SCALE, 1), 1) for i in range(B)])
    out, out_scale = boot.bootstrap_3(keys, pkg.to_device(cts.reshape(B, 2, 1, o.n)), SCALE)
    assert out.shape[2] == 3 and out_scale == SCALE
    res = pkg.to_host(out)
    for i in range(B):
        dec = o.decode(o.decrypt(sk, res[i].reshape(-1), 2, 3), 3, out_scale)
        assert np.abs(dec - zs[i]).max() < 2e-3, (i, np.abs(dec - zs[i]).max())


def test_colpacking_fast_matches_exact(pkg, env):
    """ct_ct_matrix_mul_colpacking (Ct_ct_matrix_mul.hpp:5-55) with hoisted rotations vs the SEAL-exact
    module on the same inputs: decrypted outputs agree to 1e-5 (products of O(1) values)."""
    o, be, sk = env
    rng = np.random.default_rng(9)
    limbs, cols, rows, nb = 3, 4, 128, o.n // 2 // 128   # last prime of 3 limbs is 46-bit ~ the scale
    zx, X = encrypt_batch(o, sk, rng, cols, limbs)
    zw, W = encrypt_batch(o, sk, rng, cols, limbs)
    relin = pkg.to_device(o.gen_relin_key(sk, 5))
    exact, fast = {}, {}
    need = set(b * nb for b in range(1, 16)) | set(a * 16 * nb for a in range(1, 8))
    for i, st in enumerate(sorted(need)):
        e = o.elt_from_step(st)
        k = pkg.to_device(o.gen_galois_key(sk, 600 + i, e).reshape(o.kl - 1, 2, o.kl, o.n))
        exact[e] = k
        fast[e] = be.key_prepare(k, e, max_limbs=limbs)
    for k2 in range(7):                                       # SEAL's NAF fallback also rotates right
        e = o.elt_from_step(o.n // 2 - (1 << k2) * nb)
        exact[e] = pkg.to_device(o.gen_galois_key(sk, 700 + k2, e).reshape(o.kl - 1, 2, o.kl, o.n))
    args = (pkg.to_device(X), pkg.to_device(W), SCALE, SCALE, cols, rows, cols, rows, nb)
    got_e, s_e = be.ct_ct_matrix_mul_colpacking(be.make_keys(relin=relin, galois=exact), *args)
    got_f, s_f = be.ct_ct_matrix_mul_colpacking(be.make_keys(relin=relin, galois_fast=fast), *args)
    assert got_e.shape == got_f.shape and s_e == s_f
    de, df = decrypt_batch(o, sk, pkg, got_e), decrypt_batch(o, sk, pkg, got_f)
    want = np.stack([sum(zx[j] * np.roll(zw[j], -i * nb) for j in range(cols)) for i in range(rows)])
    assert np.abs(de - want).max() < 1e-5
    assert np.abs(df - want).max() < 1e-5, np.abs(df - want).max()


def test_truncated_exact_key_is_bit_identical(pkg, env):
    """A level-truncated SEAL-layout key (key_prepare(pre_permute=False)) gives SEAL's residues bit for bit."""
    o, be, sk = env
    rng = np.random.default_rng(21)
    limbs = 5
    _, cts = encrypt_batch(o, sk, rng, 2, limbs)
    e = o.elt_from_step(3)
    k = pkg.to_device(o.gen_galois_key(sk, 900, e).reshape(o.kl - 1, 2, o.kl, o.n))
    full = be.rotate_vector_keys(be.make_keys(galois={e: k}), pkg.to_device(cts), 3)
    trunc = be.rotate_vector_keys(be.make_keys(galois={e: be.key_prepare(k, e, max_limbs=limbs, pre_permute=False)}),
                                  pkg.to_device(cts), 3)
    assert (full == trunc).all()
    exp = o.apply_galois(cts[0].reshape(-1), limbs, e, pkg.to_host(k).reshape(-1))
    assert (pkg.to_host(full[0]).reshape(-1) == exp).all()
