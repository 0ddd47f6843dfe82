"""GPU parity (bit-exact) of the RNS core against the oracle — BASELINE.json config 2:
NTT/INTT, element-wise ops, rescale, mod-switch, rotate / conjugate / relinearize (key switch)
at the repo's degree across levels.  Everything goes through the C ABI (ctypes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rand_ct(o, rng, batch, size, limbs):
    out = np.empty((batch, size, limbs, o.n), dtype=np.uint64)
    for l in range(limbs):
        out[:, :, l, :] = rng.integers(0, int(o.q[l]), (batch, size, o.n), dtype=np.uint64)
    return out


def rand_ksk(o, rng):
    kl = o.kl
    out = np.empty((kl - 1, 2, kl, o.n), dtype=np.uint64)
    for l in range(kl):
        out[:, :, l, :] = rng.integers(0, int(o.q[l]), (kl - 1, 2, o.n), dtype=np.uint64)
    return out


# ------------------------------------------------------------------------------ NTT
def test_ntt_small_all_limbs(pkg, backend_small, oracle_small):
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(1)
    x = rand_ct(o, rng, 3, 2, o.kl)
    d = pkg.to_device(x)
    be.ntt_forward_(d)
    got = pkg.to_host(d)
    for b in range(3):
        for p in range(2):
            for l in range(o.kl):
                assert (got[b, p, l] == o.ntt(l, x[b, p, l])).all(), (b, p, l)
    be.ntt_inverse_(d)
    assert (pkg.to_host(d) == x).all()
    d = pkg.to_device(x)
    be.ntt_inverse_(d)
    got = pkg.to_host(d)
    for l in range(o.kl):
        assert (got[1, 1, l] == o.intt(l, x[1, 1, l])).all()


@pytest.mark.parametrize("log_n", [13, 14, 15])
def test_ntt_other_degrees(pkg, log_n):
    from oracle import Oracle
    o = Oracle(log_n, [50, 40, 58])
    be = pkg.Backend(log_n, o.q)
    rng = np.random.default_rng(log_n)
    x = rand_ct(o, rng, 2, 1, 3)
    d = pkg.to_device(x)
    be.ntt_forward_(d)
    got = pkg.to_host(d)
    for l in range(3):
        assert (got[0, 0, l] == o.ntt(l, x[0, 0, l])).all()
    d = pkg.to_device(x)
    be.ntt_inverse_(d)
    got = pkg.to_host(d)
    for l in range(3):
        assert (got[1, 0, l] == o.intt(l, x[1, 0, l])).all()
    be.close()


def test_ntt_moai_all_36_primes(pkg, backend_moai, oracle_moai):
    """N = 65536, every prime of the repo's chain incl. the 58-bit special prime; edge values."""
    o, be = oracle_moai, backend_moai
    rng = np.random.default_rng(2)
    x = rand_ct(o, rng, 1, 1, o.kl)
    x[0, 0, :, 0] = 0
    for l in range(o.kl):
        x[0, 0, l, 1] = int(o.q[l]) - 1
    d = pkg.to_device(x)
    be.ntt_forward_(d)
    got = pkg.to_host(d)
    for l in range(o.kl):
        assert (got[0, 0, l] == o.ntt(l, x[0, 0, l])).all(), l
    d = pkg.to_device(x)
    be.ntt_inverse_(d)
    got = pkg.to_host(d)
    for l in range(o.kl):
        assert (got[0, 0, l] == o.intt(l, x[0, 0, l])).all(), l
    # single-limb batched entry point on the special prime
    y = rng.integers(0, int(o.q[35]), (4, o.n), dtype=np.uint64)
    d = pkg.to_device(y)
    be.ntt_forward_limb_(d, 35)
    got = pkg.to_host(d)
    assert (got[3] == o.ntt(35, y[3])).all()
    be.ntt_inverse_limb_(d, 35)
    assert (pkg.to_host(d) == y).all()


def test_ntt_full_size_roundtrip_and_linearity(pkg, backend_moai, oracle_moai):
    """Size-independent properties at BASELINE sizes: INTT(NTT(x)) = x and NTT(x + y) = NTT(x) + NTT(y)."""
    o, be = oracle_moai, backend_moai
    rng = np.random.default_rng(3)
    x = rand_ct(o, rng, 8, 2, 35)
    y = rand_ct(o, rng, 8, 2, 35)
    dx, dy = pkg.to_device(x), pkg.to_device(y)
    ds = be.add(dx, dy)
    be.ntt_forward_(dx)
    be.ntt_forward_(dy)
    be.ntt_forward_(ds)
    assert bool((be.add(dx, dy) == ds).all())
    be.ntt_inverse_(dx)
    assert (pkg.to_host(dx) == x).all()


# ------------------------------------------------------------------------------ element-wise
@pytest.mark.parametrize("limbs", [4, 1])
def test_elementwise_small(pkg, backend_small, oracle_small, limbs):
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(limbs)
    B = 3
    a, b = rand_ct(o, rng, B, 2, limbs), rand_ct(o, rng, B, 2, limbs)
    a[0, 0, :, :4] = 0
    b[0, 0, :, :4] = 0
    pt = rand_ct(o, rng, 1, 1, limbs)[0, 0]
    pts = np.ascontiguousarray(rand_ct(o, rng, B, 1, limbs)[:, 0])
    da, db, dpt, dpts = (pkg.to_device(v) for v in (a, b, pt, pts))
    add, sub, neg = (pkg.to_host(v) for v in (be.add(da, db), be.sub(da, db), be.negate(da)))
    mul, sq = pkg.to_host(be.multiply(da, db)), pkg.to_host(be.square(da))
    mp, ap, sp = (pkg.to_host(v) for v in (be.multiply_plain(da, dpt), be.add_plain(da, dpt), be.sub_plain(da, dpt)))
    mps = pkg.to_host(be.multiply_plain(da, dpts))
    acc = be.multiply(da, db)
    be.multiply(db, db, out=acc, accumulate=True)
    acc = pkg.to_host(acc)
    for i in range(B):
        fa, fb = a[i].reshape(-1), b[i].reshape(-1)
        assert (add[i].reshape(-1) == o.add(fa, fb, 2, limbs)).all()
        assert (sub[i].reshape(-1) == o.sub(fa, fb, 2, limbs)).all()
        assert (neg[i].reshape(-1) == o.negate(fa, 2, limbs)).all()
        assert (mul[i].reshape(-1) == o.multiply(fa, fb, limbs)).all()
        assert (sq[i].reshape(-1) == o.square(fa, limbs)).all()
        assert (mp[i].reshape(-1) == o.multiply_plain(fa, pt.reshape(-1), 2, limbs)).all()
        assert (mps[i].reshape(-1) == o.multiply_plain(fa, pts[i].reshape(-1), 2, limbs)).all()
        assert (ap[i].reshape(-1) == o.addsub_plain(0, fa, pt.reshape(-1), 2, limbs)).all()
        assert (sp[i].reshape(-1) == o.addsub_plain(1, fa, pt.reshape(-1), 2, limbs)).all()
        exp = o.add(o.multiply(fa, fb, limbs), o.multiply(fb, fb, limbs), 3, limbs)
        assert (acc[i].reshape(-1) == exp).all()


def test_scalar_ops_small(pkg, backend_small, oracle_small):
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(5)
    limbs, scale = 3, 2.0 ** 30
    a = rand_ct(o, rng, 2, 2, limbs)
    da = pkg.to_device(a)
    for v in (0.5, -1.25, 0.0, 1e-3):
        assert (be.encode_scalar_consts(v, scale, limbs) == o.encode_scalar_consts(v, scale, limbs)).all()
        pt = o.encode_scalar(v, scale, limbs)
        got = pkg.to_host(be.multiply_const(da, v, scale))
        assert (got[1].reshape(-1) == o.multiply_plain(a[1].reshape(-1), pt, 2, limbs)).all()
        got = pkg.to_host(be.add_const(da, v, scale))
        assert (got[1].reshape(-1) == o.addsub_plain(0, a[1].reshape(-1), pt, 2, limbs)).all()


# ------------------------------------------------------------------------------ rescale / mod switch / mod raise
@pytest.mark.parametrize("limbs,size", [(4, 2), (3, 3), (2, 2)])
def test_rescale_modswitch_small(pkg, backend_small, oracle_small, limbs, size):
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(limbs * 10 + size)
    a = rand_ct(o, rng, 3, size, limbs)
    da = pkg.to_device(a)
    rs = pkg.to_host(be.rescale_to_next(da))
    ms = pkg.to_host(be.mod_switch_to_next(da))
    for i in range(3):
        assert (rs[i].reshape(-1) == o.rescale(a[i].reshape(-1), size, limbs)).all()
        assert (ms[i].reshape(-1) == o.mod_switch(a[i].reshape(-1), size, limbs)).all()
    assert (pkg.to_host(be.mod_switch_to(da, 1)) == a[:, :, :1]).all()


@pytest.mark.parametrize("limbs", [35, 21, 15, 2])
def test_rescale_moai_levels(pkg, backend_moai, oracle_moai, limbs):
    o, be = oracle_moai, backend_moai
    rng = np.random.default_rng(limbs)
    a = rand_ct(o, rng, 2, 2, limbs)
    rs = pkg.to_host(be.rescale_to_next(pkg.to_device(a)))
    assert (rs[1].reshape(-1) == o.rescale(a[1].reshape(-1), 2, limbs)).all()


def test_mod_raise(pkg, backend_small, oracle_small, backend_moai, oracle_moai):
    for o, be, L in ((oracle_small, backend_small, 4), (oracle_moai, backend_moai, 35)):
        rng = np.random.default_rng(L)
        a = rand_ct(o, rng, 2, 2, 1)
        got = pkg.to_host(be.mod_raise(pkg.to_device(a), L))
        assert (got[1].reshape(-1) == o.modraise(a[1].reshape(-1), 2, L)).all()


# ------------------------------------------------------------------------------ key switching
@pytest.mark.parametrize("limbs", [4, 3, 1])
def test_keyswitch_small(pkg, backend_small, oracle_small, limbs):
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(20 + limbs)
    B = 3
    ksk = rand_ksk(o, rng)
    dk = pkg.to_device(ksk)
    a3 = rand_ct(o, rng, B, 3, limbs)
    got = pkg.to_host(be.relinearize(pkg.to_device(a3), dk))
    for i in range(B):
        assert (got[i].reshape(-1) == o.relinearize(a3[i].reshape(-1), limbs, ksk.reshape(-1))).all()
    a = rand_ct(o, rng, B, 2, limbs)
    da = pkg.to_device(a)
    for step in (1, -1, 5, 0):
        elt = be.galois_elt_from_step(step)
        assert elt == o.elt_from_step(step)
        got = pkg.to_host(be.apply_galois(da, elt, dk))
        for i in range(B):
            assert (got[i].reshape(-1) == o.apply_galois(a[i].reshape(-1), limbs, elt, ksk.reshape(-1))).all()
    for s in (3, -7, 768, 2047, 100):
        assert be.rotate_naf_steps(s) == o.naf_steps(s)


@pytest.mark.parametrize("limbs", [35, 21, 14, 3])
def test_keyswitch_moai_levels(pkg, backend_moai, oracle_moai, limbs):
    """Rotation (Galois + key switch) at the repo's parameters; the key is uniformly random
    residues — parity of the key-switch function does not depend on the key being valid."""
    o, be = oracle_moai, backend_moai
    rng = np.random.default_rng(30 + limbs)
    ksk = rand_ksk(o, rng)
    dk = pkg.to_device(ksk)
    a = rand_ct(o, rng, 2, 2, limbs)
    elt = be.galois_elt_from_step(256)
    got = pkg.to_host(be.apply_galois(pkg.to_device(a), elt, dk))
    assert (got[1].reshape(-1) == o.apply_galois(a[1].reshape(-1), limbs, elt, ksk.reshape(-1))).all()
    del dk


def test_rotate_semantics_with_real_keys(pkg, backend_small, oracle_small):
    """Decrypt-level check with valid keys: rotate (incl. NAF fallback), conjugate, multiply +
    relinearize + rescale on the GPU, decrypt/decode with the oracle."""
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(40)
    scale, limbs = 2.0 ** 30, 4
    sk = o.gen_secret(5, hamming_weight=64)
    z = (rng.normal(size=o.n // 2) + 1j * rng.normal(size=o.n // 2)) * 0.5
    ct = o.encrypt_sym(sk, 1, o.encode(z, scale, limbs), limbs)
    dct = pkg.to_device(ct.reshape(1, 2, limbs, o.n))
    keys = {}
    for i, s in enumerate((1, -1, 2, 4, 8, 0)):
        e = o.elt_from_step(s)
        keys[e] = pkg.to_device(o.gen_galois_key(sk, 100 + i, e))
    rlk = pkg.to_device(o.gen_relin_key(sk, 7))

    def dec(d, lm, sc, size=2):
        return o.decode(o.decrypt(sk, pkg.to_host(d).reshape(-1), size, lm), lm, sc)

    assert np.abs(dec(be.rotate_vector(dct, 1, keys), limbs, scale) - np.roll(z, -1)).max() < 1e-3
    assert np.abs(dec(be.rotate_vector(dct, 7, keys), limbs, scale) - np.roll(z, -7)).max() < 1e-3  # NAF: -1 + 8
    assert np.abs(dec(be.complex_conjugate(dct, keys), limbs, scale) - np.conj(z)).max() < 1e-3
    sq = be.rescale_to_next(be.relinearize(be.square(dct), rlk))
    got = dec(sq, limbs - 1, scale * scale / float(o.q[limbs - 1]))
    assert np.abs(got - z * z).max() < 1e-3
    with pytest.raises(pkg.MoaiError):
        be.rotate_vector(dct, 16, keys)  # power of two without a key: "Galois key not present"


def test_error_codes(pkg, backend_small, oracle_small):
    o, be = oracle_small, backend_small
    a = pkg.to_device(rand_ct(o, np.random.default_rng(0), 1, 2, 1))
    with pytest.raises(pkg.MoaiError) as ei:
        be.rescale_to_next(a)  # end of modulus switching chain reached (S/evaluator.cpp:1593-1596)
    assert ei.value.code == 1
    empty = be.empty(0, 2, 3, o.n)
    assert be.add(empty, empty).shape[0] == 0  # empty batch is a no-op


def test_context_lanes_from_threads(pkg):
    """moai_context_fork: four host threads, each with its own lane (stream + arena) of one context, run NTT ->
    multiply_plain -> rescale -> rotation chains concurrently; every lane's result equals the root context's, bit for
    bit (the tables — twiddles, Galois permutations — are shared and immutable)."""
    import threading
    import torch
    from oracle import Oracle
    o = Oracle(12, [40, 30, 30, 40])
    be = pkg.Backend(12, o.q)
    rng = np.random.default_rng(5)
    limbs = 3

    def rnd(size, batch=4):
        out = np.empty((batch, size, limbs, o.n), dtype=np.uint64)
        for l in range(limbs):
            out[:, :, l, :] = rng.integers(0, int(o.q[l]), (batch, size, o.n), dtype=np.uint64)
        return out

    xs = [pkg.to_device(rnd(2)) for _ in range(4)]
    pt = pkg.to_device(rnd(1, 1)[0, 0])
    ksk = np.empty((o.kl - 1, 2, o.kl, o.n), dtype=np.uint64)
    for l in range(o.kl):
        ksk[:, :, l, :] = rng.integers(0, int(o.q[l]), (o.kl - 1, 2, o.n), dtype=np.uint64)
    dk = pkg.to_device(ksk)
    elt = be.galois_elt_from_step(3)
    torch.cuda.synchronize()

    def chain(b, x):
        # the tensors are torch allocations (torch's own stream-ordered pool, not the lane's arena): drain the lane before
        # an intermediate is dropped
        y = b.multiply_plain(x, pt)
        z = b.rescale_to_next(y)
        r = b.apply_galois(z, elt, dk)
        b.synchronize()
        return r

    want = [chain(be, x) for x in xs]
    be.synchronize()
    lanes = [be.fork() for _ in range(4)]
    got, errs = [None] * 4, []

    def work(i):
        try:
            for _ in range(5):
                got[i] = chain(lanes[i], xs[i])
            lanes[i].synchronize()
        except Exception as e:   # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i in range(4):
        assert (got[i] == want[i]).all()
    for ln in lanes:
        ln.close()
    be.close()
