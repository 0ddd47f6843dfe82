"""GPU checks of the grouped-digit key switch (csrc/ksgroup.hpp): the fast-mode replacement of SEAL's
one-digit-per-prime key switch (S/evaluator.cpp:2724-3021).  The grouped keys are DERIVED on the device from
the keys SEAL's KeyGenerator makes (sums of key digits, S/keygenerator.cpp:316-371), so the test generates
stock keys with the oracle, prepares them, and compares the DECRYPTED results with the SEAL-exact operation on
the same ciphertext: same plaintext, different noise.  Stated tolerance: 1e-6 max-abs on O(1) slots at scale
2^46 (measured: ~1e-9).  Bit-exact and checked so: a hoisted grouped rotation equals the un-hoisted one (same
digits, exact integer inner products)."""
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


def encrypt_batch(o, sk, rng, B, limbs, amp=0.5):
    zs = (rng.normal(size=(B, o.n // 2)) + 1j * rng.normal(size=(B, o.n // 2))) * amp
    cts = np.stack([o.encrypt_sym(sk, 70 + i, o.encode(zs[i], SCALE, limbs), limbs) for i in range(B)])
    return zs, cts.reshape(B, 2, limbs, o.n)


def decrypt_batch(o, sk, pkg, ct, scale=SCALE):
    ct = pkg.to_host(ct)
    limbs = ct.shape[2]
    return np.stack([o.decode(o.decrypt(sk, ct[i].reshape(-1), 2, limbs), limbs, scale) for i in range(ct.shape[0])])


def test_plan_shapes(pkg, env):
    """Groups are sized so that prod(group) <= p * prod(extra primes); a key for k extra primes serves levels
    <= 17 - k."""
    o, be, sk = env
    d, kl = be.ksg_key_shape(1, 16)      # P' = 58 + 51 bits: two primes per digit
    assert kl == 18 and d == 8
    d, kl = be.ksg_key_shape(5, 12)      # P' = 58 + 255 bits: six primes per digit (51 + 46 + 46 + 3 x 51 = 296)
    assert kl == 18 and d == 2
    with pytest.raises(pkg.MoaiError):
        be.ksg_key_shape(5, 13)          # 13 + 5 > 17 data primes
    assert be.ksg_best_extra(17) == 0    # no spare prime at the top level
    assert be.ksg_best_extra(12) > 0


@pytest.mark.parametrize("limbs,k,key_limbs", [(15, 2, 15), (16, 1, 16), (9, 5, 12), (9, 2, 15), (2, 3, 8), (12, 5, 12)])
def test_grouped_rotation_matches_exact_rotation(pkg, env, limbs, k, key_limbs):
    o, be, sk = env
    rng = np.random.default_rng(limbs * 10 + k)
    steps = [1, 5, 64, o.n // 2 - 3, 7, 100]
    zs, cts = encrypt_batch(o, sk, rng, 3, limbs)
    d = pkg.to_device(cts)
    exact, grouped = {}, {}
    for i, st in enumerate(steps):
        e = o.elt_from_step(st)
        kk = pkg.to_device(o.gen_galois_key(sk, 400 + i, e).reshape(o.kl - 1, 2, o.kl, o.n))
        exact[e] = kk
        grouped[e] = [be.key_prepare_grouped(kk, e, key_limbs, k_extra=k)]
    k_exact = be.make_keys(galois=exact)
    k_grp = be.make_keys(grouped=grouped)
    many = be.rotate_many(k_grp, d, steps)                    # hoisted: one grouped decomposition
    for i, st in enumerate(steps):
        single = be.rotate_vector_keys(k_grp, d, st)          # fused pass-B + inner-product kernel
        assert (single == many[i]).all(), "hoisted and un-hoisted grouped rotations must agree bit for bit"
        ref = decrypt_batch(o, sk, pkg, be.rotate_vector_keys(k_exact, d, st))
        got = decrypt_batch(o, sk, pkg, many[i])
        want = np.roll(zs, -st, axis=1)
        assert np.abs(ref - want).max() < 1e-6
        assert np.abs(got - want).max() < 1e-6, (st, np.abs(got - want).max())
        assert np.abs(got - ref).max() < 1e-6


def test_grouped_conjugation_and_relinearisation(pkg, env):
    """multiply + relinearize (S/evaluator.cpp:1345-1400) and the complex conjugation with grouped keys against the
    same calls with SEAL's keys: decrypted results agree to 1e-6."""
    o, be, sk = env
    rng = np.random.default_rng(77)
    relin = pkg.to_device(o.gen_relin_key(sk, 5))
    relin4 = relin.reshape(o.kl - 1, 2, o.kl, o.n)
    k_exact = be.make_keys(relin=relin)
    for limbs, key_limbs, k in [(12, 12, None), (9, 12, 5), (4, 6, 3), (16, 16, 1)]:
        zx, cx = encrypt_batch(o, sk, rng, 2, limbs)
        zy, cy = encrypt_batch(o, sk, rng, 2, limbs)
        prod3 = be.multiply(pkg.to_device(cx), pkg.to_device(cy))
        gk = be.key_prepare_grouped(relin4, 0, key_limbs, k_extra=k, pre_permute=False)
        assert gk is not None, "the cost model must prefer grouped digits below the top level"
        k_grp = be.make_keys(grouped={0: [gk]})
        re_, rg = be.relinearize_keys(k_exact, prod3), be.relinearize_keys(k_grp, prod3)
        assert not (re_ == rg).all(), "the grouped key was not used"
        de, dg = decrypt_batch(o, sk, pkg, re_, SCALE * SCALE), decrypt_batch(o, sk, pkg, rg, SCALE * SCALE)
        assert np.abs(de - zx * zy).max() < 1e-6
        assert np.abs(dg - zx * zy).max() < 1e-6, (limbs, np.abs(dg - zx * zy).max())
    limbs = 12
    # conjugation
    e = o.elt_from_step(0)
    kc = pkg.to_device(o.gen_galois_key(sk, 31, e).reshape(o.kl - 1, 2, o.kl, o.n))
    zc, cc = encrypt_batch(o, sk, rng, 2, limbs)
    dc = pkg.to_device(cc)
    out = be.complex_conjugate_keys(be.make_keys(grouped={e: [be.key_prepare_grouped(kc, e, limbs, k_extra=4)]}), dc)
    assert np.abs(decrypt_batch(o, sk, pkg, out) - np.conj(zc)).max() < 1e-6


def test_merged_relinearize_rescale(pkg, env):
    """rescale_to_next(relinearize(x)) with a grouped key divides by P' q_last ONCE (ksg_moddown_rescale): decrypted
    at the lower level it matches the two SEAL-exact calls (S/evaluator.cpp:1345-1400, :1402-1481) to 1e-6 (measured
    ~1e-9), and with SEAL's key the entry point IS the two exact calls, bit for bit."""
    o, be, sk = env
    rng = np.random.default_rng(91)
    relin = pkg.to_device(o.gen_relin_key(sk, 5))
    relin4 = relin.reshape(o.kl - 1, 2, o.kl, o.n)
    k_exact = be.make_keys(relin=relin)
    for limbs, key_limbs, k in [(12, 12, None), (9, 12, 5), (4, 6, 3), (16, 16, 1), (2, 8, 3)]:
        zx, cx = encrypt_batch(o, sk, rng, 3, limbs)
        zy, cy = encrypt_batch(o, sk, rng, 3, limbs)
        prod3 = be.multiply(pkg.to_device(cx), pkg.to_device(cy))
        exact = be.rescale_to_next(be.relinearize_keys(k_exact, prod3))
        assert (be.relin_rescale_keys(k_exact, prod3) == exact).all()
        gk = be.key_prepare_grouped(relin4, 0, key_limbs, k_extra=k, pre_permute=False)
        k_grp = be.make_keys(grouped={0: [gk]})
        merged = be.relin_rescale_keys(k_grp, prod3)
        assert merged.shape == exact.shape and not (merged == exact).all()
        sc = SCALE * SCALE / float(o.q[limbs - 1])
        de, dg = decrypt_batch(o, sk, pkg, exact, sc), decrypt_batch(o, sk, pkg, merged, sc)
        assert np.abs(de - zx * zy).max() < 1e-6
        assert np.abs(dg - zx * zy).max() < 1e-6, (limbs, np.abs(dg - zx * zy).max())
        print("merged relin+rescale limbs", limbs, "k", k, "err", np.abs(dg - zx * zy).max(), "exact", np.abs(de - zx * zy).max())


@pytest.fixture(scope="module", params=["0", "1", "fpsrc"], ids=["fp64", "mma", "fpsrc"])
def env13(pkg, request):
    """N = 2^13 with the base conversion as FP64 products (default), as the opt-in tensor-core GEMM (MOAI_CONV_MMA=1)
    and with the opt-in centred-double sources + precomputed quotients (MOAI_CONV_FPSRC=1); both switches are read when
    a level's conversion tables are first built, i.e. per context."""
    import os
    from oracle import Oracle
    old = os.environ.get("MOAI_CONV_MMA")
    os.environ["MOAI_CONV_MMA"] = "1" if request.param == "1" else "0"
    os.environ["MOAI_CONV_FPSRC"] = "1" if request.param == "fpsrc" else "0"
    o = Oracle(13, BITS)
    be = pkg.Backend(13, o.q)
    sk = o.gen_secret(3, hamming_weight=64)
    yield o, be, sk
    os.environ.pop("MOAI_CONV_FPSRC", None)
    if old is None:
        os.environ.pop("MOAI_CONV_MMA", None)
    else:
        os.environ["MOAI_CONV_MMA"] = old


@pytest.mark.parametrize("limbs,k,key_limbs", [(15, 2, 15), (9, 5, 12), (9, 8, 9), (3, 1, 16), (12, 3, 14)])
def test_conversion_on_tensor_cores(pkg, env13, limbs, k, key_limbs):
    """From N = 2^13 on the base conversions of the grouped key switch (digit extension, mod-down) run as a u8 GEMM on
    the tensor cores (ConvTab::BT, csrc/ntt.cuh; 1, 2 and 3 k-steps of four sources here).  Same integers as the FP64
    products, so the decrypted rotation / relinearisation must match the SEAL-exact operation as at N = 2^12 (1e-6),
    hoisted and un-hoisted rotations agree bit for bit, and so does the merged relinearize + rescale."""
    o, be, sk = env13
    import os
    mma = os.environ.get("MOAI_CONV_MMA") == "1"
    be.profile(True)
    rng = np.random.default_rng(limbs * 10 + k)
    steps = [1, 64, o.n // 2 - 3]
    zs, cts = encrypt_batch(o, sk, rng, 2, limbs)
    d = pkg.to_device(cts)
    exact, grouped = {}, {}
    for i, st in enumerate(steps):
        e = o.elt_from_step(st)
        kk = pkg.to_device(o.gen_galois_key(sk, 400 + i, e).reshape(o.kl - 1, 2, o.kl, o.n))
        exact[e] = kk
        grouped[e] = [be.key_prepare_grouped(kk, e, key_limbs, k_extra=k)]
    k_exact = be.make_keys(galois=exact)
    relin = pkg.to_device(o.gen_relin_key(sk, 5))
    grouped[0] = [be.key_prepare_grouped(relin.reshape(o.kl - 1, 2, o.kl, o.n), 0, key_limbs, k_extra=k, pre_permute=False)]
    k_grp = be.make_keys(grouped=grouped)
    many = be.rotate_many(k_grp, d, steps)
    for i, st in enumerate(steps):
        single = be.rotate_vector_keys(k_grp, d, st)
        assert (single == many[i]).all()
        got = decrypt_batch(o, sk, pkg, many[i])
        want = np.roll(zs, -st, axis=1)
        assert np.abs(got - want).max() < 1e-6, (st, np.abs(got - want).max())
    zy, cy = encrypt_batch(o, sk, rng, 2, limbs)
    prod3 = be.multiply(d, pkg.to_device(cy))
    if limbs >= 2:
        merged = be.relin_rescale_keys(k_grp, prod3)
        sc = SCALE * SCALE / float(o.q[limbs - 1])
        err = np.abs(decrypt_batch(o, sk, pkg, merged, sc) - zs * zy).max()
        assert err < 1e-6, err
    prof = be.profile_dump()
    be.profile(False)
    assert ("k_conv_quot" in prof) == mma, sorted(prof)   # the variant under test is the one that ran
    assert ("k_conv_quot_fp" in prof) == (os.environ.get("MOAI_CONV_FPSRC") == "1"), sorted(prof)


def test_bootstrap_grouped(pkg, env):
    """The whole bootstrapping on grouped keys (every Galois key prepared for the level it is used at, the
    relinearisation key in several variants): same tolerance as the SEAL-key pipeline, 2e-3 max-abs."""
    o, be, sk = env
    boot = pkg.Bootstrapper(be, total_limbs=17)
    boot.set_hoisting(True)
    levels = boot.required_step_levels()
    fast, grouped = {}, {}
    for i, (st, lvs) in enumerate(sorted(levels.items())):
        e = o.elt_from_step(st)
        kk = pkg.to_device(o.gen_galois_key(sk, 1000 + i, e).reshape(o.kl - 1, 2, o.kl, o.n))
        for lv in lvs:
            g = be.key_prepare_grouped(kk, e, lv)
            if g is None:
                fast.setdefault(e, []).append(be.key_prepare(kk, e, max_limbs=lv))
            else:
                grouped.setdefault(e, []).append(g)
    assert grouped and fast     # the top level has no spare prime, the lower ones do
    relin = pkg.to_device(o.gen_relin_key(sk, 5))
    relin4 = relin.reshape(o.kl - 1, 2, o.kl, o.n)
    rv = [be.key_prepare_grouped(relin4, 0, lv, pre_permute=False) for lv in range(16, 3, -1)]
    grouped[0] = [v for v in rv if v is not None]
    keys = be.make_keys(relin=relin, galois_fast=fast, grouped=grouped)
    rng = np.random.default_rng(1)
    B = 3
    zs = (rng.normal(size=(B, o.n // 2)) + 1j * rng.normal(size=(B, o.n // 2))) * 0.1
    zs[2] = 0.0
    cts = np.stack([o.encrypt_sym(sk, 50 + i, o.encode(zs[i], SCALE, 1), 1) for i in range(B)])
    out, out_scale = boot.bootstrap_3(keys, pkg.to_device(cts.reshape(B, 2, 1, o.n)), SCALE)
    assert out.shape[2] == 3 and out_scale == SCALE
    res = pkg.to_host(out)
    err = 0.0
    for i in range(B):
        dec = o.decode(o.decrypt(sk, res[i].reshape(-1), 2, 3), 3, out_scale)
        err = max(err, np.abs(dec - zs[i]).max())
    print("bootstrap on grouped keys: max |out - msg| = %.3g" % err)
    assert err < 2e-3, err


def _boot_keys(pkg, o, be, sk, boot, use_grouped, use_single):
    fast, grouped, single = {}, {}, {}
    for i, (st, lvs) in enumerate(sorted(boot.required_step_levels().items())):
        e = o.elt_from_step(st)
        kk = pkg.to_device(o.gen_galois_key(sk, 1000 + i, e).reshape(o.kl - 1, 2, o.kl, o.n))
        for lv in lvs:
            if lv == 0:     # baby step of the first CoeffToSlot stage (hoisting mode 2)
                if use_single:
                    single[e] = be.key_prepare_single(kk, e)
                else:
                    fast.setdefault(e, []).append(be.key_prepare(kk, e))
                continue
            g = be.key_prepare_grouped(kk, e, lv) if use_grouped else None
            if g is None:
                fast.setdefault(e, []).append(be.key_prepare(kk, e, max_limbs=lv))
            else:
                grouped.setdefault(e, []).append(g)
    relin = pkg.to_device(o.gen_relin_key(sk, 5))
    if use_grouped:
        relin4 = relin.reshape(o.kl - 1, 2, o.kl, o.n)
        grouped[0] = [be.key_prepare_grouped(relin4, 0, lv, k_extra=k, pre_permute=False)
                      for k, lv in sorted(be.ksg_plan(range(1, o.kl - 1)).items())]
    return be.make_keys(relin=relin, galois_fast=fast, grouped=grouped, single=single), (len(fast), len(grouped), len(single))


@pytest.mark.parametrize("use_grouped,use_single", [(True, True), (False, False), (False, True)])
def test_bootstrap_lazy_moddown(pkg, env, use_grouped, use_single):
    """Hoisting mode 2: baby-step rotations stay in the key-switch basis (one mod-down per giant step, k_bsgs_ext) and
    the first CoeffToSlot stage runs on single-digit keys with baby steps only (the mod-raised c1 is one small digit).
    Same tolerance as every other bootstrapping test (2e-3 max-abs); all three key configurations: grouped + single
    (what bench.py times), SEAL digits everywhere, SEAL digits + single."""
    o, be, sk = env
    boot = pkg.Bootstrapper(be, total_limbs=17)
    boot.set_hoisting(2)
    levels = boot.required_step_levels()
    assert any(0 in lv for lv in levels.values())          # first-stage baby steps want single-digit keys
    keys, counts = _boot_keys(pkg, o, be, sk, boot, use_grouped, use_single)
    rng = np.random.default_rng(1)
    B = 3
    zs = (rng.normal(size=(B, o.n // 2)) + 1j * rng.normal(size=(B, o.n // 2))) * 0.1
    zs[2] = 0.0
    cts = np.stack([o.encrypt_sym(sk, 50 + i, o.encode(zs[i], SCALE, 1), 1) for i in range(B)])
    be.profile(True)
    out, out_scale = boot.bootstrap_3(keys, pkg.to_device(cts.reshape(B, 2, 1, o.n)), SCALE)
    prof = be.profile_dump()
    be.profile(False)
    # grouped keys: the giant steps stay in the key-switch basis as well (k_giants_sum, one division per stage)
    assert ("k_giants_sum" in prof) == use_grouped, sorted(prof)
    assert out.shape[2] == 3 and out_scale == SCALE
    res = pkg.to_host(out)
    err = 0.0
    for i in range(B):
        dec = o.decode(o.decrypt(sk, res[i].reshape(-1), 2, 3), 3, out_scale)
        err = max(err, np.abs(dec - zs[i]).max())
    print("lazy bootstrapping (grouped=%s, single=%s; keys fast/grouped/single = %s): max |out - msg| = %.3g"
          % (use_grouped, use_single, counts, err))
    assert err < 2e-3, err
