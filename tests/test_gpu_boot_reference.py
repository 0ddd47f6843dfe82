"""GPU bootstrapping, softmax_boot and single_att_block against THE REFERENCE'S OWN CODE.

oracle/_ref compiles the reference's Bootstrapper.cpp, ModularReducer.cpp, common/{Polynomial,Remez,func,Point}.cpp,
softmax.hpp and single_att_block.hpp unmodified (NTL::RR supplied over libmpfr by oracle/refbuild/ntl_shim, so the
reference's own Remez generates the EvalMod polynomial).  Both sides get the same SEAL-generated secret key,
relinearization / Galois keys and the same input ciphertexts; the outputs are decrypted with the reference's
Decryptor and compared slot by slot.

Ring degree: N = 8192.  The reference's `sfl_full_3` indexes a (2 * totlen3 + 1)-element vector up to totlen2
(Bootstrapper.cpp:2486-2488), which overruns the heap unless floor((logn - floor(logn / 3)) / 2) <= floor(logn / 3) + 1,
i.e. it only works for logn = 12 and logn = 15 of the sizes our kernels support; logn = 15 is the repo's size
(covered by test_gpu_fullsize.py against the message), logn = 12 is this file.

Bit-exact: ModRaise (integer).  By tolerance (results depend on FP64 polynomial / matrix coefficients and on which
key-switching keys are used): everything else; tolerances are stated in each test next to the measured value."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOG_N = 13
N = 1 << LOG_N
SLOTS = N // 2
NUM_BATCH = SLOTS // 128
SCALE = 2.0 ** 46
SMALL_BITS = [51] + [46] * 2 + [51] * 14 + [58]          # 17 data limbs: bootstrapping lands at chain_index 2
FULL_BITS = [51] + [46] * 20 + [51] * 14 + [58]          # the repo's chain (M/test/test_full_scheme.hpp:356-378)
TOK = 5


def _make_env(pkg, bits, extra_steps=(), att=False):
    """Reference side: SEAL context + keys + Bootstrapper built like the driver does; GPU side: Backend, exact-mode
    keys (SEAL layout) and fast-mode keys (pre-permuted, hoisted) derived from the SAME SEAL keys."""
    from oracle import SealRef, have_ref
    if not have_ref():
        pytest.fail("oracle/_ref is not built (run __graft_entry__.build() where /root/reference exists)")
    r = SealRef(LOG_N, bits, hamming_weight=64, seed=31)
    r.make_relin_key()
    ref_steps = r.boot_create()                          # runs the reference's Remez
    be = pkg.Backend(LOG_N, r.q)
    total = len(bits) - 1
    boot_exact = pkg.Bootstrapper(be, total_limbs=total)
    boot_fast = pkg.Bootstrapper(be, total_limbs=total)
    boot_fast.set_hoisting(True)
    steps = set(ref_steps) | set(boot_exact.required_steps()) | set(boot_fast.required_steps()) | set(extra_steps)
    att_steps = pkg.attention_rotation_steps(NUM_BATCH) if att else {}
    for tag in att_steps:
        steps |= set(att_steps[tag])
    steps.discard(0)
    r.make_galois_keys(sorted(steps), conjugate=True)
    r.boot_prepare()
    relin = pkg.to_device(r.export_relin_key())
    gal, fast = {}, {}
    fast_steps = set(boot_fast.required_steps()) | {0}
    for st in sorted(steps | {0}):
        e = r.elt_from_step(st)
        k = pkg.to_device(r.export_galois_key(e).reshape(r.kl - 1, 2, r.kl, r.n))
        gal[e] = k
        if st in fast_steps:
            fast.setdefault(e, []).append(be.key_prepare(k, e))
        for tag, level in (("qk", 14), ("sv", 3)):
            if st in att_steps.get(tag, ()):
                fast.setdefault(e, []).append(be.key_prepare(k, e, max_limbs=level))
    keys_exact = be.make_keys(relin=relin, galois=gal)
    keys_fast = be.make_keys(relin=relin, galois_fast=fast)
    return {"r": r, "be": be, "exact": (boot_exact, keys_exact), "fast": (boot_fast, keys_fast)}


@pytest.fixture(scope="module")
def env_small(pkg):
    return _make_env(pkg, SMALL_BITS)


@pytest.fixture(scope="module")
def env_full(pkg):
    pow2 = []
    for k in range(LOG_N - 1):
        pow2 += [1 << k, -(1 << k)]                      # keygen.create_galois_keys(gal_keys), test_full_scheme.hpp:399-400
    return _make_env(pkg, FULL_BITS, extra_steps=pow2, att=True)


def _encrypt(r, values, limbs):
    return r.encrypt(r.encode(np.asarray(values, dtype=np.complex128), SCALE, limbs), limbs, SCALE)


def _decrypt(r, ct, limbs, scale):
    ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1)
    return r.decode(r.decrypt(ct, 2, limbs, scale), limbs, scale)


def test_modraise_bit_exact_vs_reference(pkg, env_small):
    """moai_mod_raise == Bootstrapper::modraise_inplace (Bootstrapper.cpp:2938-2992), residue for residue."""
    r, be = env_small["r"], env_small["be"]
    rng = np.random.default_rng(3)
    z = (rng.normal(size=SLOTS) + 1j * rng.normal(size=SLOTS)) * 0.1
    ct = _encrypt(r, z, 1)
    exp, limbs, scale = r.boot_phase(0, ct, 1, SCALE)
    assert limbs == r.kl - 1 and scale == float(r.q[0])
    got = pkg.to_host(be.mod_raise(pkg.to_device(ct.reshape(1, 2, 1, N)), r.kl - 1))
    assert (got.reshape(-1) == exp).all()


@pytest.fixture(scope="module")
def boot_case(env_small):
    """One complex full-slot message and two real ones, bootstrapped by the reference (bootstrap_3,
    Bootstrapper.cpp:3496-3502)."""
    r = env_small["r"]
    rng = np.random.default_rng(11)
    zc = (rng.normal(size=SLOTS) + 1j * rng.normal(size=SLOTS)) * 0.1
    zr = rng.normal(size=(2, SLOTS)) * 0.1
    zr[0, :8] = [0.5, -0.5, 0.25, 0.0, 0.3, -0.3, 0.1, -0.1]
    msgs = [zc, zr[0].astype(np.complex128), zr[1].astype(np.complex128)]
    cts = [_encrypt(r, m, 1) for m in msgs]
    refs = []
    for ct in cts:
        out, limbs, scale = r.bootstrap_3(ct, SCALE)
        assert limbs == 3 and scale == SCALE             # chain_index 2 = total - 14, scale forced to final_scale (:3250)
        refs.append(_decrypt(r, out, limbs, scale))
    return msgs, cts, refs


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_bootstrap_3_vs_reference(pkg, env_small, boot_case, mode):
    """moai_bootstrap vs the reference's bootstrap_3 on the same ciphertext and keys: same output level (chain_index
    total - 14) and scale (2^46); decrypted outputs agree to 2e-5 max-abs per slot, and each side is within 2e-5 of
    the message.  Measured on a B200 (profiles/boot_reference_r2.log): |ref - msg| 2.9e-6 ... 3.2e-6, |gpu - msg|
    2.2e-6 ... 2.3e-6, |gpu - ref| 1.8e-6 ... 2.6e-6 in both modes."""
    r, be = env_small["r"], env_small["be"]
    boot, keys = env_small[mode]
    msgs, cts, refs = boot_case
    x = pkg.to_device(np.stack(cts).reshape(3, 2, 1, N))
    out, out_scale = boot.bootstrap_3(keys, x, SCALE)
    assert out.shape == (3, 2, 3, N) and out_scale == SCALE
    res = pkg.to_host(out)
    for i in range(3):
        dec = _decrypt(r, res[i], 3, out_scale)
        e_ref = np.abs(refs[i] - msgs[i]).max()
        e_gpu = np.abs(dec - msgs[i]).max()
        e_diff = np.abs(dec - refs[i]).max()
        print("bootstrap_3[%s] ct %d: |ref - msg| %.3g  |gpu - msg| %.3g  |gpu - ref| %.3g" % (mode, i, e_ref, e_gpu, e_diff))
        assert e_ref < 2e-5 and e_gpu < 2e-5 and e_diff < 2e-5, (mode, i, e_ref, e_gpu, e_diff)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_bootstrap_real_pairs_vs_reference(pkg, env_small, boot_case, mode):
    """moai_bootstrap_real (two real-slot ciphertexts per bootstrapping) vs the reference bootstrapping each of the
    two ciphertexts on its own: decrypted outputs agree to 2e-5 max-abs (measured 3.1e-6 ... 3.5e-6)."""
    r, be = env_small["r"], env_small["be"]
    boot, keys = env_small[mode]
    msgs, cts, refs = boot_case
    x = pkg.to_device(np.stack(cts[1:]).reshape(2, 2, 1, N))
    out, out_scale = boot.bootstrap_real(keys, x, SCALE)
    assert out.shape == (2, 2, 3, N) and out_scale == SCALE
    res = pkg.to_host(out)
    for i in range(2):
        dec = _decrypt(r, res[i], 3, out_scale)
        e_gpu = np.abs(dec - msgs[1 + i]).max()
        e_diff = np.abs(dec - refs[1 + i]).max()
        print("bootstrap_real[%s] ct %d: |gpu - msg| %.3g  |gpu - ref| %.3g" % (mode, i, e_gpu, e_diff))
        assert e_gpu < 2e-5 and e_diff < 2e-5, (mode, i, e_gpu, e_diff)


def _token_mask():
    mask = np.zeros(SLOTS, dtype=np.int32)
    for k in range(TOK):
        mask[k * NUM_BATCH:(k + 1) * NUM_BATCH] = 1
    return mask


@pytest.fixture(scope="module")
def softmax_case(env_full):
    """QK^T-shaped input (128 generalized diagonals, Ct_ct_matrix_mul.hpp:24-41) with scores 6 +- 0.25 on the TOK valid
    tokens, run through the reference's softmax_boot (softmax.hpp:308-581; layer 0: shift 7.5, 16 iterations).  The row
    sums of exp(S - 7.5) must stay inside (0, 2), the convergence domain of the reference's Goldschmidt inverse
    (softmax.hpp:49-82): one slot outside it overflows the whole ciphertext, in the reference as well."""
    r = env_full["r"]
    rng = np.random.default_rng(5)
    S = np.zeros((NUM_BATCH, 128, 128))
    S[:, :TOK, :TOK] = np.clip(6.0 + rng.normal(size=(NUM_BATCH, TOK, TOK)) * 0.25, 5.2, 6.8)
    sums = ((1 + (S[:, :TOK, :TOK] - 7.5) / 128.0) ** 128).sum(axis=2)
    assert 0.5 < sums.min() and sums.max() < 1.8, "test data outside the inverse's convergence domain"
    limbs = 13                                           # chain_index 12
    cts = np.empty((128, 2, limbs, N), dtype=np.uint64)
    for i in range(128):
        v = np.zeros((128, NUM_BATCH))
        for k in range(128):
            v[k] = S[:, k, (k + i) % 128]
        cts[i] = _encrypt(r, v.reshape(-1), limbs).reshape(2, limbs, N)
    mask = _token_mask()
    out, ol, osc = r.softmax_boot(cts, 128, limbs, SCALE, mask, TOK, 16, 0)
    ref = np.stack([_decrypt(r, out.reshape(128, -1)[i], ol, osc).real for i in range(128)])
    return S, cts, mask, ref, ol, osc


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_softmax_boot_vs_reference(pkg, env_full, softmax_case, mode):
    """moai_softmax_boot vs the reference header: same output level and scale, decrypted rows agree to 2e-5 max-abs
    (values are probabilities in [0, 1]; measured 1.7e-6 ... 1.8e-6); both are 4.1e-4 away from the float64 model of the
    same approximations (bound 1e-3): the bootstrapping error on the row sums, amplified by the inverse."""
    r, be = env_full["r"], env_full["be"]
    boot, keys = env_full[mode]
    S, cts, mask, ref, ol, osc = softmax_case
    out, out_scale = boot.softmax_boot(keys, pkg.to_device(cts), SCALE, mask, TOK, iters=16, layer_id=0)
    assert out.shape == (128, 2, ol, N) and out_scale == osc
    res = pkg.to_host(out)
    got = np.stack([_decrypt(r, res[i], ol, out_scale).real for i in range(128)])
    E = (1 + (S[:, :TOK, :TOK] - 7.5) / 128.0) ** 128
    P = E / (E.sum(axis=2, keepdims=True) + 1e-5)
    model = np.zeros((128, 128, NUM_BATCH))              # [diagonal i][row k][input b]
    for i in range(128):
        for k in range(TOK):
            if (k + i) % 128 < TOK:
                model[i, k] = P[:, k, (k + i) % 128]
    model = model.reshape(128, -1)
    e_diff = np.abs(got - ref).max()
    e_ref = np.abs(ref - model).max()
    e_gpu = np.abs(got - model).max()
    print("softmax_boot[%s]: |gpu - ref| %.3g  |ref - model| %.3g  |gpu - model| %.3g" % (mode, e_diff, e_ref, e_gpu))
    assert e_diff < 2e-5 and e_gpu < 1e-3 and e_ref < 1e-3, (mode, e_diff, e_ref, e_gpu)


@pytest.fixture(scope="module")
def att_case(env_full):
    """One attention head at reduced widths (hidden 48, head width 8; the pipeline is width-agnostic) through the
    reference's single_att_block (single_att_block.hpp:10-207)."""
    r = env_full["r"]
    rng = np.random.default_rng(1)
    hidden, col_W = 48, 8
    X = np.zeros((128, NUM_BATCH, hidden))
    X[:TOK] = rng.normal(size=(TOK, NUM_BATCH, hidden)) * 0.5
    WQ, WK = (rng.normal(size=(hidden, col_W)) * 0.02 for _ in range(2))
    WV = rng.normal(size=(hidden, col_W)) * 0.12
    bQ = np.full(col_W, np.sqrt(6.0 / col_W)) + rng.normal(size=col_W) * 0.02
    bK = np.full(col_W, np.sqrt(6.0 / col_W)) + rng.normal(size=col_W) * 0.02
    bV = rng.normal(size=col_W) * 0.1
    # scores Q.K ~ 6 +- 0.25: the row sums of exp(S - 7.5) must stay inside (0, 2) (see softmax_case)
    Q = np.einsum("tbh,hc->tbc", X[:TOK], WQ) + bQ
    Kk = np.einsum("tbh,hc->tbc", X[:TOK], WK) + bK
    sums = ((1 + (np.einsum("tbc,ubc->btu", Q, Kk) - 7.5) / 128.0) ** 128).sum(axis=2)
    assert 0.5 < sums.min() and sums.max() < 1.8, ("test data outside the inverse's convergence domain", sums.min(), sums.max())
    limbs = 15                                           # chain_index 14 (test_full_scheme.hpp:496-507)
    cts = np.empty((hidden, 2, limbs, N), dtype=np.uint64)
    for c in range(hidden):
        cts[c] = _encrypt(r, X[:, :, c].reshape(-1), limbs).reshape(2, limbs, N)
    mask = _token_mask()
    out, cnt, ol, osc = r.single_att_block(cts, hidden, limbs, SCALE, WQ, WK, WV, bQ, bK, bV, mask, TOK, NUM_BATCH, 16, 0)
    assert cnt == col_W
    ref = np.stack([_decrypt(r, out.reshape(cnt, -1)[i], ol, osc).real for i in range(cnt)])
    return cts, (WQ, WK, WV, bQ, bK, bV), mask, ref, ol, osc


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_single_att_block_vs_reference(pkg, env_full, att_case, mode):
    """moai_single_att_block vs the reference header on the same ciphertexts and keys: same output level / scale,
    decrypted head outputs agree to 3e-5 max-abs (outputs are O(0.6); measured 3.0e-6 ... 3.5e-6)."""
    r, be = env_full["r"], env_full["be"]
    boot, keys = env_full[mode]
    cts, (WQ, WK, WV, bQ, bK, bV), mask, ref, ol, osc = att_case
    out, out_scale = boot.single_att_block(keys, pkg.to_device(cts), SCALE, WQ, WK, WV, bQ, bK, bV, mask, TOK, NUM_BATCH,
                                           iters=16, layer_id=0)
    assert out.shape == (WQ.shape[1], 2, ol, N) and out_scale == osc
    res = pkg.to_host(out)
    got = np.stack([_decrypt(r, res[i], ol, out_scale).real for i in range(res.shape[0])])
    e_diff = np.abs(got - ref).max()
    print("single_att_block[%s]: |gpu - ref| %.3g  (|ref| max %.3g)" % (mode, e_diff, np.abs(ref).max()))
    assert e_diff < 3e-5, (mode, e_diff)
