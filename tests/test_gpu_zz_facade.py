"""The drop-in claim, end to end on the GPU: the reference's module headers, UNMODIFIED, compiled against
the header-only facade (include/facade/seal/seal.h -> include/moai_b200_seal.hpp) and bound to
libmoai_b200.so, against the same headers on the reference's real SEAL (oracle/_ref/libsealref.so) —
identical SEAL-generated keys and encryptions, every residue and all metadata bit-identical.  The case list
is the one tests/test_facade.py runs on the CPU test double (tests/facade_harness/cases.py).
softmax.hpp / Bootstrapper.h do not compile against stock SEAL here (NTL); through the facade they do, and the
bootstrapping path is checked by tolerance like tests/test_gpu_bootstrap.py (2e-3)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _driver(ref, fused=False):
    import facade_harness as facade
    if not facade.available(mock=False, fused=fused):
        pytest.skip("oracle/_ref/libfacade_driver*.so not built (needs /root/reference at build time)")
    d = facade.FacadeDriver(ref.log_n, bits=ref.bits, mock=False, fused=fused)
    assert d.lib.fd_ok(d.h)
    d.take_keys_from(ref)
    return d


@pytest.fixture(scope="module")
def small(sealref_small):
    return sealref_small, _driver(sealref_small)


@pytest.fixture(scope="module")
def deep(sealref_deep):
    return sealref_deep, _driver(sealref_deep)


def test_driver_is_bound_to_the_cuda_library(small):
    """The driver must resolve the C ABI in libmoai_b200.so, not in the CPU test double."""
    r, d = small
    assert d.lib.fd_backend_version() >= 100          # the test double answers -1
    assert "libmoai_b200.so" in open("/proc/self/maps").read()


def test_context_and_chain(small):
    from facade_harness import cases
    cases.case_context(*small)


def test_evaluator_ops_bit_exact(small):
    from facade_harness import cases
    cases.case_evaluator_ops(*small, np.random.default_rng(1))


def test_value_semantics_and_aliasing(small):
    from facade_harness import cases
    cases.case_value_semantics(*small, np.random.default_rng(4))


def test_seal_exception_rules(small):
    from facade_harness import cases
    cases.case_errors(*small, np.random.default_rng(2))


def test_decrypt_and_decode(small):
    from facade_harness import cases
    cases.case_decrypt_decode(*small, np.random.default_rng(3))


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_reference_ct_pt_matmul_header(small, variant):
    from facade_harness import cases
    cases.case_ct_pt(*small, np.random.default_rng(10 + variant), variant)


def test_reference_gelu_header(deep):
    from facade_harness import cases
    cases.case_gelu(*deep, np.random.default_rng(20))


@pytest.mark.parametrize("variant", [1, 2])
def test_reference_layernorm_header(deep, variant):
    from facade_harness import cases
    cases.case_layernorm(*deep, np.random.default_rng(30 + variant), variant)


@pytest.mark.parametrize("which", [0, 1])
def test_reference_ct_ct_matmul_header(deep, which):
    from facade_harness import cases
    cases.case_ct_ct(*deep, np.random.default_rng(40 + which), which)


def test_reference_softmax_header_exp_inverse(deep):
    from facade_harness import cases
    cases.case_exp_inverse(*deep, np.random.default_rng(50))


def test_bootstrapper_facade_preserves_message():
    """Bootstrapper.h of the facade with the reference driver's call sequence (test_full_scheme.hpp:413-448):
    decrypt(bootstrap_3(ct)) ~ decrypt(ct), max-abs slot error < 2e-3, output at L - 14 limbs, scale 2^46."""
    import facade_harness as facade
    from oracle import Oracle
    if not facade.available(mock=False):
        pytest.skip("oracle/_ref/libfacade_driver.so not built")
    bits = [51] + [46] * 2 + [51] * 14 + [58]        # same shape as the repo's chain, 17 data limbs
    o = Oracle(12, bits)
    d = facade.FacadeDriver(12, primes=o.q, mock=False)
    scale = 2.0 ** 46
    steps = d.boot_create(loge=10, logn=11, total_level=16, final_scale=scale)
    assert len(steps) > 0
    sk = o.gen_secret(3, hamming_weight=64)
    d.set_relin(o.gen_relin_key(sk, 5))
    for i, st in enumerate(sorted(set(steps + [0]))):
        e = o.elt_from_step(st)
        d.add_galois(e, o.gen_galois_key(sk, 1000 + i, e))
    rng = np.random.default_rng(1)
    z = (rng.normal(size=o.n // 2) + 1j * rng.normal(size=o.n // 2)) * 0.1
    ct = o.encrypt_sym(sk, 50, o.encode(z, scale, 1), 1)
    out, limbs, out_scale = d.bootstrap_3(ct.reshape(-1), scale, max_limbs=17)
    assert limbs == 3 and out_scale == scale
    dec = o.decode(o.decrypt(sk, out, 2, 3), 3, out_scale)
    assert np.abs(dec - z).max() < 2e-3, np.abs(dec - z).max()
    # the reference's argument checks (Bootstrapper.cpp:2939-2945)
    # the explicit batch overload bootstrap_3(vector, vector): three real-slot ciphertexts, one device call, two of
    # them sharing a bootstrapping (moai_bootstrap_real); every ciphertext comes back with its own message
    vs = rng.normal(size=(3, o.n // 2)) * 0.1
    cts = np.stack([o.encrypt_sym(sk, 80 + i, o.encode(vs[i].astype(np.complex128), scale, 1), 1) for i in range(3)])
    outs, calls = d.boot_combined(cts.reshape(-1), 3, scale, max_limbs=3, real_slots=True, max_batch=-1)
    assert calls == 1 and outs.shape == (3, 2, 3, o.n)
    for i in range(3):
        dec = o.decode(o.decrypt(sk, outs[i].reshape(-1), 2, 3), 3, scale)
        assert np.abs(dec - vs[i]).max() < 2e-3, (i, np.abs(dec - vs[i]).max())
    # the request combiner on the device: 12 ciphertexts, one bootstrap_3 call each from an OpenMP loop like the reference's
    # driver (test_full_scheme.hpp:654-660); concurrent callers are collected into batched, paired device calls
    import os
    vs = rng.normal(size=(12, o.n // 2)) * 0.1
    cts = np.stack([o.encrypt_sym(sk, 120 + i, o.encode(vs[i].astype(np.complex128), scale, 1), 1) for i in range(12)])
    outs, calls = d.boot_combined(cts.reshape(-1), 12, scale, max_limbs=3, real_slots=True, max_batch=64, linger_us=5000)
    assert outs.shape == (12, 2, 3, o.n) and 1 <= calls <= 12
    if min(os.cpu_count() or 1, int(os.environ.get("OMP_NUM_THREADS", "64"))) >= 4:
        assert calls < 12, "concurrent bootstrap_3 calls were not combined"
    for i in range(12):
        dec = o.decode(o.decrypt(sk, outs[i].reshape(-1), 2, 3), 3, scale)
        assert np.abs(dec - vs[i]).max() < 2e-3, (i, np.abs(dec - vs[i]).max())
    print("request combiner on the GPU: 12 bootstrap_3 calls from an OpenMP loop -> %d device calls" % calls)
    import ctypes as C
    two = np.zeros(2 * 2 * o.n, dtype=np.uint64)
    with pytest.raises(facade.FacadeError, match="lowest level"):
        d._chk(d.lib.fd_bootstrap_limbs(d.h, two.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int(2), C.c_double(scale)))


def test_thread_lanes_on_the_device():
    """One lane (CUDA stream + arena) per OpenMP thread on the B200: the reference-style parallel loop over a shared
    Evaluator gives the same residues as the mutex-serialised facade, bit for bit, and the decrypted values are right.
    With the 58-bit special prime of the repo's chain shape so that the grouped / fused key-switch kernels run."""
    import facade_harness as facade
    from oracle import Oracle
    if not facade.available(mock=False):
        pytest.skip("oracle/_ref/libfacade_driver.so not built")
    bits = [51] + [46] * 2 + [51] * 3 + [58]
    o = Oracle(12, bits)
    d = facade.FacadeDriver(12, primes=o.q, mock=False)
    sk = o.gen_secret(3, hamming_weight=64)
    d.set_relin(o.gen_relin_key(sk, 5))
    e = o.elt_from_step(1)
    d.add_galois(e, o.gen_galois_key(sk, 9, e))
    rng = np.random.default_rng(3)
    n_cts, limbs, scale = 24, 5, 2.0 ** 46
    zs = (rng.normal(size=(n_cts, o.n // 2)) + 1j * rng.normal(size=(n_cts, o.n // 2))) * 0.5
    x = np.stack([o.encrypt_sym(sk, 20 + i, o.encode(zs[i], scale, limbs), limbs) for i in range(n_cts)])
    serial, _ = d.parallel_chain(x.reshape(-1), n_cts, limbs, scale, lanes=False)
    serial, _ = d.parallel_chain(x.reshape(-1), n_cts, limbs, scale, lanes=False)     # timed after a warm-up pass
    ms_serial = d.last_loop_ms
    for rep in range(3):                      # races do not show every time
        lanes, threads = d.parallel_chain(x.reshape(-1), n_cts, limbs, scale, lanes=True)
        assert (serial == lanes).all(), rep
    ms_lanes = d.last_loop_ms
    sc = scale * scale / float(o.q[limbs - 1])
    for i in range(n_cts):
        dec = o.decode(o.decrypt(sk, lanes[i].reshape(-1), 2, limbs - 1), limbs - 1, sc)
        want = np.roll(zs[i] * zs[i] + zs[i] * zs[(i + 1) % n_cts], -1)
        assert np.abs(dec - want).max() < 1e-6
    print("thread lanes on the B200: %d OpenMP threads, %d ciphertexts, bit-identical to the serialised facade; "
          "loop %.2f ms with the mutex, %.2f ms with lanes" % (threads, n_cts, ms_serial, ms_lanes))


# ---- include/facade_fused first on the include path: the same module functions as fused device pipelines ----
@pytest.fixture(scope="module")
def small_fused(sealref_small):
    return sealref_small, _driver(sealref_small, fused=True)


@pytest.fixture(scope="module")
def deep_fused(sealref_deep):
    return sealref_deep, _driver(sealref_deep, fused=True)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_fused_ct_pt_matmul(small_fused, variant):
    from facade_harness import cases
    cases.case_ct_pt(*small_fused, np.random.default_rng(60 + variant), variant)


def test_fused_gelu(deep_fused):
    from facade_harness import cases
    cases.case_gelu(*deep_fused, np.random.default_rng(70))


@pytest.mark.parametrize("variant", [1, 2])
def test_fused_layernorm(deep_fused, variant):
    from facade_harness import cases
    r, d = deep_fused
    rng = np.random.default_rng(80 + variant)
    num_ct, limbs = 768, 21
    mask = np.zeros(r.n // 2, dtype=np.int32)
    mask[::16][:5] = 1
    x16, _ = cases.encrypt_batch(r, rng, 16, limbs, sigma=0.3, mask=mask)
    x = np.ascontiguousarray(np.tile(x16, (num_ct // 16, 1, 1, 1)))
    gamma, beta = rng.normal(size=num_ct), rng.normal(size=num_ct) * 0.1
    cases.same(d.layernorm(variant, x.reshape(-1), num_ct, limbs, cases.SCALE, gamma, beta, mask),
               r.layernorm(variant, x.reshape(-1), num_ct, limbs, cases.SCALE, gamma, beta, mask))


@pytest.mark.parametrize("which", [0, 1])
def test_fused_ct_ct_matmul(deep_fused, which):
    from facade_harness import cases
    cases.case_ct_ct(*deep_fused, np.random.default_rng(90 + which), which)


def test_fused_exp_inverse(deep_fused):
    from facade_harness import cases
    cases.case_exp_inverse(*deep_fused, np.random.default_rng(95))


# ---- client-side pieces: PRNG, seeded keys, wire format, Encryptor, batch_input ----
def _makers(mock):
    import facade_harness as facade
    from oracle import SealRef, have_ref
    from conftest import SMALL_BITS, SMALL_LOGN
    if not have_ref() or not facade.available(mock=mock):
        pytest.skip("oracle/_ref or the facade driver is not built")
    return (lambda seed: SealRef(SMALL_LOGN, SMALL_BITS, hamming_weight=0, seed=seed),
            lambda seed: facade.FacadeDriver(SMALL_LOGN, bits=SMALL_BITS, mock=mock, prng_seed=seed))


@pytest.mark.parametrize("case", ["case_prng", "case_key_wire_format", "case_encrypt", "case_ciphertext_wire_format",
                                  "case_batch_input", "case_keygen"])
def test_client_side(case):
    from facade_harness import cases
    if case == "case_batch_input" and not _ENCRYPT_OK:
        # the reference's batch_input encrypts inside `#pragma omp parallel for`: an exception there terminates
        # the process (the reference never catches), so it only runs once Encryptor::encrypt is known to work
        pytest.skip("Encryptor::encrypt did not pass")
    getattr(cases, case)(*_makers(mock=False))
    if case == "case_encrypt":
        globals()["_ENCRYPT_OK"] = True


_ENCRYPT_OK = False


def test_keygen_sparse_secret():
    import facade_harness as facade
    from facade_harness import cases
    from oracle import SealRef, have_ref
    from conftest import SMALL_BITS, SMALL_LOGN
    if not have_ref() or not facade.available(mock=False):
        pytest.skip("oracle/_ref or the facade driver is not built")
    cases.case_keygen_sparse(lambda seed: SealRef(SMALL_LOGN, SMALL_BITS, hamming_weight=64, seed=seed),
                             lambda seed: facade.FacadeDriver(SMALL_LOGN, bits=SMALL_BITS, mock=False, prng_seed=seed,
                                                              hamming_weight=64))


def test_reference_program_seal_ckks_test_runs_unmodified():
    """The reference's own test program SEAL_ckks_test() (M/test/test_SEAL_ckks.hpp:106-250) — KeyGenerator,
    CKKSEncoder, Encryptor, Evaluator (square, relinearize, rescale, multiply_plain, mod_switch, add), Decryptor,
    decode at N = 8192 with a {60, 40, 40, 60}-bit chain — compiled as it is against the facade (the driver includes
    the reference's whole M/include.hpp) and run: the vector it prints as computed equals the one it prints as
    expected, and the exact scales it prints are the ones stock SEAL prints for this example."""
    import re
    import facade_harness as facade
    if not facade.available(mock=False):
        pytest.skip("facade driver not built (needs /root/reference at build time)")
    ok, text = facade.reference_seal_ckks_test(mock=False)
    assert ok, text
    vecs = re.findall(r"\[ ([-0-9., ]+)\.\.\., ([-0-9., ]+) \]", text)
    assert len(vecs) == 3, text                                   # input, expected, computed
    nums = [[float(v) for v in (a + b).replace(" ", "").strip(",").split(",")] for a, b in vecs]
    expected, computed = np.array(nums[1]), np.array(nums[2])
    assert expected.shape == computed.shape == (6,)
    assert np.abs(expected - computed).max() < 1e-5, (expected, computed)
    assert abs(expected[-1] - 4.5415926) < 1e-6
    assert "Exact scale in PI*x^3: 1099512659965.7514648438" in text     # SEAL's own printed values
    assert "Exact scale in  0.4*x: 1099511775231.0197753906" in text
    assert "Modulus chain index for x3_encrypted: 0" in text and "coeff_modulus size: 200 (60 + 40 + 40 + 60) bits" in text


@pytest.mark.parametrize("log_n,bits", [(12, [40, 30, 30, 40]), (16, [51] + [46] * 20 + [51] * 14 + [58])])
def test_seed_expansion_on_device(pkg, log_n, bits):
    """moai_expand_seeds (csrc/seedexpand.cu): the uniform half of a seeded key digit / ciphertext regenerated on the
    device — BLAKE2Xb buffers in parallel, rejected words replaced in scan order — equals SEAL's
    sample_poly_uniform(Blake2xbPRNG(seed)) bit for bit (S/util/rlwe.cpp:137-166, S/randomgen.cpp:176-211; here through the
    facade's host restatement, which tests/test_facade.py pins to the real library), including the repo's chain with its
    58-bit special prime (~1000 rejections per polynomial)."""
    import facade_harness as facade
    if not facade.available(mock=False):
        pytest.skip("oracle/_ref/libfacade_driver*.so not built (needs /root/reference at build time)")
    d = facade.FacadeDriver(log_n, bits=bits, mock=False)
    be = pkg.Backend(log_n, [int(v) for v in d.q])
    rng = np.random.default_rng(5)
    seeds = rng.integers(0, 2 ** 63, size=(3, 8), dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    got = pkg.to_host(be.expand_seeds(seeds, d.kl)).reshape(3, -1)
    for i in range(3):
        exp = d.sample_uniform(seeds[i])
        assert (got[i] == exp).all(), (i, int((got[i] != exp).sum()))
    # fewer limbs than the key level (a ciphertext's c1 at its own level)
    got2 = pkg.to_host(be.expand_seeds(seeds[:1], 2)).reshape(-1)
    assert got2.shape[0] == 2 * (1 << log_n)
    be.close()
