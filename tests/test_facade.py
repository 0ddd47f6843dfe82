"""Host logic of the header-only C++ facade (include/moai_b200_seal.hpp, include/facade/) — CPU only.

The reference's module headers, UNMODIFIED, are compiled against the facade (tests/facade_harness/facade_driver.cpp)
and run on the CPU test double of the C ABI (tests/facade_harness/mock_cabi.c, a forwarder to the C oracle); the same
headers run on the reference's real SEAL (oracle/_ref/libsealref.so).  Identical SEAL-generated keys and
encryptions go to both; all residues and all metadata must match bit for bit.  What this pins is the facade:
metadata bookkeeping, SEAL's checks and exception classes, operation sequencing, value semantics, the
encoder/decoder glue.  The CUDA library itself is pinned by the `-m gpu` tests (tests/test_gpu_zz_facade.py runs
this same list on libmoai_b200.so)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _driver(ref, mock=True, fused=False):
    import facade_harness as facade
    if not facade.available(mock=mock, fused=fused):
        pytest.skip("facade driver not built (needs /root/reference at build time)")
    d = facade.FacadeDriver(ref.log_n, bits=ref.bits, mock=mock, fused=fused)
    d.take_keys_from(ref)
    return d


@pytest.fixture(scope="module")
def small(sealref_small):
    return sealref_small, _driver(sealref_small)


@pytest.fixture(scope="module")
def deep(sealref_deep):
    return sealref_deep, _driver(sealref_deep)


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference's headers")
def test_reference_headers_compile_unchanged_against_the_facade(tmp_path):
    """The reference's WHOLE source — M/include.hpp: every module header and every test / driver program, including
    all_layer_test (M/test/test_full_scheme.hpp) — compiles as it is with include/facade first on the include path;
    also with include/facade_fused in front (module functions bound to the fused pipelines).  softmax.hpp,
    single_att_block.hpp and the drivers do NOT compile against stock SEAL in this image (NTL through Bootstrapper.h)."""
    src = tmp_path / "tu.cpp"
    # M/include.hpp is the reference's umbrella header: seal/seal.h, every module header under M/source/ and every
    # test / driver header under M/test/ (all_layer_test, the matmul tests, SEAL_ckks_test)
    src.write_text("""
#include "include.hpp"
int main() { all_layer_test(); return 0; }
""")
    inc = os.path.join(ROOT, "include")
    for extra in ([], ["-I" + os.path.join(inc, "facade_fused")]):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-fopenmp", "-w"] + extra +
                              ["-I" + os.path.join(inc, "facade"), "-I" + inc, "-I" + os.path.join(REF, "include"), str(src)])


def test_facade_fails_loudly_without_a_device():
    """No CPU fallback: bound to the real library in a container without a GPU, SEALContext's constructor
    throws (the driver reports it); it never computes anything on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import facade_harness as facade
    if not facade.available(mock=False):
        pytest.skip("facade driver not built")
    with pytest.raises(facade.FacadeError):
        facade.FacadeDriver(12, bits=[40, 30, 30, 40], mock=False)


def test_context_and_chain(small):
    from facade_harness import cases
    assert small[1].lib.fd_backend_version() == -1      # these tests run on the CPU test double of the C ABI
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


def test_reference_layernorm_header(deep):
    from facade_harness import cases
    cases.case_layernorm(*deep, np.random.default_rng(31), 1)


@pytest.mark.parametrize("which", [0, 1])
def test_reference_ct_ct_matmul_header(deep, which):
    from facade_harness import cases
    cases.case_ct_ct(*deep, np.random.default_rng(40 + which), which)


def test_reference_softmax_header_exp_inverse(deep):
    from facade_harness import cases
    cases.case_exp_inverse(*deep, np.random.default_rng(50))


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_fused_ct_pt_matmul_shadow_header(sealref_small, variant):
    """include/facade_fused first on the include path: `source/matrix_mul/Ct_pt_matrix_mul.hpp` resolves to the
    fused pipeline wrapper (include/moai_b200_fused_modules.hpp); same call, same residues.  Exercises the
    pack / one-call / unpack glue all fused wrappers share."""
    from facade_harness import cases
    cases.case_ct_pt(sealref_small, _driver(sealref_small, fused=True), np.random.default_rng(60 + variant), variant)


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
    getattr(cases, case)(*_makers(mock=True))


def test_keygen_sparse_secret():
    import facade_harness as facade
    from facade_harness import cases
    from oracle import SealRef, have_ref
    from conftest import SMALL_BITS, SMALL_LOGN
    if not have_ref() or not facade.available(mock=True):
        pytest.skip("oracle/_ref or the facade driver is not built")
    cases.case_keygen_sparse(lambda seed: SealRef(SMALL_LOGN, SMALL_BITS, hamming_weight=64, seed=seed),
                             lambda seed: facade.FacadeDriver(SMALL_LOGN, bits=SMALL_BITS, mock=True, prng_seed=seed,
                                                              hamming_weight=64))


def test_reference_program_seal_ckks_test_runs_unmodified():
    """The reference's own test program SEAL_ckks_test() (M/test/test_SEAL_ckks.hpp:106-250) — KeyGenerator,
    CKKSEncoder, Encryptor, Evaluator (square, relinearize, rescale, multiply_plain, mod_switch, add), Decryptor,
    decode at N = 8192 with a {60, 40, 40, 60}-bit chain — compiled as it is against the facade (the driver includes
    the reference's whole M/include.hpp) and run: the vector it prints as computed equals the one it prints as
    expected, and the exact scales it prints are the ones stock SEAL prints for this example."""
    import re
    import facade_harness as facade
    if not facade.available(mock=True):
        pytest.skip("facade driver not built (needs /root/reference at build time)")
    ok, text = facade.reference_seal_ckks_test(mock=True)
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


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference's sources")
def test_reference_main_program_builds_unmodified_and_fails_loudly_without_gpu(tmp_path):
    """The reference's main program (test.cpp -> batch_input_test, ct_pt_matrix_mul_test, ct_ct_matrix_mul_test,
    all_layer_test) compiled and LINKED as it is against the facade (fused module headers first) and libmoai_b200.so.
    In this container there is no GPU: the executable must stop at its first SEALContext with the library's
    "no CPU fallback" error instead of computing anything on the host."""
    import torch
    inc = os.path.join(ROOT, "include")
    pkg = os.path.join(ROOT, "moai-fhe-transformerinference-public_b200")
    exe = str(tmp_path / "moai_reference_main")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O0", "-fopenmp", "-w", "-I" + os.path.join(inc, "facade_fused"),
                           "-I" + os.path.join(inc, "facade"), "-I" + inc, "-I" + os.path.join(REF, "include"),
                           os.path.join(REF, "test.cpp"), "-L" + pkg, "-lmoai_b200", "-Wl,-rpath," + pkg, "-o", exe])
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the program would run the full 12-layer workload")
    res = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, timeout=60)
    assert res.returncode != 0
    assert "BATCH ENCODE ENCRYPT" in res.stdout
    assert "no CPU fallback" in res.stderr


def test_thread_lanes_match_the_serialised_facade():
    """SEALContext::set_thread_lanes(true): every OpenMP thread works on its own lane of the context
    (moai_context_fork: its own stream and arena) instead of queueing on one mutex.  A chain that creates, hands over
    and frees ciphertexts across calls (square, multiply with the neighbour, relinearize, add, rescale, rotate) from an
    OpenMP loop gives the same residues, bit for bit, as the serialised facade — here on the CPU test double of the
    C ABI, in tests/test_gpu_zz_facade.py on the B200."""
    import facade_harness as facade
    from oracle import Oracle
    if not facade.available(mock=True):
        pytest.skip("facade driver not built")
    o = Oracle(12, [40, 30, 30, 30, 40])
    d = facade.FacadeDriver(12, primes=o.q, mock=True)
    sk = o.gen_secret(3)
    d.set_relin(o.gen_relin_key(sk, 5))
    e = o.elt_from_step(1)
    d.add_galois(e, o.gen_galois_key(sk, 9, e))
    rng = np.random.default_rng(3)
    n_cts, limbs, scale = 12, 3, 2.0 ** 30
    zs = (rng.normal(size=(n_cts, o.n // 2)) + 1j * rng.normal(size=(n_cts, o.n // 2))) * 0.5
    x = np.stack([o.encrypt_sym(sk, 20 + i, o.encode(zs[i], scale, limbs), limbs) for i in range(n_cts)])
    serial, _ = d.parallel_chain(x.reshape(-1), n_cts, limbs, scale, lanes=False)
    lanes, threads = d.parallel_chain(x.reshape(-1), n_cts, limbs, scale, lanes=True)
    assert (serial == lanes).all()
    assert threads >= 1
    # and the chain computes what it says: rot_1(x_i^2 + x_i x_{i+1}) at scale^2 / q_2
    sc = scale * scale / float(o.q[limbs - 1])
    for i in range(n_cts):
        dec = o.decode(o.decrypt(sk, lanes[i].reshape(-1), 2, limbs - 1), limbs - 1, sc)
        want = np.roll(zs[i] * zs[i] + zs[i] * zs[(i + 1) % n_cts], -1)
        assert np.abs(dec - want).max() < 1e-3          # 30-bit scale


@pytest.mark.parametrize("real_slots", [True, False])
def test_bootstrapper_request_combining_routes_every_ciphertext(real_slots):
    """Opt-in combining of concurrent bootstrap_3 calls (include/facade/Bootstrapper.h): 40 ciphertexts bootstrapped from
    an OpenMP loop like the reference's driver.  On the CPU test double the "bootstrapping" is a plumbing fake whose
    result depends on the ciphertext's own input only — so this checks the host logic: every caller gets ITS result,
    requests are actually combined (fewer device calls than ciphertexts), nothing deadlocks with many threads, small
    batches and a zero linger."""
    import facade_harness as facade
    from oracle import Oracle
    if not facade.available(mock=True):
        pytest.skip("facade driver not built")
    bits = [40] + [30] * 16 + [40]                      # 17 data limbs: the bootstrapper's level budget needs >= 15
    o = Oracle(10, bits)
    rng = np.random.default_rng(7)
    n_cts, L = 40, 17 - 14
    x = rng.integers(0, int(o.q[0]), (n_cts, 2, 1, o.n), dtype=np.uint64)
    want = np.stack([x[:, :, 0, :] % np.uint64(o.q[l]) for l in range(L)], axis=2)
    for max_batch, linger in ((64, 2000), (4, 0), (1, 0), (-1, 0)):      # -1: the explicit vector overload of bootstrap_3
        d = facade.FacadeDriver(10, primes=o.q, mock=True)
        d.boot_create(loge=10, logn=9, total_level=16, final_scale=2.0 ** 30)
        d.set_relin(np.zeros((o.kl - 1) * 2 * o.kl * o.n, dtype=np.uint64))
        got, calls = d.boot_combined(x.reshape(-1), n_cts, 2.0 ** 30, max_limbs=L, real_slots=real_slots,
                                     max_batch=max_batch, linger_us=linger)
        assert got.shape == want.shape and (got == want).all()
        assert 1 <= calls <= n_cts
        if max_batch == -1:
            assert calls == 1
        if max_batch == 1:
            assert calls == n_cts
        if max_batch == 64 and (os.cpu_count() or 1) > 1:
            assert calls < n_cts                          # concurrent callers were actually combined


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference's headers")
def test_example_server_program_end_to_end(sealref_small, tmp_path):
    """examples/server_ct_pt_matmul.cpp — the deployment split: a stock-SEAL client writes parameters, encrypted inputs
    (Ciphertext::save streams) and weights to files; the server program (facade + the reference's module header name,
    here linked against the CPU test double) writes encrypted outputs; the client reads them back with SEAL's own
    loader.  The outputs equal the reference's ct_pt_matrix_mul_wo_pre on real SEAL bit for bit."""
    import facade_harness as facade
    from facade_harness import cases
    from conftest import SMALL_BITS, SMALL_LOGN
    if not facade.available(mock=True):
        pytest.skip("facade test double not built")
    r = sealref_small
    inc = os.path.join(ROOT, "include")
    ref_out = os.path.join(ROOT, "oracle", "_ref")
    exe = str(tmp_path / "server")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-fopenmp", "-w", "-I" + os.path.join(inc, "facade_fused"),
                           "-I" + os.path.join(inc, "facade"), "-I" + inc, "-I" + os.path.join(REF, "include"),
                           os.path.join(ROOT, "examples", "server_ct_pt_matmul.cpp"), "-L" + ref_out, "-lmoai_b200_mock",
                           "-Wl,-rpath," + ref_out, "-o", exe])
    rng = np.random.default_rng(77)
    K, Cc, limbs = 4, 3, 2
    X, _ = cases.encrypt_batch(r, rng, K, limbs)
    W = rng.normal(size=(K, Cc)) * 0.3
    (tmp_path / "params.txt").write_text("%d %s\n" % (SMALL_LOGN, " ".join(str(b) for b in SMALL_BITS)))
    (tmp_path / "weights.txt").write_text("%d %d\n%s\n" % (K, Cc, " ".join(repr(float(w)) for w in W.reshape(-1))))
    with open(tmp_path / "inputs.seal", "wb") as f:
        for j in range(K):
            f.write(r.save_ciphertext(X[j].reshape(-1), 2, limbs, cases.SCALE))
    res = subprocess.run([exe, "params.txt", "inputs.seal", "weights.txt", "outputs.seal"], cwd=str(tmp_path),
                         capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    assert "3 ciphertexts at chain_index 0" in res.stdout
    exp, _ = r.ct_pt_matmul(0, X.reshape(-1), W, None, K, Cc, limbs, cases.SCALE)
    blob = (tmp_path / "outputs.seal").read_bytes()
    per = len(blob) // Cc
    for i in range(Cc):
        ct, size, l, scale = r.load_ciphertext(blob[i * per:(i + 1) * per])     # SEAL's own loader, validity checks on
        assert (size, l, scale) == (2, limbs - 1, cases.SCALE)
        assert (ct == exp.reshape(Cc, -1)[i]).all()
