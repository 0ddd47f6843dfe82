import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)  # tests/facade_harness (the name `tests` itself is taken by another package in this image)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


SMALL_BITS = [40, 30, 30, 30, 40]  # small chain used by the fast parity cases (logN = 12)
SMALL_LOGN = 12


@pytest.fixture(scope="session")
def oracle_small():
    from oracle import Oracle
    return Oracle(SMALL_LOGN, SMALL_BITS)


@pytest.fixture(scope="session")
def sealref_small():
    from oracle import SealRef, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    r = SealRef(SMALL_LOGN, SMALL_BITS, hamming_weight=0, seed=7)
    r.make_relin_key()
    r.make_galois_keys([1, 2, 4, -1, 256], conjugate=True)
    return r


def load_pkg():
    """The product package (its directory name is not a valid identifier, hence importlib)."""
    import importlib
    return importlib.import_module("moai-fhe-transformerinference-public_b200")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def backend_small(pkg, oracle_small):
    return pkg.Backend(SMALL_LOGN, oracle_small.q)


@pytest.fixture(scope="session")
def oracle_moai():
    from oracle import Oracle, MOAI_BITS
    return Oracle(16, MOAI_BITS)


@pytest.fixture(scope="session")
def backend_moai(pkg, oracle_moai):
    return pkg.Backend(16, oracle_moai.q)


DEEP_BITS = [40] + [30] * 20 + [40]  # 21 data primes: enough depth for LayerNorm (20 levels) at logN = 12


@pytest.fixture(scope="session")
def sealref_deep():
    """Real SEAL with a 21-level chain, relinearization key and every power-of-two Galois key
    (what KeyGenerator::create_galois_keys() without arguments provides, S/util/galois.cpp:106-131)."""
    from oracle import SealRef, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    r = SealRef(SMALL_LOGN, DEEP_BITS, hamming_weight=64, seed=21)
    r.make_relin_key()
    steps = []
    for k in range(0, 11):
        steps += [1 << k, -(1 << k)]
    r.make_galois_keys(steps, conjugate=True)
    return r


@pytest.fixture(scope="session")
def backend_deep(pkg, sealref_deep):
    return pkg.Backend(SMALL_LOGN, sealref_deep.q)


@pytest.fixture(scope="session")
def keys_deep(pkg, backend_deep, sealref_deep):
    r, be = sealref_deep, backend_deep
    gal = {e: pkg.to_device(r.export_galois_key(e)) for e in r.galois_elts()}
    return be.make_keys(relin=pkg.to_device(r.export_relin_key()), galois=gal)
