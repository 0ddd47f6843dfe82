"""SEAL wire format against the reference's real library (oracle/_ref): the bytes
our serializer writes are byte-identical to Ciphertext::save(compr_mode_type::none), SEAL's own
Ciphertext::load (with its validity checks, incl. the parms_id) accepts them, and our loader reads SEAL's."""
import importlib

import numpy as np
import pytest

from conftest import SMALL_BITS, SMALL_LOGN


@pytest.fixture(scope="module")
def sealio():
    return importlib.import_module("moai-fhe-transformerinference-public_b200.sealio")


def encrypt(r, rng, limbs, scale):
    top = r.kl - 1
    v = rng.normal(size=r.n // 2)
    ct = r.encrypt(r.encode_real(v, scale, top), top, scale).reshape(2, top, r.n)[:, :limbs, :]
    return v, np.ascontiguousarray(ct)


@pytest.mark.parametrize("limbs", [4, 2, 1])
def test_save_matches_seal_bytes_and_seal_loads_ours(sealio, sealref_small, limbs):
    r = sealref_small
    rng = np.random.default_rng(limbs)
    scale = 2.0 ** 30
    _, ct = encrypt(r, rng, limbs, scale)
    ref_bytes = r.save_ciphertext(ct.reshape(-1), 2, limbs, scale)
    ours = sealio.save_ciphertext(ct, [int(q) for q in r.q[:limbs]], scale)
    assert ours == ref_bytes
    raw, size, l2, s2 = r.load_ciphertext(ours)                  # SEAL's loader, full validity checks
    assert size == 2 and l2 == limbs and s2 == scale and (raw == ct.reshape(-1)).all()


def test_load_reads_seal_bytes_and_checks_parms_id(sealio, sealref_small):
    r = sealref_small
    rng = np.random.default_rng(9)
    scale = 2.0 ** 30
    _, ct = encrypt(r, rng, 3, scale)
    blob = r.save_ciphertext(ct.reshape(-1), 2, 3, scale)
    got = sealio.load_ciphertext(blob, chain_primes=[int(q) for q in r.q[:-1]])
    assert (got["residues"] == ct).all() and got["scale"] == scale and got["is_ntt_form"]
    with pytest.raises(ValueError):                               # same bytes, different prime chain
        sealio.load_ciphertext(blob, chain_primes=[int(q) + 2 for q in r.q[:-1]])
    with pytest.raises(ValueError):                               # truncated stream
        sealio.load_ciphertext(blob[:-8])


def test_size3_ciphertext_round_trip(sealio, sealref_small):
    r = sealref_small
    rng = np.random.default_rng(11)
    raw = np.stack([rng.integers(0, int(r.q[l]), (3, r.n), dtype=np.uint64) for l in range(2)], axis=1)  # [3, 2, n]
    blob = sealio.save_ciphertext(raw, [int(q) for q in r.q[:2]], 2.0 ** 60)
    assert blob == r.save_ciphertext(raw.reshape(-1), 3, 2, 2.0 ** 60)
    assert (sealio.load_ciphertext(blob)["residues"] == raw).all()
