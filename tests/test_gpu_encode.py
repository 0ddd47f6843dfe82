"""GPU parity of the device CKKS vector encoder and of the exact masked
ct-pt matmul (B3).  Bit-exact vs the oracle, which is itself pinned bit-exact vs real SEAL."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rand_cts(o, rng, count, limbs):
    out = np.empty((count, 2, limbs, o.n), dtype=np.uint64)
    for l in range(limbs):
        out[:, :, l, :] = rng.integers(0, int(o.q[l]), (count, 2, o.n), dtype=np.uint64)
    return out


@pytest.mark.parametrize("limbs,scale", [(4, 2.0 ** 30), (2, 2.0 ** 50), (3, 2.0 ** 70)])
def test_encode_small(pkg, backend_small, oracle_small, limbs, scale):
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(limbs)
    z = rng.normal(size=(3, o.n // 2)) + 1j * rng.normal(size=(3, o.n // 2))
    z[1] = 0.0
    z[2, 5:] = 0.0
    got = pkg.to_host(be.encode(z, scale, limbs))
    for i in range(3):
        assert (got[i].reshape(-1) == o.encode(z[i], scale, limbs)).all(), i
    short = rng.normal(size=100)                      # fewer values than slots, real input
    assert (pkg.to_host(be.encode(short, scale, limbs)).reshape(-1) == o.encode(short, scale, limbs)).all()
    const = np.full(o.n // 2, -0.37)                  # constant vector == scalar encode
    assert (pkg.to_host(be.encode(const, scale, limbs)).reshape(-1) == o.encode_scalar(-0.37, scale, limbs)).all()


def test_encode_moai_params(pkg, backend_moai, oracle_moai):
    o, be = oracle_moai, backend_moai
    rng = np.random.default_rng(7)
    z = (rng.normal(size=(2, o.n // 2)) + 1j * rng.normal(size=(2, o.n // 2))) * 0.3
    for limbs in (35, 2):
        got = pkg.to_host(be.encode(z, 2.0 ** 46, limbs))
        for i in range(2):
            assert (got[i].reshape(-1) == o.encode(z[i], 2.0 ** 46, limbs)).all(), (limbs, i)


def test_encode_vs_real_seal(pkg, backend_small, sealref_small):
    r, be = sealref_small, backend_small
    rng = np.random.default_rng(9)
    z = rng.normal(size=r.n // 2) + 1j * rng.normal(size=r.n // 2)
    assert (pkg.to_host(be.encode(z, 2.0 ** 30, 3)).reshape(-1) == r.encode(z, 2.0 ** 30, 3)).all()


def test_masked_matmul_small(pkg, backend_small, oracle_small):
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(11)
    K, C, limbs, scale = 7, 5, 3, 2.0 ** 30
    X = rand_cts(o, rng, K, limbs)
    W = rng.normal(size=(K, C)) * 0.2
    mask = (rng.random(o.n // 2) < 0.4).astype(np.int32)
    got = pkg.to_host(be.ct_pt_matrix_mul_wo_pre_w_mask(pkg.to_device(X), W, mask, scale))
    exp = o.ct_pt_matmul_masked(X.reshape(-1), W, mask, K, C, limbs, scale)
    assert (got.reshape(-1) == exp).all()
    # all-ones mask takes the scalar kernel and must agree with the masked oracle too
    ones = np.ones(o.n // 2, dtype=np.int32)
    got = pkg.to_host(be.ct_pt_matrix_mul_wo_pre_w_mask(pkg.to_device(X), W, ones, scale))
    assert (got.reshape(-1) == o.ct_pt_matmul_masked(X.reshape(-1), W, ones, K, C, limbs, scale)).all()


def test_masked_matmul_moai_reference_mask(pkg, backend_moai, oracle_moai):
    """The reference run's mask: input 0 has 5 tokens, the other 255 inputs are empty
    (M/test/test_full_scheme.hpp:455-457, bias_vec of Batch_encode_encrypt.hpp:40-49)."""
    o, be = oracle_moai, backend_moai
    rng = np.random.default_rng(13)
    K, C, limbs, scale = 6, 3, 2, 2.0 ** 46
    X = rand_cts(o, rng, K, limbs)
    W = rng.normal(size=(K, C)) * 0.04
    mask = np.zeros(o.n // 2, dtype=np.int32)
    for tok in range(5):
        mask[tok * 256 + 0] = 1
    got = pkg.to_host(be.ct_pt_matrix_mul_wo_pre_w_mask(pkg.to_device(X), W, mask, scale))
    exp = o.ct_pt_matmul_masked(X.reshape(-1), W, mask, K, C, limbs, scale)
    assert (got.reshape(-1) == exp).all()
