"""GPU parity of the fused ct-pt matmul (BASELINE config 1) against the
oracle and, when oracle/_ref is present, against the reference's unmodified module header run on
real SEAL.  Bit-exact (integer residues)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rand_cts(o, rng, count, limbs):
    out = np.empty((count, 2, limbs, o.n), dtype=np.uint64)
    for l in range(limbs):
        out[:, :, l, :] = rng.integers(0, int(o.q[l]), (count, 2, o.n), dtype=np.uint64)
    return out


@pytest.mark.parametrize("K,C,limbs", [(5, 11, 3), (1, 1, 2), (300, 8, 4), (70, 45, 3)])
def test_matmul_small_vs_oracle(pkg, backend_small, oracle_small, K, C, limbs):
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(K * 100 + C)
    X = rand_cts(o, rng, K, limbs)
    W = rng.normal(size=(K, C)) * 0.3
    W[0, 0] = 0.0
    W[-1, -1] = -1.0
    scale = 2.0 ** 30
    got = pkg.to_host(be.ct_pt_matrix_mul_wo_pre(pkg.to_device(X), W, scale))
    exp = o.ct_pt_matmul_scalar(X.reshape(-1), W, K, C, limbs, scale)
    assert (got.reshape(-1) == exp).all()


def test_matmul_small_vs_reference_module(pkg, backend_small, oracle_small, sealref_small):
    """Same inputs through the reference's ct_pt_matrix_mul_wo_pre / _large / _w_mask (all-ones mask)."""
    o, be, r = oracle_small, backend_small, sealref_small
    rng = np.random.default_rng(3)
    K, C, limbs, scale = 6, 128, 3, 2.0 ** 30
    X = rand_cts(o, rng, K, limbs)
    W = rng.normal(size=(K, C)) * 0.2
    got = pkg.to_host(be.ct_pt_matrix_mul_wo_pre(pkg.to_device(X), W, scale)).reshape(-1)
    for variant, mask in ((0, None), (1, None), (2, np.ones(o.n // 2, dtype=np.int32))):
        exp, _ = r.ct_pt_matmul(variant, X.reshape(-1), W, mask, K, C, limbs, scale)
        assert (got == exp).all(), variant


@pytest.mark.parametrize("K,C,limbs", [(48, 10, 2), (16, 8, 15)])
def test_matmul_moai_params_vs_oracle(pkg, backend_moai, oracle_moai, K, C, limbs):
    o, be = oracle_moai, backend_moai
    rng = np.random.default_rng(limbs)
    X = rand_cts(o, rng, K, limbs)
    W = rng.normal(size=(K, C)) * 0.04
    scale = 2.0 ** 46
    got = pkg.to_host(be.ct_pt_matrix_mul_wo_pre(pkg.to_device(X), W, scale))
    exp = o.ct_pt_matmul_scalar(X.reshape(-1), W, K, C, limbs, scale)
    assert (got.reshape(-1) == exp).all()


def test_matmul_full_c1_shape_spot_check(pkg, backend_moai, oracle_moai):
    """BASELINE config 1 at full size (768 x 768, 2 limbs): two output columns against the oracle."""
    import torch
    o, be = oracle_moai, backend_moai
    K, C, limbs, scale = 768, 768, 2, 2.0 ** 46
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    X = torch.empty((K, 2, limbs, o.n), dtype=torch.int64, device="cuda")
    for l in range(limbs):
        X[:, :, l, :] = torch.randint(0, int(o.q[l]), (K, 2, o.n), generator=g, device="cuda", dtype=torch.int64)
    W = np.random.default_rng(6).normal(size=(K, C)) * 0.04
    got = be.ct_pt_matrix_mul_wo_pre(X, W, scale)
    Xh = pkg.to_host(X).reshape(-1)
    for col in (0, 767):
        exp = o.ct_pt_matmul_scalar(Xh, W, K, C, limbs, scale, c_begin=col, c_end=col + 1)
        assert (pkg.to_host(got[col]).reshape(-1) == exp).all(), col


def test_matmul_decrypts_to_plain_product(pkg, backend_small, oracle_small):
    """End-to-end meaning: column-packed encrypted X times plaintext W decrypts to X @ W
    (closed-form style of M/test/matrix_mul/test_ct_pt_matrix_mul.hpp:38-42,88-89)."""
    o, be = oracle_small, backend_small
    rng = np.random.default_rng(8)
    K, C, limbs, scale = 12, 5, 3, 2.0 ** 30
    slots = o.n // 2
    sk = o.gen_secret(1, hamming_weight=64)
    Xp = rng.normal(size=(slots, K)) * 0.5          # slot s, feature j
    W = rng.normal(size=(K, C)) * 0.3
    cts = np.stack([o.encrypt_sym(sk, 10 + j, o.encode(Xp[:, j], scale, limbs), limbs) for j in range(K)])
    out = pkg.to_host(be.ct_pt_matrix_mul_wo_pre(pkg.to_device(cts.reshape(K, 2, limbs, o.n)), W, scale))
    out_scale = scale * scale / float(o.q[limbs - 1])
    exp = Xp @ W
    for i in range(C):
        dec = o.decode(o.decrypt(sk, out[i].reshape(-1), 2, limbs - 1), limbs - 1, out_scale).real
        assert np.abs(dec - exp[:, i]).max() < 2e-3, i


def test_matmul_bad_dimensions(pkg, backend_small, oracle_small):
    o, be = oracle_small, backend_small
    X = pkg.to_device(rand_cts(o, np.random.default_rng(0), 3, 2))
    with pytest.raises(pkg.MoaiError):
        be.ct_pt_matrix_mul_wo_pre(X, np.zeros((4, 2)), 2.0 ** 30)   # "bad dimensions of X or W"


def test_matmul_host_entry_matches_device_entry(pkg, backend_moai, oracle_moai):
    """moai_ct_pt_matrix_mul_wo_pre_host (pipelined upload / GEMM / download) returns the same residues
    as the device-pointer call."""
    import torch
    o, be = oracle_moai, backend_moai
    rng = np.random.default_rng(77)
    K, C, limbs, scale = 40, 33, 3, 2.0 ** 46
    X = rand_cts(o, rng, K, limbs)
    W = rng.normal(size=(K, C)) * 0.04
    dev = pkg.to_host(be.ct_pt_matrix_mul_wo_pre(pkg.to_device(X), W, scale))
    hX = torch.from_numpy(X.view(np.int64)).pin_memory()
    host = be.ct_pt_matrix_mul_wo_pre_host(hX, W, scale).numpy().view(np.uint64)
    assert (host == dev).all()
