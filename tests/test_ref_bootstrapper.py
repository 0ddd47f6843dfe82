"""CPU checks of the bootstrapping checker itself (no GPU): oracle/_ref now contains the reference's Bootstrapper,
ModularReducer, Polynomial, Remez, func (M/source/bootstrapping/**) compiled unmodified behind an NTL::RR-over-MPFR
shim (oracle/refbuild/ntl_shim).  These tests pin that checker before the GPU tests rely on it
(tests/test_gpu_boot_reference.py) and pin the product's EvalMod polynomial against the reference's Remez output."""
import json
import os

import numpy as np
import pytest

from test_bootstrap_plan import cosine_fit

HERE = os.path.dirname(os.path.abspath(__file__))
K, DEG, R, LOGW = 25, 59, 2, 10


def golden_poly():
    g = json.load(open(os.path.join(HERE, "golden", "evalmod_remez_K25_w10_d59_r2.json")))
    sic = float.fromhex(g["scale_inverse_coeff"])
    cheb = np.array([float.fromhex(c) for c in g["cheb_times_sic"]])
    return cheb, sic


def domain(points=201):
    w = 2.0 ** -LOGW
    return np.concatenate([i + np.linspace(-w, w, points) for i in range(-(K - 1), K)])


def test_product_cosine_matches_reference_remez():
    """csrc/bootstrap.cu fits the cosine by least squares; the reference runs a multi-interval Remez
    (common/Remez.cpp:557-586).  On the approximation domain the two polynomials agree to 5e-10 and both are within
    3e-10 of cos(2 pi (x - 1/4) / 4) (reference minimax error 1.94e-10, ours 2.62e-10)."""
    ref, sic = golden_poly()
    ours, _ = cosine_fit(K, DEG, R, LOGW)
    xs = domain()
    exact = np.cos(2 * np.pi * (xs - 0.25) / 2 ** R)
    v_ref = np.polynomial.chebyshev.chebval(xs / K, ref / sic)
    v_our = np.polynomial.chebyshev.chebval(xs / K, ours)
    assert np.abs(v_ref - exact).max() < 2.5e-10
    assert np.abs(v_our - exact).max() < 3e-10
    assert np.abs(v_our - v_ref).max() < 5e-10
    # scale_inverse_coeff^(2^r) is the reference's linear arcsin coefficient ~ 1 / (2 pi) (ModularReducer.cpp:42-47)
    assert abs(sic ** (2 ** R) - 1 / (2 * np.pi)) < 1e-6


@pytest.fixture(scope="module")
def refboot():
    from oracle import SealRef, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    r = SealRef(13, [51] + [46] * 2 + [51] * 14 + [58], hamming_weight=64, seed=5)
    r.make_relin_key()
    steps = r.boot_create()                                   # the reference's own Remez runs here
    r.make_galois_keys(steps, conjugate=True)
    r.boot_prepare()
    return r, steps


def test_reference_remez_reproduces_the_golden_polynomial(refboot):
    """The committed golden coefficients are what the reference's code generates (regenerated live)."""
    r, _ = refboot
    cheb, sic = r.boot_polynomial()
    gold, gsic = golden_poly()
    assert sic == gsic and np.array_equal(cheb, gold)


def test_reference_bootstrap_3_runs_and_is_accurate(refboot):
    """bootstrap_3 of the reference (Bootstrapper.cpp:3496-3502) at N = 8192 (logn = 12, one of the two sizes its
    sfl_full_3 supports): output at chain_index total - 14, scale 2^46, message preserved to 2e-5; ModRaise lands on
    all data limbs with scale q0; the key list is the driver's (test_full_scheme.hpp:436-443)."""
    r, steps = refboot
    assert steps[:13] == [0] + [1 << i for i in range(12)] and len(steps) == len(set(steps))
    rng = np.random.default_rng(1)
    scale = 2.0 ** 46
    z = (rng.normal(size=r.n // 2) + 1j * rng.normal(size=r.n // 2)) * 0.1
    ct = r.encrypt(r.encode(z, scale, 1), 1, scale)
    raised, limbs, sc = r.boot_phase(0, ct, 1, scale)
    assert limbs == r.kl - 1 and sc == float(r.q[0])
    out, limbs, sc = r.bootstrap_3(ct, scale)
    assert limbs == r.kl - 1 - 14 and sc == scale
    dec = r.decode(r.decrypt(out, 2, limbs, sc), limbs, sc)
    assert np.abs(dec - z).max() < 2e-5


def test_oracle_modraise_restatement_matches_reference(refboot):
    """oracle/ckks_oracle.c's ModRaise restatement == Bootstrapper::modraise_inplace (Bootstrapper.cpp:2938-2992),
    bit for bit — the restatement the GPU kernel was pinned against in round 1 is now itself pinned."""
    from oracle import Oracle
    r, _ = refboot
    o = Oracle(13, primes=[int(q) for q in r.q])
    rng = np.random.default_rng(2)
    ct = np.stack([rng.integers(0, int(r.q[0]), r.n, dtype=np.uint64) for _ in range(2)]).reshape(-1)
    exp, limbs, _ = r.boot_phase(0, ct, 1, 2.0 ** 46)
    got = o.modraise(ct, 2, limbs)
    assert np.array_equal(got, exp)


def test_golden_layer0_activations_are_consistent():
    """tests/golden/layer0_activations.npz (from /root/reference/data/layer_0): the CSV pairs are what their names
    say — softmax(QKT) = aftsoftmax per head, LayerNorm(in) = out, GELU(in) = out — so they can gate the GPU stages."""
    import math
    g = np.load(os.path.join(HERE, "golden", "layer0_activations.npz"))
    for h in range(12):
        s = g["QKT"][:, 5 * h:5 * h + 5]
        e = np.exp(s - s.max(axis=1, keepdims=True))
        assert np.abs(e / e.sum(axis=1, keepdims=True) - g["aftsoftmax"][:, 5 * h:5 * h + 5]).max() < 1e-6
        assert s.max() <= 7.5                                  # the layer-0 shift constant of softmax.hpp:324
    for name in ("ln1", "ln2"):
        x = g[name + "_in"]
        mu = x.mean(axis=1, keepdims=True)
        var = ((x - mu) ** 2).mean(axis=1, keepdims=True)
        y = (x - mu) / np.sqrt(var + 1e-12) * g[name + "_gamma"] + g[name + "_beta"]
        assert np.abs(y - g[name + "_out"]).max() < 1e-5
    gi = g["gelu_in"]
    ge = 0.5 * gi * (1 + np.vectorize(math.erf)(gi / math.sqrt(2)))
    assert np.abs(ge - g["gelu_out"]).max() < 1e-5
    assert g["selfoutput_linear"].shape == (5, 768)


def test_softmax_scale_drift_model_matches_reference_decrypted_golden():
    """tests/golden/layer0_reference_decrypted.npz holds what the reference's softmax_boot decrypts to for layer 0's
    scores (oracle/_ref, N = 8192).  The float64 model used by the full-size GPU gate — (1 + x/128)^128, Goldschmidt
    inverse, and the forced scale resets turned into the factors of softmax_scale_drift() for THAT ring's primes —
    reproduces it to 2e-4 (without the drift factors the gap is 1.1e-3), which is what licenses using the same model
    with the N = 65536 primes on the GPU."""
    from oracle import Oracle, MOAI_BITS
    from test_gpu_fullsize import softmax_scale_drift
    g = np.load(os.path.join(HERE, "golden", "layer0_activations.npz"))
    ref = np.load(os.path.join(HERE, "golden", "layer0_reference_decrypted.npz"))
    q = [float(x) for x in Oracle(13, MOAI_BITS).q]
    f_exp, f_inv, f_out = softmax_scale_drift(q)
    assert 1e-3 < f_inv - 1 < 1.3e-3
    worst, worst_plain = 0.0, 0.0
    for h in range(12):
        S = g["QKT"][:, 5 * h:5 * h + 5]
        E = (1 + (S - 7.5) / 128.0) ** 128
        model = E * f_exp / ((E * f_exp).sum(axis=1, keepdims=True) + 1e-5) * f_inv * f_out
        plain = E / (E.sum(axis=1, keepdims=True) + 1e-5)
        got = ref["softmax_ref"][:, 5 * h:5 * h + 5]
        worst = max(worst, np.abs(got - model).max())
        worst_plain = max(worst_plain, np.abs(got - plain).max())
    assert worst < 2e-4 and 5e-4 < worst_plain < 2e-3
    # and the other reference-decrypted fixtures sit where the test docstrings say they do
    assert np.abs(ref["ln1_ref"] - g["ln1_out"]).max() < 5e-4
    assert 0.1 < np.abs(ref["ln2_ref"] - g["ln2_out"]).max() < 0.15
    assert np.abs(ref["gelu_ref"] - g["gelu_out"]).max() < 0.09
    assert 1.0 < np.abs(ref["gelu_ref65536"] - g["gelu_out"][:, ref["gelu_cols65536"]]).max() < 1.2
