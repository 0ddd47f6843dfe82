"""CPU checks of the bootstrapping plan (host-only part of csrc/bootstrap.cu): the three
CoeffToSlot matrices compose to the inverse CKKS decoding map (up to the bit-reversal that
SlotToCoeff undoes), the three SlotToCoeff matrices compose to the decoding map, and the
cosine polynomial meets its accuracy target on the approximation domain."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_pkg


def plan(log_n, primes, total_limbs, direction, stage):
    lib = load_pkg().load_library()
    arr = (C.c_uint64 * len(primes))(*primes)
    nd, nch = C.c_int32(), C.c_int32()
    cheb = (C.c_double * 64)()
    rc = lib.moai_bootstrap_plan_debug(C.c_int32(log_n), arr, C.c_int32(len(primes)), C.c_int32(total_limbs),
                                       C.c_int32(direction), C.c_int32(stage), C.byref(nd), None, None, cheb,
                                       C.byref(nch))
    assert rc == 0, lib.moai_last_error()
    n = (1 << log_n) // 2
    offs = (C.c_int32 * nd.value)()
    vals = np.zeros(nd.value * n * 2)
    rc = lib.moai_bootstrap_plan_debug(C.c_int32(log_n), arr, C.c_int32(len(primes)), C.c_int32(total_limbs),
                                       C.c_int32(direction), C.c_int32(stage), C.byref(nd), offs,
                                       vals.ctypes.data_as(C.POINTER(C.c_double)), cheb, C.byref(nch))
    assert rc == 0
    d = vals.reshape(nd.value, n, 2)
    return {int(offs[k]): d[k, :, 0] + 1j * d[k, :, 1] for k in range(nd.value)}, np.array(cheb[: nch.value])


def apply(diags, v):
    n = len(v)
    out = np.zeros(n, dtype=complex)
    for off, dg in diags.items():
        out += dg * np.roll(v, -off)          # out[p] += diag[p] * v[p + off]
    return out


def bitrev_perm(n):
    bits = n.bit_length() - 1
    return np.array([int(format(i, "0%db" % bits)[::-1], 2) for i in range(n)])


@pytest.mark.parametrize("log_n", [7, 9, 12])
def test_linear_stages_compose_to_the_ckks_embedding(log_n):
    N = 1 << log_n
    n = N // 2
    primes = [1099511480321] * 17        # only the count matters for the plan
    zeta = np.exp(2j * np.pi / (2 * N))
    rot = [pow(5, j, 2 * N) for j in range(n)]
    U0 = np.array([[zeta ** ((rot[j] * k) % (2 * N)) for k in range(n)] for j in range(n)]) if n <= 256 else None
    rng = np.random.default_rng(log_n)
    c = rng.normal(size=n) + 1j * rng.normal(size=n)
    if U0 is not None:
        z = U0 @ c
    else:
        # decode through the same index arithmetic on a few columns only is too slow; use the
        # identity SlotToCoeff(CoeffToSlot(z)) instead for large n
        z = c
    K = 25
    P = bitrev_perm(n)
    w = z.copy()
    n_diags = []
    for s in range(3):
        d, _ = plan(log_n, primes, 16, 0, s)
        n_diags.append(len(d))
        w = apply(d, w)
    if U0 is not None:
        assert np.abs(w - c[P] / (2 * K)).max() < 1e-12
    back = w * (2 * K)
    for s in range(3):
        d, _ = plan(log_n, primes, 16, 1, s)
        back = apply(d, back)
    assert np.abs(back - z).max() < 1e-9 * max(1.0, np.abs(z).max())
    assert max(n_diags) <= 2 ** (((log_n - 1) + 2) // 3 + 1) - 1


def test_cosine_polynomial_accuracy():
    _, cheb = plan(7, [1099511480321] * 17, 16, 0, 0)
    assert len(cheb) == 60
    K, r, w = 25, 2, 2.0 ** -10
    xs = np.concatenate([i + w * np.linspace(-1, 1, 41) for i in range(-(K - 1), K)])
    approx = np.polynomial.chebyshev.chebval(xs / K, cheb)
    exact = np.cos(2 * np.pi * (xs - 0.25) / 2 ** r)
    assert np.abs(approx - exact).max() < 1e-8        # fit error 2.6e-10 measured
    assert np.abs(cheb).max() < 2.0                   # O(1) coefficients: benign for CKKS scales


def cosine_fit(K, deg, r, log_width=10):
    lib = load_pkg().load_library()
    cheb = (C.c_double * 128)()
    n, pl = C.c_int32(), C.c_int32()
    rc = lib.moai_bootstrap_cosine_fit_debug(C.c_int32(K), C.c_int32(deg), C.c_int32(r), C.c_int32(log_width), cheb,
                                             C.byref(n), C.byref(pl))
    if rc:
        raise ValueError(lib.moai_last_error().decode())
    return np.array(cheb[: n.value]), pl.value


def evalmod_error(K, deg, r, log_width=10):
    """max error of the fitted cosine on the approximation domain, and of sin(2 pi x) after the r double angles"""
    c, pl = cosine_fit(K, deg, r, log_width)
    w = 2.0 ** -log_width
    xs = np.concatenate([i + np.linspace(-w, w, 41) for i in range(-(K - 1), K)])
    v = np.polynomial.chebyshev.chebval(xs / K, c)
    e_poly = np.abs(v - np.cos(2 * np.pi * (xs - 0.25) / 2 ** r)).max()
    for _ in range(r):
        v = 2 * v * v - 1
    return pl, e_poly, np.abs(v - np.sin(2 * np.pi * xs)).max()


def test_evalmod_level_budget_and_fit_error():
    """The polynomial's level count follows from its degree (59 -> 6, 31 -> 5) and EvalMod must spend 8 levels in all.
    The reference's (deg 59, 2 double angles) is accurate to 1e-8 after the double angles; (31, 3), which would save
    two relinearizations, fits the budget but leaves ~3e-4 (25 periods do not fit a degree-31 polynomial) — the
    reason it is not used."""
    pl, e_poly, e_sin = evalmod_error(25, 59, 2)
    assert pl == 6 and e_poly < 1e-9 and e_sin < 1e-8
    pl, e_poly, e_sin = evalmod_error(25, 31, 3)
    assert pl == 5 and 1e-5 < e_sin < 1e-3
    for deg, r in ((31, 2), (59, 3), (47, 1)):
        with pytest.raises(ValueError, match="8 levels"):
            cosine_fit(25, deg, r)
