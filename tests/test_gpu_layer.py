"""GPU checks of the attention head, softmax with bootstrapping, and one full encoder layer
(SURVEY §8(a) B6/B7 and BASELINE config 4) at a reduced ring degree (N = 4096, 16 inputs x 128
tokens) with the repo's exact 36-prime chain shape.  These stages contain bootstrapping, so parity
is by TOLERANCE against a float64 model that applies the same approximations the reference uses
(exp = (1 + x/128)^128, Goldschmidt inverse, Newton/Goldschmidt inverse square root, degree-24
GELU polynomial).  Stated tolerances: attention head 5e-3 max-abs (measured 6.5e-4), encoder layer 0.15 max-abs on
LayerNorm-normalised outputs (O(1) values; measured 0.124 = the reference's own approximation error on these inputs).
The two fast key modes — "fast" (pre-permuted SEAL digits) and "grouped" (what bench.py times) — are additionally
compared with the SEAL-exact run on the same inputs and secret key: 1e-5 (attention) and 1e-4 (whole layer, four
bootstrapping rounds; measured 6e-6)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BITS = [51] + [46] * 20 + [51] * 14 + [58]
SCALE = 2.0 ** 46
NUM_BATCH = 16          # slots / 128 at N = 4096
TOK = 5                 # valid tokens per input (the reference run uses 5)


RESULTS = {}            # decrypted outputs per (test, key mode): the fast modes are compared with the SEAL-exact run


@pytest.fixture(scope="module", params=["exact", "fast", "grouped"])
def env(pkg, request):
    """exact: SEAL-layout keys (the bootstrapping steps + the driver's powers of two, NAF fallback for
    the rest) -> every key switch is SEAL's.  fast: hoisted rotations with pre-permuted keys for the
    exact steps the pipeline takes, truncated to the level they are used at.  grouped: what
    bench.py times — grouped-digit keys, lazy mod-down, single-digit first CoeffToSlot stage (hoisting mode 2), GELU by
    baby-step / giant-step evaluation."""
    from oracle import Oracle
    o = Oracle(12, BITS)
    be = pkg.Backend(12, o.q)
    boot = pkg.Bootstrapper(be, total_limbs=35)
    sk = o.gen_secret(7, hamming_weight=64)
    relin = pkg.to_device(o.gen_relin_key(sk, 5))

    def gen(i, st):
        e = o.elt_from_step(st)
        return e, pkg.to_device(o.gen_galois_key(sk, 3000 + i, e).reshape(o.kl - 1, 2, o.kl, o.n))

    if request.param == "exact":
        steps = set(boot.required_steps())
        for k in range(11):
            steps |= {1 << k, (o.n // 2) - (1 << k)}
        gal = dict(gen(i, st) for i, st in enumerate(sorted(steps) + [0]))
        keys = be.make_keys(relin=relin, galois=gal)
    elif request.param == "grouped":
        boot.set_hoisting(2)
        fast, grouped, single = {}, {}, {}

        def add(i, st, level):
            e, k = gen(i, st)
            if level == 0:
                single[e] = be.key_prepare_single(k, e)
                return
            g = be.key_prepare_grouped(k, e, level)
            if g is None:
                fast.setdefault(e, []).append(be.key_prepare(k, e, max_limbs=level))
            else:
                grouped.setdefault(e, []).append(g)

        i = 0
        for st, lvs in sorted(boot.required_step_levels().items()):
            for lv in lvs:
                add(i, st, lv)
                i += 1
        att = pkg.attention_rotation_steps(NUM_BATCH)
        for tag, level in (("qk", 14), ("sv", 3)):
            for j, st in enumerate(att[tag]):
                add(500 + 100 * level + j, st, level)
        relin4 = relin.reshape(o.kl - 1, 2, o.kl, o.n)
        grouped[0] = [be.key_prepare_grouped(relin4, 0, lv, k_extra=k, pre_permute=False)
                      for k, lv in sorted(be.ksg_plan(range(1, o.kl - 1)).items())]
        keys = be.make_keys(relin=relin, galois_fast=fast, grouped=grouped, single=single)
    else:
        boot.set_hoisting(True)
        fast = {}
        for i, st in enumerate(boot.required_steps() + [0]):
            e, k = gen(i, st)
            fast.setdefault(e, []).append(be.key_prepare(k, e))
        att = pkg.attention_rotation_steps(NUM_BATCH)
        for tag, level in (("qk", 14), ("sv", 3)):
            for i, st in enumerate(att[tag]):
                e, k = gen(500 + 100 * level + i, st)
                fast.setdefault(e, []).append(be.key_prepare(k, e, max_limbs=level))
        keys = be.make_keys(relin=relin, galois_fast=fast)
    mask = np.zeros(o.n // 2, dtype=np.int32)
    for k in range(TOK):
        mask[k * NUM_BATCH:(k + 1) * NUM_BATCH] = 1
    return o, be, boot, sk, keys, mask, request.param


def pack_encrypt(o, sk, X, limbs, seed):
    """X: [tokens(128), inputs(16), cols] -> cols ciphertexts, slot 16*k + b = X[k, b, col]
    (batch_input, M/source/matrix_mul/Batch_encode_encrypt.hpp:21-27)."""
    cols = X.shape[2]
    out = np.empty((cols, 2, limbs, o.n), dtype=np.uint64)
    for c in range(cols):
        v = X[:, :, c].reshape(-1)
        out[c] = o.encrypt_sym(sk, seed + c, o.encode(v, SCALE, limbs), limbs).reshape(2, limbs, o.n)
    return out


def decrypt_cols(o, sk, pkg, ct, scale):
    ct = pkg.to_host(ct)
    limbs = ct.shape[2]
    return np.stack([o.decode(o.decrypt(sk, ct[c].reshape(-1), 2, limbs), limbs, scale).real.reshape(128, NUM_BATCH)
                     for c in range(ct.shape[0])], axis=2)       # [token, input, col]


def exp128(x):
    return (1 + x / 128.0) ** 128


def goldschmidt_inverse(x, iters):
    y = 1 - x
    res = 1 + y
    for _ in range(iters):
        y = y * y
        res = res * (1 + y)
    return res


def attention_model(X, WQ, WK, WV, bQ, bK, bV, c_shift, iters=16):
    """float64 model of single_att_block on the TOK valid tokens of every input."""
    Xv = X[:TOK]                                         # [tok, inp, hidden]
    Q = np.einsum("tbh,hc->tbc", Xv, WQ) + bQ
    K = np.einsum("tbh,hc->tbc", Xv, WK) + bK
    V = np.einsum("tbh,hc->tbc", Xv, WV) + bV
    S = np.einsum("tbc,ubc->btu", Q, K)                  # [inp, tok, tok]
    E = exp128(S - c_shift)
    denom = E.sum(axis=2, keepdims=True) + 1e-5
    P = E * goldschmidt_inverse(denom, iters)
    return np.einsum("btu,ubc->tbc", P, V)               # [tok, inp, col]


def _versus_exact(name, mode, got, tol):
    """Fast modes against the SEAL-exact run of the same test (same inputs, same secret key): the only differences are
    key-switching noise and, for GELU, the evaluation order."""
    RESULTS[(name, mode)] = got
    if mode != "exact" and (name, "exact") in RESULTS:
        d = np.abs(got - RESULTS[(name, "exact")]).max()
        print("%s: %s vs exact keys, max-abs difference of the decrypted outputs %.3g (tolerance %.0e)" % (name, mode, d, tol))
        assert d < tol, d


def test_single_att_block_matches_float_model(pkg, env):
    o, be, boot, sk, keys, mask, mode = env
    rng = np.random.default_rng(1)
    hidden, col_W = 48, 8                                 # reduced widths; the pipeline is width-agnostic
    X = np.zeros((128, NUM_BATCH, hidden))
    X[:TOK] = rng.normal(size=(TOK, NUM_BATCH, hidden)) * 0.5
    # scores Q.K ~ 6 +- 0.4 so that sum_u exp(S - 7.5) stays inside (0, 2), the convergence domain of the
    # reference's Goldschmidt inverse (softmax.hpp:49-82; the shift 7.5 is its layer-0 constant, :324)
    WQ, WK = (rng.normal(size=(hidden, col_W)) * 0.03 for _ in range(2))
    WV = rng.normal(size=(hidden, col_W)) * 0.12
    bQ = np.full(col_W, np.sqrt(6.0 / col_W)) + rng.normal(size=col_W) * 0.02
    bK = np.full(col_W, np.sqrt(6.0 / col_W)) + rng.normal(size=col_W) * 0.02
    bV = rng.normal(size=col_W) * 0.1
    limbs = 15                                            # chain_index 14, like the driver (test_full_scheme.hpp:496-507)
    cts = pack_encrypt(o, sk, X, limbs, 100)
    layer_id = 0                                          # shift constant 7.5 (softmax.hpp:324)
    out, out_scale = boot.single_att_block(keys, pkg.to_device(cts), SCALE, WQ, WK, WV, bQ, bK, bV, mask, TOK, NUM_BATCH,
                                           iters=16, layer_id=layer_id)
    assert out.shape[0] == col_W and out.shape[2] == 2 and out_scale == SCALE
    got = decrypt_cols(o, sk, pkg, out, out_scale)
    exp = attention_model(X, WQ, WK, WV, bQ, bK, bV, 7.5)
    err = np.abs(got[:TOK] - exp).max()
    print("attention head (%s keys): max-abs vs the float64 model %.3g" % (mode, err))
    assert err < 5e-3, err                                # measured 6.5e-4 in all three key modes
    assert np.abs(got[TOK:]).max() < 5e-3                 # padding tokens stay (approximately) zero
    _versus_exact("att", mode, got, 1e-5)                 # measured 9.4e-7 (fast), 6.3e-7 (grouped)


def layer_model(X, w, iters=16):
    """float64 model of one encoder layer on the valid tokens, using exact LayerNorm / GELU (the
    reference's polynomial approximations of those are within the stated tolerance)."""
    import math
    hidden, heads, hd = w["hidden"], w["heads"], w["head_dim"]
    Xv = X[:TOK]
    att = []
    for h in range(heads):
        WQ = w["WQ"].reshape(heads, hidden, hd)[h]
        WK = w["WK"].reshape(heads, hidden, hd)[h]
        WV = w["WV"].reshape(heads, hidden, hd)[h]
        att.append(attention_model(X, WQ, WK, WV, w["bQ"].reshape(heads, hd)[h], w["bK"].reshape(heads, hd)[h],
                                   w["bV"].reshape(heads, hd)[h], 7.5, iters))
    A = np.concatenate(att, axis=2)
    so = A @ w["selfoutput"].reshape(hidden, hidden) + w["selfoutput_bias"]

    def ln(v, g, b):
        mu = v.mean(axis=2, keepdims=True)
        var = ((v - mu) ** 2).mean(axis=2, keepdims=True)
        return (v - mu) / np.sqrt(var) * g + b

    h1 = ln(so + Xv, w["ln1_gamma"], w["ln1_beta"])
    inter = h1 @ w["inter_weight"].reshape(hidden, -1) + w["inter_bias"]
    gelu = 0.5 * inter * (1 + np.vectorize(math.erf)(inter / math.sqrt(2)))
    fin = gelu @ w["final_weight"].reshape(-1, hidden) + w["final_bias"]
    return ln(fin + h1, w["ln2_gamma"], w["ln2_beta"])


def test_encoder_layer_end_to_end(pkg, env):
    """One full encoder layer (4 x 768 bootstrappings, 12 heads) at N = 4096, hidden = 768."""
    o, be, boot, sk, keys, mask, mode = env
    rng = np.random.default_rng(2)
    hidden, heads, hd, inter = 768, 12, 64, 3072
    X = np.zeros((128, NUM_BATCH, hidden))
    X[:TOK] = rng.normal(size=(TOK, NUM_BATCH, hidden)) * 0.5
    w = {"hidden": hidden, "heads": heads, "head_dim": hd, "inter": inter,
         # scores ~ 5.5 +- 0.2 (see the attention test); LayerNorm2's inverse-sqrt needs a variance of
         # tens to hundreds (layernorm.hpp:18-24 initial guess), hence the larger final weights
         "WQ": rng.normal(size=(heads, hidden, hd)) * 0.004, "WK": rng.normal(size=(heads, hidden, hd)) * 0.004,
         "WV": rng.normal(size=(heads, hidden, hd)) * 0.03,
         "bQ": np.sqrt(5.5 / hd) + rng.normal(size=(heads, hd)) * 0.005,
         "bK": np.sqrt(5.5 / hd) + rng.normal(size=(heads, hd)) * 0.005,
         "bV": rng.normal(size=(heads, hd)) * 0.05,
         "selfoutput": rng.normal(size=(hidden, hidden)) * 0.03, "selfoutput_bias": rng.normal(size=hidden) * 0.05,
         "ln1_gamma": 1 + rng.normal(size=hidden) * 0.05, "ln1_beta": rng.normal(size=hidden) * 0.05,
         "inter_weight": rng.normal(size=(hidden, inter)) * 0.03, "inter_bias": rng.normal(size=inter) * 0.05,
         "final_weight": rng.normal(size=(inter, hidden)) * 0.2, "final_bias": rng.normal(size=hidden) * 0.05,
         "ln2_gamma": 1 + rng.normal(size=hidden) * 0.05, "ln2_beta": rng.normal(size=hidden) * 0.05}
    limbs = 21
    cts = pack_encrypt(o, sk, X, limbs, 500)
    be.profile(True)
    out, out_scale = boot.encoder_layer(keys, pkg.to_device(cts), SCALE, w, mask, TOK, NUM_BATCH, layer_id=0,
                                        boot_chunk=64)
    stages = be.profile_dump()
    be.profile(False)
    print("per-stage ms:", {k: round(v[0], 1) for k, v in stages.items()})
    assert out.shape == (hidden, 2, 21, o.n) and out_scale == SCALE
    got = decrypt_cols(o, sk, pkg, out, out_scale)
    exp = layer_model(X, w)
    err = np.abs(got[:TOK] - exp).max()
    print("encoder layer (%s keys): max-abs vs the float64 model %.3g on O(1) outputs" % (mode, err))
    assert np.isfinite(got).all()
    # 0.124 in all three key modes: the error of the reference's OWN approximations (degree-24 GELU, Newton inverse square
    # root of LayerNorm, Goldschmidt softmax) against exact float64 functions on these inputs, not of the arithmetic —
    # the arithmetic is what the comparison with the SEAL-exact run below isolates
    assert err < 0.15, err
    _versus_exact("layer", mode, got, 1e-4)               # measured 6.2e-6 (fast), 5.9e-6 (grouped)
