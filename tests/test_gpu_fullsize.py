"""Full-size correctness gate: the repo's own parameters (N = 65536, 35 + 1 primes, Hamming-weight-192 secret, K = 25:
M/test/test_full_scheme.hpp:345-368, 413-448) with VALID keys, decrypted results.

* Bootstrapping in the mode every headline timing uses (fast: hoisted rotations,
  pre-permuted keys; plain and two-real-ciphertexts-per-bootstrapping), with the decrypted error reported PER PHASE
  (ModRaise / CoeffToSlot / EvalMod / SlotToCoeff).
* Golden gate (north_star: "decrypted layer outputs must match the reference's"): the non-linear stages of encoder
  layer 0 on the reference's own activations (tests/golden/layer0_activations.npz, made from
  /root/reference/data/layer_0/**/allresults/*.csv by tests/golden/make_layer0_golden.py): softmax_boot vs aftsoftmax.csv,
  layernorm vs real_self_output.csv, gelu_v2 vs real_intermediate_output.csv, layernorm2 vs real_final_output.csv, and
  config C1 on data/selfoutput_linear.txt with the reference run's {5, 0, ...} token mask.  The reference's modules are
  polynomial approximations; its OWN error against these CSVs (oracle/_ref at N = 4096, measured in this repo:
  layernorm 4.4e-4, layernorm2 0.13 max-abs / 5.7 % rel, gelu_v2 0.029) bounds what can be asked of a faithful
  implementation, and each tolerance below is stated next to that figure.

Keys are generated on the device (tests/devkeygen.py): genuine RLWE keys in SEAL's layout with torch's randomness.
~65 GiB of HBM; the whole module takes a few minutes on one B200."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOG_N = 16
N = 1 << LOG_N
SLOTS = N // 2
NUM_BATCH = 256
TOK = 5
SCALE = 2.0 ** 46
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layer0_activations.npz")
# what the reference's own encrypted modules output for those activations (decrypted; oracle/_ref, real SEAL):
# tests/golden/make_layer0_reference_outputs.py (N = 8192) and make_gelu_reference_fullsize.py (N = 65536)
GOLDEN_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layer0_reference_decrypted.npz")


@pytest.fixture(scope="module", params=["grouped", "fast"])
def env(pkg, request):
    """Both fast-mode key layouts: "grouped" (grouped-digit keys derived from the stock keys, lazy mod-down in the linear
    stages and single-digit keys on the first CoeffToSlot stage: csrc/ksgroup.hpp — what bench.py times) and "fast"
    (SEAL's per-prime digits, pre-permuted, one mod-down per rotation)."""
    import torch
    from devkeygen import DeviceKeyGen
    from oracle import Oracle, MOAI_BITS
    o = Oracle(LOG_N, MOAI_BITS)
    be = pkg.Backend(LOG_N, o.q)
    kg = DeviceKeyGen(pkg, be, hamming_weight=192, seed=20250991)
    boot = pkg.Bootstrapper(be, total_limbs=35)
    boot.set_hoisting(2 if request.param == "grouped" else 1)
    fast, grouped, single = {}, {}, {}
    if request.param == "grouped":
        for st, lvs in sorted(boot.required_step_levels().items()):
            e = be.galois_elt_from_step(st)
            k = kg.galois_key(e)
            for lv in lvs:
                if lv == 0:         # baby step of the first CoeffToSlot stage: single-digit key
                    single[e] = be.key_prepare_single(k, e)
                    continue
                gk = be.key_prepare_grouped(k, e, lv)
                if gk is None:      # no spare prime at the top level: SEAL's digits
                    fast.setdefault(e, []).append(be.key_prepare(k, e, max_limbs=lv))
                else:
                    grouped.setdefault(e, []).append(gk)
            del k
        relin = kg.relin_key()
        relin4 = relin.reshape(be.kl - 1, 2, be.kl, N)
        grouped[0] = [be.key_prepare_grouped(relin4, 0, lv, k_extra=k, pre_permute=False)
                      for k, lv in sorted(be.ksg_plan(range(1, be.kl - 1)).items())]
        keys = be.make_keys(relin=relin, galois_fast=fast, grouped=grouped, single=single)
    else:
        for st in boot.required_steps() + [0]:
            e = be.galois_elt_from_step(st)
            k = kg.galois_key(e)
            fast.setdefault(e, []).append(be.key_prepare(k, e))
            del k
        keys = be.make_keys(relin=kg.relin_key(), galois_fast=fast)
    torch.cuda.synchronize()
    print("\n[%s keys: %.1f GiB on the device]" % (request.param, torch.cuda.memory_allocated() / 2 ** 30))
    mask = np.zeros(SLOTS, dtype=np.int32)
    for k in range(TOK):
        mask[k * NUM_BATCH] = 1                           # bias_vec(input_len = {5, 0, ...}), Batch_encode_encrypt.hpp:39-49
    yield {"o": o, "be": be, "kg": kg, "boot": boot, "keys": keys, "mask": mask, "g": np.load(GOLDEN),
           "ref": np.load(GOLDEN_REF)}
    del keys, fast, grouped, single
    be.close()
    torch.cuda.empty_cache()


@pytest.fixture(autouse=True)
def _release_between_tests(env):
    """torch's caching allocator and the library's arena both keep what they freed; with 65 GiB of keys resident the
    GiB-sized batches of the next test need that memory back."""
    import torch
    yield
    torch.cuda.synchronize()
    env["be"].lib.moai_release_cached_memory(env["be"].h)
    torch.cuda.empty_cache()


def pack_rows(A):
    """A [TOK, cols] activations of input 0 -> [cols, SLOTS]: slot 256 k holds token k (batch_input layout,
    Batch_encode_encrypt.hpp:21-27); the other 255 inputs are empty, as in the reference run."""
    v = np.zeros((A.shape[1], SLOTS))
    for k in range(TOK):
        v[:, k * NUM_BATCH] = A[k]
    return v


def encrypt_cols(e, V, limbs, chunk=256):
    import torch
    outs = []
    for c0 in range(0, V.shape[0], chunk):
        pt = e["be"].encode(V[c0:c0 + chunk], SCALE, limbs)
        outs.append(e["kg"].encrypt(pt))
    return torch.cat(outs)


def slot_values(e, ct, scale, slots):
    """Decrypt on the device and evaluate the plaintext polynomial at the requested slots only:
    slot j <-> m(zeta^(5^j)), zeta = exp(i pi / N) (S/ckks.cpp:36-52)."""
    import torch
    be, kg = e["be"], e["kg"]
    B = ct.shape[0]
    m = kg.decrypt(ct[:, :, :1].contiguous()).view(B, N).contiguous()
    be.ntt_inverse_limb_(m, 0)
    q0 = int(be.primes[0])
    c = torch.where(m > q0 // 2, m - q0, m).to(torch.float64)
    idx = torch.arange(N, device=c.device, dtype=torch.int64)
    pos = torch.tensor([pow(5, int(j), 2 * N) for j in slots], device=c.device, dtype=torch.int64)
    ang = ((idx[:, None] * pos[None, :]) % (2 * N)).to(torch.float64) * (np.pi / N)
    re = c @ torch.cos(ang)
    im = c @ torch.sin(ang)
    return ((re + 1j * im) / scale).cpu().numpy()


VALID = [k * NUM_BATCH for k in range(TOK)]


def test_slot_evaluation_matches_decode(pkg, env):
    """Self-check of the shortcut above against the oracle's full decode."""
    rng = np.random.default_rng(0)
    z = rng.normal(size=(2, SLOTS)) * 0.3
    ct = encrypt_cols(env, z, 2)
    full = env["kg"].decrypt_decode(ct, SCALE, env["o"])
    some = slot_values(env, ct, SCALE, [0, 1, 256, 1024, SLOTS - 1])
    assert np.abs(full[:, [0, 1, 256, 1024, SLOTS - 1]] - some).max() < 1e-9
    assert np.abs(full.real - z).max() < 1e-8             # fresh-encryption noise at scale 2^46


def _bitrev_perm(n):
    bits = n.bit_length() - 1
    idx = np.arange(n)
    out = np.zeros(n, dtype=np.int64)
    for b in range(bits):
        out |= ((idx >> b) & 1) << (bits - 1 - b)
    return out


def test_bootstrap_fullsize_per_phase(pkg, env):
    """Plain full-slot bootstrapping of complex messages (C3's inputs: N(0, 0.3^2) per slot) at N = 65536 / HW-192,
    error after every phase.  Tolerances: |I| < K = 25; CoeffToSlot 1e-6 on t / (K q0); EvalMod 1e-6 on
    sin(2 pi t / q0); final message 1e-4 max-abs (the reference's bootstrap_3 measured 3e-6 ... 9e-6 at N = 8192)."""
    import torch
    o, be, kg, boot, keys = env["o"], env["be"], env["kg"], env["boot"], env["keys"]
    rng = np.random.default_rng(3)
    B = 2
    z = (rng.normal(size=(B, SLOTS)) + 1j * rng.normal(size=(B, SLOTS))) * 0.3
    ct = kg.encrypt(be.encode(z, SCALE, 1))
    q0, q1 = int(be.primes[0]), int(be.primes[1])
    K = 25

    # phase 1: ModRaise.  t = m + q0 I over the integers; recover (m0, I) from the residues mod q0 and q1
    r1, s1 = boot.bootstrap_phase_debug(keys, ct, SCALE, 1)
    assert r1.shape == (B, 2, 35, N) and s1 == float(q0)
    m2 = kg.decrypt(r1[:, :, :2].contiguous())
    be.ntt_inverse_(m2.view(B, 1, 2, N))
    m2 = m2.cpu().numpy()
    m0 = np.where(m2[:, 0] > q0 // 2, m2[:, 0] - q0, m2[:, 0])
    I = np.full(m0.shape, 1000, dtype=np.int64)
    for cand in range(-40, 41):
        hit = (m0 + cand * q0) % q1 == m2[:, 1]
        I[hit] = cand
    assert (I != 1000).all(), "ModRaise output is not m + q0 I with |I| <= 40"
    print("ModRaise: max |I| = %d (K = %d), |m0| / q0 max = %.3g" % (np.abs(I).max(), K, np.abs(m0).max() / q0))
    assert np.abs(I).max() < K
    # the input's own plaintext (decrypt at one limb) must be m0 exactly: ModRaise changes no coefficient mod q0
    m_in = kg.decrypt(ct).view(B, N).contiguous()
    be.ntt_inverse_limb_(m_in, 0)
    m_in = m_in.cpu().numpy()
    assert (np.where(m_in > q0 // 2, m_in - q0, m_in) == m0).all()
    t_over_q0 = I + m0 / q0                                # [B, N]

    # phase 2: CoeffToSlot: slots of the real / imaginary halves = t[bitrev(j)] / (K q0), t[bitrev(j) + N/2] / (K q0)
    P = _bitrev_perm(SLOTS)
    r2, s2 = boot.bootstrap_phase_debug(keys, ct, SCALE, 2)
    assert r2.shape[0] == 2 * B
    d2 = kg.decrypt_decode(r2, s2, o)
    e_cts = max(np.abs(d2[b] - t_over_q0[b][P] / K).max() for b in range(B))
    e_cts = max(e_cts, max(np.abs(d2[B + b] - t_over_q0[b][P + SLOTS] / K).max() for b in range(B)))
    print("CoeffToSlot: max |slot - t/(K q0)| = %.3g at %d limbs" % (e_cts, r2.shape[2]))
    assert e_cts < 1e-6

    # phase 3: EvalMod: slots ~ sin(2 pi t / q0)
    r3, s3 = boot.bootstrap_phase_debug(keys, ct, SCALE, 3)
    d3 = kg.decrypt_decode(r3, s3, o)
    e_mod = max(np.abs(d3[b] - np.sin(2 * np.pi * (m0[b] / q0))[P]).max() for b in range(B))
    e_mod = max(e_mod, max(np.abs(d3[B + b] - np.sin(2 * np.pi * (m0[b] / q0))[P + SLOTS]).max() for b in range(B)))
    print("EvalMod: max |slot - sin(2 pi t / q0)| = %.3g at %d limbs" % (e_mod, r3.shape[2]))
    assert e_mod < 1e-6

    # phase 4: the whole bootstrapping
    out, osc = boot.bootstrap_3(keys, ct, SCALE)
    assert out.shape == (B, 2, 21, N) and osc == SCALE
    d = kg.decrypt_decode(out, osc, o)
    e_fin = np.abs(d - z).max()
    print("bootstrap_3 (N = 65536, HW-192): max |out - msg| = %.3g" % e_fin)
    assert e_fin < 1e-4


def test_bootstrap_real_pairs_fullsize(pkg, env):
    """moai_bootstrap_real — the call the encoder layer makes 4 x per layer — on real messages shaped like the layer's
    activations: N(0, 0.3^2) full-slot rows and the LayerNorm-1 golden outputs of the 5-token sentence (|x| up to 63
    in 5 of 32768 slots).  Odd batch: one ciphertext travels alone.  Tolerance 1e-4 max-abs on O(1) values."""
    o, be, kg, boot, keys, g = env["o"], env["be"], env["kg"], env["boot"], env["keys"], env["g"]
    rng = np.random.default_rng(4)
    dense = rng.normal(size=(3, SLOTS)) * 0.3
    sparse = pack_rows(g["ln1_out"][:, :4])
    msgs = np.concatenate([dense, sparse])                 # 7 ciphertexts
    ct = kg.encrypt(be.encode(msgs, SCALE, 1))
    out, osc = boot.bootstrap_real(keys, ct, SCALE, chunk_pairs=2)
    assert out.shape == (7, 2, 21, N) and osc == SCALE
    d = kg.decrypt_decode(out, osc, o)
    err = np.abs(d - msgs).max(axis=1)
    print("bootstrap_real (N = 65536, HW-192): max |out - msg| per ciphertext =", ["%.2g" % v for v in err])
    assert err.max() < 1e-4


@pytest.mark.parametrize("name,variant,tol_abs,tol_rel,tol_ref", [("ln1", 1, 2e-3, 1e-3, 1e-4), ("ln2", 2, 0.25, 0.1, 5e-3)])
def test_layernorm_golden(pkg, env, name, variant, tol_abs, tol_rel, tol_ref):
    """layernorm / layernorm2 (layernorm.hpp:157-547) on layer 0's residual sums.
    (a) vs the REFERENCE'S decrypted output of the same module on the same activations (oracle/_ref, N = 8192):
        rel <= 1e-4 (ln1) / 5e-3 (ln2), rel = |err| / max(1, |expected|) (the modules' forced scale resets make the
        result depend on the primes, hence on the ring degree, at that level);
    (b) vs real_self_output.csv / real_final_output.csv: the reference's own code is 4.5e-4 (ln1) and 0.13 max-abs /
        5.7 % rel (ln2: its inverse square root is a low-degree iteration) away from these; tolerances ln1 2e-3 abs /
        1e-3 rel, ln2 0.25 abs / 0.1 rel."""
    be, keys, g = env["be"], env["keys"], env["g"]
    x = encrypt_cols(env, pack_rows(g[name + "_in"]), 21)
    out, osc = be.layernorm(keys, x, SCALE, g[name + "_gamma"], g[name + "_beta"], env["mask"], variant=variant)
    del x
    assert out.shape[0] == 768 and out.shape[2] == 1
    got = slot_values(env, out, osc, VALID).real.T        # [TOK, 768]
    exp = g[name + "_out"]
    err = np.abs(got - exp)
    rel = (err / np.maximum(1.0, np.abs(exp))).max()
    want = env["ref"][name + "_ref"]
    rel_ref = (np.abs(got - want) / np.maximum(1.0, np.abs(want))).max()
    print("%s golden: vs reference decrypted rel %.3g (max-abs %.3g); vs CSV max-abs %.3g, rel %.3g"
          % (name, rel_ref, np.abs(got - want).max(), err.max(), rel))
    assert rel_ref < tol_ref and err.max() < tol_abs and rel < tol_rel


def test_gelu_golden(pkg, env):
    """gelu_v2 (gelu_others.hpp:4-154) on all 3072 intermediate features of layer 0.
    (a) vs the REFERENCE'S decrypted gelu_v2 output: every entry with |x| <= 9 against the N = 8192 run (2e-3), and the
        17 columns of the N = 65536 run — all columns with an input beyond |x| = 9 plus the first eight — entry by entry
        (2e-3).  The second set matters: gelu_v2's forced scale resets leave an error that depends on the primes and
        that its degree-24 polynomial amplifies where |0.1 x| > 1 (reference at x = -14.7: 0.37 from the true GELU at
        N = 65536, 0.08 at N = 8192), and a faithful implementation reproduces exactly that.
    (b) vs real_intermediate_output.csv for |x| <= 9: 0.05 max-abs (the reference is up to 0.04 away there)."""
    be, keys, g, ref = env["be"], env["keys"], env["g"], env["ref"]
    V = pack_rows(g["gelu_in"])
    got = np.zeros((TOK, 3072))
    for c0 in range(0, 3072, 256):
        x = encrypt_cols(env, V[c0:c0 + 256], 9)
        out, osc = be.gelu_v2(keys, x, SCALE)
        del x
        assert out.shape[2] == 2
        got[:, c0:c0 + 256] = slot_values(env, out, osc, VALID).real.T
        del out
    mod = np.abs(g["gelu_in"]) <= 9.0
    e_ref = np.abs(got - ref["gelu_ref"])[mod].max()
    e_csv = np.abs(got - g["gelu_out"])[mod].max()
    cols = ref["gelu_cols65536"]
    e_full = np.abs(got[:, cols] - ref["gelu_ref65536"]).max()
    worst = np.abs(got - g["gelu_out"]).max()
    print("gelu_v2 golden: |x| <= 9: vs reference decrypted (N = 8192) %.3g, vs CSV %.3g; the %d extreme columns vs the "
          "reference at N = 65536: %.3g (they are up to %.3g from the true GELU, as the reference is)"
          % (e_ref, e_csv, len(cols), e_full, worst))
    assert e_ref < 2e-3 and e_full < 2e-3 and e_csv < 0.05


def softmax_scale_drift(q, s0=SCALE):
    """The reference's modules overwrite the scale after rescaling (`x.scale() = scale`).  The value a
    ciphertext decrypts to is then off by (true scale / 2^46), a factor fixed by the primes.  For softmax_boot:
    f_exp (exp + mask, softmax.hpp:9-47, 407-467), f_inv (inverse, :49-82, reset at :545) and f_out (:565-576)."""
    lv = 12
    s = s0 * s0 / q[lv]
    lv -= 1
    for _ in range(7):
        s = s * s / q[lv]
        lv -= 1
    s = s * s / q[lv]
    f_exp = s / s0
    sy, sr, ly = s0, s0, 20
    for _ in range(16):
        sy = sy * sy / q[ly]
        ly -= 1
        sr = sr * sy / q[ly]
    return f_exp, sr / s0, (s0 * s0 / q[ly - 1]) / s0


def test_softmax_golden(pkg, env):
    """softmax_boot (softmax.hpp:308-581, one bootstrapping inside) on layer 0's scores, all 12 heads.
    (a) vs the float64 model of exactly what the module computes: exp as (1 + (x - 7.5) / 128)^128 (:324), a 16-step
        Goldschmidt inverse of the bootstrapped row sums, and the forced scale resets, which at this ring's primes
        inflate every probability by f_inv = 1 + 9.2e-3 (1 + 1.1e-3 at N = 8192: the reference's decrypted output
        from oracle/_ref matches the same model with its own primes to 1.1e-4).  Per row the bootstrapping error on the
        row sum (8e-6 at this size) comes back multiplied by 1 / sum (sums are as small as 2e-3):
        tolerance 2e-4 + 2e-5 / sum.
    (b) vs aftsoftmax.csv (true softmax): 4e-2 max-abs; the reference is 1.2e-2 ... 2.3e-2 away (its exp)."""
    be, kg, boot, keys, g = env["be"], env["kg"], env["boot"], env["keys"], env["g"]
    f_exp, f_inv, f_out = softmax_scale_drift([float(p) for p in be.primes])
    worst_csv, worst_model, worst_ratio = 0.0, 0.0, 0.0
    for h in range(12):
        S = g["QKT"][:, 5 * h:5 * h + 5]
        want = g["aftsoftmax"][:, 5 * h:5 * h + 5]
        E = (1 + (S - 7.5) / 128.0) ** 128
        model = E * f_exp / ((E * f_exp).sum(axis=1, keepdims=True) + 1e-5) * f_inv * f_out
        V = np.zeros((128, SLOTS))
        for i in range(128):
            for k in range(TOK):
                if (k + i) % 128 < TOK:
                    V[i, k * NUM_BATCH] = S[k, (k + i) % 128]     # i-th generalized diagonal, Ct_ct_matrix_mul.hpp:24-41
        x = encrypt_cols(env, V, 13)
        out, osc = boot.softmax_boot(keys, x, SCALE, env["mask"], TOK, iters=16, layer_id=0)
        del x
        assert out.shape == (128, 2, 3, N)
        got = slot_values(env, out, osc, VALID).real      # [128 diagonals, TOK]
        P = np.zeros((TOK, TOK))
        for i in range(128):
            for k in range(TOK):
                if (k + i) % 128 < TOK:
                    P[k, (k + i) % 128] = got[i, k]
        worst_csv = max(worst_csv, np.abs(P - want).max())
        worst_model = max(worst_model, np.abs(P - model).max())
        row_tol = 2e-4 + 2e-5 / E.sum(axis=1)
        worst_ratio = max(worst_ratio, (np.abs(P - model).max(axis=1) / row_tol).max())
    print("softmax_boot golden (12 heads): max-abs vs the float64 model %.3g (worst row at %.2f of its bound "
          "2e-4 + 2e-5 / sum; scale drift f_inv - 1 = %.3g), vs aftsoftmax.csv %.3g"
          % (worst_model, worst_ratio, f_inv - 1, worst_csv))
    assert worst_ratio < 1.0 and worst_csv < 4e-2


def test_c1_selfoutput_linear_with_reference_mask(pkg, env):
    """Config C1 as the reference runs it: data/selfoutput_linear.txt (5 tokens x 768) in input 0, mask {5, 0, ...},
    ct_pt_matrix_mul_wo_pre_w_mask 768 x 768 at chain_index 1 -> 0 (test_full_scheme.hpp:601-617).  Decrypted result
    vs float64 X W: max-abs <= 1e-4 relative to max |X W| (the reference achieved 1.2e-5)."""
    be, g = env["be"], env["g"]
    X = g["selfoutput_linear"]
    rng = np.random.default_rng(20250991)
    W = rng.normal(size=(768, 768)) * 0.04
    x = encrypt_cols(env, pack_rows(X), 2)
    import time
    import torch
    out = be.ct_pt_matrix_mul_wo_pre_w_mask(x, W, env["mask"], SCALE)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = be.ct_pt_matrix_mul_wo_pre_w_mask(x, W, env["mask"], SCALE)
    torch.cuda.synchronize()
    ms_masked = (time.perf_counter() - t0) * 1e3
    ones = np.ones_like(env["mask"])
    be.ct_pt_matrix_mul_wo_pre_w_mask(x, W, ones, SCALE)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    be.ct_pt_matrix_mul_wo_pre_w_mask(x, W, ones, SCALE)
    torch.cuda.synchronize()
    ms_all = (time.perf_counter() - t0) * 1e3
    # the exact masked path encodes one plaintext per (weight, limb) like the reference; the all-valid mask is the scalar
    # case and takes the tensor-core GEMM
    # fast mode: [sum_j round(w s_w) X_j] (.) encode(mask at 2^26) — one GEMM, one plaintext
    fast = be.ct_pt_matrix_mul_wo_pre_w_mask_fast(x, W, env["mask"], SCALE)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fast = be.ct_pt_matrix_mul_wo_pre_w_mask_fast(x, W, env["mask"], SCALE)
    torch.cuda.synchronize()
    ms_fast = (time.perf_counter() - t0) * 1e3
    print("C1 768 x 768 at 2 limbs, second call: reference's 5-token mask %.1f ms exact / %.1f ms factorised, all-valid "
          "mask %.1f ms" % (ms_masked, ms_fast, ms_all))
    gotf = slot_values(env, fast, SCALE, VALID + [1, 257]).real.T
    errf = np.abs(gotf[:TOK] - X @ W).max()
    print("C1 masked matmul, factorised: max-abs %.3g, masked-out slots %.3g" % (errf, np.abs(gotf[TOK:]).max()))
    assert errf < 1e-4 * max(1.0, np.abs(X @ W).max())
    assert np.abs(gotf[TOK:]).max() < 1e-6
    assert out.shape == (768, 2, 1, N)
    got = slot_values(env, out, SCALE, VALID + [1, 257]).real.T
    exp = X @ W
    err = np.abs(got[:TOK] - exp).max()
    print("C1 masked matmul: max-abs %.3g (max |XW| %.3g), masked-out slots %.3g" % (err, np.abs(exp).max(), np.abs(got[TOK:]).max()))
    assert err < 1e-4 * max(1.0, np.abs(exp).max())
    assert np.abs(got[TOK:]).max() < 1e-6                 # other inputs' slots stay zero


GOLDEN_REST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layers_1_11_activations.npz")
# measured max errors of this repo's modules against the reference's plaintext CSVs, layer by layer (grouped keys;
# profiles/fullsize_golden_r2_layers.log), and the bound asserted for each stage.  The bounds are those of layer 0:
# what they measure is the REFERENCE'S approximation (Newton inverse square root, degree-24 GELU polynomial), which this
# repo reproduces operation for operation (bit for bit in exact mode, tests/test_gpu_modules.py).
LN1_TOL_ABS, GELU_TOL_ABS, LN2_TOL_REL = 2e-3, 0.05, 0.1


@pytest.mark.parametrize("layer", list(range(1, 12)))
def test_layers_1_to_11_golden(pkg, env, request, layer):
    """The LayerNorm / GELU / LayerNorm2 stages of encoder layers 1..11 on the reference's own activations
    (tests/golden/layers_1_11_activations.npz from /root/reference/data/layer_k/**/allresults/*.csv), against the CSV
    outputs: LayerNorm 2e-3 max-abs (measured 2.6e-5 ... 1.6e-3), gelu_v2 0.05 max-abs (measured 0.030 ... 0.037) on the
    ciphertexts — columns: the 5 tokens of a feature share one ciphertext — all of whose inputs lie inside the
    polynomial's domain |x| <= 9, LayerNorm2 0.1 relative (|err| / max(1, |expected|); measured 0.03 ... 0.08) — the same
    bounds as layer 0, where the reference's own decrypted outputs show that these distances are its approximations'
    and not the arithmetic's.  Where the reference's algorithm leaves its own domain the gate only reports:
    * gelu_v2 columns with an input beyond |x| = 9 (up to |x| = 123 in layer 10): the degree-24 polynomial in 0.1 x
      reaches 1e18 and wraps around the ciphertext modulus, which destroys every slot of that ciphertext — in the
      reference's encrypted run exactly as here;
    * LayerNorm2 of layers 10 and 11 (residual sums up to 955 / variance beyond the initial guess of its Newton inverse
      square root, layernorm.hpp:18-24): measured 1.95 and 0.126 relative.
    Run in the benchmarked key mode only."""
    if "grouped" not in request.node.name:
        pytest.skip("layers 1..11 are gated in the benchmarked key mode")
    be, keys = env["be"], env["keys"]
    g = np.load(GOLDEN_REST)
    p = "l%d_" % layer
    get = lambda n: g[p + n].astype(np.float64)
    res = {}
    for name, variant in (("ln1", 1), ("ln2", 2)):
        x = encrypt_cols(env, pack_rows(get(name + "_in")), 21)
        out, osc = be.layernorm(keys, x, SCALE, get(name + "_gamma"), get(name + "_beta"), env["mask"], variant=variant)
        del x
        got = slot_values(env, out, osc, VALID).real.T
        del out
        exp = get(name + "_out")
        res[name] = (np.abs(got - exp).max(), (np.abs(got - exp) / np.maximum(1.0, np.abs(exp))).max())
    V = pack_rows(get("gelu_in"))
    got = np.zeros((TOK, 3072))
    for c0 in range(0, 3072, 512):
        x = encrypt_cols(env, V[c0:c0 + 512], 9)
        out, osc = be.gelu_v2(keys, x, SCALE)
        del x
        got[:, c0:c0 + 512] = slot_values(env, out, osc, VALID).real.T
        del out
    cols = (np.abs(get("gelu_in")) <= 9.0).all(axis=0)     # ciphertexts whose five inputs all are inside the domain
    e_gelu = np.abs(got - get("gelu_out"))[:, cols].max()
    print("layer %d golden: ln1 max-abs %.3g (rel %.3g); gelu_v2 on the %d of 3072 in-domain columns %.3g; ln2 max-abs %.3g, "
          "rel %.3g" % (layer, res["ln1"][0], res["ln1"][1], int(cols.sum()), e_gelu, res["ln2"][0], res["ln2"][1]))
    assert res["ln1"][0] < LN1_TOL_ABS and e_gelu < GELU_TOL_ABS
    assert np.isfinite(res["ln2"][1])
    if layer <= 9:
        assert res["ln2"][1] < LN2_TOL_REL
