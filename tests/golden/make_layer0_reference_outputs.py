#!/usr/bin/env python
"""Writes tests/golden/layer0_reference_decrypted.npz: what the REFERENCE'S OWN encrypted modules output (decrypted)
for the layer-0 activations of tests/golden/layer0_activations.npz.

The reference's unmodified headers (softmax.hpp incl. its Bootstrapper, layernorm.hpp, gelu_others.hpp) run inside
oracle/_ref on real SEAL at N = 8192 (the smallest ring its bootstrapper supports, see tests/test_gpu_boot_reference.py)
with the repo's 36-prime chain, scale 2^46, the 5-token sentence in input 0 and the {5, 0, ...} token mask — the
run M/test/test_full_scheme.hpp:339-1087 performs at N = 65536.  The modules act slot-wise, so the decrypted values in
the 5 valid slots do not depend on the ring degree beyond noise (1e-6).  These are the "reference's decrypted outputs"
BASELINE.json's north_star asks the GPU path to match.  About 20 minutes on 8 cores.
Run where /root/reference exists: python tests/golden/make_layer0_reference_outputs.py"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

LOG_N, TOK, SCALE = 13, 5, 2.0 ** 46
N, SLOTS = 1 << LOG_N, (1 << LOG_N) // 2
NB = SLOTS // 128
BITS = [51] + [46] * 20 + [51] * 14 + [58]
VALID = [k * NB for k in range(TOK)]

oracle.build_ref()
g = np.load(os.path.join(HERE, "layer0_activations.npz"))
r = oracle.SealRef(LOG_N, BITS, hamming_weight=64, seed=2025)
r.set_threads()
r.make_relin_key()
steps = r.boot_create()
r.make_galois_keys(steps, conjugate=True)
r.boot_prepare()
mask = np.zeros(SLOTS, dtype=np.int32)
mask[VALID] = 1


def enc(v, limbs):
    return r.encrypt(r.encode(np.asarray(v, dtype=np.complex128), SCALE, limbs), limbs, SCALE)


def pack(A, limbs):
    out = np.empty((A.shape[1], 2 * limbs * N), dtype=np.uint64)
    for c in range(A.shape[1]):
        v = np.zeros(SLOTS)
        v[VALID] = A[:, c]
        out[c] = enc(v, limbs)
    return out


def dec_valid(flat, count, limbs, scale):
    o = flat.reshape(count, -1)
    return np.stack([r.decode(r.decrypt(o[c], 2, limbs, scale), limbs, scale).real[VALID] for c in range(count)], axis=1)


res = {}
t0 = time.time()
for name, variant in (("ln1", 1), ("ln2", 2)):
    out, ol, osc = r.layernorm(variant, pack(g[name + "_in"], 21).reshape(-1), 768, 21, SCALE, g[name + "_gamma"],
                               g[name + "_beta"], mask)
    res[name + "_ref"] = dec_valid(out, 768, ol, osc)                       # [TOK, 768]
    print(name, "max |ref - csv|", np.abs(res[name + "_ref"] - g[name + "_out"]).max(), "%.0f s" % (time.time() - t0), flush=True)
gel = []
for c0 in range(0, 3072, 256):
    out, ol, osc = r.gelu_v2(pack(g["gelu_in"][:, c0:c0 + 256], 9).reshape(-1), 256, 9, SCALE)
    gel.append(dec_valid(out, 256, ol, osc))
    print("gelu", c0, "%.0f s" % (time.time() - t0), flush=True)
res["gelu_ref"] = np.concatenate(gel, axis=1)                              # [TOK, 3072]
print("gelu max |ref - csv|", np.abs(res["gelu_ref"] - g["gelu_out"]).max(), flush=True)
sm = np.zeros((TOK, 60))
for h in range(12):
    S = g["QKT"][:, 5 * h:5 * h + 5]
    cts = np.empty((128, 2 * 13 * N), dtype=np.uint64)
    for i in range(128):
        v = np.zeros(SLOTS)
        for k in range(TOK):
            if (k + i) % 128 < TOK:
                v[k * NB] = S[k, (k + i) % 128]                            # i-th generalized diagonal of QK^T
        cts[i] = enc(v, 13)
    out, ol, osc = r.softmax_boot(cts, 128, 13, SCALE, mask, TOK, 16, 0)
    o = out.reshape(128, -1)
    for i in range(128):
        d = r.decode(r.decrypt(o[i], 2, ol, osc), ol, osc).real
        for k in range(TOK):
            if (k + i) % 128 < TOK:
                sm[k, 5 * h + (k + i) % 128] = d[k * NB]
    print("softmax head", h, "max |ref - csv|", np.abs(sm[:, 5 * h:5 * h + 5] - g["aftsoftmax"][:, 5 * h:5 * h + 5]).max(),
          "%.0f s" % (time.time() - t0), flush=True)
res["softmax_ref"] = sm
np.savez_compressed(os.path.join(HERE, "layer0_reference_decrypted.npz"), **res)
print("done in %.0f s" % (time.time() - t0))
