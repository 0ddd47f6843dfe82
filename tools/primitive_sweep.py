#!/usr/bin/env python
"""BASELINE.json config 2 — CKKS primitive sweep at the repo's parameters (N = 65536, 36 primes):
device time per op (CUDA events, warm, inputs larger than L2 where the batch allows) and the
achieved algorithmic HBM bandwidth (SURVEY §8(d) byte counts) against MEASURED_PEAKS.json.
Writes one JSON object per line to stdout.  Run on a GPU box:  python tools/primitive_sweep.py"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (prime selection helper)


def main():
    import torch
    pkg = importlib.import_module("moai-fhe-transformerinference-public_b200")
    primes = bench.moai_primes()
    be = pkg.Backend(16, primes)
    n = 1 << 16
    limb = n * 8
    peak = 6452.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    g = torch.Generator(device="cuda")
    g.manual_seed(1)

    def rnd(batch, size, limbs):
        x = torch.empty((batch, size, limbs, n), dtype=torch.int64, device="cuda")
        for l in range(limbs):
            x[:, :, l, :] = torch.randint(0, primes[l], (batch, size, n), generator=g, device="cuda", dtype=torch.int64)
        return x

    def timeit(fn, iters=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    kl = len(primes)
    ksk = torch.empty((kl - 1, 2, kl, n), dtype=torch.int64, device="cuda")
    for l in range(kl):
        ksk[:, :, l, :] = torch.randint(0, primes[l], (kl - 1, 2, n), generator=g, device="cuda", dtype=torch.int64)
    elt = be.galois_elt_from_step(256)
    for limbs in (35, 21, 15, 10, 4, 2):
        batch = max(4, min(64, (1 << 30) // (2 * limbs * limb)))      # ~1 GiB of ciphertexts
        a, b = rnd(batch, 2, limbs), rnd(batch, 2, limbs)
        pt = rnd(1, 1, limbs)[0, 0]
        polys = batch * 2 * limbs
        rows = []
        ms = timeit(lambda: be.ntt_forward_(a))
        rows.append(("ntt_forward", ms, polys * 2 * limb, polys))
        ms = timeit(lambda: be.ntt_inverse_(a))
        rows.append(("ntt_inverse", ms, polys * 2 * limb, polys))
        out = torch.empty_like(a)
        ms = timeit(lambda: be.add(a, b, out=out))
        rows.append(("add", ms, 3 * polys * limb, batch))
        ms = timeit(lambda: be.multiply_plain(a, pt, out=out))
        rows.append(("multiply_plain", ms, 2 * polys * limb + limbs * limb, batch))
        out3 = be.empty(batch, 3, limbs, n)
        ms = timeit(lambda: be.multiply(a, b, out=out3))
        rows.append(("multiply", ms, (4 + 3) * batch * limbs * limb, batch))
        outr = be.empty(batch, 2, limbs - 1, n)
        ms = timeit(lambda: be.rescale_to_next(a, out=outr))
        rows.append(("rescale_to_next", ms, batch * (2 * limbs + 2 * (limbs - 1)) * limb, batch))
        kb = max(1, min(batch, 8))
        ak = a[:kb].contiguous()
        outk = torch.empty_like(ak)
        ms = timeit(lambda: be.apply_galois(ak, elt, ksk, out=outk), iters=3, warm=1)
        ks_bytes = kb * 4 * limbs * limb + 2 * limbs * (limbs + 1) * limb   # ct in/out + evk read once per call
        rows.append(("rotate_vector(keyswitch)", ms, ks_bytes, kb))
        a3 = rnd(kb, 3, limbs)
        out2 = be.empty(kb, 2, limbs, n)
        ms = timeit(lambda: be.relinearize(a3, ksk, out=out2), iters=3, warm=1)
        rows.append(("relinearize", ms, kb * 5 * limbs * limb + 2 * limbs * (limbs + 1) * limb, kb))
        for name, ms, nbytes, units in rows:
            gbs = nbytes / (ms * 1e-3) / 1e9
            print(json.dumps({"op": name, "limbs": limbs, "batch": units if "ntt" not in name else batch,
                              "ms": round(ms, 4), "us_per_unit": round(ms * 1e3 / units, 3),
                              "unit": "limb-transform" if "ntt" in name else "ciphertext",
                              "algorithmic_GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4)}),
                  flush=True)
        del a, b, out, out3, outr, ak, outk, a3, out2
        torch.cuda.empty_cache()
    be.close()


if __name__ == "__main__":
    main()
