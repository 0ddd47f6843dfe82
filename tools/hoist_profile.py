#!/usr/bin/env python
"""Times the hoisted baby-step rotations of one bootstrapping linear stage (rotate_many, 15 steps) at the
repo's parameters.  usage: python tools/hoist_profile.py [limbs] [batch] [steps]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    pkg = importlib.import_module("moai-fhe-transformerinference-public_b200")
    primes = bench.moai_primes()
    be = pkg.Backend(16, primes)
    limbs = int(sys.argv[1]) if len(sys.argv) > 1 else 35
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 15
    n, kl = 1 << 16, len(primes)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    x = torch.empty((batch, 2, limbs, n), dtype=torch.int64, device="cuda")
    for l in range(limbs):
        x[:, :, l, :] = torch.randint(0, primes[l], (batch, 2, n), generator=g, device="cuda", dtype=torch.int64)
    gal = {}
    steps = list(range(1, nsteps + 1))
    for st in steps:
        key = torch.empty((kl - 1, 2, kl, n), dtype=torch.int64, device="cuda")
        for l in range(kl):
            key[:, :, l, :] = torch.randint(0, primes[l], (kl - 1, 2, n), generator=g, device="cuda", dtype=torch.int64)
        gal[be.galois_elt_from_step(st)] = key
    keys = be.make_keys(galois_fast=gal)
    be.rotate_many(keys, x, steps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        be.rotate_many(keys, x, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"op": "rotate_many (hoisted)", "limbs": limbs, "batch": batch, "steps": nsteps,
                      "multi_key": os.environ.get("MOAI_KSM_MULTI", "1"), "ms": round(ms, 2),
                      "us_per_rotation_per_ct": round(ms * 1000 / batch / nsteps, 1)}))
    be.close()


if __name__ == "__main__":
    main()
