#!/usr/bin/env python
"""Small driver for ncu: one SEAL-exact rotation (key switch) and one hoisted rotate_many of a batch
at the repo's parameters.  usage: python tools/ks_profile.py [limbs] [batch]"""
import importlib
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
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    n, kl = 1 << 16, len(primes)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    x = torch.empty((batch, 2, limbs, n), dtype=torch.int64, device="cuda")
    for l in range(limbs):
        x[:, :, l, :] = torch.randint(0, primes[l], (batch, 2, n), generator=g, device="cuda", dtype=torch.int64)
    key = torch.empty((kl - 1, 2, kl, n), dtype=torch.int64, device="cuda")
    for l in range(kl):
        key[:, :, l, :] = torch.randint(0, primes[l], (kl - 1, 2, n), generator=g, device="cuda", dtype=torch.int64)
    elt = be.galois_elt_from_step(1)
    reps = int(os.environ.get("KS_REPS", "2"))
    for _ in range(reps):
        be.apply_galois(x, elt, key)
    torch.cuda.synchronize()
    be.close()
    print("ok")


if __name__ == "__main__":
    main()
