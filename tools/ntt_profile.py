#!/usr/bin/env python
"""Small driver for ncu: forward + inverse NTT of a ciphertext batch at the repo's parameters."""
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
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    x = torch.empty((batch, 2, limbs, 1 << 16), dtype=torch.int64, device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    for l in range(limbs):
        x[:, :, l, :] = torch.randint(0, primes[l], (batch, 2, 1 << 16), generator=g, device="cuda", dtype=torch.int64)
    for _ in range(3):
        be.ntt_forward_(x)
        be.ntt_inverse_(x)
    torch.cuda.synchronize()
    be.close()
    print("ok")


if __name__ == "__main__":
    main()
