#!/usr/bin/env python
"""Per-kernel device time of ONE module call at the repo's parameters on synthetic residues (timing only):
   python tools/module_profile.py gelu --batch 256 --limbs 9
   python tools/module_profile.py relin --batch 64 --limbs 28"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["gelu", "relin", "relin_rescale", "rotate", "rescale"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--limbs", type=int, default=9)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--seal-digits", action="store_true", help="SEAL's per-prime digits instead of grouped keys")
    args = ap.parse_args()
    pkg = importlib.import_module("moai-fhe-transformerinference-public_b200")
    primes = bench.moai_primes()
    be = pkg.Backend(16, primes)
    n, kl = 1 << 16, len(primes)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)

    def rand_key(levels=kl - 1):
        ids = list(range(levels)) + [kl - 1]
        k = torch.empty((levels, 2, levels + 1, n), dtype=torch.int64, device="cuda")
        for pos, l in enumerate(ids):
            k[:, :, pos, :] = torch.randint(0, primes[l], (levels, 2, n), generator=g, device="cuda", dtype=torch.int64)
        return k

    def rand_ct(size, limbs):
        x = torch.empty((args.batch, size, limbs, n), dtype=torch.int64, device="cuda")
        for l in range(limbs):
            x[:, :, l, :] = torch.randint(0, primes[l], (args.batch, size, n), generator=g, device="cuda", dtype=torch.int64)
        return x

    grouped = {}
    if not args.seal_digits:
        grouped[0] = [be.random_grouped_key(k, lv, g) for k, lv in sorted(be.ksg_plan(range(1, kl - 1)).items())]
    e1 = be.galois_elt_from_step(1)
    fast = {}
    if args.what == "rotate":
        k = be.ksg_best_extra(args.limbs)
        if k and not args.seal_digits:
            grouped[e1] = [be.random_grouped_key(k, args.limbs, g)]
        else:
            fast[e1] = [rand_key(args.limbs)]
    keys = be.make_keys(relin=rand_key(), galois_fast=fast, grouped=grouped)
    if args.what == "gelu":
        x = rand_ct(2, args.limbs)
        run = lambda: be.gelu_v2(keys, x, 2.0 ** 46)
    elif args.what == "relin":
        x = rand_ct(3, args.limbs)
        run = lambda: be.relinearize_keys(keys, x)
    elif args.what == "relin_rescale":
        x = rand_ct(3, args.limbs)
        run = lambda: be.relin_rescale_keys(keys, x)
    elif args.what == "rescale":
        x = rand_ct(2, args.limbs)
        run = lambda: be.rescale_to_next(x)
    else:
        x = rand_ct(2, args.limbs)
        run = lambda: be.rotate_vector_keys(keys, x, 1)
    run()
    torch.cuda.synchronize()
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        run()
    e1_.record()
    torch.cuda.synchronize()
    l0 = be.launch_count()
    be.profile(True)
    run()
    dump = be.profile_dump()
    be.profile(False)
    ms = e0.elapsed_time(e1_) / args.iters
    kern = {k: [round(v[0], 3), v[1]] for k, v in sorted(dump.items(), key=lambda kv: -kv[1][0]) if k.startswith("k_")}
    print(json.dumps({"what": args.what, "batch": args.batch, "limbs": args.limbs, "grouped": not args.seal_digits,
                      "ms": round(ms, 3), "us_per_ct": round(ms * 1e3 / args.batch, 2), "launches": be.launch_count() - l0,
                      "kernel_ms_sum": round(sum(v[0] for v in kern.values()), 3), "kernels_ms": kern}))
    be.close()


if __name__ == "__main__":
    main()
