#!/usr/bin/env python
"""BASELINE.json config 3 — bootstrapping microbench at the repo's parameters (N = 65536, 35 + 1
primes, full slots).  Times Bootstrapper::bootstrap_3 on a batch of ciphertexts with CUDA events and
prints per-ciphertext milliseconds plus the projected per-layer / per-input cost
(4 x 768 + 12 bootstrappings per layer, M/test/test_full_scheme.hpp:654,758,991,1081; softmax.hpp:536).
Key material is uniformly random residues: the timing of a key switch does not depend on key values
(correctness with valid keys is covered by tests/test_gpu_bootstrap.py).
usage: python tools/bootstrap_bench.py [--batch 16] [--keys required|pow2]"""
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
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--keys", default="required", choices=["required", "pow2"])
    ap.add_argument("--fast", action="store_true", help="hoisted baby steps with pre-permuted keys (fast mode)")
    ap.add_argument("--real", action="store_true", help="real-slot messages, two per bootstrapping (moai_bootstrap_real); "
                                                        "--batch counts ciphertexts, so batch/2 bootstrappings run")
    ap.add_argument("--grouped", action="store_true", help="fast mode with grouped-digit keys (csrc/ksgroup.hpp): every "
                                                           "Galois key in the digit layout its level prefers, the "
                                                           "relinearisation key in one variant per layout")
    ap.add_argument("--lazy", action="store_true", help="with --grouped: hoisting mode 2 (baby-step rotations stay in the "
                                                        "key-switch basis, single-digit keys on the first CoeffToSlot stage)")
    args = ap.parse_args()
    args.grouped = args.grouped or args.lazy
    args.fast = args.fast or args.grouped
    pkg = importlib.import_module("moai-fhe-transformerinference-public_b200")
    primes = bench.moai_primes()
    be = pkg.Backend(16, primes)
    n, kl = 1 << 16, len(primes)
    boot = pkg.Bootstrapper(be, total_limbs=35)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)

    def rand_key():
        k = torch.empty((kl - 1, 2, kl, n), dtype=torch.int64, device="cuda")
        for l in range(kl):
            k[:, :, l, :] = torch.randint(0, primes[l], (kl - 1, 2, n), generator=g, device="cuda", dtype=torch.int64)
        return k

    if args.fast:
        boot.set_hoisting(2 if args.lazy else 1)
    if args.keys == "required":
        steps = boot.required_steps()
    else:
        steps = [1 << k for k in range(15)] + [(n // 2) - (1 << k) for k in range(15)]
    gal = {}
    for st in steps + [0]:
        gal[be.galois_elt_from_step(st)] = rand_key()
    # random residues stand in for key material; a pre-permuted key is the same size and layout
    key_bytes = 0
    if args.grouped:
        gal, grouped, single = {}, {}, {}
        for st, lvs in boot.required_step_levels().items():
            e = be.galois_elt_from_step(st)
            for lv in lvs:
                if lv == 0:   # first CoeffToSlot stage, baby step: single-digit key [1, 2, kl, n]
                    single[e] = rand_key()[:1].clone()
                    key_bytes += single[e].numel() * 8
                    continue
                k = be.ksg_best_extra(lv)
                if k == 0:   # no spare prime at this level: SEAL's digits, level-truncated
                    t = rand_key()[:lv, :, :lv + 1, :].contiguous()
                    gal.setdefault(e, []).append(t)
                    key_bytes += t.numel() * 8
                else:
                    gk = be.random_grouped_key(k, lv, g)
                    grouped.setdefault(e, []).append(gk)
                    key_bytes += gk.t.numel() * 8
        grouped[0] = [be.random_grouped_key(k, lv, g) for k, lv in sorted(be.ksg_plan(range(1, 35)).items())]
        key_bytes += sum(gk.t.numel() * 8 for gk in grouped[0])
        keys = be.make_keys(relin=rand_key(), galois_fast=gal, grouped=grouped, single=single)
    else:
        keys = be.make_keys(relin=rand_key(), galois_fast=gal) if args.fast else be.make_keys(relin=rand_key(), galois=gal)
    x = torch.randint(0, primes[0], (args.batch, 2, 1, n), generator=g, device="cuda", dtype=torch.int64)
    run = (lambda: boot.bootstrap_real(keys, x, 2.0 ** 46, chunk_pairs=32)) if args.real else \
        (lambda: boot.bootstrap_3(keys, x, 2.0 ** 46))
    run()                                        # warm-up: encodes the linear-transform plaintexts once
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = bench.ClockSampler(0)
    sampler.start()
    e0.record()
    for _ in range(args.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    be.profile(True)                             # per-phase breakdown from one more (event-bracketed) call
    run()
    dump = be.profile_dump()
    phases = {k: round(v[0] / args.batch, 2) for k, v in dump.items() if k.startswith("boot_")}
    phases["allocator"] = {k: (round(v[0], 1), v[1]) for k, v in dump.items() if k.startswith("alloc_")}
    # live per-kernel device time of that call (deferred CUDA-event pairs around every launch): [ms, units]
    kernels = {k: [round(v[0], 2), v[1]] for k, v in sorted(dump.items(), key=lambda kv: -kv[1][0]) if k.startswith("k_")}
    be.profile(False)
    ms = e0.elapsed_time(e1) / args.iters
    per_ct = ms / args.batch
    per_layer_s = per_ct * 3084 / 1000.0
    print(json.dumps({"op": "bootstrap_real (two real-slot ciphertexts per bootstrapping)" if args.real else "bootstrap_3", "batch": args.batch, "keys": args.keys, "mode": ("fast (hoisted, grouped digits%s)" % (", lazy mod-down + single-digit first stage" if args.lazy else "") if args.grouped else "fast (hoisted)") if args.fast else "exact (SEAL key switch)",
                      "grouped_key_GiB": round(key_bytes / 2 ** 30, 2),
                      "galois_keys": len(gal), "clocks": clocks,
                      "ms_per_batch": round(ms, 2), "ms_per_ciphertext": round(per_ct, 2), "phase_ms_per_ciphertext": phases,
                      "projected_bootstrap_s_per_layer": round(per_layer_s, 1),
                      "projected_bootstrap_s_per_input_12_layers": round(per_layer_s * 12 / 256, 2),
                      "reference_bootstrap_s_per_input_12_layers": 384.8,
                      "kernels_ms_one_call": kernels,
                      "gpu_mem_GiB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}))
    be.close()


if __name__ == "__main__":
    main()
