#!/usr/bin/env python
"""BASELINE.json config 4 — one full BERT-base encoder layer (12 heads, hidden 768, FFN 3072;
4 x 768 + 12 bootstrappings) on one packed batch of 256 inputs x 128 tokens at the repo's CKKS
parameters (N = 65536, 35 + 1 primes), all 128 token slots valid, on ONE B200.
Inputs, weights and evaluation keys are synthetic (uniform residues / Gaussian weights): kernel
timing does not depend on the values; correctness of the same pipeline is covered at N = 4096 by
tests/test_gpu_layer.py.  Prints the per-stage device time (the rows of P:Table 3) and the
amortized seconds per input extrapolated to 12 identical layers."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

REFERENCE_TABLE3 = {  # P:Table 3, seconds per input summed over 12 layers
    "attention": 37.4 + 40.3 + 53.3 + 1.4, "att_qkv_matmul": 37.4, "att_qk_colpacking": 40.3, "att_softmax_boot": 53.3,
    "att_sv_diagpacking": 1.4, "selfoutput_matmul": 1.7, "bootstrap_1": 95.4, "layernorm_1": 0.6,
    "bootstrap_2": 95.8, "intermediate_matmul": 44.1, "gelu": 3.3, "final_matmul": 7.1, "bootstrap_3": 98.8,
    "layernorm_2": 0.6, "bootstrap_4": 94.8}


class Args:
    def __init__(self, layers=1, mode="fast"):
        self.layers, self.mode = layers, mode


def setup(args, device=0, seed=11):
    """Backend, bootstrapper, synthetic keys / weights / input for one packed batch on `device`."""
    import torch
    pkg = importlib.import_module("moai-fhe-transformerinference-public_b200")
    primes = bench.moai_primes()
    torch.cuda.set_device(device)
    be = pkg.Backend(16, primes, device=device)
    n, kl = 1 << 16, len(primes)
    boot = pkg.Bootstrapper(be, total_limbs=35)
    g = torch.Generator(device=torch.device("cuda", device))
    g.manual_seed(seed)

    def rand_key(levels=kl - 1):
        """uniform residues in a key's layout: SEAL's [kl-1, 2, kl, n], or truncated [L, 2, L + 1, n]"""
        ids = list(range(levels)) + [kl - 1]
        k = torch.empty((levels, 2, levels + 1, n), dtype=torch.int64, device=torch.device("cuda", device))
        for pos, l in enumerate(ids):
            k[:, :, pos, :] = torch.randint(0, primes[l], (levels, 2, n), generator=g, device=torch.device("cuda", device), dtype=torch.int64)
        return k

    if args.mode == "exact":
        steps = set(boot.required_steps())
        # of the driver's default power-of-two Galois keys only the multiples of num_batch = 256 are ever
        # used (QK^T rotates by i*256, softmax*V by multiples of 256): 2^8 .. 2^14 and their negatives
        att = set()
        for k in range(8, 15):
            att |= {1 << k, (n // 2) - (1 << k)}
        gal = {be.galois_elt_from_step(st): rand_key() for st in sorted(steps) + [0]}
        for st in sorted(att):          # used at <= 14 limbs only: SEAL-exact keys truncated to 14 levels
            e = be.galois_elt_from_step(st)
            if e not in gal:
                gal[e] = rand_key(14)
        keys = be.make_keys(relin=rand_key(), galois=gal)
        n_keys = len(gal)
    elif args.mode == "grouped":
        # fast mode with grouped-digit keys (csrc/ksgroup.hpp): every Galois key in the digit layout the cost model
        # prefers at the level it is used at, the relinearisation key in one variant per layout
        boot.set_hoisting(int(os.environ.get("MOAI_HOISTING", "2")))   # 2: lazy mod-down + single-digit first stage
        dev = torch.device("cuda", device)
        gal, grouped, single = {}, {}, {}

        def add(st, level):
            e = be.galois_elt_from_step(st)
            if level == 0:      # baby step of the first CoeffToSlot stage: single-digit key [1, 2, kl, n]
                single[e] = rand_key()[:1].clone()
                return
            k = be.ksg_best_extra(level)
            if k == 0:
                gal.setdefault(e, []).append(rand_key(level))
            else:
                grouped.setdefault(e, []).append(be.random_grouped_key(k, level, g))

        for st, lvs in sorted(boot.required_step_levels().items()):
            for lv in lvs:
                add(st, lv)
        att = pkg.attention_rotation_steps(256)
        for tag, level in (("qk", 14), ("sv", 3)):
            for st in att[tag]:
                add(st, level)
        grouped[0] = [be.random_grouped_key(k, lv, g) for k, lv in sorted(be.ksg_plan(range(1, kl - 1)).items())]
        keys = be.make_keys(relin=rand_key(), galois_fast=gal, grouped=grouped, single=single)
        n_keys = sum(len(v) for v in gal.values()) + sum(len(v) for v in grouped.values()) + len(single)
    else:
        boot.set_hoisting(True)
        gal = {}
        for st in boot.required_steps() + [0]:
            gal.setdefault(be.galois_elt_from_step(st), []).append(rand_key())
        att = pkg.attention_rotation_steps(256)
        for tag, level in (("qk", 14), ("sv", 3)):     # keys truncated to the level they are used at
            for st in att[tag]:
                gal.setdefault(be.galois_elt_from_step(st), []).append(rand_key(level))
        keys = be.make_keys(relin=rand_key(), galois_fast=gal)
        n_keys = sum(len(v) for v in gal.values())
    key_gib = torch.cuda.memory_allocated() / 2 ** 30
    hidden, heads, hd, inter = 768, 12, 64, 3072
    rng = np.random.default_rng(20250991)
    w = {"hidden": hidden, "heads": heads, "head_dim": hd, "inter": inter,
         "WQ": rng.normal(size=(heads, hidden, hd)) * 0.04, "WK": rng.normal(size=(heads, hidden, hd)) * 0.04,
         "WV": rng.normal(size=(heads, hidden, hd)) * 0.04, "bQ": rng.normal(size=(heads, hd)) * 0.04,
         "bK": rng.normal(size=(heads, hd)) * 0.04, "bV": rng.normal(size=(heads, hd)) * 0.04,
         "selfoutput": rng.normal(size=(hidden, hidden)) * 0.04, "selfoutput_bias": rng.normal(size=hidden) * 0.04,
         "ln1_gamma": np.ones(hidden), "ln1_beta": np.zeros(hidden),
         "inter_weight": rng.normal(size=(hidden, inter)) * 0.04, "inter_bias": rng.normal(size=inter) * 0.04,
         "final_weight": rng.normal(size=(inter, hidden)) * 0.04, "final_bias": rng.normal(size=hidden) * 0.04,
         "ln2_gamma": np.ones(hidden), "ln2_beta": np.zeros(hidden)}
    x = torch.empty((hidden, 2, 21, n), dtype=torch.int64, device=torch.device("cuda", device))
    for l in range(21):
        x[:, :, l, :] = torch.randint(0, primes[l], (hidden, 2, n), generator=g, device=torch.device("cuda", device), dtype=torch.int64)
    mask = np.ones(n // 2, dtype=np.int32)     # all 128 tokens of all 256 inputs valid
    return {"be": be, "boot": boot, "keys": keys, "w": w, "x": x, "mask": mask, "key_gib": key_gib, "n_keys": n_keys,
            "hidden": hidden, "n": n}


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=1,
                    help="encoder layers run back to back on the same packed batch (BASELINE config 5 = 12)")
    ap.add_argument("--mode", default="grouped", choices=["grouped", "fast", "exact"],
                    help="grouped: fast mode with grouped-digit keys (default); fast: hoisted rotations, pre-permuted "
                         "level-truncated keys; exact: SEAL-identical key switches")
    args = ap.parse_args()
    st = setup(args)
    be, boot, keys, w, x, mask = st["be"], st["boot"], st["keys"], st["w"], st["x"], st["mask"]
    key_gib, n_keys, hidden, n = st["key_gib"], st["n_keys"], st["hidden"], st["n"]
    boot_chunk = int(os.environ.get("MOAI_BOOT_CHUNK", "32"))
    be.profile(True)
    sampler = bench.ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    layer_ms = []
    try:
        for layer_id in range(args.layers):     # each layer's output (chain_index 20) is the next layer's input
            el0, el1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            el0.record()
            out, out_scale = boot.encoder_layer(keys, x, 2.0 ** 46, w, mask, 128, 256, layer_id=layer_id % 12,
                                                boot_chunk=boot_chunk, inplace=True)
            el1.record()
            torch.cuda.synchronize()
            layer_ms.append(el0.elapsed_time(el1))
            x = out
    except Exception as exc:      # report how far the run got
        print(json.dumps({"error": str(exc), "stages_done_ms": {k: round(v[0], 1) for k, v in be.profile_dump().items()},
                          "layers_done_ms": layer_ms, "evaluation_keys_GiB": round(key_gib, 1)}))
        raise
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    stages = be.profile_dump()
    be.profile(False)
    clocks = sampler.stop()
    rows = {}
    for k, (v, cnt) in stages.items():
        if k == "ctpt_gemm":
            continue
        v = v / args.layers
        rows[k] = {"ms": round(v, 1), "s_per_input_12_layers": round(v / 1000.0 * 12 / 256, 3),
                   "reference_s_per_input_12_layers": REFERENCE_TABLE3.get(k)}
    total = ms / 1000.0 / args.layers
    print(json.dumps({"workload": "C4/C5: %d BERT-base encoder layer(s) back to back, 256 inputs x 128 tokens, N=65536, "
                                  "1 x B200" % args.layers,
                      "mode": args.mode, "evaluation_keys_GiB": round(key_gib, 1), "layers": args.layers,
                      "seconds_per_layer": [round(v / 1000.0, 2) for v in layer_ms],
                      "layer_seconds": round(total, 2), "host_wall_seconds": round(wall, 2),
                      "amortized_s_per_input_12_layers": round(total * 12 / 256, 3),
                      "reference_s_per_input_12_layers": 574.6, "galois_keys": n_keys, "clocks": clocks,
                      "gpu_mem_GiB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                      "launches": be.launch_count(), "stages": rows}))
    assert out.shape == (hidden, 2, 21, n) and out_scale == 2.0 ** 46
    be.close()


if __name__ == "__main__":
    main()
