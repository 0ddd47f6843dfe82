#!/usr/bin/env python
"""Partitioning A check: ONE packed batch over the ranks of a torchrun job.

Every rank builds the same synthetic encoder-layer problem (N = 4096 by default, the repo's 36-prime chain shape,
hidden 768, 12 heads, FFN 3072; uniformly random residues as ciphertexts and keys — bit-equality of two evaluation
orders does not need decryptable data), runs the layer once with the communicator (its share of the heads / columns /
bootstrapping pairs + NCCL all-gathers) and once alone, and requires the two outputs to be BIT-IDENTICAL on every rank.
Prints one JSON line from rank 0 with the timings and the gathered volume.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
      tools/partition_check.py [--log-n 12]"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=12)
    args = ap.parse_args()
    pkg = importlib.import_module("moai-fhe-transformerinference-public_b200")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    primes = bench.moai_primes()          # congruent to 1 mod 2^17: valid NTT primes for every N <= 65536
    log_n, n, kl = args.log_n, 1 << args.log_n, len(primes)
    be = pkg.Backend(log_n, primes, device=local_rank)
    boot = pkg.Bootstrapper(be, total_limbs=35)
    boot.set_hoisting(True)
    g = torch.Generator(device=dev)
    g.manual_seed(7)                      # the SAME problem on every rank

    def rand_key(levels=kl - 1):
        ids = list(range(levels)) + [kl - 1]
        k = torch.empty((levels, 2, levels + 1, n), dtype=torch.int64, device=dev)
        for pos, l in enumerate(ids):
            k[:, :, pos, :] = torch.randint(0, primes[l], (levels, 2, n), generator=g, device=dev, dtype=torch.int64)
        return k

    num_batch = (n // 2) // 128
    gal = {}
    for st in boot.required_steps() + [0]:
        gal.setdefault(be.galois_elt_from_step(st), []).append(rand_key())
    att = pkg.attention_rotation_steps(num_batch)
    for tag, level in (("qk", 14), ("sv", 3)):
        for st in att[tag]:
            gal.setdefault(be.galois_elt_from_step(st), []).append(rand_key(level))
    keys = be.make_keys(relin=rand_key(), galois_fast=gal)
    hidden, heads, hd, inter = 768, 12, 64, 3072
    rng = np.random.default_rng(3)
    w = {"hidden": hidden, "heads": heads, "head_dim": hd, "inter": inter,
         "WQ": rng.normal(size=(heads, hidden, hd)) * 0.04, "WK": rng.normal(size=(heads, hidden, hd)) * 0.04,
         "WV": rng.normal(size=(heads, hidden, hd)) * 0.04, "bQ": rng.normal(size=(heads, hd)) * 0.04,
         "bK": rng.normal(size=(heads, hd)) * 0.04, "bV": rng.normal(size=(heads, hd)) * 0.04,
         "selfoutput": rng.normal(size=(hidden, hidden)) * 0.04, "selfoutput_bias": rng.normal(size=hidden) * 0.04,
         "ln1_gamma": np.ones(hidden), "ln1_beta": np.zeros(hidden),
         "inter_weight": rng.normal(size=(hidden, inter)) * 0.04, "inter_bias": rng.normal(size=inter) * 0.04,
         "final_weight": rng.normal(size=(inter, hidden)) * 0.04, "final_bias": rng.normal(size=hidden) * 0.04,
         "ln2_gamma": np.ones(hidden), "ln2_beta": np.zeros(hidden)}
    x0 = torch.empty((hidden, 2, 21, n), dtype=torch.int64, device=dev)
    for l in range(21):
        x0[:, :, l, :] = torch.randint(0, primes[l], (hidden, 2, n), generator=g, device=dev, dtype=torch.int64)
    mask = np.ones(n // 2, dtype=np.int32)
    cw, keep = boot.layer_weights(w)

    def layer(x):
        aux = torch.empty_like(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for stage in range(4):
            boot.encoder_layer_stage(keys, stage, x, aux, 2.0 ** 46, cw, mask, 128, num_batch, layer_id=0, boot_chunk=32)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    be.comm_init(dist)
    xs = x0.clone()
    ms_sharded = layer(xs)
    gathers, rx = be.comm_stats()
    be._chk(be.lib.moai_comm_destroy(be.h))
    xa = x0.clone()
    ms_alone = layer(xa)
    same = bool(torch.equal(xs, xa))
    flags = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    t = torch.tensor([ms_sharded, ms_alone], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "partitioning A: one packed batch over %d GPUs vs one GPU" % world, "log_n": log_n,
                          "bit_identical_on_every_rank": bool(flags.item()), "layer_ms_sharded": float(t[0]),
                          "layer_ms_alone": float(t[1]), "speedup": float(t[1] / t[0]), "all_gathers": gathers,
                          "received_GiB_per_rank": rx / 2 ** 30}))
    be.close()
    dist.destroy_process_group()
    if not flags.item():
        sys.exit(1)


if __name__ == "__main__":
    main()
