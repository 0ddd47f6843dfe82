"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: unit sharding is a disjoint cover,
the max-over-ranks reduction and the ragged all-gather of ciphertext column shards work."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_cols, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    par = importlib.import_module("moai-fhe-transformerinference-public_b200.parallel")
    b, e = par.shard_range(total_cols, rank, world)
    # each rank "computes" its columns: value = column index, shape [cols, 2, 1, 4]
    local = torch.arange(b, e, dtype=torch.int64).view(-1, 1, 1, 1).expand(-1, 2, 1, 4).contiguous()
    full = par.gather_columns(local, total_cols)
    ok = full.shape[0] == total_cols and bool((full[:, 0, 0, 0] == torch.arange(total_cols)).all())
    mx = par.max_over_ranks(10.0 + rank)
    q.put((rank, b, e, ok, mx))
    dist.destroy_process_group()


def test_shard_range_is_a_disjoint_cover():
    import importlib
    load_pkg()
    par = importlib.import_module("moai-fhe-transformerinference-public_b200.parallel")
    for total in (0, 1, 7, 768, 3084):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    assert par.amortized_seconds_per_input(1000.0, 256, 8, replicas=True) == 1.0 / 2048
    assert par.amortized_seconds_per_input(1000.0, 256, 8, replicas=False) == 1.0 / 256


def test_two_rank_gather_and_max():
    load_pkg()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert [r[1:3] for r in res] == [(0, 4), (4, 7)]          # ragged split 4 + 3
    assert all(r[3] for r in res)
    assert all(r[4] == 11.0 for r in res)
