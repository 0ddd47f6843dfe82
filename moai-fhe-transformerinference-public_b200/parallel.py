"""Multi-GPU host logic (one process per GPU, torch.distributed).

The reference has no distributed code: its only parallelism is OpenMP over independent ciphertexts
(SURVEY §2.3).  The 256 inputs of a packed batch share every ciphertext, so the independent units
are (a) whole packed batches ("replicas", weak scaling, no data-path collective) and (b) ciphertext
indices — output columns of a ct-pt matmul, heads, bootstrappings — whose results are gathered with
one all-gather of raw uint64 limbs (no reduction op: sums are modular).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced split of `total` independent units; rank r gets [begin, end)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_over_ranks(value, device="cpu"):
    """Device-timed milliseconds -> max over ranks (the number bench.py reports)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_columns(local, total_cols):
    """All-gather the column shards of a ciphertext vector.  local: [cols_r, 2, limbs, n] int64
    (uint64 bits); returns [total_cols, 2, limbs, n] on every rank.  Shards may be ragged."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total_cols, r, world) for r in range(world)]
    widest = max(e - b for b, e in sizes)
    pad = torch.zeros((widest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * widest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * widest: r * widest + (e - b)] for r, (b, e) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def amortized_seconds_per_input(ms_per_step, inputs_per_rank, world, replicas=True):
    """bench.py's metric: replicas process world * inputs_per_rank inputs per step; a sharded
    batch processes inputs_per_rank inputs with all ranks cooperating."""
    inputs = inputs_per_rank * world if replicas else inputs_per_rank
    return (ms_per_step / 1000.0) / inputs
