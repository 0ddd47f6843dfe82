"""Multi-GPU checks; they need at least two visible GPUs and are skipped (with that reason) on a one-GPU box.
Run them with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_partitioning_a_is_bit_identical_to_one_gpu():
    """One packed batch over 2 GPUs (heads / columns / bootstrapping pairs sharded, NCCL all-gathers of limbs,
    moai_comm_*): the layer output equals the one-GPU output bit for bit on every rank (tools/partition_check.py)."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "partition_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    print(line)
    assert res["bit_identical_on_every_rank"] and res["all_gathers"] >= 6


def test_two_devices_in_one_process(pkg):
    """Two contexts on two devices in ONE process (moai_context_create takes a device): the arenas are keyed by
    (device, stream), so a block cached on device 0 is never handed to device 1 — same op, same residues on both."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch
    from oracle import Oracle
    o = Oracle(12, [40, 30, 30, 40])
    rng = np.random.default_rng(0)
    a = np.empty((4, 2, 3, o.n), dtype=np.uint64)
    for l in range(3):
        a[:, :, l, :] = rng.integers(0, int(o.q[l]), (4, 2, o.n), dtype=np.uint64)
    outs = []
    for rounds in range(2):                      # second round reuses the blocks the first one cached
        for dev in (0, 1):
            torch.cuda.set_device(dev)
            be = pkg.Backend(12, o.q, device=dev)
            x = pkg.to_device(a, device="cuda:%d" % dev)
            y = be.rescale_to_next(be.add(x, x))
            assert y.device.index == dev
            outs.append(pkg.to_host(y))
            del x, y
            be.close()
    for r in outs[1:]:
        assert np.array_equal(r, outs[0])
    torch.cuda.set_device(0)
