"""Host logic of the row-sharded path (vsim_b200/shard.py) with world_size-2 and -4 gloo groups on CPU:
shard ranges + slice assembly reproduce the full mul_mat bit for bit (each rank's slice computed by the
oracle -- test infrastructure -- exactly as its GPU would)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import Oracle, Q4_0, Q4_1
from vsim_b200.shard import all_gather_rows, assemble, shard_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, qt, n, k, m, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = Oracle()
        rng = np.random.default_rng(7)                      # every rank: the same tensor and activations
        wq = orc.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
        x = rng.standard_normal((m, k)).astype(np.float32)
        r0, nr = shard_rows(n, world, rank)
        y_loc = orc.mul_mat(qt, wq[r0:r0 + nr], nr, k, x)   # this rank's output columns
        y = all_gather_rows(torch.from_numpy(np.ascontiguousarray(y_loc))).numpy()
        y_full = orc.mul_mat(qt, wq, n, k, x)
        q.put((rank, bool(np.array_equal(y.view(np.uint32), y_full.view(np.uint32)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("qt", [Q4_0, Q4_1])
def test_sharded_all_gather_equals_full(world, qt):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, qt, 64, 256, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res)


def test_shard_rows_and_assemble():
    assert shard_rows(50400, 8, 3) == (3 * 6300, 6300)
    assert shard_rows(16, 1, 0) == (0, 16)
    with pytest.raises(ValueError):
        shard_rows(50257, 8, 0)                              # GPT-2's vocab does not shard: replicas only
    with pytest.raises(ValueError):
        shard_rows(64, 4, 4)
    s = torch.arange(2 * 3 * 4, dtype=torch.float32).view(2, 3, 4)     # [world][M][N/world]
    y = assemble(s)
    assert y.shape == (3, 8)
    assert torch.equal(y[1], torch.cat([s[0, 1], s[1, 1]]))
