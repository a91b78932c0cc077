"""Row-sharded decode program list on two GPUs (b200q4_chain_create_sharded): the gathered output vectors on EVERY rank equal a
single-GPU b200q4_mul_mat_dev on the unsharded tensor bit for bit -- no reduction crosses GPUs (imax.c:1185-1188 splits by
output rows), the all-gather is the reducer warps' {value, tag} stores into every rank's pool."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, str(ROOT))
    ok, msg = True, ""
    try:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
        import vsim_b200 as vs
        from vsim_b200.shard import ShardedChain, shard_rows
        vs.init(rank)
        L = vs.load_library()
        L.b200q4_set_exact(0)
        rng = np.random.default_rng(77)                    # the same on every rank
        shapes = [(4096, 4096), (2048, 1024), (8192, 4096)]
        fulls, shards, xs = [], [], []
        for n, k in shapes:
            raw = rng.integers(0, 256, size=(n, (k // 32) * 20), dtype=np.uint8)
            raw.reshape(n, k // 32, 20)[:, :, :4] = np.full((n, k // 32), 0.01, np.float32).view(np.uint8).reshape(n, k // 32, 4)
            r0, nr = shard_rows(n, world, rank)
            fulls.append(vs.Q4Weight(0, raw, n, k))
            shards.append(vs.Q4Weight(0, raw, n, k, row0=r0, nrows=nr))
            xs.append(torch.from_numpy(rng.standard_normal(k).astype(np.float32)).cuda())
        ch = ShardedChain(L, shards, [x.data_ptr() for x in xs])
        for rep in range(3):
            ch.run()
            torch.cuda.synchronize()
            dist.barrier()                                 # every rank's stores have landed
            for i, (n, k) in enumerate(shapes):
                y1 = torch.empty(n, device="cuda")
                assert L.b200q4_mul_mat_dev(fulls[i].handle, xs[i].data_ptr(), 1, 0, y1.data_ptr(), 0, None) == 0
                L.b200q4_sync(None)
                got = ch.gathered(i)
                if not torch.equal(got.view(torch.int32), y1.view(torch.int32)):
                    ok, msg = False, f"rank {rank} rep {rep} op {i}: max diff {float((got - y1).abs().max())}"
            dist.barrier()
        ch.free()
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:                                # noqa: BLE001
        ok, msg = False, f"rank {rank}: {e!r}"
    q.put((rank, ok, msg))


@pytest.mark.gpu
def test_sharded_chain_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
