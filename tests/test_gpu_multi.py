"""Fused all-gather epilogue (b200q4_mul_mat_dev_allgather) on >= 2 B200s of one box: every rank ends up
with the full Y, bit-identical to the single-GPU result.  Skipped on a single-GPU box
(run with: gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok, msg = True, ""
    try:
        import vsim_b200 as vs
        from oracle import Oracle
        from vsim_b200.shard import PeerOutput, shard_rows
        vs.init(rank)
        L = vs.load_library()
        orc = Oracle()
        po = PeerOutput(L, 64, 1024)
        # K = 2048: one K-slice per tile (the GEMV epilogue stores to the peers itself);
        # K = 8192: K-sliced launches (the finalize kernel adds the parked pieces, stores to the peers and signals)
        for n, k in ((1024, 2048), (512, 8192)):
            for qt in (0, 1):
                rng = np.random.default_rng(11 + qt + k)
                wq = orc.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
                full = vs.Q4Weight(qt, wq, n, k)
                r0, nr = shard_rows(n, world, rank)
                shard = vs.Q4Weight(qt, wq, n, k, row0=r0, nrows=nr)
                for m in (1, 2, 5, 8, 20, 64):               # decode kernel epilogue (<= 8) and tile kernel + scatter
                    x = rng.standard_normal((m, k)).astype(np.float32)
                    d_x = torch.from_numpy(x).cuda()
                    for rep in range(3):                      # both ping-pong buffers, repeated sequence numbers
                        y = po.mul_mat(shard, d_x, m, n_total=n).clone()
                    torch.cuda.synchronize()
                    y_one = vs.mul_mat(full, x)               # single-GPU result of the same call
                    got = y.cpu().numpy()
                    if not np.array_equal(got.view(np.uint32), y_one.view(np.uint32)):
                        # slices of a shard are summed over differently placed stream-K pieces (tile kernel): rounding only
                        if m <= 8 or np.abs(got - y_one).max() > 6e-6 * np.abs(y_one).max():
                            ok, msg = False, f"rank {rank} qt {qt} K {k} M {m}: max diff {np.abs(got - y_one).max()}"
                full.free(); shard.free()
        dist.barrier()
    except Exception as e:                                # noqa: BLE001
        ok, msg = False, f"rank {rank}: {e!r}"
    finally:
        q.put((rank, ok, msg))
        dist.destroy_process_group()


def test_fused_all_gather_two_or_more_gpus():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), [m for _, ok, m in res if not ok]
