"""Row-sharded decode program on >= 2 B200s: every rank's logits equal the single-GPU program's bit for bit (no reduction
crosses GPUs), for several steps incl. the chunked-attention regime.  Skipped on a single-GPU box
(run with: gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi_neox.py -m gpu)."""
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


def _worker(rank, world, port, path, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok, msg = True, ""
    try:
        import vsim_b200 as vs
        from vsim_b200.shard import ShardedNeox
        vs.init(rank)
        L = vs.load_library()
        L.b200q4_set_exact(0)
        one = L.b200neox_load(str(path).encode(), 160, 8)          # the single-GPU program on this rank's GPU
        assert one and L.b200neox_decode_is_program(one) == 1, L.b200neox_last_error()
        sh = ShardedNeox(L, str(path), n_ctx=160)
        nv = sh.n_vocab
        ref = np.empty(nv, np.float32)
        tok = 3
        for n_past in range(0, 80):                                # past one attention chunk
            t = np.array([tok], np.int32)
            assert L.b200neox_eval(one, t.ctypes.data, 1, n_past, 1, ref.ctypes.data) == 0, L.b200neox_last_error()
            got = sh.eval(tok, n_past)
            if not np.array_equal(got.view(np.uint32), ref.view(np.uint32)):
                ok, msg = False, f"rank {rank} step {n_past}: max diff {np.abs(got - ref).max()} of {np.abs(ref).max()}"
                break
            tok = int(ref.argmax())
        sh.free()
        L.b200neox_free(one)
        dist.barrier()
    except Exception as e:                                # noqa: BLE001
        ok, msg = False, f"rank {rank}: {e!r}"
    finally:
        q.put((rank, ok, msg))
        dist.destroy_process_group()


@pytest.mark.parametrize("par_res", [1, 0])
def test_sharded_program_equals_single_gpu(tmp_path, par_res):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import Q4_0
    from tests.helpers.neox_file import write_neox
    path = tmp_path / "m.bin"
    write_neox(path, Q4_0, n_vocab=256, n_embd=256, n_head=4, n_layer=3, n_rot=32, seed=5, par_res=par_res)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), [m for _, ok, m in res if not ok]
