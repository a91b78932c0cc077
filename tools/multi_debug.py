"""Development aid: 2-rank fused all-gather debug (run under torchrun)."""
import os, sys
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import vsim_b200 as vs
from oracle import Oracle
from vsim_b200.shard import PeerOutput, shard_rows
vs.init(rank); L = vs.load_library(); orc = Oracle()
n, k = 1024, 2048
po = PeerOutput(L, 64, n)
for qt in (0, 1):
    rng = np.random.default_rng(11 + qt)
    wq = orc.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    full = vs.Q4Weight(qt, wq, n, k)
    r0, nr = shard_rows(n, world, rank)
    shard = vs.Q4Weight(qt, wq, n, k, row0=r0, nrows=nr)
    for m in (1, 8, 20, 33, 64, 40, 64):
        x = rng.standard_normal((m, k)).astype(np.float32)
        d_x = torch.from_numpy(x).cuda()
        for rep in range(3):
            y = po.mul_mat(shard, d_x, m).clone()
        torch.cuda.synchronize()
        y_one = vs.mul_mat(full, x)
        got = y.cpu().numpy()
        d = np.abs(got - y_one)
        half = n // world
        print(f"rank {rank} qt {qt} M {m}: max diff own-slice {d[:, r0:r0+nr].max():.3g} other {np.delete(d, np.s_[r0:r0+nr], axis=1).max():.3g} "
              f"bad rows {np.where(d.max(axis=1) > 1e-3)[0][:10]}", flush=True)
    full.free(); shard.free()
dist.barrier(); dist.destroy_process_group()
