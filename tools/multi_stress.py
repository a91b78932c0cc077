"""Development aid: stress the fused all-gather on N ranks (run under torchrun): many calls, random M, check every result."""
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
n, k, qt = 1024, 2048, 0
po = PeerOutput(L, 64, n)
rng = np.random.default_rng(5)
wq = orc.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
full = vs.Q4Weight(qt, wq, n, k)
r0, nr = shard_rows(n, world, rank)
shard = vs.Q4Weight(qt, wq, n, k, row0=r0, nrows=nr)
ms = [1, 8, 20, 33, 64, 40]
xs = {m: rng.standard_normal((m, k)).astype(np.float32) for m in ms}
want = {m: torch.from_numpy(vs.mul_mat(full, xs[m])).cuda() for m in ms}
dxs = {m: torch.from_numpy(xs[m]).cuda() for m in ms}
torch.cuda.synchronize(); dist.barrier()
order = np.random.default_rng(99).choice(ms, size=int(sys.argv[1]) if len(sys.argv) > 1 else 600)
bad = 0
errs = []
for it, m in enumerate(order):
    m = int(m)
    y = po.mul_mat(shard, dxs[m], m)
    d = (y - want[m]).abs().max()
    errs.append((it, m, d))
torch.cuda.synchronize()
for it, m, d in errs:
    v = float(d)
    if v > (0 if m <= 8 else 2e-5):
        bad += 1
        if bad <= 8:
            print(f"rank {rank}: call {it} M={m} max diff {v:.3g}", flush=True)
print(f"rank {rank}: {len(order)} calls, {bad} bad", flush=True)
dist.barrier(); dist.destroy_process_group()
