"""The bench headline's launch for a profiler: OPS independent Q4_0 M=1 mul_mats (K=4096, N=16384, 10 distinct weight tensors,
every op with its own fp32 input row) as persistent launches of q4_prog_kernel.
    ncu --set full --clock-control none --import-source on -k regex:q4_prog --launch-skip 2 -c 1 -o OUT python tools/prog_prof.py [OPS]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from tools.chain_bench import rand_q4

ops = int(sys.argv[1]) if len(sys.argv) > 1 else 100
vsim_b200.init(0)
L = vsim_b200.load_library()
rng = np.random.default_rng(1)
n, k, nw = 16384, 4096, 10
ws = [vsim_b200.Q4Weight(0, rand_q4(rng, 0, n, k), n, k) for _ in range(nw)]
x = torch.randn(ops, k, device="cuda")
y = torch.zeros(ops, n, device="cuda")
ch = vsim_b200.Chain([ws[i % nw] for i in range(ops)], [x[i].data_ptr() for i in range(ops)], [y[i].data_ptr() for i in range(ops)], [-1] * ops)
for _ in range(4):
    ch.run()
L.b200q4_sync(None)
print("ok", float(y.abs().sum()))
ch.free()
