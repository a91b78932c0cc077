"""Development aid: what a launch of the decode program kernel costs beyond its steps -- lists of 1 ... 40 independent headline
steps (Q4_0, K=4096, N=16384, own input rows), one launch on an idle GPU and ten back to back (b200q4_chain_time).
B200Q4_PROG_NOCOOP=1 launches without the cooperative attribute (no difference measured)."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import vsim_b200
from tools.chain_bench import rand_q4
vsim_b200.init(0)
L = vsim_b200.load_library()
rng = np.random.default_rng(1)
n, k = 16384, 4096
wl = [vsim_b200.Q4Weight(0, rand_q4(rng, 0, n, k), n, k) for _ in range(10)]
xin = torch.randn(64, k, device="cuda"); yout = torch.zeros(64, n, device="cuda")
st = torch.cuda.Stream()
for nops in (1, 2, 5, 10, 20, 40):
    ws = [wl[i % 10] for i in range(nops)]
    ch = vsim_b200.Chain(ws, [xin[i].data_ptr() for i in range(nops)], [yout[i].data_ptr() for i in range(nops)], [-1] * nops)
    ch.run(st.cuda_stream); torch.cuda.synchronize()
    cold = []
    for _ in range(5):
        torch.cuda.synchronize()
        cold.append(L.b200q4_chain_time(ch.handle, None, 1, st.cuda_stream) * 1e3)
    torch.cuda.synchronize()
    warm = L.b200q4_chain_time(ch.handle, None, 10, st.cuda_stream) * 1e3 / 10
    print(f"ops {nops:3d}: one launch on an idle GPU {np.median(cold):8.2f} us ({np.median(cold)/nops:6.2f} per op)   ten launches back to back {warm:8.2f} us per launch ({warm/nops:6.2f} per op)", flush=True)
