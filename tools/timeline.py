"""Development aid: per-CTA timeline of the decode kernel (needs B200Q4_DBG=1)."""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np, torch
os.environ["B200Q4_DBG"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from vsim_b200 import lib as vlib
qt = int(sys.argv[1]) if len(sys.argv) > 1 else 0
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
k = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
print(f"== timeline qt={qt} M={m} N={n} K={k}")
vsim_b200.init(0)
L = vsim_b200.load_library()
rng = np.random.default_rng(0)
ws = []
for i in range(10):
    raw = rng.integers(0, 256, size=(n, vlib.row_bytes(qt, k)), dtype=np.uint8)
    if qt == 0:
        raw.reshape(n, k // 32, 20)[:, :, :4] = np.full((n, k // 32), 0.01, np.float32).view(np.uint8).reshape(n, k // 32, 4)
    else:
        raw[:, :(k // 32) * 8] = np.full((n, k // 32 * 2), 0.01, np.float32).view(np.uint8).reshape(n, -1)
    ws.append(vsim_b200.Q4Weight(qt, raw, n, k))
arr = (C.c_void_p * 10)(*[w.handle for w in ws])
x = torch.randn(m, k, device="cuda"); y = torch.empty(m, n, device="cuda")
torch.cuda.synchronize()
L.b200q4_time_mul_mat_dev.restype = C.c_double
t = L.b200q4_time_mul_mat_dev(arr, 10, x.data_ptr(), m, y.data_ptr(), 50)
print(f"{t*1e3:.2f} us/call")
buf = np.zeros((256, 16), np.uint64)
L.b200q4_debug_timeline.argtypes = [C.c_void_p]
assert L.b200q4_debug_timeline(buf.ctypes.data) == 0
b = buf[:148].astype(np.int64)
t0 = b[:, 0].min()
names = {0: "kernel start", 1: "producer: static tasks issued", 2: "producer: past griddepsync", 3: "consumers: past griddepsync",
         4: "consumers: prologue done", 5: "consumer0: first stage landed", 10: "producer: last copy issued", 9: "consumer0: last stage landed (or stop)",
         6: "consumer0: loop done", 11: "reducer: last stage's partials seen", 12: "reducer: last stage summed",
         13: "reducer: stop marker seen", 7: "reducer done", 8: "producer done"}
for sidx, nm in names.items():
    v = (b[:, sidx] - t0) / 1e3
    print(f"{nm:34s} min {v.min():7.2f}  med {np.median(v):7.2f}  max {v.max():7.2f} us")
print("per-CTA span (reducer done - start): med %.2f max %.2f" % (np.median((b[:, 7] - b[:, 0]) / 1e3), ((b[:, 7] - b[:, 0]) / 1e3).max()))

r = buf.reshape(-1)[3000:3000 + 96].astype(np.int64).reshape(24, 4)
print("consumer warp 0 of CTA 0, per iteration: stage landed / iteration done / stages in the iteration")
c0 = b[0, 0]
for i in range(24):
    if r[i, 0] == 0: break
    print(f"  iter {i:2d}: {(r[i,0]-c0)/1e3:7.2f} {(r[i,1]-c0)/1e3:7.2f}   x{r[i,2]}")
pass
