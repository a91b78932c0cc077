"""Development aid: per-step clock stamps of every warp role of q4_tile_kernel (CTA 0, first 32 chunk steps).
Needs a -DTK_TRACE build:
    B200Q4_DEFS=TK_TRACE B200Q4_LIB_OUT=vsim_b200/libb200q4_trace.so python -m vsim_b200.build --force
    B200Q4_LIB=vsim_b200/libb200q4_trace.so python tools/tile_trace.py <qt> <M>
"""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np, torch
os.environ["B200Q4_DBG"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from vsim_b200 import lib as vlib
qt = int(sys.argv[1]) if len(sys.argv) > 1 else 0
m = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n, k = 16384, 4096
vsim_b200.init(0)
L = vsim_b200.load_library()
rng = np.random.default_rng(0)
ws = []
for i in range(10):
    raw = rng.integers(0, 256, size=(n, vlib.row_bytes(qt, k)), dtype=np.uint8)
    ws.append(vsim_b200.Q4Weight(qt, raw, n, k))
arr = (C.c_void_p * 10)(*[w.handle for w in ws])
x = torch.randn(m, k, device="cuda"); y = torch.empty(m, n, device="cuda")
torch.cuda.synchronize()
t = L.b200q4_time_mul_mat_dev(arr, 10, x.data_ptr(), m, y.data_ptr(), 20)
print(f"qt={qt} M={m}: {t*1e3:.2f} us/call")
buf = np.zeros((256, 16), np.uint64)
L.b200q4_debug_timeline.argtypes = [C.c_void_p]
assert L.b200q4_debug_timeline(buf.ctypes.data) == 0
b = buf.reshape(-1)[: 11 * 32].reshape(11, 32).astype(np.int64)
t0 = b[8, 0]
print("kernel-level (cycles after the activation producer started): kernel entry %d, prologue done %d, pack kernel done (act producer) %d, "
      "epilogue past the dependency sync %d, raw producer at the final barrier %d, epilogue warp 0: last piece accumulated %d, "
      "its store / stream-K combine done %d" % tuple(int(b[8, k] - t0) for k in (7, 8, 1, 2, 3, 5, 6)))
hdr = ["raw issued", "conv: raw_full", "conv: conv_empty", "conv: done", "act issued", "mma: conv_full", "mma: act_full", "mma: committed", "epi: tm_full", "epi: chunk read"]
rows = [0, 1, 2, 3, 9, 10, 4, 5, 6, 7]
print("step " + " ".join(f"{h:>17s}" for h in hdr))
for s in range(32):
    if b[0, s] == 0 and s > 0:
        break
    print(f"{s:4d} " + " ".join(f"{(b[r, s] - t0) if b[r, s] else -1:17d}" for r in rows))
