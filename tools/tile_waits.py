"""Development aid: where each warp role of q4_tile_kernel waits.  Needs a -DB200Q4_TIMELINE build:
    B200Q4_TIMELINE=1 python -m vsim_b200.build --force && python tools/tile_waits.py <qt> <M>
"""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np, torch
os.environ["B200Q4_DBG"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from vsim_b200 import lib as vlib
qt = int(sys.argv[1]) if len(sys.argv) > 1 else 0
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n, k = 16384, 4096
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
t = L.b200q4_time_mul_mat_dev(arr, 10, x.data_ptr(), m, y.data_ptr(), 20)
print(f"qt={qt} M={m}: {t*1e3:.2f} us/call")
buf = np.zeros((256, 16), np.uint64)
L.b200q4_debug_timeline.argtypes = [C.c_void_p]
assert L.b200q4_debug_timeline(buf.ctypes.data) == 0
b = buf[:148].astype(np.float64)
names = ["raw producer TOTAL", "  wait raw_empty", "act producer TOTAL", "  wait act_empty", "MMA issuer TOTAL", "  wait conv_full",
         "  wait act_full", "  wait tm_empty", "converter TOTAL", "  wait raw_full", "  wait conv_empty", "epilogue TOTAL",
         "  wait conv_full", "  wait act_full", "  wait tm_full", "-"]
for i, nm in enumerate(names):
    print(f"{nm:22s} median {np.median(b[:, i]):10.0f} cycles   max {b[:, i].max():10.0f}")
