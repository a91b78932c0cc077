"""Development aid: b200q4_mul_mat through the C ABI with pinned host buffers (bench.py's e2e leg), us per call."""
import ctypes as C, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200 as vs
from vsim_b200 import lib as vlib
n, k, m = 16384, 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 1
vs.init(0); L = vs.load_library()
rng = np.random.default_rng(0)
ws = []
for i in range(10):
    raw = rng.integers(0, 256, size=(n, k // 32, 20), dtype=np.uint8)
    raw[:, :, :4] = np.frombuffer(np.float32(0.01).tobytes(), np.uint8)
    ws.append(vs.Q4Weight(0, raw.reshape(n, -1), n, k))
hx = L.b200q4_host_alloc(m * k * 4); hy = L.b200q4_host_alloc(m * n * 4)
x = rng.standard_normal((m, k)).astype(np.float32); C.memmove(hx, x.ctypes.data, x.nbytes)
for i in range(50): L.b200q4_mul_mat(ws[i % 10].handle, hx, m, k * 4, hy, n * 4)
t0 = time.perf_counter()
N = 2000
for i in range(N): L.b200q4_mul_mat(ws[i % 10].handle, hx, m, k * 4, hy, n * 4)
dt = time.perf_counter() - t0
print(f"M={m}: {dt / N * 1e6:.2f} us/call")
