"""Development aid: the reference's ggml_mul_mat + ggml_graph_compute with the hook served by the GPU (libggml_b200.so) at
full GPT-J shapes, several thread counts, vs the oracle."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import Oracle, REF_DIR
orc = Oracle()
L = C.CDLL(str(REF_DIR / "libggml_b200.so"))
L.ref_mul_mat.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
L.ref_mul_mat.restype = C.c_double
rng = np.random.default_rng(0)
for n, k in ((4096, 4096), (16384, 4096), (4096, 16384), (50400, 4096)):
    raw = rng.integers(0, 256, size=(n, k // 32, 20), dtype=np.uint8)
    raw[:, :, :4] = np.frombuffer(np.float32(1.0 / (4.6 * np.sqrt(k))).tobytes(), np.uint8)
    wq = np.ascontiguousarray(raw.reshape(n, -1))
    for m in (1, 5):
        x = rng.standard_normal((m, k)).astype(np.float32)
        yr = orc.mul_mat(0, wq, n, k, x, nth=8)
        for nth in (1, 4, 16):
            y = np.empty((m, n), np.float32)
            rc = L.ref_mul_mat(0, wq.ctypes.data, n, k, x.ctypes.data, m, y.ctypes.data, nth, 1)
            print(f"N={n} K={k} M={m} nth={nth}: rc={rc:.3g} max err / max|y| = {np.abs(y - yr).max() / np.abs(yr).max():.3e}", flush=True)
