"""Development aid: full-size GPT-J shapes with random-nibble weights (code 0 included) vs the oracle, M = 1 and 5."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200 as vs
from oracle import Oracle
vs.init(0); orc = Oracle()
rng = np.random.default_rng(0)
for n, k in ((4096, 4096), (16384, 4096), (4096, 16384), (50400, 4096)):
    raw = rng.integers(0, 256, size=(n, k // 32, 20), dtype=np.uint8)
    raw[:, :, :4] = np.frombuffer(np.float32(1.0 / (4.6 * np.sqrt(k))).tobytes(), np.uint8)
    wq = raw.reshape(n, -1)
    w = vs.Q4Weight(0, wq, n, k)
    for m in (1, 5):
        x = rng.standard_normal((m, k)).astype(np.float32)
        y = vs.mul_mat(w, x)
        yr = orc.mul_mat(0, wq, n, k, x, nth=8)
        err = np.abs(y - yr).max() / np.abs(yr).max()
        y2 = vs.mul_mat(w, x)
        print(f"N={n} K={k} M={m}: max err / max|y| = {err:.3e}  repeat-identical {np.array_equal(y, y2)}", flush=True)
    w.free()
