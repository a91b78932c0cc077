"""Development aid: randomized shape sweep of the tile kernel against the oracle (stream-K edge cases)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200 as gpu
from oracle import Oracle

gpu.init(0)
orc = Oracle()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
shapes = [(512, 2048, 64), (1024, 2048, 64), (512, 2048, 20), (1024, 2048, 20)]
for _ in range(60):
    n = int(rng.choice([16, 100, 128, 200, 512, 640, 1024, 2000, 4096]))
    k = int(rng.choice([32, 128, 512, 1024, 2048, 4096, 8192]))
    m = int(rng.choice([9, 16, 20, 33, 64, 65, 100, 128, 130, 256]))
    if n * k * m < 6e9:
        shapes.append((n, k, m))
bad_total = 0
for n, k, m in shapes:
    for qt in (0, 1):
        wq = orc.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
        x = rng.standard_normal((m, k)).astype(np.float32)
        w = gpu.Q4Weight(qt, wq, n, k)
        y = gpu.mul_mat(w, x)
        y2 = gpu.mul_mat(w, x)
        yr = orc.mul_mat(qt, wq, n, k, x, nth=8)
        err = np.abs(y - yr)
        bad = int((err > 1e-3 * np.abs(yr) + 1e-5 * np.abs(yr).max(axis=1, keepdims=True)).sum())
        rep = np.array_equal(y.view(np.uint32), y2.view(np.uint32))
        if bad or not rep:
            bad_total += 1
            print(f"FAIL qt={qt} N={n} K={k} M={m}: bad={bad}/{err.size} max={err.max():.3g} repeatable={rep}", flush=True)
        w.free()
print("shapes", len(shapes), "failures", bad_total)
