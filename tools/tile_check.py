"""Development aid: run the tcgen05 tile kernel over a list of shapes and print its error against the
oracle (no asserts -- one GPU round trip should show every failing shape).  Run on a GPU box:
    timeout -s KILL 300 python tools/tile_check.py [quick]
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200 as gpu                      # noqa: E402
from oracle import Oracle, Q4_0, Q4_1        # noqa: E402

SHAPES = [(128, 128, 16), (128, 128, 3), (16, 64, 5), (200, 256, 9), (333, 2048, 5), (640, 1024, 16), (300, 512, 17),
          (128, 256, 40), (1000, 4096, 64), (777, 1056, 100), (512, 4096, 128), (384, 1024, 300), (4096, 4096, 8),
          (16384, 4096, 8), (2048, 16384, 4)]


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    gpu.init(0)
    orc = Oracle()
    print(gpu.device_info(), flush=True)
    for qt, tag in ((Q4_0, "q4_0"), (Q4_1, "q4_1")):
        for n, k, m in (SHAPES[:4] if quick else SHAPES):
            rng = np.random.default_rng(n * 131 + k * 7 + m)
            wq = orc.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
            x = rng.standard_normal((m, k)).astype(np.float32)
            w = gpu.Q4Weight(qt, wq, n, k)
            t0 = time.time()
            y = gpu.mul_mat(w, x)
            dt = time.time() - t0
            rows = np.arange(n) if n * k * m < 3e9 else np.sort(rng.choice(n, 256, replace=False))
            yr = orc.mul_mat(qt, wq[rows], len(rows), k, x, nth=8)
            err = np.abs(y[:, rows].astype(np.float64) - yr)
            rel = np.linalg.norm(err) / (np.linalg.norm(yr) + 1e-300)
            bad = (err > 1e-3 * np.abs(yr) + 1e-5 * np.abs(yr).max(axis=1, keepdims=True)).sum()
            print(f"{tag} N={n} K={k} M={m} kernel={gpu.kernel_for(qt, m, n, k)} rel={rel:.3e} bad={bad}/{err.size} "
                  f"nan={np.isnan(y).sum()} t={dt * 1e3:.1f}ms", flush=True)
            if bad and n <= 1024:
                i = np.unravel_index(err.argmax(), err.shape)
                print("   worst at (m,n)=", i, "got", y[:, rows][i], "want", yr[i], flush=True)
                br = np.where(err.max(axis=0) > 1e-3 * np.abs(yr).max())[0]
                bm = np.where(err.max(axis=1) > 1e-3 * np.abs(yr).max())[0]
                print("   bad n:", br[:24], "... count", len(br), " bad m:", bm[:24], "count", len(bm), flush=True)
            w.free()


if __name__ == "__main__":
    main()
