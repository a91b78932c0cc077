"""Development aid: offline weight quantizer (row f-4) -- device (incl. PCIe both ways) vs the oracle's scalar CPU port."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200 as vs
from oracle import quantize_weights
vs.init(0)
rng = np.random.default_rng(0)
n, k = 50400, 4096
w = (rng.random((n, k), dtype=np.float32) - np.float32(0.5)) * np.float32(0.1)
for qt in (0, 1):
    vs.quantize_weights(qt, w[:64])
    t0 = time.perf_counter(); rows, hist = vs.quantize_weights(qt, w); t1 = time.perf_counter()
    c0 = time.perf_counter(); want, whist = quantize_weights(qt, w[:4096]); c1 = time.perf_counter()
    ok = np.array_equal(rows[:4096], want)
    print(f"q4_{qt}: {n}x{k} ({w.nbytes/1e9:.2f} GB fp32) device {1e3*(t1-t0):.0f} ms = {w.nbytes/(t1-t0)/1e9:.1f} GB/s of fp32 in; "
          f"CPU port {1e3*(c1-c0)*n/4096:.0f} ms (scaled from 4096 rows, 1 thread); first 4096 rows identical: {ok}; hist sum {hist.sum()}")
