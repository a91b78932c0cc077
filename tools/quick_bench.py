"""Quick device-side timing of the microbench points (development aid, not the bench contract)."""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from vsim_b200 import lib as vlib

PEAK = 6454.6
try:
    PEAK = json.load(open(Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


def main():
    shapes = [(16384, 4096)]
    if "--all" in sys.argv:
        shapes += [(4096, 4096), (4096, 16384), (50400, 4096), (6144, 24576)]
    if "--bloom" in sys.argv:                              # BLOOM-7b1 shapes (BASELINE.json configs[3]: Q4_1 prefill of 2048 tokens)
        shapes = [(4096, 4096), (16384, 4096), (4096, 16384), (250880, 4096)]
    if "--neox" in sys.argv:
        shapes = [(6144, 6144), (24576, 6144), (6144, 24576), (5120, 5120), (20480, 5120)]
    ms_list = [1, 8, 16, 64] if "--m" not in sys.argv else [int(v) for v in sys.argv[sys.argv.index("--m") + 1].split(",")]
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    rng = np.random.default_rng(1)
    for n, k in shapes:
        for qt in ((1,) if "--bloom" in sys.argv else (0, 1)):
            rb = vlib.row_bytes(qt, k)
            nw = max(2, int(np.ceil(400e6 / (n * rb))))
            ws = []
            for i in range(nw):
                raw = rng.integers(0, 256, size=(n, rb), dtype=np.uint8)
                # plausible scales: overwrite the fp32 fields with small positive numbers
                if qt == 0:
                    v = raw.reshape(n, k // 32, 20)
                    v[:, :, :4] = np.full((n, k // 32), 0.01, np.float32).view(np.uint8).reshape(n, k // 32, 4)
                else:
                    raw[:, : (k // 32) * 8] = np.full((n, k // 32 * 2), 0.01, np.float32).view(np.uint8).reshape(n, -1)
                ws.append(vsim_b200.Q4Weight(qt, raw, n, k))
            arr = (C.c_void_p * nw)(*[w.handle for w in ws])
            for m in ms_list:
                x = torch.randn(m, k, device="cuda", dtype=torch.float32)
                y = torch.empty(m, n, device="cuda", dtype=torch.float32)
                torch.cuda.synchronize()
                L.b200q4_time_mul_mat_dev(arr, nw, x.data_ptr(), m, y.data_ptr(), 20)
                t = L.b200q4_time_mul_mat_dev(arr, nw, x.data_ptr(), m, y.data_ptr(), 200)
                if t < 0:
                    print("ERR", vlib._err(L)); continue
                bytes_alg = n * rb + 4 * m * k + 4 * m * n
                gbs = bytes_alg / (t * 1e-3) / 1e9
                print(f"N={n} K={k} type=q4_{qt} M={m:3d}: {t*1e3:8.2f} us/call  {gbs:8.1f} GB/s  {gbs/PEAK*100:5.1f}% of measured HBM peak  "
                      f"{2*m*n*k/(t*1e-3)/1e12:7.2f} TFLOP/s  [{vsim_b200.kernel_for(qt, m, n, k)}]", flush=True)
            for w in ws:
                w.free()


if __name__ == "__main__":
    main()
