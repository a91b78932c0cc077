"""Development aid: a handful of device-resident mul_mat calls at one (type, M) for ncu.
    python tools/tile_prof.py <qt> <M> [N K]
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from vsim_b200 import lib as vlib


def main():
    qt, m = int(sys.argv[1]), int(sys.argv[2])
    n, k = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (16384, 4096)
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    rng = np.random.default_rng(1)
    rb = vlib.row_bytes(qt, k)
    ws = []
    for i in range(4):
        raw = rng.integers(0, 256, size=(n, rb), dtype=np.uint8)
        if qt == 0:
            v = raw.reshape(n, k // 32, 20)
            v[:, :, :4] = np.full((n, k // 32), 0.01, np.float32).view(np.uint8).reshape(n, k // 32, 4)
        else:
            raw[:, : (k // 32) * 8] = np.full((n, k // 32 * 2), 0.01, np.float32).view(np.uint8).reshape(n, -1)
        ws.append(vsim_b200.Q4Weight(qt, raw, n, k))
    arr = (C.c_void_p * 4)(*[w.handle for w in ws])
    x = torch.randn(m, k, device="cuda", dtype=torch.float32)
    y = torch.empty(m, n, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    t = L.b200q4_time_mul_mat_dev(arr, 4, x.data_ptr(), m, y.data_ptr(), 8)
    print(f"qt={qt} M={m} N={n} K={k}: {t * 1e3:.2f} us/call [{vsim_b200.kernel_for(qt, m, n, k)}]")


if __name__ == "__main__":
    main()
