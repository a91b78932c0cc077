"""Per-op timeline of one decode step of the device-resident GPT-NeoX program (development aid).  Needs a -DB200Q4_TIMELINE build:
    B200Q4_TIMELINE=1 B200Q4_LIB_OUT=vsim_b200/libb200q4_tl.so python -m vsim_b200.build --force
    B200Q4_LIB=vsim_b200/libb200q4_tl.so B200Q4_DBG=1 python tools/neox_timeline.py [gptj6b|neox20b] [n_past] [first layer shown]"""
import ctypes as C
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from tools.synth_model import SHAPES, write_model

name = sys.argv[1] if len(sys.argv) > 1 else "gptj6b"
n_past = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lay = int(sys.argv[3]) if len(sys.argv) > 3 else 10
path = Path("/dev/shm" if Path("/dev/shm").is_dir() else "/tmp") / f"b200neox_tl_{name}.bin"
write_model(path, *SHAPES[name])
vsim_b200.init(0)
L = vsim_b200.load_library()
m = L.b200neox_load(str(path).encode(), 512, 8)
assert m, L.b200neox_last_error()
path.unlink()
L.b200q4_set_exact(0)
nv = L.b200neox_n_vocab(m)
out = np.empty(nv, np.float32)
for p in range(n_past - 3, n_past + 1):
    t = np.array([7], np.int32)
    assert L.b200neox_eval(m, t.ctypes.data, 1, p, 1, out.ctypes.data) == 0, L.b200neox_last_error()
L.b200neox_debug_timeline.restype = C.c_int
L.b200neox_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
cap = 148 * 4096 * 8 + 148 * 16
buf = np.zeros(cap, np.uint64)
kinds = np.zeros(4096, np.int32)
nops = C.c_int(0)
g = L.b200neox_debug_timeline(m, buf.ctypes.data, buf.size, C.byref(nops), kinds.ctypes.data, 4096)
assert g > 0, g
n = nops.value
t = buf[: g * n * 8].reshape(g, n, 8).astype(np.int64)
t0 = t[:, :, 0][t[:, :, 0] > 0].min()
t = np.where(t > 0, t - t0, -1)
per_layer = 7
print(f"{name} n_past {n_past}: {n} ops, grid {g};  us since the first producer stamp")
print(" op kind      N | prod start med | reach med/max | dep ok med | pack med/max | warp0 done med/max | published med/max | step")
prev = None
tot_end = 0.0
for o in range(n):
    col = lambda i: t[:, o, i][t[:, o, i] >= 0]
    med = lambda i: float(np.median(col(i))) / 1e3 if col(i).size else float("nan")
    mx = lambda i: float(col(i).max()) / 1e3 if col(i).size else float("nan")
    k = int(kinds[o]) & 15
    end = mx(5) if k == 0 else mx(3)
    if end == end:
        tot_end = max(tot_end, end)
    if lay * per_layer <= o < (lay + 2) * per_layer or o >= n - 3:
        step = (end - prev) if (prev is not None and end == end) else float("nan")
        print(f"{o:3d} {['mv','attn','sum'][k]:5s} {int(kinds[o]) >> 4:6d} | {med(0):8.2f} | {med(1):8.2f} {mx(1):8.2f} | {med(2):8.2f} | {med(3):8.2f} {mx(3):8.2f} | {med(4):8.2f} {mx(4):8.2f} | {med(5):8.2f} {mx(5):8.2f} | {step:6.2f}")
    if k == 1 and lay * per_layer <= o < (lay + 2) * per_layer:
        has = t[:, o, 0] >= 0
        if has.any():
            seg = [float(np.median(t[has, o, i] - t[has, o, 1])) / 1e3 for i in (0, 2, 4, 5, 6, 7, 3)]
            print(f"      attention on the {int(has.sum())} CTAs with an item, us after reaching the op: item start {seg[0]:.2f} | q/k/v+rope {seg[1]:.2f} | scores {seg[2]:.2f} | softmax {seg[3]:.2f} | PV {seg[4]:.2f} | stored {seg[5]:.2f} | left {seg[6]:.2f}")
    if end == end:
        prev = end
print(f"last stamp at {tot_end:.1f} us")
L.b200neox_free(m)
