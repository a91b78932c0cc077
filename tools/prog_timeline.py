"""Per-op timeline of the decode program kernel (development aid).  Needs a -DB200Q4_TIMELINE build:
    B200Q4_TIMELINE=1 B200Q4_LIB_OUT=vsim_b200/libb200q4_tl.so python -m vsim_b200.build --force
    B200Q4_LIB=vsim_b200/libb200q4_tl.so B200Q4_DBG=1 python tools/prog_timeline.py [dep|l2|l2i|gptj|neox20b]
stamps per (CTA, op): 0 producer starts issuing, 7 producer done issuing, 1 consumers reach the op, 2 dependency satisfied,
3 pack built, 4 consumer warp 0 done with its last stage, 6 reducer sees the op's first stage, 5 CTA published."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from tools.chain_bench import rand_q4, model_chain

vsim_b200.init(0)
L = vsim_b200.load_library()
L.b200q4_chain_timeline.restype = C.c_int
L.b200q4_chain_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
rng = np.random.default_rng(1)
what = sys.argv[1] if len(sys.argv) > 1 else "dep"
show = None
if what in ("l2", "l2i"):
    n = k = 4096; nops = 12
    w = vsim_b200.Q4Weight(0, rand_q4(rng, 0, n, k), n, k)
    ws = [w] * nops
    ys = [torch.zeros(n, device="cuda") for _ in range(2)]
    if what == "l2":
        xs = [ys[(i + 1) & 1].data_ptr() for i in range(nops)]
        yp = [ys[i & 1].data_ptr() for i in range(nops)]
        ch = vsim_b200.Chain(ws, xs, yp, [-1] + list(range(nops - 1)))
    else:
        ch = vsim_b200.Chain(ws, [ys[1].data_ptr()] * nops, [ys[0].data_ptr()] * nops, [-1] * nops)
elif what == "hl":                                        # bench.py's headline launch: independent steps, own input rows, rotating weights
    n, k = 16384, 4096
    nops = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    wl_ = [vsim_b200.Q4Weight(0, rand_q4(rng, 0, n, k), n, k) for _ in range(10)]
    ws = [wl_[i % 10] for i in range(nops)]
    xin = torch.randn(nops, k, device="cuda")
    yout = torch.zeros(nops, n, device="cuda")
    ch = vsim_b200.Chain(ws, [xin[i].data_ptr() for i in range(nops)], [yout[i].data_ptr() for i in range(nops)], [-1] * nops)
elif what in ("gptj", "neox20b"):
    ch, _, nops = model_chain(rng, what, NL=4)
    show = list(range(6, 19))
else:
    n, k, nops = 16384, 4096, 10
    ws = [vsim_b200.Q4Weight(0, rand_q4(rng, 0, n, k), n, k) for _ in range(nops)]
    x0 = torch.randn(k, device="cuda")
    ys = [torch.zeros(n, device="cuda") for _ in range(nops)]
    xs = [x0.data_ptr()] + [ys[i - 1].data_ptr() for i in range(1, nops)]
    yp = [y.data_ptr() for y in ys]
    ch = vsim_b200.Chain(ws, xs, yp, [-1] + list(range(nops - 1)))
for _ in range(4):
    ch.run()
L.b200q4_sync(None)
info = ch.info()
buf = np.zeros(148 * info["ops"] * 8 + 148 * 16 + 2048 * 8, np.uint64)
g = L.b200q4_chain_timeline(ch.handle, buf.ctypes.data, buf.size)
assert g > 0, g
acc = buf[g * info["ops"] * 8: g * info["ops"] * 8 + g * 16].reshape(g, 16).astype(np.float64) / 1.9e3     # cycles -> us at ~1.9 GHz
t = buf[: g * info["ops"] * 8].reshape(g, info["ops"], 8).astype(np.int64)
t0 = t[:, 0, 0].min()
t = np.where(t > 0, t - t0, -1)
print(f"{what}: grid {g}, {info}")
print(" op | prod start min/med | prod done med | reach med/max | dep ok med | pack med/max | reducer 1st med | warp0 done med/max | published med/max | op time")
prev = 0
for o in (show or range(nops)):
    col = lambda i: t[:, o, i][t[:, o, i] >= 0]
    med = lambda i: float(np.median(col(i))) / 1e3 if col(i).size else float("nan")
    mx = lambda i: float(col(i).max()) / 1e3 if col(i).size else float("nan")
    mn = lambda i: float(col(i).min()) / 1e3 if col(i).size else float("nan")
    print(f"{o:3d} | {mn(0):7.2f} {med(0):7.2f} | {med(7):7.2f} | {med(1):7.2f} {mx(1):7.2f} | {med(2):7.2f} | {med(3):7.2f} {mx(3):7.2f} | {med(6):7.2f} | {med(4):7.2f} {mx(4):7.2f} | {med(5):7.2f} {mx(5):7.2f} | {mx(5) - prev:6.2f}")
    prev = mx(5)

names = ["producer waits for a free slot", "consumer warp 0 waits for a full stage", "consumer warp 0 in pg_begin_op", "reducer waits for consumers", "reducer total", "consumer warp 0 computing its units"]
print(f"  stages per CTA: median {np.median(acc[:, 6]) * 1.9e3:.0f};  waited-for stages: {np.median(acc[:, 9]) * 1.9e3:.0f}, their issue-to-arrival latency: {np.median(acc[:, 8] / np.maximum(acc[:, 9], 1e-9)):.2f} us")
print(f"  consumer warp 0 waits for the first stage of an op (peek): median {np.median(acc[:, 7]):8.2f} us")
for i, nm in enumerate(names):
    print(f"  {nm}: median {np.median(acc[:, i]):8.2f} us   min {acc[:, i].min():8.2f}   max {acc[:, i].max():8.2f}   (per launch of {nops} ops)")

# per-stage log of CTA 0 (clock64 of its SM, 1.965 GHz)
sl = buf[g * info["ops"] * 8 + g * 16:][: 1024 * 8].reshape(1024, 8).astype(np.int64)
nst = int((sl[:, 0] > 0).sum())
c0 = sl[0, 0]
print(f"CTA 0: {nst} ring entries;  us since the first issue:  issue | full seen | warp0 done | warp15 done | reducer sees | released | op(END) units")
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for i in range(min(nst, lim)):
    r = sl[i]
    f = lambda v: f"{(v - c0) / 1965.0:8.2f}" if v > 0 else "     nan"
    o = int(r[6]); end = "E" if o & (1 << 30) else " "
    print(f"  {i:3d}  {f(r[0])} {f(r[1])} {f(r[2])} {f(r[3])} {f(r[4])} {f(r[5])}   {o & ((1 << 30) - 1):3d}{end} {int(np.int64(r[7])):3d}")

wl = buf[g * info["ops"] * 8 + g * 16 + 1024 * 8:][: 1024 * 8].reshape(1024, 8).astype(np.int64)
print("warp 15 of CTA 0, cycles: wait | decode | octet | store | arrive | loop-back to next top")
for i in range(min(nst - 1, lim)):
    r = wl[i]
    if r[0] <= 0: continue
    print(f"  {i:3d}  {r[1]-r[0]:6d} {r[2]-r[1]:6d} {r[3]-r[2]:6d} {r[4]-r[3]:6d} {r[5]-r[4]:6d} {wl[i+1][0]-r[5]:6d}")
