"""Development timing of the decode program kernel (q4_prog_kernel) through b200q4_chain_*: per-op microseconds of
   dep     a chain of dependent microbench-shape mul_mats (op i reads op i-1's output), weights rotated (> 3x L2)
   indep   the same ops without dependencies (pure weight streaming through the resident CTAs)
   l2      20 dependent ops on ONE small tensor (L2-resident): the consumers' compute rate without HBM
   gptj    the 169 mul_mats of a GPT-J-6B-shaped token with the graph's real dependencies (mul_mats only)
python tools/chain_bench.py [dep indep l2 gptj neox20b]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200

PEAK = 6454.6
try:
    PEAK = json.load(open(Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


def rand_q4(rng, qt, n, k):
    rb = (k // 32) * (20 if qt == 0 else 24)
    raw = rng.integers(0, 256, size=(n, rb), dtype=np.uint8)
    if qt == 0:
        raw.reshape(n, k // 32, 20)[:, :, :4] = np.full((n, k // 32), 0.01, np.float32).view(np.uint8).reshape(n, k // 32, 4)
    else:
        raw[:, : (k // 32) * 8] = np.full((n, k // 32 * 2), 0.01, np.float32).view(np.uint8).reshape(n, -1)
    return raw


def time_chain(L, chain, reps, warm=3):
    for _ in range(warm):
        chain.run()
    L.b200q4_sync(None)
    st = time_chain.stream = getattr(time_chain, "stream", None) or torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        chain.run(st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def micro(L, rng, qt, dep, n=16384, k=4096, nw=10, distinct_x=False):
    ws = [vsim_b200.Q4Weight(qt, rand_q4(rng, qt, n, k), n, k) for _ in range(nw)]
    x0 = torch.randn(k, device="cuda")
    xd = [torch.randn(k, device="cuda") for _ in range(nw)]
    ys = [torch.zeros(n, device="cuda") for _ in range(nw)]
    xs = [x0.data_ptr()] + [(ys[i - 1] if dep else (xd[i] if distinct_x else x0)).data_ptr() for i in range(1, nw)]
    ch = vsim_b200.Chain(ws, xs, [y.data_ptr() for y in ys], [-1] + [(i - 1 if dep else -1) for i in range(1, nw)])
    ms = time_chain(L, ch, 30)
    wb = n * (k // 32) * (20 if qt == 0 else 24) + 4 * k + 4 * n
    us = ms * 1e3 / nw
    print(f"micro q4_{qt} N={n} K={k} {'dependent' if dep else 'independent, distinct inputs' if distinct_x else 'independent, one shared input (one pack)'} x{nw}: {us:.2f} us/op  {wb / us / 1e3:.0f} GB/s  "
          f"{wb / us / 1e3 / PEAK * 100:.1f}% of measured peak   {ch.info()}", flush=True)
    ch.free()
    for w in ws:
        w.free()


def l2(L, rng):
    n = k = 4096
    w = vsim_b200.Q4Weight(0, rand_q4(rng, 0, n, k), n, k)
    ys = [torch.zeros(n, device="cuda") for _ in range(2)]
    nops = 20
    xs = [ys[(i + 1) & 1].data_ptr() for i in range(nops)]
    ch = vsim_b200.Chain([w] * nops, xs, [ys[i & 1].data_ptr() for i in range(nops)], [i - 1 for i in range(nops)])
    ms = time_chain(L, ch, 30)
    print(f"l2-resident 4096x4096 x{nops} dependent: {ms * 1e3 / nops:.2f} us/op ({n * k * 0.625 / (ms / nops * 1e-3) / 1e9:.0f} GB/s of weights from L2)", flush=True)
    ch.free()
    xs = [ys[1].data_ptr()] * nops
    ch = vsim_b200.Chain([w] * nops, xs, [ys[0].data_ptr()] * nops, [-1] * nops)
    ms = time_chain(L, ch, 30)
    print(f"l2-resident 4096x4096 x{nops} independent: {ms * 1e3 / nops:.2f} us/op ({n * k * 0.625 / (ms / nops * 1e-3) / 1e9:.0f} GB/s of weights from L2)", flush=True)
    ch.free()
    w.free()


def model_chain(rng, name, NL=None):
    NL0, E, F, V = {"gptj": (28, 4096, 16384, 50400), "neox20b": (44, 6144, 24576, 50432)}[name]
    NL = NL or NL0
    host = [rand_q4(rng, 0, n, k) for n, k in [(E, E)] * 4 + [(F, E), (E, F)]]
    shapes = [(E, E)] * 4 + [(F, E), (E, F)]
    layers = [[vsim_b200.Q4Weight(0, host[j], *shapes[j]) for j in range(6)] for _ in range(NL)]
    head = vsim_b200.Q4Weight(0, rand_q4(rng, 0, V, E), V, E)
    x = torch.randn(E, device="cuda")
    q, k_, v, att, ffo = (torch.zeros(E, device="cuda") for _ in range(5))
    ff = torch.zeros(F, device="cuda")
    logits = torch.zeros(V, device="cuda")
    ws, xs, ys, deps = [], [], [], []
    last = -1
    for lw in layers:                                          # q k v fc_in | fc_out dense   (use_parallel_residual = 1 order)
        base = len(ws)
        ws += [lw[0], lw[1], lw[2], lw[4], lw[5], lw[3]]
        xs += [x.data_ptr()] * 4 + [ff.data_ptr(), v.data_ptr()]
        ys += [q.data_ptr(), k_.data_ptr(), v.data_ptr(), ff.data_ptr(), ffo.data_ptr(), att.data_ptr()]
        deps += [last, -1, -1, -1, base + 3, base + 2]
        last = base + 5
    ws.append(head); xs.append(att.data_ptr()); ys.append(logits.data_ptr()); deps.append(last)
    ch = vsim_b200.Chain(ws, xs, ys, deps)
    ch.keep = (layers, head, x, q, k_, v, att, ffo, ff, logits)
    return ch, sum(h.size for h in host) * NL + V * (E // 32) * 20, len(ws)


def model(L, rng, name):
    ch, wbytes, nws = model_chain(rng, name)
    ws = [None] * nws
    ms = time_chain(L, ch, 10)
    print(f"{name}-shaped token, {len(ws)} mul_mats in one launch: {ms:.3f} ms/token  {1e3 / ms:.0f} tok/s  "
          f"{wbytes / ms / 1e6:.0f} GB/s = {wbytes / ms / 1e6 / PEAK * 100:.1f}% of measured peak   {ch.info()}", flush=True)
    ch.free()


def main():
    what = sys.argv[1:] or ["dep", "indep", "l2", "gptj"]
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    rng = np.random.default_rng(1)
    if "dep" in what:
        micro(L, rng, 0, True); micro(L, rng, 1, True)
    if "indep" in what:
        micro(L, rng, 0, False)
        micro(L, rng, 0, False, distinct_x=True)
        micro(L, rng, 0, False, nw=1)
        micro(L, rng, 0, False, nw=40)
    if "l2" in what:
        l2(L, rng)
    for name in ("gptj", "neox20b"):
        if name in what:
            model(L, rng, name)


if __name__ == "__main__":
    main()
