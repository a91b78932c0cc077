#!/usr/bin/env python
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libggml_ref.so).

Run in the build container (where /root/reference is mounted):

    make -C oracle ref && python tests/golden/make_golden.py

The reference has no golden vectors of its own; these
files pin the hot path's behaviour to the outputs of the reference's own
code: quantize_row_q4_0/_q4_1 (ggml.c:209-298), ggml_mul_mat +
ggml_graph_compute on Q4_0/Q4_1 weights (ggml.c:4891-5453, imax.c:1133-1230), and the
offline weight quantizers ggml_quantize_q4_0/_q4_1 (utils.cpp:424-541), all called through ctypes.  Inputs are stored next to the outputs, so the
fixtures do not depend on numpy's RNG staying stable.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import Q4_0, Q4_1, Reference, reference_quantize_weights  # noqa: E402

OUT = Path(__file__).resolve().parent


def edge_rows(k: int) -> np.ndarray:
    """Rows (K floats each) whose blocks exercise the quantizers' corners."""
    rng = np.random.default_rng(20261018)
    nb = k // 32
    rows = []
    r = np.zeros(k, np.float32)                                   # all-zero blocks -> d = 0, codes 8 / 0
    rows.append(r)
    rows.append(np.full(k, 0.75, np.float32))                     # constant block (q4_1: d = 0)
    rows.append(-np.abs(rng.standard_normal(k)).astype(np.float32) - 0.1)   # all-negative blocks
    rows.append(np.abs(rng.standard_normal(k)).astype(np.float32) + 0.1)    # all-positive blocks
    # exact .5 ties: amax = 7 -> d = 1, id = 1, values k + 0.5 round away from zero
    t = np.tile(np.array([7, -7, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 3.5, -3.5, 4.5, -4.5, 5.5, -5.5, 6.5, -6.5,
                          6.49999, -6.49999, 0.49999997, -0.49999997, 0, -0.0, 1, -1, 2, -2, 3, -3, 4, -4, 5, -5],
                         np.float32), nb)
    rows.append(t)
    # q4_1 ties: min = 0, max = 15 -> d = 1
    t1 = np.tile(np.array([0, 15, 0.5, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5, 8.5, 9.5, 10.5, 11.5, 12.5, 13.5,
                           14.5, 14.49999, 0.49999997, 7, 8, 9, 1, 2, 3, 4, 5, 6, 10, 11, 12, 13], np.float32), nb)
    rows.append(t1)
    rows.append((rng.standard_normal(k) * 1e30).astype(np.float32))          # huge scale
    rows.append((rng.standard_normal(k) * 1e-30).astype(np.float32))         # tiny but normal scale
    rows.append((rng.standard_normal(k) * 3e-39).astype(np.float32))         # denormal amax -> 1/d = inf corner
    rows.append((rng.standard_normal(k) * 1e-44).astype(np.float32))         # d underflows towards 0
    one_hot = np.zeros(k, np.float32); one_hot[5::32] = -3.25                 # single non-zero per block
    rows.append(one_hot)
    ramp = np.tile(np.linspace(-1, 1, 32, dtype=np.float32), nb)              # monotone ramp
    rows.append(ramp)
    mixed = rng.standard_normal(k).astype(np.float32)
    mixed[::64] = 0.0; mixed[32:64] = 0.0                                      # a zero block inside a normal row
    rows.append(mixed)
    rows.append(rng.uniform(-0.05, 0.05, k).astype(np.float32))               # weight-like
    rows.append(rng.standard_normal(k).astype(np.float32))                    # activation-like
    rows.append((rng.standard_normal(k) * np.exp(rng.uniform(-20, 20, k))).astype(np.float32))  # wild dynamic range
    return np.stack(rows)


def main() -> None:
    ref = Reference()
    rng = np.random.default_rng(1234)

    # ---- quantizer known-answer tests ------------------------------------
    qk = {}
    for k in (64, 256, 4096):
        x = np.concatenate([edge_rows(k), rng.standard_normal((8, k)).astype(np.float32)])
        qk[f"x_k{k}"] = x
        qk[f"q4_0_k{k}"] = ref.quantize_rows(Q4_0, x)
        qk[f"q4_1_k{k}"] = ref.quantize_rows(Q4_1, x)
    np.savez_compressed(OUT / "quantize_kat.npz", **qk)

    # ---- operator known-answer tests -------------------------------------
    mm = {}
    cases = [  # (name, N, K, Ms)
        ("a", 96, 512, (1, 2, 3, 5, 8, 9, 64)),      # N not a multiple of any tile size
        ("b", 40, 4096, (1, 8)),                     # microbench K
        ("c", 130, 64, (1, 7, 17)),                  # smallest legal K (nb = 2), ragged N
        ("d", 256, 1280, (1, 4, 33, 130)),           # nb = 40 (not a power of two), M past one tile
    ]
    for name, n, k, ms in cases:
        wf = rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32)
        wf[3] = 0.0                                   # an all-zero weight row
        mm[f"{name}_shape"] = np.array([n, k], np.int32)
        for qt, tag in ((Q4_0, "q4_0"), (Q4_1, "q4_1")):
            wq = ref.quantize_rows(qt, wf)            # the reference's own row format
            mm[f"{name}_{tag}_w"] = wq
            for m in ms:
                x = rng.standard_normal((m, k)).astype(np.float32)
                if m >= 3:
                    x[1] = 0.0                        # an all-zero activation row
                mm[f"{name}_x_m{m}"] = x if f"{name}_x_m{m}" not in mm else mm[f"{name}_x_m{m}"]
                x = mm[f"{name}_x_m{m}"]
                y1 = ref.mul_mat(qt, wq, n, k, x, nth=1)
                y4 = ref.mul_mat(qt, wq, n, k, x, nth=4)
                assert np.array_equal(y1, y4), "reference result depends on thread count?!"
                mm[f"{name}_{tag}_y_m{m}"] = y1
    np.savez_compressed(OUT / "mul_mat_kat.npz", **mm)

    # ---- transposed-src0 branch (vec_mad; row f-2 of SURVEY section 8) ------------
    tt = {}
    n, k = 128, 96
    wf = rng.uniform(-0.05, 0.05, (k, n)).astype(np.float32)     # K rows of N values
    tt["shape"] = np.array([n, k], np.int32)
    for qt, tag in ((Q4_0, "q4_0"), (Q4_1, "q4_1")):
        wq = ref.quantize_rows(qt, wf)
        tt[f"{tag}_wt"] = wq
        for m in (1, 5):
            x = rng.standard_normal((m, k)).astype(np.float32)
            tt[f"x_m{m}"] = x if f"x_m{m}" not in tt else tt[f"x_m{m}"]
            x = tt[f"x_m{m}"]
            for nth in (1, 3):
                tt[f"{tag}_y_m{m}_t{nth}"] = ref.mul_mat_transposed(qt, wq, n, k, x, nth=nth)
    np.savez_compressed(OUT / "mul_mat_transposed_kat.npz", **tt)

    # ---- offline weight quantizers (utils.cpp ggml_quantize_q4_0/_q4_1; row f-4) -----------------
    wk = {}
    for k in (64, 256):
        x = np.concatenate([edge_rows(k), rng.uniform(-0.05, 0.05, (8, k)).astype(np.float32)])
        wk[f"x_k{k}"] = x
        for qt, tag in ((Q4_0, "q4_0"), (Q4_1, "q4_1")):
            rows, hist = reference_quantize_weights(qt, x)
            wk[f"{tag}_k{k}"] = rows
            wk[f"{tag}_hist_k{k}"] = hist
    np.savez_compressed(OUT / "quantize_weights_kat.npz", **wk)

    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
