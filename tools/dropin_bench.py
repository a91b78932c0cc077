#!/usr/bin/env python
"""Model-level drop-in measurement: the reference's OWN main program (vsim.cpp NeoX graph + ggml.c executor), once as
the unmodified CPU build (oracle/_ref/vsim-ubuntu) and once linked against libb200q4.so instead of imax.c
(oracle/_ref/vsim-b200), greedy-decoding the same random-init model written in the reference's q4 file format.

    python tools/dropin_bench.py [--model gptj6b|neox20b|mid] [--threads T] [--cpu-n 4,20] [--gpu-n 16,272] [--keep]

ms/token = (T_PREDICT(n_predict = n1) - T_PREDICT(n_predict = n0)) / (n1 - n0), T_PREDICT being the reference's own
timer around gptneox_eval in its generate loop (vsim.cpp:805-810, printed by monitor.c with 0.1 s resolution -- the
token counts are chosen so that the difference is seconds): model load, weight upload and the prompt cancel out.
Everything except the Q4 mul_mats (norms, rope, attention, softmax, gelu, sampling) runs on the host CPU in BOTH builds -- this is what swapping the IMAX hook alone buys inside the reference program (SURVEY.md
section 8, rows a-b); a device-resident graph is row f-1.  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import struct
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "oracle" / "_ref"
SHAPES = {  # name: (n_vocab, n_embd, n_head, n_layer, n_rot)
    "gptj6b": (50400, 4096, 16, 28, 64),
    "neox20b": (50432, 6144, 64, 44, 24),
    "mid": (8192, 2048, 16, 8, 32),
    "gpt2_117m": (50257, 768, 12, 12, 16),       # BASELINE.json configs[0] as a NeoX-shaped stand-in (the reference ships no GPT-2 driver)
    "gptj6b_4l": (50400, 4096, 16, 4, 64),     # GPT-J widths, 4 layers: how far rounding-level differences have grown by then
    "gptj6b_12l": (50400, 4096, 16, 12, 64),
}


def write_model(path: Path, n_vocab, n_embd, n_head, n_layer, n_rot, seed=0, gain=0.5) -> int:
    """Random-init GPT-NeoX in the reference's on-disk format (vsim.cpp:119-138 header, :163-173 vocab, :375-448
    tensors), Q4_0: random nibbles, block scales gain/(4.6 sqrt(K)).  gain < 1 keeps the random network contractive:
    with unit-gain random matrices a 28-layer stack is chaotic -- rounding-level differences between two correct
    mul_mat implementations (or two thread counts) grow to O(1) in the logits, which says nothing about either."""
    rng = np.random.default_rng(seed)

    def t2(f, name, k, n):
        nm = name.encode()
        f.write(struct.pack("<iii", 2, len(nm), 2) + struct.pack("<ii", k, n) + nm)
        d = np.float32(gain / (4.6 * np.sqrt(k)))
        rows = 4096
        for r0 in range(0, n, rows):                      # in slabs: the 20B lm_head alone is 190 MB
            nr = min(rows, n - r0)
            raw = rng.integers(0, 256, size=(nr, k // 32, 20), dtype=np.uint8)
            raw[:, :, :4] = np.frombuffer(d.tobytes(), np.uint8)
            f.write(raw.tobytes())

    def t1(f, name, n, mean=0.0, scale=0.02):
        nm = name.encode()
        v = (mean + rng.standard_normal(n) * scale).astype(np.float32)
        f.write(struct.pack("<iii", 1, len(nm), 0) + struct.pack("<i", n) + nm + v.tobytes())

    with open(path, "wb") as f:
        f.write(struct.pack("<i", 0x67676d6c))
        f.write(struct.pack("<iiiiiii", n_vocab, n_embd, n_head, n_layer, n_rot, 1, 2))
        for i in range(n_vocab):
            tok = f"t{i}".encode()
            f.write(struct.pack("<I", len(tok)) + tok)
        t2(f, "gpt_neox.embed_in.weight", n_embd, n_vocab)
        t1(f, "gpt_neox.final_layer_norm.weight", n_embd, 1.0)
        t1(f, "gpt_neox.final_layer_norm.bias", n_embd)
        t2(f, "embed_out.weight", n_embd, n_vocab)
        for l in range(n_layer):
            p = f"gpt_neox.layers.{l}."
            t1(f, p + "input_layernorm.weight", n_embd, 1.0)
            t1(f, p + "input_layernorm.bias", n_embd)
            t1(f, p + "post_attention_layernorm.weight", n_embd, 1.0)
            t1(f, p + "post_attention_layernorm.bias", n_embd)
            for nm in ("query", "key", "value", "dense"):
                t2(f, p + f"attention.{nm}.weight", n_embd, n_embd)
                t1(f, p + f"attention.{nm}.bias", n_embd)
            t2(f, p + "mlp.dense_h_to_4h.weight", n_embd, 4 * n_embd)
            t1(f, p + "mlp.dense_h_to_4h.bias", 4 * n_embd)
            t2(f, p + "mlp.dense_4h_to_h.weight", 4 * n_embd, n_embd)
            t1(f, p + "mlp.dense_4h_to_h.bias", n_embd)
    return path.stat().st_size


GPU_ENV: dict = {}


def prompt_logits(binary: Path, model: Path, threads: int, timeout: int):
    cmd = [str(binary), "gptneox", "-m", str(model), "--prompt", "1 2 3 4 5", "--seed", "42", "--threads", str(threads),
           "--n_predict", "1", "--top_k", "1", "--top_p", "1.0", "--temp", "1.0", "--repeat_penalty", "1.0", "--return_logits"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env={**os.environ, **(GPU_ENV if binary.name == "vsim-b200" else {})})
    rows = [np.array(l.split()[1:-1] if l.rstrip().endswith("<END|>") else l.split()[1:], dtype=np.float64)
            for l in r.stdout.splitlines() if l.startswith("logits:")]
    if r.returncode != 0 or not rows:
        raise RuntimeError(f"{binary.name} --return_logits failed: {r.stderr[-800:]}")
    return rows[0]


def run(binary: Path, model: Path, n_predict: int, threads: int, timeout: int):
    global GPU_ENV
    cmd = [str(binary), "gptneox", "-m", str(model), "--prompt", "1 2 3 4 5", "--seed", "42", "--threads", str(threads),
           "--n_predict", str(n_predict), "--top_k", "1", "--top_p", "1.0", "--temp", "1.0", "--repeat_penalty", "1.0"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env={**os.environ, **(GPU_ENV if binary.name.startswith("vsim-b200") else {})})
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"{binary.name} failed: {r.stderr[-800:]}{r.stdout[-800:]}")
    body = r.stdout.split("<|BEGIN>")[1].split("<END|>")[0]
    m = re.search(r"T_PREDICT\s*:\s*([0-9.]+)s", r.stdout)
    if not m:
        raise RuntimeError("no T_PREDICT line in the program's time report")
    return float(m.group(1)), dt, [int(t) for t in body.split()]


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gptj6b", choices=sorted(SHAPES))
    ap.add_argument("--threads", type=int, default=min(os.cpu_count() or 1, 16), help="ggml threads of the CPU build")
    ap.add_argument("--gpu-threads", type=int, default=4,
                    help="ggml threads of the GPU-backed build (its host work is small: the spin barriers of a 16-thread pool cost more than they buy)")
    ap.add_argument("--gain", type=float, default=0.5, help="per-matrix gain of the random model (see write_model)")
    ap.add_argument("--cpu-n", default="4,20", help="n_predict pair for the CPU build")
    ap.add_argument("--gpu-n", default="16,272", help="n_predict pair for the GPU-backed build (n_ctx is 512)")
    ap.add_argument("--exact", action="store_true", help="GPU-backed build in exact mode (B200Q4_EXACT=1: bit-identical mul_mats)")
    ap.add_argument("--cpu-only", action="store_true", help="only the unmodified CPU build (no GPU needed): BASELINE.json configs[0]")
    ap.add_argument("--keep", action="store_true", help="keep the model file")
    ap.add_argument("--timeout", type=int, default=900)
    args = ap.parse_args()
    GPU_ENV["B200Q4_EXACT"] = "1" if args.exact else "0"      # (the hook's own default is exact mode)
    cpu_bin, gpu_bin = REF / "vsim-ubuntu", REF / "vsim-b200"
    if not (cpu_bin.exists() and gpu_bin.exists()):
        print(json.dumps({"unavailable": "oracle/_ref/vsim-ubuntu / vsim-b200 not built (make -C oracle needs /root/reference)"}))
        return 0
    shm = Path("/dev/shm")
    base = shm if shm.is_dir() and os.access(shm, os.W_OK) else Path("/tmp")
    model = base / f"b200q4_dropin_{args.model}.bin"
    nv, ne, nh, nl, nr = SHAPES[args.model]
    t0 = time.perf_counter()
    size = write_model(model, nv, ne, nh, nl, nr, gain=args.gain)
    t_write = time.perf_counter() - t0
    try:
        res = {}
        if args.cpu_only:
            n0, n1 = (int(v) for v in args.cpu_n.split(","))
            pa, wa, _ = run(cpu_bin, model, n0, args.threads, args.timeout)
            pb, wb, toks = run(cpu_bin, model, n1, args.threads, args.timeout)
            ms = (pb - pa) / (n1 - n0) * 1e3
            print(json.dumps({"what": "unmodified reference main program (oracle/_ref/vsim-ubuntu: vsim.cpp + ggml.c + imax.c, Makefile-ubuntu flags), "
                                      "greedy decode, batch 1, CPU only",
                              "model": f"{args.model}: n_vocab {nv}, n_embd {ne}, n_head {nh}, n_layer {nl}, n_rot {nr}, Q4_0, random-init, {size / 1e6:.0f} MB file",
                              "threads": args.threads, "host_cores": os.cpu_count(), "n_predict": [n0, n1], "t_predict_s": [pa, pb],
                              "ms_per_token": ms, "tok_per_s": 1e3 / ms, "tokens_decoded": len(toks), "first_tokens": toks[:12]}))
            return 0
        lg = prompt_logits(gpu_bin, model, args.gpu_threads, args.timeout)
        lc = prompt_logits(cpu_bin, model, args.threads, args.timeout)
        legs = [("b200_hook", gpu_bin, args.gpu_n, args.gpu_threads), ("cpu_reference", cpu_bin, args.cpu_n, args.threads)]
        dev_bin = REF / "vsim-b200dev"
        if dev_bin.exists():                               # the same program on the device-resident graph (vsim.cpp + 3-line patch)
            legs.insert(1, ("b200_device_graph", dev_bin, args.gpu_n, 1))
        for tag, b, pair, th in legs:
            n0, n1 = (int(v) for v in pair.split(","))
            pa, wa, _ = run(b, model, n0, th, args.timeout)
            pb, wb, toks_b = run(b, model, n1, th, args.timeout)
            res[tag] = {"ms_per_token": (pb - pa) / (n1 - n0) * 1e3, "wall_s": [wa, wb], "t_predict_s": [pa, pb],
                        "n_predict": [n0, n1], "tokens": toks_b}
        g, c = res["b200_hook"], res["cpu_reference"]
        line = {
            "what": "reference main program (vsim.cpp NeoX graph + ggml.c executor), greedy decode, batch 1: unmodified CPU build vs the "
                    "same sources with the IMAX hook served by libb200q4.so (only the Q4 mul_mats run on the GPU)",
            "model": f"{args.model}: n_vocab {nv}, n_embd {ne}, n_head {nh}, n_layer {nl}, n_rot {nr}, Q4_0, random-init, {size / 1e9:.2f} GB file",
            "threads": {"cpu_reference": args.threads, "b200_hook": args.gpu_threads}, "host_cores": os.cpu_count(),
            "prompt_logits_max_abs_diff_over_max_abs": float(np.abs(lg - lc).max() / np.abs(lc).max()),
            "prompt_logits_argmax": {"cpu_reference": int(lc.argmax()), "b200_hook": int(lg.argmax())},
            "random_model_gain": args.gain, "b200_mode": "exact (B200Q4_EXACT=1, the hook's default)" if args.exact else "fast kernels (B200Q4_EXACT=0)",
            "n_predict": {"cpu_reference": c["n_predict"], "b200_hook": g["n_predict"]}, "prompt_tokens": 5,
            "t_predict_s": {"cpu_reference": c["t_predict_s"], "b200_hook": g["t_predict_s"]},
            "cpu_reference_ms_per_token": c["ms_per_token"], "b200_hook_ms_per_token": g["ms_per_token"],
            "tok_per_s": {"cpu_reference": 1e3 / c["ms_per_token"], "b200_hook": 1e3 / g["ms_per_token"]},
            "speedup": c["ms_per_token"] / g["ms_per_token"],
            "b200_device_graph_ms_per_token": res.get("b200_device_graph", {}).get("ms_per_token"),
            "b200_device_graph_first_tokens": res.get("b200_device_graph", {}).get("tokens", [])[:12],
            "wall_s": {"cpu_reference": c["wall_s"], "b200_hook": g["wall_s"], "write_model": t_write},
            "greedy_tokens_identical_on_common_prefix": c["tokens"] == g["tokens"][:len(c["tokens"])],
            "first_tokens": {"cpu_reference": c["tokens"][:12], "b200_hook": g["tokens"][:12]},
        }
        print(json.dumps(line))
    finally:
        if not args.keep:
            model.unlink(missing_ok=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
