"""Device-resident GPT-NeoX decode (b200neox_eval, row f-1): ms/token on a random-init model written in the reference's
file format.  Decode steps run as ONE persistent launch (q4_prog_kernel); B200NEOX_PROGRAM=0 gives the per-op kernels.
    python tools/neox_bench.py [gptj6b|neox20b|mid|tiny] [n_tokens] [n_past_start]"""
import json, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from tools.synth_model import SHAPES, write_model, q4_weight_bytes
name = sys.argv[1] if len(sys.argv) > 1 else "gptj6b"
ntok = int(sys.argv[2]) if len(sys.argv) > 2 else 32
start = int(sys.argv[3]) if len(sys.argv) > 3 else 5
path = Path("/dev/shm" if Path("/dev/shm").is_dir() else "/tmp") / f"b200neox_{name}.bin"
t0 = time.perf_counter()
size = write_model(path, *SHAPES[name])
t_write = time.perf_counter() - t0
vsim_b200.init(0)
L = vsim_b200.load_library()
t0 = time.perf_counter()
m = L.b200neox_load(str(path).encode(), 512, 8)
assert m, L.b200neox_last_error()
t_load = time.perf_counter() - t0
path.unlink()
nv = L.b200neox_n_vocab(m)
out = np.empty(nv, np.float32)
peak = 6454.6
try:
    peak = json.load(open(Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
wb = q4_weight_bytes(SHAPES[name][0], SHAPES[name][1], SHAPES[name][3])
res = {"model": f"{name}: {SHAPES[name]}, Q4_0, random-init, {size / 1e9:.2f} GB file", "write_s": t_write, "load_and_upload_s": t_load,
       "tokens_timed": ntok, "program": bool(L.b200neox_decode_is_program(m)), "weight_bytes_per_token": wb}
L.b200q4_set_exact(0)
prompt = np.array([1, 2, 3, 4, 5], np.int32)
assert L.b200neox_eval(m, prompt.ctypes.data, 5, 0, 1, out.ctypes.data) == 0, L.b200neox_last_error()
for n0 in (start, 256, 500):
    if n0 + ntok > 512:
        ntok2 = 512 - n0
    else:
        ntok2 = ntok
    n_past, nxt = n0, int(out.argmax())
    toks = []
    t = np.array([nxt], np.int32)
    L.b200neox_eval(m, t.ctypes.data, 1, n_past, 1, out.ctypes.data)           # warm
    t0 = time.perf_counter()
    for _ in range(ntok2):
        t = np.array([nxt], np.int32)
        assert L.b200neox_eval(m, t.ctypes.data, 1, n_past, 1, out.ctypes.data) == 0, L.b200neox_last_error()
        n_past += 1
        nxt = int(out.argmax())
        toks.append(nxt)
    dt = time.perf_counter() - t0
    ms = dt / ntok2 * 1e3
    res[f"n_past_{n0}"] = {"ms_per_token": ms, "tok_per_s": 1e3 / ms, "weight_GBps": wb / ms / 1e6, "frac_hbm_peak": wb / ms / 1e6 / peak,
                           "first_tokens": toks[:6]}
L.b200neox_free(m)
print(json.dumps(res))
