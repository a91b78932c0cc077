"""Device-resident GPT-NeoX decode (b200neox_eval, row f-1): ms/token on a random-init model written in the reference's
file format -- the same model tools/dropin_bench.py runs through the reference program.
    python tools/neox_bench.py [gptj6b|neox20b|mid] [n_tokens]"""
import json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
import dropin_bench as d
import vsim_b200
name = sys.argv[1] if len(sys.argv) > 1 else "gptj6b"
ntok = int(sys.argv[2]) if len(sys.argv) > 2 else 32
path = Path("/dev/shm" if Path("/dev/shm").is_dir() else "/tmp") / f"b200neox_{name}.bin"
size = d.write_model(path, *d.SHAPES[name])
vsim_b200.init(0)
L = vsim_b200.load_library()
t0 = time.perf_counter()
m = L.b200neox_load(str(path).encode(), 512, 8)
assert m, L.b200neox_last_error()
t_load = time.perf_counter() - t0
nv = L.b200neox_n_vocab(m)
out = np.empty(nv, np.float32)
res = {"model": f"{name}: {d.SHAPES[name]}, Q4_0, random-init, {size / 1e9:.2f} GB file", "load_and_upload_s": t_load, "tokens_timed": ntok}
for mode, tag in ((0, "fast_kernels"), (1, "exact_mode")):
    L.b200q4_set_exact(mode)
    prompt = np.array([1, 2, 3, 4, 5], np.int32)
    assert L.b200neox_eval(m, prompt.ctypes.data, 5, 0, 1, out.ctypes.data) == 0, L.b200neox_last_error()
    n_past, nxt = 5, int(out.argmax())
    toks = []
    t0 = time.perf_counter()
    for _ in range(ntok):
        t = np.array([nxt], np.int32)
        assert L.b200neox_eval(m, t.ctypes.data, 1, n_past, 1, out.ctypes.data) == 0, L.b200neox_last_error()
        n_past += 1
        nxt = int(out.argmax())
        toks.append(nxt)
    dt = time.perf_counter() - t0
    res[tag] = {"ms_per_token": dt / ntok * 1e3, "tok_per_s": ntok / dt, "first_tokens": toks[:8]}
L.b200q4_set_exact(0)
L.b200neox_free(m)
path.unlink()
print(json.dumps(res))
