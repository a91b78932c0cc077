"""Development aid: decode program (one launch per token) against the per-op fast kernels on small random models."""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vsim_b200
from oracle import Q4_0, Q4_1
from tests.helpers.neox_file import write_neox

vsim_b200.init(0)
L = vsim_b200.load_library()
L.b200q4_set_exact(0)
tmp = Path("/tmp")


def run(tag, **kw):
    n_ctx = kw.pop("n_ctx", 96)
    steps = kw.pop("steps", 3)
    path = tmp / f"dbg_{tag}.bin"
    qt = kw.pop("qt", Q4_0)
    write_neox(path, qt, **kw)
    os.environ["B200NEOX_PROGRAM"] = "1"
    a = L.b200neox_load(str(path).encode(), n_ctx, 8)
    os.environ["B200NEOX_PROGRAM"] = "0"
    b = L.b200neox_load(str(path).encode(), n_ctx, 8)
    assert a and b, L.b200neox_last_error()
    assert L.b200neox_decode_is_program(a) == 1, L.b200neox_last_error()
    nv = kw["n_vocab"]

    def ev(m, toks, n_past):
        t = np.asarray(toks, np.int32); out = np.empty(nv, np.float32)
        assert L.b200neox_eval(m, t.ctypes.data, len(toks), n_past, 1, out.ctypes.data) == 0, L.b200neox_last_error()
        return out
    from oracle.neox_graph import NeoxOracle
    g = NeoxOracle(path, n_ctx=n_ctx)
    prompt = [1, 2, 3, 4, 5]
    ev(a, prompt, 0); lb = ev(b, prompt, 0); lo = g.eval(prompt, 0, 1)
    print(f"{tag}: prompt per-op vs oracle {np.abs(lb - lo).max() / np.abs(lo).max():.2e}")
    n_past, nxt = 5, int(lo.argmax())
    errs, errb = [], []
    for s in range(min(steps, 8)):
        la, lb, lo = ev(a, [nxt], n_past), ev(b, [nxt], n_past), g.eval([nxt], n_past, 1)
        errs.append(float(np.abs(la - lo).max() / np.abs(lo).max()))
        errb.append(float(np.abs(lb - lo).max() / np.abs(lo).max()))
        n_past += 1; nxt = int(lo.argmax())
    print(f"{tag}: {kw}\n   program vs oracle: {['%.1e' % e for e in errs]}\n   per-op  vs oracle: {['%.1e' % e for e in errb]}", flush=True)
    L.b200neox_free(a); L.b200neox_free(b)


run("l0", n_vocab=128, n_embd=128, n_head=4, n_layer=0, n_rot=16, seed=9)
run("l1", n_vocab=128, n_embd=128, n_head=4, n_layer=1, n_rot=16, seed=9)
run("l1_norot", n_vocab=128, n_embd=128, n_head=4, n_layer=1, n_rot=0, seed=9)
run("l1_h1", n_vocab=128, n_embd=128, n_head=1, n_layer=1, n_rot=16, seed=9)
run("l3", n_vocab=128, n_embd=128, n_head=4, n_layer=3, n_rot=16, seed=9)
run("l2_256", n_vocab=160, n_embd=256, n_head=4, n_layer=2, n_rot=32, seed=11)
run("l2_256_pr0", n_vocab=160, n_embd=256, n_head=4, n_layer=2, n_rot=32, seed=11, par_res=0)
run("l2_q41", n_vocab=160, n_embd=256, n_head=4, n_layer=2, n_rot=32, seed=11, qt=Q4_1)
for pr in (1, 0):
    for q in (Q4_0, Q4_1):
        run(f"t3_{q}_{pr}", n_vocab=160, n_embd=256, n_head=4, n_layer=3, n_rot=32, seed=11, par_res=pr, qt=q, n_ctx=160, steps=8)
run("l1_256", n_vocab=160, n_embd=256, n_head=4, n_layer=1, n_rot=32, seed=11)
run("l0_256", n_vocab=160, n_embd=256, n_head=4, n_layer=0, n_rot=32, seed=11)
run("l2_512", n_vocab=160, n_embd=512, n_head=8, n_layer=2, n_rot=32, seed=12)
