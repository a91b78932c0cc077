"""Development aid: run the GPU-backed reference program on the random GPT-J-shaped model with B200Q4_DUMP_DIR set, then
check EVERY dumped mul_mat call (real activations of the graph) against the oracle with the weights read back from the file."""
import os, struct, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import dropin_bench as d
from oracle import Oracle
orc = Oracle()
name = sys.argv[1] if len(sys.argv) > 1 else 'gptj6b'
nv, ne, nh, nl, nr = d.SHAPES[name]
p = Path('/dev/shm/chk.bin'); d.write_model(p, nv, ne, nh, nl, nr)
dump = Path(tempfile.mkdtemp(dir='/dev/shm'))
cmd = [str(d.REF / 'vsim-b200'), "gptneox", "-m", str(p), "--prompt", "1 2 3 4 5", "--seed", "42", "--threads", "4", "--n_predict", "2",
       "--top_k", "1", "--top_p", "1.0", "--temp", "1.0", "--repeat_penalty", "1.0"]
r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "B200Q4_EXACT": "0", "B200Q4_DUMP_DIR": str(dump)})
print("rc", r.returncode, r.stderr[-300:])
# Q4 tensors of the file in order
blob = np.memmap(p, dtype=np.uint8, mode='r')
off = 4 + 28
for i in range(nv):
    ln = struct.unpack_from("<I", blob, off)[0]; off += 4 + ln
tens = {}
while off < blob.size:
    nd, ln, ft = struct.unpack_from("<iii", blob, off); off += 12
    dims = struct.unpack_from("<" + "i" * nd, blob, off); off += 4 * nd
    nm = bytes(blob[off:off + ln]).decode(); off += ln
    if nd == 2:
        k, n = dims; nbytes = n * (k // 32) * 20
        tens[nm] = (off, n, k); off += nbytes
    else:
        off += dims[0] * 4
order = []
for l in range(nl):
    pre = f"gpt_neox.layers.{l}."
    # graph order: the K and V copies into the KV memory are expanded first (vsim.cpp:557-558), Q follows
    order += [pre + "attention.key.weight", pre + "attention.value.weight", pre + "attention.query.weight", pre + "attention.dense.weight",
              pre + "mlp.dense_h_to_4h.weight", pre + "mlp.dense_4h_to_h.weight"]
order.append("embed_out.weight")
files = sorted(dump.glob("call_*.bin"))
print(len(files), "calls dumped")
worst = 0.0
for ci, f in enumerate(files):
    raw = np.fromfile(f, dtype=np.uint8)
    t, n, k, m = struct.unpack_from("<iiii", raw, 0)
    x = raw[16:16 + m * k * 4].view(np.float32).reshape(m, k)
    y = raw[16 + m * k * 4:].view(np.float32).reshape(m, n)
    nm = order[ci % len(order)]
    o, n2, k2 = tens[nm]
    assert (n2, k2) == (n, k), (nm, n2, k2, n, k)
    wq = np.ascontiguousarray(blob[o:o + n * (k // 32) * 20]).reshape(n, -1)
    yr = orc.mul_mat(0, wq, n, k, np.ascontiguousarray(x), nth=8)
    err = np.abs(y - yr).max() / (np.abs(yr).max() + 1e-30)
    worst = max(worst, err)
    flag = "  <<<<<< BAD" if err > 1e-4 or not np.isfinite(y).all() else ""
    if flag or ci < 8 or ci % 40 == 0:
        print(f"call {ci:3d} {nm:45s} N={n} K={k} M={m}: err/max|y| {err:.3e}  |x|max {np.abs(x).max():.3g} zeros {np.mean(x == 0):.2f}{flag}", flush=True)
print("worst", worst)
