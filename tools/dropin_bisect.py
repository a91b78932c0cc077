"""Development aid: run vsim-b200 (GPU hook) and vsim-oracle (same patched ggml.c, CPU oracle hook) on the same random
model with B200Q4_DUMP_DIR set, and report the first mul_mat call whose INPUT or OUTPUT differs between the two programs."""
import os, struct, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import dropin_bench as d
name = sys.argv[1] if len(sys.argv) > 1 else 'gptj6b_4l'
EXTRA = {"B200Q4_EXACT": "1"} if "--exact" in sys.argv else {"B200Q4_EXACT": "0"}
p = Path('/dev/shm/bis.bin'); d.write_model(p, *d.SHAPES[name])
dumps = {}
for tag in ("vsim-b200", "vsim-oracle"):
    dump = Path(tempfile.mkdtemp(dir='/dev/shm'))
    cmd = [str(d.REF / tag), "gptneox", "-m", str(p), "--prompt", "1 2 3 4 5", "--seed", "42", "--threads", "4", "--n_predict", "1",
           "--top_k", "1", "--top_p", "1.0", "--temp", "1.0", "--repeat_penalty", "1.0", "--return_logits"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env={**os.environ, "B200Q4_DUMP_DIR": str(dump), **EXTRA})
    rows = [np.array(l.split()[1:-1] if l.rstrip().endswith("<END|>") else l.split()[1:], dtype=np.float64) for l in r.stdout.splitlines() if l.startswith("logits:")]
    dumps[tag] = (dump, rows[0] if rows else None)
    print(tag, "rc", r.returncode, "calls", len(list(dump.glob('call_*.bin'))), r.stderr[-200:])
la, lb = dumps["vsim-b200"][1], dumps["vsim-oracle"][1]
print("logits (with dumping on): max diff / max", np.abs(la - lb).max() / np.abs(lb).max())
fa, fb = sorted(dumps["vsim-b200"][0].glob("call_*.bin")), sorted(dumps["vsim-oracle"][0].glob("call_*.bin"))
shown = 0
for ci, (a, b) in enumerate(zip(fa, fb)):
    ra, rb = np.fromfile(a, dtype=np.uint8), np.fromfile(b, dtype=np.uint8)
    t, n, k, m = struct.unpack_from("<iiii", ra, 0)
    xa, xb = ra[16:16 + m * k * 4].view(np.float32), rb[16:16 + m * k * 4].view(np.float32)
    ya, yb = ra[16 + m * k * 4:].view(np.float32), rb[16 + m * k * 4:].view(np.float32)
    dx = np.abs(xa - xb).max() / (np.abs(xb).max() + 1e-30)
    dy = np.abs(ya - yb).max() / (np.abs(yb).max() + 1e-30)
    thr = 0.0 if EXTRA["B200Q4_EXACT"] == "1" else 1e-4
    if dx > thr or dy > thr or ci < 3:
        print(f"call {ci:3d} N={n} K={k} M={m}: input diff {dx:.3e} ({int((xa != xb).sum())} elements)  output diff {dy:.3e} ({int((ya != yb).sum())} elements)", flush=True)
        shown += 1
        if shown > 14: break
