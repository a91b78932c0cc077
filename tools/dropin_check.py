"""Development aid: logits of the CPU reference binary vs the GPU-backed one on a random model, several thread counts."""
import os, sys, subprocess
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import numpy as np
import dropin_bench as d
from pathlib import Path
p = Path('/dev/shm/chk.bin')
d.write_model(p, *d.SHAPES[sys.argv[1] if len(sys.argv) > 1 else 'gptj6b_4l'])
def logits(binary, threads, env=None):
    cmd = [str(binary), "gptneox", "-m", str(p), "--prompt", "1 2 3 4 5", "--seed", "42", "--threads", str(threads), "--n_predict", "1",
           "--top_k", "1", "--top_p", "1.0", "--temp", "1.0", "--repeat_penalty", "1.0", "--return_logits"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, **(env or {})})
    rows = [np.array(l.split()[1:-1] if l.rstrip().endswith("<END|>") else l.split()[1:], dtype=np.float64) for l in r.stdout.splitlines() if l.startswith("logits:")]
    if not rows:
        print("NO LOGITS rc", r.returncode, "stderr:", r.stderr[-600:], flush=True)
        return np.zeros(50400)
    return rows[0]
res = {}
for tag, b, env in (("cpu", d.REF / 'vsim-ubuntu', None), ("gpu-fast", d.REF / 'vsim-b200', {"B200Q4_EXACT": "0"}), ("gpu-exact", d.REF / 'vsim-b200', {"B200Q4_EXACT": "1"})):
    for th in (1, 4, 16):
        res[(tag, th)] = logits(b, th, env)
for k, v in res.items():
    ref = res[("cpu", k[1])]
    print(k, "argmax", v.argmax(), "max|diff vs cpu at the SAME thread count| / max", np.abs(v - ref).max() / np.abs(ref).max(),
          "| vs cpu/1:", np.abs(v - res[("cpu", 1)]).max() / np.abs(ref).max(), flush=True)
p.unlink()
