"""Row f-1 groundwork (CPU): the restatement of the reference's whole GPT-NeoX forward pass (oracle/neox_graph.py)
against the UNMODIFIED reference program (oracle/_ref/vsim-ubuntu): prompt logits equal in every printed digit, greedy
continuation identical, at 1 and 3 ggml threads (the thread count changes the summation order of the V^T x softmax
product, and the restatement follows it)."""
import numpy as np
import pytest

from oracle import Q4_0, Q4_1, REF_DIR

REF_BIN = REF_DIR / "vsim-ubuntu"


@pytest.mark.parametrize("par_res", [1, 0])
@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("qt,tag", [(Q4_0, "q4_0"), (Q4_1, "q4_1")])
def test_forward_pass_restatement_equals_reference_program(tmp_path, qt, tag, threads, par_res):
    if not REF_BIN.exists():
        pytest.skip("oracle/_ref/vsim-ubuntu not built")
    from oracle.neox_graph import NeoxOracle
    from tests.helpers.neox_file import write_neox
    from tests.test_dropin import _logits, _run, _tokens
    model = tmp_path / f"g-{tag}.bin"
    write_neox(model, qt, n_vocab=128, n_embd=128, n_head=4, n_layer=3, n_rot=16, seed=9, par_res=par_res)
    prompt = [1, 2, 3, 4, 5]
    ref_logits = _logits(_run(REF_BIN, model, logits=True, threads=threads))[0]
    out = _run(REF_BIN, model, n_predict=6, threads=threads).replace("use_parallel_residual == 0", "")   # (printed per layer, vsim.cpp:637)
    ref_tokens = _tokens(out)
    g = NeoxOracle(model)
    g.eval(prompt, 0, threads)                               # the program's warm-up eval (vsim.cpp:793) writes the KV memory first
    logits = g.eval(prompt, 0, threads)
    printed = np.array([float("%.8f" % v) for v in logits])  # the program prints %.8f
    assert np.array_equal(printed, ref_logits), np.abs(printed - ref_logits).max()
    toks, n_past = list(prompt), len(prompt)
    nxt = int(np.argmax(logits))
    for _ in range(6):
        toks.append(nxt)
        logits = g.eval([nxt], n_past, threads)
        n_past += 1
        nxt = int(np.argmax(logits))
    assert toks == ref_tokens, (toks, ref_tokens)
