"""Row f-1, first slice (GPU): the device-resident GPT-NeoX forward pass (b200neox_eval) against the CPU restatement
of gptneox_eval (oracle/neox_graph.py, itself pinned to the unmodified reference program by
tests/test_neox_graph_oracle.py).  With exact-mode mul_mats the logits must be BIT-IDENTICAL -- prompt batch, decode
steps, and the reference's dependence on its thread count; with the fast kernels they stay within 1e-3 of the scale on
this small, well-conditioned model."""
import ctypes as C

import numpy as np
import pytest

from oracle import Q4_0, Q4_1

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("qt,tag", [(Q4_0, "q4_0"), (Q4_1, "q4_1")])
def test_device_forward_pass_equals_the_restatement(tmp_path, qt, tag, threads):
    import vsim_b200
    from oracle.neox_graph import NeoxOracle
    from tests.helpers.neox_file import write_neox
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    model = tmp_path / f"d-{tag}.bin"
    write_neox(model, qt, n_vocab=128, n_embd=128, n_head=4, n_layer=3, n_rot=16, seed=9)
    g = NeoxOracle(model)
    m = L.b200neox_load(str(model).encode(), 64, 8)
    assert m, L.b200neox_last_error()
    old = L.b200q4_set_exact(1)
    try:
        nv = L.b200neox_n_vocab(m)
        assert nv == 128

        def dev(tokens, n_past):
            t = np.asarray(tokens, np.int32)
            out = np.empty(nv, np.float32)
            rc = L.b200neox_eval(m, t.ctypes.data, len(tokens), n_past, threads, out.ctypes.data)
            assert rc == 0, L.b200neox_last_error()
            return out

        prompt = [1, 2, 3, 4, 5]
        g.eval(prompt, 0, threads); dev(prompt, 0)              # the reference program's warm-up eval
        lo, ld = g.eval(prompt, 0, threads), dev(prompt, 0)
        assert np.array_equal(lo.view(np.uint32), ld.view(np.uint32)), np.abs(lo - ld).max()
        n_past, nxt = len(prompt), int(np.argmax(ld))
        for _ in range(6):
            lo, ld = g.eval([nxt], n_past, threads), dev([nxt], n_past)
            assert np.array_equal(lo.view(np.uint32), ld.view(np.uint32)), (n_past, np.abs(lo - ld).max())
            n_past += 1
            nxt = int(np.argmax(ld))
        # fast kernels: same graph, logits within 1e-3 of the scale
        L.b200q4_set_exact(0)
        lf = dev([nxt], n_past)
        lo = g.eval([nxt], n_past, threads)
        assert np.abs(lf - lo).max() <= 1e-3 * np.abs(lo).max()
    finally:
        L.b200q4_set_exact(old)
        L.b200neox_free(m)
