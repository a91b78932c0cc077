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


@pytest.mark.parametrize("par_res", [1, 0])
@pytest.mark.parametrize("qt,tag", [(Q4_0, "q4_0"), (Q4_1, "q4_1")])
def test_decode_program_follows_the_restatement(tmp_path, qt, tag, par_res):
    """Decode steps as ONE persistent launch (q4_prog_kernel): prompt through the per-op kernels, then greedy decode steps
    through the program; logits within 1e-3 of the scale of the restated gptneox_eval at every step (both residual
    variants, vsim.cpp:636-696), and the KV memory it leaves behind serves later steps."""
    import vsim_b200
    from oracle.neox_graph import NeoxOracle
    from tests.helpers.neox_file import write_neox
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    model = tmp_path / f"p-{tag}-{par_res}.bin"
    write_neox(model, qt, n_vocab=160, n_embd=256, n_head=4, n_layer=3, n_rot=32, seed=11, par_res=par_res)
    g = NeoxOracle(model, n_ctx=160)
    m = L.b200neox_load(str(model).encode(), 160, 8)
    assert m, L.b200neox_last_error()
    old = L.b200q4_set_exact(0)
    try:
        assert L.b200neox_decode_is_program(m) == 1, L.b200neox_last_error()
        nv = L.b200neox_n_vocab(m)

        def dev(tokens, n_past):
            t = np.asarray(tokens, np.int32)
            out = np.empty(nv, np.float32)
            rc = L.b200neox_eval(m, t.ctypes.data, len(tokens), n_past, 1, out.ctypes.data)
            assert rc == 0, L.b200neox_last_error()
            return out

        prompt = [1, 2, 3, 4, 5]
        lo, ld = g.eval(prompt, 0, 1), dev(prompt, 0)
        assert np.abs(ld - lo).max() <= 1e-3 * np.abs(lo).max()
        n_past, nxt = len(prompt), int(np.argmax(lo))
        launches = vsim_b200.launch_count()
        for step in range(70):                        # past one attention chunk (64 positions): the merged-partials path too
            lo, ld = g.eval([nxt], n_past, 1), dev([nxt], n_past)
            assert np.abs(ld - lo).max() <= 1e-3 * np.abs(lo).max(), (step, np.abs(ld - lo).max(), np.abs(lo).max())
            n_past += 1
            nxt = int(np.argmax(lo))
        assert vsim_b200.launch_count() - launches == 70      # one launch per token
    finally:
        L.b200q4_set_exact(old)
        L.b200neox_free(m)


def test_exact_mode_serial_residual_graph_is_bit_identical(tmp_path):
    """use_parallel_residual = 0 (vsim.cpp:636-665) through the per-op kernels in exact mode: the restatement's bits."""
    import vsim_b200
    from oracle.neox_graph import NeoxOracle
    from tests.helpers.neox_file import write_neox
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    model = tmp_path / "s.bin"
    write_neox(model, Q4_0, n_vocab=128, n_embd=128, n_head=4, n_layer=2, n_rot=16, seed=3, par_res=0)
    g = NeoxOracle(model)
    m = L.b200neox_load(str(model).encode(), 64, 8)
    assert m, L.b200neox_last_error()
    old = L.b200q4_set_exact(1)
    try:
        def dev(tokens, n_past):
            t = np.asarray(tokens, np.int32)
            out = np.empty(128, np.float32)
            assert L.b200neox_eval(m, t.ctypes.data, len(tokens), n_past, 1, out.ctypes.data) == 0, L.b200neox_last_error()
            return out
        lo, ld = g.eval([1, 2, 3, 4, 5], 0, 1), dev([1, 2, 3, 4, 5], 0)
        assert np.array_equal(lo.view(np.uint32), ld.view(np.uint32))
        lo, ld = g.eval([7], 5, 1), dev([7], 5)
        assert np.array_equal(lo.view(np.uint32), ld.view(np.uint32))
    finally:
        L.b200q4_set_exact(old)
        L.b200neox_free(m)


def test_load_rejects_wrong_shapes(tmp_path):
    """'has wrong shape in model file' (vsim.cpp:375-448): a tensor that does not match the hyper-parameters is refused."""
    import struct
    import vsim_b200
    from tests.helpers.neox_file import write_neox
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    model = tmp_path / "ok.bin"
    write_neox(model, Q4_0, n_vocab=64, n_embd=64, n_head=2, n_layer=1, n_rot=8, seed=1)
    raw = bytearray(model.read_bytes())
    for field, value in ((1, 128), (2, 0), (2, 3), (4, 7)):   # n_embd lie, n_head = 0, n_head not dividing, odd n_rot
        bad = bytearray(raw)
        struct.pack_into("<i", bad, 4 + 4 * field, value)
        pth = tmp_path / f"bad{field}_{value}.bin"
        pth.write_bytes(bytes(bad))
        assert not L.b200neox_load(str(pth).encode(), 32, 4)
        assert L.b200neox_last_error()
    m = L.b200neox_load(str(model).encode(), 32, 4)
    assert m
    out = np.empty(64, np.float32)
    t = np.asarray([64], np.int32)                    # token id out of range
    assert L.b200neox_eval(m, t.ctypes.data, 1, 0, 1, out.ctypes.data) != 0
    L.b200neox_free(m)
