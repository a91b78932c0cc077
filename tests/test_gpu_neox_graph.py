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
def test_decode_program_equals_the_per_op_kernels(tmp_path, monkeypatch, qt, tag, par_res):
    """Decode steps as ONE persistent launch (q4_prog_kernel) against the per-op kernels the tests above pin to the
    restated gptneox_eval, both residual variants (vsim.cpp:636-696).  The KV memory is brought to n positions by an
    exact-mode prompt (bit-identical to the reference on both instances), then ONE fast decode step is compared: the
    two paths share the mul_mat arithmetic and follow the same graph arithmetic (double-accumulated scores, fp16 exp and
    gelu tables, layer-norm in double), so they agree to rounding -- 1e-5 of the scale -- at every context length, across
    the 64-position attention chunks too.  (Steps are compared one at a time because the reference graph re-quantizes
    its activations to 4 bits at every mul_mat: a rounding-level difference that flips one code shows up as an O(1e-2)
    jump and stays in the KV memory -- the reference differs from itself that way between thread counts, DESIGN.md.)
    The program must also be deterministic: the same step twice gives the same bits."""
    import vsim_b200
    from oracle.neox_graph import NeoxOracle
    from tests.helpers.neox_file import write_neox
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    model = tmp_path / f"p-{tag}-{par_res}.bin"
    write_neox(model, qt, n_vocab=160, n_embd=256, n_head=4, n_layer=3, n_rot=32, seed=11, par_res=par_res)
    monkeypatch.setenv("B200NEOX_PROGRAM", "1")
    a = L.b200neox_load(str(model).encode(), 160, 136)
    monkeypatch.setenv("B200NEOX_PROGRAM", "0")
    b = L.b200neox_load(str(model).encode(), 160, 136)
    assert a and b, L.b200neox_last_error()
    g = NeoxOracle(model, n_ctx=160)
    old = L.b200q4_set_exact(0)
    try:
        assert L.b200neox_decode_is_program(a) == 1, L.b200neox_last_error()
        assert L.b200neox_decode_is_program(b) == 0
        nv = L.b200neox_n_vocab(a)

        def ev(m, tokens, n_past):
            t = np.asarray(tokens, np.int32)
            out = np.empty(nv, np.float32)
            rc = L.b200neox_eval(m, t.ctypes.data, len(tokens), n_past, 1, out.ctypes.data)
            assert rc == 0, L.b200neox_last_error()
            return out

        rng = np.random.default_rng(3)
        worst_oracle = 0.0
        for n in (5, 8, 9, 40, 63, 64, 65, 130):
            prompt = rng.integers(0, nv, n).tolist()
            L.b200q4_set_exact(1)
            ev(a, prompt, 0); ev(b, prompt, 0)                 # identical KV memories, the reference's bits
            L.b200q4_set_exact(0)
            tok = int(rng.integers(0, nv))
            launches = vsim_b200.launch_count()
            la = ev(a, [tok], n)
            assert vsim_b200.launch_count() - launches == 1    # one launch per token
            lb = ev(b, [tok], n)
            scale = np.abs(lb).max()
            assert np.abs(la - lb).max() <= 1e-5 * scale, (n, np.abs(la - lb).max(), scale)
            la2 = ev(a, [tok], n)                              # same step again (it rewrites position n with the same values)
            assert np.array_equal(la.view(np.uint32), la2.view(np.uint32)), n
            if n in (5, 65):                                   # and the direct link to the restated reference graph
                g.eval(prompt, 0, 1)
                lo = g.eval([tok], n, 1)
                worst_oracle = max(worst_oracle, float(np.abs(la - lo).max() / np.abs(lo).max()))
        assert worst_oracle <= 1e-3, worst_oracle
        # a greedy run: every step one launch, finite logits, token ids in range
        n_past, nxt = 131, int(np.argmax(la))
        for _ in range(20):
            lg = ev(a, [nxt], n_past)
            assert np.all(np.isfinite(lg))
            n_past += 1
            nxt = int(np.argmax(lg))
    finally:
        L.b200q4_set_exact(old)
        L.b200neox_free(a)
        L.b200neox_free(b)


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
