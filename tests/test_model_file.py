"""The model-file reader (b200q4_model_load; SURVEY section 8 row f-3, first half): the reference's on-disk format
(vsim.cpp:108-458) indexed on the host -- no GPU needed -- and, on a B200, its Q4 tensors uploaded straight from the
file mapping."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import Q4_0, Q4_1, row_bytes


def _lib():
    import vsim_b200
    return vsim_b200.load_library()


def _tensors(L, m):
    out = []
    for i in range(L.b200q4_model_tensor_count(m)):
        name, nd, ft, ne0, ne1 = C.c_char_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        off, nb = C.c_size_t(), C.c_size_t()
        assert L.b200q4_model_tensor_info(m, i, C.byref(name), C.byref(nd), C.byref(ft), C.byref(ne0), C.byref(ne1),
                                          C.byref(off), C.byref(nb)) == 0
        out.append((name.value.decode(), nd.value, ft.value, ne0.value, ne1.value, off.value, nb.value))
    return out


@pytest.mark.parametrize("qt", [Q4_0, Q4_1])
def test_index_of_a_neox_file(tmp_path, qt):
    from tests.helpers.neox_file import write_neox
    L = _lib()
    path = tmp_path / "m.bin"
    write_neox(path, qt, n_vocab=128, n_embd=128, n_head=4, n_layer=2, n_rot=16)
    m = L.b200q4_model_load(str(path).encode(), 7, 0, 0)
    assert m, L.b200q4_last_error()
    try:
        assert [L.b200q4_model_header(m, i) for i in range(7)] == [128, 128, 4, 2, 16, 1, 2 + qt]
        ts = _tensors(L, m)
        assert len(ts) == 4 + 2 * 16
        names = [t[0] for t in ts]
        assert names[0] == "gpt_neox.embed_in.weight" and names[3] == "embed_out.weight"
        assert "gpt_neox.layers.1.mlp.dense_4h_to_h.weight" in names
        blob = path.read_bytes()
        for name, nd, ft, ne0, ne1, off, nb in ts:
            if nd == 2:
                assert ft == 2 + qt and nb == ne1 * row_bytes(qt, ne0)
            else:
                assert ft == 0 and nb == ne0 * 4
            assert blob[off - len(name):off].decode() == name          # the record's name sits right before its data
        assert ts[-1][5] + ts[-1][6] == len(blob)                       # the index covers the file exactly
        i = names.index("gpt_neox.layers.0.mlp.dense_h_to_4h.weight")
        assert (ts[i][3], ts[i][4]) == (128, 512)                       # ne0 = K, ne1 = N rows
        data = C.string_at(L.b200q4_model_tensor_data(m, i), ts[i][6])
        assert data == blob[ts[i][5]:ts[i][5] + ts[i][6]]
        assert L.b200q4_model_weight(m, i) is None                      # nothing uploaded with upload == 0
    finally:
        L.b200q4_model_free(m)


def test_quantize_tool_style_file_with_vocab_count_and_bad_files(tmp_path):
    """quantize_{gptj,gpt2,bloom}.cpp write 6 header ints and a vocabulary with a leading count."""
    L = _lib()
    w = np.arange(2 * 20, dtype=np.uint8).tobytes()                     # one 2 x 32 q4_0 tensor
    body = struct.pack("<i", 0x67676d6c) + struct.pack("<iiiiii", 3, 32, 2, 1, 8, 2) + struct.pack("<i", 3)
    for tok in (b"a", b"bc", b""):
        body += struct.pack("<I", len(tok)) + tok
    body += struct.pack("<iii", 2, 1, 2) + struct.pack("<ii", 32, 2) + b"w" + w
    body += struct.pack("<iii", 1, 1, 0) + struct.pack("<i", 3) + b"b" + np.float32([1, 2, 3]).tobytes()
    p = tmp_path / "j.bin"
    p.write_bytes(body)
    m = L.b200q4_model_load(str(p).encode(), 6, 1, 0)
    assert m, L.b200q4_last_error()
    ts = _tensors(L, m)
    assert [(t[0], t[1], t[2], t[3], t[4], t[6]) for t in ts] == [("w", 2, 2, 32, 2, 40), ("b", 1, 0, 3, 1, 12)]
    L.b200q4_model_free(m)
    for bad, what in ((b"\\x00" * 64, "magic"), (body[:-5], "truncated"), (body[:60], "truncated")):
        q = tmp_path / "bad.bin"
        q.write_bytes(bad)
        assert not L.b200q4_model_load(str(q).encode(), 6, 1, 0)
        assert what in L.b200q4_last_error().decode()
    assert not L.b200q4_model_load(str(tmp_path / "missing.bin").encode(), 6, 1, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("qt", [Q4_0, Q4_1])
def test_upload_straight_from_the_mapping(tmp_path, oracle, qt):
    import vsim_b200
    from tests.helpers.neox_file import write_neox
    vsim_b200.init(0)
    L = _lib()
    L.b200q4_set_exact(0)
    path = tmp_path / "m.bin"
    write_neox(path, qt, n_vocab=128, n_embd=256, n_head=4, n_layer=1, n_rot=16)
    m = L.b200q4_model_load(str(path).encode(), 7, 0, 1)
    assert m, L.b200q4_last_error()
    try:
        ts = _tensors(L, m)
        rng = np.random.default_rng(0)
        checked = 0
        for i, (name, nd, ft, k, n, off, nb) in enumerate(ts):
            h = L.b200q4_model_weight(m, i)
            if nd != 2:
                assert h is None
                continue
            rows = np.frombuffer(C.string_at(L.b200q4_model_tensor_data(m, i), nb), np.uint8).reshape(n, -1)
            back = np.zeros_like(rows)
            assert L.b200q4_weight_download(h, back.ctypes.data, 0) == 0 and np.array_equal(back, rows)
            x = rng.standard_normal((3, k)).astype(np.float32)
            y = np.zeros((3, n), np.float32)
            assert L.b200q4_mul_mat(h, x.ctypes.data, 3, 0, y.ctypes.data, 0) == 0
            yr = oracle.mul_mat(qt, rows, n, k, x, nth=2)
            assert np.all(np.abs(y - yr) <= 1e-3 * np.abs(yr) + 1e-5 * np.abs(yr).max())
            checked += 1
        assert checked == 2 + 6
    finally:
        L.b200q4_model_free(m)
