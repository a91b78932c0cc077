"""Row f-1 groundwork (CPU): the restatement of the NON-hot ops of the reference's NeoX graph (oracle/neox_ops_oracle.c)
pinned BIT FOR BIT to the compiled reference -- one-op graphs through ggml_graph_compute in oracle/_ref/libggml_ref.so.
A device-resident graph executor will be checked against these functions."""
import ctypes as C

import numpy as np
import pytest

from oracle import HERE, REF_DIR, Q4_0, Q4_1


@pytest.fixture(scope="module")
def libs():
    so = REF_DIR / "libggml_ref.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libggml_ref.so not built (needs /root/reference at build time)")
    R = C.CDLL(str(so))
    O = C.CDLL(str(HERE / "libq4oracle.so"))
    R.ref_op.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    R.ref_get_rows.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    R.ggml_fp32_to_fp16.argtypes, R.ggml_fp32_to_fp16.restype = [C.c_float], C.c_uint16
    R.ggml_fp16_to_fp32.argtypes, R.ggml_fp16_to_fp32.restype = [C.c_uint16], C.c_float
    O.nxo_fp32_to_fp16.argtypes, O.nxo_fp32_to_fp16.restype = [C.c_float], C.c_uint16
    O.nxo_fp16_to_fp32.argtypes, O.nxo_fp16_to_fp32.restype = [C.c_uint16], C.c_float
    O.nxo_norm_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    O.nxo_gelu.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
    O.nxo_soft_max_rows.argtypes = [C.c_void_p, C.c_int, C.c_int]
    O.nxo_rope_neox.argtypes = [C.c_void_p] + [C.c_int] * 6
    O.nxo_diag_mask_inf.argtypes = [C.c_void_p] + [C.c_int] * 4
    O.nxo_get_rows_q4.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    R.ref_op(1, np.zeros(32, np.float32).ctypes.data, 32, 1, 1, 0, 0, 0, np.zeros(32, np.float32).ctypes.data, 1)  # ggml_init: tables
    return R, O


def ref_op(R, op, x, a=0, b=0, c=0, nth=1):
    x = np.ascontiguousarray(x, np.float32)
    shp = x.shape + (1,) * (3 - x.ndim)                    # numpy [ne2][ne1][ne0]
    x3 = x.reshape(shp if x.ndim == 3 else (1,) * (3 - x.ndim) + x.shape)
    ne2, ne1, ne0 = x3.shape
    y = np.empty_like(x3)
    assert R.ref_op(op, x3.ctypes.data, ne0, ne1, ne2, a, b, c, y.ctypes.data, nth) == 0
    return y.reshape(x.shape)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_fp16_conversions_all_halves_and_many_floats(libs):
    R, O = libs
    for h in range(1 << 16):
        a, b = R.ggml_fp16_to_fp32(h), O.nxo_fp16_to_fp32(h)
        assert (np.isnan(a) and np.isnan(b)) or np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), h
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(20000).astype(np.float32) * np.exp(rng.uniform(-20, 12, 20000)).astype(np.float32),
                         np.float32([0, -0.0, 65504, 65519.99, 65520, 70000, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8, 8.9e-8, 6.1e-5, 6.0975e-5,
                                     np.inf, -np.inf, 1.0009765, 1.00048828125, 1.00146484375])])
    # exact halfway cases between adjacent halves (ties to even)
    hs = rng.integers(0, 0x7BFF, 4000).astype(np.uint16)
    lo = np.array([R.ggml_fp16_to_fp32(int(h)) for h in hs], np.float64)
    hi = np.array([R.ggml_fp16_to_fp32(int(h) + 1) for h in hs], np.float64)
    xs = np.concatenate([xs, ((lo + hi) / 2).astype(np.float32), -((lo + hi) / 2).astype(np.float32)])
    for v in xs:
        assert R.ggml_fp32_to_fp16(float(v)) == O.nxo_fp32_to_fp16(float(v)), float(v)


@pytest.mark.parametrize("nth", [1, 3])
def test_norm_gelu_softmax_bit_exact(libs, nth):
    R, O = libs
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((7, 4096)) * 3).astype(np.float32)
    x[2] *= 1e-3
    x[3] += 100.0
    y = np.empty_like(x)
    O.nxo_norm_rows(x.ctypes.data, 7, 4096, y.ctypes.data)
    assert np.array_equal(bits(y), bits(ref_op(R, 0, x, nth=nth)))
    g = (rng.standard_normal((5, 16384)) * 4).astype(np.float32)
    g[0, :8] = [0, -0.0, 1e-9, -1e-9, 20, -20, 60000, -60000]      # beyond fp16 range the reference itself asserts
    yg = np.empty_like(g)
    O.nxo_gelu(g.ctypes.data, g.size, yg.ctypes.data)
    assert np.array_equal(bits(yg), bits(ref_op(R, 1, g, nth=nth)))
    s = (rng.standard_normal((6, 517)) * 5).astype(np.float32)
    s[1, 100:] = -np.inf                                   # what diag_mask_inf leaves behind
    s[2] = 3.0
    ys = s.copy()
    O.nxo_soft_max_rows(ys.ctypes.data, 6, 517)
    assert np.array_equal(bits(ys), bits(ref_op(R, 2, s, nth=nth)))
    assert np.allclose(ys.sum(axis=1), 1.0, atol=2e-3)


@pytest.mark.parametrize("mode,n_past", [(0, 0), (0, 7), (1, 0), (1, 5), (1, 11)])
def test_rope_neox_bit_exact(libs, mode, n_past):
    R, O = libs
    rng = np.random.default_rng(2)
    ne0, ne1, ne2, n_rot = 64, 4, 12, 32                   # d_head, heads, positions; rotary dims < d_head (NeoX: n_rot 24..64)
    x = rng.standard_normal((ne2, ne1, ne0)).astype(np.float32)
    y = x.copy()
    O.nxo_rope_neox(y.ctypes.data, ne0, ne1, ne2, n_past, n_rot, mode)
    yr = ref_op(R, 3, x, n_past, n_rot, mode)
    if mode == 1 and n_past > 0:                           # rows below n_past are not written by the op (in place in the graph)
        assert np.array_equal(bits(y[n_past:]), bits(yr[n_past:]))
        assert np.array_equal(y[:n_past], x[:n_past])
    else:
        assert np.array_equal(bits(y), bits(yr))
    assert np.array_equal(y[..., n_rot:], x[..., n_rot:])  # dims past n_rot pass through


def test_diag_mask_and_get_rows(libs, oracle):
    R, O = libs
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 5, 9)).astype(np.float32)  # [heads][N][n_past + N], n_past = 4
    y = x.copy()
    O.nxo_diag_mask_inf(y.ctypes.data, 9, 5, 3, 4)
    assert np.array_equal(bits(y), bits(ref_op(R, 4, x, 4)))
    for qt in (Q4_0, Q4_1):
        wq = oracle.quantize_rows(qt, rng.uniform(-1, 1, (50, 256)).astype(np.float32))
        ids = np.array([0, 49, 7, 7, 23], np.int32)
        a, b = np.empty((5, 256), np.float32), np.empty((5, 256), np.float32)
        O.nxo_get_rows_q4(qt, wq.ctypes.data, 256, ids.ctypes.data, 5, a.ctypes.data)
        assert R.ref_get_rows(qt, wq.ctypes.data, 50, 256, ids.ctypes.data, 5, b.ctypes.data) == 0
        assert np.array_equal(bits(a), bits(b))
