"""Parity of the CUDA path (through the C ABI) with the oracle.  Needs a B200.

Bars (north_star):
  * quantize_row_q4_0/_q4_1: byte-for-byte equal to the reference row image;
  * mul_mat: every output within 1e-3 relative of the oracle (plus a 1e-5*rowmax
    floor for outputs that cancel to ~0) -- TOL below; observed error is ~1e-6
    because the device sums each 32-element block exactly in int32.
"""
import numpy as np
import pytest

from oracle import Q4_0, Q4_1, row_bytes

pytestmark = pytest.mark.gpu

TYPES = [(Q4_0, "q4_0"), (Q4_1, "q4_1")]
RTOL = 1e-3          # north_star: <= 1e-3 relative on every output
FLOOR = 1e-5         # x max|y_ref| of the row: outputs that cancel to ~0


@pytest.fixture(scope="module")
def gpu():
    import vsim_b200
    vsim_b200.init(0)
    vsim_b200.load_library().b200q4_set_exact(0)           # these are the tests of the default (fast) kernels
    assert not vsim_b200.kernel_for(0, 1, 64, 64).endswith("_exact")
    yield vsim_b200
    # leave the context up: other modules share the process


def assert_close(y, yr, what=""):
    assert y.shape == yr.shape, (y.shape, yr.shape)
    if y.size == 0:
        return 0.0
    rowmax = np.abs(yr).max(axis=1, keepdims=True)
    tol = RTOL * np.abs(yr) + FLOOR * rowmax + 1e-30
    err = np.abs(y.astype(np.float64) - yr.astype(np.float64))
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} outputs off, worst {err.max():.3e} vs tol {tol[err.argmax() // y.shape[1], 0]:.3e}"
    nrm = np.linalg.norm(yr.astype(np.float64)) + 1e-300
    return float(np.linalg.norm(err) / nrm)


def exact64(oracle, qt, wq, n, k, x):
    """The mathematically exact operator in float64 from the quantized operands."""
    aq = oracle.quantize_rows(qt, x)
    wd = oracle.dequantize_rows(qt, wq, k).astype(np.float64)
    ad = oracle.dequantize_rows(qt, aq, k).astype(np.float64)
    return ad @ wd.T


# ---- quantizers: bit-exact -----------------------------------------------------
@pytest.mark.parametrize("k", [64, 256, 4096])
@pytest.mark.parametrize("qt,tag", TYPES)
def test_quantize_golden_bit_exact(gpu, golden, qt, tag, k):
    g = golden["quantize_kat"]
    got = gpu.quantize_rows(qt, g[f"x_k{k}"])
    want = g[f"{tag}_k{k}"]
    assert np.array_equal(got, want), f"{(got != want).sum()} bytes differ"


@pytest.mark.parametrize("qt,tag", TYPES)
def test_quantize_random_bit_exact_vs_oracle(gpu, oracle, qt, tag):
    rng = np.random.default_rng(31337)
    for k, rows in ((32, 3), (64, 5), (4096, 64), (16384, 9), (24576, 2)):
        x = (rng.standard_normal((rows, k)) * np.exp(rng.uniform(-12, 12, (rows, 1)))).astype(np.float32)
        x[0, :32] = 0.0
        x[-1, -32:] = -0.0
        assert np.array_equal(gpu.quantize_rows(qt, x), oracle.quantize_rows(qt, x)), (k, rows)


def test_quantize_signed_zero_min_q4_1(gpu, oracle):
    x = np.zeros((2, 64), np.float32)
    x[0, 3] = -0.0; x[0, 7] = 0.0; x[0, 10:20] = 1.5       # min is a zero; first zero seen is -0.0
    x[1, 0] = 0.0; x[1, 1] = -0.0; x[1, 40:] = 2.0
    assert np.array_equal(gpu.quantize_rows(Q4_1, x), oracle.quantize_rows(Q4_1, x))


def test_quantize_empty(gpu):
    assert gpu.quantize_rows(Q4_0, np.zeros((0, 64), np.float32)).shape == (0, 40)


# ---- mul_mat against golden vectors (reference outputs) ------------------------
@pytest.mark.parametrize("qt,tag", TYPES)
def test_mul_mat_golden(gpu, golden, qt, tag):
    g = golden["mul_mat_kat"]
    worst = 0.0
    n_cases = 0
    for key in g.files:
        if f"_{tag}_y_m" not in key:
            continue
        name = key.split("_")[0]
        m = int(key.rsplit("_m", 1)[1])
        n, k = (int(v) for v in g[f"{name}_shape"])
        w = gpu.Q4Weight(qt, g[f"{name}_{tag}_w"], n, k)
        y = gpu.mul_mat(w, g[f"{name}_x_m{m}"])
        worst = max(worst, assert_close(y, g[key], f"{name} {tag} M={m}"))
        w.free()
        n_cases += 1
    assert n_cases >= 10
    assert worst < 1e-5, worst        # normwise error is at rounding level, far inside the bar


@pytest.mark.parametrize("qt,tag", TYPES)
@pytest.mark.parametrize("n,k,m", [(16, 64, 1), (17, 96, 2), (1, 32, 1), (50, 160, 8), (333, 2048, 5),
                                   (4096, 4096, 1), (2048, 5120, 9), (512, 16384, 8), (256, 24576, 3),
                                   (640, 1024, 16), (300, 512, 17), (128, 256, 40),
                                   # K > 4096: K-sliced decode launches (2..6 piece classes, ragged last slice)
                                   (700, 5120, 1), (300, 6144, 2), (130, 4128, 1), (257, 20480, 4), (64, 8192, 8),
                                   (1000, 24576, 1), (33, 4160, 7), (2000, 16384, 1), (6144, 24576, 2)])
def test_mul_mat_random_vs_oracle(gpu, oracle, qt, tag, n, k, m):
    rng = np.random.default_rng(n * 131 + k * 7 + m)
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    y = gpu.mul_mat(w, x)
    yr = oracle.mul_mat(qt, wq, n, k, x, nth=8)
    rel = assert_close(y, yr, f"{tag} N={n} K={k} M={m}")
    # and closer to the exact result than the 1e-3 bar by orders of magnitude
    ex = exact64(oracle, qt, wq, n, k, x)
    assert np.linalg.norm(y - ex) / (np.linalg.norm(ex) + 1e-300) < 2e-6
    assert rel < 1e-5
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_mul_mat_strided_host_buffers(gpu, oracle, qt, tag):
    """src1 rows at nb11, dst rows at nb1 (ggml.c:4938-4949 strides are honoured)."""
    rng = np.random.default_rng(5)
    n, k, m = 100, 256, 6
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    xbig = rng.standard_normal((m, k + 24)).astype(np.float32)
    ybig = np.full((m, n + 12), 7.0, np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    gpu.mul_mat(w, xbig[:, :k], out=ybig[:, :n])
    assert_close(ybig[:, :n], oracle.mul_mat(qt, wq, n, k, np.ascontiguousarray(xbig[:, :k])))
    assert np.all(ybig[:, n:] == 7.0)                      # padding untouched
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_wdata_entry_equals_fp32_entry_bitwise(gpu, oracle, qt, tag):
    """CPU-quantized rows (params->wdata of an unpatched ggml.c) give the same bits as device
    quantization -- for the fused-quantization decode kernel (M <= 2) as well as the packed path:
    the codes the GEMV consumes are the oracle's codes."""
    rng = np.random.default_rng(77)
    for n, k, m in ((200, 768, 9), (200, 768, 1), (333, 4096, 2), (64, 64, 1), (48, 96, 2)):
        wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
        x = (rng.standard_normal((m, k)) * np.exp(rng.uniform(-6, 6, (m, 1)))).astype(np.float32)
        x[0, :32] = 0.0
        w = gpu.Q4Weight(qt, wq, n, k)
        y1 = gpu.mul_mat(w, x)
        y2 = gpu.mul_mat_wdata(w, oracle.quantize_rows(qt, x), m)
        if m == 1:
            # M = 1 with fp32 input runs the fused "octet" kernel, the wdata entry the two-row kernel:
            # same codes, different (fixed) summation trees -> equal to fp32 rounding.  One wrong
            # 4-bit code would move an output by ~1e-3 of the row scale.
            assert np.abs(y1 - y2).max() <= 4e-6 * np.abs(y2).max(), (n, k, m)
        else:
            assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32)), (n, k, m)
        w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_weight_retile_roundtrip(gpu, oracle, qt, tag):
    rng = np.random.default_rng(9)
    for n, k in ((16, 128), (37, 96), (1000, 4096)):
        wq = oracle.quantize_rows(qt, rng.standard_normal((n, k)).astype(np.float32))
        w = gpu.Q4Weight(qt, wq, n, k)
        assert np.array_equal(w.download(), wq)            # the device re-tiling loses nothing
        assert w.device_bytes >= wq.size
        w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_row_shards_concatenate_to_full(gpu, oracle, qt, tag):
    """Row-wise weight shards (multi-GPU split, imax.c:1185-1188) reproduce the full result exactly."""
    rng = np.random.default_rng(11)
    n, k, m = 192, 512, 4
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k)).astype(np.float32)
    full = gpu.Q4Weight(qt, wq, n, k)
    yf = gpu.mul_mat(full, x)
    parts = []
    for g in range(4):
        sh = gpu.Q4Weight(qt, wq, n, k, row0=g * 48, nrows=48)
        parts.append(gpu.mul_mat(sh, x))
        sh.free()
    assert_close(np.concatenate(parts, axis=1), yf)
    full.free()


def test_empty_and_degenerate_shapes(gpu, oracle):
    wq = oracle.quantize_rows(Q4_0, np.ones((8, 64), np.float32))
    w = gpu.Q4Weight(Q4_0, wq, 8, 64)
    assert gpu.mul_mat(w, np.zeros((0, 64), np.float32)).shape == (0, 8)
    y = gpu.mul_mat(w, np.zeros((3, 64), np.float32))
    assert np.all(y == 0)
    w.free()
    w0 = gpu.Q4Weight(Q4_0, np.zeros((0,), np.uint8), 0, 64)
    assert gpu.mul_mat(w0, np.ones((2, 64), np.float32)).shape == (2, 0)
    w0.free()


def test_bad_arguments_fail_loudly(gpu):
    with pytest.raises(ValueError):
        gpu.Q4Weight(Q4_0, np.zeros(10, np.uint8), 4, 64)
    with pytest.raises(gpu.B200Q4Error):
        gpu.Q4Weight(Q4_0, np.zeros(4 * 20, np.uint8), 4, 48 // 48 * 33)   # K % 32 != 0


@pytest.mark.parametrize("qt,tag", TYPES)
def test_run_to_run_determinism_and_batch_consistency(gpu, oracle, qt, tag):
    rng = np.random.default_rng(3)
    n, k = 4096, 4096
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((8, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    y8a = gpu.mul_mat(w, x)
    y8b = gpu.mul_mat(w, x)
    assert np.array_equal(y8a.view(np.uint32), y8b.view(np.uint32))        # fixed-order partial sums
    # a row inside a pair is independent of its partner: identical bits whatever the other row is
    y2a = gpu.mul_mat(w, x[2:4])
    y2b = gpu.mul_mat(w, np.stack([x[7], x[3]]))
    assert np.array_equal(y2a[1:2].view(np.uint32), y2b[1:2].view(np.uint32))
    # alone (octet kernel), in a pair (chain4) or in a batch of 8 (one IMMA per block) the
    # summation trees differ: equal to fp32 rounding
    y1 = gpu.mul_mat(w, x[3:4])
    y1b = gpu.mul_mat(w, x[3:4])
    assert np.array_equal(y1.view(np.uint32), y1b.view(np.uint32))
    for other in (y2a[1:2], y8a[3:4]):
        assert_close(y1, other)
        assert np.abs(y1 - other).max() <= 4e-6 * np.abs(other).max()
    # scaling the activations by a power of two scales the result exactly
    y2 = gpu.mul_mat(w, x * np.float32(4.0))
    assert np.array_equal((y8a * np.float32(4.0)).view(np.uint32), y2.view(np.uint32))
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
@pytest.mark.parametrize("n,k,m", [(1024, 16384, 1), (100, 24576, 1), (333, 8192, 2), (50, 160, 1), (4096, 4096, 1),
                                   (77, 96, 2), (2048, 16384, 2), (16, 32, 1)])
def test_pack_kernel_path_equals_fused_prologue(gpu, oracle, qt, tag, n, k, m):
    """Decode, M <= 2: activations quantized once by the pack kernel (pre-quantized / unaligned inputs, or the
    B200Q4_FUSE_MAX_K knob) or inside the CTAs' prologues (default) -- the codes and the summation order are the
    same, so the bits are.  K > 4096 also exercises K-slicing (piece classes, parked pieces, batched arrival)."""
    import ctypes as C
    L = gpu.load_library()
    L.b200q4_debug_set_fuse_max_k.restype = C.c_int
    L.b200q4_debug_set_fuse_max_k.argtypes = [C.c_int]
    rng = np.random.default_rng(n + k + m)
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    old = L.b200q4_debug_set_fuse_max_k(0)                 # always the pack kernel
    try:
        y_pack = gpu.mul_mat(w, x)
        L.b200q4_debug_set_fuse_max_k(1 << 30)             # always the fused prologue
        y_fused = gpu.mul_mat(w, x)
    finally:
        L.b200q4_debug_set_fuse_max_k(old)
    y_default = gpu.mul_mat(w, x)
    assert np.array_equal(y_pack.view(np.uint32), y_fused.view(np.uint32))
    assert np.array_equal(y_default.view(np.uint32), y_fused.view(np.uint32))
    assert_close(y_pack, oracle.mul_mat(qt, wq, n, k, x, nth=8), f"{tag} N={n} K={k} M={m} pack path")
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_k_sliced_many_pieces_per_cta(gpu, oracle, qt, tag):
    """K = 12288 (three K-slices), N = 30000: each CTA parks ~38 pieces for the finalize kernel; twice on the same
    tensor (workspace slots and ticket counters are reused)."""
    rng = np.random.default_rng(99)
    n, k = 30000, 12288
    wf = rng.random((n, k), dtype=np.float32)
    wf -= np.float32(0.5)
    wf *= np.float32(0.1)
    wq = oracle.quantize_rows(qt, wf)
    del wf
    x = rng.standard_normal((1, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    y = gpu.mul_mat(w, x)
    y2 = gpu.mul_mat(w, x)
    assert np.array_equal(y.view(np.uint32), y2.view(np.uint32))
    assert assert_close(y, oracle.mul_mat(qt, wq, n, k, x, nth=8), f"{tag} N={n} K={k}") < 1e-5
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
@pytest.mark.parametrize("n,k,m", [(4096, 4096, 1), (1000, 16384, 1), (333, 2048, 2), (512, 8192, 8), (100, 256, 5)])
def test_zero_copy_host_path_equals_copy_path(gpu, oracle, qt, tag, n, k, m):
    """b200q4_mul_mat on PINNED host buffers (b200q4_host_alloc): X is pulled by a fetch kernel, Y is stored by the
    decode epilogue straight into the caller's (strided) buffer, completion is a polled host word.  Same bits as
    the staged-copy path that pageable buffers take; bytes outside the Y rows stay untouched."""
    import ctypes as C
    L = gpu.load_library()
    rng = np.random.default_rng(n + 3 * k + m)
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    y_copy = gpu.mul_mat(w, x)                             # numpy arrays are pageable: staged copies
    pad = 8                                                # y rows n + pad floats apart
    L.b200q4_host_alloc.restype = C.c_void_p
    hx = L.b200q4_host_alloc(m * k * 4)
    hy = L.b200q4_host_alloc(m * (n + pad) * 4)
    try:
        C.memmove(hx, x.ctypes.data, m * k * 4)
        ybuf = np.ctypeslib.as_array(C.cast(hy, C.POINTER(C.c_float)), shape=(m, n + pad))
        ybuf[:] = 7.0
        for _ in range(3):                                 # the completion word counts up
            rc = L.b200q4_mul_mat(w.handle, C.c_void_p(hx), m, k * 4, C.c_void_p(hy), (n + pad) * 4)
            assert rc == 0, gpu.lib._err(L)
        assert np.array_equal(ybuf[:, :n].view(np.uint32), y_copy.view(np.uint32))
        assert np.all(ybuf[:, n:] == 7.0)
        assert_close(ybuf[:, :n].copy(), oracle.mul_mat(qt, wq, n, k, x, nth=8), f"{tag} zero-copy N={n} K={k} M={m}")
    finally:
        L.b200q4_host_free(C.c_void_p(hx))
        L.b200q4_host_free(C.c_void_p(hy))
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
@pytest.mark.parametrize("n,k,m", [(700, 5120, 1), (300, 6144, 2), (130, 4128, 1), (64, 8192, 8), (1000, 8192, 1), (50, 4160, 5)])
def test_k_slicing_forced_on_short_rows(gpu, oracle, qt, tag, n, k, m):
    """By default rows of K <= 8192 (M <= 2) stay whole; forcing K-slicing from K > 2048 on must give the same
    result up to the rounding of the piece-wise sums (and the oracle's bar either way)."""
    import ctypes as C
    L = gpu.load_library()
    L.b200q4_debug_set_slice_min_c.restype = C.c_int
    L.b200q4_debug_set_slice_min_c.argtypes = [C.c_int]
    rng = np.random.default_rng(n + k + m)
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    y_default = gpu.mul_mat(w, x)
    old = L.b200q4_debug_set_slice_min_c(16)
    try:
        y_sliced = gpu.mul_mat(w, x)
    finally:
        L.b200q4_debug_set_slice_min_c(old)
    yr = oracle.mul_mat(qt, wq, n, k, x, nth=8)
    assert_close(y_default, yr, f"{tag} default")
    assert_close(y_sliced, yr, f"{tag} sliced")
    assert np.abs(y_sliced - y_default).max() <= 4e-6 * np.abs(yr).max()
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_next_weight_hint_changes_nothing_but_timing(gpu, oracle, qt, tag):
    """b200q4_hint_next_weight is a pure performance hint: same bits with no hint, the right hint, a hint for a tensor of
    another shape (K-sliced), another type, a batch size the tile kernel serves, and a hinted tensor freed before the call."""
    L = gpu.load_library()
    rng = np.random.default_rng(12)
    n, k = 1500, 4096
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((2, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    others = [gpu.Q4Weight(qt, oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (300, 16384)).astype(np.float32)), 300, 16384),
              gpu.Q4Weight(1 - qt, oracle.quantize_rows(1 - qt, rng.uniform(-0.05, 0.05, (40, 64)).astype(np.float32)), 40, 64)]
    base = [gpu.mul_mat(w, x[:m]) for m in (1, 2)]
    for hint, hm in ((w, 1), (others[0], 1), (others[0], 8), (others[1], 2), (w, 64)):
        for i, m in enumerate((1, 2)):
            assert L.b200q4_hint_next_weight(hint.handle, hm) == 0
            assert np.array_equal(gpu.mul_mat(w, x[:m]).view(np.uint32), base[i].view(np.uint32))
    gone = gpu.Q4Weight(qt, wq[:64], 64, k)
    assert L.b200q4_hint_next_weight(gone.handle, 1) == 0
    gone.free()                                            # the library drops a hint that points at a freed tensor
    assert np.array_equal(gpu.mul_mat(w, x[:1]).view(np.uint32), base[0].view(np.uint32))
    assert L.b200q4_hint_next_weight(None, 1) == 0
    for o in others:
        o.free()
    w.free()


# ---- offline weight quantizers on the device (row f-4) ---------------------------------------------
@pytest.mark.parametrize("k", [64, 256])
@pytest.mark.parametrize("qt,tag", TYPES)
def test_weight_quantizer_golden_bit_exact(gpu, golden, qt, tag, k):
    g = golden["quantize_weights_kat"]
    rows, hist = gpu.quantize_weights(qt, g[f"x_k{k}"])
    assert np.array_equal(rows, g[f"{tag}_k{k}"])
    assert np.array_equal(hist, g[f"{tag}_hist_k{k}"])


@pytest.mark.parametrize("qt,tag", TYPES)
def test_weight_quantizer_random_vs_oracle_incl_strided_input(gpu, qt, tag):
    from oracle import quantize_weights
    rng = np.random.default_rng(8)
    big = (rng.standard_normal((3000, 4096 + 64)) * 0.05).astype(np.float32)
    big[7] = 0.0
    big[8] = -np.abs(big[8]) - 1.0
    x = big[:, :4096]                                      # rows 4160 floats apart
    rows, hist = gpu.quantize_weights(qt, x)
    want, whist = quantize_weights(qt, np.ascontiguousarray(x))
    assert np.array_equal(rows, want)
    assert np.array_equal(hist, whist) and hist.sum() == x.size


# ---- BASELINE.json microbench shape, full size -----------------------------------
@pytest.mark.parametrize("qt,tag", TYPES)
@pytest.mark.parametrize("m", [1, 8])
def test_microbench_shape_full_size_vs_oracle(gpu, oracle, qt, tag, m):
    rng = np.random.default_rng(1234)
    n, k = 16384, 4096
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = np.random.default_rng(5678).standard_normal((m, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    y = gpu.mul_mat(w, x)
    yr = oracle.mul_mat(qt, wq, n, k, x, nth=8)
    assert assert_close(y, yr, f"{tag} microbench M={m}") < 1e-5
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_microbench_shape_m64_sampled_rows(gpu, oracle, qt, tag):
    """M=64 at full N,K: oracle on a 512-row sample of W (whole check would take ~10 s of CPU)."""
    rng = np.random.default_rng(4321)
    n, k, m = 16384, 4096, 64
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    y = gpu.mul_mat(w, x)
    rows = np.sort(rng.choice(n, 512, replace=False))
    yr = oracle.mul_mat(qt, wq[rows], 512, k, x, nth=8)
    assert assert_close(y[:, rows], yr, f"{tag} M=64 sampled") < 1e-5
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_microbench_shape_m512_sampled_rows(gpu, oracle, qt, tag):
    """The M=512 microbench point (BASELINE configs[1], tcgen05 tile kernel) at full N, K: oracle on a 128-row sample of
    W that hits every 128-row MMA tile position (first / last rows of tiles, the tensor's last tile), all 512 columns."""
    rng = np.random.default_rng(99)
    n, k, m = 16384, 4096, 512
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    y = gpu.mul_mat(w, x)
    rows = np.unique(np.concatenate([rng.choice(n, 96, replace=False), [0, 15, 16, 127, 128, 8191, 8192, n - 129, n - 128, n - 17, n - 16, n - 1]]))
    yr = oracle.mul_mat(qt, wq[rows], len(rows), k, x, nth=8)
    assert assert_close(y[:, rows], yr, f"{tag} M=512 sampled") < 1e-5
    # size-independent properties at the full size: a repeated call gives the same bits (fixed summation order); rows of X
    # that are equal give equal rows of Y up to rounding -- rows 0 and 511 sit in different M-tiles, whose stream-K pieces
    # are added in a different order, so the last bits may differ -- and changing one row of X changes no other row of Y
    x2 = x.copy(); x2[511] = x2[0]
    y2 = gpu.mul_mat(w, x2)
    assert np.array_equal(gpu.mul_mat(w, x2).view(np.uint32), y2.view(np.uint32))
    assert np.abs(y2[511] - y2[0]).max() <= 1e-5 * np.abs(y2[0]).max()
    assert np.array_equal(y2[1:511].view(np.uint32), y[1:511].view(np.uint32))
    w.free()


# ---- transposed-src0 branch (row f-2) ----------------------------------------------
@pytest.mark.parametrize("qt,tag", TYPES)
def test_transposed_branch_golden_bit_exact(gpu, golden, qt, tag):
    g = golden["mul_mat_transposed_kat"]
    n, k = (int(v) for v in g["shape"])
    wt = gpu.Q4Weight(qt, g[f"{tag}_wt"], n, k, transposed=True)
    for m in (1, 5):
        y = gpu.mul_mat_transposed(wt, g[f"x_m{m}"])
        # same k-order accumulation as a single-thread reference run -> identical bits
        assert np.array_equal(y.view(np.uint32), g[f"{tag}_y_m{m}_t1"].view(np.uint32))
        assert_close(y, g[f"{tag}_y_m{m}_t3"])             # 3-thread reference differs only in rounding
    wt.free()
