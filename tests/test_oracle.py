"""The oracle (our C restatement) against the reference's own outputs.

* always: against tests/golden/*.npz (produced by the compiled reference,
  tests/golden/make_golden.py);
* when oracle/_ref/libggml_ref.so is present: live, byte for byte, on fresh
  seeded inputs, including the ggml_graph_compute path through the IMAX hook.
"""
import numpy as np
import pytest

from oracle import Q4_0, Q4_1, np_quantize_rows_q4_0, np_quantize_rows_q4_1, row_bytes

TYPES = [(Q4_0, "q4_0"), (Q4_1, "q4_1")]


@pytest.mark.parametrize("k", [64, 256, 4096])
@pytest.mark.parametrize("qt,tag", TYPES)
def test_quantize_matches_golden(oracle, golden, qt, tag, k):
    g = golden["quantize_kat"]
    x = g[f"x_k{k}"]
    want = g[f"{tag}_k{k}"]
    got = oracle.quantize_rows(qt, x)
    assert got.shape == want.shape == (x.shape[0], row_bytes(qt, k))
    assert np.array_equal(got, want)                       # bit-exact: scales and nibbles


@pytest.mark.parametrize("k", [64, 256, 4096])
def test_numpy_restatement_matches_golden(golden, k):
    g = golden["quantize_kat"]
    x = g[f"x_k{k}"]
    assert np.array_equal(np_quantize_rows_q4_0(x), g[f"q4_0_k{k}"])
    assert np.array_equal(np_quantize_rows_q4_1(x), g[f"q4_1_k{k}"])


def _mm_cases(g):
    for key in g.files:
        if "_y_m" in key:
            name, t0, t1, _, mtag = key.split("_")
            yield name, f"{t0}_{t1}", int(mtag[1:])


@pytest.mark.parametrize("qt,tag", TYPES)
def test_mul_mat_matches_golden_bit_exact(oracle, golden, qt, tag):
    g = golden["mul_mat_kat"]
    n_checked = 0
    for name, t, m in _mm_cases(g):
        if t != tag:
            continue
        n, k = (int(v) for v in g[f"{name}_shape"])
        y = oracle.mul_mat(qt, g[f"{name}_{tag}_w"], n, k, g[f"{name}_x_m{m}"], nth=1)
        want = g[f"{name}_{tag}_y_m{m}"]
        # same operation order as the reference -> identical fp32 bits
        assert np.array_equal(y.view(np.uint32), want.view(np.uint32)), (name, m)
        n_checked += 1
    assert n_checked >= 10


@pytest.mark.parametrize("qt,tag", TYPES)
def test_mul_mat_thread_count_invariant(oracle, golden, qt, tag):
    g = golden["mul_mat_kat"]
    n, k = (int(v) for v in g["a_shape"])
    x = g["a_x_m9"]
    y1 = oracle.mul_mat(qt, g[f"a_{tag}_w"], n, k, x, nth=1)
    for nth in (2, 3, 8):
        assert np.array_equal(y1, oracle.mul_mat(qt, g[f"a_{tag}_w"], n, k, x, nth=nth))


@pytest.mark.parametrize("qt,tag", TYPES)
def test_transposed_branch_matches_golden(oracle, golden, qt, tag):
    g = golden["mul_mat_transposed_kat"]
    n, k = (int(v) for v in g["shape"])
    for m in (1, 5):
        for nth in (1, 3):
            y = oracle.mul_mat_transposed(qt, g[f"{tag}_wt"], n, k, g[f"x_m{m}"], nth=nth)
            assert np.array_equal(y.view(np.uint32), g[f"{tag}_y_m{m}_t{nth}"].view(np.uint32))


@pytest.mark.parametrize("qt,tag", TYPES)
def test_wdata_is_the_quantized_activation_image(oracle, golden, qt, tag):
    g = golden["mul_mat_kat"]
    n, k = (int(v) for v in g["a_shape"])
    x = g["a_x_m5"]
    _, wd = oracle.mul_mat(qt, g[f"a_{tag}_w"], n, k, x, return_wdata=True)
    assert np.array_equal(wd, oracle.quantize_rows(qt, x))


def test_empty_inputs(oracle):
    w = np.zeros((0,), np.uint8)
    y = oracle.mul_mat(Q4_0, w, 0, 64, np.zeros((3, 64), np.float32))
    assert y.shape == (3, 0)
    wq = oracle.quantize_rows(Q4_0, np.ones((4, 64), np.float32))
    y = oracle.mul_mat(Q4_0, wq, 4, 64, np.zeros((0, 64), np.float32))
    assert y.shape == (0, 4)


def test_dequantize_roundtrip_error_bound(oracle):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((4, 256)).astype(np.float32)
    for qt in (Q4_0, Q4_1):
        q = oracle.quantize_rows(qt, x)
        xd = oracle.dequantize_rows(qt, q, 256)
        blk = np.abs(x).reshape(4, 8, 32).max(axis=2) if qt == Q4_0 else np.ptp(x.reshape(4, 8, 32), axis=2)
        step = blk / (7.0 if qt == Q4_0 else 15.0)
        err = np.abs(xd - x).reshape(4, 8, 32).max(axis=2)
        assert np.all(err <= 0.5 * step * 1.0001 + 1e-7)
        # quantize(dequantize(q)) is idempotent on the codes for q4_0 (scale is reproduced exactly
        # only when amax is hit; check codes through a second quantization of the dequantized row)
        q2 = oracle.quantize_rows(qt, xd)
        xd2 = oracle.dequantize_rows(qt, q2, 256)
        assert np.allclose(xd2, xd, rtol=0, atol=float(step.max()) * 0.51)


# ---- live against the compiled reference -------------------------------------
@pytest.mark.parametrize("qt", [Q4_0, Q4_1])
def test_quantize_live_vs_reference(oracle, reference, qt):
    rng = np.random.default_rng(99)
    for k in (64, 1024, 5120):
        x = (rng.standard_normal((16, k)) * np.exp(rng.uniform(-8, 8, (16, 1)))).astype(np.float32)
        x[3, :32] = 0
        assert np.array_equal(oracle.quantize_rows(qt, x), reference.quantize_rows(qt, x))
        q = reference.quantize_rows(qt, x)
        assert np.array_equal(oracle.dequantize_rows(qt, q, k).view(np.uint32),
                              reference.dequantize_rows(qt, q, k).view(np.uint32))


@pytest.mark.parametrize("qt", [Q4_0, Q4_1])
@pytest.mark.parametrize("m,nth", [(1, 1), (5, 3), (9, 8)])
def test_mul_mat_live_vs_reference(oracle, reference, qt, m, nth):
    rng = np.random.default_rng(4242 + m)
    n, k = 200, 768
    wq = reference.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k)).astype(np.float32)
    want = reference.mul_mat(qt, wq, n, k, x, nth=nth)      # ggml_graph_compute, IMAX hook for Q4_0
    got = oracle.mul_mat(qt, wq, n, k, x, nth=nth)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


# ---- offline weight quantizers (utils.cpp:424-541) -------------------------
@pytest.mark.parametrize("k", [64, 256])
@pytest.mark.parametrize("qt,tag", TYPES)
def test_weight_quantizer_matches_golden(golden, qt, tag, k):
    from oracle import quantize_weights
    g = golden["quantize_weights_kat"]
    rows, hist = quantize_weights(qt, g[f"x_k{k}"])
    assert np.array_equal(rows, g[f"{tag}_k{k}"])
    assert np.array_equal(hist, g[f"{tag}_hist_k{k}"])
    assert hist.sum() == g[f"x_k{k}"].size


def test_weight_quantizer_q4_1_keeps_the_numeric_limits_min_quirk(oracle):
    """utils.cpp:503 starts the running max at numeric_limits<float>::min() (smallest POSITIVE float): an all-negative
    block gets max = 1.17549435e-38, unlike quantize_row_q4_1 (ggml.c:269, -FLT_MAX)."""
    from oracle import quantize_weights
    x = -np.abs(np.random.default_rng(0).standard_normal((1, 32))).astype(np.float32) - 0.5
    w, _ = quantize_weights(Q4_1, x)
    a = oracle.quantize_rows(Q4_1, x)
    mn, d = w.view(np.float32)[0, 0], w.view(np.float32)[0, 1]
    assert mn == x.min() and d == np.float32((np.float32(1.17549435e-38) - x.min()) / np.float32(15.0))
    assert not np.array_equal(w, a)
    # Q4_0: the two quantizers are the same function
    assert np.array_equal(quantize_weights(Q4_0, x)[0], oracle.quantize_rows(Q4_0, x))


@pytest.mark.parametrize("qt,tag", TYPES)
def test_weight_quantizer_live_vs_reference(qt, tag):
    from oracle import REF_DIR, quantize_weights, reference_quantize_weights
    if not (REF_DIR / "libutils_ref.so").exists():
        pytest.skip("oracle/_ref/libutils_ref.so not built (needs /root/reference at build time)")
    rng = np.random.default_rng(77)
    x = (rng.standard_normal((64, 1024)) * np.exp(rng.uniform(-3, 3, (64, 1)))).astype(np.float32)
    x[5] = 0.0
    x[6] = -np.abs(x[6])
    a, ha = quantize_weights(qt, x)
    b, hb = reference_quantize_weights(qt, x)
    assert np.array_equal(a, b) and np.array_equal(ha, hb)
