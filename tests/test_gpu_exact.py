"""Exact mode (b200q4_set_exact / B200Q4_EXACT=1): the CUDA path computes the reference's own fp32 rounding sequence
(ggml_vec_dot_q4_0/_q4_1, ggml.c:484-510 / 525-552 == imax.c:1210-1228), so Y must equal the oracle -- and therefore
the compiled reference -- BIT FOR BIT, for any shape, any M, weights with every nibble value.  Needs a B200."""
import ctypes as C

import numpy as np
import pytest

from oracle import Q4_0, Q4_1, row_bytes

pytestmark = pytest.mark.gpu
TYPES = [(Q4_0, "q4_0"), (Q4_1, "q4_1")]


@pytest.fixture(scope="module")
def gpu():
    import vsim_b200
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    L.b200q4_set_exact.restype = C.c_int
    L.b200q4_set_exact.argtypes = [C.c_int]
    old = L.b200q4_set_exact(1)
    yield vsim_b200
    L.b200q4_set_exact(old)


def random_nibble_weights(rng, qt, n, k):
    """every code 0..15 (the reference's quantizer never emits Q4_0 code 0, a model file may), positive and negative scales"""
    nb = k // 32
    if qt == Q4_0:
        raw = rng.integers(0, 256, size=(n, nb, 20), dtype=np.uint8)
        raw[:, :, :4] = (rng.standard_normal((n, nb)) * 0.01).astype(np.float32).view(np.uint8).reshape(n, nb, 4)
        return np.ascontiguousarray(raw.reshape(n, -1))
    raw = rng.integers(0, 256, size=(n, nb * 24), dtype=np.uint8)
    raw[:, : nb * 8] = (rng.standard_normal((n, nb * 2)) * 0.01).astype(np.float32).view(np.uint8).reshape(n, -1)
    return raw


@pytest.mark.parametrize("qt,tag", TYPES)
@pytest.mark.parametrize("n,k,m", [(16, 32, 1), (17, 96, 2), (1, 64, 1), (50, 160, 8), (333, 2048, 5), (4096, 4096, 1),
                                   (1000, 5120, 9), (130, 4128, 3), (257, 16384, 4), (64, 24576, 1), (300, 512, 17), (2048, 4096, 64)])
def test_exact_mode_is_bit_identical_to_the_oracle(gpu, oracle, qt, tag, n, k, m):
    rng = np.random.default_rng(n * 7 + k + m)
    wq = random_nibble_weights(rng, qt, n, k)
    x = rng.standard_normal((m, k)).astype(np.float32)
    x[0, :32] = 0.0                                        # an all-zero block: codes 8 / 0, scale 0
    w = gpu.Q4Weight(qt, wq, n, k)
    assert gpu.kernel_for(qt, m, n, k) == f"{tag}_exact"
    y = gpu.mul_mat(w, x)
    yr = oracle.mul_mat(qt, wq, n, k, x, nth=8)
    assert np.array_equal(y.view(np.uint32), yr.view(np.uint32)), f"{tag} N={n} K={k} M={m}: {np.abs(y - yr).max()}"
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_exact_mode_quantizer_weights_and_wdata_entry(gpu, oracle, qt, tag):
    """weights from the reference quantizer; rows pre-quantized by the CPU INIT phase (params->wdata) give the same bits"""
    rng = np.random.default_rng(5)
    n, k, m = 777, 1024, 6
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = (rng.standard_normal((m, k)) * 3).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    yr = oracle.mul_mat(qt, wq, n, k, x, nth=4)
    assert np.array_equal(gpu.mul_mat(w, x).view(np.uint32), yr.view(np.uint32))
    assert np.array_equal(gpu.mul_mat_wdata(w, oracle.quantize_rows(qt, x), m).view(np.uint32), yr.view(np.uint32))
    w.free()


def test_exact_mode_full_size_microbench_shape(gpu, oracle):
    rng = np.random.default_rng(1234)
    n, k = 16384, 4096
    wq = oracle.quantize_rows(Q4_0, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = np.random.default_rng(5678).standard_normal((8, k)).astype(np.float32)
    w = gpu.Q4Weight(Q4_0, wq, n, k)
    y = gpu.mul_mat(w, x)
    assert np.array_equal(y.view(np.uint32), oracle.mul_mat(Q4_0, wq, n, k, x, nth=8).view(np.uint32))
    w.free()
