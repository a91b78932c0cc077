"""The tcgen05 tile kernel (q4_tile.cuh: M >= 9 by default) against the oracle.  Needs a B200.

Same bars as test_gpu_parity.py: every output within 1e-3 relative of the oracle (+1e-5*rowmax floor);
observed normwise error ~1e-6 -- the 32-element block sums are exact on the tensor core (e4m3 operands
holding small integers), only the fp32 scale products round.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import Q4_0, Q4_1

pytestmark = pytest.mark.gpu

TYPES = [(Q4_0, "q4_0"), (Q4_1, "q4_1")]
RTOL, FLOOR = 1e-3, 1e-5


@pytest.fixture(scope="module")
def gpu():
    import vsim_b200
    vsim_b200.init(0)
    vsim_b200.load_library().b200q4_set_exact(0)           # the tile kernel is a fast-mode kernel
    return vsim_b200


@pytest.fixture()
def force_tile(gpu):
    """Route every M >= 1 through the tile kernel for the duration of a test."""
    L = gpu.load_library()
    L.b200q4_debug_set_tile_min_m.restype = C.c_int
    L.b200q4_debug_set_tile_min_m.argtypes = [C.c_int]
    old = L.b200q4_debug_set_tile_min_m(1)
    yield
    L.b200q4_debug_set_tile_min_m(old)


def check(y, yr, what=""):
    assert y.shape == yr.shape
    rowmax = np.abs(yr).max(axis=1, keepdims=True)
    err = np.abs(y.astype(np.float64) - yr.astype(np.float64))
    bad = err > RTOL * np.abs(yr) + FLOOR * rowmax + 1e-30
    assert not bad.any(), f"{what}: {bad.sum()} outputs off, worst {err.max():.3e}"
    return float(np.linalg.norm(err) / (np.linalg.norm(yr.astype(np.float64)) + 1e-300))


def make(oracle, qt, n, k, m, seed, wscale=0.05, spread=0.0):
    rng = np.random.default_rng(seed)
    wq = oracle.quantize_rows(qt, rng.uniform(-wscale, wscale, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k))
    if spread:
        x = x * np.exp(rng.uniform(-spread, spread, (m, 1))) + rng.uniform(-2, 2, (m, 1))
    return wq, x.astype(np.float32)


# every tile width (16/32/64/128 rows), ragged N (not a multiple of 16 or 128), chunk counts that do
# not divide evenly among the stream-K pieces, tensors smaller and larger than one wave of CTAs
SHAPES = [(128, 128, 16), (16, 32, 9), (17, 96, 12), (200, 256, 9), (333, 2048, 17), (640, 1024, 33), (300, 512, 65),
          (1000, 4096, 64), (777, 1056, 100), (512, 4096, 128), (384, 1024, 300), (4096, 4096, 10), (130, 16384, 24),
          (20000, 256, 40)]


@pytest.mark.parametrize("qt,tag", TYPES)
@pytest.mark.parametrize("n,k,m", SHAPES)
def test_tile_vs_oracle(gpu, oracle, qt, tag, n, k, m):
    wq, x = make(oracle, qt, n, k, m, n * 131 + k * 7 + m, spread=2.0)
    assert gpu.kernel_for(qt, m, n, k).endswith("tile_tcgen05")
    w = gpu.Q4Weight(qt, wq, n, k)
    y = gpu.mul_mat(w, x)
    rel = check(y, oracle.mul_mat(qt, wq, n, k, x, nth=8), f"{tag} N={n} K={k} M={m}")
    assert rel < 1e-5, rel
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
@pytest.mark.parametrize("m", [1, 2, 3, 5, 8])
def test_tile_small_m_forced(gpu, oracle, force_tile, qt, tag, m):
    """M below the default threshold, forced through the tile kernel (padding columns, 16-row tiles)."""
    n, k = 700, 1280
    wq, x = make(oracle, qt, n, k, m, 99 + m)
    w = gpu.Q4Weight(qt, wq, n, k)
    assert check(gpu.mul_mat(w, x), oracle.mul_mat(qt, wq, n, k, x, nth=4), f"{tag} forced M={m}") < 1e-5
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_tile_equals_decode_kernel_to_rounding(gpu, oracle, force_tile, qt, tag):
    """The two kernels consume the same codes and sum each block exactly: results agree to fp32 rounding."""
    n, k, m = 2048, 4096, 8
    wq, x = make(oracle, qt, n, k, m, 5)
    w = gpu.Q4Weight(qt, wq, n, k)
    L = gpu.load_library()
    y_tile = gpu.mul_mat(w, x)
    old = L.b200q4_debug_set_tile_min_m(0)
    y_dec = gpu.mul_mat(w, x)
    L.b200q4_debug_set_tile_min_m(old)
    assert np.abs(y_tile - y_dec).max() <= 6e-6 * np.abs(y_dec).max()
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_tile_deterministic_and_scale_exact(gpu, oracle, qt, tag):
    n, k, m = 3000, 2048, 40                 # 24 tiles x 16 chunks over 148 CTAs: most tiles are split into pieces
    wq, x = make(oracle, qt, n, k, m, 17)
    w = gpu.Q4Weight(qt, wq, n, k)
    y1, y2 = gpu.mul_mat(w, x), gpu.mul_mat(w, x)
    assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32))          # pieces are combined in a fixed order
    y4 = gpu.mul_mat(w, x * np.float32(4.0))
    assert np.array_equal((y1 * np.float32(4.0)).view(np.uint32), y4.view(np.uint32))   # power-of-two scaling is exact
    # a row's result does not depend on which other rows share its tile
    ya = gpu.mul_mat(w, x[7:23])
    assert np.array_equal(ya.view(np.uint32), y1[7:23].view(np.uint32))
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
@pytest.mark.parametrize("n,k,m", [(130, 16384, 24), (128, 8192, 9), (700, 4096, 70)])
def test_tile_stream_k_many_owners_repeatable(gpu, oracle, qt, tag, n, k, m):
    """Tiles whose chunks are spread over many CTAs (up to 74 owners per tile): the owner of the tile's first chunk combines
    the parked pieces in CTA order and re-arms the tile's counter -- five launches in a row give the same bits, and they
    are the oracle's values."""
    wq, x = make(oracle, qt, n, k, m, n + k + m)
    w = gpu.Q4Weight(qt, wq, n, k)
    ys = [gpu.mul_mat(w, x) for _ in range(5)]
    for y in ys[1:]:
        assert np.array_equal(y.view(np.uint32), ys[0].view(np.uint32))
    assert check(ys[0], oracle.mul_mat(qt, wq, n, k, x, nth=4), f"{tag} many owners N={n} K={k} M={m}") < 1e-5
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_tile_wdata_entry_bitwise(gpu, oracle, qt, tag):
    """Rows quantized by the reference's INIT phase (params->wdata) give the same bits as device quantization."""
    n, k, m = 500, 768, 70
    wq, x = make(oracle, qt, n, k, m, 23, spread=3.0)
    x[0, :32] = 0.0
    w = gpu.Q4Weight(qt, wq, n, k)
    y1 = gpu.mul_mat(w, x)
    y2 = gpu.mul_mat_wdata(w, oracle.quantize_rows(qt, x), m)
    assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32))
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_tile_strided_output(gpu, oracle, qt, tag):
    n, k, m = 100, 256, 20
    wq, x = make(oracle, qt, n, k, m, 3)
    ybig = np.full((m, n + 12), 7.0, np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    gpu.mul_mat(w, x, out=ybig[:, :n])
    check(ybig[:, :n], oracle.mul_mat(qt, wq, n, k, x))
    assert np.all(ybig[:, n:] == 7.0)
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_tile_extreme_scales_and_zero_blocks(gpu, oracle, qt, tag):
    """Blocks of zeros, constant blocks, 1e30 / 1e-30 magnitudes: the bf16 x3 scale split is exact for any finite fp32."""
    rng = np.random.default_rng(8)
    n, k, m = 256, 512, 16
    wf = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    wf[:, 32:64] = 0.0
    wf[5, :] *= np.float32(1e-20)
    wf[6, :] *= np.float32(1e15)
    wq = oracle.quantize_rows(qt, wf)
    x = rng.standard_normal((m, k)).astype(np.float32)
    x[1] *= np.float32(1e20); x[2] *= np.float32(1e-25); x[3, 64:96] = 3.25; x[4] = 0.0
    w = gpu.Q4Weight(qt, wq, n, k)
    y = gpu.mul_mat(w, x)
    yr = oracle.mul_mat(qt, wq, n, k, x)
    assert np.isfinite(y).all()
    check(y, yr, tag)
    w.free()


@pytest.mark.parametrize("qt,tag", TYPES)
def test_prefill_2048_tokens_sampled(gpu, oracle, qt, tag):
    """BASELINE configs[3] shape class: 2048-token prefill against one n_embd x n_embd weight; oracle on sampled rows."""
    rng = np.random.default_rng(2048)
    n, k, m = 4096, 4096, 2048
    wq = oracle.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    x = rng.standard_normal((m, k)).astype(np.float32)
    w = gpu.Q4Weight(qt, wq, n, k)
    y = gpu.mul_mat(w, x)
    rows = np.sort(rng.choice(n, 48, replace=False))
    ms = np.sort(rng.choice(m, 96, replace=False))
    yr = oracle.mul_mat(qt, wq[rows], len(rows), k, x[ms], nth=8)
    assert check(y[np.ix_(ms, rows)], yr, f"{tag} prefill") < 1e-5
    # size-independent property at full size: a row permutation of W permutes Y's columns (to fp32
    # rounding: a row that moves to another tile is summed over differently placed stream-K pieces)
    perm = rng.permutation(n)
    wp = gpu.Q4Weight(qt, wq[perm], n, k)
    yp = gpu.mul_mat(wp, x)
    assert np.abs(yp - y[:, perm]).max() <= 6e-6 * np.abs(y).max()
    w.free(); wp.free()
