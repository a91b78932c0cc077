"""Decode program kernel (q4_prog_kernel): chains of dependent M = 1 mul_mats in ONE persistent launch must give the
bits of the per-op decode kernel (b200q4_mul_mat_dev, M = 1), which tests/test_gpu_parity.py pins to the oracle; and
the oracle directly on small shapes.  Shapes cover whole tiles, short tiles (K < 2048), ragged N, K-sliced rows
(K > 8192) feeding a dependent op, and both quant types."""
import numpy as np
import pytest

from oracle import Oracle, Q4_0, Q4_1

pytestmark = pytest.mark.gpu


def _setup():
    import torch
    import vsim_b200
    vsim_b200.init(0)
    L = vsim_b200.load_library()
    L.b200q4_set_exact(0)
    return torch, vsim_b200, L


def _rand_weight(vs, orc, rng, qt, n, k):
    wq = orc.quantize_rows(qt, rng.uniform(-0.05, 0.05, (n, k)).astype(np.float32))
    return wq, vs.Q4Weight(qt, wq, n, k)


def _per_op(L, torch, w, x, n):
    y = torch.empty(n, device="cuda", dtype=torch.float32)
    rc = L.b200q4_mul_mat_dev(w.handle, x.data_ptr(), 1, 0, y.data_ptr(), 0, None)
    assert rc == 0
    L.b200q4_sync(None)
    return y


@pytest.mark.parametrize("qt", [Q4_0, Q4_1])
@pytest.mark.parametrize("shapes", [
    [(1024, 2048)],                                  # one op, whole tiles
    [(100, 128)],                                    # short tiles, ragged N
    [(4096, 4096), (8192, 4096), (4096, 8192)],      # a dependent chain, two-stage tiles
    [(2048, 1024), (12288, 2048), (1000, 12288), (320, 992)],   # K-sliced op (K = 12288) in the middle, ragged everything
    [(16400, 16384), (64, 16384)],                   # sliced op whose input is a sliced op's output
])
def test_chain_equals_per_op_kernel(qt, shapes):
    torch, vs, L = _setup()
    orc = Oracle()
    rng = np.random.default_rng(len(shapes) * 7 + qt)
    ws, hosts = [], []
    for n, k in shapes:
        wq, w = _rand_weight(vs, orc, rng, qt, n, k)
        ws.append(w); hosts.append(wq)
    x0 = torch.from_numpy(rng.standard_normal(shapes[0][1]).astype(np.float32)).cuda()
    # reference run: one launch per op; op i reads the first K_i outputs of op i-1
    ref, x = [], x0
    for (n, k), w in zip(shapes, ws):
        y = _per_op(L, torch, w, x, n)
        ref.append(y)
        x = y
    ys = [torch.zeros(n + 8, device="cuda", dtype=torch.float32) for n, _ in shapes]
    xs = [x0.data_ptr()] + [ys[i - 1].data_ptr() for i in range(1, len(shapes))]
    chain = vs.Chain(ws, xs, [y.data_ptr() for y in ys], [-1] + list(range(len(shapes) - 1)))
    for rep in range(3):                             # the counters clear themselves: every run is the first run
        for y in ys:
            y.zero_()
        chain.run()
        L.b200q4_sync(None)
        for i, ((n, k), y) in enumerate(zip(shapes, ys)):
            assert torch.equal(y[:n].view(torch.int32), ref[i].view(torch.int32)), (rep, i, (y[:n] - ref[i]).abs().max().item())
            assert float(y[n:].abs().max()) == 0.0   # nothing written past N
    # and against the oracle itself on the first op
    n, k = shapes[0]
    if n * k <= 1 << 23:
        yo = orc.mul_mat(qt, hosts[0], n, k, x0.cpu().numpy()[None, :], nth=1)[0]
        yd = ys[0][:n].cpu().numpy()
        assert np.all(np.abs(yd - yo) <= 1e-3 * np.abs(yo) + 1e-5 * np.abs(yo).max())
    chain.free()
    for w in ws:
        w.free()


def test_chain_independent_ops_and_repeated_weights():
    """Q/K/V-style: three ops reading the same X with no dependency between them, then one that needs all three."""
    torch, vs, L = _setup()
    orc = Oracle()
    rng = np.random.default_rng(5)
    E = 2048
    ws = [_rand_weight(vs, orc, rng, Q4_0, E, E)[1] for _ in range(3)] + [_rand_weight(vs, orc, rng, Q4_0, 512, 3 * E)[1]]
    x0 = torch.from_numpy(rng.standard_normal(E).astype(np.float32)).cuda()
    qkv = torch.zeros(3 * E, device="cuda", dtype=torch.float32)
    out = torch.zeros(512, device="cuda", dtype=torch.float32)
    ref = torch.cat([_per_op(L, torch, ws[i], x0, E) for i in range(3)])
    ref_out = _per_op(L, torch, ws[3], ref, 512)
    chain = vs.Chain(ws, [x0.data_ptr()] * 3 + [qkv.data_ptr()],
                     [qkv.data_ptr() + 4 * E * i for i in range(3)] + [out.data_ptr()], [-1, -1, -1, 2])
    for _ in range(2):
        chain.run()
        L.b200q4_sync(None)
        assert torch.equal(qkv.view(torch.int32), ref.view(torch.int32))
        assert torch.equal(out.view(torch.int32), ref_out.view(torch.int32))
    chain.free()
    for w in ws:
        w.free()


@pytest.mark.parametrize("qt", [Q4_0, Q4_1])
def test_chain_independent_steps_with_their_own_inputs(qt):
    """bench.py's headline shape of list: independent ops, every one with its OWN plain input vector -- the activation packs
    are prepared ahead of time, between two stages of the previous op (no pause; no barrier where every stage covers the
    same 128 quant blocks: K = 4096 and K-slices of 4096).  Mixed with geometries that do need the barrier (K = 2048,
    ragged K = 992), a K-sliced op, a dependent op in the middle, ops sharing an input, and tiny ops that have no stage
    at all on most thread blocks.  Every output equals the per-call kernel's bit for bit, on every repetition."""
    torch, vs, L = _setup()
    orc = Oracle()
    rng = np.random.default_rng(31 + qt)
    shapes = [(4096, 4096), (4096, 4096), (1024, 2048), (4096, 4096), (96, 992), (2048, 16384), (4096, 4096), (32, 64), (16384, 4096),
              (1024, 2048), (1024, 2048)]
    ws = [_rand_weight(vs, orc, rng, qt, n, k)[1] for n, k in shapes]
    xs = [torch.from_numpy(rng.standard_normal(k).astype(np.float32)).cuda() for _, k in shapes]
    xs[10] = xs[9]                                       # the last two share an input (one pack)
    ys = [torch.zeros(n, device="cuda", dtype=torch.float32) for n, _ in shapes]
    xptr = [x.data_ptr() for x in xs]
    xptr[3] = ys[1].data_ptr()                           # op 3 reads op 1's output: a true dependency between the independent ones
    deps = [-1] * len(shapes)
    ref = []
    for i, ((n, k), w) in enumerate(zip(shapes, ws)):
        xin = ref[1] if i == 3 else xs[i]
        ref.append(_per_op(L, torch, w, xin, n))
    chain = vs.Chain(ws, xptr, [y.data_ptr() for y in ys], deps)
    for rep in range(3):
        for y in ys:
            y.zero_()
        chain.run()
        L.b200q4_sync(None)
        for i, y in enumerate(ys):
            assert torch.equal(y.view(torch.int32), ref[i].view(torch.int32)), (rep, i, shapes[i], (y - ref[i]).abs().max().item())
    chain.free()
    for w in ws:
        w.free()


def test_chain_rejects_bad_arguments():
    torch, vs, L = _setup()
    orc = Oracle()
    rng = np.random.default_rng(6)
    w0 = _rand_weight(vs, orc, rng, Q4_0, 64, 64)[1]
    w1 = _rand_weight(vs, orc, rng, Q4_1, 64, 64)[1]
    x = torch.zeros(64, device="cuda"); y = torch.zeros(64, device="cuda")
    with pytest.raises(vs.B200Q4Error):              # mixed quant types
        vs.Chain([w0, w1], [x.data_ptr()] * 2, [y.data_ptr()] * 2, [-1, 0])
    with pytest.raises(vs.B200Q4Error):              # forward dependency
        vs.Chain([w0, w0], [x.data_ptr()] * 2, [y.data_ptr()] * 2, [1, -1])
    with pytest.raises(vs.B200Q4Error):              # misaligned pointer
        vs.Chain([w0], [x.data_ptr() + 4], [y.data_ptr()], [-1])
    w0.free(); w1.free()
