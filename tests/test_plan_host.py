"""Host logic of the decode launch plan (no GPU): for any shape the tasks a launch hands out -- three static ones per
CTA, then per-class tickets -- cover every 16-row x 4-block unit of the weight tensor exactly once, K-slices are whole
quant-block octets, and everything fits the 227 KB of shared memory.  The enumeration below restates the formulas of
q4_gemv_pipe_kernel's producer (task_range / issue_piece) and of prefetch_next_launch."""
import ctypes as C

import numpy as np
import pytest

SHAPES = [(16, 32), (17, 96), (1, 64), (50, 160), (333, 2048), (4096, 4096), (16384, 4096), (50400, 4096), (4096, 16384),
          (6144, 6144), (24576, 6144), (6144, 24576), (5120, 5120), (5120, 20480), (130, 4128), (33, 4160), (257, 20480),
          (250880, 4096), (64, 65536), (2000, 1920), (7, 448)]


def plan(n, k, m, qt=0):
    import vsim_b200
    L = vsim_b200.load_library()
    L.b200q4_debug_plan.restype = C.c_longlong
    L.b200q4_debug_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    out = (C.c_int * 12)()
    smem = L.b200q4_debug_plan(qt, n, k, m, out)
    assert smem > 0, (n, k, m)
    keys = ("MP", "PP", "TS", "ntasks", "grid", "stages", "nbp", "gnbp", "b8", "m1", "C", "n_tiles")
    return dict(zip(keys, out)), smem


def task_span(p, b_cls_idx, t):
    """(first unit, unit count) of task t in the sequence of a CTA of class `cls` -- the producer's formulas"""
    cls = b_cls_idx
    C_, TS, PP = p["C"], p["TS"], p["PP"]
    U = p["n_tiles"] * C_
    if PP > 1:                                           # sliced: piece `cls` of tile t
        return t * C_ + cls * TS, min(TS, C_ - cls * TS)
    if C_ >= 16:                                         # one whole tile
        return t * C_, C_
    return t * TS, min(TS, U - t * TS)                   # several short tiles


@pytest.mark.parametrize("m", [1, 2, 5, 8])
@pytest.mark.parametrize("n,k", SHAPES)
def test_tasks_cover_every_unit_exactly_once(n, k, m):
    for qt in (0, 1):
        p, smem = plan(n, k, m, qt)
        assert smem <= 232448 and p["stages"] >= 2
        C_, TS, PP, grid = p["C"], p["TS"], p["PP"], p["grid"]
        U = p["n_tiles"] * C_
        assert 1 <= grid <= 148 and p["ntasks"] >= grid
        if PP > 1:
            assert TS % 2 == 0 and TS <= 32 and PP * TS >= C_ and (PP - 1) * TS < C_      # whole octets, one stage or less
            assert p["nbp"] == TS * 4 and p["gnbp"] == PP * TS * 4 and p["ntasks"] == p["n_tiles"] * PP
            assert grid >= PP                                                           # every class has a CTA
        seen = np.zeros(U, np.int32)
        limit = p["n_tiles"] if PP > 1 else p["ntasks"]
        for cls in range(PP):
            ncls = (grid - cls + PP - 1) // PP if PP > 1 else grid
            done = set()
            for idx in range(ncls):                      # static tasks of the class's CTAs
                for s in range(3):
                    t = idx + s * ncls
                    if t < limit:
                        done.add(t)
            done |= set(range(3 * ncls, limit))          # then tickets, in order
            assert done == set(range(limit))
            for t in done:
                u0, cnt = task_span(p, cls, t)
                assert cnt >= 1 and u0 >= 0 and u0 + cnt <= U
                seen[u0:u0 + cnt] += 1
        assert np.all(seen == 1), (n, k, m, qt, int((seen != 1).sum()))


def test_default_slicing_rule():
    assert plan(4096, 4096, 1)[0]["PP"] == 1 and plan(4096, 8192, 1)[0]["PP"] == 1          # M <= 2: whole tiles up to two stages
    assert plan(4096, 8192 + 128, 1)[0]["PP"] == 3 and plan(4096, 16384, 1)[0]["PP"] == 4
    assert plan(4096, 6144, 8)[0]["PP"] == 2 and plan(4096, 4096, 8)[0]["PP"] == 1          # 8-row kernel: from K > 4096
    p1, p8 = plan(16384, 4096, 1)[0], plan(16384, 4096, 8)[0]
    assert (p1["MP"], p1["m1"], p1["b8"], p1["stages"]) == (2, 1, 0, 5)                     # octet mode, compact codes, 5 stages
    assert (plan(16384, 4096, 2)[0]["b8"], p8["MP"], p8["stages"]) == (1, 8, 3)
