"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/b200q4.h declares (and the reference's hook names), refuses to
compute without a GPU, and the ggml struct mirror matches the reference's header."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import vsim_b200
from vsim_b200 import lib as vlib

ROOT = Path(__file__).resolve().parents[1]


def _declared_in_header():
    names = set()
    for hdr in ("b200q4.h", "b200neox.h", "b200q4_dev.h"):   # every header of include/ (the last one: bench / test helpers)
        text = (ROOT / "include" / hdr).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(b200(?:q4|neox)_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_symbols_all_exported():
    L = vsim_b200.load_library()
    declared = _declared_in_header()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in b200q4.h but not exported by libb200q4.so"
    assert sorted(vlib.ABI_SYMBOLS) == declared


def test_reference_hook_symbols_exported():
    L = vsim_b200.load_library()
    for name in vlib.HOOK_SYMBOLS:       # imax.c:52, imax.c:1133 + the new Q4_1 twin
        assert hasattr(L, name), name


def test_abi_version_and_kernel_names():
    L = vsim_b200.load_library()
    assert L.b200q4_abi_version() == 1
    assert vsim_b200.kernel_for(vsim_b200.Q4_0, 1, 16384, 4096) == "q4_0_gemv_pipe_chain4"
    assert vsim_b200.kernel_for(vsim_b200.Q4_1, 8, 16384, 4096) == "q4_1_gemv_pipe_m8"
    assert vsim_b200.kernel_for(vsim_b200.Q4_1, 16, 16384, 4096) == "q4_1_tile_tcgen05"
    assert vsim_b200.kernel_for(vsim_b200.Q4_0, 512, 16384, 4096) == "q4_0_tile_tcgen05"
    assert vsim_b200.kernel_for(7, 1, 16, 64) == "invalid"


def test_no_cpu_fallback_without_gpu():
    """Without a usable sm_100 device every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU boxes")
    L = vsim_b200.load_library()
    assert L.b200q4_init(0) == -2                       # B200Q4_ENODEV
    assert b"no CPU fallback" in L.b200q4_last_error()
    x = np.zeros((1, 64), np.float32)
    out = np.zeros((1, 40), np.uint8)
    assert L.b200q4_quantize_rows(0, x.ctypes.data, 1, 64, 256, out.ctypes.data) == -2
    assert not L.b200q4_weight_upload(0, out.ctypes.data, 1, 64, 40)
    with pytest.raises(vsim_b200.B200Q4Error):
        vsim_b200.quantize_rows(vsim_b200.Q4_0, x)


def test_ggml_struct_mirror_matches_reference(reference):
    """include/b200q4_ggml_abi.h vs the real struct ggml_tensor (ggml.h:275-305)."""
    size, o_type, o_ndims, o_ne, o_nb, o_src0, o_src1, o_data = reference.tensor_layout()

    class T(C.Structure):
        _fields_ = [("type", C.c_int), ("n_dims", C.c_int), ("ne", C.c_int * 4), ("nb", C.c_size_t * 4),
                    ("op", C.c_int), ("is_param", C.c_bool), ("grad", C.c_void_p), ("src0", C.c_void_p),
                    ("src1", C.c_void_p), ("opt", C.c_void_p * 4), ("n_tasks", C.c_int), ("perf_runs", C.c_int),
                    ("perf_cycles", C.c_int64), ("perf_time_us", C.c_int64), ("data", C.c_void_p),
                    ("padding", C.c_char * 8)]
    assert C.sizeof(T) == size == 160
    assert (T.type.offset, T.n_dims.offset, T.ne.offset, T.nb.offset) == (o_type, o_ndims, o_ne, o_nb)
    assert (T.src0.offset, T.src1.offset, T.data.offset) == (o_src0, o_src1, o_data)
    # and the C mirror itself states the same layout
    hdr = (ROOT / "include" / "b200q4_ggml_abi.h").read_text()
    for field in ("int     type;", "int     ne[B200Q4_GGML_MAX_DIMS];", "size_t  nb[B200Q4_GGML_MAX_DIMS];", "void   *data;"):
        assert field in hdr


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (tier rule 3)."""
    for p in (ROOT / "vsim_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".c", ".h"}:
            txt = p.read_text()
            assert "oracle" not in txt.replace("no CPU fallback", ""), f"{p} mentions the oracle"
