"""ctypes binding of include/b200q4.h (host buffers are numpy arrays).

Mirrors the C ABI one to one; argument meaning and error behaviour are those
of the header.  Every failure of the library surfaces as B200Q4Error carrying
b200q4_last_error(); nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

Q4_0, Q4_1 = 0, 1
QK = 32
_LIB = None
_INIT_DEVICE = None

# every symbol include/b200q4.h declares (tests/test_abi.py checks header <-> this list <-> the .so)
ABI_SYMBOLS = [
    "b200q4_abi_version", "b200q4_init", "b200q4_shutdown", "b200q4_last_error", "b200q4_device_info",
    "b200q4_weight_upload", "b200q4_weight_upload_rows", "b200q4_weight_free", "b200q4_weight_rows",
    "b200q4_weight_cols", "b200q4_weight_type", "b200q4_weight_device_bytes", "b200q4_weight_download",
    "b200q4_mul_mat", "b200q4_mul_mat_wdata", "b200q4_quantize_rows", "b200q4_weight_upload_transposed",
    "b200q4_mul_mat_transposed", "b200q4_mul_mat_dev", "b200q4_quantize_rows_dev", "b200q4_sync",
    "b200q4_mul_mat_dev_allgather", "b200q4_wait_flags", "b200q4_set_exact", "b200q4_exact_enabled", "b200q4_quantize_weights", "b200q4_hint_next_weight",
    "b200q4_model_load", "b200q4_model_header", "b200q4_model_tensor_count", "b200q4_model_tensor_info",
    "b200q4_model_weight", "b200q4_model_tensor_data", "b200q4_model_free",
    "b200neox_load", "b200neox_n_vocab", "b200neox_eval", "b200neox_free", "b200neox_last_error", "b200neox_decode_is_program", "b200neox_load_sharded", "b200neox_sharded_pool_bytes",
    "b200q4_chain_create", "b200q4_chain_run", "b200q4_chain_info", "b200q4_chain_free",
    "b200q4_chain_sharded_ctr_bytes", "b200q4_chain_create_sharded",
    "b200q4_host_alloc", "b200q4_host_free", "b200q4_launch_count", "b200q4_launch_count_reset",
    "b200q4_kernel_for",
    # include/b200q4_dev.h: bench / test / development helpers, not part of the product ABI
    "b200q4_time_mul_mat_dev", "b200q4_debug_plan", "b200q4_debug_set_tile_min_m", "b200q4_debug_set_fuse_max_k",
    "b200q4_debug_set_slice_min_c", "b200q4_debug_timeline", "b200q4_chain_timeline", "b200q4_chain_time",
]
HOOK_SYMBOLS = ["init_xmax", "imax_ggml_compute_forward_mul_mat_q4_0_f32",
                "imax_ggml_compute_forward_mul_mat_q4_1_f32", "b200q4_hook_shutdown"]


class B200Q4Error(RuntimeError):
    pass


def row_bytes(qtype: int, k: int) -> int:
    return (k // QK) * (20 if qtype == Q4_0 else 24)


def load_library(build: bool = True) -> C.CDLL:
    """dlopen vsim_b200/libb200q4.so (building it with nvcc first if stale)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    from . import build as _b
    import os
    if os.environ.get("B200Q4_LIB"):              # development: an experimental build of the same sources
        path = os.environ["B200Q4_LIB"]
    elif build:
        path = _b.build_library()
    else:
        path = _b.LIB
    if not Path(path).exists():
        raise B200Q4Error(f"{path} missing: build it with `python -m vsim_b200.build` (no CPU fallback exists)")
    L = C.CDLL(str(path))
    vp, sz, i, f32p = C.c_void_p, C.c_size_t, C.c_int, C.c_void_p
    sig = {
        "b200q4_abi_version": (i, []),
        "b200q4_init": (i, [i]),
        "b200q4_shutdown": (None, []),
        "b200q4_last_error": (C.c_char_p, []),
        "b200q4_device_info": (i, [C.c_char_p, sz, C.POINTER(i), C.POINTER(sz)]),
        "b200q4_weight_upload": (vp, [i, vp, i, i, sz]),
        "b200q4_weight_upload_rows": (vp, [i, vp, i, i, sz, i, i]),
        "b200q4_weight_upload_transposed": (vp, [i, vp, i, i, sz]),
        "b200q4_weight_free": (None, [vp]),
        "b200q4_weight_rows": (i, [vp]),
        "b200q4_weight_cols": (i, [vp]),
        "b200q4_weight_type": (i, [vp]),
        "b200q4_weight_device_bytes": (sz, [vp]),
        "b200q4_weight_download": (i, [vp, vp, sz]),
        "b200q4_mul_mat": (i, [vp, f32p, i, sz, f32p, sz]),
        "b200q4_mul_mat_wdata": (i, [vp, vp, i, f32p, sz]),
        "b200q4_quantize_rows": (i, [i, f32p, i, i, sz, vp]),
        "b200q4_mul_mat_transposed": (i, [vp, f32p, i, sz, f32p, sz]),
        "b200q4_mul_mat_dev": (i, [vp, vp, i, sz, vp, sz, vp]),
        "b200q4_quantize_rows_dev": (i, [i, vp, i, i, sz, vp, vp]),
        "b200q4_sync": (i, [vp]),
        "b200q4_set_exact": (i, [i]),
        "b200q4_exact_enabled": (i, []),
        "b200q4_hint_next_weight": (i, [vp, i]),
        "b200q4_model_load": (vp, [C.c_char_p, i, i, i]),
        "b200q4_model_header": (i, [vp, i]),
        "b200q4_model_tensor_count": (i, [vp]),
        "b200q4_model_tensor_info": (i, [vp, i, C.POINTER(C.c_char_p), C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i),
                                         C.POINTER(sz), C.POINTER(sz)]),
        "b200q4_model_weight": (vp, [vp, i]),
        "b200q4_model_tensor_data": (vp, [vp, i]),
        "b200q4_model_free": (None, [vp]),
        "b200neox_load": (vp, [C.c_char_p, i, i]),
        "b200neox_n_vocab": (i, [vp]),
        "b200neox_eval": (i, [vp, vp, i, i, i, vp]),
        "b200neox_free": (None, [vp]),
        "b200neox_decode_is_program": (i, [vp]),
        "b200neox_load_sharded": (vp, [C.c_char_p, i, i, i, C.POINTER(vp), sz]),
        "b200neox_sharded_pool_bytes": (sz, [i, i, i]),
        "b200q4_chain_create": (vp, [i, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i)]),
        "b200q4_chain_sharded_ctr_bytes": (sz, [i]),
        "b200q4_chain_create_sharded": (vp, [i, C.POINTER(vp), C.POINTER(vp), i, i, C.POINTER(vp), sz, C.POINTER(sz)]),
        "b200q4_chain_run": (i, [vp, vp]),
        "b200q4_chain_info": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(sz)]),
        "b200q4_chain_free": (None, [vp]),
        "b200neox_last_error": (C.c_char_p, []),
        "b200q4_quantize_weights": (i, [i, f32p, C.c_longlong, i, sz, vp, vp]),
        "b200q4_mul_mat_dev_allgather": (i, [vp, vp, i, sz, C.POINTER(vp), i, i, i, sz, C.POINTER(vp), vp, vp]),
        "b200q4_wait_flags": (i, [vp, i, vp, vp]),
        "b200q4_host_alloc": (vp, [sz]),
        "b200q4_host_free": (None, [vp]),
        "b200q4_launch_count": (C.c_uint64, []),
        "b200q4_launch_count_reset": (None, []),
        "b200q4_kernel_for": (C.c_char_p, [i, i, i, i]),
        "b200q4_time_mul_mat_dev": (C.c_double, [C.POINTER(vp), i, vp, i, vp, i]),
        "b200q4_chain_time": (C.c_double, [vp, vp, i, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _LIB = L
    return L


def _err(L) -> str:
    return (L.b200q4_last_error() or b"").decode(errors="replace")


def _check(L, rc: int, what: str) -> None:
    if rc != 0:
        raise B200Q4Error(f"{what} failed (rc={rc}): {_err(L)}")


def init(device: int = 0) -> None:
    global _INIT_DEVICE
    L = load_library()
    _check(L, L.b200q4_init(device), "b200q4_init")
    _INIT_DEVICE = device


def is_initialised() -> bool:
    return _INIT_DEVICE is not None


def shutdown() -> None:
    global _INIT_DEVICE
    if _LIB is not None:
        _LIB.b200q4_shutdown()
    _INIT_DEVICE = None


def device_info():
    L = load_library()
    name = C.create_string_buffer(256)
    sms, hbm = C.c_int(0), C.c_size_t(0)
    _check(L, L.b200q4_device_info(name, 256, C.byref(sms), C.byref(hbm)), "b200q4_device_info")
    return name.value.decode(), sms.value, hbm.value


def launch_count() -> int:
    return int(load_library().b200q4_launch_count())


def launch_count_reset() -> None:
    load_library().b200q4_launch_count_reset()


def kernel_for(qtype: int, m: int, n: int, k: int) -> str:
    return load_library().b200q4_kernel_for(qtype, m, n, k).decode()


class Q4Weight:
    """One Q4_0/Q4_1 weight tensor resident in HBM (b200q4_weight*)."""

    def __init__(self, qtype: int, rows: np.ndarray, n: int, k: int, *, row0: int = 0, nrows: int | None = None,
                 transposed: bool = False):
        L = self._L = load_library()
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        self.qtype, self.k, self.transposed = qtype, k, transposed
        qlen, nr = (n, k) if transposed else (k, n)
        rb = row_bytes(qtype, qlen)
        if rows.size != nr * rb:
            raise ValueError(f"weight bytes {rows.size} != {nr} rows x {rb} B")
        ptr = rows.ctypes.data if rows.size else C.addressof(C.create_string_buffer(1))
        if transposed:
            self.n = n
            self.handle = L.b200q4_weight_upload_transposed(qtype, ptr, n, k, rb)
        elif nrows is None and row0 == 0:
            self.n = n
            self.handle = L.b200q4_weight_upload(qtype, ptr, n, k, rb)
        else:
            self.n = n - row0 if nrows is None else nrows
            self.handle = L.b200q4_weight_upload_rows(qtype, ptr, n, k, rb, row0, self.n)
        if not self.handle:
            raise B200Q4Error(f"b200q4_weight_upload failed: {_err(L)}")

    @property
    def device_bytes(self) -> int:
        return int(self._L.b200q4_weight_device_bytes(self.handle))

    def download(self) -> np.ndarray:
        qlen, nr = (self.n, self.k) if self.transposed else (self.k, self.n)
        out = np.empty((nr, row_bytes(self.qtype, qlen)), dtype=np.uint8)
        _check(self._L, self._L.b200q4_weight_download(self.handle, out.ctypes.data, out.shape[1]), "weight_download")
        return out

    def free(self) -> None:
        if getattr(self, "handle", None):
            self._L.b200q4_weight_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Chain:
    """b200q4_chain: n dependent M = 1 mul_mats (device pointers) executed by one persistent launch."""

    def __init__(self, weights, x_ptrs, y_ptrs, deps=None):
        L = self._L = load_library()
        n = len(weights)
        ws = (C.c_void_p * n)(*[w.handle for w in weights])
        xs = (C.c_void_p * n)(*x_ptrs)
        ys = (C.c_void_p * n)(*y_ptrs)
        dp = (C.c_int * n)(*deps) if deps is not None else None
        self.handle = L.b200q4_chain_create(n, ws, xs, ys, dp)
        if not self.handle:
            raise B200Q4Error(f"b200q4_chain_create failed: {_err(L)}")
        self.n = n

    def run(self, stream=None) -> None:
        _check(self._L, self._L.b200q4_chain_run(self.handle, stream), "b200q4_chain_run")

    def info(self):
        n, st, sm = C.c_int(0), C.c_int(0), C.c_size_t(0)
        self._L.b200q4_chain_info(self.handle, C.byref(n), C.byref(st), C.byref(sm))
        return {"ops": n.value, "stages": st.value, "smem": sm.value}

    def free(self) -> None:
        if getattr(self, "handle", None):
            self._L.b200q4_chain_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def mul_mat(w: Q4Weight, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """Y[M][N] = ggml mul_mat(W, X): host fp32 in, host fp32 out, through b200q4_mul_mat."""
    L = load_library()
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    if x.shape[1] != w.k or (x.strides[1] != 4 and x.shape[1] > 1):
        if x.shape[1] != w.k:
            raise ValueError(f"X has K={x.shape[1]}, weight has K={w.k}")
        x = np.ascontiguousarray(x)
    m = x.shape[0]
    y = out if out is not None else np.empty((m, w.n), dtype=np.float32)
    xs = x.strides[0] if m > 1 else w.k * 4
    ys = y.strides[0] if m > 1 else max(w.n, 1) * 4
    _check(L, L.b200q4_mul_mat(w.handle, x.ctypes.data, m, xs, y.ctypes.data, ys), "b200q4_mul_mat")
    return y


def mul_mat_wdata(w: Q4Weight, wdata: np.ndarray, m: int) -> np.ndarray:
    L = load_library()
    wdata = np.ascontiguousarray(wdata, dtype=np.uint8)
    if wdata.size != m * row_bytes(w.qtype, w.k):
        raise ValueError("wdata size mismatch")
    y = np.empty((m, w.n), dtype=np.float32)
    _check(L, L.b200q4_mul_mat_wdata(w.handle, wdata.ctypes.data, m, y.ctypes.data, max(w.n, 1) * 4), "b200q4_mul_mat_wdata")
    return y


def quantize_weights(qtype: int, w: np.ndarray):
    """ggml_quantize_q4_0/_q4_1 (utils.cpp:424-541) on the device: fp32 rows [n][k] -> (reference-format rows, histogram int64[16])."""
    L = load_library()
    w = np.asarray(w, dtype=np.float32)
    if w.ndim != 2 or w.strides[1] != 4:
        w = np.ascontiguousarray(w)
    n, k = w.shape
    out = np.zeros((n, row_bytes(qtype, k)), np.uint8)
    hist = np.zeros(16, np.int64)
    rc = L.b200q4_quantize_weights(qtype, w.ctypes.data, n, k, w.strides[0] if n > 1 else 0,
                                   out.ctypes.data, hist.ctypes.data)
    if rc:
        raise B200Q4Error(f"b200q4_quantize_weights failed (rc={rc}): {_err(L)}")
    return out, hist


def mul_mat_transposed(wt: Q4Weight, x: np.ndarray) -> np.ndarray:
    L = load_library()
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, wt.k) if wt.k else np.zeros((np.asarray(x).shape[0], 0), np.float32)
    m = x.shape[0]
    y = np.empty((m, wt.n), dtype=np.float32)
    _check(L, L.b200q4_mul_mat_transposed(wt.handle, x.ctypes.data, m, max(wt.k, 1) * 4, y.ctypes.data, max(wt.n, 1) * 4),
           "b200q4_mul_mat_transposed")
    return y


def quantize_rows(qtype: int, x: np.ndarray) -> np.ndarray:
    """quantize_row_q4_0/_q4_1 over the rows of x on the GPU; returns the reference's byte image."""
    L = load_library()
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    m, k = x.shape
    out = np.empty((m, row_bytes(qtype, k)), dtype=np.uint8)
    _check(L, L.b200q4_quantize_rows(qtype, x.ctypes.data, m, k, k * 4, out.ctypes.data), "b200q4_quantize_rows")
    return out
