"""CPU oracle for vsim's Q4_0/Q4_1 mul_mat hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``vsim_b200``) never does; it fails loudly when its CUDA library is missing.

Two checkers live here:

* :class:`Oracle` -- ``libq4oracle.so``, our C restatement (``q4_oracle.c``) of
  ggml.c:209-298 (quantizers), :472-553 (block dots), :4891-5453 + imax.c:1182-1230
  (the operator).  Always available (gcc is in the image).
* :class:`Reference` -- ``_ref/libggml_ref.so``, the UNMODIFIED reference
  compiled from ``/root/reference`` by ``oracle/Makefile`` (present in the
  build container, travels to the GPU box as a built artefact).  Used to pin
  the restatement and as the "reference" CPU baseline.

Parity status: pinned (see q4_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
QK = 32
Q4_0, Q4_1 = 0, 1          # enum ggml_type, ggml.h:217-226


def row_bytes(qtype: int, k: int) -> int:
    """Bytes of one quantized row (GGML_TYPE_SIZE / GGML_BLCK_SIZE, ggml.c:895-915)."""
    return (k // QK) * (20 if qtype == Q4_0 else 24)


def build(ref: bool = True) -> None:
    """Compile libq4oracle.so (and _ref/ when /root/reference is mounted)."""
    target = ["all"] if ref else [str(HERE / "libq4oracle.so")]
    subprocess.run(["make", "-s", "-C", str(HERE), *target], check=True)


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Oracle:
    """ctypes view of libq4oracle.so (our restatement)."""

    def __init__(self):
        so = HERE / "libq4oracle.so"
        if not so.exists() or so.stat().st_mtime < (HERE / "q4_oracle.c").stat().st_mtime:
            build(ref=False)
        L = self.lib = C.CDLL(str(so))
        for n in ("q4o_quantize_row_q4_0", "q4o_quantize_row_q4_1"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            getattr(L, n).restype = None
        for n in ("q4o_dequantize_row_q4_0", "q4o_dequantize_row_q4_1"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            getattr(L, n).restype = None
        for n in ("q4o_vec_dot_q4_0", "q4o_vec_dot_q4_1"):
            getattr(L, n).argtypes = [C.c_int, C.c_void_p, C.c_void_p]
            getattr(L, n).restype = C.c_float
        L.q4o_mul_mat.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                  C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.q4o_mul_mat.restype = C.c_int
        L.q4o_mul_mat_transposed.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                             C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.q4o_mul_mat_transposed.restype = C.c_int

    # -- quantizers ---------------------------------------------------------
    def quantize_rows(self, qtype: int, x: np.ndarray) -> np.ndarray:
        """x: [R][K] fp32 -> [R][row_bytes] uint8 in the reference's row format."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x[None, :]
        r, k = x.shape
        assert k % QK == 0
        out = np.empty((r, row_bytes(qtype, k)), dtype=np.uint8)
        fn = self.lib.q4o_quantize_row_q4_0 if qtype == Q4_0 else self.lib.q4o_quantize_row_q4_1
        for i in range(r):
            fn(x[i].ctypes.data, out[i].ctypes.data, k)
        return out

    def dequantize_rows(self, qtype: int, q: np.ndarray, k: int) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.uint8).reshape(-1, row_bytes(qtype, k))
        out = np.empty((q.shape[0], k), dtype=np.float32)
        fn = self.lib.q4o_dequantize_row_q4_0 if qtype == Q4_0 else self.lib.q4o_dequantize_row_q4_1
        for i in range(q.shape[0]):
            fn(q[i].ctypes.data, out[i].ctypes.data, k)
        return out

    def vec_dot(self, qtype: int, k: int, w_row: np.ndarray, a_row: np.ndarray) -> np.float32:
        fn = self.lib.q4o_vec_dot_q4_0 if qtype == Q4_0 else self.lib.q4o_vec_dot_q4_1
        w_row = np.ascontiguousarray(w_row, dtype=np.uint8)
        a_row = np.ascontiguousarray(a_row, dtype=np.uint8)
        return np.float32(fn(k, w_row.ctypes.data, a_row.ctypes.data))

    # -- the operator -------------------------------------------------------
    def mul_mat(self, qtype: int, w: np.ndarray, n: int, k: int, x: np.ndarray,
                nth: int = 1, return_wdata: bool = False):
        """Y[M][N] = mul_mat(W[N][K] quantized bytes, X[M][K] fp32), reference semantics."""
        w = np.ascontiguousarray(w, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, k)
        m = x.shape[0]
        assert w.size == n * row_bytes(qtype, k)
        y = np.empty((m, n), dtype=np.float32)
        wd = np.empty((m, row_bytes(qtype, k)), dtype=np.uint8) if return_wdata else None
        rc = self.lib.q4o_mul_mat(qtype, w.ctypes.data, n, k, x.ctypes.data, m, k,
                                  y.ctypes.data, n, nth, wd.ctypes.data if wd is not None else None)
        if rc != 0:
            raise ValueError("q4o_mul_mat: bad arguments")
        return (y, wd) if return_wdata else y

    def mul_mat_transposed(self, qtype: int, wt: np.ndarray, n: int, k: int, x: np.ndarray,
                           nth: int = 1) -> np.ndarray:
        """src0 stored as K rows of N quantized values (nb01 < nb00 branch)."""
        wt = np.ascontiguousarray(wt, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, k)
        m = x.shape[0]
        y = np.empty((m, n), dtype=np.float32)
        rc = self.lib.q4o_mul_mat_transposed(qtype, wt.ctypes.data, n, k, x.ctypes.data, m, k,
                                             y.ctypes.data, n, nth)
        if rc != 0:
            raise ValueError("q4o_mul_mat_transposed: bad arguments")
        return y


def quantize_weights(qtype: int, w: np.ndarray):
    """Oracle restatement of the reference's OFFLINE weight quantizers (utils.cpp:424-541): rows [n][k] fp32 ->
    (reference-format rows uint8 [n][row_bytes], histogram int64[16])."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    n, k = w.shape
    L = C.CDLL(str(HERE / "libq4oracle.so"))
    L.q4o_quantize_weights.argtypes = [C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p]
    L.q4o_quantize_weights.restype = C.c_size_t
    out = np.zeros((n, row_bytes(qtype, k)), np.uint8)
    hist = np.zeros(16, np.int64)
    L.q4o_quantize_weights(qtype, w.ctypes.data, n, k, out.ctypes.data, hist.ctypes.data)
    return out, hist


def reference_quantize_weights(qtype: int, w: np.ndarray):
    """The UNMODIFIED reference functions ggml_quantize_q4_0/_q4_1 of utils.cpp (oracle/_ref/libutils_ref.so)."""
    so = REF_DIR / "libutils_ref.so"
    if not so.exists():
        raise FileNotFoundError(f"{so} missing: run `make -C oracle ref` where /root/reference is mounted")
    L = C.CDLL(str(so))
    C.c_int.in_dll(L, "do_print").value = 0                 # utils.cpp:423: silences the debug printf of the first block
    fn = getattr(L, "_Z18ggml_quantize_q4_0PfPviiiPl" if qtype == Q4_0 else "_Z18ggml_quantize_q4_1PfPviiiPl")
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    fn.restype = C.c_size_t
    w = np.ascontiguousarray(w, dtype=np.float32)
    n, k = w.shape
    out = np.zeros((n, row_bytes(qtype, k)), np.uint8)
    hist = np.zeros(16, np.int64)
    fn(w.ctypes.data, out.ctypes.data, n * k, k, 32, hist.ctypes.data)
    return out, hist


def have_reference() -> bool:
    return (REF_DIR / "libggml_ref.so").exists()


class Reference:
    """ctypes view of _ref/libggml_ref.so: the reference's own compiled code."""

    def __init__(self):
        so = REF_DIR / "libggml_ref.so"
        if not so.exists():
            raise FileNotFoundError(f"{so} missing: run `make -C oracle ref` where /root/reference is mounted")
        L = self.lib = C.CDLL(str(so))
        for n in ("quantize_row_q4_0", "quantize_row_q4_1", "dequantize_row_q4_0", "dequantize_row_q4_1"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            getattr(L, n).restype = None
        L.ref_mul_mat.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int, C.c_int]
        L.ref_mul_mat.restype = C.c_double
        L.ref_mul_mat_steps.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_mul_mat_steps.restype = C.c_double
        L.ref_mul_mat_transposed.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                             C.c_int, C.c_void_p, C.c_int]
        L.ref_mul_mat_transposed.restype = C.c_double
        L.ref_tensor_layout.argtypes = [C.POINTER(C.c_size_t)]
        L.ref_tensor_layout.restype = None

    def quantize_rows(self, qtype: int, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x[None, :]
        r, k = x.shape
        out = np.empty((r, row_bytes(qtype, k)), dtype=np.uint8)
        fn = self.lib.quantize_row_q4_0 if qtype == Q4_0 else self.lib.quantize_row_q4_1
        for i in range(r):
            fn(x[i].ctypes.data, out[i].ctypes.data, k)
        return out

    def dequantize_rows(self, qtype: int, q: np.ndarray, k: int) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.uint8).reshape(-1, row_bytes(qtype, k))
        out = np.empty((q.shape[0], k), dtype=np.float32)
        fn = self.lib.dequantize_row_q4_0 if qtype == Q4_0 else self.lib.dequantize_row_q4_1
        for i in range(q.shape[0]):
            fn(q[i].ctypes.data, out[i].ctypes.data, k)
        return out

    def mul_mat(self, qtype: int, w: np.ndarray, n: int, k: int, x: np.ndarray,
                nth: int = 1, reps: int = 1, return_time: bool = False):
        """ggml_mul_mat + ggml_graph_compute of the reference (through its IMAX hook for Q4_0)."""
        w = np.ascontiguousarray(w, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, k)
        m = x.shape[0]
        y = np.empty((m, n), dtype=np.float32)
        us = self.lib.ref_mul_mat(qtype, w.ctypes.data, n, k, x.ctypes.data, m, y.ctypes.data, nth, reps)
        if us < 0:
            raise RuntimeError("ref_mul_mat failed")
        return (y, us) if return_time else y

    def mul_mat_steps(self, qtype: int, w: np.ndarray, n: int, k: int, x: np.ndarray, nth: int,
                      warmup: int, steps: int):
        """warmup + steps ggml_graph_compute calls on one resident graph; returns (Y, seconds for the timed steps)."""
        w = np.ascontiguousarray(w, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, k)
        m = x.shape[0]
        y = np.empty((m, n), dtype=np.float32)
        us = self.lib.ref_mul_mat_steps(qtype, w.ctypes.data, n, k, x.ctypes.data, m, y.ctypes.data, nth, warmup, steps)
        if us < 0:
            raise RuntimeError("ref_mul_mat_steps failed")
        return y, us * 1e-6

    def mul_mat_transposed(self, qtype: int, wt: np.ndarray, n: int, k: int, x: np.ndarray,
                           nth: int = 1) -> np.ndarray:
        wt = np.ascontiguousarray(wt, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, k)
        m = x.shape[0]
        y = np.empty((m, n), dtype=np.float32)
        self.lib.ref_mul_mat_transposed(qtype, wt.ctypes.data, n, k, x.ctypes.data, m, y.ctypes.data, nth)
        return y

    def tensor_layout(self):
        out = (C.c_size_t * 8)()
        self.lib.ref_tensor_layout(out)
        return list(out)


# ---------------------------------------------------------------------------
# numpy restatement of the two quantizers (vectorised).
# Used as a third, independent voice in tests/test_oracle.py.
# ---------------------------------------------------------------------------
def _round_to_int_np(v: np.ndarray) -> np.ndarray:
    r = np.where(np.isfinite(v), np.trunc(v.astype(np.float64) + np.copysign(0.5, v.astype(np.float64))), np.nan)
    bad = ~(np.abs(r) < 2147483648.0)
    return np.where(bad, np.int64(-2147483648), np.nan_to_num(r).astype(np.int64))


def np_quantize_rows_q4_0(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    r, k = x.shape
    nb = k // QK
    xb = x.reshape(r, nb, QK)
    amax = np.abs(xb).max(axis=2)
    with np.errstate(all="ignore"):
        d = (amax / np.float32(7.0)).astype(np.float32)
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
        v = (xb * idv[:, :, None]).astype(np.float32)
    codes = ((_round_to_int_np(v).astype(np.int8).astype(np.int16) + 8) & 0xFF).astype(np.uint8)
    out = np.empty((r, nb, 20), dtype=np.uint8)
    out[:, :, :4] = d.view(np.uint8).reshape(r, nb, 4)
    out[:, :, 4:] = codes[:, :, 0::2] | (codes[:, :, 1::2] << 4)
    return out.reshape(r, nb * 20)


def np_quantize_rows_q4_1(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    r, k = x.shape
    nb = k // QK
    xb = x.reshape(r, nb, QK)
    mn = xb.min(axis=2)
    mx = xb.max(axis=2)
    with np.errstate(all="ignore"):
        d = ((mx - mn).astype(np.float32) / np.float32(15.0)).astype(np.float32)
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
        v = ((xb - mn[:, :, None]).astype(np.float32) * idv[:, :, None]).astype(np.float32)
    codes = (_round_to_int_np(v) & 0xFF).astype(np.uint8)
    out = np.empty((r, nb * 24), dtype=np.uint8)
    out[:, : nb * 4] = mn.astype(np.float32).view(np.uint8).reshape(r, nb * 4)
    out[:, nb * 4: nb * 8] = d.view(np.uint8).reshape(r, nb * 4)
    out[:, nb * 8:] = (codes[:, :, 0::2] | (codes[:, :, 1::2] << 4)).reshape(r, nb * 16)
    return out
