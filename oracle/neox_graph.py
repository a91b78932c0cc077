"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's whole GPT-NeoX forward pass
(gptneox_eval, vsim.cpp:470-747) on top of the op restatements of q4_oracle.c / neox_ops_oracle.c.

Row f-1 groundwork: a device-resident graph executor has to reproduce exactly this.  The
restatement is pinned to the UNMODIFIED reference program: tests/test_neox_graph_oracle.py compares the prompt logits
(every printed digit) and the greedy continuation with oracle/_ref/vsim-ubuntu at several thread counts -- the thread
count matters, because ggml's transposed f32 mul_mat (V^T x softmax, ggml.c:4517-4553) gives every thread a range of
positions and adds the per-thread partial sums afterwards (ggml.c:4463-4481).
"""
from __future__ import annotations

import ctypes as C
import struct
from pathlib import Path

import numpy as np

from . import HERE, Oracle, Q4_0, Q4_1, row_bytes

_F = np.float32


class NeoxFile:
    """The reference's on-disk model format (vsim.cpp:119-138 header, :163-173 vocabulary, :375-448 tensors)."""

    def __init__(self, path):
        b = np.fromfile(path, dtype=np.uint8)
        magic, = struct.unpack_from("<i", b, 0)
        assert magic == 0x67676d6c
        (self.n_vocab, self.n_embd, self.n_head, self.n_layer, self.n_rot, self.par_res, self.ftype) = struct.unpack_from("<7i", b, 4)
        off = 32
        for _ in range(self.n_vocab):
            ln, = struct.unpack_from("<I", b, off)
            off += 4 + ln
        self.t = {}
        while off < b.size:
            nd, ln, ft = struct.unpack_from("<3i", b, off)
            off += 12
            dims = struct.unpack_from("<%di" % nd, b, off)
            off += 4 * nd
            name = bytes(b[off:off + ln]).decode()
            off += ln
            if nd == 1:
                self.t[name] = b[off:off + dims[0] * 4].view(np.float32).copy()
                off += dims[0] * 4
            else:
                k, n = dims
                qt = ft - 2
                nb = n * row_bytes(qt, k)
                self.t[name] = (qt, n, k, np.ascontiguousarray(b[off:off + nb]).reshape(n, -1))
                off += nb
        self.qtype = self.ftype - 2


class NeoxOracle:
    def __init__(self, path, n_ctx: int = 512):
        self.f = NeoxFile(path)
        self.o = Oracle()
        L = self.L = C.CDLL(str(HERE / "libq4oracle.so"))
        L.nxo_norm_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.nxo_gelu.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.nxo_soft_max_rows.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.nxo_rope_neox.argtypes = [C.c_void_p] + [C.c_int] * 6
        L.nxo_get_rows_q4.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        f = self.f
        self.mem_k = np.zeros((f.n_layer, n_ctx, f.n_embd), _F)       # vsim.cpp:318-330: F32 KV memory
        self.mem_v = np.zeros((f.n_layer, n_ctx, f.n_embd), _F)

    # ---- op wrappers -----------------------------------------------------------------------------
    def _norm(self, x):
        y = np.empty_like(x)
        self.L.nxo_norm_rows(x.ctypes.data, x.shape[0], x.shape[1], y.ctypes.data)
        return y

    def _ln(self, x, name):                      # norm, then mul(repeat(w), cur), then add(.., repeat(b))   (vsim.cpp:523-533)
        return (self.f.t[name + ".weight"] * self._norm(x)).astype(_F) + self.f.t[name + ".bias"]

    def _mm(self, name, x):                      # ggml_mul_mat on a Q4 weight: the hot path's own oracle
        qt, n, k, wq = self.f.t[name]
        return self.o.mul_mat(qt, wq, n, k, np.ascontiguousarray(x, _F), nth=1)

    def _gelu(self, x):
        x = np.ascontiguousarray(x, _F)
        y = np.empty_like(x)
        self.L.nxo_gelu(x.ctypes.data, x.size, y.ctypes.data)
        return y

    # ---- gptneox_eval (vsim.cpp:470-747) ------------------------------------------------------------
    def eval(self, tokens, n_past: int, n_threads: int = 1) -> np.ndarray:
        f = self.f
        N, E, H = len(tokens), f.n_embd, f.n_head
        d = E // H
        P = n_past + N
        ids = np.asarray(tokens, np.int32)
        qt, nv, k, wte = f.t["gpt_neox.embed_in.weight"]
        x = np.empty((N, E), _F)
        self.L.nxo_get_rows_q4(qt, wte.ctypes.data, E, ids.ctypes.data, N, x.ctypes.data)          # ggml_get_rows(wte, embd)
        scale = _F(1.0 / np.sqrt(np.float64(_F(E) / _F(H))))                                      # vsim.cpp:590: 1.0f/sqrt(float(n_embd)/n_head)
        for l in range(f.n_layer):
            p = f"gpt_neox.layers.{l}."
            cur = self._ln(x, p + "input_layernorm")
            q = self._mm(p + "attention.query.weight", cur) + f.t[p + "attention.query.bias"]
            kk = self._mm(p + "attention.key.weight", cur) + f.t[p + "attention.key.bias"]
            v = self._mm(p + "attention.value.weight", cur) + f.t[p + "attention.value.bias"]
            self.mem_k[l, n_past:P] = kk                                                          # cpy into the KV memory (:557-558)
            self.mem_v[l, n_past:P] = v
            qr = np.ascontiguousarray(q.reshape(N, H, d))
            self.L.nxo_rope_neox(qr.ctypes.data, d, H, N, n_past, f.n_rot, 0)                      # Q: mode 0, positions n_past + i
            km = self.mem_k[l, :P].reshape(P, H, d)                                                # K: the memory view itself, mode 1, IN PLACE
            assert km.flags.c_contiguous and km.base is not None
            self.L.nxo_rope_neox(km.ctypes.data, d, H, P, n_past, f.n_rot, 1)
            # K * Q (ggml.c:4484-4516, vec_dot_f32 :399-434: float products accumulated sequentially in double)
            prod = (km.transpose(1, 0, 2)[:, None, :, :] * qr.transpose(1, 0, 2)[:, :, None, :]).astype(_F)   # [H][N][P][d]
            kq = np.cumsum(prod.astype(np.float64), axis=-1)[..., -1].astype(_F)
            kq = (kq * scale).astype(_F)                                                           # ggml_scale
            for n in range(N):                                                                     # ggml_diag_mask_inf(n_past)
                kq[:, n, n_past + n + 1:] = -np.inf
            kq = np.ascontiguousarray(kq)
            self.L.nxo_soft_max_rows(kq.ctypes.data, H * N, P)
            # V^T * softmax: the transposed branch -- thread k owns positions [k*dc, (k+1)*dc), float accumulation in
            # position order from +0, then dst = part[0]; dst += part[k] for k = 1 .. nth-1   (MUL_MAT n_tasks = n_threads)
            vm = self.mem_v[l, :P].reshape(P, H, d)
            dc = (P + n_threads - 1) // n_threads
            kqv = None
            for t in range(n_threads):
                part = np.zeros((H, N, d), _F)
                for ic in range(t * dc, min((t + 1) * dc, P)):
                    part = (part + (vm[ic][:, None, :] * kq[:, :, ic][:, :, None]).astype(_F)).astype(_F)
                kqv = part if kqv is None else (kqv + part).astype(_F)
            merged = np.ascontiguousarray(kqv.transpose(1, 0, 2).reshape(N, E))                    # permute(0,2,1,3) + cpy
            att = f.t[p + "attention.dense.bias"] + self._mm(p + "attention.dense.weight", merged)
            if f.par_res == 0:                                                                     # vsim.cpp:636-665
                ff = self._ln((att + x).astype(_F), p + "post_attention_layernorm")                # inpFF = norm(cur + inpL)
            else:                                                                                  # vsim.cpp:666-696
                ff = self._ln(x, p + "post_attention_layernorm")
            ff = f.t[p + "mlp.dense_h_to_4h.bias"] + self._mm(p + "mlp.dense_h_to_4h.weight", ff)
            ff = self._gelu(ff)
            ff = f.t[p + "mlp.dense_4h_to_h.bias"] + self._mm(p + "mlp.dense_4h_to_h.weight", ff)
            if f.par_res == 0:
                x = (ff + x).astype(_F)                                                            # inpL = inpFF + inpL  (attention output only feeds the MLP)
            else:
                ff = (att + ff).astype(_F)
                x = (x + ff).astype(_F)
        x = self._ln(x, "gpt_neox.final_layer_norm")
        return self._mm("embed_out.weight", x)[N - 1]                                              # logits of the last token
