"""Writes a random-init GPT-NeoX model straight into the reference's on-disk q4 format
(vsim.cpp:119-138 header, :163-173 vocab, :375-448 tensors), quantizing the
2-D weights row by row with the oracle's quantize_row (test infrastructure).  Used by the
end-to-end drop-in test: the UNMODIFIED reference binary vs the same sources linked to libb200q4.so.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

from oracle import Oracle, Q4_0, Q4_1


def write_neox(path: Path, qtype: int, n_vocab=256, n_embd=256, n_head=4, n_layer=2, n_rot=32, seed=0, par_res=1) -> None:
    orc = Oracle()
    rng = np.random.default_rng(seed)
    ftype = 2 if qtype == Q4_0 else 3

    def t2(name, k, n, scale=0.08):          # Linear.weight[N][K], quantized along K
        w = (rng.standard_normal((n, k)) * scale).astype(np.float32)
        data = orc.quantize_rows(qtype, w).tobytes()
        nm = name.encode()
        return struct.pack("<iii", 2, len(nm), ftype) + struct.pack("<ii", k, n) + nm + data

    def t1(name, n, mean=0.0, scale=0.02):
        v = (mean + rng.standard_normal(n) * scale).astype(np.float32)
        nm = name.encode()
        return struct.pack("<iii", 1, len(nm), 0) + struct.pack("<i", n) + nm + v.tobytes()

    out = [struct.pack("<i", 0x67676d6c),
           struct.pack("<iiiiiii", n_vocab, n_embd, n_head, n_layer, n_rot, par_res, ftype)]
    for i in range(n_vocab):
        tok = f"t{i}".encode()
        out.append(struct.pack("<I", len(tok)) + tok)
    out.append(t2("gpt_neox.embed_in.weight", n_embd, n_vocab, 0.5))
    out.append(t1("gpt_neox.final_layer_norm.weight", n_embd, 1.0))
    out.append(t1("gpt_neox.final_layer_norm.bias", n_embd))
    out.append(t2("embed_out.weight", n_embd, n_vocab))
    for l in range(n_layer):
        p = f"gpt_neox.layers.{l}."
        out.append(t1(p + "input_layernorm.weight", n_embd, 1.0))
        out.append(t1(p + "input_layernorm.bias", n_embd))
        out.append(t1(p + "post_attention_layernorm.weight", n_embd, 1.0))
        out.append(t1(p + "post_attention_layernorm.bias", n_embd))
        for nm in ("query", "key", "value", "dense"):
            out.append(t2(p + f"attention.{nm}.weight", n_embd, n_embd))
            out.append(t1(p + f"attention.{nm}.bias", n_embd))
        out.append(t2(p + "mlp.dense_h_to_4h.weight", n_embd, 4 * n_embd))
        out.append(t1(p + "mlp.dense_h_to_4h.bias", 4 * n_embd))
        out.append(t2(p + "mlp.dense_4h_to_h.weight", 4 * n_embd, n_embd))
        out.append(t1(p + "mlp.dense_4h_to_h.bias", n_embd))
    Path(path).write_bytes(b"".join(out))
