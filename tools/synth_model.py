"""Random-init GPT-NeoX model files in the reference's on-disk format (vsim.cpp:119-138 header, :163-173 vocabulary,
:375-448 tensors), written straight as q4 bytes -- an f32 intermediate of a 6B / 20B model would be
23 / 80 GB.  Used by bench.py's full-graph decode legs and tools/neox_bench.py.  Every 2-D tensor gets random nibbles and
the block scale gain / (4.6 sqrt(K)) (gain < 1 keeps a deep random network contractive); the bytes of one layer are
generated once and written for every layer (distinct HBM copies after upload, which is what the weight stream needs)."""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

SHAPES = {  # name: (n_vocab, n_embd, n_head, n_layer, n_rot)
    "gptj6b": (50400, 4096, 16, 28, 64),
    "neox20b": (50432, 6144, 64, 44, 24),
    "pythia12b": (50288, 5120, 40, 36, 32),
    "mid": (8192, 2048, 16, 8, 32),
    "gpt2_117m": (50257, 768, 12, 12, 16),       # BASELINE.json configs[0] as a NeoX-shaped stand-in (the reference ships no GPT-2 driver)
    "tiny": (512, 256, 4, 2, 32),
}


def write_model(path: Path, n_vocab, n_embd, n_head, n_layer, n_rot, qtype=0, par_res=1, seed=0, gain=0.5) -> int:
    rng = np.random.default_rng(seed)
    cache = {}

    def q4_bytes(k, n):
        key = (k, n)
        if key not in cache:
            nb = k // 32
            d = np.float32(gain / (4.6 * np.sqrt(k)))
            if qtype == 0:
                raw = rng.integers(0, 256, size=(min(n, 4096), nb, 20), dtype=np.uint8)
                raw[:, :, :4] = np.frombuffer(d.tobytes(), np.uint8)
            else:                                      # q4_1 row: min[nb] | d[nb] | nibbles[nb][16]
                raw = rng.integers(0, 256, size=(min(n, 4096), nb * 24), dtype=np.uint8)
                raw[:, : nb * 4] = np.frombuffer(np.full(nb, -7.5 * d, np.float32).tobytes(), np.uint8)
                raw[:, nb * 4: nb * 8] = np.frombuffer(np.full(nb, d, np.float32).tobytes(), np.uint8)
            cache[key] = raw.tobytes()
        return cache[key]

    def t2(f, name, k, n):
        nm = name.encode()
        f.write(struct.pack("<iii", 2, len(nm), 2 + qtype) + struct.pack("<ii", k, n) + nm)
        slab = q4_bytes(k, n)
        rows = min(n, 4096)
        rb = len(slab) // rows
        left = n
        while left > 0:                                   # the same 4096-row slab again and again
            nr = min(rows, left)
            f.write(slab[: nr * rb])
            left -= nr

    def t1(f, name, n, mean=0.0, scale=0.02):
        nm = name.encode()
        v = (mean + rng.standard_normal(n) * scale).astype(np.float32)
        f.write(struct.pack("<iii", 1, len(nm), 0) + struct.pack("<i", n) + nm + v.tobytes())

    with open(path, "wb") as f:
        f.write(struct.pack("<i", 0x67676d6c))
        f.write(struct.pack("<iiiiiii", n_vocab, n_embd, n_head, n_layer, n_rot, par_res, 2 + qtype))
        for i in range(n_vocab):
            tok = f"t{i}".encode()
            f.write(struct.pack("<I", len(tok)) + tok)
        t2(f, "gpt_neox.embed_in.weight", n_embd, n_vocab)
        t1(f, "gpt_neox.final_layer_norm.weight", n_embd, 1.0)
        t1(f, "gpt_neox.final_layer_norm.bias", n_embd)
        t2(f, "embed_out.weight", n_embd, n_vocab)
        for l in range(n_layer):
            p = f"gpt_neox.layers.{l}."
            t1(f, p + "input_layernorm.weight", n_embd, 1.0)
            t1(f, p + "input_layernorm.bias", n_embd)
            t1(f, p + "post_attention_layernorm.weight", n_embd, 1.0)
            t1(f, p + "post_attention_layernorm.bias", n_embd)
            for nm in ("query", "key", "value", "dense"):
                t2(f, p + f"attention.{nm}.weight", n_embd, n_embd)
                t1(f, p + f"attention.{nm}.bias", n_embd)
            t2(f, p + "mlp.dense_h_to_4h.weight", n_embd, 4 * n_embd)
            t1(f, p + "mlp.dense_h_to_4h.bias", 4 * n_embd)
            t2(f, p + "mlp.dense_4h_to_h.weight", 4 * n_embd, n_embd)
            t1(f, p + "mlp.dense_4h_to_h.bias", n_embd)
    return Path(path).stat().st_size


def q4_weight_bytes(n_vocab, n_embd, n_layer, qtype=0) -> int:
    """Q4 bytes one decode token streams (every 2-D tensor except the embedding matrix, which is read one row of)."""
    per = 20 if qtype == 0 else 24
    blocks = n_layer * (4 * n_embd * n_embd + 8 * n_embd * n_embd) // 32 + n_vocab * n_embd // 32
    return blocks * per
