"""Row-sharded Q4 mul_mat across the GPUs of one node.

The split is the one the reference's threads use (imax.c:1185-1188): rank g of G owns weight rows
[g*N/G, (g+1)*N/G) of every tensor, X is replicated (each rank quantizes it locally, bit-identically),
and the fp32 output slices are all-gathered.  No reduction crosses GPUs, so the result equals the
single-GPU one bit for bit.

Two exchange paths:
  * PeerOutput    -- the product path on NVLink: symmetric (peer-mapped) Y buffers; the mul_mat kernel's
                     epilogue stores its slice into every rank's Y and publishes a sequence flag
                     (b200q4_mul_mat_dev_allgather + b200q4_wait_flags).  No collective call per mul_mat.
  * all_gather_rows -- torch.distributed all-gather of [M][N/G] slices (NCCL on GPUs, gloo in the CPU
                     tests of this host logic); the baseline the fused path is measured against.
torch is used for device memory and process groups only.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from .lib import B200Q4Error, _err


def shard_rows(n: int, world: int, rank: int) -> tuple[int, int]:
    """(first row, row count) of rank's shard.  N must divide evenly (every listed model does)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    if n % world:
        raise ValueError(f"{n} weight rows do not split evenly over {world} ranks: this tensor stays replicated")
    return rank * (n // world), n // world


def assemble(slices: torch.Tensor) -> torch.Tensor:
    """[world][M][N/world] gathered slices -> Y [M][N] (rank r's rows are output columns [r*N/G, (r+1)*N/G))."""
    world, m, ns = slices.shape
    return slices.permute(1, 0, 2).reshape(m, world * ns)


def all_gather_rows(y_local: torch.Tensor, group=None) -> torch.Tensor:
    """Collective baseline: all-gather the [M][N/world] slices and assemble Y [M][N] on every rank."""
    world = dist.get_world_size(group)
    m, ns = y_local.shape
    out = torch.empty((world * m, ns), dtype=y_local.dtype, device=y_local.device)       # rank-major concatenation
    dist.all_gather_into_tensor(out, y_local.contiguous(), group=group)
    return assemble(out.view(world, m, ns))


class PeerOutput:
    """Symmetric Y buffers + arrival flags for the fused all-gather epilogue.

    Two Y buffers alternate: rank A may begin storing exchange s+1 into rank B while B still reads
    exchange s as its next input; A cannot reach exchange s+2 (the same buffer again) before B has
    finished s+1, because every call first waits for the previous exchange to have landed from all ranks.
    The exchange counter lives on the device (calls can be captured in a CUDA graph); this object mirrors
    it on the host to know which buffer a call's result is in."""

    def __init__(self, lib, max_m: int, n_total: int, group=None):
        import torch.distributed._symmetric_memory as symm
        self.L = lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("at most 8 ranks (one NVSwitch domain)")
        self.max_m, self.n_total = max_m, n_total
        dev = torch.device("cuda", torch.cuda.current_device())
        self.y, self.y_hdl = [], []
        for _ in range(2):
            t = symm.empty(max_m * n_total, dtype=torch.float32, device=dev)
            self.y.append(t)
            self.y_hdl.append(symm.rendezvous(t, self.group))
        self.flags = symm.empty(16, dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.f_hdl = symm.rendezvous(self.flags, self.group)
        self.seq_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier(self.group)
        vp = C.c_void_p
        self._yptrs = (vp * (2 * self.world))(*[int(p) for h in self.y_hdl for p in h.buffer_ptrs])     # [buffer][rank]
        self._fptrs = (vp * self.world)(*[int(p) for p in self.f_hdl.buffer_ptrs])
        self.seq = 0                                 # host mirror of the device counter
        self._st = torch.cuda.Stream(device=dev)     # used when the caller is on the legacy default stream

    def _stream(self, stream):
        cur = stream if stream is not None else torch.cuda.current_stream()
        # The C ABI reads a NULL stream as "the library's own stream", which is not ordered with torch's
        # legacy default stream: hop onto a private stream, ordered both ways with the caller's.
        hop = cur.cuda_stream == 0
        if hop:
            self._st.wait_stream(cur)
        return cur, hop, (self._st.cuda_stream if hop else cur.cuda_stream)

    def _check(self, rc):
        if rc:
            from .lib import B200Q4Error, _err
            raise B200Q4Error(f"fused all-gather call failed (rc={rc}): {_err(self.L)}")

    def mul_mat(self, w, d_x: torch.Tensor, m: int, n_total: int | None = None, stream=None, wait: bool = True) -> torch.Tensor:
        """Y[M][n_total] = mul_mat over the row shards of all ranks; `w` is this rank's shard (n_total/world rows).
        wait=False: the next call on this object consumes the result (it waits for the exchange itself);
        wait=True: the stream is held until every rank's slice has landed (for any other consumer)."""
        n_total = n_total or self.n_total
        if m > self.max_m or n_total > self.n_total or w.n * self.world != n_total:
            raise ValueError("shape does not fit the symmetric buffers / shard size")
        cur, hop, st = self._stream(stream)
        self.seq += 1
        b = self.seq & 1
        self._check(self.L.b200q4_mul_mat_dev_allgather(w.handle, d_x.data_ptr(), m, w.k * 4, self._yptrs, self.world, self.rank,
                                                        self.rank * w.n, n_total * 4, self._fptrs, self.seq_ctr.data_ptr(), st))
        if wait:
            self._check(self.L.b200q4_wait_flags(self.flags.data_ptr(), self.world, self.seq_ctr.data_ptr(), st))
        if hop:
            cur.wait_stream(self._st)
        return self.y[b][: m * n_total].view(m, n_total)

    def wait(self, stream=None):
        """Hold the stream until the latest exchange has landed from every rank."""
        cur, hop, st = self._stream(stream)
        self._check(self.L.b200q4_wait_flags(self.flags.data_ptr(), self.world, self.seq_ctr.data_ptr(), st))
        if hop:
            cur.wait_stream(self._st)


class ShardedNeox:
    """b200neox_load_sharded behind torch symmetric memory: one process per GPU, every rank calls eval() with the same
    token; decode steps are ONE persistent launch per rank, the ranks' thread blocks exchange their output slices
    directly ({value, tag} stores into every rank's pool over NVLink)."""

    def __init__(self, lib, path: str, n_ctx: int = 512, group=None):
        import struct
        import torch.distributed._symmetric_memory as symm
        self.L = lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        with open(path, "rb") as f:
            magic, n_vocab, n_embd, n_head, n_layer = struct.unpack("<5i", f.read(20))
        if magic != 0x67676d6c:
            raise ValueError("not a ggml model file")
        self.n_vocab = n_vocab
        need = int(lib.b200neox_sharded_pool_bytes(n_vocab, n_embd, n_layer))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.pool = symm.empty((need + 3) // 4, dtype=torch.int32, device=dev)
        self.pool.zero_()
        self.hdl = symm.rendezvous(self.pool, self.group)
        torch.cuda.synchronize()
        dist.barrier(self.group)
        ptrs = (C.c_void_p * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        self.m = lib.b200neox_load_sharded(str(path).encode(), n_ctx, self.rank, self.world, ptrs, need)
        if not self.m:
            raise RuntimeError((lib.b200neox_last_error() or b"").decode(errors="replace"))
        torch.cuda.synchronize()
        dist.barrier(self.group)                          # nobody launches before every rank's program and counters exist
        import numpy as np
        self._out = np.empty(n_vocab, np.float32)
        self._tok = np.zeros(1, np.int32)

    def eval(self, token: int, n_past: int):
        """One decode step on every rank; returns the logits (host numpy view, valid until the next call)."""
        self._tok[0] = token
        rc = self.L.b200neox_eval(self.m, self._tok.ctypes.data, 1, n_past, 1, self._out.ctypes.data)
        if rc:
            raise RuntimeError((self.L.b200neox_last_error() or b"").decode(errors="replace"))
        return self._out

    def free(self):
        if getattr(self, "m", None):
            torch.cuda.synchronize()
            dist.barrier(self.group)                      # no rank frees its pool while a peer's kernel may still store into it
            self.L.b200neox_free(self.m)
            self.m = None


class ShardedChain:
    """b200q4_chain_create_sharded behind torch symmetric memory: one process per GPU; op i is this rank's row shard of a
    tensor with world * N_i rows, and one persistent launch per rank runs the whole list -- the reducer warps store every
    output slice as {value, tag} pairs into EVERY rank's pool over NVLink (the all-gather is part of the mul_mat)."""

    def __init__(self, lib, weights, x_ptrs, group=None):
        import torch.distributed._symmetric_memory as symm
        self.L = lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        n = self.n = len(weights)
        off = int(lib.b200q4_chain_sharded_ctr_bytes(n))
        self.offs, self.rows = [], []
        for w in weights:
            self.offs.append(off)
            self.rows.append(w.n * self.world)
            off += (2 * 4 * w.n * self.world + 255) & ~255
        dev = torch.device("cuda", torch.cuda.current_device())
        self.pool = symm.empty(off // 4, dtype=torch.int32, device=dev)
        self.pool.zero_()
        self.hdl = symm.rendezvous(self.pool, self.group)
        torch.cuda.synchronize()
        dist.barrier(self.group)
        ptrs = (C.c_void_p * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        ws = (C.c_void_p * n)(*[w.handle for w in weights])
        xs = (C.c_void_p * n)(*x_ptrs)
        yo = (C.c_size_t * n)(*self.offs)
        self.handle = lib.b200q4_chain_create_sharded(n, ws, xs, self.rank, self.world, ptrs, off, yo)
        if not self.handle:
            raise B200Q4Error(f"b200q4_chain_create_sharded failed: {_err(lib)}")
        torch.cuda.synchronize()
        dist.barrier(self.group)                          # nobody launches before every rank's list and counters exist

    def run(self, stream=None) -> None:
        rc = self.L.b200q4_chain_run(self.handle, stream)
        if rc:
            raise B200Q4Error(f"b200q4_chain_run failed (rc={rc}): {_err(self.L)}")

    def gathered(self, i: int) -> torch.Tensor:
        """The full output vector of op i as it stands in THIS rank's pool (values of the {value, tag} pairs)."""
        v = self.pool.view(torch.float32)[self.offs[i] // 4: self.offs[i] // 4 + 2 * self.rows[i]]
        return v[0::2]

    def free(self):
        if getattr(self, "handle", None):
            torch.cuda.synchronize()
            dist.barrier(self.group)                      # no rank frees its pool while a peer's kernel may still store into it
            self.L.b200q4_chain_free(self.handle)
            self.handle = None
