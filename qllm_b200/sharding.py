"""Column (N) sharding of a QuantLinear across the GPUs of one node.

Every output column depends on one column of qweight/qzeros/scales and on all of x, so rank r keeps a
contiguous column range and the only exchange is one all-gather of y[:, shard] per layer.  Shard
boundaries are multiples of 32 columns (64 for Marlin) so packed qzeros words, AWQ words and Marlin
64-column blocks are never split.  The reference has no multi-GPU path.
"""
import ctypes

import torch
import torch.nn as nn

from ._lib import Layer, PeerSync, check, lib
from .q_layers import (QuantLinearGPTQ, QuantLinearHQQ, QuantLinearMarlin, WQLinear_GEMM, _workspace)


def shard_cols(N: int, world: int, rank: int, gran: int = 32):
    """[c0, c1) owned by `rank`: whole `gran`-column tiles, balanced to within one tile."""
    if N % gran:
        raise ValueError(f"N={N} is not a multiple of the shard granularity {gran}")
    tiles = N // gran
    return tiles * rank // world * gran, tiles * (rank + 1) // world * gran


def shard_granularity(layer) -> int:
    return 64 if isinstance(layer, QuantLinearMarlin) else 32


def shard_layer(layer, rank: int, world: int):
    """A new layer of the same class holding columns shard_cols(N, world, rank) of `layer`'s buffers."""
    N, K, b = layer.outfeatures, layer.infeatures, layer.bits
    c0, c1 = shard_cols(N, world, rank, shard_granularity(layer))
    n = c1 - c0
    has_bias = layer.bias is not None
    gs = layer.groupsize
    if isinstance(layer, QuantLinearMarlin):
        new = QuantLinearMarlin.__new__(QuantLinearMarlin)
        nn.Module.__init__(new)
        new._init_common(b, gs, K, n, torch.float16)
        new.group_size, new.pack_mode, new.g_idx, new.qzeros = gs, "MARLIN", None, None
        new.register_buffer("qweight", layer.qweight[:, 2 * c0:2 * c1].contiguous())
        new.register_buffer("scales", layer.scales[:, c0:c1].contiguous())
    elif isinstance(layer, WQLinear_GEMM):
        new = WQLinear_GEMM(b, gs, K, n, has_bias, dtype=layer.dtype)
        new.qweight = layer.qweight[:, c0 // 8:c1 // 8].contiguous()
        new.qzeros = layer.qzeros[:, c0 // 8:c1 // 8].contiguous()
        new.scales = layer.scales[:, c0:c1].contiguous()
    elif isinstance(layer, QuantLinearHQQ):
        new = QuantLinearHQQ(b, gs, K, n, has_bias, dtype=layer.dtype)
        new.qweight = layer.qweight[:, c0:c1].contiguous()
        new.qzeros = layer.qzeros[:, c0:c1].contiguous()
        new.scales = layer.scales[:, c0:c1].contiguous()
    elif isinstance(layer, QuantLinearGPTQ):
        new = QuantLinearGPTQ(b, gs, K, n, has_bias, dtype=layer.dtype)
        new.qweight = layer.qweight[:, c0:c1].contiguous()
        new.qzeros = layer.qzeros[:, c0 * b // 32:c1 * b // 32].contiguous()
        new.scales = layer.scales[:, c0:c1].contiguous()
        new.g_idx = layer.g_idx.clone()                     # replicated
        new.zero_bias = layer.zero_bias
    else:
        raise TypeError(type(layer))
    new.bias = layer.bias[c0:c1].contiguous() if has_bias else None
    new.col0, new.full_n = c0, N
    return new


class ColumnShardedLinear(nn.Module):
    """y = all_gather_N( x @ W[:, shard] ).  `local` is this rank's shard (see shard_layer)."""

    def __init__(self, local, full_n: int, rank: int, world: int, group=None):
        super().__init__()
        self.local, self.full_n, self.rank, self.world, self.group = local, full_n, rank, world, group
        gran = shard_granularity(local)
        self.ranges = [shard_cols(full_n, world, r, gran) for r in range(world)]
        self.equal = len({c1 - c0 for c0, c1 in self.ranges}) == 1

    def forward(self, x):
        import torch.distributed as dist
        part = self.local(x)                                   # [..., n_local] through the C ABI
        lead = part.shape[:-1]
        part2 = part.reshape(-1, part.shape[-1])
        M = part2.shape[0]
        if self.world == 1:
            return part
        wmax = max(c1 - c0 for c0, c1 in self.ranges)
        if part2.shape[1] != wmax:
            part2 = torch.nn.functional.pad(part2, (0, wmax - part2.shape[1]))
        flat = torch.empty(self.world * M, wmax, dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(flat, part2.contiguous(), group=self.group)
        gathered = flat.view(self.world, M, wmax)
        if self.equal:
            y = gathered.permute(1, 0, 2).reshape(M, self.full_n)
        else:
            y = torch.cat([gathered[r, :, : c1 - c0] for r, (c0, c1) in enumerate(self.ranges)], dim=1)
        return y.reshape(lead + (self.full_n,))


def sharded_forward_into_peers(local, x, peer_outputs, n_offset: int):
    """Fused-epilogue form (b200q_linear_sharded): compute this rank's columns and store them at
    column n_offset of every buffer in `peer_outputs` ([M, ldy] fp16 tensors; peers' buffers must be mapped
    into this process, e.g. torch.distributed._symmetric_memory).  The caller owns the cross-rank barrier."""
    desc = local._descriptor()
    x2 = x.reshape(-1, x.shape[-1])
    M = x2.shape[0]
    ptrs = (ctypes.c_void_p * len(peer_outputs))(*[t.data_ptr() for t in peer_outputs])
    need = lib.b200q_workspace_bytes(ctypes.byref(desc), M)
    ws = _workspace(x.device, need)
    st = lib.b200q_linear_sharded(ctypes.byref(desc), x2.data_ptr(), M, x2.stride(0), ptrs, len(peer_outputs),
                                  peer_outputs[0].stride(0), n_offset, ws.data_ptr(), ws.numel(),
                                  torch.cuda.current_stream(x.device).cuda_stream)
    check(st, "b200q_linear_sharded")


# ---------------------------------------------------------------------------------------------------------------
# Fused all-gather + cross-GPU hand-off (b200q_linear_group_sharded): no collective kernel between layers.

class PeerArena:
    """One symmetric-memory allocation per rank, mapped into every rank of the node over NVLink
    (torch.distributed._symmetric_memory): `n_slots` u64 hand-off counters followed by caller-carved output replicas.
    ptr(r, off) is the address of byte `off` of rank r's arena as seen from this process."""
    COUNTER_BYTES = 8

    def __init__(self, payload_bytes: int, n_slots: int = 1024, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.counter_bytes = (n_slots * self.COUNTER_BYTES + 255) & ~255
        self.nbytes = self.counter_bytes + ((payload_bytes + 255) & ~255)
        self.buf = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group.group_name if hasattr(self.group, "group_name") else self.group)
        self.base = [int(p) for p in self.hdl.buffer_ptrs]
        if len(self.base) != self.world or self.base[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned inconsistent buffer pointers")
        self.buf.zero_()
        torch.cuda.synchronize()
        dist.barrier(self.group)
        self._cursor = self.counter_bytes
        # word 0: local step number; words 1..: per-call execution counts (B200Q_PEER_NODE_EPOCH, maintained by the kernels)
        self.epoch = torch.zeros(4096, dtype=torch.int64, device=self.device)
        self._counters = (ctypes.c_void_p * self.world)(*self.base)                 # counters sit at offset 0

    def carve(self, nbytes: int) -> int:
        """Reserve `nbytes` (256-byte aligned) of payload; returns the offset, identical on every rank."""
        off = self._cursor
        self._cursor += (nbytes + 255) & ~255
        if self._cursor > self.nbytes:
            raise MemoryError("PeerArena payload exhausted")
        return off

    def ptr(self, r: int, off: int) -> int:
        return self.base[r] + off

    def local_view(self, off: int, shape, dtype=torch.float16):
        n = 1
        for s in shape:
            n *= s
        return self.buf[off:off + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)

    def sync_desc(self, wait_slot: int = -1, wait_count: int = 0, post_slot: int = -1, flags: int = 0,
                  tag_stride: int = 1, y_seq: int = 0, x_seq: int = 0) -> PeerSync:
        s = PeerSync()
        s.n_peers, s.self_rank = self.world, self.rank
        s.counters = ctypes.cast(self._counters, ctypes.POINTER(ctypes.c_void_p))
        s.epoch = self.epoch.data_ptr()
        s.wait_slot, s.wait_count, s.post_slot, s.flags = wait_slot, wait_count, post_slot, flags
        s.tag_stride, s.y_seq, s.x_seq = tag_stride, y_seq, x_seq
        return s

    def untag(self, off: int, M: int, N: int, out: torch.Tensor, tag_stride: int, seq: int, stream=None):
        """out[M, N] fp16 = the tagged replica at arena offset `off`, once every word carries the tag of call `seq`."""
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        s = self.sync_desc(tag_stride=tag_stride, x_seq=seq)
        check(lib.b200q_peer_untag(self.ptr(self.rank, off), N, out.data_ptr(), out.stride(0), M, N, ctypes.byref(s), st),
              "b200q_peer_untag")

    def advance(self, stream=None):
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        check(lib.b200q_peer_epoch_advance(self.epoch.data_ptr(), st), "b200q_peer_epoch_advance")

    def wait(self, wait_slot: int, wait_count: int, stream=None):
        """Make `stream` wait until the awaited output is complete locally (for consumers outside the engine)."""
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        s = self.sync_desc(wait_slot, wait_count, -1)
        check(lib.b200q_peer_wait(ctypes.byref(s), st), "b200q_peer_wait")

    def poisoned(self) -> bool:
        """True when a kernel gave up waiting for a peer (2 s) -- the outputs of that step are invalid."""
        return bool(self.buf[:8].view(torch.int64).item() != 0)


def sharded_group_posts(local_layers, M: int) -> int:
    """Storing CTAs (= posts per peer) of one fused call on this rank's shards."""
    descs = [l._decode_descriptor(M) for l in local_layers]
    arr = (ctypes.POINTER(Layer) * len(descs))(*[ctypes.pointer(d) for d in descs])
    n = lib.b200q_sharded_posts(arr, len(descs), M)
    check(min(n, 0), "b200q_sharded_posts")
    return n


class LocalArena(PeerArena):
    """The same arena on ONE GPU (world 1, plain device memory, no process group): lets a chain of QuantLinears hand its
    activations over as tagged words on a single device -- the consumer kernel's load stage then follows the data, not
    the kernel boundary (B200Q_PEER_NODE_EPOCH)."""

    def __init__(self, payload_bytes: int, n_slots: int = 1024, device=None):
        self.group, self.rank, self.world = None, 0, 1
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.counter_bytes = (n_slots * self.COUNTER_BYTES + 255) & ~255
        self.nbytes = self.counter_bytes + ((payload_bytes + 255) & ~255)
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        self.base = [self.buf.data_ptr()]
        self._cursor = self.counter_bytes
        self.epoch = torch.zeros(4096, dtype=torch.int64, device=self.device)
        self._counters = (ctypes.c_void_p * 1)(*self.base)


def sharded_group_forward(arena: PeerArena, local_layers, x, out_offsets, full_ns, col0s, sync: PeerSync, ws, stream=None,
                          x_ptr=None, M=None, ldx=None):
    """One launch: this rank's column shards of sibling layers `local_layers` (shared x [M, K]) stored at column
    col0s[i] of the replicas at arena offset out_offsets[i] ([M, full_ns[i]] fp16, or uint32 words with
    PEER_Y_TAGGED) on EVERY rank, with the hand-off described by `sync` (PeerArena.sync_desc).  Tagged activations
    (PEER_X_TAGGED) are passed as x_ptr / M / ldx (a uint32 [M, ldx] replica)."""
    n, P = len(local_layers), arena.world
    if x_ptr is None:                                       # plain fp16 activations
        x2 = x.reshape(-1, x.shape[-1])
        M, x_ptr, ldx = x2.shape[0], x2.data_ptr(), x2.stride(0)
    descs = [l._decode_descriptor(M) for l in local_layers]
    arr = (ctypes.POINTER(Layer) * n)(*[ctypes.pointer(d) for d in descs])
    yp = (ctypes.c_void_p * (n * P))(*[arena.ptr(r, out_offsets[i]) for i in range(n) for r in range(P)])
    ld = (ctypes.c_int64 * n)(*full_ns)
    no = (ctypes.c_int64 * n)(*col0s)
    st = torch.cuda.current_stream(arena.device).cuda_stream if stream is None else stream
    check(lib.b200q_linear_group_sharded(arr, n, x_ptr, M, ldx, yp, ld, no, ctypes.byref(sync),
                                         ws.data_ptr(), ws.numel(), st), "b200q_linear_group_sharded")
