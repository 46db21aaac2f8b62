"""Column (N) sharding of a QuantLinear across the GPUs of one node.

Every output column depends on one column of qweight/qzeros/scales and on all of x, so rank r keeps a
contiguous column range and the only exchange is one all-gather of y[:, shard] per layer.  Shard
boundaries are multiples of 32 columns (64 for Marlin) so packed qzeros words, AWQ words and Marlin
64-column blocks are never split.  The reference has no multi-GPU path.
"""
import ctypes

import torch
import torch.nn as nn

from ._lib import check, lib
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
