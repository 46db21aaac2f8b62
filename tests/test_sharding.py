"""Column sharding: host logic on CPU (world_size-2 gloo) and the fused peer-store epilogue on one GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import qllm_b200
from oracle import qlinear_oracle as O
from qllm_b200 import sharding
from tests.util import layer_from_dict, oracle_forward, rel_err


def test_shard_cols_partition():
    for N, world, gran in [(4096, 8, 32), (11008, 8, 32), (11008, 4, 32), (4096, 2, 64), (13824, 8, 64)]:
        rs = [sharding.shard_cols(N, world, r, gran) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert all((c1 - c0) % gran == 0 for c0, c1 in rs)
        w = [c1 - c0 for c0, c1 in rs]
        assert max(w) - min(w) <= gran


@pytest.mark.parametrize("layout,bits,gs", [("GPTQ", 4, 64), ("GPTQ", 2, 64), ("GPTQ", 3, 32), ("HQQ", 4, 64), ("GEMM", 4, 64),
                                             ("MARLIN", 4, 128)])
def test_shard_layer_slices_are_the_right_columns(layout, bits, gs):
    K, N, world = 256, 512, 4
    L = O.make_layer(layout, bits, gs, K, N, seed=9, bias=(layout != "MARLIN"), act_order=(layout == "GPTQ" and bits == 4))
    full = layer_from_dict(L, device="cpu")
    for r in range(world):
        sh = sharding.shard_layer(full, r, world)
        c0, c1 = sh.col0, sh.col0 + sh.outfeatures
        q, z, s, gi = O.unpack_layer(layout, bits, gs, K, c1 - c0, sh.qweight.numpy(),
                                     None if sh.qzeros is None else sh.qzeros.numpy(), sh.scales.numpy(),
                                     sh.g_idx.numpy() if layout == "GPTQ" else None)
        assert np.array_equal(q, L["q"][:, c0:c1])
        assert np.array_equal(np.asarray(z), np.asarray(L["z"])[:, c0:c1])
        assert np.array_equal(s.view(np.uint16), L["s"][:, c0:c1].view(np.uint16))
        if L["bias"] is not None:
            assert np.array_equal(sh.bias.numpy(), L["bias"][c0:c1])


def _gloo_worker(rank, world, port, N, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K = 128
        L = O.make_layer("GPTQ", 4, 64, K, N, seed=3)
        full = layer_from_dict(L, device="cpu")
        local = sharding.shard_layer(full, rank, world)
        W = O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "engine").astype(np.float32)
        c0, c1 = local.col0, local.col0 + local.outfeatures
        # stand-in for the CUDA kernel on CPU ranks: the TEST computes the shard with the oracle
        local.forward = lambda x: (x.float() @ torch.from_numpy(W[:, c0:c1])).to(x.dtype)
        mod = sharding.ColumnShardedLinear(local, N, rank, world)
        x = torch.randn(3, K, generator=torch.Generator().manual_seed(0))
        y = mod(x)
        ref = x.float() @ torch.from_numpy(W)
        ret[rank] = float((y.float() - ref).abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [256, 160])          # equal shards, and uneven (5 tiles over 2 ranks)
def test_column_sharded_linear_world2_gloo(N):
    world = 2
    ret = mp.Manager().dict()
    port = 29500 + (os.getpid() % 2000) + (1 if N == 160 else 0)
    mp.spawn(_gloo_worker, args=(world, port, N, ret), nprocs=world, join=True)
    assert len(ret) == world and all(v < 1e-5 for v in ret.values())


@pytest.mark.gpu
@pytest.mark.parametrize("layout,M", [("GEMM", 1), ("GPTQ", 4), ("GPTQ", 64), ("MARLIN", 2)])
def test_fused_peer_store_epilogue(layout, M):
    """b200q_linear_sharded writes this shard's columns into every peer buffer (two local buffers here)."""
    K, N, world = 512, 1024, 4
    gs = 128
    L = O.make_layer(layout, 4, gs, K, N, seed=21)
    full = layer_from_dict(L, device="cpu")
    x = np.random.default_rng(1).standard_normal((M, K)).astype(np.float16)
    xd = torch.from_numpy(x).cuda()
    outs = [torch.zeros(M, N, dtype=torch.float16, device="cuda") for _ in range(2)]
    for r in range(world):
        local = sharding.shard_layer(full, r, world).cuda()
        sharding.sharded_forward_into_peers(local, xd, outs, local.col0)
    torch.cuda.synchronize()
    ref = oracle_forward(L, x)
    for o in outs:
        assert rel_err(o.float().cpu().numpy(), ref) < 1e-3
    assert torch.equal(outs[0], outs[1])
