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


class _VirtualArena:
    """Two 'ranks' on one GPU for the hand-off protocol: plain device buffers stand in for symmetric memory.  Calls are
    issued rank after rank on ONE stream, so every wait is already satisfied when its kernel runs (a spinning kernel
    could otherwise starve its producer of SMs on a single device)."""

    def __init__(self, world, payload):
        import ctypes
        from qllm_b200._lib import PeerSync
        self.world = world
        self.bufs = [torch.zeros(2048 + payload, dtype=torch.uint8, device="cuda") for _ in range(world)]
        self.epochs = [torch.zeros(16, dtype=torch.int64, device="cuda") for _ in range(world)]
        self._ctr = (ctypes.c_void_p * world)(*[b.data_ptr() for b in self.bufs])
        self._ct, self._PS = ctypes, PeerSync

    def sync(self, rank, wait_slot, wait_count, post_slot):
        s = self._PS()
        s.n_peers, s.self_rank = self.world, rank
        s.counters = self._ct.cast(self._ctr, self._ct.POINTER(self._ct.c_void_p))
        s.epoch = self.epochs[rank].data_ptr()
        s.wait_slot, s.wait_count, s.post_slot = wait_slot, wait_count, post_slot
        return s

    def counter(self, rank, slot):
        return int(self.bufs[rank][:2048].view(torch.int64)[slot].item())


@pytest.mark.gpu
@pytest.mark.parametrize("layout,M,Ns", [("GPTQ", 1, (1024, 1024, 1024)), ("GEMM", 2, (1536, 1536)), ("GPTQ", 1, (512,)),
                                         ("GPTQ", 1, (192, 192))])         # 96-column shards: one and a half tiles
def test_fused_handoff_group_sharded_virtual_ranks(layout, M, Ns):
    """b200q_linear_group_sharded: shards of sibling layers land in every replica, every storing CTA posts once on
    every peer, a consumer that waits on the slot sees epoch * posts, counters only grow across steps, slot 0 (the
    time-out poison) stays clear; results equal the oracle."""
    import ctypes
    from qllm_b200 import Layer, check, lib
    K, gs, world = 512, 128, 2
    Ls = [O.make_layer(layout, 4, gs, K, N, seed=31 + i) for i, N in enumerate(Ns)]
    fulls = [layer_from_dict(L, device="cpu") for L in Ls]
    shards = [[sharding.shard_layer(f, r, world).cuda() for f in fulls] for r in range(world)]
    x = np.random.default_rng(3).standard_normal((M, K)).astype(np.float16)
    xd = torch.from_numpy(x).cuda()
    payload = sum(M * N * 2 + 256 for N in Ns)
    A = _VirtualArena(world, payload)
    offs, o = [], 2048
    for N in Ns:
        offs.append(o)
        o += (M * N * 2 + 255) & ~255
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    posts = [sharding.sharded_group_posts(shards[r], M) for r in range(world)]
    assert all(p > 0 for p in posts)
    n = len(Ns)
    for step in (1, 2, 3):
        for r in range(world):
            check(lib.b200q_peer_epoch_advance(A.epochs[r].data_ptr(), st))
        for r in range(world):
            descs = [l._decode_descriptor(M) for l in shards[r]]
            arr = (ctypes.POINTER(Layer) * n)(*[ctypes.pointer(d) for d in descs])
            yp = (ctypes.c_void_p * (n * world))(*[A.bufs[q].data_ptr() + offs[i] for i in range(n) for q in range(world)])
            ld = (ctypes.c_int64 * n)(*Ns)
            no = (ctypes.c_int64 * n)(*[l.col0 for l in shards[r]])
            s = A.sync(r, -1, 0, 5)
            check(lib.b200q_linear_group_sharded(arr, n, xd.data_ptr(), M, xd.stride(0), yp, ld, no, ctypes.byref(s),
                                                 ws.data_ptr(), ws.numel(), st))
        for r in range(world):                      # consumers: wait on slot 5 (already complete), post nothing
            s = A.sync(r, 5, posts[1 - r], -1)
            check(lib.b200q_peer_wait(ctypes.byref(s), st))
            descs = [shards[r][0]._decode_descriptor(M)]
            arr = (ctypes.POINTER(Layer) * 1)(ctypes.pointer(descs[0]))
            scratch = torch.zeros(M, Ns[0], dtype=torch.float16, device="cuda")
            yp = (ctypes.c_void_p * world)(*[scratch.data_ptr()] * world)
            ld, no = (ctypes.c_int64 * 1)(Ns[0]), (ctypes.c_int64 * 1)(shards[r][0].col0)
            check(lib.b200q_linear_group_sharded(arr, 1, xd.data_ptr(), M, xd.stride(0), yp, ld, no, ctypes.byref(s),
                                                 ws.data_ptr(), ws.numel(), st))
        torch.cuda.synchronize()
        for r in range(world):
            assert A.counter(r, 5) == step * posts[1 - r]
            assert A.counter(r, 0) == 0
    for i, (L, N) in enumerate(zip(Ls, Ns)):
        ref = oracle_forward(L, x)
        outs = [A.bufs[r][offs[i]:offs[i] + M * N * 2].view(torch.float16).view(M, N) for r in range(world)]
        assert torch.equal(outs[0], outs[1])
        assert rel_err(outs[0].float().cpu().numpy(), ref) < 1e-3


@pytest.mark.gpu
def test_peer_wait_times_out_instead_of_hanging():
    """A wait whose producer never posts gives up after its bound and poisons slot 0 (no GPU hang)."""
    import ctypes
    from qllm_b200 import check, lib
    A = _VirtualArena(2, 0)
    check(lib.b200q_peer_epoch_advance(A.epochs[0].data_ptr(), torch.cuda.current_stream().cuda_stream))
    s = A.sync(0, 7, 3, -1)
    check(lib.b200q_peer_wait(ctypes.byref(s), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert A.counter(0, 0) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("node_epoch", [0, 4])          # 4 = B200Q_PEER_NODE_EPOCH: the step comes from the call's own word
@pytest.mark.parametrize("layout,M,act", [("GPTQ", 1, False), ("GEMM", 2, False), ("GPTQ", 1, True), ("GPTQ", 2, True)])
def test_tagged_activation_chain_virtual_ranks(layout, M, act, node_epoch):
    """Flag-in-data hand-off: layer A's shards are written as tagged words (fp16 | step tag << 16) into every replica,
    layer B reads the tagged replica as its x, b200q_peer_untag returns plain fp16.  Two steps: the tag follows the
    epoch.  Checked against the oracle of the two-layer chain (x_B = fp16(y_A))."""
    import ctypes
    from qllm_b200 import Layer, check, lib
    from qllm_b200._lib import PEER_X_TAGGED, PEER_Y_TAGGED
    K, NA, NB, gs, world = 512, 1024, 512, 128, 2
    # act: the consumer is a desc_act layer -- it gathers its tagged words through x_perm
    LA, LB = O.make_layer(layout, 4, gs, K, NA, seed=41), O.make_layer(layout, 4, gs, NA, NB, seed=42, act_order=act)
    fA, fB = layer_from_dict(LA, device="cpu"), layer_from_dict(LB, device="cpu")
    sA = [sharding.shard_layer(fA, r, world).cuda() for r in range(world)]
    sB = [sharding.shard_layer(fB, r, world).cuda() for r in range(world)]
    A = _VirtualArena(world, 2 * M * (NA + NB) * 4 + 1024)
    offA, offB = 2048, 2048 + ((M * NA * 4 + 255) & ~255)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(5)

    def call(r, shard, x_ptr, ldx, off, N, flags, y_seq=0, x_seq=0):
        d = shard._decode_descriptor(M)
        arr = (ctypes.POINTER(Layer) * 1)(ctypes.pointer(d))
        yp = (ctypes.c_void_p * world)(*[A.bufs[q].data_ptr() + off for q in range(world)])
        ld, no = (ctypes.c_int64 * 1)(N), (ctypes.c_int64 * 1)(shard.col0)
        s = A.sync(r, -1, 0, -1)
        s.flags, s.tag_stride, s.y_seq, s.x_seq = flags | node_epoch, 3, y_seq, x_seq
        check(lib.b200q_linear_group_sharded(arr, 1, x_ptr, M, ldx, yp, ld, no, ctypes.byref(s), ws.data_ptr(), ws.numel(), st))

    for step in (1, 2, 3):
        x = rng.standard_normal((M, K)).astype(np.float16)
        xd = torch.from_numpy(x).cuda()
        for r in range(world):
            check(lib.b200q_peer_epoch_advance(A.epochs[r].data_ptr(), st))
        for r in range(world):
            call(r, sA[r], xd.data_ptr(), xd.stride(0), offA, NA, PEER_Y_TAGGED, y_seq=1)
        for r in range(world):
            call(r, sB[r], A.bufs[r].data_ptr() + offA, NA, offB, NB, PEER_Y_TAGGED | PEER_X_TAGGED, y_seq=2, x_seq=1)
        outs = []
        for r in range(world):
            out = torch.zeros(M, NB, dtype=torch.float16, device="cuda")
            s = A.sync(r, -1, 0, -1)
            s.tag_stride, s.x_seq = 3, 2
            check(lib.b200q_peer_untag(A.bufs[r].data_ptr() + offB, NB, out.data_ptr(), NB, M, NB, ctypes.byref(s), st))
            outs.append(out)
        torch.cuda.synchronize()
        words = A.bufs[0][offA:offA + M * NA * 4].view(torch.int32).cpu().numpy().astype(np.uint32)
        assert np.all(words >> 16 == 3 * step + 1)
        yA = oracle_forward(LA, x).astype(np.float16)
        assert rel_err((words & 0xffff).astype(np.uint16).view(np.float16).reshape(M, NA), oracle_forward(LA, x)) < 1e-3
        ref = oracle_forward(LB, (words & 0xffff).astype(np.uint16).view(np.float16).reshape(M, NA))
        assert torch.equal(outs[0], outs[1])
        assert rel_err(outs[0].float().cpu().numpy(), ref) < 1e-3
        assert A.counter(0, 0) == 0 and A.counter(1, 0) == 0
        if node_epoch:                      # each call's own word follows the step; the arrival words are back at zero
            assert [int(v) for v in A.epochs[0][:4].cpu()] == [step, 0, step, step]
            assert int(ws[4000:4016].view(torch.int32).abs().sum().item()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("K,N,world", [(4096, 4096, 8), (4096, 11008, 8), (11008, 4096, 8), (4096, 11008, 4)])
def test_tagged_shards_at_llama7b_shapes_all_ranks_on_one_gpu(K, N, world):
    """The column shards bench.py --gpus 8 / 4 runs (512-, 1376- = 21.5-tile and 2752-column shards, K up to 11008), every
    rank executed in turn on one GPU: all shards land tagged in both checked replicas and equal the oracle."""
    import ctypes
    from qllm_b200 import Layer, check, lib
    from qllm_b200._lib import PEER_Y_TAGGED
    gs, M = 128, 1
    L = O.make_layer("GEMM", 4, gs, K, N, seed=K + N + world)
    full = layer_from_dict(L, device="cpu")
    x = np.random.default_rng(9).standard_normal((M, K)).astype(np.float16)
    xd = torch.from_numpy(x).cuda()
    A = _VirtualArena(world, M * N * 4 + 256)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for r in range(world):
        check(lib.b200q_peer_epoch_advance(A.epochs[r].data_ptr(), st))
    for r in range(world):
        shard = sharding.shard_layer(full, r, world).cuda()
        d = shard._decode_descriptor(M)
        arr = (ctypes.POINTER(Layer) * 1)(ctypes.pointer(d))
        yp = (ctypes.c_void_p * world)(*[A.bufs[q].data_ptr() + 2048 for q in range(world)])
        ld, no = (ctypes.c_int64 * 1)(N), (ctypes.c_int64 * 1)(shard.col0)
        s = A.sync(r, -1, 0, -1)
        s.flags, s.tag_stride, s.y_seq = PEER_Y_TAGGED, 5, 2
        check(lib.b200q_linear_group_sharded(arr, 1, xd.data_ptr(), M, xd.stride(0), yp, ld, no, ctypes.byref(s),
                                             ws.data_ptr(), ws.numel(), st))
        del shard
    torch.cuda.synchronize()
    ref = oracle_forward(L, x)
    for q in (0, world - 1):
        words = A.bufs[q][2048:2048 + M * N * 4].view(torch.int32).cpu().numpy().astype(np.uint32)
        assert np.all(words >> 16 == 7)
        assert rel_err((words & 0xffff).astype(np.uint16).view(np.float16).reshape(M, N), ref) < 1e-3
