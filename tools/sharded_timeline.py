"""Phase timeline of the fused column-sharded decode (torchrun, N >= 2): per launch on rank 0, when its CTAs started,
when the local upstream kernel was done, when the peers' posts had arrived, when math / reduction / stores+post ended.
torchrun --nproc-per-node 2 tools/sharded_timeline.py [--blocks 3]"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import qllm_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=3)
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
bench.BLOCKS = a.blocks
dev = torch.device("cuda", lr)
blocks = bench.build_model(dev, rank, world, "GEMM")
step = bench.FusedShardedStep(blocks, dev, 1, rank, world)
lib = qllm_b200.lib
s = torch.cuda.Stream()
buf = torch.zeros(1 << 20, dtype=torch.int64, device=dev)
with torch.cuda.stream(s):
    step.run(s.cuda_stream)
    s.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    lib.b200q_debug_set_timeline(buf.data_ptr(), buf.numel() * 8)
    with torch.cuda.graph(g, stream=s):
        step.run(torch.cuda.current_stream().cuda_stream)
    lib.b200q_debug_set_timeline(None, 0)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    buf.zero_()
    g.replay()
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    t = buf.cpu().numpy().reshape(-1, 8)
    used = np.nonzero(t[:, 0])[0]
    rows = t[used].astype(np.float64)
    t0 = rows[:, 0].min()
    tag = t[used][:, 7].astype(np.uint64)
    bounds = [i for i in range(len(rows)) if (int(tag[i]) & 0xffffffff) == 0] + [len(rows)]
    names = ["qkv", "o", "gate|up", "down"] * a.blocks
    cols = ["start", "prologue", "peers_done", "local_done", "math_done", "cluster_reduced", "stored+posted"]
    for li in range(len(bounds) - 1):
        blk = rows[bounds[li]:bounds[li + 1]]
        out = []
        for j in (0, 1, 3, 2, 4, 5, 6):
            col = blk[:, j]
            col = col[col > 0]
            if len(col):
                out.append(f"{cols[j]}[{(col.min()-t0)/1e3:6.2f},{(np.median(col)-t0)/1e3:6.2f},{(col.max()-t0)/1e3:6.2f}]")
        print(f"{li:2d} {names[li] if li < len(names) else '?':8s} ctas={len(blk):3d} " + " ".join(out))
dist.destroy_process_group()
