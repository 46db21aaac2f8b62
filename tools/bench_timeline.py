"""Phase timeline of the bench's own decode graph (sibling groups, LlamaDecoderLayer dependency order) for a few
decoder blocks: per launch, when its CTAs started, when the upstream result became visible, when the math and the
reduction ended.  python tools/bench_timeline.py [--blocks 3]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import qllm_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=3)
ap.add_argument("--layout", default="GEMM")
a = ap.parse_args()
bench.BLOCKS = a.blocks
dev = torch.device("cuda:0")
blocks = bench.build_model(dev, 0, 1, a.layout)
step = bench.DecodeStep(blocks, dev, 1, 0, 1)
lib = qllm_b200.lib
s = torch.cuda.Stream()
buf = torch.zeros(1 << 20, dtype=torch.int64, device=dev)
with torch.cuda.stream(s):
    step.run(s.cuda_stream)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    lib.b200q_debug_set_timeline(buf.data_ptr(), buf.numel() * 8)
    n0 = lib.b200q_launch_count()
    with torch.cuda.graph(g, stream=s):
        step.run(torch.cuda.current_stream().cuda_stream)
    nl = lib.b200q_launch_count() - n0
    lib.b200q_debug_set_timeline(None, 0)
    g.replay()
    torch.cuda.synchronize()
    buf.zero_()
    g.replay()
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 8)
used = np.nonzero(t[:, 0])[0]
# launches are appended back to back; split where the start stamp jumps backwards is not reliable -> use plans
names = ["qkv", "o", "gate|up", "down"] * a.blocks
import ctypes
sizes = []
for b in blocks:
    for grp in ([b["q"], b["k"], b["v"]], [b["o"]], [b["gate"], b["up"]], [b["down"]]):
        # CTAs of the launch = rows with a start stamp; recover from the plan of each layer group is not exposed -> infer below
        sizes.append(None)
t0 = t[used][:, 0].min()
rows = t[used].astype(np.float64)
# launch boundaries: slot 7 = (grid size << 32 | CTA index); a launch starts at CTA index 0
tag = t[used][:, 7].astype(np.uint64)
bounds = [i for i in range(len(rows)) if (int(tag[i]) & 0xffffffff) == 0]
bounds.append(len(rows))
print(f"{nl} launches captured, {len(bounds) - 1} inferred")
cols = ["start", "prologue", "upstream_done", "-", "math_done", "cluster_reduced", "stored"]
prev_done = None
for li in range(len(bounds) - 1):
    blk = rows[bounds[li]:bounds[li + 1]]
    out = []
    for j in (0, 1, 2, 4, 5, 6):
        col = blk[:, j]
        col = col[col > 0]
        if len(col):
            out.append(f"{cols[j]}[{(col.min()-t0)/1e3:6.2f},{(np.median(col)-t0)/1e3:6.2f},{(col.max()-t0)/1e3:6.2f}]")
    nm = names[li] if li < len(names) else "?"
    print(f"{li:2d} {nm:8s} ctas={len(blk):3d} " + " ".join(out))
