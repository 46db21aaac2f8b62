"""Phase timeline of the decode kernel across a chain of launches replayed from ONE CUDA graph (PDL edges, no host
launch gaps), from the per-CTA %globaltimer stamps of b200q_debug_set_timeline.
python tools/timeline.py [--layout GEMM] [--shape 4096x4096] [--launches 12]"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qllm_b200  # noqa: E402
from tools.microbench import rand_layer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layout", default="GEMM")
ap.add_argument("--shape", default="4096x4096")
ap.add_argument("--launches", type=int, default=12)
a = ap.parse_args()
K, N = (int(v) for v in a.shape.split("x"))
dev = torch.device("cuda:0")
copies = 24
layers = [rand_layer(a.layout, 4, 128, K, N, dev, s) for s in range(copies)]
x = torch.randn(1, K, dtype=torch.float16, device=dev)
y = torch.empty(1, N, dtype=torch.float16, device=dev)
descs = [l._decode_descriptor(1) for l in layers]
ws = torch.zeros(1 << 22, dtype=torch.uint8, device=dev)
lib = qllm_b200.lib
buf = torch.zeros(1 << 20, dtype=torch.int64, device=dev)


def chain(st, n):
    for i in range(n):
        qllm_b200.check(lib.b200q_linear(ctypes.byref(descs[i % copies]), x.data_ptr(), 1, K, y.data_ptr(), N, ws.data_ptr(), ws.numel(), st))


chain(torch.cuda.current_stream().cuda_stream, copies)
torch.cuda.synchronize()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    g = torch.cuda.CUDAGraph()
    lib.b200q_debug_set_timeline(buf.data_ptr(), buf.numel() * 8)     # stamp offsets are fixed at capture time
    with torch.cuda.graph(g, stream=s):
        chain(torch.cuda.current_stream().cuda_stream, a.launches)
    lib.b200q_debug_set_timeline(None, 0)
    g.replay()
    torch.cuda.synchronize()
    buf.zero_()
    g.replay()
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 8)
used = np.nonzero(t[:, 0])[0]
nct = len(used) // a.launches
print(f"{a.layout} {K}x{N}: {nct} CTAs per launch (CUDA graph replay)")
t0 = t[used][:, 0].min()
names = ["start", "ring+table", "upstream_done", "-", "math_done", "cluster_reduced", "stored"]
for li in range(a.launches):
    blk = t[used[li * nct:(li + 1) * nct]].astype(np.float64)
    row = []
    for j, nm in enumerate(names):
        col = blk[:, j]
        col = col[col > 0]
        if len(col):
            row.append(f"{nm}[{(col.min()-t0)/1e3:6.2f},{(np.median(col)-t0)/1e3:6.2f},{(col.max()-t0)/1e3:6.2f}]")
    print(f"launch {li:2d}: " + " ".join(row))
