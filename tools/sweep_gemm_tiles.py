"""Token-tile width x cluster K-split sweep of the tcgen05 GEMM (diagnostic options gemm_force_tt / gemm_force_ksplit):
the data behind the tile policy (tc_pick) in gemm_tcgen05.cu.  Every configuration is checked against the (128, 1) output."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qllm_b200  # noqa: E402
from tools.microbench import rand_layer  # noqa: E402


def opt(name, v):
    qllm_b200.check(qllm_b200.lib.b200q_debug_set_option(name.encode(), float(v)))


def run(layers, x, M, K, N, iters):
    dev = x.device
    y = torch.empty(M, N, dtype=torch.float16, device=dev)
    descs = [l._decode_descriptor(M) for l in layers]
    need = qllm_b200.lib.b200q_workspace_bytes(ctypes.byref(descs[0]), M)
    ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    n = len(layers)
    with torch.cuda.stream(s):
        def call(i):
            qllm_b200.check(qllm_b200.lib.b200q_gemm(ctypes.byref(descs[i % n]), x.data_ptr(), M, K, y.data_ptr(), N, ws.data_ptr(),
                                                     ws.numel(), s.cuda_stream))
        call(0)
        torch.cuda.synchronize()
        y0 = y.clone()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(iters):
                call(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters, y0


def main():
    dev = torch.device("cuda:0")
    shapes = [(4096, 4096), (4096, 11008), (11008, 4096)]
    Ms = [int(m) for m in (sys.argv[1].split(",") if len(sys.argv) > 1 else "256,512,1024,2048".split(","))]
    combos = [(128, 1), (128, 2), (256, 1), (256, 2), (0, 0)]          # (0, 0): the library's own choice
    for K, N in shapes:
        copies = max(2, int(200e6 // (K * N // 2)) + 1)
        layers = [rand_layer("GPTQ", 4, 128, K, N, dev, s) for s in range(copies)]
        for M in Ms:
            x = torch.randn(M, K, dtype=torch.float16, device=dev)
            ref = None
            for tt, ks in combos:
                opt("gemm_force_tt", tt)
                opt("gemm_force_ksplit", ks)
                us, y = run(layers, x, M, K, N, 2 * copies)
                if ref is None:
                    ref = y
                err = ((y.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
                print(json.dumps(dict(K=K, N=N, M=M, tt=tt, ksplit=ks, us=round(us, 2), TFLOPs=round(2.0 * M * N * K / us / 1e6, 1),
                                      rel_vs_128x1=round(err, 6))), flush=True)
            opt("gemm_force_tt", 0)
            opt("gemm_force_ksplit", 0)


if __name__ == "__main__":
    main()
