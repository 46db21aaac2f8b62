"""Per-shape kernel timing through the C ABI (CUDA events, L2-cold: rotates over weight copies
whose total size exceeds the 126 MB L2).  Prints achieved algorithmic GB/s and TFLOP/s vs the
measured peaks.  Usage: python tools/microbench.py [--layout GEMM] [--m 1] [--graph]"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qllm_b200  # noqa: E402
from qllm_b200 import q_layers  # noqa: E402

PEAKS = {"hbm_gbs": 6578.3, "bf16_tflops": 1662.8}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass


def rand_layer(layout, bits, gs, K, N, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    ri = lambda *s: torch.randint(-2**31, 2**31 - 1, s, dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    G = K // gs
    sc = (torch.rand(G, N, device=dev, generator=g) * 0.4 + 0.8) * (1.0 / (6.5 * K ** 0.5))
    if layout == "GEMM":
        l = qllm_b200.WQLinear_GEMM(bits, gs, K, N, False, dtype=torch.float16)
        l.qweight, l.qzeros = ri(K, N // 8), ri(G, N // 8)
    elif layout == "MARLIN":
        l = qllm_b200.QuantLinearMarlin(bits, gs, K, N, False, dtype=torch.float16)
        l.qweight = ri(K // 16, 2 * N)
    elif layout == "HQQ":
        l = qllm_b200.QuantLinearHQQ(bits, gs, K, N, False, dtype=torch.float16)
        l.qweight = ri(K * bits // 32, N)
        l.qzeros = torch.randint(0, 2 ** bits, (G, N), device=dev, generator=g).to(torch.float16)
    else:
        l = qllm_b200.QuantLinearGPTQ(bits, gs, K, N, False, dtype=torch.float16)
        l.qweight, l.qzeros = ri(K * bits // 32, N), ri(G, N * bits // 32)
        l.g_idx = l.g_idx.to(dev)
        if layout == "GPTQ_ACT":                          # desc_act checkpoint: rows of every group scattered over K
            l.g_idx = l.g_idx[torch.randperm(K, device=dev, generator=g)].contiguous()
    l.scales = sc.to(torch.float16)
    return l.to(dev)


def alg_bytes(layout, bits, gs, K, N, M):
    G = K // gs
    z = 0 if layout == "MARLIN" else (G * N * 2 if layout == "HQQ" else G * N * bits // 8)
    return K * N * bits // 8 + G * N * 2 + z + M * K * 2 + M * N * 2 + (K * 4 if layout == "GPTQ_ACT" else 0)


def time_shape(layout, bits, gs, K, N, M, iters=200, use_graph=False, force=None):
    dev = torch.device("cuda:0")
    wbytes = K * N * bits // 8
    copies = max(2, int(200e6 // wbytes) + 1)
    layers = [rand_layer(layout, bits, gs, K, N, dev, s) for s in range(copies)]
    x = torch.randn(M, K, dtype=torch.float16, device=dev)
    y = torch.empty(M, N, dtype=torch.float16, device=dev)
    descs = [l._decode_descriptor(M) for l in layers]
    need = qllm_b200.lib.b200q_workspace_bytes(ctypes.byref(descs[0]), M)
    ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=dev)
    fn = {None: qllm_b200.lib.b200q_linear, "gemv": qllm_b200.lib.b200q_gemv, "gemm": qllm_b200.lib.b200q_gemm}[force]
    st = torch.cuda.current_stream().cuda_stream

    def call(i):
        qllm_b200.check(fn(ctypes.byref(descs[i % copies]), x.data_ptr(), M, K, y.data_ptr(), N, ws.data_ptr(), ws.numel(), st))

    for i in range(copies):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if use_graph:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            st = s.cuda_stream
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for i in range(iters):
                    call(i)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for i in range(iters):
            call(i)
        e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    b = alg_bytes(layout, bits, gs, K, N, M)
    gbs = b / us / 1e3
    tf = 2.0 * M * N * K / us / 1e6
    kern = qllm_b200.lib.b200q_select_kernel(ctypes.byref(descs[0]), M)
    return dict(layout=layout, bits=bits, group=gs, K=K, N=N, M=M, us=round(us, 3), GBps=round(gbs, 1),
                hbm_frac=round(gbs / PEAKS["hbm_gbs"], 3), TFLOPs=round(tf, 2),
                tc_frac=round(tf / PEAKS["bf16_tflops"], 3), kernel=kern, graph=use_graph)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", default="GEMM,GPTQ,MARLIN")
    ap.add_argument("--m", default="1")
    ap.add_argument("--shapes", default="4096x4096,4096x11008,11008x4096")
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--group", type=int, default=128)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--force", default=None)
    a = ap.parse_args()
    for layout in a.layouts.split(","):
        for shp in a.shapes.split(","):
            K, N = (int(v) for v in shp.split("x"))
            for M in (int(v) for v in a.m.split(",")):
                print(json.dumps(time_shape(layout, a.bits, a.group, K, N, M, a.iters, a.graph, a.force)), flush=True)
