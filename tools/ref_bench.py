"""Time the reference's own CUDA kernels (oracle/_ref, sm_100 SASS of csrc/awq_cuda + csrc/ort_cuda) next to
the engine on the same packed weights: the 'kernels to beat' table of BASELINE.md §5."""
import ctypes
import glob
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import qllm_b200  # noqa: E402
from tools.microbench import rand_layer  # noqa: E402


def timeit(fn, copies, reps=4):
    """fn(i) enqueues one call on layer copy i.  All copies x reps calls are captured in one CUDA graph, so the number is
    device time per call with no Python / launch-API overhead on either side and cold weights (copies exceed L2)."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(copies):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                for i in range(copies):
                    fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * copies)


def time_eager(fn, copies, reps=8):
    """Eager calls on the (legacy) default stream, CUDA events around the loop: for kernels that ignore the current stream and
    so cannot be captured.  Includes whatever launch overhead the caller's Python adds (both sides get the same loop)."""
    for i in range(copies):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    for _ in range(reps):
        for i in range(copies):
            fn(i)
    e1.record(torch.cuda.default_stream())
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * copies)


def main():
    import awq_inference_engine as awq
    import ort_ops as ort
    dev = torch.device("cuda:0")
    rows = []
    for K, N in ((4096, 4096), (4096, 11008), (11008, 4096)):
        copies = max(2, int(200e6 // (K * N // 2)) + 1)
        for M in (1, 512):
            x = torch.randn(M, K, dtype=torch.float16, device=dev)
            la = [rand_layer("GEMM", 4, 128, K, N, dev, s) for s in range(copies)]
            la_native = [(l.qweight, l.scales, l.qzeros) for l in la]          # the engine releases these at its first forward
            lg = [rand_layer("GPTQ", 4, 128, K, N, dev, s) for s in range(copies)]
            r = {"K": K, "N": N, "M": M, "timing": "cuda graph, device time per call"}
            r["ref_awq_gemm_us"] = timeit(lambda i: awq.gemm_forward_cuda(x, la_native[i][0], la_native[i][1], la_native[i][2], 8), copies)
            r["b200q_awq_us"] = timeit(lambda i: la[i](x), copies)
            if M <= 8:
                # ort_ops.gemv launches on the legacy default stream (dq_gemv.cu:166): a CUDA graph does not capture it, so it is
                # timed eagerly, with the engine timed the same way beside it
                r["ref_ort_gemv_eager_us"] = time_eager(lambda i: ort.gemv(x, lg[i].qweight, lg[i].scales, lg[i].qzeros, None, 128, 4, K, 0), copies)
                r["b200q_gptq_eager_us"] = time_eager(lambda i: lg[i](x), copies)
            else:
                r["ref_ort_dequant_matmul_us"] = timeit(
                    lambda i: torch.matmul(x, ort.dequant(lg[i].qweight, lg[i].scales, lg[i].qzeros, None, 128, 4, K, 0)), copies)
            r["b200q_gptq_us"] = timeit(lambda i: lg[i](x), copies)
            # the reference Marlin kernel raises cudaErrorIllegalInstruction on B200 (sm_100) and poisons the context: not timed
            rows.append({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()})
            print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
