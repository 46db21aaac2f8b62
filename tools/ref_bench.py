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


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


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
            lg = [rand_layer("GPTQ", 4, 128, K, N, dev, s) for s in range(copies)]
            lm = None
            it = [0]

            def nxt(ls):
                it[0] += 1
                return ls[it[0] % copies]
            r = {"K": K, "N": N, "M": M}
            r["ref_awq_gemm_us"] = timeit(lambda: (lambda l: awq.gemm_forward_cuda(x, l.qweight, l.scales, l.qzeros, 8))(nxt(la)))
            r["b200q_awq_us"] = timeit(lambda: nxt(la)(x))
            if M <= 8:
                r["ref_ort_gemv_us"] = timeit(lambda: (lambda l: ort.gemv(x, l.qweight, l.scales, l.qzeros, None, 128, 4, K, 0))(nxt(lg)))
            else:
                r["ref_ort_dequant_matmul_us"] = timeit(lambda: (lambda l: torch.matmul(x, ort.dequant(l.qweight, l.scales, l.qzeros, None, 128, 4, K, 0)))(nxt(lg)))
            r["b200q_gptq_us"] = timeit(lambda: nxt(lg)(x))
            if False and lm:            # the reference Marlin kernel raises cudaErrorIllegalInstruction on B200 (sm_100)
                ws = torch.zeros(N // 128 * 16, dtype=torch.int32, device=dev)
                C = torch.empty(M, N, dtype=torch.float16, device=dev)
                r["ref_marlin_us"] = timeit(lambda: (lambda l: awq.mul(x, l.qweight, C, l.scales, ws, -1, -1, -1, 16))(nxt(lm)))
                r["b200q_marlin_us"] = timeit(lambda: nxt(lm)(x))
            rows.append({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()})
            print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
