"""Determinism stress of the CTA-pair split-K GEMM: N back-to-back calls per configuration, every output compared with the
first (bit-exact) and with the unsplit kernel (1e-3)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qllm_b200
from tools.microbench import rand_layer

def opt(k, v):
    qllm_b200.check(qllm_b200.lib.b200q_debug_set_option(k.encode(), float(v)))

def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    dev = torch.device("cuda:0")
    bad = 0
    for K, N, M in ((11008, 4096, 300), (4096, 4096, 512), (4096, 11008, 200), (4096, 4096, 129), (11008, 4096, 1000)):
        layer = rand_layer("GPTQ", 4, 128, K, N, dev, 3)
        x = torch.randn(M, K, dtype=torch.float16, device=dev)
        opt("gemm_force_tt", 128); opt("gemm_force_ksplit", 1)
        base = layer(x).clone()
        for tt in (128, 256):
            opt("gemm_force_tt", tt); opt("gemm_force_ksplit", 2)
            first = layer(x).clone()
            rel = ((first.float() - base.float()).abs().max() / base.float().abs().max()).item()
            ndiff = 0
            for _ in range(reps):
                y = layer(x)
                if not torch.equal(y, first):
                    ndiff += 1
            torch.cuda.synchronize()
            print(f"K={K} N={N} M={M} tt={tt} cs=2: rel vs unsplit {rel:.2e}, {ndiff}/{reps} calls differ from the first", flush=True)
            bad += ndiff
        opt("gemm_force_tt", 0); opt("gemm_force_ksplit", 0)
    print("STRESS", "FAILED" if bad else "OK")

if __name__ == "__main__":
    main()
