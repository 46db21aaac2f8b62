"""Repeated calls of the tcgen05 GEMM (library's own tile choice): every output against the first (bit-exact) and against a
float64 matmul over the dequantised weights.  Used with and without compute-sanitizer, and with B200Q_LIB=<other build>."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qllm_b200
from tools.microbench import rand_layer

def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda:0")
    bad = 0
    for K, N, M in ((4096, 4096, 512), (11008, 4096, 1000), (4096, 11008, 200), (1024, 512, 300)):
        layer = rand_layer("GPTQ", 4, 128, K, N, dev, 3)
        x = torch.randn(M, K, dtype=torch.float16, device=dev)
        W = layer.dequantize().double()
        ref = x.double() @ W
        first = layer(x).clone()
        rel = ((first.double() - ref).abs().max() / ref.abs().max()).item()
        ndiff = 0
        for _ in range(reps):
            if not torch.equal(layer(x), first):
                ndiff += 1
        torch.cuda.synchronize()
        print(f"K={K} N={N} M={M}: rel vs float64 {rel:.2e}, {ndiff}/{reps} calls differ from the first", flush=True)
        bad += ndiff + (rel > 1e-3)
    print("STRESS", "FAILED" if bad else "OK")

if __name__ == "__main__":
    main()
