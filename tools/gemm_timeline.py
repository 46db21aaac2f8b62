"""Per-k-block phase stamps of CTA (0,0) of the tcgen05 GEMM (b200q_debug_set_timeline)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qllm_b200
from tools.microbench import rand_layer
M, K, N = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 4096, 4096
dev = torch.device("cuda:0")
l = rand_layer("GPTQ", 4, 128, K, N, dev, 0)
x = torch.randn(M, K, dtype=torch.float16, device=dev)
for _ in range(3): l(x)
buf = torch.zeros(1 << 16, dtype=torch.int64, device=dev)
qllm_b200.lib.b200q_debug_set_timeline(buf.data_ptr(), buf.numel() * 8)
l(x); torch.cuda.synchronize()
qllm_b200.lib.b200q_debug_set_timeline(None, 0)
t = buf.cpu().numpy().reshape(-1, 8)[: K // 64].astype(np.float64)
t0 = t[t > 0].min()
names = ["dq:in_landed", "dq:alu_done", "dq:a_free", "dq:st_issued", "mma:ready", "mma:issued"]
print("kb  " + "  ".join(f"{n:>13s}" for n in names) + "   (us since first stamp)")
for kb in list(range(0, 12)) + list(range(K // 64 - 4, K // 64)):
    print(f"{kb:3d} " + "  ".join(f"{(t[kb, j] - t0) / 1e3:13.2f}" for j in range(6)))
d = np.diff(t[:, 5])
print("median us per k-block (mma issue to issue):", np.median(d) / 1e3)
