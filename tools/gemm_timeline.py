"""Per-k-block phase stamps (SM cycles, %clock64) of CTA (0,0) of the tcgen05 GEMM (b200q_debug_set_timeline)."""
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
nkb = K // 64
tall = buf.cpu().numpy().reshape(-1, 8).astype(np.float64)
t, ends = tall[:nkb], tall[nkb]
t0 = t[t > 0].min()
print("whole-kernel stamps of CTA (0,0), cycles relative to the first k-block stamp: entry %.0f | barriers+TMEM %.0f | tables %.0f | "
      "accumulators complete %.0f | epilogue stored %.0f" % tuple(ends[j] - t0 for j in range(5)))
names = ["dq:w_landed", "dq:alu_done", "dq:a_free", "dq:signalled", "mma:x_landed", "mma:a_landed", "mma:issued", "mma:committed"]
print("kb  " + "  ".join(f"{n:>13s}" for n in names) + "   (SM cycles since first stamp)")
for kb in list(range(0, 16)) + list(range(nkb // 2, nkb // 2 + 8)) + list(range(nkb - 6, nkb)):
    print(f"{kb:3d} " + "  ".join(f"{(t[kb, j] - t0):13.0f}" for j in range(8)))
mid = slice(8, nkb - 4)
print("median cycles per k-block (mma issue to issue):", np.median(np.diff(t[:, 6])[mid]))
print("median cycles: wait X", np.median((t[:, 4] - np.roll(t[:, 7], 1))[mid]), "| wait A", np.median((t[:, 5] - t[:, 4])[mid]),
      "| issue 4 MMAs", np.median((t[:, 6] - t[:, 5])[mid]), "| 2 commits", np.median((t[:, 7] - t[:, 6])[mid]))
print("median cycles dequant team: LDS+ALU", np.median((t[:, 1] - t[:, 0])[mid]), "| wait A stage free", np.median((t[:, 2] - t[:, 1])[mid]),
      "| st+wait::st+arrive", np.median((t[:, 3] - t[:, 2])[mid]), "| A signalled -> MMA saw it", np.median((t[:, 5] - t[:, 3])[mid]))
print("total cycles first stamp -> last commit:", t[:, 7].max() - t0)
