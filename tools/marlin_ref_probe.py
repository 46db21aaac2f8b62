"""Does the reference's own Marlin kernel (csrc/awq_cuda/quantization/marlin_cuda_kernel.cu, compiled unmodified for
compute_100 / sm_100 by oracle/Makefile.ref) run on B200?  Each shape runs in its own process (a faulting kernel poisons
the CUDA context); the log is the evidence behind "no reference-CUDA comparison for pack_mode=MARLIN"."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, "%s"); sys.path.insert(0, "%s/oracle/_ref")
import numpy as np
from oracle import qlinear_oracle as O
import awq_inference_engine as awq
M, K, N, gs = %d, %d, %d, %d
L = O.make_layer("MARLIN", 4, gs, K, N, seed=1)
A = torch.randn(M, K, dtype=torch.float16, device="cuda")
B = torch.from_numpy(L["qweight"]).cuda(); s = torch.from_numpy(L["scales"]).cuda()
C = torch.zeros(M, N, dtype=torch.float16, device="cuda")
ws = torch.zeros(N // 128 * 16, dtype=torch.int32, device="cuda")
awq.mul(A, B, C, s, ws, -1, -1, -1, 8)            # quant_linear_marlin.py:142-146
torch.cuda.synchronize()
W = O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "engine")
ref = A.double().cpu().numpy() @ W.astype(np.float64)
print("ok rel_err", float(np.abs(C.double().cpu().numpy() - ref).max() / np.abs(ref).max()))
'''
for (M, K, N, gs) in ((1, 512, 256, 128), (16, 512, 256, 128), (200, 1024, 512, 128), (1, 256, 256, -1), (1, 1024, 512, 128),
                      (16, 1024, 512, 128), (1, 4096, 4096, 128), (16, 4096, 4096, 128), (512, 4096, 4096, 128), (64, 2048, 1024, -1)):
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT, M, K, N, gs)], capture_output=True, text=True, timeout=300)
    err = [l for l in r.stderr.strip().splitlines() if "Error" in l or "error" in l]
    tail = (r.stdout.strip().splitlines() or [""])[-1] + " | " + (err[-1][:200] if err else "")
    print(f"M={M} K={K} N={N} group={gs}: rc={r.returncode} {tail}", flush=True)
