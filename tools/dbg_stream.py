"""Localise a decode-kernel mismatch: error per output column block / per M row for a forced plan."""
import sys, os, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qllm_b200
from oracle import qlinear_oracle as O
from tests.util import layer_from_dict, oracle_forward
lib = qllm_b200.lib
layout, K, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
L = O.make_layer(layout, 4, 128, K, N, seed=K + N)
layer = layer_from_dict(L)
for M in (1, 3):
    x = np.random.default_rng(3).standard_normal((M, K)).astype(np.float16)
    xt = torch.from_numpy(x).cuda()
    ref = oracle_forward(L, x)
    for opt, val in (("st_cluster", 0), ("st_cluster", 1), ("st_cluster", 2), ("st_cluster", 3), ("st_cluster", 4), ("st_cluster", 8)):
        lib.b200q_debug_set_option(opt.encode(), float(val))
        plan = (ctypes.c_int32 * 4)()
        lib.b200q_debug_decode_plan(ctypes.byref(layer._descriptor()), M, plan)
        y = layer(xt).float().cpu().numpy()
        err = np.abs(y - ref) / np.abs(ref).max()
        blocks = err.reshape(M, -1, 32).max(axis=2)
        print(f"M={M} {opt}={val} plan={list(plan)} max_err={err.max():.2e} bad 32-col blocks:", [(int(m), int(b)) for m, b in zip(*np.nonzero(blocks > 1e-3))][:12])
    lib.b200q_debug_set_option(b"st_cluster", 0.0)
