"""Decode kernel: device time per call for every cluster size (K split) per layout / shape, in one process."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qllm_b200
from tools.microbench import time_shape
lib = qllm_b200.lib
lib.b200q_debug_set_option(b"fma_max_m", 0)          # rp kernel only
shapes = ((4096, 4096), (4096, 11008), (11008, 4096))
for lay in sys.argv[1].split(",") if len(sys.argv) > 1 else ("GEMM", "GPTQ", "MARLIN"):
    for K, N in shapes:
        row = {"layout": lay, "K": K, "N": N}
        for c in range(1, 9):
            lib.b200q_debug_set_option(b"force_cluster", c)
            try:
                row[f"c{c}"] = round(time_shape(lay, 4, 128, K, N, 1, 300, True)["us"], 2)
            except Exception as e:
                row[f"c{c}"] = str(e)[:40]
        lib.b200q_debug_set_option(b"force_cluster", 0)
        row["auto"] = round(time_shape(lay, 4, 128, K, N, 1, 300, True)["us"], 2)
        print(json.dumps(row), flush=True)
