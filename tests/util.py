"""Shared helpers for the GPU parity tests (oracle is the checker; product code is qllm_b200)."""
import ctypes

import numpy as np
import torch

from oracle import qlinear_oracle as O

CLS = {"GPTQ": "QuantLinearGPTQ", "HQQ": "QuantLinearHQQ", "GEMM": "WQLinear_GEMM", "MARLIN": "QuantLinearMarlin",
       "GEMV": "WQLinear_GEMV", "ORT": "QuantLinearORT"}


def layer_from_dict(L, device="cuda", dtype=torch.float16):
    import qllm_b200
    cls = getattr(qllm_b200, CLS[L["layout"]])
    gs = L["group_size"]
    layer = cls(L["bits"], gs, L["K"], L["N"], L["bias"] is not None, dtype=dtype)
    layer.qweight = torch.from_numpy(L["qweight"])
    if L["qzeros"] is not None:
        layer.qzeros = torch.from_numpy(np.ascontiguousarray(L["qzeros"]))
    layer.scales = torch.from_numpy(L["scales"])
    if L["layout"] in ("GPTQ", "ORT"):
        layer.g_idx = torch.from_numpy(L["g_idx"].astype(np.int32))
    if L["bias"] is not None:
        layer.bias = torch.from_numpy(L["bias"])
    return layer.to(device)


def oracle_forward(L, x, mode="engine"):
    W = O.dequant(L["q"], L["z"], L["s"], L["g_idx"], mode)
    return O.matmul_ref(x, W, L["bias"], acc=np.float64)


def rel_err(y, ref):
    return float(np.abs(np.asarray(y, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))
