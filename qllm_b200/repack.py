"""`--pack_mode` conversion on the GPU.

The reference converts a model between pack modes (GEMM / GPTQ / MARLIN; `qllm --load ... --pack_mode ...`) by
dequantising every layer to fp16 and re-quantising it with torch ops on one device
(repack_to_new_mode, qllm/auto_model_quantization.py:115-147: unpack() -> fp16 weight -> target.pack(), with a
round((W + z s) / s) in the middle).  The integers, zeros and scales never change, so the conversion is a pure
re-layout: here every layer goes  source layout --b200q_repack_gptq4--> K-packed words --b200q_repack_from_gptq4-->
target layout, exact by construction (no rounding anywhere) and HBM-bound.

    convert_layer(layer, "MARLIN")            -> a new QuantLinear of the target class (same device)
    repack_to_new_mode(model, "GPTQ")         -> the model with every b200q QuantLinear converted in place
"""
import ctypes

import torch
import torch.nn as nn

from ._lib import LAYOUT_AWQ_GEMM, LAYOUT_AWQ_GEMV, LAYOUT_GPTQ, LAYOUT_HQQ, LAYOUT_MARLIN, Layer, check, lib
from . import codec
from .q_layers import (QuantLinearGPTQ, QuantLinearHQQ, QuantLinearMarlin, QuantLinearORT, WQLinear_GEMM, WQLinear_GEMV, _B200QuantLinearBase,
                       _set_op_by_name, select_quant_linear)

_TARGETS = {"GPTQ": QuantLinearGPTQ, "GEMM": WQLinear_GEMM, "MARLIN": QuantLinearMarlin, "GEMV": WQLinear_GEMV, "ORT": QuantLinearORT}


def _kpacked(layer):
    """(qweight [K/8, N], qzeros [G, N/8], scales fp16 [G, N]) of a 4-bit layer, on its device (exact)."""
    if layer._detect_act_order():
        raise ValueError("act-order (g_idx) layers are converted by the quantiser, not by a re-layout")
    if layer._layout == LAYOUT_GPTQ:
        if layer._detect_act_order():
            raise ValueError("act-order (g_idx) layers have no AWQ / Marlin form (quant_linear_awq.py:96-103, quant_linear_marlin.py:96-98)")
        return layer.qweight, layer.qzeros, layer.scales.to(torch.float16)
    desc = layer._gemm_descriptor()                   # builds (or reuses) the exact K-packed re-layout
    del desc
    return layer._shadow


def convert_layer(layer: _B200QuantLinearBase, new_pack_mode: str) -> _B200QuantLinearBase:
    new_pack_mode = new_pack_mode.upper()
    if new_pack_mode not in _TARGETS:
        raise NotImplementedError(f"pack_mode {new_pack_mode}")
    cls = _TARGETS[new_pack_mode]
    if type(layer) is cls:
        return layer
    if layer.bits != 4 or layer._layout == LAYOUT_HQQ:
        raise NotImplementedError("pack-mode conversion covers the 4-bit integer-zero layouts (GPTQ, GEMM, GEMV, MARLIN)")
    if not layer.qweight.is_cuda and not getattr(layer, "_consolidated", False):
        raise RuntimeError("convert_layer runs on the GPU: move the layer to a CUDA device first")
    K, N, gs = layer.infeatures, layer.outfeatures, layer.groupsize
    qw, qz, sc = _kpacked(layer)
    dev = qw.device
    new = cls(4, gs, K, N, layer.bias is not None, dtype=layer.dtype)
    new.bias = layer.bias
    if new_pack_mode == "GPTQ":
        new.qweight, new.qzeros, new.scales = qw.clone(), qz.clone(), sc.to(layer.dtype).clone()
        new.g_idx = new.g_idx.to(dev)
        return new
    if new_pack_mode == "MARLIN":
        z = codec.gptq_unpack_qzeros(qz, 4, N)
        if not bool((z == 8).all().item()):
            raise ValueError("pack_mode=MARLIN needs symmetric quantisation (every zero point == 8, quant_linear_marlin.py:99)")
    d = Layer()
    d.layout, d.bits, d.group_size, d.K, d.N, d.zero_bias = LAYOUT_GPTQ, 4, gs, K, N, 0
    d.qweight, d.qzeros, d.scales = qw.data_ptr(), qz.data_ptr(), sc.data_ptr()
    def alloc(t):                                     # outputs are OR-ed / memset in whole 32-bit words
        nbytes = t.numel() * t.element_size()
        return torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=dev), nbytes
    bq, nq = alloc(new.qweight)
    bz, nz = alloc(new.qzeros) if isinstance(new.qzeros, torch.Tensor) else (None, 0)
    out_sc = torch.empty(new.scales.shape, dtype=torch.float16, device=dev)
    check(lib.b200q_repack_from_gptq4(ctypes.byref(d), new._layout, bq.data_ptr(), None if bz is None else bz.data_ptr(),
                                      out_sc.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "b200q_repack_from_gptq4")
    new.qweight = bq.view(torch.uint8)[:nq].view(new.qweight.dtype).reshape(new.qweight.shape)
    new.scales = out_sc.to(new.scales.dtype)
    if bz is not None:
        new.qzeros = bz.view(torch.uint8)[:nz].view(new.qzeros.dtype).reshape(new.qzeros.shape)
    for name, buf in list(new.named_buffers()):       # remaining buffers (e.g. Marlin's legacy workspace) follow the device
        if buf.device != dev:
            setattr(new, name, buf.to(dev))
    return new


def repack_to_new_mode(model: nn.Module, new_pack_mode: str) -> nn.Module:
    """Same contract as the reference's repack_to_new_mode (auto_model_quantization.py:115-147): every QuantLinear of the
    model becomes the class select_quant_linear picks for `new_pack_mode`; model.quant_config.version follows when present."""
    names = [n for n, m in model.named_modules() if isinstance(m, _B200QuantLinearBase) and not isinstance(m, QuantLinearHQQ)]
    for n in names:
        mod = model.get_submodule(n)
        _set_op_by_name(model, n, convert_layer(mod, new_pack_mode))
    qc = getattr(model, "quant_config", None)
    if qc is not None and hasattr(qc, "version"):
        qc.version = new_pack_mode.upper()
    return model
