"""Loader shim: `qllm --load <checkpoint>` for the b200q engine.

The reference's `AutoQuantizedModelForCausalLM.from_quantized` (qllm/modeling/base.py:225-322) cannot run against
transformers 5.x (`transformers.modeling_utils.no_init_weights` is gone, base.py:247; `accelerate` is absent, base.py:11;
`QuantLinearMarlin.__init__` rejects the `dtype=` kwarg the loader passes).  This module restates what that
function does for a checkpoint on disk -- nothing else of the reference's control plane:

  1. quantisation config, in the reference's search order (config.py:81-126): quant_config.json, quantize_config.json,
     config.json["quantization_config"]; a config without "version" is an AutoGPTQ checkpoint (zeros stored as z - 1);
     per-layer (wbits, groupsize) from quant_config_by_layer.json (config.py:70-79);
  2. the model skeleton from its HF config on the meta device (no 13 GB of throw-away fp16 weights, no init);
  3. every nn.Linear that has a `.qweight` tensor in the checkpoint becomes the QuantLinear class picked by
     select_quant_linear (utils/modelutils.py:44-68) through make_mixbits_quant_linear (:161-182);
  4. safetensors / .bin shards are assigned buffer by buffer (`load_state_dict(strict=False, assign=True)`, the
     reference's _load_check_point, base.py:118-172), AutoGPTQ zeros are rewritten once
     (handle_qzeros_for_autogptq, quant_linear_gptq.py:119-134);
  5. the model moves to the GPU and sibling projections are fused (qllm_b200.fuse_siblings).

`save_quantized` writes the same on-disk format back (the reference's save path: base.py:324-365) and exists so that
the round trip can be tested without a network.
"""
import glob
import json
import os
from dataclasses import dataclass, field

import torch
import torch.nn as nn

from .q_layers import (QuantLinearGPTQ, _B200QuantLinearBase, fuse_siblings, make_mixbits_quant_linear,
                       select_quant_linear)


@dataclass
class QuantConfig:
    """What BaseQuantizeConfig exposes to the loader (config.py:11-126)."""
    bits: int
    group_size: int
    quant_method: str
    version: str                      # pack mode: GPTQ | GEMM | MARLIN | AUTO ...
    autogptq: bool                    # zeros stored as z - 1 (COMPATIBLE_WITH_AUTOGPTQ)
    by_op: dict = field(default_factory=dict)
    raw: dict = field(default_factory=dict)


def load_quant_config(path: str) -> QuantConfig:
    raw = None
    for name in ("quant_config.json", "quantize_config.json"):
        f = os.path.join(path, name)
        if os.path.exists(f):
            raw = json.load(open(f))
            break
    if raw is None:
        f = os.path.join(path, "config.json")
        if os.path.exists(f):
            raw = json.load(open(f)).get("quantization_config")
            if raw is not None and raw.get("use_exllama", False):
                raise ValueError("use_exllama checkpoints are not supported (as in the reference, config.py:98)")
    if raw is None:
        raise FileNotFoundError(f"no quant_config.json / quantize_config.json / quantization_config under {path}")
    bits = raw.get("w_bit", raw.get("bits"))
    group = raw.get("q_group_size", raw.get("group_size"))
    method = raw.get("quant_method")
    if method == "vptq":
        raise NotImplementedError("VPTQ checkpoints are outside the b200q hot path")
    if bits is None or group is None:
        raise ValueError("quantisation config lacks bits / group_size")
    autogptq = bool(raw.get("COMPATIBLE_WITH_AUTOGPTQ"))
    if "version" not in raw:                      # GPTQ-for-LLaMa / AutoGPTQ checkpoint (config.py:111-116)
        method, version, autogptq = "gptq", "GPTQ", True
    else:
        version = str(raw["version"]).upper()
        method = raw.get("quant_method", "awq")
    by_op = {"groupsize": int(group), "wbits": int(bits)}
    f = os.path.join(path, "quant_config_by_layer.json")
    if os.path.exists(f):
        by_op = json.load(open(f))
    return QuantConfig(int(bits), int(group), str(method).lower(), version, autogptq, by_op, raw)


def checkpoint_files(path: str):
    """Weight shards in the order the reference tries them (base.py:52-116): safetensors first, then .bin."""
    for index, pattern in (("model.safetensors.index.json", "*.safetensors"), ("pytorch_model.bin.index.json", "*.bin")):
        idx = os.path.join(path, index)
        if os.path.exists(idx):
            names = sorted(set(json.load(open(idx))["weight_map"].values()))
            return [os.path.join(path, n) for n in names]
        files = sorted(glob.glob(os.path.join(path, pattern)))
        if files:
            return files
    raise FileNotFoundError(f"no *.safetensors / *.bin weights under {path}")


def _shard_keys(f: str):
    if f.endswith(".safetensors"):
        from safetensors import safe_open
        with safe_open(f, framework="pt") as h:
            return list(h.keys())
    return list(torch.load(f, map_location="meta", weights_only=True).keys())


def _load_shard(f: str):
    if f.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(f, device="cpu")
    return torch.load(f, map_location="cpu", weights_only=True)


def _rebuild_meta_buffers(model):
    """Non-persistent buffers (rotary inv_freq) are computed in __init__ and therefore still on the meta device:
    re-create the owning module from the model config on the CPU and take its buffers."""
    for name, mod in list(model.named_modules()):
        metas = [b for b, t in mod.named_buffers(recurse=False) if t.is_meta]
        if not metas:
            continue
        try:
            fresh = type(mod)(config=model.config)
        except Exception as e:                                            # pragma: no cover - unknown architecture
            raise RuntimeError(f"cannot materialise buffers {metas} of {name} ({type(mod).__name__}): {e}")
        for b in metas:
            mod._buffers[b] = getattr(fresh, b).detach().clone()


def from_quantized(path: str, device="cuda", dtype=torch.float16, fuse: bool = True, config=None):
    """Load a GPTQ / AWQ / HQQ / Marlin checkpoint directory into an HF causal LM whose QuantLinears run on libb200q."""
    from transformers import AutoConfig, AutoModelForCausalLM
    qc = load_quant_config(path)
    cfg = config if config is not None else AutoConfig.from_pretrained(path)
    if hasattr(cfg, "quantization_config"):
        try:
            delattr(cfg, "quantization_config")        # the HF quantiser registry must not try to handle the checkpoint
        except Exception:
            pass
    with torch.device("meta"):
        model = AutoModelForCausalLM.from_config(cfg, dtype=dtype)
    files = checkpoint_files(path)
    keys = [k for f in files for k in _shard_keys(f)]
    quantised = {k[:-len(".qweight")] for k in keys if k.endswith(".qweight")}
    linears = {n for n, m in model.named_modules() if isinstance(m, nn.Linear)}
    names = sorted(linears & quantised)
    if "groupsize" not in qc.by_op:                    # per-layer table: only the layers it lists (base.py:265-269)
        names = [n for n in names if n in qc.by_op]
    target = select_quant_linear(qc.version, qc.bits, qc.quant_method)
    make_mixbits_quant_linear(model, names, qc.by_op, target_layer=target)
    missing, unexpected = set(), []
    for f in files:
        res = model.load_state_dict(_load_shard(f), strict=False, assign=True)
        unexpected += [k for k in res.unexpected_keys if not k.endswith(".bias")]
    model.tie_weights()
    _rebuild_meta_buffers(model)
    missing = [n for n, p in list(model.named_parameters()) + list(model.named_buffers()) if p.is_meta]
    if missing:
        raise RuntimeError(f"checkpoint {path} lacks tensors for: {missing[:8]}{' ...' if len(missing) > 8 else ''}")
    for m in model.modules():
        if isinstance(m, _B200QuantLinearBase):
            m._desc = None
            if qc.autogptq and isinstance(m, QuantLinearGPTQ):
                m.handle_qzeros_for_autogptq()
    model.quant_config, model.quant_config_by_layer, model.unexpected_keys = qc, qc.by_op, unexpected
    model.eval()
    if dtype is not None:                              # assign=True keeps the checkpoint's dtypes (an fp32 checkpoint would stay fp32)
        model = model.to(dtype)
    if device is not None and str(device) != "cpu":
        model = model.to(device)
    if fuse:
        fuse_siblings(model)
    return model


def save_quantized(model, path: str, qc: QuantConfig, autogptq_zeros: bool = False):
    """Write `model` (QuantLinears already packed) in the reference's on-disk format: config.json, model.safetensors,
    quant_config.json (+ quant_config_by_layer.json for mixed-bit models).  `autogptq_zeros=True` writes an AutoGPTQ
    style checkpoint instead: quantize_config.json without "version" and qzeros holding z - 1."""
    from safetensors.torch import save_file
    from . import codec
    os.makedirs(path, exist_ok=True)
    model.config.save_pretrained(path)
    sd = {}
    for k, v in model.state_dict().items():
        if v.is_meta:
            continue
        sd[k] = v.detach().cpu().contiguous().clone()
    if autogptq_zeros:
        for name, m in model.named_modules():
            if isinstance(m, QuantLinearGPTQ):
                z = codec.gptq_unpack_qzeros(m.qzeros.cpu(), m.bits, m.outfeatures)
                sd[name + ".qzeros"] = codec.gptq_pack_qzeros((z - 1) & m.maxq, m.bits).contiguous()
        json.dump({"bits": qc.bits, "group_size": qc.group_size, "desc_act": False}, open(os.path.join(path, "quantize_config.json"), "w"))
    else:
        json.dump({"w_bit": qc.bits, "q_group_size": qc.group_size, "version": qc.version, "quant_method": qc.quant_method,
                   "zero_point": True}, open(os.path.join(path, "quant_config.json"), "w"))
        if "groupsize" not in qc.by_op:
            json.dump(qc.by_op, open(os.path.join(path, "quant_config_by_layer.json"), "w"))
    tied = getattr(model.config, "tie_word_embeddings", False)
    if tied:
        sd.pop("lm_head.weight", None)
    save_file(sd, os.path.join(path, "model.safetensors"), metadata={"format": "pt"})
