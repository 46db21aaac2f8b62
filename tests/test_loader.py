"""Loader shim: a checkpoint directory in the reference's on-disk format -> HF model whose
QuantLinears are the b200q classes.  Host logic runs on CPU; the forward parity check needs the GPU."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

import qllm_b200
from qllm_b200 import loader


def _tiny_llama(seed=0, hidden=256, inter=512, layers=2):
    from transformers import AutoModelForCausalLM, LlamaConfig
    torch.manual_seed(seed)
    cfg = LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=128, max_position_embeddings=64, tie_word_embeddings=False)
    return AutoModelForCausalLM.from_config(cfg, dtype=torch.float16).eval()


def _rtn(w, bits, gs, symmetric=False):
    """Round-to-nearest group quantiser: -> scales [N, G], zeros [N, G] (the quantiser-side layout pack() takes)."""
    N, K = w.shape
    wg = w.float().reshape(N, K // gs, gs)
    maxq = 2 ** bits - 1
    if symmetric:
        s = (wg.abs().amax(-1) / (maxq // 2)).clamp_min(1e-5)
        return s.to(torch.float16), torch.full_like(s, (maxq + 1) // 2)
    lo, hi = wg.amin(-1), wg.amax(-1)
    s = ((hi - lo) / maxq).clamp_min(1e-5).to(torch.float16).float()
    z = torch.round(-lo / s).clamp(0, maxq)
    return s.to(torch.float16), z


def _quantise(model, pack_mode, method, bits, gs):
    cls = qllm_b200.select_quant_linear(pack_mode, bits, method)
    for name, m in list(model.named_modules()):
        if isinstance(m, nn.Linear) and ".layers." in name:
            s, z = _rtn(m.weight.data, bits, gs, symmetric=(pack_mode == "MARLIN"))
            q = cls(bits, gs, m.in_features, m.out_features, False, dtype=torch.float16)
            q.pack(m, s, z)
            parent = model.get_submodule(name.rsplit(".", 1)[0])
            setattr(parent, name.rsplit(".", 1)[1], q)
    return model


CASES = [("GPTQ", "gptq", 4, 64), ("GEMM", "awq", 4, 128), ("GPTQ", "hqq", 4, 64), ("GPTQ", "gptq", 8, 128)]


@pytest.mark.parametrize("pack_mode,method,bits,gs", CASES)
def test_save_load_round_trip_cpu(tmp_path, pack_mode, method, bits, gs):
    model = _quantise(_tiny_llama(), pack_mode, method, bits, gs)
    qc = loader.QuantConfig(bits, gs, method, pack_mode, False, {"groupsize": gs, "wbits": bits})
    loader.save_quantized(model, str(tmp_path), qc)
    got = loader.from_quantized(str(tmp_path), device="cpu", fuse=True)
    want_cls = qllm_b200.select_quant_linear(pack_mode, bits, method)
    qlayers = [m for m in got.modules() if isinstance(m, want_cls)]
    assert len(qlayers) == 2 * 7
    assert not any(isinstance(m, nn.Linear) for n, m in got.named_modules() if ".layers." in n)
    assert isinstance(got.lm_head, nn.Linear)                                    # no .qweight in the checkpoint: stays fp16
    a, b = model.state_dict(), got.state_dict()
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k].cpu(), b[k].cpu()), k
    assert all(not t.is_meta for t in list(got.parameters()) + list(got.buffers()))
    assert sum(1 for m in got.modules() if getattr(m, "_sibling_group", None) is not None) == 2 * 5   # q,k,v + gate,up
    assert got.quant_config.version == pack_mode and got.quant_config.bits == bits


def test_autogptq_checkpoint_zeros_are_rewritten(tmp_path):
    """No "version" key => AutoGPTQ: qzeros hold z - 1 on disk and are rewritten to z at load (quant_linear_gptq.py:119-134)."""
    model = _quantise(_tiny_llama(1), "GPTQ", "gptq", 4, 64)
    qc = loader.QuantConfig(4, 64, "gptq", "GPTQ", True, {"groupsize": 64, "wbits": 4})
    loader.save_quantized(model, str(tmp_path), qc, autogptq_zeros=True)
    assert os.path.exists(tmp_path / "quantize_config.json") and not os.path.exists(tmp_path / "quant_config.json")
    from safetensors.torch import load_file
    disk = load_file(str(tmp_path / "model.safetensors"))
    name = "model.layers.0.self_attn.q_proj.qzeros"
    assert not torch.equal(disk[name], model.state_dict()[name])                 # z - 1 on disk
    got = loader.from_quantized(str(tmp_path), device="cpu")
    assert got.quant_config.autogptq
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), got.state_dict()[k].cpu()), k


def test_mixed_bit_table_and_config_search_order(tmp_path):
    """quant_config_by_layer.json: only listed layers are swapped, each with its own (wbits, groupsize)."""
    model = _tiny_llama(2)
    table = {}
    for name, m in list(model.named_modules()):
        if isinstance(m, nn.Linear) and ".mlp." in name:
            bits, gs = (8, 128) if "down_proj" in name else (4, 64)
            s, z = _rtn(m.weight.data, bits, gs)
            q = qllm_b200.QuantLinearGPTQ(bits, gs, m.in_features, m.out_features, False, dtype=torch.float16)
            q.pack(m, s, z)
            setattr(model.get_submodule(name.rsplit(".", 1)[0]), name.rsplit(".", 1)[1], q)
            table[name] = {"wbits": bits, "groupsize": gs}
    qc = loader.QuantConfig(4, 64, "gptq", "GPTQ", False, table)
    loader.save_quantized(model, str(tmp_path), qc)
    got = loader.from_quantized(str(tmp_path), device="cpu")
    assert isinstance(got.model.layers[0].self_attn.q_proj, nn.Linear)
    d = got.model.layers[1].mlp.down_proj
    assert isinstance(d, qllm_b200.QuantLinearGPTQ) and d.bits == 8 and d.groupsize == 128
    assert got.model.layers[1].mlp.up_proj.bits == 4
    # search order: quant_config.json wins over config.json["quantization_config"]
    cfg = json.load(open(tmp_path / "config.json"))
    cfg["quantization_config"] = {"bits": 2, "group_size": 32, "version": "GEMM"}
    json.dump(cfg, open(tmp_path / "config.json", "w"))
    assert loader.load_quant_config(str(tmp_path)).bits == 4
    os.remove(tmp_path / "quant_config.json")
    assert loader.load_quant_config(str(tmp_path)).bits == 2


def test_missing_tensors_are_reported(tmp_path):
    model = _quantise(_tiny_llama(3), "GPTQ", "gptq", 4, 64)
    loader.save_quantized(model, str(tmp_path), loader.QuantConfig(4, 64, "gptq", "GPTQ", False, {"groupsize": 64, "wbits": 4}))
    from safetensors.torch import load_file, save_file
    sd = load_file(str(tmp_path / "model.safetensors"))
    sd.pop("model.norm.weight")
    save_file(sd, str(tmp_path / "model.safetensors"))
    with pytest.raises(RuntimeError, match="lacks tensors"):
        loader.from_quantized(str(tmp_path), device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("pack_mode,method,bits,gs", [("GPTQ", "gptq", 4, 64), ("GEMM", "awq", 4, 128), ("MARLIN", "gptq", 4, 128)])
def test_loaded_model_forward_matches_dequantised_fp16_model(tmp_path, pack_mode, method, bits, gs):
    """End to end on the GPU: checkpoint -> from_quantized -> logits (prefill and one decode step through the engine)
    against the same architecture holding the dequantised fp16 weights in nn.Linear."""
    base = _tiny_llama(4)
    ref = _tiny_llama(4)
    model = _quantise(base, pack_mode, method, bits, gs)
    for name, m in model.named_modules():
        if isinstance(m, qllm_b200.q_layers._B200QuantLinearBase):
            w, _, _ = m.unpack()
            ref.get_submodule(name).weight.data.copy_(w)
    loader.save_quantized(model, str(tmp_path), loader.QuantConfig(bits, gs, method, pack_mode, False, {"groupsize": gs, "wbits": bits}))
    got = loader.from_quantized(str(tmp_path), device="cuda")
    ref = ref.cuda()
    n0 = qllm_b200.lib.b200q_launch_count()
    for T in (17, 1):                                   # prefill-sized (tcgen05 GEMM) and decode-sized (decode kernel) calls
        ids = torch.randint(0, 128, (1, T), generator=torch.Generator().manual_seed(T)).cuda()
        with torch.no_grad():
            a, b = got(ids).logits.float(), ref(ids).logits.float()
        assert torch.isfinite(a).all()
        assert ((a - b).abs().max() / b.abs().max()).item() < 2e-2
    assert qllm_b200.lib.b200q_launch_count() > n0      # the engine ran, not a fallback
    out = got.generate(torch.tensor([[1, 2, 3]]).cuda(), max_new_tokens=4, do_sample=False)
    assert out.shape == (1, 7)


# ---- a checkpoint written in the reference's own schema by the reference's own code (tests/golden/make_ckpt_fixture.py) ----
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_ref_gptq")


@pytest.mark.parametrize("sub", ["st", "bin"])
def test_reference_written_checkpoint_loads(sub):
    """Sharded safetensors + model.safetensors.index.json, and sharded .bin + pytorch_model.bin.index.json: quantize_config.json
    (GPTQConfig.to_dict()), quant_config_by_layer.json and config.json["quantization_config"] as the reference writes them
    (modeling/base.py:324-336); every packed buffer arrives bit-identical to what the reference's pack() produced."""
    from safetensors.torch import load_file
    got = loader.from_quantized(os.path.join(FIX, sub), device="cpu", dtype=torch.float16, fuse=False)
    assert got.quant_config.version == "GPTQ" and got.quant_config.bits == 4 and got.quant_config.group_size == 32
    assert not got.quant_config.autogptq                        # "version" present: zeros are stored as z, not z - 1
    ref = {}
    for f in sorted(os.listdir(os.path.join(FIX, "st"))):
        if f.endswith(".safetensors"):
            ref.update(load_file(os.path.join(FIX, "st", f)))
    qlayers = {n: m for n, m in got.named_modules() if isinstance(m, qllm_b200.QuantLinearGPTQ)}
    assert len(qlayers) == 14
    for n, m in qlayers.items():
        assert torch.equal(m.qweight, ref[n + ".qweight"]) and torch.equal(m.qzeros, ref[n + ".qzeros"])
        assert torch.equal(m.scales.float(), ref[n + ".scales"].float())          # fp32 in the checkpoint, exact in fp16
        assert torch.equal(m.g_idx, ref[n + ".g_idx"])
    assert got.lm_head.weight.dtype == torch.float16 and not any(p.is_meta for p in got.parameters())


@pytest.mark.gpu
@pytest.mark.parametrize("sub", ["st", "bin"])
def test_reference_written_checkpoint_logits(sub):
    """End to end against the REFERENCE's own forward: logits.npz holds what the reference model (its QuantLinearGPTQ.forward
    on CPU, fp32) computed for these token ids; the engine (fp16 model, CUDA kernels) must agree to fp16 accuracy."""
    d = np.load(os.path.join(FIX, "logits.npz"))
    got = loader.from_quantized(os.path.join(FIX, sub), device="cuda", dtype=torch.float16)
    n0 = qllm_b200.lib.b200q_launch_count()
    with torch.no_grad():
        a = got(torch.from_numpy(d["ids"]).cuda()).logits.float().cpu().numpy()
        a1 = got(torch.from_numpy(d["ids"][:, :1]).cuda()).logits.float().cpu().numpy()       # decode-sized call
    assert qllm_b200.lib.b200q_launch_count() > n0
    ref = d["logits"]
    assert np.abs(a - ref).max() / np.abs(ref).max() < 5e-3
    assert np.abs(a1[0, 0] - ref[0, 0]).max() / np.abs(ref).max() < 5e-3
