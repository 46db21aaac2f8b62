"""Write a tiny Llama checkpoint in the REFERENCE's on-disk schema, using the reference's own code for everything the
schema defines (authoring container only; needs /root/reference):

  * every decoder Linear is quantised and packed by the reference's QuantLinearGPTQ.pack (quant_linear_gptq.py / compress_weight.py);
  * quantize_config.json is GPTQConfig(...).to_dict() (quantization/config_builder.py:34-46), quant_config_by_layer.json the
    per-layer table, config.json carries quantization_config -- what AutoQuantizedModelForCausalLM.save_pretrained writes
    (modeling/base.py:324-336);
  * weights are saved by HF save_pretrained as SHARDED safetensors (+ model.safetensors.index.json); a second directory holds
    the same tensors as sharded .bin files (+ pytorch_model.bin.index.json), the other format _load_check_point accepts
    (modeling/base.py:118-172);
  * logits.npz: the reference model's own CPU forward (QuantLinearGPTQ.forward -> DequantizeLinearBlockWise + matmul, fp32)
    on fixed token ids -- the golden output the loader test compares the engine with.

    python tests/golden/make_ckpt_fixture.py        # writes tests/golden/ckpt_ref_gptq/{st,bin}/ and logits.npz
"""
import json
import os
import shutil
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference, REF  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ckpt_ref_gptq")
BITS, GS = 4, 32


def main():
    gptq, hqq, awq, marlin, cw = import_reference()
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_config_builder", os.path.join(REF, "qllm", "quantization", "config_builder.py"))
    cb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cb)
    from transformers import AutoModelForCausalLM, LlamaConfig
    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2,
                      vocab_size=64, max_position_embeddings=32, tie_word_embeddings=False)
    model = AutoModelForCausalLM.from_config(cfg, dtype=torch.float32).eval()
    table = {}
    for name, m in list(model.named_modules()):
        if isinstance(m, torch.nn.Linear) and ".layers." in name:
            N, K = m.weight.shape
            wg = m.weight.data.reshape(N, K // GS, GS)
            lo, hi = wg.amin(-1), wg.amax(-1)
            s = ((hi - lo) / 15).clamp_min(1e-5).half().float()
            z = torch.round(-lo / s).clamp(0, 15)
            q = gptq.QuantLinearGPTQ(BITS, GS, K, N, False, dtype=torch.float32)
            q.pack(m, s, z, None)                                  # the reference's packer
            parent = model.get_submodule(name.rsplit(".", 1)[0])
            setattr(parent, name.rsplit(".", 1)[1], q)
            table[name] = {"wbits": BITS, "groupsize": GS}
    qc = cb.GPTQConfig(damp_percent=0.01, group_size=GS, desc_act=False, bits=BITS, sym=False, allow_mix_bits=False,
                       true_sequential=False, static_groups=False, version="GPTQ", quant_method="gptq").to_dict()
    ids = torch.tensor([[1, 5, 9, 33, 2, 60, 7, 7, 12]])
    with torch.no_grad():
        logits = model(ids).logits.float().numpy()
    shutil.rmtree(OUT, ignore_errors=True)
    st, bn = os.path.join(OUT, "st"), os.path.join(OUT, "bin")
    model.config.quantization_config = qc
    model.save_pretrained(st, safe_serialization=True, max_shard_size="40KB")
    for d in (st, bn):
        os.makedirs(d, exist_ok=True)
        json.dump(table, open(os.path.join(d, "quant_config_by_layer.json"), "w"), indent=4)
        json.dump(qc, open(os.path.join(d, "quantize_config.json"), "w"), indent=4)
    shutil.copy(os.path.join(st, "config.json"), os.path.join(bn, "config.json"))
    # the .bin form: same tensors, two shards + index (what an older HF save_pretrained(safe_serialization=False) writes)
    sd = {k: v.detach().clone().contiguous() for k, v in model.state_dict().items()}
    keys = sorted(sd)
    half = len(keys) // 2
    wm = {}
    for i, part in enumerate((keys[:half], keys[half:])):
        fn = f"pytorch_model-{i + 1:05d}-of-00002.bin"
        torch.save({k: sd[k] for k in part}, os.path.join(bn, fn))
        wm.update({k: fn for k in part})
    json.dump({"metadata": {}, "weight_map": wm}, open(os.path.join(bn, "pytorch_model.bin.index.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(OUT, "logits.npz"), ids=ids.numpy(), logits=logits)
    tot = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(OUT) for f in fs)
    print("fixture written:", OUT, f"{tot / 1024:.0f} KB", sorted(os.listdir(st)))


if __name__ == "__main__":
    main()
