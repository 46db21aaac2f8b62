"""Generate golden vectors by EXECUTING the reference's own Python implementation on CPU.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference has no tests / known-answer vectors for this path, so these
fixtures -- packed buffers produced by the reference's `pack()` and outputs of the reference's
`forward()` / `unpack()` on CPU -- are what pins `oracle/qlinear_oracle.py`.
Nothing from the reference is copied: its modules are imported from where they lie.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("QLLM_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    """qllm/modeling/__init__.py pulls accelerate/texttable/primefac (absent here); register empty
    package shells so that only q_layers/* and utils/logger.py are executed."""
    for name, sub in (("qllm", "qllm"), ("qllm.modeling", "qllm/modeling"), ("qllm.utils", "qllm/utils")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m
    # ext_package_checker.py:18-20 calls torch.cuda.get_device_properties(0) unguarded
    torch.cuda.get_device_properties = lambda i=0: types.SimpleNamespace(major=10, minor=0)
    import importlib
    comm = types.ModuleType("qllm.utils.comm_utils")
    sys.modules["qllm.utils.comm_utils"] = comm
    sys.modules["qllm.utils"].comm_utils = comm
    gptq = importlib.import_module("qllm.modeling.q_layers.quant_linear_gptq")
    hqq = importlib.import_module("qllm.modeling.q_layers.quant_linear_hqq")
    awq = importlib.import_module("qllm.modeling.q_layers.quant_linear_awq")
    marlin = importlib.import_module("qllm.modeling.q_layers.quant_linear_marlin")
    cw = importlib.import_module("qllm.modeling.q_layers.compress_weight")
    return gptq, hqq, awq, marlin, cw


class _TorchCpuProxy:
    """`torch` look-alike whose .device("cuda") is the CPU: QuantLinearMarlin.pack hard-codes
    torch.device("cuda") (quant_linear_marlin.py:103)."""

    def __getattr__(self, k):
        if k == "device":
            return lambda *_a, **_k: torch.device("cpu")
        return getattr(torch, k)


def synth(K, N, G, bits, seed, sym=False):
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(N, K, generator=g) * 0.02                    # nn.Linear.weight layout [N, K]
    maxq = 2 ** bits - 1
    Wg = W.reshape(N, G, K // G)
    if sym:
        s = (Wg.abs().amax(-1) / ((maxq + 1) // 2 - 1)).clamp_min(1e-4)
        z = torch.full_like(s, (maxq + 1) // 2)
    else:
        mx, mn = Wg.amax(-1), Wg.amin(-1)
        s = ((mx - mn) / maxq).clamp_min(1e-4)
        z = torch.round(-mn / s).clamp(0, maxq)
    return W, s.half().float(), z                                # scales/zeros [N, G]


def main():
    gptq, hqq, awq, marlin, cw = import_reference()
    os.environ["COMPATIBLE_WITH_AUTOGPTQ"] = "0"
    torch.manual_seed(0)

    # ---- 1. raw bit-stream codec, every bit width (compress_weight.py:46-92) -----------------
    codec = {}
    for bits in range(2, 9):
        g = torch.Generator().manual_seed(100 + bits)
        q = torch.randint(0, 2 ** bits, (96, 24), generator=g, dtype=torch.int32)
        packed = torch.zeros((96 * bits // 32, 24), dtype=torch.int32)
        cw.general_pack_on_row(packed, q, bits)
        back = torch.zeros_like(q)
        cw.general_unpack_on_row(packed, back, bits)
        assert (back == q).all()
        codec[f"q{bits}"] = q.numpy()
        codec[f"p{bits}"] = packed.numpy()
    np.savez_compressed(os.path.join(OUT, "codec.npz"), **codec)

    # ---- 2. GPTQ layers: pack + CPU forward (quant_linear_gptq.py) ---------------------------
    out = {}
    cases = [(4, 128, 256, 64, False), (4, 32, 128, 64, True), (2, 64, 128, 64, False),
             (3, 32, 128, 64, False), (8, 128, 256, 32, False), (4, -1, 128, 64, False)]
    for ci, (bits, gs, K, N, act) in enumerate(cases):
        G = 1 if gs == -1 else K // gs
        W, s, z = synth(K, N, G, bits, 200 + ci)
        lin = torch.nn.Linear(K, N, bias=True)
        lin.weight.data = W.clone()
        lin.bias.data = torch.randn(N) * 0.1
        layer = gptq.QuantLinearGPTQ(bits, gs, K, N, True, dtype=torch.float32)
        g_idx = None
        if act:
            perm = torch.randperm(K, generator=torch.Generator().manual_seed(7))
            g_idx = (torch.arange(K) // gs)[perm].to(torch.int32)
        layer.bias = lin.bias.data.clone()
        layer.pack(lin, s, z, g_idx)
        x = torch.randn(5, K, generator=torch.Generator().manual_seed(300 + ci))
        layer.scales = layer.scales.float()
        y32 = layer(x)                                             # fp32 torch path on CPU
        fw, s_u, z_u = layer.unpack()
        p = f"c{ci}_"
        out.update({p + "meta": np.array([bits, gs, K, N, int(act)]), p + "qweight": layer.qweight.numpy(),
                    p + "qzeros": layer.qzeros.numpy(), p + "scales": layer.scales.half().numpy(),
                    p + "g_idx": layer.g_idx.numpy(), p + "bias": layer.bias.half().numpy(),
                    p + "x": x.half().numpy(), p + "y32": y32.detach().numpy(),
                    p + "unpack_w": fw.float().numpy(), p + "unpack_z": z_u.numpy()})
    # AutoGPTQ z-1 storage + loader fix-up (quant_linear_gptq.py:119-134)
    layer = gptq.QuantLinearGPTQ(4, 64, 128, 64, False, dtype=torch.float16)
    zt = torch.randint(0, 16, (2, 64), generator=torch.Generator().manual_seed(9), dtype=torch.int32)
    os.environ["COMPATIBLE_WITH_AUTOGPTQ"] = "1"
    layer.pack_qzeros(zt, "cpu")
    os.environ["COMPATIBLE_WITH_AUTOGPTQ"] = "0"
    out["autogptq_z"] = zt.numpy()
    out["autogptq_stored"] = layer.qzeros.numpy().copy()
    layer.handle_qzeros_for_autogptq()
    out["autogptq_fixed"] = layer.qzeros.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "gptq.npz"), **out)

    # ---- 3. HQQ layers (fp16 zeros; quant_linear_hqq.py) --------------------------------------
    out = {}
    for ci, (bits, gs, K, N) in enumerate([(4, 64, 128, 64), (2, 16, 64, 32), (3, 64, 128, 32), (8, 128, 128, 32)]):
        G = K // gs
        W, s, z = synth(K, N, G, bits, 400 + ci)
        lin = torch.nn.Linear(K, N, bias=False)
        lin.weight.data = W.clone()
        layer = hqq.QuantLinearHQQ(bits, gs, K, N, False, dtype=torch.float32)
        layer.pack(lin, s, z, None)
        layer.scales = layer.scales.float()
        layer.qzeros = layer.qzeros.float()
        x = torch.randn(3, K, generator=torch.Generator().manual_seed(500 + ci))
        y32 = layer(x)
        p = f"c{ci}_"
        out.update({p + "meta": np.array([bits, gs, K, N]), p + "qweight": layer.qweight.numpy(),
                    p + "qzeros": layer.qzeros.half().numpy(), p + "scales": layer.scales.half().numpy(),
                    p + "x": x.half().numpy(), p + "y32": y32.detach().numpy()})
    np.savez_compressed(os.path.join(OUT, "hqq.npz"), **out)

    # ---- 4. AWQ GEMM layout: pack + unpack (no CPU forward exists; quant_linear_awq.py) -------
    out = {}
    for ci, (gs, K, N) in enumerate([(128, 256, 64), (32, 64, 128)]):
        G = K // gs
        W, s, z = synth(K, N, G, 4, 600 + ci)
        lin = torch.nn.Linear(K, N, bias=False)
        lin.weight.data = W.clone()
        layer = awq.WQLinear_GEMM(4, gs, K, N, False, dtype=torch.float16)
        layer.pack(lin, s, z, None)
        fw, s_u, z_u = layer.unpack()
        p = f"c{ci}_"
        out.update({p + "meta": np.array([4, gs, K, N]), p + "qweight": layer.qweight.numpy(),
                    p + "qzeros": layer.qzeros.numpy(), p + "scales": layer.scales.numpy(),
                    p + "unpack_w": fw.float().numpy(), p + "unpack_z": z_u.numpy(),
                    p + "int_w": layer.unpack_qweight("cpu").numpy()})
    np.savez_compressed(os.path.join(OUT, "awq.npz"), **out)

    # ---- 4b. AWQ GEMV layout: the reference's own packer (quant_linear_awq.py:186-254).  accelerate_pack_on_device is
    #      declared without @classmethod and passes `linear.weight.device` where the ctor expects a dtype, so it is
    #      driven with cls given explicitly and a stand-in `linear` whose weight.device IS the dtype ----------------
    out = {}
    for ci, (gs, K, N) in enumerate([(128, 256, 64), (64, 256, 32), (32, 128, 64), (128, 2048, 32)]):
        G = K // gs
        W, s, z = synth(K, N, G, 4, 650 + ci)
        # the reference's GEMV packer does not clamp (a rounded 16 would spill into the next nibble): give it weights that
        # sit exactly on the 4-bit grid
        se, ze = s.repeat_interleave(gs, dim=1), z.repeat_interleave(gs, dim=1)
        W = (torch.clamp(torch.round(W / se + ze), 0, 15) - ze) * se
        fake = types.SimpleNamespace(in_features=K, out_features=N, bias=None,
                                     weight=types.SimpleNamespace(data=W.clone(), device=torch.float16))
        layer = awq.WQLinear_GEMV.accelerate_pack_on_device(awq.WQLinear_GEMV, fake, 4, gs, False, s, z)
        intw = torch.round((W + (z * s).repeat_interleave(gs, dim=1)) / s.repeat_interleave(gs, dim=1)).to(torch.int32)   # test input record
        p = f"c{ci}_"
        out.update({p + "meta": np.array([4, gs, K, N]), p + "qweight": layer.qweight.numpy(), p + "qzeros": layer.qzeros.numpy(),
                    p + "scales": layer.scales.numpy(), p + "int_w": intw.t().contiguous().numpy(),
                    p + "int_z": z.to(torch.int32).t().contiguous().numpy(), p + "nat_scales": s.half().t().contiguous().numpy()})
    np.savez_compressed(os.path.join(OUT, "awq_gemv.npz"), **out)

    # ---- 4c. ORT MatMulNBits blobs: the reference's QuantLinearORT.pack + its torch forward (quant_linear_onnxruntime.py) ----
    import importlib
    ort = importlib.import_module("qllm.modeling.q_layers.quant_linear_onnxruntime")
    out = {}
    for ci, (gs, K, N) in enumerate([(128, 256, 64), (32, 128, 32), (64, 384, 64)]):
        G = K // gs
        W, s, z = synth(K, N, G, 4, 680 + ci)
        se, ze = s.repeat_interleave(gs, dim=1), z.repeat_interleave(gs, dim=1)
        W = (torch.clamp(torch.round(W / se + ze), 0, 15) - ze) * se           # on the 4-bit grid (the packer does not clamp)
        lin = torch.nn.Linear(K, N, bias=False)
        lin.weight.data = W.clone()
        layer = ort.QuantLinearORT(4, gs, K, N, False, dtype=torch.float32)
        layer.pack(lin, s, z.to(torch.int32), None)          # integer zeros: a float tensor would select the float-zero blob form (:119,:130)
        x = torch.randn(3, K, generator=torch.Generator().manual_seed(ci))
        y = layer(x)
        intw = torch.round((W + ze * se) / se).to(torch.int32)
        p = f"c{ci}_"
        out.update({p + "meta": np.array([4, gs, K, N]), p + "qweight": layer.qweight.numpy(), p + "qzeros": layer.qzeros.numpy(),
                    p + "scales": layer.scales.numpy(), p + "int_w": intw.t().contiguous().numpy(),
                    p + "int_z": z.to(torch.int32).t().contiguous().numpy(), p + "nat_scales": s.t().contiguous().numpy(),
                    p + "x": x.numpy(), p + "y32": y.detach().numpy()})
    np.savez_compressed(os.path.join(OUT, "ort.npz"), **out)

    # ---- 5. Marlin layout: pack only (unpack is NotImplemented; quant_linear_marlin.py) -------
    out = {}
    marlin.torch = _TorchCpuProxy()
    for ci, (gs, K, N) in enumerate([(128, 256, 256), (-1, 128, 256)]):
        G = 1 if gs == -1 else K // gs
        W, s, z = synth(K, N, G, 4, 700 + ci, sym=True)
        lin = torch.nn.Linear(K, N, bias=False)
        lin.weight.data = W.half()
        layer = marlin.QuantLinearMarlin(4, gs if gs != -1 else K, K, N, False)
        layer.pack(lin, s.half(), None, None)
        # the integer the reference quantised to, restated here only to record the test input
        gsz = K if gs == -1 else gs
        wq = torch.clamp(torch.round(lin.weight.data.t().reshape(G, gsz, N)          # fp16 division, as
                                     / s.half().t().reshape(G, 1, N)).int() + 8, 0, 15).reshape(K, N)  # :110-113
        p = f"c{ci}_"
        out.update({p + "meta": np.array([4, gsz, K, N]), p + "qweight": layer.qweight.numpy(),
                    p + "scales": layer.scales.numpy(), p + "nat_scales": s.half().t().contiguous().numpy(),
                    p + "int_w": wq.numpy()})
    marlin.torch = torch
    np.savez_compressed(os.path.join(OUT, "marlin.npz"), **out)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
