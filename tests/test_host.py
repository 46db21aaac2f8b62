"""CPU-only tests: the C-ABI library loads and exports every symbol include/b200q.h declares (no compute
without a GPU), and the host-side mirror of the reference's plug-in interface behaves like it."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import qllm_b200
from oracle import qlinear_oracle as O
from qllm_b200 import codec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200q.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200q_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 14
    lib = ctypes.CDLL(os.path.join(ROOT, "qllm_b200", "libb200q.so"))
    for n in names:
        assert hasattr(lib, n), f"libb200q.so does not export {n}"
    assert qllm_b200.lib.b200q_version() == 100
    assert qllm_b200.lib.b200q_gemv_max_m() == 8
    assert b"workspace" in qllm_b200.lib.b200q_strerror(-5)


def test_descriptor_struct_matches_header_layout():
    from qllm_b200._lib import Layer
    assert ctypes.sizeof(Layer) == 6 * 4 + 6 * 8
    assert Layer.qweight.offset == 24 and Layer.bias.offset == 56


def test_select_quant_linear_table():
    s = qllm_b200.select_quant_linear
    assert s("GPTQ", 4, "gptq") is qllm_b200.QuantLinearGPTQ
    assert s("GEMM", 4, "awq") is qllm_b200.WQLinear_GEMM
    assert s("AUTO", 4, "gptq") is qllm_b200.WQLinear_GEMM          # modelutils.py:61 (4-bit + engine present)
    assert s("AUTO", 3, "gptq") is qllm_b200.QuantLinearGPTQ
    assert s("MARLIN", 4, "gptq") is qllm_b200.QuantLinearMarlin
    assert s("GPTQ", 4, "hqq") is qllm_b200.QuantLinearHQQ
    assert s("ORT", 4, "gptq") is qllm_b200.QuantLinearORT            # modelutils.py:50-51
    assert s("GEMV", 4, "awq") is qllm_b200.WQLinear_GEMV             # not in the reference's table (dead there); offered here
    with pytest.raises(NotImplementedError):
        s("GPTQ", 4, "vptq")


@pytest.mark.parametrize("cls,bits,gs,K,N,shapes", [
    ("QuantLinearGPTQ", 4, 128, 4096, 11008, {"qweight": (512, 11008), "qzeros": (32, 1376), "scales": (32, 11008), "g_idx": (4096,)}),
    ("QuantLinearGPTQ", 3, 64, 128, 64, {"qweight": (12, 64), "qzeros": (2, 6), "scales": (2, 64)}),
    ("QuantLinearHQQ", 2, 64, 128, 64, {"qweight": (8, 64), "qzeros": (2, 64), "scales": (2, 64)}),
    ("WQLinear_GEMM", 4, 128, 4096, 4096, {"qweight": (4096, 512), "qzeros": (32, 512), "scales": (32, 4096)}),
    ("QuantLinearMarlin", 4, 128, 4096, 4096, {"qweight": (256, 8192), "scales": (32, 4096)}),
])
def test_buffer_names_and_shapes_match_reference(cls, bits, gs, K, N, shapes):
    """state_dict keys/shapes are what the reference registers (quant_linear_*.py __init__)."""
    layer = getattr(qllm_b200, cls)(bits, gs, K, N, True, dtype=torch.float16)
    sd = layer.state_dict()
    for name, shp in shapes.items():
        assert tuple(sd[name].shape) == shp, name
    assert "bias" in sd and sd["bias"].shape == (N,)
    assert layer.infeatures == K and layer.outfeatures == N and layer.bits == bits and layer.groupsize == gs


def test_make_mixbits_quant_linear_swaps_named_layers():
    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(128, 64, bias=False)
            self.blk = torch.nn.ModuleList([torch.nn.Linear(128, 256, bias=True)])
            self.head = torch.nn.Linear(64, 10)
    m = Toy().half()
    qllm_b200.make_mixbits_quant_linear(m, ["a", "blk.0"], {"a": {"wbits": 4, "groupsize": 64}, "blk.0": {"wbits": 8, "groupsize": 128}},
                                        target_layer=qllm_b200.QuantLinearGPTQ)
    assert isinstance(m.a, qllm_b200.QuantLinearGPTQ) and m.a.bits == 4 and m.a.groupsize == 64 and m.a.bias is None
    assert isinstance(m.blk[0], qllm_b200.QuantLinearGPTQ) and m.blk[0].bits == 8 and m.blk[0].bias is not None
    assert isinstance(m.head, torch.nn.Linear)


@pytest.mark.parametrize("cls,layout,bits,gs", [("QuantLinearGPTQ", "GPTQ", 4, 32), ("QuantLinearGPTQ", "GPTQ", 3, 32),
                                                 ("QuantLinearHQQ", "HQQ", 4, 64), ("WQLinear_GEMM", "GEMM", 4, 64),
                                                 ("QuantLinearMarlin", "MARLIN", 4, 128)])
def test_pack_unpack_roundtrip_matches_oracle_layout(cls, layout, bits, gs):
    """pack() of a weight that is exactly representable reproduces the oracle's packed bytes; unpack() inverts it."""
    K, N = 256, 256
    L = O.make_layer(layout, bits, gs, K, N, seed=5)
    W = O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "exact").astype(np.float32)            # [K, N], exact grid points
    lin = torch.nn.Linear(K, N, bias=False)
    lin.weight.data = torch.from_numpy(W.T.copy())
    if layout == "MARLIN":
        lin = lin.half()
    layer = getattr(qllm_b200, cls)(bits, gs, K, N, False, dtype=torch.float16)
    scales_ng = torch.from_numpy(L["s"].astype(np.float32).T.copy())                            # quantiser's [N, G]
    zeros_ng = None if layout == "MARLIN" else torch.from_numpy(np.asarray(L["z"]).astype(np.float32).T.copy())
    layer.pack(lin, scales_ng, zeros_ng, None)
    assert np.array_equal(layer.qweight.numpy(), L["qweight"])
    if layout in ("GPTQ", "GEMM"):
        assert np.array_equal(layer.qzeros.numpy(), L["qzeros"])
    assert np.array_equal(layer.scales.numpy().view(np.uint16), L["scales"].view(np.uint16))
    w, s, z = layer.unpack()
    assert np.allclose(w.float().numpy(), W.T, atol=2e-3 * np.abs(W).max())


def test_forward_refuses_cpu_tensors():
    layer = qllm_b200.QuantLinearGPTQ(4, 128, 128, 64, False, dtype=torch.float16)
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 128, dtype=torch.float16))


def test_autogptq_zero_fixup_host():
    d = np.load(os.path.join(ROOT, "tests", "golden", "gptq.npz"))
    layer = qllm_b200.QuantLinearGPTQ(4, 64, 128, 64, False, dtype=torch.float16)
    layer.qzeros = torch.from_numpy(d["autogptq_stored"].copy())
    layer.handle_qzeros_for_autogptq()
    assert np.array_equal(layer.qzeros.numpy(), d["autogptq_fixed"])


def test_bench_accounting_matches_the_survey_figures():
    """bench.py's algorithmic bytes are SURVEY section 8(d)'s: 0.51953 B/weight at int4 g128, 8,732,672 B for 4096^2 at M = 1,
    3.3645e9 B per Llama-2-7B token; column shards tile the output exactly."""
    import bench
    assert bench.alg_bytes(4096, 4096, 1) == 8732672
    assert bench.alg_bytes(4096, 11008, 1) == 23455232
    total = bench.BLOCKS * sum(bench.alg_bytes(K, N, 1) for _, K, N in bench.SHAPES)
    assert abs(total - 3.3645e9) / 3.3645e9 < 2e-3          # the survey's per-weight figure leaves out the 5 MB of x / y per token
    for world in (1, 2, 4, 8):
        for _, _, N in bench.SHAPES:
            cuts = [bench.shard_cols(N, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert all((c1 - c0) % 32 == 0 and c1 > c0 for c0, c1 in cuts)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/b200q.h compiles as C99 (no C++-isms, no torch types) and a C program links
    against libb200q.so -- what a cgo / JNI / ctypes binding relies on."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use_b200q.c"
    src.write_text('#include "b200q.h"\n'
                   "int main(void) { b200q_layer l; b200q_peer_sync s; (void)l; (void)s;\n"
                   "  return (b200q_version() == B200Q_VERSION && b200q_gemv_max_m() > 0 && b200q_strerror(B200Q_ERR_SHAPE) != 0) ? 0 : 1; }\n")
    libdir = os.path.join(ROOT, "qllm_b200")
    exe = tmp_path / "use_b200q"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-L", libdir, "-lb200q", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_bench_handoff_schedule():
    """bench.py's fused-step call sequence: unique y_seq below the tag stride, and every consumer awaits the tag of the call
    that wrote its input last (sibling groups and the one-call-per-layer form of act-order models)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    F = bench.FusedShardedStep
    for calls, x_of in ((F.CALLS, F.X_OF), (F.CALLS_ACT, F.X_OF_ACT)):
        sched, stride = bench.handoff_schedule(calls, x_of, 3)
        seqs = [c[3] for c in sched]
        assert seqs == list(range(1, len(calls) * 3 + 1)) and stride == len(calls) * 3 + 1
        last_writer = {}
        for blk, names, x_name, y_seq, x_seq in sched:
            if x_name is None:
                assert blk == 0 and x_seq == 0               # the token's input is plain fp16
            else:
                assert x_seq == last_writer[x_name] and 0 < x_seq < y_seq
            for n in names:
                last_writer[n] = y_seq
        assert sched[-1][1] == ("down",)
        # every block after the first reads the previous block's down
        firsts = [c for c in sched if c[0] == 1 and x_of[(c[3] - 1) % len(calls)] is None]
        assert firsts and all(c[2] == "down" and c[4] == len(calls) for c in firsts)
