"""Pins oracle/qlinear_oracle.py against vectors produced by the reference's own Python code
(tests/golden/make_golden.py). Integer work is bit-exact; float work within stated tolerance."""
import os

import numpy as np
import pytest

from oracle import qlinear_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("bits", range(2, 9))
def test_bitstream_codec_matches_reference(golden_dir, bits):
    d = _load(golden_dir, "codec.npz")
    q, p = d[f"q{bits}"], d[f"p{bits}"]
    assert np.array_equal(O.pack_rows(q, bits), p)            # compress_weight.py:46-51
    assert np.array_equal(O.unpack_rows(p, bits, q.shape[0]), q)   # compress_weight.py:87-92


def _gptq_cases(d):
    i = 0
    while f"c{i}_meta" in d:
        yield i
        i += 1


def test_gptq_unpack_and_forward(golden_dir):
    d = _load(golden_dir, "gptq.npz")
    for i in _gptq_cases(d):
        bits, gs, K, N, act = (int(v) for v in d[f"c{i}_meta"])
        p = f"c{i}_"
        q, z, s, gi = O.unpack_layer("GPTQ", bits, gs, K, N, d[p + "qweight"], d[p + "qzeros"],
                                     d[p + "scales"], d[p + "g_idx"])
        assert np.array_equal(z, d[p + "unpack_z"]), f"case {i}: qzeros unpack"
        # re-pack is the identity on the reference's bytes
        assert np.array_equal(O.gptq_pack_qweight(q, bits), d[p + "qweight"])
        assert np.array_equal(O.gptq_pack_qzeros(z, bits), d[p + "qzeros"])
        # reference unpack() returns fp16(q*s - z*s)^T in fp32 scales here -> compare in fp32
        W = O.dequant(q, z, s, gi, "exact")
        assert np.allclose(W.T, d[p + "unpack_w"], rtol=0, atol=2e-3 * np.abs(W).max())
        # reference fp32 CPU forward (DequantizeLinearBlockWise + matmul + bias)
        y = O.matmul_ref(d[p + "x"], W, d[p + "bias"], acc=np.float64)
        ref = d[p + "y32"]
        assert np.abs(y - ref).max() <= 1e-3 * np.abs(ref).max(), f"case {i}: forward"


def test_autogptq_zero_fixup(golden_dir):
    d = _load(golden_dir, "gptq.npz")
    z = d["autogptq_z"]
    assert np.array_equal(O.gptq_pack_qzeros(z, 4, zero_bias=1), d["autogptq_stored"])
    assert np.array_equal(O.autogptq_fix_qzeros(d["autogptq_stored"], 4, z.shape[1]), d["autogptq_fixed"])
    assert np.array_equal(O.gptq_unpack_qzeros(d["autogptq_fixed"], 4, z.shape[1]), z)


def test_hqq_forward(golden_dir):
    d = _load(golden_dir, "hqq.npz")
    i = 0
    while f"c{i}_meta" in d:
        bits, gs, K, N = (int(v) for v in d[f"c{i}_meta"])
        p = f"c{i}_"
        y = O.forward("HQQ", bits, gs, K, N, d[p + "x"], d[p + "qweight"], d[p + "qzeros"], d[p + "scales"],
                      mode="exact", acc=np.float64)
        ref = d[p + "y32"]
        assert np.abs(y - ref).max() <= 1e-3 * np.abs(ref).max(), f"case {i}"
        i += 1
    assert i == 4


def test_awq_layout(golden_dir):
    d = _load(golden_dir, "awq.npz")
    for i in range(2):
        bits, gs, K, N = (int(v) for v in d[f"c{i}_meta"])
        p = f"c{i}_"
        q = O.awq_unpack_qweight(d[p + "qweight"])
        z = O.awq_unpack_qzeros(d[p + "qzeros"])
        # reference unpack_qweight returns [K, N]; unpack_qzeros returns [G, N]
        assert np.array_equal(q, d[p + "int_w"])
        assert np.array_equal(z, d[p + "unpack_z"])
        assert np.array_equal(O.awq_pack_qweight(q), d[p + "qweight"])
        assert np.array_equal(O.awq_pack_qzeros(z), d[p + "qzeros"])
        W = O.dequant(q, z, d[p + "scales"], O.default_g_idx(K, gs), "torch")
        assert np.array_equal(W.T.astype(np.float32), d[p + "unpack_w"])     # fp16 bit-exact


def test_marlin_layout(golden_dir):
    d = _load(golden_dir, "marlin.npz")
    for i in range(2):
        bits, gs, K, N = (int(v) for v in d[f"c{i}_meta"])
        p = f"c{i}_"
        qw, sp = O.marlin_pack(d[p + "int_w"], d[p + "nat_scales"], gs)
        assert np.array_equal(qw, d[p + "qweight"])
        assert np.array_equal(sp.view(np.uint16), d[p + "scales"].view(np.uint16))
        q, s = O.marlin_unpack(d[p + "qweight"], d[p + "scales"], gs, K)
        assert np.array_equal(q, d[p + "int_w"])
        assert np.array_equal(s.view(np.uint16), d[p + "nat_scales"].view(np.uint16))


def test_marlin_closed_form_word00():
    """SURVEY Appendix A.7 worked example: word [0,0] nibbles = (k,n):
    (0,0),(8,0),(0,8),(8,8),(1,0),(9,0),(1,8),(9,8)."""
    K, N = 16, 64
    q = (np.arange(K)[:, None] * 0 + 0).astype(np.int32) + np.zeros((K, N), np.int32)
    marks = [(0, 0), (8, 0), (0, 8), (8, 8), (1, 0), (9, 0), (1, 8), (9, 8)]
    for v, (k, n) in enumerate(marks):
        q[k, n] = v + 1
    qw, _ = O.marlin_pack(q, np.ones((1, N), np.float16), K)
    w = int(qw.view(np.uint32)[0, 0])
    assert [(w >> (4 * i)) & 0xF for i in range(8)] == list(range(1, 9))


@pytest.mark.parametrize("layout,bits", [("GPTQ", 2), ("GPTQ", 3), ("GPTQ", 4), ("GPTQ", 8),
                                          ("HQQ", 4), ("GEMM", 4), ("MARLIN", 4)])
def test_make_layer_roundtrip(layout, bits):
    L = O.make_layer(layout, bits, 64 if layout != "MARLIN" else 128, 256, 256, seed=3,
                     act_order=(layout == "GPTQ" and bits == 4))
    q, z, s, gi = O.unpack_layer(layout, bits, L["group_size"], 256, 256, L["qweight"], L["qzeros"],
                                 L["scales"], L["g_idx"])
    assert np.array_equal(q, L["q"])
    assert np.array_equal(np.asarray(z), np.asarray(L["z"]))
    assert np.array_equal(s.view(np.uint16), L["s"].view(np.uint16))


def test_cpu_baseline_port_pinned_to_reference_forward(golden_dir):
    """oracle/cpu_baseline.py (bench.py's cpu_baseline / --impl reference arm) against the reference's own
    QuantLinearGPTQ.forward outputs and unpack() weights stored in gptq.npz (quant_linear_gptq.py:13-52, :136-143)."""
    import torch
    from oracle import cpu_baseline as C
    d = _load(golden_dir, "gptq.npz")
    pinned = 0
    for i in _gptq_cases(d):
        bits, gs, K, N, act = (int(v) for v in d[f"c{i}_meta"])
        if bits not in (2, 4, 8) or act:                     # the port restates the 2/4/8-bit, g_idx-free branch
            continue
        p = f"c{i}_"
        qw, qz = torch.from_numpy(d[p + "qweight"]), torch.from_numpy(d[p + "qzeros"])
        sc = torch.from_numpy(d[p + "scales"]).float()
        W = C.dequantize_blockwise(qw, sc, qz, gs, bits, K)
        ref_w = d[p + "unpack_w"].T                          # reference unpack(): [N, K] fp32
        assert np.abs(W.numpy() - ref_w).max() <= 2e-3 * np.abs(ref_w).max(), f"case {i}: weights"
        x = torch.from_numpy(d[p + "x"]).float()
        y = C.quant_linear_forward(x, qw, sc, qz, gs, bits, K).numpy() + d[p + "bias"].astype(np.float32)
        ref = d[p + "y32"]
        assert np.abs(y - ref).max() <= 1e-3 * np.abs(ref).max(), f"case {i}: forward"
        # fp16 arithmetic (what the bench times) stays within fp16 accumulation error of the same outputs
        y16 = C.quant_linear_forward(x.half(), qw, sc.half(), qz, gs, bits, K).float().numpy() + d[p + "bias"].astype(np.float32)
        assert np.abs(y16 - ref).max() <= 2e-2 * np.abs(ref).max(), f"case {i}: fp16 forward"
        pinned += 1
    assert pinned >= 3


def test_awq_gemv_layout_matches_reference_packer(golden_dir):
    """WQLinear_GEMV buffers written by the reference's own accelerate_pack_on_device (quant_linear_awq.py:186-254):
    the oracle's unpack recovers the integers bit-exactly and its pack reproduces the reference's bytes."""
    d = _load(golden_dir, "awq_gemv.npz")
    i = 0
    while f"c{i}_meta" in d:
        bits, gs, K, N = (int(v) for v in d[f"c{i}_meta"])
        p = f"c{i}_"
        q, z, s = O.awq_gemv_unpack(d[p + "qweight"], d[p + "qzeros"], d[p + "scales"], K, gs)
        assert np.array_equal(q, d[p + "int_w"]), f"case {i}: qweight"
        assert np.array_equal(z, d[p + "int_z"]), f"case {i}: qzeros"
        assert np.array_equal(s.view(np.uint16), d[p + "nat_scales"].view(np.uint16)), f"case {i}: scales"
        qw, qz, sp = O.awq_gemv_pack(q, z, s, gs)
        assert np.array_equal(qw, d[p + "qweight"]) and np.array_equal(qz, d[p + "qzeros"])
        assert np.array_equal(sp.view(np.uint16), d[p + "scales"].view(np.uint16))
        assert d[p + "qzeros"].shape[1] == O.awq_gemv_zeros_width(K, gs)
        i += 1
    assert i == 4


def test_ort_blob_layout_matches_reference(golden_dir):
    """QuantLinearORT buffers written by the reference's pack() and outputs of its torch forward
    (dequantize_blockwise_4bits + matmul, quant_linear_onnxruntime.py:31-82): unpack / pack bit-exact, forward within 1e-3."""
    d = _load(golden_dir, "ort.npz")
    i = 0
    while f"c{i}_meta" in d:
        bits, gs, K, N = (int(v) for v in d[f"c{i}_meta"])
        p = f"c{i}_"
        q, z, s = O.ort_unpack(d[p + "qweight"], d[p + "qzeros"], d[p + "scales"], K, N, gs)
        assert np.array_equal(q, d[p + "int_w"]) and np.array_equal(z, d[p + "int_z"])
        assert np.array_equal(s, d[p + "nat_scales"])
        qw, qz, sp = O.ort_pack(q, z, s, gs)
        assert np.array_equal(qw, d[p + "qweight"]) and np.array_equal(qz, d[p + "qzeros"]) and np.array_equal(sp, d[p + "scales"])
        W = O.dequant(q, z, s, O.default_g_idx(K, gs), "exact")
        y = O.matmul_ref(d[p + "x"], W, None, acc=np.float64)
        assert np.abs(y - d[p + "y32"]).max() <= 1e-3 * np.abs(d[p + "y32"]).max()
        i += 1
    assert i == 3
