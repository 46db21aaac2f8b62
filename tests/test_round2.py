"""Round-2 host-shim / hygiene tests (GPU): one packed copy for AWQ / Marlin layers with an exact round trip back to the
checkpoint format, cache invalidation on in-place updates, empty batches through fused sibling groups, NaN / Inf
propagation in the integer decode path, B200Q_ERR_ARCH plumbing, TMA descriptor cache."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import qlinear_oracle as O
from tests.util import layer_from_dict, oracle_forward, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.mark.parametrize("layout,gs,K,N", [("GEMM", 128, 512, 256), ("GEMM", 64, 1024, 128), ("MARLIN", 128, 512, 256), ("MARLIN", -1, 256, 512)])
def test_one_packed_copy_and_exact_round_trip(layout, gs, K, N):
    """After the first forward an AWQ / Marlin layer holds ONE packed copy (the K-packed re-layout); state_dict() returns
    the checkpoint's own bytes, rebuilt by b200q_repack_from_gptq4 (bit-exact), and load_state_dict() takes them back."""
    import qllm_b200
    L = O.make_layer(layout, 4, gs, K, N, seed=K + N, bias=True)
    layer = layer_from_dict(L)
    x = torch.randn(3, K, dtype=torch.float16, device="cuda")
    y0 = layer(x)
    assert layer.qweight.numel() == 0 and layer.scales.numel() == 0          # checkpoint-format buffers released
    for M in (1, 5, 40):                                                      # every kernel class runs on the remaining copy
        xm = torch.randn(M, K, dtype=torch.float16, device="cuda")
        assert rel_err(layer(xm).float().cpu().numpy(), oracle_forward(L, xm.cpu().numpy())) < TOL
    sd = layer.state_dict()
    assert np.array_equal(sd["qweight"].cpu().numpy(), L["qweight"])
    assert np.array_equal(sd["scales"].cpu().numpy().view(np.uint16), np.asarray(L["scales"]).view(np.uint16))
    if layout == "GEMM":
        assert np.array_equal(sd["qzeros"].cpu().numpy(), L["qzeros"])
    # a different checkpoint loaded into the same module: buffers regain their shapes, outputs follow the new weights
    L2 = O.make_layer(layout, 4, gs, K, N, seed=K + N + 1, bias=True)
    sd2 = {k: v.clone() for k, v in layer_from_dict(L2).state_dict().items()}
    layer.load_state_dict(sd2)
    assert layer.qweight.numel() > 0
    assert rel_err(layer(x).float().cpu().numpy(), oracle_forward(L2, x.cpu().numpy())) < TOL
    w, s, z = layer.unpack()                                                  # unpack() works on a consolidated layer too
    assert w.shape == (N, K)


def test_in_place_update_invalidates_cached_relayouts():
    """copy_() into a buffer keeps data_ptr() but must drop the descriptor / act-order re-layout built from the old bytes."""
    K, N = 512, 128
    La = O.make_layer("GPTQ", 4, 128, K, N, seed=1, act_order=True)
    Lb = O.make_layer("GPTQ", 4, 128, K, N, seed=2, act_order=True)
    layer = layer_from_dict(La)
    x = torch.randn(1, K, dtype=torch.float16, device="cuda")
    assert rel_err(layer(x).float().cpu().numpy(), oracle_forward(La, x.cpu().numpy())) < TOL
    with torch.no_grad():
        layer.qweight.copy_(torch.from_numpy(Lb["qweight"]))
        layer.qzeros.copy_(torch.from_numpy(Lb["qzeros"]))
        layer.scales.copy_(torch.from_numpy(Lb["scales"]))
        layer.g_idx.copy_(torch.from_numpy(Lb["g_idx"].astype(np.int32)))
    assert rel_err(layer(x).float().cpu().numpy(), oracle_forward(Lb, x.cpu().numpy())) < TOL


def test_empty_batch_through_fused_siblings():
    """An MoE expert that receives zero routed tokens: x of shape (0, K) on a model with fuse_siblings() installed."""
    import qllm_b200

    class Expert(nn.Module):
        def __init__(self):
            super().__init__()
            self.w1 = layer_from_dict(O.make_layer("GPTQ", 4, 128, 256, 128, seed=3))
            self.w3 = layer_from_dict(O.make_layer("GPTQ", 4, 128, 256, 128, seed=4))

    m = Expert()
    assert qllm_b200.fuse_siblings(m) == 1
    x0 = torch.zeros(0, 256, dtype=torch.float16, device="cuda")
    assert m.w1(x0).shape == (0, 128) and m.w3(x0).shape == (0, 128)
    x1 = torch.randn(2, 256, dtype=torch.float16, device="cuda")
    a, b = m.w1(x1), m.w3(x1)
    assert a.shape == b.shape == (2, 128)


def test_fuse_siblings_skips_act_order_before_first_forward():
    import qllm_b200

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = layer_from_dict(O.make_layer("GPTQ", 4, 128, 256, 128, seed=5, act_order=True))
            self.k_proj = layer_from_dict(O.make_layer("GPTQ", 4, 128, 256, 128, seed=6, act_order=True))
            self.v_proj = layer_from_dict(O.make_layer("GPTQ", 4, 128, 256, 128, seed=7, act_order=True))

    assert qllm_b200.fuse_siblings(Attn()) == 0        # act_order is None until the first forward: detected from g_idx instead


@pytest.mark.parametrize("layout", ["GPTQ", "GEMM"])
def test_integer_decode_path_propagates_non_finite(layout):
    """The reference's fp16 FMA kernels propagate NaN / Inf; the integer (base-128 digit) path must not turn them into
    finite numbers (fmaxf drops NaN, float2int saturates Inf)."""
    K, N = 1024, 256
    layer = layer_from_dict(O.make_layer(layout, 4, 128, K, N, seed=8))
    for M in (1, 2):
        for bad in (float("nan"), float("inf"), float("-inf")):
            x = torch.randn(M, K, dtype=torch.float16, device="cuda")
            x[M - 1, 333] = bad
            y = layer(x)
            assert not torch.isfinite(y[M - 1]).any(), f"{layout} M={M} {bad}: finite outputs"
            if M == 2:
                assert torch.isfinite(y[0]).all()              # the other token is untouched


def test_reverse_repack_rejects_bad_targets():
    import qllm_b200
    from qllm_b200._lib import LAYOUT_HQQ
    layer = layer_from_dict(O.make_layer("GPTQ", 4, 128, 256, 128, seed=9))
    d = layer._descriptor()
    out = torch.zeros(256 * 128 // 8, dtype=torch.int32, device="cuda")
    st = qllm_b200.lib.b200q_repack_from_gptq4(ctypes.byref(d), LAYOUT_HQQ, out.data_ptr(), out.data_ptr(), out.data_ptr(), None)
    assert st == -3


def test_gemm_first_call_under_graph_capture():
    """No allocation / synchronisation inside the launch path: the very first tcgen05 GEMM of a process may be captured."""
    layer = layer_from_dict(O.make_layer("GPTQ", 4, 128, 512, 256, seed=10))
    x = torch.randn(64, 512, dtype=torch.float16, device="cuda")
    layer._descriptor()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        y_eager = layer(x)                      # sets the kernel attribute (cudaFuncSetAttribute is not capturable)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            y = layer(x)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(y, y_eager)


# ---- the AWQ-GEMV layout (WQLinear_GEMV) -------------------------------------------------------
@pytest.mark.parametrize("gs,K,N", [(128, 512, 128), (64, 512, 64), (32, 256, 96), (128, 4096, 256)])
def test_awq_gemv_layout(gs, K, N):
    L = O.make_layer("GEMV", 4, gs, K, N, seed=K + N + gs, bias=True)
    layer = layer_from_dict(L)
    q, z = layer.unpack_int()                                             # read in place (generic element decoders)
    assert np.array_equal(q.cpu().numpy(), L["q"]) and np.array_equal(z.cpu().numpy(), L["z"])
    W = layer.dequantize().cpu().numpy()
    assert np.array_equal(W.view(np.uint16), O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "engine").view(np.uint16))
    for M in (1, 2, 7, 40, 300):                                          # every kernel class, on the exact K-packed re-layout
        x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
        y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
        assert rel_err(y, oracle_forward(L, x)) < TOL
    sd = layer.state_dict()                                               # checkpoint bytes restored exactly
    assert np.array_equal(sd["qweight"].cpu().numpy(), L["qweight"])
    assert np.array_equal(sd["qzeros"].cpu().numpy(), L["qzeros"])
    assert np.array_equal(sd["scales"].cpu().numpy().view(np.uint16), L["scales"].view(np.uint16))


def test_awq_gemv_vs_reference_cuda():
    """The reference's own WQLinear_GEMV kernels (gemv_forward_cuda, gemv_cuda.cu:60-186; gemmv2_forward_cuda) on the same bytes."""
    import glob, os, sys
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    if not glob.glob(os.path.join(ref_dir, "awq_inference_engine*.so")):
        pytest.skip("oracle/_ref not built")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import awq_inference_engine as awq
    K, N, gs = 1024, 512, 128
    L = O.make_layer("GEMV", 4, gs, K, N, seed=77)
    layer = layer_from_dict(L)
    for M in (1, 4):
        x = torch.randn(M, K, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(M))
        ref = awq.gemv_forward_cuda(x, layer.qweight, layer.scales, layer.qzeros, gs) if layer.qweight.numel() else None
        if ref is None:
            qw, qz, sc = layer._native_tensors()
            ref = awq.gemv_forward_cuda(x, qw, sc, qz, gs)
        torch.cuda.synchronize()
        y = layer(x)
        torch.cuda.synchronize()
        err = ((y.double() - ref.double()).abs().max() / ref.double().abs().max()).item()
        assert err < 2e-3, err


# ---- pack-mode conversion as an exact integer re-layout on the GPU -----------------------------------
@pytest.mark.parametrize("src,dst", [("GPTQ", "GEMM"), ("GPTQ", "GEMV"), ("GEMM", "GPTQ"), ("GEMM", "MARLIN"), ("MARLIN", "GEMM"),
                                      ("GEMV", "GEMM"), ("MARLIN", "GPTQ"), ("GPTQ", "MARLIN")])
def test_pack_mode_conversion_is_exact(src, dst):
    """convert_layer == what the TARGET class's own pack() would have written for the same integers (oracle packers,
    pinned to the reference's pack() by tests/golden), and the converted layer computes the same outputs."""
    import qllm_b200
    K, N, gs = 512, 256, 128
    sym = "MARLIN" in (src, dst)
    rng = np.random.default_rng(11)
    base = O.make_layer("MARLIN" if sym else "GPTQ", 4, gs, K, N, seed=42)          # fixes q, s (and z == 8 when symmetric)
    q, s = base["q"], base["s"]
    z = np.full((K // gs, N), 8, dtype=np.int32) if sym else base["z"]

    def packed(layout):
        if layout == "GPTQ":
            return O.gptq_pack_qweight(q, 4), O.gptq_pack_qzeros(z, 4), s
        if layout == "GEMM":
            return O.awq_pack_qweight(q), O.awq_pack_qzeros(z), s
        if layout == "GEMV":
            return O.awq_gemv_pack(q, z, s, gs)
        qw, sp = O.marlin_pack(q, s, gs)
        return qw, None, sp

    qw, qz, sc = packed(src)
    L = dict(layout=src, bits=4, group_size=gs, K=K, N=N, bias=None, g_idx=O.default_g_idx(K, gs), qweight=qw, qzeros=qz, scales=sc)
    layer = layer_from_dict(L)
    new = qllm_b200.convert_layer(layer, dst)
    assert type(new).__name__ == {"GPTQ": "QuantLinearGPTQ", "GEMM": "WQLinear_GEMM", "GEMV": "WQLinear_GEMV", "MARLIN": "QuantLinearMarlin"}[dst]
    tw, tz, ts = packed(dst)
    assert np.array_equal(new.qweight.cpu().numpy(), tw)
    if tz is not None:
        assert np.array_equal(new.qzeros.cpu().numpy(), tz)
    assert np.array_equal(new.scales.cpu().numpy().view(np.uint16), np.asarray(ts).view(np.uint16))
    x = torch.randn(3, K, dtype=torch.float16, device="cuda")
    ref = O.matmul_ref(x.cpu().numpy(), O.dequant(q, z, s, O.default_g_idx(K, gs), "engine"), None, acc=np.float64)
    assert rel_err(new(x).float().cpu().numpy(), ref) < TOL


def test_marlin_conversion_needs_symmetric_zeros():
    import qllm_b200
    layer = layer_from_dict(O.make_layer("GPTQ", 4, 128, 256, 256, seed=3))
    with pytest.raises(ValueError):
        qllm_b200.convert_layer(layer, "MARLIN")


def test_repack_to_new_mode_on_a_module_tree():
    import qllm_b200

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = layer_from_dict(O.make_layer("GEMM", 4, 128, 256, 128, seed=1))
            self.b = layer_from_dict(O.make_layer("GEMM", 4, 128, 128, 256, seed=2))

        def forward(self, x):
            return self.b(self.a(x))

    m = Block()
    x = torch.randn(2, 256, dtype=torch.float16, device="cuda")
    y0 = m(x)
    qllm_b200.repack_to_new_mode(m, "GPTQ")
    assert isinstance(m.a, qllm_b200.QuantLinearGPTQ) and isinstance(m.b, qllm_b200.QuantLinearGPTQ)
    assert torch.allclose(m(x).float(), y0.float(), rtol=0, atol=2e-3 * y0.float().abs().max().item())


# ---- ORT MatMulNBits blobs (QuantLinearORT) ----------------------------------------------------------
@pytest.mark.parametrize("gs,K,N,act", [(128, 512, 128, False), (32, 256, 96, False), (64, 192, 64, False), (128, 512, 64, True)])
def test_ort_blob_layout(gs, K, N, act):
    L = O.make_layer("ORT", 4, gs, K, N, seed=K + N + gs, bias=True, act_order=act)
    layer = layer_from_dict(L)
    q, z = layer.unpack_int()
    assert np.array_equal(q.cpu().numpy(), L["q"]) and np.array_equal(z.cpu().numpy(), L["z"])
    W = layer.dequantize().cpu().numpy()
    assert np.array_equal(W.view(np.uint16), O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "engine").view(np.uint16))
    for M in (1, 6, 40):
        x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
        y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
        assert rel_err(y, oracle_forward(L, x)) < TOL
    if not act:                                                           # one packed copy; checkpoint bytes restored exactly
        assert layer.qweight.numel() == 0
        sd = layer.state_dict()
        assert np.array_equal(sd["qweight"].cpu().numpy(), L["qweight"]) and np.array_equal(sd["qzeros"].cpu().numpy(), L["qzeros"])
        assert np.array_equal(sd["scales"].cpu().numpy().view(np.uint16), L["scales"].view(np.uint16))


def test_ort_conversion_targets():
    import qllm_b200
    L = O.make_layer("GEMM", 4, 128, 256, 128, seed=5)
    new = qllm_b200.convert_layer(layer_from_dict(L), "ORT")
    qw, qz, sp = O.ort_pack(L["q"], L["z"], L["s"], 128)
    assert np.array_equal(new.qweight.cpu().numpy(), qw) and np.array_equal(new.qzeros.cpu().numpy(), qz)
    assert np.array_equal(new.scales.cpu().numpy().view(np.uint16), sp.view(np.uint16))


# ---- element-wise neighbours fused into the Linear (b200q_linear_ex) ----------------------------------
@pytest.mark.parametrize("layout,act", [("GPTQ", False), ("GEMM", False), ("GPTQ", True), ("HQQ", False)])
@pytest.mark.parametrize("M", [1, 2, 5, 40, 300])
def test_fused_silu_mul_and_residual_match_the_separate_ops(layout, act, M):
    """down_proj(silu(gate) * up) + residual through b200q_linear_ex == the three separate fp16 ops around forward():
    the fused input is rounded like silu-then-mul in fp16, the residual is added to the rounded output."""
    K, N = 768, 256
    L = O.make_layer(layout, 4, 128, K, N, seed=K + M, bias=True, act_order=act, float_zeros=(layout == "HQQ"))
    layer = layer_from_dict(L)
    g = torch.Generator(device="cuda").manual_seed(M)
    gate = torch.randn(M, K, dtype=torch.float16, device="cuda", generator=g) * 2
    up = torch.randn(M, K, dtype=torch.float16, device="cuda", generator=g)
    res = torch.randn(M, N, dtype=torch.float16, device="cuda", generator=g)
    x_ref = torch.nn.functional.silu(gate) * up
    y_plain = layer(x_ref)
    # (1) residual only: bit-identical to the separate add
    assert torch.equal(layer.forward_fused(x_ref, residual=res), y_plain + res)
    # (2) silu * up folded in: the fused input equals torch's fp16 silu / mul up to 1 ulp of fp16 (expf implementations),
    #     so compare against the oracle on the engine-side definition and against the unfused result loosely
    y = layer.forward_fused(gate, x_mul=up, residual=res)
    assert ((y.float() - (y_plain + res).float()).abs().max() / (y_plain + res).float().abs().max()).item() < 2e-3
    xs = x_ref.float().cpu().numpy().astype(np.float16)
    ref = oracle_forward(L, xs) + res.float().cpu().numpy()
    assert rel_err(y.float().cpu().numpy(), ref) < 2e-3


def test_fused_mlp_helper():
    import qllm_b200
    H, I = 256, 512
    gate = layer_from_dict(O.make_layer("GPTQ", 4, 128, H, I, seed=1))
    up = layer_from_dict(O.make_layer("GPTQ", 4, 128, H, I, seed=2))
    down = layer_from_dict(O.make_layer("GPTQ", 4, 128, I, H, seed=3))
    for M in (1, 3, 33):
        x = torch.randn(M, H, dtype=torch.float16, device="cuda")
        want = down(torch.nn.functional.silu(gate(x)) * up(x)) + x
        got = qllm_b200.fused_mlp(gate, up, down, x, residual=x)
        assert ((got.float() - want.float()).abs().max() / want.float().abs().max()).item() < 2e-3


@pytest.mark.parametrize("bits", [3, 5, 6])
def test_act_order_relayout_any_bit_width(bits):
    """desc_act checkpoints at 3 / 5 / 6 bits: the row-permuted re-layout re-packs the 32-row bit-stream blocks exactly."""
    K, N, gs = 256, 64, 64
    L = O.make_layer("GPTQ", bits, gs, K, N, seed=bits, act_order=True)
    layer = layer_from_dict(L)
    for M in (1, 20):
        x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
        y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
        assert rel_err(y, oracle_forward(L, x)) < TOL
    qw, perm = layer._ao
    q = O.gptq_unpack_qweight(qw.cpu().numpy(), bits, K)
    assert np.array_equal(q, L["q"][perm.cpu().numpy()])            # packed row j holds original row perm[j]


# ---- sibling GEMMs at prefill sizes: b200q_linear_group with the flag release (gemm_tcgen05.cu) -------------------
@pytest.mark.gpu
@pytest.mark.parametrize("M", [65, 200, 512])
def test_sibling_group_prefill_matches_single_calls(M):
    """q|k|v through one b200q_linear_group call at M > 64 against three b200q_linear calls (same kernel; a sibling may get
    another tile shape, i.e. another fp32 summation order: 1e-3) and bit-identical from call to call; the flag / counter
    words in the workspace are back at zero afterwards, also after CUDA-graph replays."""
    import qllm_b200
    from qllm_b200 import q_layers
    K = 2048
    layers = [layer_from_dict(O.make_layer("GPTQ", 4, 128, K, N, seed=N)) for N in (1024, 512, 1536)]
    x = torch.randn(M, K, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(M))
    single = [q_layers._B200QuantLinearBase.forward(l, x) for l in layers]
    torch.cuda.synchronize()
    first = None
    for _ in range(3):
        outs = qllm_b200.linear_group(layers, x)
        torch.cuda.synchronize()
        if first is None:
            first = [o.clone() for o in outs]
        for a, b, c in zip(outs, single, first):
            assert torch.equal(a, c)
            assert rel_err(a.float().cpu().numpy(), b.float().cpu().numpy()) < 1e-3
    single = first
    ws = q_layers._workspace(x.device, 4096)
    assert int(ws[:4096].view(torch.int32).abs().sum().item()) == 0
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        qllm_b200.linear_group(layers, x)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            outs = qllm_b200.linear_group(layers, x)
        for _ in range(4):
            for o in outs:
                o.zero_()
            g.replay()
            s.synchronize()
            for a, b in zip(outs, single):
                assert torch.equal(a, b)
        wsg = q_layers._workspace(x.device, 4096)
        assert int(wsg[:4096].view(torch.int32).abs().sum().item()) == 0


@pytest.mark.gpu
def test_sibling_group_prefill_switch_off():
    import qllm_b200
    from qllm_b200 import q_layers
    K, M = 1024, 300
    layers = [layer_from_dict(O.make_layer("GPTQ", 4, 128, K, N, seed=N + 1)) for N in (512, 768)]
    x = torch.randn(M, K, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    on = qllm_b200.linear_group(layers, x)
    qllm_b200.check(qllm_b200.lib.b200q_debug_set_option(b"gemm_siblings", 0.0))
    try:
        off = qllm_b200.linear_group(layers, x)
    finally:
        qllm_b200.check(qllm_b200.lib.b200q_debug_set_option(b"gemm_siblings", 1.0))
    torch.cuda.synchronize()
    for a, b in zip(on, off):
        assert rel_err(a.float().cpu().numpy(), b.float().cpu().numpy()) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("K,N,M", [(4096, 4096, 512), (11008, 4096, 300), (4096, 4096, 129)])
def test_gemm_cluster_split_k_matches_unsplit(K, N, M):
    """The CTA-pair split of K (halves exchanged through distributed shared memory) against the unsplit kernel and the
    float64 arbiter; the pair sums its two fp32 halves in a fixed order, so repeated calls are bit-identical."""
    import qllm_b200
    layer = layer_from_dict(O.make_layer("GPTQ", 4, 128, K, N, seed=K + N + M))
    x = torch.randn(M, K, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    opt = lambda k, v: qllm_b200.check(qllm_b200.lib.b200q_debug_set_option(k, float(v)))
    try:
        opt(b"gemm_force_tt", 128); opt(b"gemm_force_ksplit", 1)
        base = layer(x).clone()
        outs = {}
        for tt in (128, 256):
            opt(b"gemm_force_tt", tt); opt(b"gemm_force_ksplit", 2)
            outs[tt] = layer(x).clone()
            assert torch.equal(outs[tt], layer(x))
    finally:
        opt(b"gemm_force_tt", 0); opt(b"gemm_force_ksplit", 0)
    torch.cuda.synchronize()
    ref = x.double() @ layer.dequantize().double()
    for y in (base, outs[128], outs[256]):
        assert ((y.double() - ref).abs().max() / ref.abs().max()).item() < 1e-3


# ---- bf16 activations through the C ABI (b200q_fusion.act_dtype) ------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("layout,bits,gs,M", [("GPTQ", 4, 128, 1), ("GPTQ", 4, 128, 2), ("GEMM", 4, 128, 1), ("GPTQ", 4, 128, 5),
                                              ("GPTQ", 8, 128, 3), ("GPTQ", 3, 128, 2), ("GPTQ", 4, 128, 100), ("HQQ", 4, 64, 300),
                                              ("MARLIN", 4, 128, 1), ("GPTQ", 4, 128, 64)])
def test_bf16_forward_equals_reference_casts(layout, bits, gs, M):
    """bf16 in / bf16 out == the reference's handling of a bf16 model (auto_cast to fp16, out.to(x.dtype),
    quant_linear_awq.py:29-36,:146), bit for bit, in the decode kernel (native) and through the conversion passes."""
    K, N = 1024, 512
    layer = layer_from_dict(O.make_layer(layout, bits, gs, K, N, seed=bits + M))
    x = (torch.randn(M, K, device="cuda", generator=torch.Generator(device="cuda").manual_seed(M)) * 3).to(torch.bfloat16)
    x[0, 5] = 7.0e4                        # beyond fp16: the reference's cast makes it inf
    x[-1, 9] = 3.0e-6                      # fp16 sub-normal after the cast
    y = layer(x)
    assert y.dtype == torch.bfloat16 and y.shape == (M, N)
    ref = layer(x.to(torch.float16)).to(torch.bfloat16)
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(y.float(), nan=1e30), torch.nan_to_num(ref.float(), nan=1e30))


@pytest.mark.gpu
@pytest.mark.parametrize("M", [1, 2, 6, 40, 128, 600])
def test_bf16_fused_mlp_and_residual(M):
    """b200q_linear_ex in bf16: silu(gate) * up and the residual add rounded as the model's bf16 torch ops round them."""
    import torch.nn.functional as F
    K, N = 1024, 512
    layer = layer_from_dict(O.make_layer("GPTQ", 4, 128, K, N, seed=M + 40))
    g = torch.Generator(device="cuda").manual_seed(M)
    gate = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    up = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    res = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    y = layer.forward_fused(gate, x_mul=up, residual=res)
    assert y.dtype == torch.bfloat16
    ref = res + layer((F.silu(gate) * up).to(torch.float16)).to(torch.bfloat16)
    torch.cuda.synchronize()
    err = (y.float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
    assert err < 2.0 ** -7                 # expf vs torch's exp may move a bf16 rounding by one ulp here and there
    y2 = layer.forward_fused(gate, residual=res)
    ref2 = res + layer(gate.to(torch.float16)).to(torch.bfloat16)
    torch.cuda.synchronize()
    assert torch.equal(y2, ref2)
