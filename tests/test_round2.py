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
