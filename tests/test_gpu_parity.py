"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on seeded inputs.

Bars: integer unpack bit-exact; dequantised fp16 weights bit-exact vs the oracle's "engine"
rounding fp16((q-z)*s); matmul outputs within 1e-3 of max|y| of the float64 oracle (the
tolerance BASELINE.json's north_star states)."""
import numpy as np
import pytest
import torch

from oracle import qlinear_oracle as O
from tests.util import layer_from_dict, oracle_forward, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3

UNPACK_CASES = [("GPTQ", b, 32, 128, 64) for b in range(2, 9)] + [
    ("GPTQ", 4, 128, 256, 96), ("HQQ", 4, 64, 128, 64), ("HQQ", 3, 32, 64, 32), ("GEMM", 4, 64, 128, 128),
    ("MARLIN", 4, 128, 256, 256), ("MARLIN", 4, -1, 128, 256)]


@pytest.mark.parametrize("layout,bits,gs,K,N", UNPACK_CASES)
def test_unpack_bit_exact(layout, bits, gs, K, N):
    L = O.make_layer(layout, bits, gs, K, N, seed=K + N + bits, act_order=(layout == "GPTQ" and bits == 4))
    layer = layer_from_dict(L)
    q, z = layer.unpack_int()
    assert np.array_equal(q.cpu().numpy(), L["q"])
    if z is not None:
        assert np.array_equal(z.cpu().numpy(), np.asarray(L["z"]).astype(np.int32))
    W = layer.dequantize().cpu().numpy()
    ref = O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "engine")
    assert np.array_equal(W.view(np.uint16), ref.view(np.uint16))


def test_unpack_autogptq_zero_bias():
    L = O.make_layer("GPTQ", 4, 64, 128, 64, seed=5)
    L2 = dict(L)
    L2["qzeros"] = O.gptq_pack_qzeros(L["z"], 4, zero_bias=1)      # AutoGPTQ stores z-1
    layer = layer_from_dict(L2)
    layer.zero_bias = 1
    _, z = layer.unpack_int()
    assert np.array_equal(z.cpu().numpy(), L["z"])
    layer2 = layer_from_dict(L2)
    layer2.handle_qzeros_for_autogptq()                              # loader-style rewrite
    assert np.array_equal(layer2.qzeros.cpu().numpy(), L["qzeros"])


# (layout, bits, group, K, N): shapes chosen to hit partial tiles, several k-splits, several groups
GEMV_CASES = [
    ("GPTQ", 4, 128, 1024, 512), ("GPTQ", 4, 32, 512, 96), ("GPTQ", 4, -1, 512, 64), ("GPTQ", 4, 64, 2048, 160),
    ("GPTQ", 2, 64, 1024, 128), ("GPTQ", 2, 16, 512, 64), ("GPTQ", 8, 128, 512, 128), ("GPTQ", 8, 32, 256, 96),
    ("HQQ", 4, 64, 1024, 256), ("HQQ", 2, 64, 512, 64), ("HQQ", 8, 128, 512, 64),
    ("GEMM", 4, 128, 1024, 1024), ("GEMM", 4, 32, 512, 576), ("GEMM", 4, 64, 2048, 64),
    ("MARLIN", 4, 128, 1024, 512), ("MARLIN", 4, -1, 512, 256), ("MARLIN", 4, 128, 2048, 768),
]


@pytest.mark.parametrize("layout,bits,gs,K,N", GEMV_CASES)
@pytest.mark.parametrize("M", [1, 3, 8])
def test_decode_kernel_vs_oracle(layout, bits, gs, K, N, M):
    import qllm_b200
    L = O.make_layer(layout, bits, gs, K, N, seed=K * 7 + N + M, bias=(M == 3),
                     float_zeros=(layout == "HQQ" and bits == 4))
    layer = layer_from_dict(L)
    x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
    y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
    ref = oracle_forward(L, x)
    assert rel_err(y, ref) < TOL
    # these shapes must take the mma decode kernel, not the generic fallback
    import ctypes
    assert qllm_b200.lib.b200q_select_kernel(ctypes.byref(layer._descriptor()), M) == 1


GENERIC_CASES = [("GPTQ", 3, 32, 256, 64, False), ("GPTQ", 5, 64, 128, 64, False), ("GPTQ", 6, 32, 128, 32, False),
                 ("GPTQ", 7, 32, 128, 32, False), ("GPTQ", 4, 32, 256, 64, True), ("GPTQ", 3, 32, 128, 32, True),
                 ("HQQ", 3, 64, 256, 64, False), ("GPTQ", 8, 64, 128, 64, True)]


@pytest.mark.parametrize("layout,bits,gs,K,N,act", GENERIC_CASES)
@pytest.mark.parametrize("M", [1, 7, 16, 40])
def test_generic_kernel_vs_oracle(layout, bits, gs, K, N, act, M):
    L = O.make_layer(layout, bits, gs, K, N, seed=K + N + M + bits, act_order=act, bias=True)
    layer = layer_from_dict(L)
    x = np.random.default_rng(M + 1).standard_normal((M, K)).astype(np.float16)
    y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
    assert rel_err(y, oracle_forward(L, x)) < TOL


def test_workspace_left_zeroed_and_reusable():
    from qllm_b200 import q_layers
    L = O.make_layer("GEMM", 4, 128, 2048, 512, seed=1)
    layer = layer_from_dict(L)
    x = torch.randn(2, 2048, dtype=torch.float16, device="cuda")
    y1 = layer(x)
    y2 = layer(x)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)                       # deterministic split-K reduction
    for ws in q_layers._workspaces.values():
        assert int(ws[:4096].count_nonzero()) == 0   # arrival counters are reset by the last CTA


def test_strided_and_batched_inputs():
    L = O.make_layer("GPTQ", 4, 128, 512, 256, seed=2)
    layer = layer_from_dict(L)
    xb = torch.randn(2, 3, 1024, dtype=torch.float16, device="cuda")
    x = xb[..., :512]                                  # row stride 1024, not contiguous
    y = layer(x)
    assert y.shape == (2, 3, 256)
    ref = oracle_forward(L, x.reshape(-1, 512).cpu().numpy())
    assert rel_err(y.reshape(-1, 256).float().cpu().numpy(), ref) < TOL
    xbf = x.to(torch.bfloat16)                         # bf16 activations: cast to fp16 and back
    ybf = layer(xbf)
    assert ybf.dtype == torch.bfloat16


def test_error_codes():
    import ctypes
    import qllm_b200
    from qllm_b200._lib import Layer
    lib = qllm_b200.lib
    d = Layer()
    assert lib.b200q_workspace_bytes(ctypes.byref(d), 1) == 0
    L = O.make_layer("GEMM", 4, 128, 256, 64, seed=3)
    layer = layer_from_dict(L)
    desc = layer._descriptor()
    x = torch.zeros(1, 256, dtype=torch.float16, device="cuda")
    y = torch.zeros(1, 64, dtype=torch.float16, device="cuda")
    st = lib.b200q_linear(ctypes.byref(desc), x.data_ptr(), 1, 256, y.data_ptr(), 64, None, 0, None)
    assert st == -5 and b"workspace" in lib.b200q_strerror(st)
    st = lib.b200q_linear(ctypes.byref(desc), None, 1, 256, y.data_ptr(), 64, None, 0, None)
    assert st == -1
    with pytest.raises(RuntimeError):
        layer.cpu()(torch.zeros(1, 256, dtype=torch.float16))


@pytest.mark.parametrize("layout,K,N", [("GEMM", 4096, 4096), ("GPTQ", 4096, 11008), ("GPTQ", 11008, 4096),
                                         ("MARLIN", 4096, 4096)])
def test_full_size_properties(layout, K, N):
    """BASELINE sizes, where the numpy oracle is too slow: size-independent properties.
    (1) one-hot activations read back single rows of W bit-exactly (unpack at full size);
    (2) linearity: f(a*x1 + x2) == a*f(x1) + f(x2) within fp16 rounding;
    (3) agreement with an fp32 torch matmul over the engine's own dequantised weights."""
    gs = 128
    rng = np.random.default_rng(K + N)
    L = dict(layout=layout, bits=4, group_size=gs, K=K, N=N, bias=None, g_idx=O.default_g_idx(K, gs))
    if layout == "MARLIN":
        L["qweight"] = rng.integers(-2**31, 2**31, size=(K // 16, 2 * N), dtype=np.int64).astype(np.int32)
        L["qzeros"] = None
    elif layout == "GEMM":
        L["qweight"] = rng.integers(-2**31, 2**31, size=(K, N // 8), dtype=np.int64).astype(np.int32)
        L["qzeros"] = rng.integers(-2**31, 2**31, size=(K // gs, N // 8), dtype=np.int64).astype(np.int32)
    else:
        L["qweight"] = rng.integers(-2**31, 2**31, size=(K // 8, N), dtype=np.int64).astype(np.int32)
        L["qzeros"] = rng.integers(-2**31, 2**31, size=(K // gs, N // 8), dtype=np.int64).astype(np.int32)
    L["scales"] = rng.uniform(0.002, 0.012, size=(K // gs, N)).astype(np.float16)
    layer = layer_from_dict(L)
    W = layer.dequantize()
    rows = [0, 1, 7, gs, K // 2 + 3, K - 1]
    x = torch.zeros(len(rows), K, dtype=torch.float16, device="cuda")
    for i, r in enumerate(rows):
        x[i, r] = 1.0
    y = layer(x)
    assert torch.equal(y, W[rows])
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn(1, K, dtype=torch.float16, device="cuda", generator=g)
    x2 = torch.randn(1, K, dtype=torch.float16, device="cuda", generator=g)
    f1, f2, f3 = layer(x1).float(), layer(x2).float(), layer(2 * x1 + x2).float()
    assert (f3 - (2 * f1 + f2)).abs().max() <= 4e-3 * f3.abs().max()
    xm = torch.randn(8, K, dtype=torch.float16, device="cuda", generator=g)
    ref = xm.double() @ W.double()
    assert ((layer(xm).double() - ref).abs().max() / ref.abs().max()).item() < TOL
    assert ((layer(xm[:1]).double() - ref[:1]).abs().max() / ref.abs().max()).item() < TOL


# ---- tcgen05 GEMM (M > 8, or forced) ---------------------------------------------------------
GEMM_CASES = [("GPTQ", 4, 128, 512, 256), ("GPTQ", 4, 64, 1024, 384), ("GPTQ", 4, 32, 256, 128),
              ("HQQ", 4, 64, 512, 128), ("GPTQ", 4, -1, 256, 128), ("GPTQ", 4, 128, 2048, 160)]


def _dump(name, **arrs):
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed(os.path.join("gpurun_out", name), **arrs)


@pytest.mark.parametrize("layout,bits,gs,K,N", GEMM_CASES)
@pytest.mark.parametrize("M", [9, 33, 64, 100, 128, 257])
def test_tcgen05_gemm_vs_oracle(layout, bits, gs, K, N, M):
    import ctypes
    import qllm_b200
    L = O.make_layer(layout, bits, gs, K, N, seed=K + 3 * N + M, bias=(M % 2 == 1),
                     float_zeros=(layout == "HQQ"))
    layer = layer_from_dict(L)
    assert qllm_b200.lib.b200q_select_kernel(ctypes.byref(layer._descriptor()), M) == 2
    x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
    y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
    ref = oracle_forward(L, x)
    err = rel_err(y, ref)
    if not err < TOL:
        _dump(f"gemm_fail_{layout}_{gs}_{K}_{N}_{M}.npz", y=y, ref=ref, x=x, q=L["q"], z=np.asarray(L["z"]), s=L["s"])
    assert err < TOL


def test_tcgen05_gemm_forced_small_m_and_identity():
    """b200q_gemm at M < 8 (zero-filled token tile) and a one-hot X that reads W back bit-exactly:
    pins the k order inside the TMEM A operand and the n order of the epilogue."""
    import ctypes
    import qllm_b200
    from qllm_b200 import q_layers
    K, N = 256, 256
    L = O.make_layer("GPTQ", 4, 128, K, N, seed=77)
    layer = layer_from_dict(L)
    desc = layer._descriptor()
    W = O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "engine")
    for M in (1, 5, K):
        x = torch.zeros(M, K, dtype=torch.float16, device="cuda")
        rows = list(range(M)) if M == K else [3, 77, 128, 200, 255][:M]
        for i, r in enumerate(rows):
            x[i, r] = 1.0
        y = torch.empty(M, N, dtype=torch.float16, device="cuda")
        ws = q_layers._workspace(x.device, 1 << 20)
        st = qllm_b200.lib.b200q_gemm(ctypes.byref(desc), x.data_ptr(), M, K, y.data_ptr(), N, ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream().cuda_stream)
        qllm_b200.check(st)
        torch.cuda.synchronize()
        got = y.cpu().numpy()
        if not np.array_equal(got.view(np.uint16), W[rows].view(np.uint16)):
            _dump(f"gemm_identity_fail_M{M}.npz", y=got, W=W, rows=np.array(rows))
        assert np.array_equal(got.view(np.uint16), W[rows].view(np.uint16))


@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 11008)])
def test_tcgen05_gemm_full_size(K, N):
    gs = 128
    rng = np.random.default_rng(K + N)
    L = dict(layout="GPTQ", bits=4, group_size=gs, K=K, N=N, bias=None, g_idx=O.default_g_idx(K, gs))
    L["qweight"] = rng.integers(-2**31, 2**31, size=(K // 8, N), dtype=np.int64).astype(np.int32)
    L["qzeros"] = rng.integers(-2**31, 2**31, size=(K // gs, N // 8), dtype=np.int64).astype(np.int32)
    L["scales"] = rng.uniform(0.002, 0.012, size=(K // gs, N)).astype(np.float16)
    layer = layer_from_dict(L)
    W = layer.dequantize()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(512, K, dtype=torch.float16, device="cuda", generator=g)
    y = layer(x)
    ref = x.double() @ W.double()
    assert ((y.double() - ref).abs().max() / ref.abs().max()).item() < TOL


@pytest.mark.parametrize("layout,gs,K,N", [("GEMM", 128, 512, 256), ("GEMM", 64, 1024, 384), ("MARLIN", 128, 512, 256),
                                            ("MARLIN", -1, 256, 256)])
@pytest.mark.parametrize("M", [17, 128, 300])
def test_awq_marlin_prefill_through_exact_repack(layout, gs, K, N, M):
    """M > 8 on AWQ/Marlin layers: one-time integer re-layout (bit-exact) + tcgen05 GEMM."""
    L = O.make_layer(layout, 4, gs, K, N, seed=K + N + M, bias=True)
    layer = layer_from_dict(L)
    x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
    y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
    assert rel_err(y, oracle_forward(L, x)) < TOL
    qw, qz, sc = layer._shadow
    q = O.gptq_unpack_qweight(qw.cpu().numpy(), 4, K)
    z = O.gptq_unpack_qzeros(qz.cpu().numpy(), 4, N)
    assert np.array_equal(q, L["q"]) and np.array_equal(z, np.asarray(L["z"]).astype(np.int32))
    assert np.array_equal(sc.cpu().numpy().view(np.uint16), L["s"].view(np.uint16))


@pytest.mark.parametrize("layout,bits,gs,K,N", [("GPTQ", 8, 128, 512, 256), ("GPTQ", 2, 64, 512, 128), ("GPTQ", 2, 16, 256, 128),
                                                ("HQQ", 8, 128, 256, 128), ("HQQ", 2, 64, 512, 256), ("GPTQ", 8, 32, 256, 160)])
@pytest.mark.parametrize("M", [20, 128, 260])
def test_tcgen05_gemm_2bit_8bit(layout, bits, gs, K, N, M):
    import ctypes
    import qllm_b200
    L = O.make_layer(layout, bits, gs, K, N, seed=K + N + M + bits, bias=True, float_zeros=(layout == "HQQ" and bits == 8))
    layer = layer_from_dict(L)
    assert qllm_b200.lib.b200q_select_kernel(ctypes.byref(layer._descriptor()), M) == 2
    x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
    y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
    assert rel_err(y, oracle_forward(L, x)) < TOL


# ---- streaming decode kernel: sibling groups, plan variations, legacy path ------------------------
GROUP_CASES = [("GEMM", 4, 128, 1024, (512, 256, 256)), ("GEMM", 4, 64, 512, (1024, 1024)), ("GPTQ", 4, 128, 1024, (256, 96, 160)),
               ("GPTQ", 2, 64, 512, (128, 64)), ("GPTQ", 8, 128, 512, (128, 128, 64)), ("HQQ", 4, 64, 1024, (256, 128)),
               ("MARLIN", 4, 128, 1024, (512, 256, 256)), ("GEMM", 4, 128, 4096, (4096, 4096, 4096))]


@pytest.mark.parametrize("layout,bits,gs,K,Ns", GROUP_CASES)
@pytest.mark.parametrize("M", [1, 2, 8])
def test_decode_sibling_group_vs_oracle(layout, bits, gs, K, Ns, M):
    """b200q_linear_group (q/k/v-, gate/up-style siblings in one launch) == the per-layer oracle results."""
    import qllm_b200
    if K * max(Ns) > 2 ** 22 and M != 1:
        pytest.skip("full-size group: M=1 only (oracle time)")
    Ls = [O.make_layer(layout, bits, gs, K, N, seed=K + 13 * N + i, bias=(i == 1), float_zeros=(layout == "HQQ"))
          for i, N in enumerate(Ns)]
    layers = [layer_from_dict(L) for L in Ls]
    x = np.random.default_rng(M + K).standard_normal((M, K)).astype(np.float16)
    for l in layers:
        l._decode_descriptor(M)                                  # one-time re-layout (AWQ / Marlin) is not part of the call
    n0 = qllm_b200.lib.b200q_launch_count()
    ys = qllm_b200.linear_group(layers, torch.from_numpy(x).cuda())
    assert qllm_b200.lib.b200q_launch_count() - n0 == 1          # fused: one launch for the whole group
    for L, y in zip(Ls, ys):
        assert rel_err(y.float().cpu().numpy(), oracle_forward(L, x)) < TOL
    # and the same as the one-layer-at-a-time path up to the summation order of the K split
    for layer, y in zip(layers, ys):
        y1 = layer(torch.from_numpy(x).cuda()).float()
        assert (y1 - y.float()).abs().max() <= 1e-3 * y1.abs().max()


def test_decode_group_mixed_falls_back_per_layer():
    import qllm_b200
    La = O.make_layer("GEMM", 4, 128, 512, 256, seed=1)
    Lb = O.make_layer("GEMM", 4, 64, 512, 256, seed=2)               # different group size: not fusable
    layers = [layer_from_dict(La), layer_from_dict(Lb)]
    x = np.random.default_rng(0).standard_normal((1, 512)).astype(np.float16)
    for l in layers:
        l._decode_descriptor(1)
    n0 = qllm_b200.lib.b200q_launch_count()
    ys = qllm_b200.linear_group(layers, torch.from_numpy(x).cuda())
    assert qllm_b200.lib.b200q_launch_count() - n0 == 2
    for L, y in zip((La, Lb), ys):
        assert rel_err(y.float().cpu().numpy(), oracle_forward(L, x)) < TOL


@pytest.mark.parametrize("opt,val", [("st_cluster", 1), ("st_cluster", 3), ("st_cluster", 8), ("st_depth", 2), ("st_depth", 16),
                                     ("st_tpc", 2), ("st_tpc", 4), ("st_target", 400)])
@pytest.mark.parametrize("layout,K,N", [("GEMM", 2048, 640), ("GPTQ", 2048, 416), ("MARLIN", 2048, 320)])
def test_stream_kernel_plan_variations(opt, val, layout, K, N):
    """Every (cluster, ring depth, tiles-per-CTA) the planner can pick gives the same answer."""
    import qllm_b200
    lib = qllm_b200.lib
    if layout == "MARLIN":
        N = 512                                                      # the Marlin ctor wants N % 256 == 0
    L = O.make_layer(layout, 4, 128, K, N, seed=K + N)
    layer = layer_from_dict(L)
    x = np.random.default_rng(3).standard_normal((3, K)).astype(np.float16)
    xt = torch.from_numpy(x).cuda()
    base = layer(xt)
    assert lib.b200q_debug_set_option(opt.encode(), float(val)) == 0
    try:
        y = layer(xt)
        y1 = layer(xt[:1])
    finally:
        lib.b200q_debug_set_option(opt.encode(), 0.0 if opt != "st_target" else 120.0)
    ref = oracle_forward(L, x)
    assert rel_err(y.float().cpu().numpy(), ref) < TOL
    assert rel_err(y1.float().cpu().numpy(), ref[:1]) < TOL
    assert rel_err(base.float().cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("layout,bits,gs,K,N", [("GEMM", 4, 128, 1024, 1024), ("GPTQ", 4, 128, 1024, 512), ("GPTQ", 2, 64, 1024, 128),
                                                ("MARLIN", 4, 128, 1024, 512), ("GEMM", 4, 128, 11008, 4096)])
def test_pre_streaming_decode_kernels_still_agree(layout, bits, gs, K, N):
    """The one remaining non-streaming decode kernel (gemv_mma.cu: cp.async.bulk + mbarrier pipeline, the fallback for the
    group sizes the streaming kernels do not tile) on shapes the streaming kernels also cover: same oracle, same bar."""
    import qllm_b200
    lib = qllm_b200.lib
    L = O.make_layer(layout, bits, gs, K, N, seed=K + N + 1) if K * N < 2 ** 22 else None
    if L is None:
        rng = np.random.default_rng(1)
        L = dict(layout=layout, bits=4, group_size=gs, K=K, N=N, bias=None, g_idx=O.default_g_idx(K, gs))
        L["qweight"] = rng.integers(-2**31, 2**31, size=(K, N // 8), dtype=np.int64).astype(np.int32)
        L["qzeros"] = rng.integers(-2**31, 2**31, size=(K // gs, N // 8), dtype=np.int64).astype(np.int32)
        L["scales"] = rng.uniform(0.002, 0.012, size=(K // gs, N)).astype(np.float16)
    layer = layer_from_dict(L)
    x = torch.randn(2, K, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    y_stream = layer(x)
    assert lib.b200q_debug_set_option(b"stream", 0.0) == 0
    try:
        y_old = layer(x)
    finally:
        lib.b200q_debug_set_option(b"stream", 1.0)
    W = layer.dequantize()
    ref = x.double() @ W.double()
    for y in (y_stream, y_old):
        assert ((y.double() - ref).abs().max() / ref.abs().max()).item() < TOL


def test_fuse_siblings_is_transparent_to_the_caller():
    """HF-style attention / MLP modules calling q_proj(x), k_proj(x), v_proj(x) one after the other get the fused
    launch without any change to their forward code, and any other calling pattern still gives the right answer."""
    import qllm_b200

    class Attn(torch.nn.Module):
        def __init__(self, mods):
            super().__init__()
            self.q_proj, self.k_proj, self.v_proj, self.o_proj = mods

        def forward(self, h):
            q, k, v = self.q_proj(h), self.k_proj(h), self.v_proj(h)
            return self.o_proj(q + k + v)

    K = 512
    Ls = [O.make_layer("GEMM", 4, 128, K, K, seed=40 + i) for i in range(4)]
    attn = Attn([layer_from_dict(L) for L in Ls])
    h = torch.randn(1, 1, K, dtype=torch.float16, device="cuda")
    plain = attn(h)
    n0 = qllm_b200.lib.b200q_launch_count()
    assert qllm_b200.fuse_siblings(attn) == 1
    fused = attn(h)
    assert qllm_b200.lib.b200q_launch_count() - n0 == 2              # qkv in one launch + o_proj
    assert (plain.float() - fused.float()).abs().max() <= 2e-3 * plain.float().abs().max()
    # out-of-order / different-input calls bypass the parked results
    h2 = torch.randn(1, 1, K, dtype=torch.float16, device="cuda")
    k_only = attn.k_proj(h2)
    assert torch.equal(k_only, qllm_b200.q_layers._B200QuantLinearBase.forward(attn.k_proj, h2))
    q1 = attn.q_proj(h)
    v2 = attn.v_proj(h2)                                             # parked v belongs to h, not h2
    assert torch.equal(v2, qllm_b200.q_layers._B200QuantLinearBase.forward(attn.v_proj, h2))
    h.add_(1.0)                                                      # in-place update bumps the version counter
    k3 = attn.k_proj(h)
    assert torch.equal(k3, qllm_b200.q_layers._B200QuantLinearBase.forward(attn.k_proj, h))


# ---- integer-tensor-path decode kernel (gemv_imma.cu): plan variations, M = 2, partial tiles, fallbacks ----
@pytest.mark.parametrize("opt,val", [("im_cluster", 1), ("im_cluster", 3), ("im_cluster", 8), ("im_depth", 2), ("im_depth", 4), ("im_tpc", 2),
                                     ("im_target", 40), ("im_target", 290)])
@pytest.mark.parametrize("layout,gs,K,N", [("GPTQ", 128, 2048, 416), ("GPTQ", 32, 1024, 160), ("HQQ", 64, 2048, 256), ("GEMM", 128, 2048, 640),
                                           ("MARLIN", 128, 2048, 512), ("GPTQ", -1, 1024, 96)])
def test_imma_kernel_plan_variations(opt, val, layout, gs, K, N):
    import qllm_b200
    lib = qllm_b200.lib
    L = O.make_layer(layout, 4, gs, K, N, seed=K + N + 7, bias=True, float_zeros=(layout == "HQQ"))
    layer = layer_from_dict(L)
    x = np.random.default_rng(5).standard_normal((2, K)).astype(np.float16)
    x[0, 5] = 37.0                                                   # an outlier: the digit split is scaled per 128-k part
    x[1, K // 2] = -0.001
    xt = torch.from_numpy(x).cuda()
    assert lib.b200q_debug_set_option(opt.encode(), float(val)) == 0
    try:
        y2 = layer(xt)
        y1 = layer(xt[:1])
    finally:
        lib.b200q_debug_set_option(opt.encode(), 148.0 if opt == "im_target" else 0.0)
    ref = oracle_forward(L, x)
    assert rel_err(y2.float().cpu().numpy(), ref) < TOL
    assert rel_err(y1.float().cpu().numpy(), ref[:1]) < TOL


def test_imma_kernel_matches_fp16_path_and_handles_extremes():
    """Same layer through the integer path and (imma = 0) the fp16-sub-normal path; activations spanning fp16's range."""
    import qllm_b200
    lib = qllm_b200.lib
    K, N = 4096, 1024
    L = O.make_layer("GPTQ", 4, 128, K, N, seed=99)
    layer = layer_from_dict(L)
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((1, K)) * np.exp(rng.uniform(-9, 5, size=(1, K)))).astype(np.float16)   # 6 decades of magnitude
    x[0, :128] = 0.0                                                 # an all-zero part
    xt = torch.from_numpy(x).cuda()
    y_int = layer(xt).float().cpu().numpy()
    lib.b200q_debug_set_option(b"imma", 0.0)
    try:
        y_f16 = layer(xt).float().cpu().numpy()
    finally:
        lib.b200q_debug_set_option(b"imma", 1.0)
    ref = oracle_forward(L, x)
    assert rel_err(y_int, ref) < TOL and rel_err(y_f16, ref) < TOL
    assert rel_err(y_int, ref) <= rel_err(y_f16, ref) + 2e-4        # the fixed-point split loses nothing against fp16 MMA inputs


# ---------------------------------------------------------------------------------------------------------------
# act-order (desc_act) checkpoints: exact row-permuted re-layout + activations gathered through x_perm

@pytest.mark.gpu
@pytest.mark.parametrize("bits,gs,K,N", [(4, 128, 1024, 512), (4, 64, 512, 256), (8, 128, 512, 256), (2, 64, 512, 256)])
@pytest.mark.parametrize("M", [1, 2, 5, 16, 64, 300])
def test_act_order_relayout_vs_oracle(bits, gs, K, N, M):
    import ctypes
    import qllm_b200
    L = O.make_layer("GPTQ", bits, gs, K, N, seed=7 * K + N + bits, act_order=True, bias=(M == 2))
    layer = layer_from_dict(L)
    x = np.random.default_rng(M + 3).standard_normal((M, K)).astype(np.float16)
    y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
    assert rel_err(y, oracle_forward(L, x)) < TOL
    fast = layer._fast_descriptor()
    assert fast.x_perm and not fast.g_idx                       # the re-layout is in use, not the generic g_idx kernel
    want = 1 if M <= 8 else 2
    assert qllm_b200.lib.b200q_select_kernel(ctypes.byref(fast), M) == want


@pytest.mark.gpu
def test_act_order_repack_is_the_exact_row_permutation():
    K, N, gs = 512, 256, 128
    L = O.make_layer("GPTQ", 4, gs, K, N, seed=99, act_order=True)
    layer = layer_from_dict(L)
    layer._fast_descriptor()
    qw_perm, perm = layer._ao
    perm = perm.cpu().numpy()
    assert np.array_equal(L["g_idx"][perm], np.arange(K) // gs)               # groups contiguous after the permutation
    q_perm = O.unpack_rows(qw_perm.cpu().numpy(), 4, K) if hasattr(O, "unpack_rows") else None
    if q_perm is None:
        from qllm_b200 import codec
        q_perm = codec.unpack_rows(qw_perm.cpu(), 4, K).numpy()
    assert np.array_equal(q_perm, L["q"][perm])                                # bit-exact
    # one-hot activations read single rows of W back through the gather: row k of W = y(e_k)
    for k in (0, 17, K - 1):
        e = torch.zeros(1, K, dtype=torch.float16, device="cuda")
        e[0, k] = 1.0
        W = O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "engine")
        assert np.array_equal(layer(e).cpu().numpy()[0], W[k].astype(np.float16))


@pytest.mark.gpu
@pytest.mark.parametrize("M", [1, 12])
def test_irregular_g_idx_stays_on_the_generic_kernel(M):
    """g_idx maps whose groups do not all own group_size rows cannot be made contiguous: generic kernel, same answer."""
    K, N, gs = 256, 64, 32
    L = O.make_layer("GPTQ", 4, gs, K, N, seed=5, act_order=True)
    g = L["g_idx"].copy()
    g[g == 1] = 0                                                              # group 0 now owns 64 rows, group 1 none
    L = O.make_layer("GPTQ", 4, gs, K, N, seed=5, act_order=True)
    L["g_idx"] = g.astype(np.int32)
    layer = layer_from_dict(L)
    x = np.random.default_rng(1).standard_normal((M, K)).astype(np.float16)
    y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
    assert rel_err(y, oracle_forward(L, x)) < TOL
    assert not layer._fast_descriptor().x_perm


@pytest.mark.gpu
def test_generic_g_idx_kernel_still_agrees(monkeypatch):
    from qllm_b200 import q_layers
    monkeypatch.setattr(q_layers, "ACTORDER_RELAYOUT", False)
    L = O.make_layer("GPTQ", 4, 128, 512, 256, seed=3, act_order=True)
    layer = layer_from_dict(L)
    for M in (1, 9):
        x = np.random.default_rng(M).standard_normal((M, 512)).astype(np.float16)
        assert rel_err(layer(torch.from_numpy(x).cuda()).float().cpu().numpy(), oracle_forward(L, x)) < TOL
    assert not layer._fast_descriptor().x_perm


@pytest.mark.gpu
@pytest.mark.parametrize("layout,bits,gs,K,N", [("HQQ", 4, 64, 14336, 256), ("GPTQ", 4, 32, 8192, 384), ("HQQ", 2, 64, 14336, 128),
                                                ("GPTQ", 8, 64, 14336, 128)])
@pytest.mark.parametrize("M", [16, 200])
def test_tcgen05_gemm_long_k_small_groups_chunked_tables(layout, bits, gs, K, N, M):
    """More group constants than fit in shared memory (Mixtral w2: K = 14336, g64): the GEMM stages its (scale, zero)
    tables per chunk of k-blocks instead of falling back to the generic kernel."""
    import ctypes
    import qllm_b200
    L = O.make_layer(layout, bits, gs, K, N, seed=K + N + bits + M, float_zeros=(layout == "HQQ"))
    layer = layer_from_dict(L)
    assert qllm_b200.lib.b200q_select_kernel(ctypes.byref(layer._descriptor()), M) == 2
    x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
    y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
    assert rel_err(y, oracle_forward(L, x)) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("layout,bits,gs,K,N", [("GPTQ", 4, 128, 4096, 512), ("GPTQ", 4, 128, 11008, 512), ("HQQ", 4, 64, 8192, 256),
                                                ("GPTQ", 8, 128, 4096, 384), ("MARLIN", 4, 128, 4096, 512)])
@pytest.mark.parametrize("M", [9, 33, 64])
def test_tcgen05_gemm_split_k_small_m(layout, bits, gs, K, N, M):
    """M <= 64: K is split over gridDim.z so that every SM dequantises (partial fp32 tiles, last-arriver reduction in
    split order): same answer as the oracle, bit-identical run to run, arrival counters left zeroed, and equal (to
    fp32 summation order) to the unsplit kernel."""
    import qllm_b200
    from qllm_b200 import q_layers
    L = O.make_layer(layout, bits, gs, K, N, seed=K + N + M, bias=(M == 33), float_zeros=(layout == "HQQ"))
    layer = layer_from_dict(L)
    x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
    xd = torch.from_numpy(x).cuda()
    y1 = layer(xd)
    y2 = layer(xd)
    torch.cuda.synchronize()
    ref = oracle_forward(L, x)
    assert rel_err(y1.float().cpu().numpy(), ref) < TOL
    assert torch.equal(y1, y2)
    for ws in q_layers._workspaces.values():
        assert int(ws[:4096].count_nonzero()) == 0
    qllm_b200.lib.b200q_debug_set_option(b"gemm_splitk", 0.0)
    try:
        y0 = layer(xd)
        torch.cuda.synchronize()
    finally:
        qllm_b200.lib.b200q_debug_set_option(b"gemm_splitk", 1.0)
    assert rel_err(y0.float().cpu().numpy(), ref) < TOL
    assert (y0.float() - y1.float()).abs().max().item() <= 2e-3 * float(np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("layout,gs,K,N", [("GPTQ", 64, 1024, 256), ("GPTQ", 128, 4096, 384), ("HQQ", 64, 2048, 256), ("GPTQ", 32, 512, 160)])
@pytest.mark.parametrize("M", [1, 5, 16, 200])
def test_three_bit_layers_run_on_the_tensor_core_gemm(layout, gs, K, N, M):
    """3-bit (32 values straddling three words) has a tcgen05 producer; with no 3-bit decode kernel, M <= 8 takes the
    split-K GEMM as well instead of the per-element generic kernel."""
    import ctypes
    import qllm_b200
    L = O.make_layer(layout, 3, gs, K, N, seed=K + N + M, bias=(M == 5), float_zeros=(layout == "HQQ"))
    layer = layer_from_dict(L)
    assert qllm_b200.lib.b200q_select_kernel(ctypes.byref(layer._descriptor()), M) == 2
    x = np.random.default_rng(M).standard_normal((M, K)).astype(np.float16)
    y = layer(torch.from_numpy(x).cuda()).float().cpu().numpy()
    assert rel_err(y, oracle_forward(L, x)) < TOL
    if M == 1:                                    # one-hot rows read W back bit-exactly through the 3-bit unpack
        W = O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "engine").astype(np.float16)
        for k in (0, 31, 32, K - 1):
            e = torch.zeros(1, K, dtype=torch.float16, device="cuda")
            e[0, k] = 1.0
            got = layer(e).cpu().numpy()[0]
            want = W[k] if L["bias"] is None else (W[k].astype(np.float32) + L["bias"].astype(np.float32)).astype(np.float16)
            assert np.array_equal(got, want)
