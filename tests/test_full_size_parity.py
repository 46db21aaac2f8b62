"""Full-size parity (BASELINE.json configs 2-5 shapes): the engine against the numpy oracle and against the
reference's own CUDA kernels (oracle/_ref) on identical packed buffers.

The oracle half under test is the unpack + dequant (numpy, from the PACKED buffers, nothing of the engine);
the contraction x @ W over the oracle's weights is done in float64 -- numpy for <= 64 rows, torch-on-GPU float64 for
the 512 / 2048-row cases (a float64 matmul is the arbiter, not a product path).  Tolerance: 1e-3 of max|y|
(north_star)."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

from oracle import qlinear_oracle as O
from tests.util import layer_from_dict

pytestmark = pytest.mark.gpu
TOL = 1e-3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
HAVE_REF = bool(glob.glob(os.path.join(REF, "awq_inference_engine*.so")))


def _rand_words(rng, *shape):
    return rng.integers(-2 ** 31, 2 ** 31, size=shape, dtype=np.int64).astype(np.int32)


def _packed_layer(layout, bits, gs, K, N, seed, act_order=False, float_zeros=False):
    """Random PACKED buffers (what a checkpoint holds) + the oracle's view of them."""
    rng = np.random.default_rng(seed)
    G = K // gs
    L = dict(layout=layout, bits=bits, group_size=gs, K=K, N=N, bias=None, g_idx=O.default_g_idx(K, gs))
    if layout == "GEMM":
        L["qweight"], L["qzeros"] = _rand_words(rng, K, N // 8), _rand_words(rng, G, N // 8)
    elif layout == "HQQ":
        L["qweight"] = _rand_words(rng, K * bits // 32, N)
        z = rng.integers(0, 1 << bits, size=(G, N)).astype(np.float16)
        if float_zeros:
            z = (z + rng.uniform(-0.5, 0.5, size=z.shape)).astype(np.float16)
        L["qzeros"] = z
    else:
        L["qweight"], L["qzeros"] = _rand_words(rng, K * bits // 32, N), _rand_words(rng, G, N * bits // 32)
        if act_order:
            L["g_idx"] = O.default_g_idx(K, gs)[rng.permutation(K)].astype(np.int32)
    L["scales"] = rng.uniform(0.002, 0.012, size=(G, N)).astype(np.float16)
    q, z, s, gi = O.unpack_layer(layout, bits, gs, K, N, L["qweight"], L["qzeros"], L["scales"], L["g_idx"])
    L.update(q=q, z=z, s=s)
    return L


def _oracle_W(L):
    return O.dequant(L["q"], L["z"], L["s"], L["g_idx"], "engine")


def _check(layer, L, Ms, seed=0):
    W = _oracle_W(L)
    Wd = torch.from_numpy(W).cuda().double()
    g = torch.Generator(device="cuda").manual_seed(seed)
    for M in Ms:
        x = torch.randn(M, L["K"], dtype=torch.float16, device="cuda", generator=g)
        y = layer(x).double()
        ref = x.double() @ Wd                                   # float64 contraction over the ORACLE's weights
        err = ((y - ref).abs().max() / ref.abs().max()).item()
        assert err < TOL, f"M={M}: rel err {err}"
        rows = min(M, 8)                                        # pure-numpy arbiter on a few rows
        ref_np = O.matmul_ref(x[:rows].cpu().numpy(), W, None, acc=np.float64)
        err = float(np.abs(y[:rows].cpu().numpy() - ref_np).max() / np.abs(ref_np).max())
        assert err < TOL, f"M={M} (numpy rows): rel err {err}"


# ---- config 3: Llama-2-7B GPTQ prefill shapes through the tcgen05 GEMM -------------------------------------
@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 11008), (11008, 4096)])
def test_prefill_gptq_vs_oracle(K, N):
    L = _packed_layer("GPTQ", 4, 128, K, N, seed=K + N)
    _check(layer_from_dict(L), L, (512, 2048))


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 11008), (11008, 4096)])
@pytest.mark.parametrize("M", [1, 512, 2048])
def test_awq_vs_reference_cuda_full_size(K, N, M):
    """The reference's gemm_forward_cuda (gemm_cuda_gen.cu:1102-1161) on the same AWQ bytes, decode and prefill."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import awq_inference_engine as awq
    L = _packed_layer("GEMM", 4, 128, K, N, seed=K + N + 1)
    layer = layer_from_dict(L)
    x = torch.randn(M, K, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(M))
    ref = awq.gemm_forward_cuda(x, layer.qweight, layer.scales, layer.qzeros, 8)   # before the first forward releases the AWQ buffers
    torch.cuda.synchronize()
    y = layer(x)
    torch.cuda.synchronize()
    err = ((y.double() - ref.double()).abs().max() / ref.double().abs().max()).item()
    # the reference rounds its 8 split-K partials to fp16 before summing them (gemm_cuda_gen.cu:1114-1160): 2e-3
    assert err < 2e-3, err
    Wd = torch.from_numpy(_oracle_W(L)).cuda().double()
    arb = x.double() @ Wd
    assert ((y.double() - arb).abs().max() / arb.abs().max()).item() < TOL


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("K,N", [(4096, 4096), (11008, 4096)])
def test_gptq_vs_reference_ort_dequant_full_size(K, N):
    """ort_ops.dequant + matmul (quant_linear_gptq.py:81-85) at full size, M = 512."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import ort_ops as ort
    L = _packed_layer("GPTQ", 4, 128, K, N, seed=K + N + 2)
    layer = layer_from_dict(L)
    Wref = ort.dequant(layer.qweight, layer.scales, layer.qzeros, None, 128, 4, K, 0)
    x = torch.randn(512, K, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    y = layer(x)
    ref = x.double() @ Wref.double()
    torch.cuda.synchronize()
    assert ((y.double() - ref).abs().max() / ref.abs().max()).item() < TOL


# ---- config 2 decode shapes against the oracle (AWQ / GPTQ / Marlin bytes) ---------------------------------
@pytest.mark.parametrize("layout", ["GEMM", "GPTQ"])
@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 11008), (11008, 4096)])
def test_decode_full_size_vs_oracle(layout, K, N):
    L = _packed_layer(layout, 4, 128, K, N, seed=K + N + 3)
    _check(layer_from_dict(L), L, (1, 2, 8, 16))


# ---- config 4: Llama-2-13B act-order (g_idx) ------------------------------------------------------------
@pytest.mark.parametrize("K,N", [(5120, 5120), (5120, 13824), (13824, 5120)])
def test_act_order_13b_full_size(K, N):
    L = _packed_layer("GPTQ", 4, 128, K, N, seed=K + N + 4, act_order=True)
    _check(layer_from_dict(L), L, (1, 16, 512))


# ---- config 5: Mixtral expert FFN, HQQ 2/3/4/8-bit -------------------------------------------------------
@pytest.mark.parametrize("bits,gs", [(4, 64), (2, 64), (3, 64), (8, 128)])
@pytest.mark.parametrize("K,N", [(4096, 14336), (14336, 4096)])
def test_mixtral_hqq_full_size(bits, gs, K, N):
    L = _packed_layer("HQQ", bits, gs, K, N, seed=K + N + bits, float_zeros=(bits == 4))
    _check(layer_from_dict(L), L, (1, 16, 512))
