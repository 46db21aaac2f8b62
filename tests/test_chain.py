"""Decode chain (b200q_chain_plan / b200q_chain_run, csrc/decode_chain.cu): every step's output against the numpy oracle
evaluated on the activations the engine itself produced for the previous step (so errors cannot accumulate into the
tolerance), plus determinism, the zeroed-workspace contract and NaN propagation."""
import numpy as np
import pytest
import torch

from oracle import qlinear_oracle as O
from tests.util import layer_from_dict, oracle_forward, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _mk(layout, gs, K, N, seed, bias=False, float_zeros=False, act_order=False):
    L = O.make_layer(layout, 4, gs, K, N, seed=seed, bias=bias, float_zeros=float_zeros, act_order=act_order)
    return L, layer_from_dict(L)


def _run_chain(specs, M, x0, wiring):
    """specs: list of lists of (L, layer); wiring[i] = (step j, layer idx, col offset) the input of step i comes from, or None
    (external x0).  Returns (ys, xs) as numpy: outputs per step/layer and the input each step actually consumed."""
    import qllm_b200
    dev = torch.device("cuda")
    ys = [[torch.full((M, L["N"]), float("nan"), dtype=torch.float16, device=dev) for L, _ in step] for step in specs]
    steps, xs = [], []
    for i, step in enumerate(specs):
        K = step[0][0]["K"]
        if wiring[i] is None:
            x = x0
        else:
            j, li, c0 = wiring[i]
            x = ys[j][li][:, c0:c0 + K]
        xs.append(x)
        steps.append(([layer for _, layer in step], x, ys[i]))
    chain = qllm_b200.DecodeChain(steps, M=M)
    chain.run()
    torch.cuda.synchronize()
    assert chain.error_code() == 0
    return chain, ys, xs


def _check_steps(specs, ys, xs):
    for i, step in enumerate(specs):
        x = xs[i].float().cpu().numpy().astype(np.float16)
        for (L, _), y in zip(step, ys[i]):
            ref = oracle_forward(L, x)
            got = y.float().cpu().numpy()
            assert np.isfinite(got).all(), f"step {i}: non-finite output"
            err = rel_err(got, ref)
            assert err < TOL, f"step {i} ({L['layout']} {L['K']}x{L['N']}): rel err {err}"


@pytest.mark.parametrize("M", [1, 2])
def test_chain_llama_like_block(M):
    """q|k|v -> o (reads v) -> gate|up (reads o) -> down (reads gate) -> next q|k|v (reads down): all inputs but the first
    are produced inside the chain."""
    H, I = 512, 1280
    specs = [[_mk("GPTQ", 128, H, H, 1), _mk("GPTQ", 128, H, H, 2), _mk("GPTQ", 128, H, H, 3, bias=True)],
             [_mk("GPTQ", 128, H, H, 4)],
             [_mk("GPTQ", 128, H, I, 5, bias=True), _mk("GPTQ", 128, H, I, 6)],
             [_mk("GPTQ", 128, I, H, 7)],
             [_mk("GPTQ", 128, H, H, 8), _mk("GPTQ", 128, H, H, 9), _mk("GPTQ", 128, H, H, 10)]]
    wiring = [None, (0, 2, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0)]
    x0 = torch.randn(M, H, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(M))
    chain, ys, xs = _run_chain(specs, M, x0, wiring)
    _check_steps(specs, ys, xs)
    # determinism + workspace contract: a second run is bit-identical and the counter region is zero again
    first = [[y.clone() for y in step] for step in ys]
    chain.run()
    torch.cuda.synchronize()
    for a, b in zip(first, ys):
        for ya, yb in zip(a, b):
            assert torch.equal(ya, yb)
    assert int(chain.workspace[:4096].count_nonzero()) == 0


@pytest.mark.parametrize("layout,gs,K,N", [("GPTQ", 64, 512, 256), ("GPTQ", 256, 512, 192), ("GPTQ", -1, 768, 96),
                                            ("HQQ", 64, 512, 128), ("HQQ", 128, 1024, 160), ("GEMM", 128, 512, 512),
                                            ("MARLIN", 128, 512, 256), ("GPTQ", 128, 2048, 32)])
def test_chain_single_step_layouts(layout, gs, K, N):
    """One-step chains over the supported (layout, group) combinations, incl. half tiles (N % 64 == 32) and fp16 zeros."""
    spec = [[_mk(layout, gs, K, N, seed=K + N, bias=True, float_zeros=(layout == "HQQ"))]]
    for M in (1, 2):
        x0 = torch.randn(M, K, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(K + M))
        _, ys, xs = _run_chain(spec, M, x0, [None])
        _check_steps(spec, ys, xs)


def test_chain_act_order_relayout():
    """desc_act checkpoint: row-permuted re-layout, x gathered through x_perm inside the chain's x stage (both from an
    external x and from the previous step's partial sums)."""
    H = 512
    specs = [[_mk("GPTQ", 128, H, H, 21, act_order=True)], [_mk("GPTQ", 128, H, 256, 22, act_order=True)]]
    x0 = torch.randn(1, H, dtype=torch.float16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    _, ys, xs = _run_chain(specs, 1, x0, [None, (0, 0, 0)])
    _check_steps(specs, ys, xs)


def test_chain_matches_per_layer_calls_full_size():
    """Llama-2-7B block shapes (AWQ checkpoints): the chain against the per-layer API on the same inputs."""
    import qllm_b200
    H, I, M = 4096, 11008, 1
    rng = np.random.default_rng(5)

    def awq(K, N, seed):
        import qllm_b200
        l = qllm_b200.WQLinear_GEMM(4, 128, K, N, False, dtype=torch.float16)
        g = torch.Generator().manual_seed(seed)
        l.qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, (K, N // 8), dtype=torch.int64, generator=g).to(torch.int32)
        l.qzeros = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 128, N // 8), dtype=torch.int64, generator=g).to(torch.int32)
        l.scales = ((torch.rand(K // 128, N, generator=g) * 0.4 + 0.8) / (6.5 * K ** 0.5)).to(torch.float16)
        return l.cuda()

    q, k, v, o = (awq(H, H, s) for s in range(4))
    gate, up, down = awq(H, I, 4), awq(H, I, 5), awq(I, H, 6)
    dev = torch.device("cuda")
    h = torch.randn(M, H, dtype=torch.float16, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    z = lambda n: torch.zeros(M, n, dtype=torch.float16, device=dev)
    yq, yk, yv, yo, yg, yu, yd = z(H), z(H), z(H), z(H), z(I), z(I), z(H)
    chain = qllm_b200.DecodeChain([([q, k, v], h, [yq, yk, yv]), ([o], yv, [yo]), ([gate, up], yo, [yg, yu]), ([down], yg, [yd])], M=M)
    chain.run()
    torch.cuda.synchronize()
    assert chain.error_code() == 0
    for layer, x, y in ((q, h, yq), (k, h, yk), (v, h, yv), (o, yv, yo), (gate, yo, yg), (up, yo, yu), (down, yg, yd)):
        ref = layer(x)
        W = layer.dequantize().double()
        exact = x.double() @ W
        scale = exact.abs().max().item()
        assert ((y.double() - exact).abs().max().item() / scale) < TOL
        assert ((y.double() - ref.double()).abs().max().item() / scale) < TOL


def test_chain_nan_propagates():
    spec = [[_mk("GPTQ", 128, 512, 128, 31)]]
    x0 = torch.randn(1, 512, dtype=torch.float16, device="cuda")
    x0[0, 77] = float("nan")
    _, ys, _ = _run_chain(spec, 1, x0, [None])
    assert torch.isnan(ys[0][0]).all()
    x0[0, 77] = float("inf")
    _, ys, _ = _run_chain(spec, 1, x0, [None])
    assert not torch.isfinite(ys[0][0]).any()


def test_chain_rejects_unsupported():
    import qllm_b200
    L, layer = _mk("GPTQ", 32, 256, 64, 41)               # group 32: no chain kernel
    x = torch.zeros(1, 256, dtype=torch.float16, device="cuda")
    y = torch.zeros(1, 64, dtype=torch.float16, device="cuda")
    with pytest.raises(ValueError):
        qllm_b200.DecodeChain([([layer], x, [y])], M=1)
