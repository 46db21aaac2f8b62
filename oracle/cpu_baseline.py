"""CPU restatement (torch ops, all host threads) of the reference's only CPU-capable QuantLinear
path: QuantLinearGPTQ.forward -> DequantizeLinearBlockWise -> torch.matmul
(qllm/modeling/q_layers/quant_linear_gptq.py:13-52, :81-85).  TEST / BASELINE INFRASTRUCTURE ONLY:
used by bench.py's cpu_baseline leg and `--impl reference`; never by the product path.

The reference's AWQ and Marlin layers have no CPU forward at all (quant_linear_awq.py:142-148 needs
the CUDA extension; the orchestrator repacks to GPTQ instead, auto_model_quantization.py:47-54), so
the CPU baseline for any pack mode is this GPTQ-layout path on the same logical weights.
"""
import time

import torch


def dequantize_blockwise(qweight, scales, qzeros, groupsize, bits, in_features):
    """int32 [K*b/32, N], scales [G,N], qzeros int32 [G, N*b/32] -> W [K, N] in scales.dtype.
    Same arithmetic as the reference: W = s*q - s*z with q, z unpacked by shift/and (2/4/8-bit)."""
    assert bits in (2, 4, 8)
    if groupsize == -1:                              # per-channel: one group spans K (quant_linear_gptq.py:98)
        groupsize = in_features
    per = 32 // bits
    shifts = torch.arange(0, 32, bits, dtype=torch.int32)
    mask = (1 << bits) - 1
    small = torch.int16 if bits == 8 else torch.int8
    z = ((qzeros.unsqueeze(2) >> shifts.view(1, 1, per)) & mask).to(small)          # [G, N/per, per]
    z = z.reshape(qzeros.shape[0], 1, -1)                                              # [G, 1, N]
    q = ((qweight.unsqueeze(1) >> shifts.view(1, per, 1)) & mask).to(small)          # [K/per, per, N]
    q = q.reshape(-1, groupsize, q.shape[-1])                                          # [G, group, N]
    s = scales.reshape(-1, 1, scales.shape[-1])
    w = s * q - (z * s).to(s.dtype)
    return w.reshape(in_features, -1)


def quant_linear_forward(x, qweight, scales, qzeros, groupsize, bits, in_features):
    return torch.matmul(x, dequantize_blockwise(qweight, scales, qzeros, groupsize, bits, in_features))


def make_block(hidden=4096, inter=11008, group=128, bits=4, dtype=torch.float16, seed=0):
    """Random packed GPTQ-layout buffers for one Llama decoder block's 7 QuantLinears."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(hidden, hidden)] * 4 + [(hidden, inter)] * 2 + [(inter, hidden)]
    layers = []
    for K, N in shapes:
        qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K * bits // 32, N), dtype=torch.int64, generator=g).to(torch.int32)
        qz = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // group, N * bits // 32), dtype=torch.int64, generator=g).to(torch.int32)
        sc = ((torch.rand(K // group, N, generator=g) * 0.4 + 0.8) / (6.5 * K ** 0.5)).to(dtype)
        layers.append((K, N, qw, sc, qz))
    return layers


def time_block(M=1, repeats=3, dtype=torch.float16, group=128, bits=4, hidden=4096, inter=11008):
    """Seconds for one decoder block's 7 QuantLinear forwards on CPU (best of `repeats` after one
    warm-up), chained like the GPU bench: q,k,v(h); o(v); gate,up(o); down(gate)."""
    layers = make_block(hidden, inter, group, bits, dtype)
    h = torch.randn(M, hidden).to(dtype)

    def run():
        fw = lambda i, x: quant_linear_forward(x, layers[i][2], layers[i][3], layers[i][4], group, bits, layers[i][0])
        q, k, v = fw(0, h), fw(1, h), fw(2, h)
        o = fw(3, v)
        gt, up = fw(4, o), fw(5, o)
        return fw(6, gt), q, k, up

    run()
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t0)
    return best
