"""Host-side (load/convert-time) codec for the packed formats, in torch so it runs on CPU or GPU.

Written from the format specification (reference: compress_weight.py:10-92,
quant_linear_awq.py:95-140, quant_linear_marlin.py:18-42,:95-137), not from the reference's code.
It is load-time plumbing behind QuantLinear.pack()/unpack(); the hot path never calls it.
"""
import torch

AWQ_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)


def pack_rows(vals: torch.Tensor, bits: int) -> torch.Tensor:
    """int [R, C], values in [0, 2^bits) -> int32 [R*bits/32, C]; LSB-first bit-stream along dim 0."""
    R, C = vals.shape
    if (R * bits) % 32:
        raise ValueError("rows * bits must be a multiple of 32")
    v = vals.to(torch.int64) & ((1 << bits) - 1)
    if 32 % bits == 0:
        per = 32 // bits
        sh = (torch.arange(per, device=v.device, dtype=torch.int64) * bits).view(1, per, 1)
        words = (v.view(R // per, per, C) << sh).sum(dim=1)
    else:
        bit = torch.arange(bits, device=v.device, dtype=torch.int64).view(1, bits, 1)
        stream = ((v.unsqueeze(1) >> bit) & 1).reshape(R * bits // 32, 32, C)
        pos = torch.arange(32, device=v.device, dtype=torch.int64).view(1, 32, 1)
        words = (stream << pos).sum(dim=1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    return words.to(torch.int32).contiguous()


def unpack_rows(packed: torch.Tensor, bits: int, rows: int) -> torch.Tensor:
    """int32 [R*bits/32, C] -> int32 [R, C]."""
    p = packed.to(torch.int64) & 0xFFFFFFFF
    W, C = p.shape
    if 32 % bits == 0:
        per = 32 // bits
        sh = (torch.arange(per, device=p.device, dtype=torch.int64) * bits).view(1, per, 1)
        out = ((p.unsqueeze(1) >> sh) & ((1 << bits) - 1)).reshape(W * per, C)
    else:
        pos = torch.arange(32, device=p.device, dtype=torch.int64).view(1, 32, 1)
        stream = ((p.unsqueeze(1) >> pos) & 1).reshape(W * 32 // bits, bits, C)
        bit = torch.arange(bits, device=p.device, dtype=torch.int64).view(1, bits, 1)
        out = (stream << bit).sum(dim=1)
    return out[:rows].to(torch.int32).contiguous()


# ---- GPTQ / HQQ -------------------------------------------------------------------------------
def gptq_pack_qzeros(z: torch.Tensor, bits: int, zero_bias: int = 0) -> torch.Tensor:
    zz = (z.to(torch.int64) - zero_bias) & ((1 << bits) - 1)
    return pack_rows(zz.t().contiguous(), bits).t().contiguous()


def gptq_unpack_qzeros(qzeros: torch.Tensor, bits: int, N: int, zero_bias: int = 0) -> torch.Tensor:
    z = unpack_rows(qzeros.t().contiguous(), bits, N).t()
    return ((z + zero_bias) & ((1 << bits) - 1)).to(torch.int32).contiguous()


# ---- AWQ GEMM ---------------------------------------------------------------------------------
def _awq_pack(v: torch.Tensor) -> torch.Tensor:
    R, N = v.shape
    order = torch.tensor(AWQ_ORDER, device=v.device)
    x = (v.to(torch.int64) & 0xF).view(R, N // 8, 8)[:, :, order]
    sh = (torch.arange(8, device=v.device, dtype=torch.int64) * 4).view(1, 1, 8)
    w = (x << sh).sum(dim=2)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.to(torch.int32).contiguous()


def _awq_unpack(p: torch.Tensor) -> torch.Tensor:
    R, C = p.shape
    sh = (torch.arange(8, device=p.device, dtype=torch.int64) * 4).view(1, 1, 8)
    nib = ((p.to(torch.int64).unsqueeze(2) & 0xFFFFFFFF) >> sh) & 0xF
    out = torch.empty_like(nib)
    out[:, :, torch.tensor(AWQ_ORDER, device=p.device)] = nib
    return out.reshape(R, C * 8).to(torch.int32)


awq_pack_qweight = _awq_pack
awq_unpack_qweight = _awq_unpack
awq_pack_qzeros = _awq_pack
awq_unpack_qzeros = _awq_unpack


# ---- AWQ GEMV layout (WQLinear_GEMV, quant_linear_awq.py:156-265) ------------------------------
def awq_gemv_zeros_width(in_features: int, group_size: int, pack_num: int = 8) -> int:
    """calculate_zeros_width (quant_linear_awq.py:15-27)."""
    if group_size >= 128:
        mult = 1
    elif group_size == 64:
        mult = 2
    elif group_size == 32:
        mult = 4
    else:
        raise NotImplementedError(group_size)
    base = (in_features // group_size + pack_num - 1) // pack_num
    return (base + mult - 1) // mult * mult


def awq_gemv_pack(q: torch.Tensor, z: torch.Tensor, scales: torch.Tensor, group_size: int):
    """q int [K,N], z int [G,N], scales [G,N] -> qweight i32 [N, K/8] (nibble i = k 8w+i), qzeros i32 [N, ZW]
    (nibble i of word c = group 8c+i), scales [N, 8 ZW] (zero padded)."""
    K, N = q.shape
    G = z.shape[0]
    zw = awq_gemv_zeros_width(K, group_size)
    qweight = pack_rows(q, 4).t().contiguous()
    zp = torch.zeros((zw * 8, N), dtype=torch.int32, device=z.device)
    zp[:G] = z.to(torch.int32)
    qzeros = pack_rows(zp, 4).t().contiguous()
    sp = torch.zeros((N, zw * 8), dtype=scales.dtype, device=scales.device)
    sp[:, :G] = scales.t()
    return qweight, qzeros, sp


def awq_gemv_unpack(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, K: int, group_size: int):
    """-> (q int32 [K,N], z int32 [G,N], scales [G,N])."""
    G = K // group_size
    q = unpack_rows(qweight.t().contiguous(), 4, K)
    z = unpack_rows(qzeros.t().contiguous(), 4, qzeros.shape[1] * 8)[:G]
    return q, z.contiguous(), scales[:, :G].t().contiguous()


# ---- ORT MatMulNBits blobs (QuantLinearORT, quant_linear_onnxruntime.py:85-174), 4-bit ----------
def ort_pack(q: torch.Tensor, z: torch.Tensor, scales: torch.Tensor, group_size: int):
    """q int [K,N], z int [G,N], scales [G,N] -> qweight u8 [N, G, group/2], qzeros u8 [N * ceil(G/2)], scales [N * G]
    (pack_on_device, quant_linear_onnxruntime.py:112-150)."""
    K, N = q.shape
    G = z.shape[0]
    qt = q.t().to(torch.uint8)                                   # [N, K]
    qweight = (qt[:, 0::2] | (qt[:, 1::2] << 4)).reshape(N, G, group_size // 2).contiguous()
    zt = z.t().to(torch.uint8)                                   # [N, G]
    if G & 1:
        zt = torch.nn.functional.pad(zt, (0, 1))
    qzeros = (zt[:, 0::2] | (zt[:, 1::2] << 4)).reshape(-1).contiguous()
    return qweight, qzeros, scales.t().contiguous().reshape(-1)


def ort_unpack(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, K: int, N: int, group_size: int):
    """-> (q int32 [K,N], z int32 [G,N], scales [G,N])."""
    G = K // group_size
    b = qweight.reshape(N, K // 2).to(torch.int32)
    q = torch.stack((b & 0xF, (b >> 4) & 0xF), dim=2).reshape(N, K).t().contiguous()
    zb = qzeros.reshape(N, (G + 1) // 2).to(torch.int32)
    z = torch.stack((zb & 0xF, (zb >> 4) & 0xF), dim=2).reshape(N, -1)[:, :G].t().contiguous()
    return q, z, scales.reshape(N, G).t().contiguous()


# ---- Marlin -----------------------------------------------------------------------------------
def _marlin_perms():
    perm = []
    for lane in range(32):
        col, q = lane // 4, lane % 4
        one = [16 * r + col + 8 * blk for blk in (0, 1) for r in (2 * q, 2 * q + 1, 2 * q + 8, 2 * q + 9)]
        for j in range(4):
            perm.extend(p + 256 * j for p in one)
    perm = torch.tensor(perm).view(-1, 8)[:, list(AWQ_ORDER)].reshape(-1)
    sp = torch.tensor([i + 8 * j for i in range(8) for j in range(8)])
    sps = torch.tensor([2 * i + j for i in range(4) for j in (0, 1, 8, 9, 16, 17, 24, 25)])
    return perm, sp, sps


_PERM, _SPERM, _SPERM1 = _marlin_perms()


def marlin_pack(q: torch.Tensor, scales: torch.Tensor, group_size: int):
    """q int [K,N] in [0,15]; scales [G,N] natural -> (qweight int32 [K/16, 2N], permuted scales)."""
    K, N = q.shape
    dev = q.device
    w = q.to(torch.int64).view(K // 16, 16, N // 16, 16).permute(0, 2, 1, 3).reshape(K // 16, N * 16)
    w = w.view(-1, _PERM.numel())[:, _PERM.to(dev)].view(K // 16, N * 16)
    sh = (torch.arange(8, device=dev, dtype=torch.int64) * 4).view(1, 1, 8)
    packed = ((w.view(K // 16, N * 2, 8) & 0xF) << sh).sum(dim=2)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed).to(torch.int32)
    sp = _SPERM1 if group_size == K else _SPERM
    s = scales.reshape(-1, sp.numel())[:, sp.to(dev)].reshape(-1, N).contiguous()
    return packed.contiguous(), s


def marlin_unpack(qweight: torch.Tensor, scales: torch.Tensor, group_size: int, K: int):
    R, C = qweight.shape
    N = C // 2
    dev = qweight.device
    sh = (torch.arange(8, device=dev, dtype=torch.int64) * 4).view(1, 1, 8)
    res = (((qweight.to(torch.int64) & 0xFFFFFFFF).unsqueeze(2) >> sh) & 0xF).reshape(R, C * 8)
    w = torch.empty_like(res).view(-1, _PERM.numel())
    w[:, _PERM.to(dev)] = res.view(-1, _PERM.numel())
    w = w.view(R, N // 16, 16, 16).permute(0, 2, 1, 3).reshape(K, N)
    sp = _SPERM1 if group_size == K else _SPERM
    s = torch.empty_like(scales).view(-1, sp.numel())
    s[:, sp.to(dev)] = scales.reshape(-1, sp.numel())
    return w.to(torch.int32).contiguous(), s.reshape(-1, N).contiguous()


def quantize_weight(weight_t: torch.Tensor, scales_t: torch.Tensor, zeros_t: torch.Tensor, g_idx: torch.Tensor,
                    maxq: int) -> torch.Tensor:
    """round((W + z*s)/s) per g_idx (the reference's _quant_weight contract, compress_weight.py:98-103).
    weight_t [K,N], scales_t/zeros_t [G,N] -> int32 [K,N] clamped to [0, maxq]."""
    s = scales_t[g_idx.long()]
    z = zeros_t[g_idx.long()]
    return torch.round((weight_t + z * s) / s).clamp_(0, maxq).to(torch.int32)
