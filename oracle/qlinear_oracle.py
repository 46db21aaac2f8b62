"""CPU oracle for QLLM's QuantLinear hot path.  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch numpy restatement of the reference's *algorithm* (on-disk
packed formats, dequantisation arithmetic, y = x @ W) for the path named by BASELINE.json.
It exists so that tests / smoke() / bench.py's cpu_baseline leg can check the CUDA engine;
nothing in the product package (`qllm_b200/`) may import it.

Parity pinning: the reference ships no tests or golden vectors.  The oracle is
pinned instead against outputs of the reference's own Python implementation executed in the
authoring container (tests/golden/make_golden.py imports /root/reference and stores packed
buffers + reference forward/unpack results in tests/golden/*.npz; tests/test_oracle_golden.py
replays them bit-exactly for integer work and to 1 fp16 ulp-ish tolerance for float work).

Reference citations (all paths relative to /root/reference):
  bit-stream codec ............ qllm/modeling/q_layers/compress_weight.py:10-92
  GPTQ dequant (+g_idx) ....... qllm/modeling/q_layers/quant_linear_gptq.py:13-52
  HQQ dequant (fp16 zeros) .... qllm/modeling/q_layers/quant_linear_hqq.py:8-28
  AWQ nibble interleave ....... qllm/modeling/q_layers/quant_linear_awq.py:95-140
  Marlin tile permutation ..... qllm/modeling/q_layers/quant_linear_marlin.py:18-42,:95-137
  AWQ CUDA rounding ........... csrc/awq_cuda/quantization/gemm_cuda_gen.cu:153-176
  ort CUDA rounding ........... csrc/ort_cuda/dq_gemv.cu:266-272,:292-307
"""
from __future__ import annotations

import numpy as np

AWQ_ORDER = np.array([0, 2, 4, 6, 1, 3, 5, 7])          # quant_linear_awq.py:107
AWQ_INV = np.argsort(AWQ_ORDER)                         # column 8c+i sits in nibble AWQ_INV[i]


# --------------------------------------------------------------------------------------
# LSB-first bit-stream codec along axis 0 (compress_weight.py:10-92).
# value r of a column occupies stream bits [r*bits, (r+1)*bits); 32 stream bits per int32 word.
# --------------------------------------------------------------------------------------
def pack_rows(vals: np.ndarray, bits: int) -> np.ndarray:
    """int [R, C] with values in [0, 2^bits) -> int32 [R*bits/32, C]."""
    vals = np.asarray(vals)
    R, C = vals.shape
    assert (R * bits) % 32 == 0, "row count must fill whole 32-bit words"
    v = vals.astype(np.uint64) & ((1 << bits) - 1)
    nwords = R * bits // 32
    out = np.zeros((nwords, C), dtype=np.uint64)
    for r in range(R):
        pos = r * bits
        w, o = divmod(pos, 32)
        out[w] |= (v[r] << o) & 0xFFFFFFFF
        if o + bits > 32:                                # straddles into the next word (odd bits)
            out[w + 1] |= v[r] >> (32 - o)
    return out.astype(np.uint32).view(np.int32)


def unpack_rows(packed: np.ndarray, bits: int, nrows: int | None = None) -> np.ndarray:
    """int32 [R*bits/32, C] -> int32 [R, C] (compress_weight.py:54-84)."""
    p = np.ascontiguousarray(packed).view(np.uint32).astype(np.uint64)
    nwords, C = p.shape
    R = nwords * 32 // bits if nrows is None else nrows
    mask = (1 << bits) - 1
    out = np.empty((R, C), dtype=np.int32)
    if 32 % bits == 0:
        per = 32 // bits
        for j in range(per):
            out[j::per] = ((p >> (bits * j)) & mask)[: (R - j + per - 1) // per]
        return out
    for r in range(R):
        pos = r * bits
        w, o = divmod(pos, 32)
        v = p[w] >> o
        if o + bits > 32:
            v = v | (p[w + 1] << (32 - o))
        out[r] = (v & mask).astype(np.int32)
    return out


# --------------------------------------------------------------------------------------
# GPTQ / HQQ layout
# --------------------------------------------------------------------------------------
def gptq_pack_qweight(q: np.ndarray, bits: int) -> np.ndarray:
    """q int [K, N] -> qweight int32 [K*bits/32, N] (bit-stream along K)."""
    return pack_rows(q, bits)


def gptq_unpack_qweight(qweight: np.ndarray, bits: int, K: int) -> np.ndarray:
    return unpack_rows(qweight, bits, K)


def gptq_pack_qzeros(z: np.ndarray, bits: int, zero_bias: int = 0) -> np.ndarray:
    """z int [G, N] -> qzeros int32 [G, N*bits/32] (bit-stream along N); stored value is
    (z - zero_bias) & mask (compress_weight.py:156-172; zero_bias=1 for AutoGPTQ files)."""
    zz = (np.asarray(z).astype(np.int64) - zero_bias) & ((1 << bits) - 1)
    return np.ascontiguousarray(pack_rows(zz.T, bits).T)


def gptq_unpack_qzeros(qzeros: np.ndarray, bits: int, N: int, zero_bias: int = 0) -> np.ndarray:
    z = unpack_rows(np.ascontiguousarray(np.asarray(qzeros).T), bits, N).T
    return ((z + zero_bias) & ((1 << bits) - 1)).astype(np.int32)


def autogptq_fix_qzeros(qzeros: np.ndarray, bits: int, N: int) -> np.ndarray:
    """QuantLinearGPTQ.handle_qzeros_for_autogptq (quant_linear_gptq.py:119-134): stored z-1 -> z."""
    z = gptq_unpack_qzeros(qzeros, bits, N, zero_bias=1)
    return gptq_pack_qzeros(z, bits)


# --------------------------------------------------------------------------------------
# AWQ "GEMM" layout (Appendix A.6): 4-bit, nibble j of word [k, c] = q[k, 8c + AWQ_ORDER[j]]
# --------------------------------------------------------------------------------------
def _awq_pack_rowwise(vals: np.ndarray) -> np.ndarray:
    """int [R, N] -> int32 [R, N/8] with the AWQ interleave inside each 8-column word."""
    R, N = vals.shape
    assert N % 8 == 0
    v = (np.asarray(vals).astype(np.uint64) & 0xF).reshape(R, N // 8, 8)[:, :, AWQ_ORDER]
    sh = (4 * np.arange(8, dtype=np.uint64)).reshape(1, 1, 8)
    return (v << sh).sum(axis=2).astype(np.uint32).view(np.int32)


def _awq_unpack_rowwise(packed: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(packed).view(np.uint32).astype(np.uint64)
    R, C = p.shape
    sh = (4 * np.arange(8, dtype=np.uint64)).reshape(1, 1, 8)
    nib = ((p[:, :, None] >> sh) & 0xF).astype(np.int32)      # nib[.., j] = column AWQ_ORDER[j]
    out = np.empty((R, C, 8), dtype=np.int32)
    out[:, :, AWQ_ORDER] = nib
    return out.reshape(R, C * 8)


def awq_pack_qweight(q: np.ndarray) -> np.ndarray:
    """q int [K, N] -> qweight int32 [K, N/8]."""
    return _awq_pack_rowwise(q)


def awq_unpack_qweight(qweight: np.ndarray) -> np.ndarray:
    return _awq_unpack_rowwise(qweight)


def awq_pack_qzeros(z: np.ndarray) -> np.ndarray:
    """z int [G, N] -> qzeros int32 [G, N/8] (same interleave)."""
    return _awq_pack_rowwise(z)


def awq_unpack_qzeros(qzeros: np.ndarray) -> np.ndarray:
    return _awq_unpack_rowwise(qzeros)


# --------------------------------------------------------------------------------------
# Marlin layout (Appendix A.7). Symmetric int4, stored value = q_signed + 8.
# --------------------------------------------------------------------------------------
def _marlin_perms():
    """Restates quant_linear_marlin.py:18-39 (_get_perms)."""
    perm = []
    for i in range(32):
        perm1 = []
        col = i // 4
        for block in (0, 1):
            for row in (2 * (i % 4), 2 * (i % 4) + 1, 2 * (i % 4 + 4), 2 * (i % 4 + 4) + 1):
                perm1.append(16 * row + col + 8 * block)
        for j in range(4):
            perm.extend(p + 256 * j for p in perm1)
    perm = np.array(perm).reshape(-1, 8)[:, AWQ_ORDER].ravel()
    scale_perm = [i + 8 * j for i in range(8) for j in range(8)]
    scale_perm_single = [2 * i + j for i in range(4) for j in (0, 1, 8, 9, 16, 17, 24, 25)]
    return perm, np.array(scale_perm), np.array(scale_perm_single)


MARLIN_PERM, MARLIN_SCALE_PERM, MARLIN_SCALE_PERM_SINGLE = _marlin_perms()


def marlin_pack(q: np.ndarray, scales: np.ndarray, group_size: int):
    """q int [K, N] in [0,15] (value = signed+8), scales fp16 [G, N] natural order
    -> (qweight int32 [K/16, 2N], scales fp16 [G, N] Marlin-permuted).
    Restates QuantLinearMarlin.pack (quant_linear_marlin.py:110-137) from the integer stage on."""
    K, N = q.shape
    assert K % 16 == 0 and N % 64 == 0
    tile = 16
    w = np.asarray(q).astype(np.int64).reshape(K // tile, tile, N // tile, tile)
    w = w.transpose(0, 2, 1, 3).reshape(K // tile, N * tile)
    res = w.reshape(-1, MARLIN_PERM.size)[:, MARLIN_PERM].reshape(w.shape)
    packed = np.zeros((res.shape[0], res.shape[1] // 8), dtype=np.uint64)
    for i in range(8):
        packed |= (res[:, i::8].astype(np.uint64) & 0xF) << np.uint64(4 * i)
    s = np.asarray(scales)
    if group_size != K:
        s = s.reshape(-1, MARLIN_SCALE_PERM.size)[:, MARLIN_SCALE_PERM]
    else:
        s = s.reshape(-1, MARLIN_SCALE_PERM_SINGLE.size)[:, MARLIN_SCALE_PERM_SINGLE]
    s = np.ascontiguousarray(s.reshape(-1, N))
    return packed.astype(np.uint32).view(np.int32), s


def marlin_unpack(qweight: np.ndarray, scales: np.ndarray, group_size: int, K: int):
    """Inverse of marlin_pack (the reference has none: quant_linear_marlin.py:139-140).
    Returns (q int32 [K, N] in [0,15], scales fp16 [G, N] natural order)."""
    p = np.ascontiguousarray(qweight).view(np.uint32).astype(np.uint64)
    R, C = p.shape
    N = C // 2
    res = np.empty((R, C * 8), dtype=np.int32)
    for i in range(8):
        res[:, i::8] = ((p >> np.uint64(4 * i)) & 0xF).astype(np.int32)
    w = np.empty_like(res).reshape(-1, MARLIN_PERM.size)
    w[:, MARLIN_PERM] = res.reshape(-1, MARLIN_PERM.size)
    w = w.reshape(R, N // 16, 16, 16).transpose(0, 2, 1, 3).reshape(K, N)
    s = np.asarray(scales)
    if group_size != K:
        t = np.empty_like(s).reshape(-1, MARLIN_SCALE_PERM.size)
        t[:, MARLIN_SCALE_PERM] = s.reshape(-1, MARLIN_SCALE_PERM.size)
    else:
        t = np.empty_like(s).reshape(-1, MARLIN_SCALE_PERM_SINGLE.size)
        t[:, MARLIN_SCALE_PERM_SINGLE] = s.reshape(-1, MARLIN_SCALE_PERM_SINGLE.size)
    return w, np.ascontiguousarray(t.reshape(-1, N))


# --------------------------------------------------------------------------------------
# Dequantisation arithmetic — the reference has several roundings for the same math
# (SURVEY Appendix B). W is [K, N]; group of row k is g_idx[k] (default k // group_size).
# --------------------------------------------------------------------------------------
def default_g_idx(K: int, group_size: int) -> np.ndarray:
    return (np.arange(K) // group_size).astype(np.int32)


def dequant(q: np.ndarray, z: np.ndarray, s: np.ndarray, g_idx: np.ndarray, mode: str = "engine"):
    """q int [K,N]; z int or float [G,N]; s fp16 [G,N]; returns W [K,N].

    mode:
      "engine" : fp16( (q - z) * s ) with one rounding  -- what the B200 engine feeds its MMAs and
                 what the AWQ/Marlin CUDA kernels compute (gemm_cuda_gen.cu:153-176).
      "torch"  : fp16(fp16(s*q) - fp16(s*z))            -- DequantizeLinearBlockWise /
                 DequantAndUnpack in fp16 (quant_linear_gptq.py:39-49, quant_linear_hqq.py:21-23).
      "ort"    : fp16(fma(q, s, -fp16(s*z)))            -- ort_ops.dequant (dq_gemv.cu:266-272).
      "exact"  : float64 (q - z) * s                    -- arbiter.
    """
    qk = np.asarray(q)
    sg = np.asarray(s)[g_idx]
    zg = np.asarray(z)[g_idx]
    if mode == "exact":
        return (qk.astype(np.float64) - zg.astype(np.float64)) * sg.astype(np.float64)
    if mode == "engine":
        d = qk.astype(np.float32) - zg.astype(np.float32)
        if np.issubdtype(np.asarray(z).dtype, np.floating):
            d = d.astype(np.float16).astype(np.float32)     # (q - z) is rounded to fp16 first
        return (d * sg.astype(np.float32)).astype(np.float16)
    s16 = sg.astype(np.float16)
    if mode == "torch":
        a = (s16 * qk.astype(np.float16)).astype(np.float16)
        b = (np.asarray(z).astype(np.float16) * np.asarray(s).astype(np.float16)).astype(np.float16)[g_idx]
        return (a - b).astype(np.float16)
    if mode == "ort":
        b = (np.asarray(z).astype(np.float16) * np.asarray(s).astype(np.float16)).astype(np.float16)[g_idx]
        return (qk.astype(np.float64) * s16.astype(np.float64) - b.astype(np.float64)).astype(np.float16)
    raise ValueError(mode)


def matmul_ref(x: np.ndarray, W: np.ndarray, bias: np.ndarray | None = None, acc=np.float32):
    """y = x @ W (+ bias), accumulated in `acc`, returned in `acc` (caller rounds)."""
    y = np.asarray(x).astype(acc) @ np.asarray(W).astype(acc)
    if bias is not None:
        y = y + np.asarray(bias).astype(acc)
    return y


# --------------------------------------------------------------------------------------
# AWQ GEMV layout (WQLinear_GEMV, quant_linear_awq.py:156-265; kernels gemv_cuda.cu:60-186)
# --------------------------------------------------------------------------------------
def awq_gemv_zeros_width(K: int, group_size: int) -> int:
    """calculate_zeros_width, quant_linear_awq.py:15-27."""
    mult = 1 if group_size >= 128 else {64: 2, 32: 4}[group_size]
    base = (K // group_size + 7) // 8
    return (base + mult - 1) // mult * mult


def awq_gemv_pack(q: np.ndarray, z: np.ndarray, s: np.ndarray, group_size: int):
    """q [K,N], z [G,N], s [G,N] -> qweight [N, K/8] (word w of row n: nibble i = q[8w+i, n], order_map 0..7,
    quant_linear_awq.py:224-232), qzeros [N, ZW] (nibble i of word c = z[8c+i, n], :242-253), scales [N, 8 ZW] (:204-210)."""
    K, N = q.shape
    G = z.shape[0]
    zw = awq_gemv_zeros_width(K, group_size)
    qweight = np.ascontiguousarray(pack_rows(q, 4).T)
    zp = np.zeros((zw * 8, N), dtype=np.int32)
    zp[:G] = z
    qzeros = np.ascontiguousarray(pack_rows(zp, 4).T)
    sp = np.zeros((N, zw * 8), dtype=np.float16)
    sp[:, :G] = np.asarray(s).T
    return qweight, qzeros, sp


def awq_gemv_unpack(qweight: np.ndarray, qzeros: np.ndarray, scales: np.ndarray, K: int, group_size: int):
    G = K // group_size
    q = unpack_rows(np.ascontiguousarray(np.asarray(qweight).T), 4, K)
    z = unpack_rows(np.ascontiguousarray(np.asarray(qzeros).T), 4, qzeros.shape[1] * 8)[:G]
    return q, np.ascontiguousarray(z), np.ascontiguousarray(np.asarray(scales)[:, :G].T)


# --------------------------------------------------------------------------------------
# ORT MatMulNBits blobs (QuantLinearORT.pack_on_device, quant_linear_onnxruntime.py:112-150; dequantize_blockwise_4bits :48-82)
# --------------------------------------------------------------------------------------
def ort_pack(q: np.ndarray, z: np.ndarray, s: np.ndarray, group_size: int):
    """q [K,N], z [G,N], s [G,N] -> qweight u8 [N, G, group/2] (low nibble = even k), qzeros u8 [N * ceil(G/2)], scales [N * G]."""
    K, N = q.shape
    G = z.shape[0]
    qt = q.T.astype(np.uint8)
    qweight = np.ascontiguousarray((qt[:, 0::2] | (qt[:, 1::2] << 4)).reshape(N, G, group_size // 2))
    zt = z.T.astype(np.uint8)
    if G & 1:
        zt = np.pad(zt, ((0, 0), (0, 1)))
    qzeros = np.ascontiguousarray((zt[:, 0::2] | (zt[:, 1::2] << 4)).reshape(-1))
    return qweight, qzeros, np.ascontiguousarray(np.asarray(s).T).reshape(-1)


def ort_unpack(qweight: np.ndarray, qzeros: np.ndarray, scales: np.ndarray, K: int, N: int, group_size: int):
    G = K // group_size
    b = np.asarray(qweight).reshape(N, K // 2).astype(np.int32)
    q = np.stack((b & 0xF, (b >> 4) & 0xF), axis=2).reshape(N, K).T
    zb = np.asarray(qzeros).reshape(N, (G + 1) // 2).astype(np.int32)
    z = np.stack((zb & 0xF, (zb >> 4) & 0xF), axis=2).reshape(N, -1)[:, :G].T
    return np.ascontiguousarray(q), np.ascontiguousarray(z), np.ascontiguousarray(np.asarray(scales).reshape(N, G).T)


# --------------------------------------------------------------------------------------
# One-call forward per pack mode, from the *packed* buffers (what a checkpoint holds)
# --------------------------------------------------------------------------------------
def unpack_layer(layout: str, bits: int, group_size: int, K: int, N: int, qweight, qzeros, scales,
                 g_idx=None, zero_bias: int = 0):
    """-> (q int32 [K,N], z [G,N] (int32 or float), s fp16 [G,N] natural order, g_idx int32 [K])."""
    layout = layout.upper()
    gs = K if group_size == -1 else group_size
    gi = default_g_idx(K, gs) if g_idx is None else np.asarray(g_idx).astype(np.int32)
    if layout == "GPTQ":
        return (gptq_unpack_qweight(qweight, bits, K), gptq_unpack_qzeros(qzeros, bits, N, zero_bias),
                np.asarray(scales), gi)
    if layout == "HQQ":
        return gptq_unpack_qweight(qweight, bits, K), np.asarray(qzeros), np.asarray(scales), gi
    if layout in ("GEMM", "AWQ"):
        assert bits == 4
        return awq_unpack_qweight(qweight), awq_unpack_qzeros(qzeros), np.asarray(scales), gi
    if layout == "GEMV":
        assert bits == 4
        q, z, s = awq_gemv_unpack(qweight, qzeros, scales, K, gs)
        return q, z, s, gi
    if layout == "ORT":
        assert bits == 4
        q, z, s = ort_unpack(qweight, qzeros, scales, K, N, gs)
        return q, z, s, gi
    if layout == "MARLIN":
        assert bits == 4
        q, s = marlin_unpack(qweight, scales, gs, K)
        return q, np.full((K // gs, N), 8, dtype=np.int32), s, gi
    raise ValueError(layout)


def forward(layout, bits, group_size, K, N, x, qweight, qzeros, scales, g_idx=None, bias=None,
            zero_bias=0, mode="engine", acc=np.float32):
    q, z, s, gi = unpack_layer(layout, bits, group_size, K, N, qweight, qzeros, scales, g_idx, zero_bias)
    W = dequant(q, z, s, gi, mode)
    return matmul_ref(x, W, bias, acc)


# --------------------------------------------------------------------------------------
# Synthetic layers (SURVEY §8(d) recipe) — deterministic, used by tests and bench
# --------------------------------------------------------------------------------------
def make_layer(layout: str, bits: int, group_size: int, K: int, N: int, seed: int = 0,
               act_order: bool = False, float_zeros: bool = False, bias: bool = False):
    """Returns dict(qweight, qzeros, scales, g_idx, bias, q, z, s) with packed buffers in `layout`."""
    rng = np.random.default_rng(seed)
    layout = layout.upper()
    gs = K if group_size == -1 else group_size
    G = K // gs
    q = rng.integers(0, 1 << bits, size=(K, N), dtype=np.int32)
    s = rng.uniform(0.002, 0.012, size=(G, N)).astype(np.float16)
    g_idx = default_g_idx(K, gs)
    if act_order:
        g_idx = g_idx[rng.permutation(K)].astype(np.int32)
    out = dict(layout=layout, bits=bits, group_size=gs, K=K, N=N, q=q, s=s, g_idx=g_idx,
               bias=(rng.standard_normal(N).astype(np.float16) if bias else None))
    if layout == "MARLIN":
        z = np.full((G, N), 8, dtype=np.int32)
        qw, sp = marlin_pack(q, s, gs)
        out.update(qweight=qw, qzeros=None, scales=sp, z=z)
    elif layout in ("GEMM", "AWQ"):
        z = rng.integers(0, 1 << bits, size=(G, N), dtype=np.int32)
        out.update(qweight=awq_pack_qweight(q), qzeros=awq_pack_qzeros(z), scales=s, z=z)
    elif layout == "ORT":
        z = rng.integers(0, 1 << bits, size=(G, N), dtype=np.int32)
        qw, qz, sp = ort_pack(q, z, s, gs)
        out.update(qweight=qw, qzeros=qz, scales=sp, z=z)
    elif layout == "GEMV":
        z = rng.integers(0, 1 << bits, size=(G, N), dtype=np.int32)
        qw, qz, sp = awq_gemv_pack(q, z, s, gs)
        out.update(qweight=qw, qzeros=qz, scales=sp, z=z)
    elif layout == "HQQ":
        z = rng.integers(0, 1 << bits, size=(G, N)).astype(np.float16)
        if float_zeros:
            z = (z + rng.uniform(-0.5, 0.5, size=z.shape)).astype(np.float16)
        out.update(qweight=gptq_pack_qweight(q, bits), qzeros=z, scales=s, z=z)
    elif layout == "GPTQ":
        z = rng.integers(0, 1 << bits, size=(G, N), dtype=np.int32)
        out.update(qweight=gptq_pack_qweight(q, bits), qzeros=gptq_pack_qzeros(z, bits), scales=s, z=z)
    else:
        raise ValueError(layout)
    return out
