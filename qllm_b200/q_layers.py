"""Host-side mirror of QLLM's QuantLinear plug-in interface, backed by libb200q.so.

Class names, constructor signatures, attribute and buffer names/shapes/dtypes are those of the
reference's `qllm/modeling/q_layers/*` (so `load_state_dict` of any GPTQ/AWQ/HQQ/Marlin checkpoint
fills them unchanged and `make_mixbits_quant_linear` can instantiate them):

    QuantLinearGPTQ    quant_linear_gptq.py:92-143
    WQLinear_GEMM      quant_linear_awq.py:38-153
    QuantLinearMarlin  quant_linear_marlin.py:60-146   (+ the `dtype=` kwarg the reference's ctor lacks,
                                                        SURVEY §3.4, and an `unpack()`)
    QuantLinearHQQ     quant_linear_hqq.py:48-80
    select_quant_linear / make_mixbits_quant_linear    utils/modelutils.py:44-68, :161-182

`forward` goes through the C ABI only (include/b200q.h); there is no torch/CPU fallback: calling
forward on CPU tensors raises, exactly like the reference's AWQ/Marlin layers.
"""
import ctypes
import math
import os

import torch
import torch.nn as nn

from . import codec
from ._lib import (LAYOUT_AWQ_GEMM, LAYOUT_AWQ_GEMV, LAYOUT_GPTQ, LAYOUT_HQQ, LAYOUT_MARLIN, LAYOUT_ORT, Fusion, Layer, check, lib)

_RELAYOUT = (LAYOUT_AWQ_GEMM, LAYOUT_MARLIN, LAYOUT_AWQ_GEMV, LAYOUT_ORT)      # layouts that run on their exact K-packed re-layout

_workspaces = {}

# Decode (M <= 2) on AWQ-GEMM / Marlin layers runs on the layer's exact K-packed re-layout (the integer-tensor-path
# kernel, csrc/gemv_imma.cu); B200Q_DECODE_RELAYOUT=0 keeps decode on the checkpoint bytes (fp16-path kernels).
DECODE_RELAYOUT = os.environ.get("B200Q_DECODE_RELAYOUT", "1") != "0"
# Act-order (desc_act) GPTQ layers run on their exact row-permuted re-layout with activations gathered through the
# permutation (b200q_repack_actorder + b200q_layer.x_perm); B200Q_ACTORDER_RELAYOUT=0 keeps them on the generic kernel.
ACTORDER_RELAYOUT = os.environ.get("B200Q_ACTORDER_RELAYOUT", "1") != "0"
# AWQ-GEMM / Marlin layers run on ONE packed copy: the exact K-packed re-layout built at first use, after which the
# checkpoint-format buffers are released (state_dict() restores them through b200q_repack_from_gptq4).
# B200Q_KEEP_NATIVE=1 keeps both copies resident (+0.5 B/weight) and lets 3 <= M <= 8 read the checkpoint bytes.
KEEP_NATIVE = os.environ.get("B200Q_KEEP_NATIVE", "0") == "1"


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Zero-initialised scratch shared by all layers on (device, current stream); the engine leaves it
    zeroed after every call (b200q.h), so it is allocated once and only ever grown."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


class _B200QuantLinearBase(nn.Module):
    """Shared forward plumbing: builds the `b200q_layer` descriptor lazily and calls b200q_linear."""
    _layout = None

    def _init_common(self, bits, groupsize, infeatures, outfeatures, dtype):
        self.dtype = torch.get_default_dtype() if dtype is None else dtype
        self.infeatures = infeatures
        self.outfeatures = outfeatures
        self.bits = bits
        self.groupsize = groupsize if groupsize != -1 else infeatures
        self.maxq = 2 ** bits - 1
        self.orig_fp_weight = None
        self.act_order = None
        self.zero_bias = 0            # the loader rewrites AutoGPTQ zeros, so kernels always see 0
        self._desc = None
        self._desc_key = None

    # -- descriptor ---------------------------------------------------------------------------
    def _tensors(self):
        g_idx = self.g_idx if (self.act_order and isinstance(getattr(self, "g_idx", None), torch.Tensor)) else None
        return self.qweight, getattr(self, "qzeros", None), self.scales, g_idx, self.bias

    def _detect_act_order(self):
        g = getattr(self, "g_idx", None)
        if self._layout not in (LAYOUT_GPTQ, LAYOUT_ORT) or not isinstance(g, torch.Tensor):
            return False
        trivial = torch.arange(self.infeatures, device=g.device, dtype=torch.int64) // self.groupsize
        return bool((g.to(torch.int64) != trivial).any().item())

    def _descriptor(self):
        if self.act_order is None:
            self.act_order = self._detect_act_order()
        if getattr(self, "_consolidated", False):
            return self._shadow_desc                      # the only packed copy left: the K-packed re-layout (same q, z, s)
        qw, qz, sc, gi, bias = self._tensors()
        key = self._buffer_key()
        if self._desc is None or key != self._desc_key:
            if not qw.is_cuda:
                raise RuntimeError("qllm_b200 QuantLinear.forward needs CUDA buffers (no CPU fallback)")
            if sc.dtype != torch.float16:
                # scales arrive in the model dtype; the engine computes in fp16 (as the reference's
                # kernels do after their bf16->fp16 cast: quant_linear_awq.py:29-36, ort_ops.cc:79-90)
                self.scales = sc = sc.to(torch.float16)
            if self._layout == LAYOUT_HQQ and qz.dtype != torch.float16:
                self.qzeros = qz = qz.to(torch.float16)
            if bias is not None and bias.dtype != torch.float16:
                self._bias16 = bias.to(torch.float16)
            else:
                self._bias16 = bias
            for t in (qw, qz, sc, gi, self._bias16):
                if t is not None and not t.is_contiguous():
                    raise ValueError("QuantLinear buffers must be contiguous")
            d = Layer()
            d.layout, d.bits, d.group_size = self._layout, self.bits, self.groupsize
            d.K, d.N, d.zero_bias = self.infeatures, self.outfeatures, self.zero_bias
            d.qweight = qw.data_ptr()
            d.qzeros = qz.data_ptr() if qz is not None else None
            d.scales = sc.data_ptr()
            d.g_idx = gi.data_ptr() if gi is not None else None
            d.bias = self._bias16.data_ptr() if self._bias16 is not None else None
            self._desc = d
            self._desc_key = self._buffer_key()
        return self._desc

    def _buffer_key(self):
        """Identity AND version of every buffer: an in-place update (load_state_dict without assign=True, .copy_())
        keeps data_ptr() but bumps _version, and must invalidate the descriptor and every derived re-layout."""
        return tuple((0, 0) if t is None else (t.data_ptr(), t._version) for t in self._tensors())

    # -- act-order: groups made contiguous once, activations gathered at run time ------------------
    def _fast_descriptor(self):
        """The descriptor forward() hands to the engine for GPTQ/HQQ layers: the checkpoint buffers, or -- act-order
        checkpoints whose groups all own exactly `groupsize` rows -- {row-permuted qweight, original qzeros/scales,
        g_idx = NULL, x_perm = stable argsort(g_idx)}, built once (exact integer re-layout on the GPU)."""
        desc = self._descriptor()
        if not (ACTORDER_RELAYOUT and self.act_order and self._layout == LAYOUT_GPTQ):
            return desc
        key = self._desc_key
        if getattr(self, "_ao_key", None) != key:
            dev = self.qweight.device
            g = self.g_idx.to(device=dev, dtype=torch.int64)
            perm = torch.argsort(g, stable=True)
            K, gs = self.infeatures, self.groupsize
            regular = K % gs == 0 and K % 2 == 0 and bool((g[perm] == torch.arange(K, device=dev) // gs).all().item())
            self._ao_desc = None
            if regular:
                perm32 = perm.to(torch.int32).contiguous()
                qw = torch.empty_like(self.qweight)
                check(lib.b200q_repack_actorder(ctypes.byref(desc), perm32.data_ptr(), qw.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream), "b200q_repack_actorder")
                torch.cuda.current_stream(dev).synchronize()    # one-time: the kernels prefetch weights ahead of their stream dependency
                d = Layer()
                d.layout, d.bits, d.group_size, d.K, d.N, d.zero_bias = desc.layout, desc.bits, desc.group_size, desc.K, desc.N, desc.zero_bias
                d.qweight, d.qzeros, d.scales, d.g_idx, d.bias, d.x_perm = qw.data_ptr(), desc.qzeros, desc.scales, None, desc.bias, perm32.data_ptr()
                self._ao, self._ao_desc = (qw, perm32), d
            self._ao_key = key
        return self._ao_desc if self._ao_desc is not None else desc

    # -- K-packed shadow for the tensor-core GEMM (AWQ / Marlin until their native producers exist) --
    def _gemm_descriptor(self):
        """b200q_layer for the kernels that need K-packed words (tcgen05 GEMM, integer-path decode, decode chain).
        GPTQ/HQQ: the checkpoint buffers themselves.  AWQ/Marlin: a one-time exact integer re-layout
        (b200q_repack_gptq4); unless B200Q_KEEP_NATIVE=1 the checkpoint-format buffers are then released, so the
        layer holds ONE packed copy (state_dict() restores the checkpoint format through b200q_repack_from_gptq4)."""
        if self._layout not in _RELAYOUT or (self._layout == LAYOUT_ORT and self.act_order):
            return self._fast_descriptor()
        if getattr(self, "_consolidated", False):
            return self._shadow_desc
        desc = self._descriptor()
        key = self._desc_key
        if getattr(self, "_shadow_key", None) != key:
            dev = self.qweight.device
            K, N, G = self.infeatures, self.outfeatures, self.infeatures // self.groupsize
            qw = torch.empty((K // 8, N), dtype=torch.int32, device=dev)
            qz = torch.empty((G, N // 8), dtype=torch.int32, device=dev)
            sc = torch.empty((G, N), dtype=torch.float16, device=dev)
            check(lib.b200q_repack_gptq4(ctypes.byref(desc), qw.data_ptr(), qz.data_ptr(), sc.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream), "b200q_repack_gptq4")
            torch.cuda.current_stream(dev).synchronize()        # one-time: the kernels prefetch weights ahead of their stream dependency
            d = Layer()
            d.layout, d.bits, d.group_size, d.K, d.N, d.zero_bias = LAYOUT_GPTQ, 4, self.groupsize, K, N, 0
            d.qweight, d.qzeros, d.scales, d.g_idx, d.bias = qw.data_ptr(), qz.data_ptr(), sc.data_ptr(), None, desc.bias
            self._shadow, self._shadow_desc, self._shadow_key = (qw, qz, sc), d, key
            if not KEEP_NATIVE:
                self._release_native()
        return self._shadow_desc

    # -- one packed copy: release / restore the checkpoint-format buffers --------------------------------
    def _release_native(self):
        self._native_meta = {n: (tuple(getattr(self, n).shape), getattr(self, n).dtype) for n in ("qweight", "qzeros", "scales")
                             if isinstance(getattr(self, n, None), torch.Tensor)}
        dev = self.qweight.device
        for n, (_, dt) in self._native_meta.items():
            setattr(self, n, torch.empty(0, dtype=dt, device=dev))
        self._consolidated = True
        self._desc = None

    def _native_tensors(self):
        """Checkpoint-format (qweight, qzeros, scales) rebuilt from the K-packed copy (exact)."""
        qw, qz, sc = self._shadow
        dev = qw.device
        meta = self._native_meta

        def alloc(name):                                  # outputs are OR-ed / memset in whole 32-bit words
            shape, dt = meta[name]
            nbytes = int(torch.tensor(shape).prod().item()) * torch.empty(0, dtype=dt).element_size()
            return torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=dev), shape, dt, nbytes

        bq, sq, dq, nq = alloc("qweight")
        bz = alloc("qzeros") if "qzeros" in meta else None
        out_sc = torch.empty(meta["scales"][0], dtype=torch.float16, device=dev)
        check(lib.b200q_repack_from_gptq4(ctypes.byref(self._shadow_desc), self._layout, bq.data_ptr(),
                                          None if bz is None else bz[0].data_ptr(), out_sc.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream), "b200q_repack_from_gptq4")
        view = lambda b, shape, dt, n: b.view(torch.uint8)[:n].view(dt).reshape(shape)
        return view(bq, sq, dq, nq), (None if bz is None else view(*bz)), out_sc.to(meta["scales"][1])

    def _restore_native(self):
        if getattr(self, "_consolidated", False):
            qw, qz, sc = self._native_tensors()
            self.qweight, self.scales = qw, sc
            if qz is not None:
                self.qzeros = qz
            self._consolidated = False
            self._shadow = self._shadow_desc = self._shadow_key = None
            self._desc = None

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        if getattr(self, "_consolidated", False):
            qw, qz, sc = self._native_tensors()
            super()._save_to_state_dict(destination, prefix, keep_vars)
            destination[prefix + "qweight"], destination[prefix + "scales"] = qw.contiguous(), sc
            if qz is not None:
                destination[prefix + "qzeros"] = qz
            return
        super()._save_to_state_dict(destination, prefix, keep_vars)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        self._restore_native()                      # buffers regain their checkpoint shapes before they are overwritten
        self._desc = None
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _decode_descriptor(self, M):
        """Descriptor the decode kernels should read at batch M: the checkpoint buffers, or -- AWQ-GEMM / Marlin at
        M <= 2 -- the one-time exact K-packed re-layout that the integer-tensor-path kernel consumes."""
        if self._layout in _RELAYOUT and ((DECODE_RELAYOUT and M <= 2) or not KEEP_NATIVE or self._layout in (LAYOUT_AWQ_GEMV, LAYOUT_ORT)
                                                                 or getattr(self, "_consolidated", False)):
            return self._gemm_descriptor()
        return self._fast_descriptor()

    # -- forward ------------------------------------------------------------------------------
    def __call__(self, x):
        if NVTX:             # B200Q_NVTX=1: one range per QuantLinear call (shows up in nsys / ncu --nvtx timelines)
            torch.cuda.nvtx.range_push(f"b200q.{type(self).__name__} {tuple(x.shape)}x{self.infeatures}->{self.outfeatures} w{self.bits}g{self.groupsize}")
            try:
                return self._call(x)
            finally:
                torch.cuda.nvtx.range_pop()
        return self._call(x)

    def _call(self, x):
        grp = getattr(self, "_sibling_group", None)
        if grp is not None:
            return grp.forward_of(self, x)
        return super().__call__(x)

    def forward(self, x):
        if x.dtype == torch.bfloat16 and x.is_cuda:      # bf16 model: the library converts (in the decode kernel itself at M <= 2)
            return self.forward_fused(x)
        desc = self._decode_descriptor(8) if self._layout in _RELAYOUT else self._fast_descriptor()
        out_shape = x.shape[:-1] + (self.outfeatures,)
        x2 = x.reshape(-1, x.shape[-1])
        if x2.dtype != torch.float16:
            x2 = x2.to(torch.float16)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        M = x2.shape[0]
        y = torch.empty((M, self.outfeatures), dtype=torch.float16, device=x.device)
        if M > 0:
            if M > lib.b200q_gemv_max_m() and lib.b200q_select_kernel(ctypes.byref(desc), M) != 2:
                desc = self._gemm_descriptor()
            elif M <= 2:
                desc = self._decode_descriptor(M)
            need = lib.b200q_workspace_bytes(ctypes.byref(desc), M)
            ws = _workspace(x.device, need)
            st = lib.b200q_linear(ctypes.byref(desc), x2.data_ptr(), M, x2.stride(0), y.data_ptr(), y.stride(0),
                                  ws.data_ptr(), ws.numel(), torch.cuda.current_stream(x.device).cuda_stream)
            check(st, type(self).__name__ + ".forward")
        if y.dtype != x.dtype:
            y = y.to(x.dtype)
        return y.reshape(out_shape)

    def forward_fused(self, x, x_mul=None, residual=None):
        """forward() with the Linear's element-wise neighbours fused in (b200q_linear_ex):
        y = (silu(x) * x_mul if x_mul is not None else x) @ W + bias (+ residual).  Same results as the separate fp16 ops."""
        out_shape = x.shape[:-1] + (self.outfeatures,)
        # bf16 in, bf16 out without a cast on this side (b200q_fusion.act_dtype): every operand must be bf16 then
        bf16 = x.dtype == torch.bfloat16 and all(t is None or t.dtype == torch.bfloat16 for t in (x_mul, residual))
        act = torch.bfloat16 if bf16 else torch.float16
        def prep(t):
            t2 = t.reshape(-1, t.shape[-1])
            if t2.dtype != act:
                t2 = t2.to(act)
            return t2 if t2.stride(-1) == 1 else t2.contiguous()
        x2 = prep(x)
        M = x2.shape[0]
        fu = Fusion()
        fu.act_dtype = 1 if bf16 else 0
        keep = []
        if x_mul is not None:
            xm = prep(x_mul)
            if xm.stride(0) != x2.stride(0):
                xm, x2 = xm.contiguous(), x2.contiguous()
            keep.append(xm)
            fu.x_mul = xm.data_ptr()
        if residual is not None:
            r2 = prep(residual)
            keep.append(r2)
            fu.residual, fu.ldres = r2.data_ptr(), r2.stride(0)
        y = torch.empty((M, self.outfeatures), dtype=act, device=x.device)
        if M > 0:
            desc = self._decode_descriptor(8) if self._layout in _RELAYOUT else self._fast_descriptor()
            if M > lib.b200q_gemv_max_m() and lib.b200q_select_kernel(ctypes.byref(desc), M) != 2:
                desc = self._gemm_descriptor()
            elif M <= 2:
                desc = self._decode_descriptor(M)
            ws = _workspace(x.device, lib.b200q_workspace_bytes_ex(ctypes.byref(desc), M, ctypes.byref(fu)))
            check(lib.b200q_linear_ex(ctypes.byref(desc), x2.data_ptr(), M, x2.stride(0), y.data_ptr(), y.stride(0), ctypes.byref(fu),
                                      ws.data_ptr(), ws.numel(), torch.cuda.current_stream(x.device).cuda_stream),
                  type(self).__name__ + ".forward_fused")
        if y.dtype != x.dtype:
            y = y.to(x.dtype)
        return y.reshape(out_shape)

    def dequantize(self) -> torch.Tensor:
        """fp16 W[K, N] (= nn.Linear.weight.T) straight from the packed buffers (b200q_dequant)."""
        desc = self._descriptor()
        w = torch.empty((self.infeatures, self.outfeatures), dtype=torch.float16, device=self.qweight.device)
        check(lib.b200q_dequant(ctypes.byref(desc), w.data_ptr(), torch.cuda.current_stream(w.device).cuda_stream))
        return w

    def unpack_int(self):
        """(q int32 [K,N], z int32 [G,N] or None) via b200q_unpack -- the bit-exact gate."""
        desc = self._descriptor()
        dev = self.qweight.device
        q = torch.empty((self.infeatures, self.outfeatures), dtype=torch.int32, device=dev)
        G = self.infeatures // self.groupsize
        z = None if self._layout == LAYOUT_HQQ else torch.empty((G, self.outfeatures), dtype=torch.int32, device=dev)
        check(lib.b200q_unpack(ctypes.byref(desc), q.data_ptr(), None if z is None else z.data_ptr(),
                               torch.cuda.current_stream(dev).cuda_stream))
        return q, z

    def _default_g_idx(self):
        return (torch.arange(self.infeatures, dtype=torch.int32) // self.groupsize).to(torch.int32)

    def extra_repr(self):
        return "infeatures={}, outfeatures={}, bias={}, bits={}, groupsize={}, pack_mode={}".format(
            self.infeatures, self.outfeatures, self.bias is not None, self.bits, self.groupsize, self.pack_mode)


def _pack_device():
    return "cuda" if torch.cuda.is_available() else "cpu"


class QuantLinearGPTQ(_B200QuantLinearBase):
    """pack_mode=GPTQ: qweight i32 [K*b/32, N], qzeros i32 [G, N*b/32], scales [G,N], g_idx i32 [K]."""
    _layout = LAYOUT_GPTQ

    def __init__(self, bits, groupsize, infeatures, outfeatures, bias, dtype=None):
        super().__init__()
        if bits not in [2, 3, 4, 5, 6, 7, 8]:
            raise NotImplementedError("Only 2,3,4,5,6,7,8 bits are supported.")
        self._init_common(bits, groupsize, infeatures, outfeatures, dtype)
        self.pack_mode = "GPTQ"
        G = math.ceil(infeatures / self.groupsize)
        self.register_buffer("qweight", torch.zeros((infeatures // 32 * bits, outfeatures), dtype=torch.int32))
        self.register_buffer("qzeros", torch.zeros((G, outfeatures // 32 * bits), dtype=torch.int32))
        self.register_buffer("scales", torch.zeros((G, outfeatures), dtype=self.dtype))
        self.register_buffer("g_idx", self._default_g_idx())
        if bias:
            self.register_buffer("bias", torch.zeros((outfeatures), dtype=self.dtype))
        else:
            self.bias = None

    def handle_qzeros_for_autogptq(self):
        """AutoGPTQ stores z-1; rewrite to z once at load (reference: quant_linear_gptq.py:119-134)."""
        if self.qzeros.numel() == 0:
            return
        z = codec.gptq_unpack_qzeros(self.qzeros, self.bits, self.outfeatures, zero_bias=1)
        self.qzeros = codec.gptq_pack_qzeros(z, self.bits).to(self.qzeros.device)
        self._desc = None

    def pack(self, linear, scales, zeros, g_idx=None):
        """scales/zeros are the quantiser's [N, G] tensors (reference contract, compress_weight.py:204-210)."""
        dev = _pack_device()
        g = self._default_g_idx() if g_idx is None else g_idx.to(torch.int32).cpu()
        s_t = scales.t().contiguous().to(dev).float()
        z_t = zeros.t().contiguous().to(dev).float()
        q = codec.quantize_weight(linear.weight.data.t().to(dev).float(), s_t, z_t, g.to(dev), self.maxq)
        self.qweight = codec.pack_rows(q, self.bits).cpu()
        self.qzeros = codec.gptq_pack_qzeros(z_t.round().to(torch.int32), self.bits).cpu()
        self.scales = s_t.to(self.dtype).cpu()
        self.g_idx = g
        if linear.bias is not None:
            self.bias = linear.bias.detach().clone().to(self.dtype).cpu()
        self._desc, self.act_order = None, None

    def unpack(self):
        """-> (fp16 weight [N, K], scales [G, N], zeros int [G, N]) like CompressWeight.unpack."""
        q = codec.unpack_rows(self.qweight, self.bits, self.infeatures)
        z = codec.gptq_unpack_qzeros(self.qzeros, self.bits, self.outfeatures)
        gi = self.g_idx.long().to(q.device)
        s = self.scales.float()
        w = ((q.float() - z.float()[gi]) * s[gi]).to(torch.float16)
        return w.t().contiguous().cpu(), self.scales.cpu(), z.cpu()


class QuantLinearHQQ(_B200QuantLinearBase):
    """HQQ: qweight as GPTQ, qzeros/scales floating [G, N] (quant_linear_hqq.py:48-80)."""
    _layout = LAYOUT_HQQ

    def __init__(self, bits, groupsize, infeatures, outfeatures, bias, dtype=None):
        super().__init__()
        if bits not in [2, 3, 4, 5, 6, 7, 8]:
            raise NotImplementedError("Only 2,3,4,5,6,7,8 bits are supported.")
        self._init_common(bits, groupsize, infeatures, outfeatures, dtype)
        self.pack_mode = "HQQ"
        G = math.ceil(infeatures / self.groupsize)
        self.g_idx = self._default_g_idx()      # plain attribute, not a buffer (quant_linear_hqq.py:60)
        self.register_buffer("qweight", torch.zeros((infeatures // 32 * bits, outfeatures), dtype=torch.int32))
        self.register_buffer("qzeros", torch.zeros((G, outfeatures), dtype=self.dtype))
        self.register_buffer("scales", torch.zeros((G, outfeatures), dtype=self.dtype))
        if bias:
            self.register_buffer("bias", torch.zeros((outfeatures), dtype=self.dtype))
        else:
            self.bias = None

    def pack(self, linear, scales, zeros, g_idx=None):
        dev = _pack_device()
        s_t = scales.t().contiguous().to(dev).float()
        z_t = zeros.t().contiguous().to(dev).float()
        q = codec.quantize_weight(linear.weight.data.t().to(dev).float(), s_t, z_t, self._default_g_idx().to(dev), self.maxq)
        self.qweight = codec.pack_rows(q, self.bits).cpu()
        self.qzeros = z_t.to(self.dtype).cpu()
        self.scales = s_t.to(self.dtype).cpu()
        if linear.bias is not None:
            self.bias = linear.bias.detach().clone().to(self.dtype).cpu()
        self._desc = None

    def unpack(self):
        q = codec.unpack_rows(self.qweight, self.bits, self.infeatures)
        gi = self._default_g_idx().long().to(q.device)
        w = ((q.float() - self.qzeros.float()[gi]) * self.scales.float()[gi]).to(torch.float16)
        return w.t().contiguous().cpu(), self.scales.cpu(), self.qzeros.cpu()


class WQLinear_GEMM(_B200QuantLinearBase):
    """pack_mode=GEMM (AWQ): qweight i32 [K, N/8], qzeros i32 [G, N/8], nibble order 0,2,4,6,1,3,5,7."""
    _layout = LAYOUT_AWQ_GEMM

    def __init__(self, w_bit, group_size, in_features, out_features, bias, dtype=None):
        super().__init__()
        if w_bit not in [4]:
            raise NotImplementedError("Only 4-bit are supported for now.")
        self._init_common(w_bit, group_size, in_features, out_features, dtype)
        self.w_bit = w_bit
        self.group_size = self.groupsize
        self.pack_mode = "GEMM"
        assert in_features % self.group_size == 0
        assert out_features % (32 // w_bit) == 0
        self.g_idx = self._default_g_idx()
        self.register_buffer("qweight", torch.zeros((in_features, out_features // 8), dtype=torch.int32))
        self.register_buffer("qzeros", torch.zeros((in_features // self.group_size, out_features // 8), dtype=torch.int32))
        self.register_buffer("scales", torch.zeros((in_features // self.group_size, out_features), dtype=self.dtype))
        if bias:
            self.register_buffer("bias", torch.zeros((out_features), dtype=self.dtype))
        else:
            self.bias = None

    def pack(self, linear, scales, zeros, g_idx=None):
        if g_idx is not None:
            triv = self._default_g_idx()
            assert torch.equal(g_idx.cpu().to(torch.int32), triv), "AWQ GEMM layout has no act-order"
        dev = _pack_device()
        s_t = scales.t().contiguous().to(dev).float()
        z_t = zeros.t().contiguous().to(dev).float()
        q = codec.quantize_weight(linear.weight.data.t().to(dev).float(), s_t, z_t, self._default_g_idx().to(dev), self.maxq)
        self.qweight = codec.awq_pack_qweight(q).cpu()
        self.qzeros = codec.awq_pack_qzeros(z_t.round().to(torch.int32)).cpu()
        self.scales = s_t.to(self.dtype).cpu()
        if linear.bias is not None:
            self.bias = linear.bias.detach().clone().to(self.dtype).cpu()
        self._desc = None

    def unpack(self):
        if getattr(self, "_consolidated", False):
            qw, qz, sc = self._native_tensors()
        else:
            qw, qz, sc = self.qweight, self.qzeros, self.scales
        q = codec.awq_unpack_qweight(qw)
        z = codec.awq_unpack_qzeros(qz)
        gi = self._default_g_idx().long().to(q.device)
        w = ((q.float() - z.float()[gi]) * sc.float()[gi]).to(torch.float16)
        return w.t().contiguous().cpu(), sc.cpu(), z.cpu()


class WQLinear_GEMV(_B200QuantLinearBase):
    """pack_mode=GEMV (AWQ): qweight i32 [N, K/8], qzeros i32 [N, ZW], scales [N, 8 ZW] (quant_linear_awq.py:156-265).
    The reference's dispatch never selects it (utils/modelutils.py:52-67) but its kernels are built and exported
    (gemv_forward_cuda / gemmv2_forward_cuda, pybind_awq.cpp:17-18), so checkpoints in this layout exist."""
    _layout = LAYOUT_AWQ_GEMV

    def __init__(self, w_bit, group_size, in_features, out_features, bias, dtype=None):
        super().__init__()
        if w_bit not in [4]:
            raise NotImplementedError("Only 4-bit are supported for now.")
        self._init_common(w_bit, group_size, in_features, out_features, dtype)
        self.in_features, self.out_features = in_features, out_features
        self.w_bit = w_bit
        self.group_size = self.groupsize
        self.split_k_iters = 8
        self.pack_mode = "GEMV"
        assert in_features % self.group_size == 0
        assert out_features % (32 // w_bit) == 0
        zw = codec.awq_gemv_zeros_width(in_features, self.group_size)
        self.g_idx = self._default_g_idx()
        self.register_buffer("qweight", torch.zeros((out_features, in_features // 8), dtype=torch.int32))
        self.register_buffer("qzeros", torch.zeros((out_features, zw), dtype=torch.int32))
        self.register_buffer("scales", torch.zeros((out_features, zw * 8), dtype=self.dtype))
        if bias:
            self.register_buffer("bias", torch.zeros((out_features), dtype=self.dtype))
        else:
            self.bias = None

    def pack(self, linear, scales, zeros, g_idx=None):
        dev = _pack_device()
        s_t = scales.t().contiguous().to(dev).float()
        z_t = zeros.t().contiguous().to(dev).float()
        q = codec.quantize_weight(linear.weight.data.t().to(dev).float(), s_t, z_t, self._default_g_idx().to(dev), self.maxq)
        qw, qz, sc = codec.awq_gemv_pack(q, z_t.round().to(torch.int32), s_t.to(self.dtype), self.group_size)
        self.qweight, self.qzeros, self.scales = qw.cpu(), qz.cpu(), sc.cpu()
        if linear.bias is not None:
            self.bias = linear.bias.detach().clone().to(self.dtype).cpu()
        self._desc = None

    def unpack(self):
        if getattr(self, "_consolidated", False):
            qw, qz, sc = self._native_tensors()
        else:
            qw, qz, sc = self.qweight, self.qzeros, self.scales
        q, z, s = codec.awq_gemv_unpack(qw, qz, sc, self.infeatures, self.group_size)
        gi = self._default_g_idx().long().to(q.device)
        w = ((q.float() - z.float()[gi]) * s.float()[gi]).to(torch.float16)
        return w.t().contiguous().cpu(), s.cpu(), z.cpu()


class QuantLinearORT(_B200QuantLinearBase):
    """pack_mode=ORT: com.microsoft::MatMulNBits blobs (quant_linear_onnxruntime.py:85-174), 4-bit: qweight u8 [N, G, group/2],
    qzeros u8 [N * ceil(G/2)] (two zero points per byte), scales [N * G], g_idx i32 [K].  The reference runs it through
    ort_ops.Dequantize4Bits (csrc/ort_cuda/dq.cu:79-213) or a torch dequant, then torch.matmul (:31-43)."""
    _layout = LAYOUT_ORT

    def __init__(self, bits, groupsize, infeatures, outfeatures, bias, dtype=None):
        super().__init__()
        if bits != 4:
            raise NotImplementedError("only 4bit is supported by ONNXRUNTIME for now.")     # quant_linear_onnxruntime.py:114
        self._init_common(bits, groupsize, infeatures, outfeatures, dtype)
        self.pack_mode = "ORT"
        G = infeatures // self.groupsize
        self.register_buffer("qweight", torch.zeros((outfeatures, G, self.groupsize // 2), dtype=torch.uint8))
        self.register_buffer("qzeros", torch.zeros((G + (G & 1)) * (outfeatures // 8 * bits), dtype=torch.uint8))
        self.register_buffer("scales", torch.zeros(math.ceil(infeatures / self.groupsize) * outfeatures, dtype=self.dtype))
        self.register_buffer("g_idx", self._default_g_idx())
        if bias:
            self.register_buffer("bias", torch.zeros((outfeatures), dtype=self.dtype))
        else:
            self.bias = None

    def pack(self, linear, scales, zeros, g_idx=None):
        dev = _pack_device()
        g = self._default_g_idx() if g_idx is None else g_idx.to(torch.int32).cpu()
        s_t = scales.t().contiguous().to(dev).float()
        z_t = zeros.t().contiguous().to(dev).float()
        q = codec.quantize_weight(linear.weight.data.t().to(dev).float(), s_t, z_t, g.to(dev), self.maxq)
        qw, qz, sc = codec.ort_pack(q, z_t.round().to(torch.int32), s_t.to(self.dtype), self.groupsize)
        self.qweight, self.qzeros, self.scales, self.g_idx = qw.cpu(), qz.cpu(), sc.cpu(), g
        if linear.bias is not None:
            self.bias = linear.bias.detach().clone().to(self.dtype).cpu()
        self._desc, self.act_order = None, None

    def unpack(self):
        if getattr(self, "_consolidated", False):
            qw, qz, sc = self._native_tensors()
        else:
            qw, qz, sc = self.qweight, self.qzeros, self.scales
        q, z, s = codec.ort_unpack(qw, qz, sc, self.infeatures, self.outfeatures, self.groupsize)
        gi = self.g_idx.long().to(q.device)
        w = ((q.float() - z.float()[gi]) * s.float()[gi]).to(torch.float16)
        return w.t().contiguous().cpu(), s.cpu(), z.cpu()


class QuantLinearMarlin(_B200QuantLinearBase):
    """pack_mode=MARLIN: symmetric int4, qweight i32 [K/16, 2N], scales fp16 [G, N] permuted."""
    _layout = LAYOUT_MARLIN

    def __init__(self, bits, group_size, infeatures, outfeatures, bias, dtype=None):
        super().__init__()
        if bits not in [4]:
            raise NotImplementedError("Only 4 bits are supported.")
        if infeatures % 128 != 0 or outfeatures % 256 != 0:
            raise ValueError("`infeatures` must be divisible by 128 and `outfeatures` by 256.")
        if group_size not in [-1, 128] and group_size != infeatures:
            raise ValueError("Only group_size -1 and 128 are supported.")
        self._init_common(bits, group_size, infeatures, outfeatures, dtype)
        self.group_size = self.groupsize
        self.pack_mode = "MARLIN"
        self.register_buffer("qweight", torch.zeros((infeatures // 16, outfeatures * 16 // 8), dtype=torch.int32))
        self.register_buffer("scales", torch.zeros((infeatures // self.group_size, outfeatures), dtype=torch.float16))
        # kept for state-dict / attribute compatibility; the engine uses its own shared scratch
        self.register_buffer("workspace", torch.zeros(outfeatures // 128 * 16, dtype=torch.int32), persistent=False)
        self.g_idx = None
        self.qzeros = None
        if bias:
            self.register_buffer("bias", torch.zeros((outfeatures), dtype=torch.float16))
        else:
            self.bias = None

    def pack(self, linear, scales, zeros, g_idx=None):
        assert zeros is None or torch.all(zeros == 8), "only symmetric quantisation is supported (z == 8)"
        dev = _pack_device()
        s_t = scales.t().contiguous().to(dev).to(torch.float16)                 # [G, N]
        gi = self._default_g_idx().long().to(dev)
        w = linear.weight.data.t().to(dev).to(torch.float16)
        q = torch.clamp(torch.round(w / s_t[gi]).to(torch.int32) + 8, 0, 15)    # fp16 division as the reference
        qw, sp = codec.marlin_pack(q, s_t, self.group_size)
        self.qweight = qw.cpu()
        self.scales = sp.cpu()
        if linear.bias is not None:
            self.bias = linear.bias.detach().clone().to(torch.float16).cpu()
        self._desc = None

    def unpack(self):
        """The reference raises NotImplementedError here (quant_linear_marlin.py:139-140)."""
        if getattr(self, "_consolidated", False):
            qw, _, sc = self._native_tensors()
        else:
            qw, sc = self.qweight, self.scales
        q, s = codec.marlin_unpack(qw, sc, self.group_size, self.infeatures)
        gi = self._default_g_idx().long().to(q.device)
        w = ((q.float() - 8.0) * s.float()[gi]).to(torch.float16)
        return w.t().contiguous().cpu(), s.cpu(), torch.full_like(s, 8, dtype=torch.int32).cpu()


NVTX = os.environ.get("B200Q_NVTX", "0") not in ("", "0")
GROUP_GEMM_MIN_M = 64      # above this the library runs sibling GEMMs with the flag release (below: plain per-layer calls)


def linear_group(layers, x):
    """[layer(x) for layer in layers] for sibling QuantLinears that share their input (q/k/v, gate/up), as one
    launch of b200q_linear_group at decode sizes (identical results; falls back per layer inside the library)."""
    K = layers[0].infeatures
    x2 = x.reshape(-1, x.shape[-1])
    M = x2.shape[0]
    if M == 0 or any(l.infeatures != K for l in layers) or (lib.b200q_gemv_max_m() < M <= GROUP_GEMM_MIN_M):
        return [_B200QuantLinearBase.forward(l, x) for l in layers]      # not l(x): that would re-enter the sibling group
    if M > lib.b200q_gemv_max_m():
        # prefill sizes: one tcgen05 GEMM per sibling, the later ones released by a flag instead of the kernel boundary
        descs = [l._decode_descriptor(8) if l._layout in _RELAYOUT else l._fast_descriptor() for l in layers]
        descs = [l._gemm_descriptor() if lib.b200q_select_kernel(ctypes.byref(d), M) != 2 else d for l, d in zip(layers, descs)]
    else:
        descs = [l._decode_descriptor(M) for l in layers]
    if x2.dtype != torch.float16:
        x2 = x2.to(torch.float16)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    n = len(layers)
    ys = [torch.empty((M, l.outfeatures), dtype=torch.float16, device=x.device) for l in layers]
    LP = ctypes.POINTER(Layer)
    arr = (LP * n)(*[ctypes.pointer(d) for d in descs])
    yp = (ctypes.c_void_p * n)(*[y.data_ptr() for y in ys])
    ld = (ctypes.c_int64 * n)(*[y.stride(0) for y in ys])
    need = max(4096, max(lib.b200q_workspace_bytes(ctypes.byref(d), M) for d in descs))
    ws = _workspace(x.device, need)
    check(lib.b200q_linear_group(arr, n, x2.data_ptr(), M, x2.stride(0), yp, ld, ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream(x.device).cuda_stream), "b200q_linear_group")
    outs = []
    for l, y in zip(layers, ys):
        if y.dtype != x.dtype:
            y = y.to(x.dtype)
        outs.append(y.reshape(x.shape[:-1] + (l.outfeatures,)))
    return outs


def fused_mlp(gate_proj, up_proj, down_proj, x, residual=None):
    """LlamaMLP.forward -- down_proj(silu(gate_proj(x)) * up_proj(x)) (+ residual) -- in two engine calls at decode sizes:
    gate|up as one sibling-group launch, then down_proj with the activation folded into its x stage and the skip
    connection into its epilogue (no element-wise kernels, no extra activation round trips)."""
    g, u = linear_group([gate_proj, up_proj], x)
    return down_proj.forward_fused(g, x_mul=u, residual=residual)


class _SiblingGroup:
    """Lazy fusion of sibling QuantLinears behind unmodified callers (HF LlamaAttention calls q_proj(x), k_proj(x),
    v_proj(x) one after the other on the same tensor): the first sibling's forward computes all of them in one
    b200q_linear_group launch and parks the others' results, which are handed out when their forward is called
    with the same tensor (checked by identity and version counter); any other call takes the plain path."""

    def __init__(self, layers):
        self.layers = list(layers)
        self.key, self.parked = None, {}

    def forward_of(self, layer, x):
        key = (id(x), x._version, x.data_ptr(), tuple(x.shape))
        if self.key == key and id(layer) in self.parked:
            return self.parked.pop(id(layer))
        rows = x.reshape(-1, x.shape[-1]).shape[0]
        if layer is not self.layers[0] or rows == 0 or lib.b200q_gemv_max_m() < rows <= GROUP_GEMM_MIN_M or not x.is_cuda:
            return _B200QuantLinearBase.forward(layer, x)
        outs = linear_group(self.layers, x)
        self.key = key
        self._x = x                      # keeps id(x) from being recycled while results are parked
        self.parked = {id(l): y for l, y in zip(self.layers[1:], outs[1:])}
        return outs[0]


SIBLING_SETS = (("q_proj", "k_proj", "v_proj"), ("gate_proj", "up_proj"), ("w1", "w3"))


def fuse_siblings(model, sibling_sets=SIBLING_SETS):
    """Install lazy sibling fusion on every parent module that holds a full sibling set of b200q QuantLinears with
    the same (class, bits, groupsize, infeatures).  Returns the number of groups installed.  The model's own
    forward code is untouched (drop-in); call once after the checkpoint is loaded."""
    n = 0
    for parent in model.modules():
        for names in sibling_sets:
            subs = [getattr(parent, nm, None) for nm in names]
            if not all(isinstance(m, _B200QuantLinearBase) for m in subs):
                continue
            a = subs[0]
            if any(type(m) is not type(a) or m.bits != a.bits or m.groupsize != a.groupsize or m.infeatures != a.infeatures
                   or m._detect_act_order() for m in subs):      # act_order itself is only filled in by the first forward
                continue
            grp = _SiblingGroup(subs)
            for m in subs:
                m._sibling_group = grp
            n += 1
    return n


def select_quant_linear(pack_mode: str, wbits: int, quant_method: str):
    """Same decision table as the reference (utils/modelutils.py:44-68), minus the back-ends that are
    out of scope (VPTQ, ORT); AUTO prefers the AWQ layout for 4-bit as the reference does on sm>=75."""
    pack_mode = pack_mode.upper()
    quant_method = quant_method.lower()
    if quant_method == "vptq":
        raise NotImplementedError(f"quant_method={quant_method} is outside the b200q hot path")
    if pack_mode == "ORT":
        return QuantLinearORT
    if quant_method == "hqq":
        return QuantLinearHQQ
    if pack_mode == "MARLIN":
        return QuantLinearMarlin
    if pack_mode == "GEMV":                      # not reachable in the reference's table; offered for checkpoints in that layout
        return WQLinear_GEMV
    if pack_mode == "GEMM" or (pack_mode == "AUTO" and wbits == 4):
        return WQLinear_GEMM
    return QuantLinearGPTQ


def _set_op_by_name(layer, name, new_module):
    levels = name.split(".")
    mod = layer
    for lv in levels[:-1]:
        mod = mod[int(lv)] if lv.isdigit() else getattr(mod, lv)
    setattr(mod, levels[-1], new_module)


def make_mixbits_quant_linear(module, replaced_names, quant_info: dict, name="", target_layer=None):
    """Swap the named nn.Linear modules for `target_layer` instances (utils/modelutils.py:161-182);
    per-layer (wbits, groupsize) when `quant_info` is keyed by layer name."""
    dtype = next(iter(module.parameters())).dtype
    for module_name, sub in list(module.named_modules()):
        if module_name not in replaced_names:
            continue
        if "groupsize" in quant_info and "wbits" in quant_info:
            bits, groupsize = quant_info["wbits"], quant_info["groupsize"]
        else:
            bits, groupsize = quant_info[module_name]["wbits"], quant_info[module_name]["groupsize"]
        new = target_layer(bits, groupsize, sub.in_features, sub.out_features, sub.bias is not None, dtype=dtype)
        new.bias = sub.bias.data if sub.bias is not None else None
        _set_op_by_name(module, module_name, new)
