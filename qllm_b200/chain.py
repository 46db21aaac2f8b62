"""Decode chain: a recorded run of sibling-group QuantLinear calls executed by ONE persistent launch
(include/b200q.h: b200q_chain_plan / b200q_chain_run; csrc/decode_chain.cu).

Each step has the semantics of `linear_group(layers, x)` writing into caller-owned outputs; a step whose `x` is a view
into an output of the previous step consumes it inside the kernel.  The reference issues one kernel launch per
QuantLinear.forward (quant_linear_awq.py:142-148, quant_linear_gptq.py:55-85); at batch 1 that is launch-bound.

    chain = DecodeChain([([q, k, v], h, [yq, yk, yv]), ([o], yv, [yo]), ([gate, up], yo, [yg, yu]), ([down], yg, [yd])], M=1)
    chain.run()          # enqueues on the current stream; CUDA-graph capturable
"""
import ctypes

import torch

from ._lib import ChainStep, Layer, check, lib


class DecodeChain:
    def __init__(self, steps, M=1):
        if not steps:
            raise ValueError("empty chain")
        self.M = int(M)
        self._keep = []                               # ctypes arrays / descriptors / tensors referenced by the plan
        n = len(steps)
        arr = (ChainStep * n)()
        dev = None
        for i, (layers, x, ys) in enumerate(steps):
            if len(layers) != len(ys) or not layers:
                raise ValueError("each step needs one output per layer")
            if x.dtype != torch.float16 or any(y.dtype != torch.float16 for y in ys):
                raise ValueError("chain activations and outputs are fp16")
            if x.dim() != 2 or x.shape[0] != self.M or x.stride(1) != 1:
                raise ValueError("x must be [M, K] with unit column stride")
            dev = x.device
            descs = [l._decode_descriptor(self.M) for l in layers]
            lp = (ctypes.POINTER(Layer) * len(layers))(*[ctypes.pointer(d) for d in descs])
            yp = (ctypes.c_void_p * len(layers))(*[y.data_ptr() for y in ys])
            ld = (ctypes.c_int64 * len(layers))(*[y.stride(0) for y in ys])
            self._keep += [descs, lp, yp, ld, x, ys, layers]
            arr[i].layers, arr[i].n_layers = lp, len(layers)
            arr[i].x, arr[i].ldx = x.data_ptr(), x.stride(0)
            arr[i].y, arr[i].ldy = yp, ld
        nbytes = lib.b200q_chain_plan_bytes(arr, n, self.M)
        if nbytes == 0:
            raise ValueError("these layers cannot run as a decode chain (see b200q_chain_plan); use linear_group per step")
        self.plan_host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=False)
        ws = ctypes.c_size_t(0)
        with torch.cuda.device(dev):
            check(lib.b200q_chain_plan(arr, n, self.M, self.plan_host.data_ptr(), nbytes, ctypes.byref(ws)), "b200q_chain_plan")
        self.plan_dev = self.plan_host.to(dev)
        self.workspace = torch.zeros(max(ws.value, 4096), dtype=torch.uint8, device=dev)
        self.device, self.n_steps = dev, n
        torch.cuda.current_stream(dev).synchronize()

    def run(self, stream=None):
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        check(lib.b200q_chain_run(self.plan_host.data_ptr(), self.plan_dev.data_ptr(), self.workspace.data_ptr(),
                                  self.workspace.numel(), s), "b200q_chain_run")

    def error_code(self):
        """Non-zero when a wait inside the kernel timed out (the kernel then traps): word 514 of the counter region."""
        return int(self.workspace[:4096].view(torch.int32)[514].item())
