#!/usr/bin/env python
"""bench.py -- driver contract benchmark for the b200q QuantLinear hot path.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): one decode step
(batch 1) of a random-init Llama-2-7B, int4 g128, AWQ pack_mode=GEMM -- the 224 QuantLinear calls of
the 32 decoder blocks (q,k,v,o 4096->4096; gate,up 4096->11008; down 11008->4096), chained by the data
dependencies of LlamaDecoderLayer (q,k,v read the block input; o reads v; gate,up read o; down reads gate), each
going through the C ABI exactly as QuantLinear.forward does: siblings that share their input (q/k/v, gate/up) as one
b200q_linear_group launch (the host shim's fuse_siblings), the others as b200q_linear -- 128 launches, 224 layers.
A "step" = one token through all 224 layers.  3.4 GB of distinct packed weights: far larger than L2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200q|reference] [--config decode7b|prefill7b|act13b|mixtral]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1: column-sharded)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, INTER, BLOCKS, GROUP, BITS = 4096, 11008, 32, 128, 4
METRIC = "llama2_7b_int4_g128_quantlinear_decode_tokens_per_s"
SHAPES = [("q", HIDDEN, HIDDEN), ("k", HIDDEN, HIDDEN), ("v", HIDDEN, HIDDEN), ("o", HIDDEN, HIDDEN),
          ("gate", HIDDEN, INTER), ("up", HIDDEN, INTER), ("down", INTER, HIDDEN)]
LAYOUT, WORKLOAD = "GEMM", None

# BASELINE.json configs: [1] = decode7b (the configuration the metric is quoted on, default), [2] = prefill7b,
# [3] = act13b, [4] = mixtral.  configs[0] (CPU plumbing) is the reference arm.
CONFIGS = {
    "decode7b": dict(hidden=4096, inter=11008, blocks=32, layout="GEMM", metric=METRIC),
    "prefill7b": dict(hidden=4096, inter=11008, blocks=32, layout="GPTQ", metric="llama2_7b_int4_g128_gptq_quantlinear_prefill_m512_tflops"),
    "act13b": dict(hidden=5120, inter=13824, blocks=40, layout="GPTQ_ACT",
                   metric="llama2_13b_int4_g128_gptq_actorder_quantlinear_decode_tokens_per_s"),
    "mixtral": dict(hidden=4096, inter=14336, blocks=1, layout="HQQ", metric="mixtral_8x7b_expert_ffn_hqq_decode_expert_tokens_per_s"),
}


def select_config(name):
    global HIDDEN, INTER, BLOCKS, METRIC, SHAPES, LAYOUT
    c = CONFIGS[name]
    HIDDEN, INTER, BLOCKS, METRIC, LAYOUT = c["hidden"], c["inter"], c["blocks"], c["metric"], c["layout"]
    SHAPES = [("q", HIDDEN, HIDDEN), ("k", HIDDEN, HIDDEN), ("v", HIDDEN, HIDDEN), ("o", HIDDEN, HIDDEN),
              ("gate", HIDDEN, INTER), ("up", HIDDEN, INTER), ("down", INTER, HIDDEN)]


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        p.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
        p["source"] = "measured"
    except Exception:
        pass
    return p


def alg_bytes(K, N, M, layout="GEMM"):
    G = K // GROUP
    z = 0 if layout == "MARLIN" else G * N * BITS // 8
    return K * N * BITS // 8 + G * N * 2 + z + M * K * 2 + M * N * 2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path (torch dequant + matmul,
    restated in oracle/cpu_baseline.py because /root/reference does not travel; pinned to the reference's
    own forward outputs by tests/test_oracle_golden.py), on ALL host threads (torchrun exports OMP_NUM_THREADS=1).
    A timed step is the whole token: the 7 QuantLinears of a decoder block x 32 blocks (the blocks reuse one
    block's packed buffers -- CPU time does not depend on the values); warm-up steps run one block each."""
    if rank != 0:
        return
    import torch
    from oracle import cpu_baseline as C
    torch.set_num_threads(os.cpu_count() or 1)
    steps, warm = max(1, args.steps), max(1, args.warmup)
    layers = C.make_block(HIDDEN, INTER, GROUP, BITS, torch.float16)
    h = torch.randn(1, HIDDEN).to(torch.float16)

    def block(x):
        fw = lambda i, x: C.quant_linear_forward(x, layers[i][2], layers[i][3], layers[i][4], GROUP, BITS, layers[i][0])
        q, k, v = fw(0, x), fw(1, x), fw(2, x)
        o = fw(3, v)
        gt, up = fw(4, o), fw(5, o)
        return fw(6, gt)

    for _ in range(warm):
        block(h)
    # bound the run: whole tokens while they fit ~150 s, else fewer blocks per step (reported)
    t0 = time.perf_counter()
    block(h)
    t_block = time.perf_counter() - t0
    blocks_per_step = BLOCKS if t_block * BLOCKS * steps <= 150.0 else max(1, int(150.0 / (t_block * steps)))
    t0 = time.perf_counter()
    for _ in range(steps):
        x = h
        for _ in range(blocks_per_step):
            x = block(x)
    dt = (time.perf_counter() - t0) / steps
    tok_s = blocks_per_step / (dt * BLOCKS)
    cores = torch.get_num_threads()
    ref_cuda = None
    try:                                               # labelled extra, not the arm's value: the reference's own CUDA kernel on this GPU
        ref_cuda = reference_cuda_chain()
    except Exception as e:
        ref_cuda = {"error": f"{type(e).__name__}: {e}"[:200]}
    sample = (f"{blocks_per_step} of {BLOCKS} decoder blocks ({7 * blocks_per_step} QuantLinears, M=1, fp16) timed per step"
              + ("" if blocks_per_step == BLOCKS else f", x{BLOCKS / blocks_per_step:.2f} extrapolated"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": "Llama-2-7B int4 g128 QuantLinear decode, M=1 (reference torch-CPU dequant+matmul, GPTQ layout)",
                   "layers": 7 * BLOCKS, "blocks_timed_per_step": blocks_per_step, "extrapolated": blocks_per_step != BLOCKS},
        "cpu_baseline": {"value": tok_s, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_cuda_kernels": ref_cuda,
    }), flush=True)


def reference_cuda_chain(steps=5):
    """Extra beside the CPU arm: the same 224-layer decode chain through the reference's OWN CUDA kernel --
    awq_inference_engine.gemm_forward_cuda (quant_linear_awq.py:142-148, split_k_iters = 8) compiled unmodified for sm_100 into
    oracle/_ref -- on random AWQ-format buffers, one CUDA graph, CUDA events.  None when there is no GPU or no oracle/_ref."""
    import glob
    import torch
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not torch.cuda.is_available() or not glob.glob(os.path.join(ref, "awq_inference_engine*.so")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import awq_inference_engine as awq
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    ri = lambda *sh: torch.randint(-2 ** 31, 2 ** 31 - 1, sh, dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    blocks = [{name: (ri(K, N // 8), ((torch.rand(K // GROUP, N, device=dev, generator=g) * 0.4 + 0.8) / (6.5 * K ** 0.5)).to(torch.float16),
                      ri(K // GROUP, N // 8)) for name, K, N in SHAPES} for _ in range(BLOCKS)]
    x0 = torch.randn(1, HIDDEN, device=dev, dtype=torch.float16)
    f = lambda x, t: awq.gemm_forward_cuda(x, t[0], t[1], t[2], 8)

    def token():
        x = x0
        for b in blocks:
            q, k, v = f(x, b["q"]), f(x, b["k"]), f(x, b["v"])
            o = f(v, b["o"])
            gt, up = f(o, b["gate"]), f(o, b["up"])
            x = f(gt, b["down"])
        return x

    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        token()
        s.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            out = token()
        for _ in range(2):
            graph.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        s.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": 1e3 / ms, "unit": "tokens/s", "ms_per_step": ms, "kernel": "awq_inference_engine.gemm_forward_cuda (reference csrc/awq_cuda, sm_100 build)",
            "layers": len(SHAPES) * BLOCKS, "outputs_finite": bool(torch.isfinite(out.float()).all().item()),
            "what": "the reference's own CUDA AWQ kernel over the same chain on this GPU (CUDA graph, inputs resident); the arm's value stays the CPU path"}


# ---------------------------------------------------------------------------------------------
def shard_cols(N, world, rank, gran=32):
    """Contiguous column shard [c0, c1) of width multiple of `gran` (equal when N % (world*gran) == 0)."""
    tiles = N // gran
    t0, t1 = tiles * rank // world, tiles * (rank + 1) // world
    return t0 * gran, t1 * gran


def build_model(dev, rank, world, layout="GEMM"):
    import torch
    import qllm_b200
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    ri = lambda *s: torch.randint(-2 ** 31, 2 ** 31 - 1, s, dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    blocks = []
    for _ in range(BLOCKS):
        layers = {}
        for name, K, N in SHAPES:
            c0, c1 = shard_cols(N, world, rank)
            n = c1 - c0
            G = K // GROUP
            if layout == "GEMM":
                l = qllm_b200.WQLinear_GEMM(BITS, GROUP, K, n, False, dtype=torch.float16)
                l.qweight, l.qzeros = ri(K, n // 8), ri(G, n // 8)
            else:
                l = qllm_b200.QuantLinearGPTQ(BITS, GROUP, K, n, False, dtype=torch.float16)
                l.qweight, l.qzeros = ri(K * BITS // 32, n), ri(G, n * BITS // 32)
                l.g_idx = l.g_idx.to(dev)
                if layout == "GPTQ_ACT":                  # desc_act checkpoint: every group's rows scattered over K (same map on every rank)
                    gp = torch.Generator(device=dev).manual_seed(99 + K)
                    l.g_idx = l.g_idx[torch.randperm(K, device=dev, generator=gp)].contiguous()
            l.scales = ((torch.rand(G, n, device=dev, generator=g) * 0.4 + 0.8) / (6.5 * K ** 0.5)).to(torch.float16)
            l = l.to(dev)
            l.col0, l.full_n = c0, N
            layers[name] = l
        blocks.append(layers)
    return blocks


class DecodeStep:
    """The 224 chained QuantLinear calls of one token, issued through the C ABI."""

    def __init__(self, blocks, dev, M, rank, world):
        import torch
        import qllm_b200
        self.lib, self.blocks, self.M, self.world, self.rank = qllm_b200.lib, blocks, M, world, rank
        self.fuse = os.environ.get("B200Q_BENCH_NO_GROUP") is None
        f16 = dict(dtype=torch.float16, device=dev)
        self.h = torch.zeros(M, HIDDEN, **f16)
        self.bufs = {n: torch.zeros(M, N, **f16) for n, _, N in SHAPES}
        self.part = {n: torch.zeros(M, max(shard_cols(N, world, r)[1] - shard_cols(N, world, r)[0] for r in range(world)), **f16)
                     for n, _, N in SHAPES} if world > 1 else None
        need = max(self.lib.b200q_workspace_bytes(ctypes.byref(l._decode_descriptor(M)), M) for l in blocks[0].values())
        self.ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=dev)
        self.torch = torch

    def _call(self, layer, x, name, stream):
        from qllm_b200 import check
        desc = layer._decode_descriptor(self.M)
        if self.world == 1:
            y = self.bufs[name]
            check(self.lib.b200q_linear(ctypes.byref(desc), x.data_ptr(), self.M, x.stride(0), y.data_ptr(), y.stride(0),
                                        self.ws.data_ptr(), self.ws.numel(), stream))
            return y
        # column shard -> local slice, then one all-gather of the output (equal shards)
        import torch.distributed as dist
        part = self.part[name]
        check(self.lib.b200q_linear(ctypes.byref(desc), x.data_ptr(), self.M, x.stride(0), part.data_ptr(), part.stride(0),
                                    self.ws.data_ptr(), self.ws.numel(), stream))
        y = self.bufs[name]
        if self.M == 1 and part.shape[1] * self.world == y.shape[1]:
            dist.all_gather_into_tensor(y.view(-1), part.view(-1))
        else:
            gathered = [self.torch.empty_like(part) for _ in range(self.world)]
            dist.all_gather(gathered, part)
            off = 0
            for r, t in enumerate(gathered):
                c0, c1 = shard_cols(y.shape[1], self.world, r)
                y[:, c0:c1].copy_(t[:, : c1 - c0])
        return y

    def _group(self, layers, x, names, stream):
        """Sibling layers sharing x (q/k/v, gate/up): one b200q_linear_group launch on one GPU; per-layer calls
        (each followed by its all-gather) when column-sharded."""
        if self.world > 1 or not self.fuse:
            return [self._call(l, x, n, stream) for l, n in zip(layers, names)]
        from qllm_b200 import check, Layer
        n = len(layers)
        descs = [l._decode_descriptor(self.M) for l in layers]
        ys = [self.bufs[nm] for nm in names]
        arr = (ctypes.POINTER(Layer) * n)(*[ctypes.pointer(d) for d in descs])
        yp = (ctypes.c_void_p * n)(*[y.data_ptr() for y in ys])
        ld = (ctypes.c_int64 * n)(*[y.stride(0) for y in ys])
        check(self.lib.b200q_linear_group(arr, n, x.data_ptr(), self.M, x.stride(0), yp, ld, self.ws.data_ptr(),
                                          self.ws.numel(), stream))
        return ys

    def run(self, stream):
        h = self.h
        for b in self.blocks:
            q, k, v = self._group([b["q"], b["k"], b["v"]], h, ["q", "k", "v"], stream)
            o = self._call(b["o"], v, "o", stream)
            gt, up = self._group([b["gate"], b["up"]], o, ["gate", "up"], stream)
            h = self._call(b["down"], gt, "down", stream)
        return h


class ChainDecodeStep:
    """The same 224 QuantLinear calls as DecodeStep, recorded once as b200q decode chains (b200q_chain_plan) and replayed
    with b200q_chain_run: one persistent launch per `span` decoder blocks (span = 32: the whole token is one launch).
    Every layer's fp16 output is still written to its buffer, exactly as the per-call path does."""

    def __init__(self, blocks, dev, M, span):
        import torch
        import qllm_b200
        self.M, self.fuse = M, True
        f16 = dict(dtype=torch.float16, device=dev)
        self.h = torch.zeros(M, HIDDEN, **f16)
        # per-block output buffers (a real model's activations are distinct tensors per block as well)
        self.chains, steps, x = [], [], self.h
        for i, b in enumerate(blocks):
            y = {n: torch.zeros(M, N, **f16) for n, _, N in SHAPES}
            steps.append(([b["q"], b["k"], b["v"]], x, [y["q"], y["k"], y["v"]]))
            steps.append(([b["o"]], y["v"], [y["o"]]))
            steps.append(([b["gate"], b["up"]], y["o"], [y["gate"], y["up"]]))
            steps.append(([b["down"]], y["gate"], [y["down"]]))
            x = y["down"]
            if (i + 1) % span == 0 or i + 1 == len(blocks):
                self.chains.append(qllm_b200.DecodeChain(steps, M=M))
                steps = []
        self.out = x

    def run(self, stream):
        for c in self.chains:
            c.run(stream)
        return self.out


def handoff_schedule(calls, x_of, n_blocks):
    """Call sequence of the fused sharded step: [(block, names, x_name, y_seq, x_seq)], tag_stride.  y_seq numbers the calls of
    a token from 1; x_name is None for the first block's input (plain fp16 h), else the buffer the call reads; x_seq is the
    y_seq of the call that wrote that buffer last -- what the consumer's awaited tag is built from."""
    nc = len(calls)
    seq_of, block_in, out = {}, None, []
    for i in range(n_blocks):
        for j, names in enumerate(calls):
            slot = 1 + nc * i + j
            x_name = x_of[j] if x_of[j] is not None else block_in
            out.append((i, names, x_name, slot, seq_of.get(x_name, 0)))
            for n in names:
                seq_of[n] = slot
        block_in = "down"
    return out, nc * n_blocks + 1


class FusedShardedStep:
    """N > 1: every call is one b200q_linear_group_sharded launch -- this rank's column shards, stored into every
    rank's replica over NVLink, with the cross-GPU hand-off inside the kernels: tagged activations (flag-in-data:
    a consumer lane spins on exactly the words it needs; no fence, atomic or barrier), or -- B200Q_BENCH_COUNTERS=1 --
    the counter protocol (post / wait).  No collective and no host work between layers; the whole token (epoch advance
    + 128 launches + the final untag / wait) is one CUDA graph."""
    CALLS = (("q", "k", "v"), ("o",), ("gate", "up"), ("down",))
    X_OF = (None, "v", "o", "gate")          # input of call j (None: the block input = previous block's down / h)
    # act-order layers carry their own x_perm, so siblings do not share a launch: one call per layer
    CALLS_ACT = (("q",), ("k",), ("v",), ("o",), ("gate",), ("up",), ("down",))
    X_OF_ACT = (None, None, None, "v", "o", "o", "gate")

    def __init__(self, blocks, dev, M, rank, world):
        import torch
        import torch.distributed as dist
        import qllm_b200
        from qllm_b200.sharding import LocalArena, PeerArena, sharded_group_posts
        self.lib, self.blocks, self.M, self.world, self.rank, self.torch = qllm_b200.lib, blocks, M, world, rank, torch
        if LAYOUT == "GPTQ_ACT":
            self.CALLS, self.X_OF = self.CALLS_ACT, self.X_OF_ACT
        self.tagged = os.environ.get("B200Q_BENCH_COUNTERS") is None or world == 1
        esz = 4 if self.tagged else 2
        full = {n: N for n, _, N in SHAPES}
        payload = sum(((M * N * esz + 255) & ~255) for N in full.values())
        self.arena = PeerArena(payload, n_slots=8 + 7 * len(blocks)) if world > 1 else LocalArena(payload, n_slots=8 + 7 * len(blocks), device=dev)
        self.off = {n: self.arena.carve(M * N * esz) for n, N in full.items()}
        self.bufs = {n: self.arena.local_view(self.off[n], (M, N), torch.int32 if self.tagged else torch.float16) for n, N in full.items()}
        self.full = full
        self.h = torch.zeros(M, HIDDEN, dtype=torch.float16, device=dev)
        self.out = torch.zeros(M, HIDDEN, dtype=torch.float16, device=dev)
        mine = [sharded_group_posts([blocks[0][n] for n in names], M) for names in self.CALLS]
        every = [mine]
        if world > 1:
            every = [None] * world
            dist.all_gather_object(every, mine)
        self.wait_counts = [sum(every[r][j] for r in range(world) if r != rank) for j in range(len(self.CALLS))]
        need = max(self.lib.b200q_workspace_bytes(ctypes.byref(l._decode_descriptor(M)), M) for l in blocks[0].values())
        self.ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=dev)
        self.fuse = True
        # load the kernels locally (lazy module loading can take seconds) before any rank waits on a peer (2 s bound)
        for l in blocks[0].values():
            l(torch.zeros(M, l.infeatures, dtype=torch.float16, device=dev))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(self, stream):
        from qllm_b200.sharding import sharded_group_forward
        from qllm_b200._lib import PEER_NODE_EPOCH, PEER_X_TAGGED, PEER_Y_TAGGED
        node = PEER_NODE_EPOCH if os.environ.get("B200Q_BENCH_NO_NODE_EPOCH") is None else 0
        A, M = self.arena, self.M
        A.advance(stream)
        x_name, wait_slot, wait_count = None, -1, 0
        nc = len(self.CALLS)
        schedule, stride = handoff_schedule(self.CALLS, self.X_OF, len(self.blocks))
        for i, names, x_name, slot, x_seq in schedule:
            b, j = self.blocks[i], (slot - 1) % nc
            layers = [b[n] for n in names]
            offs, fulls, col0s = [self.off[n] for n in names], [self.full[n] for n in names], [l.col0 for l in layers]
            if self.tagged:
                flags = PEER_Y_TAGGED | (PEER_X_TAGGED if x_name is not None else 0) | node
                sync = A.sync_desc(flags=flags, tag_stride=stride, y_seq=slot, x_seq=x_seq)
            else:
                sync = A.sync_desc(wait_slot, wait_count, slot)
            if x_name is None:
                sharded_group_forward(A, layers, self.h, offs, fulls, col0s, sync, self.ws, stream)
            else:
                xb = self.bufs[x_name]
                sharded_group_forward(A, layers, None, offs, fulls, col0s, sync, self.ws, stream,
                                      x_ptr=xb.data_ptr(), M=M, ldx=xb.stride(0))
            wait_slot, wait_count = slot, self.wait_counts[j]
        # the host (or lm_head) reads the complete hidden state as plain fp16
        if self.tagged:
            A.untag(self.off["down"], M, HIDDEN, self.out, stride, schedule[-1][3], stream)
        else:
            A.wait(wait_slot, wait_count, stream)
            self.out.copy_(self.bufs["down"])
        return self.out


def run_b200q(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import qllm_b200
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    P = peaks()
    steps, warm = max(1, args.steps), max(3, args.warmup)
    blocks = build_model(dev, rank, world, LAYOUT)
    M = 1
    fused_err = None
    step = None
    # act-order layers can read tagged activations (gathered word by word through x_perm) but that load stage is slow
    # (N = 1: 140 vs 368 tokens/s): their N > 1 default stays the NCCL all-gather per layer, B200Q_BENCH_ACT_FUSED=1 selects the fused path
    if world > 1 and os.environ.get("B200Q_BENCH_NCCL") is None and (LAYOUT != "GPTQ_ACT" or os.environ.get("B200Q_BENCH_ACT_FUSED")):
        try:
            step = FusedShardedStep(blocks, dev, M, rank, world)
        except Exception as e:                            # symmetric memory unavailable: NCCL all-gather per layer
            fused_err = f"{type(e).__name__}: {e}"[:200]
        ok = torch.tensor([0 if step is None else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            step = None
    fused_sharded = step is not None
    tagged_n1 = False
    if world == 1 and args.handoff == "tagged" and args.chain == 0:
        # single GPU, same kernels as the N > 1 path with one peer: activations travel between the QuantLinears as tagged words
        # and a consumer's load stage follows the data instead of the kernel boundary (B200Q_PEER_NODE_EPOCH)
        step = FusedShardedStep(blocks, dev, M, 0, 1)
        tagged_n1 = True
    chain_span = 0
    if step is None and world == 1 and args.chain > 0:
        step = ChainDecodeStep(blocks, dev, M, args.chain)
        chain_span = args.chain
    if step is None:
        step = DecodeStep(blocks, dev, M, rank, world)
    h0 = (torch.randn(M, HIDDEN, generator=torch.Generator().manual_seed(7)) * 1.0).to(torch.float16)
    h0_pinned = h0.pin_memory()
    y_pinned = torch.empty(M, HIDDEN, dtype=torch.float16).pin_memory()
    step.h.copy_(h0)

    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        out = step.run(s.cuda_stream)                 # eager warm-up (sets kernel attributes)
    s.synchronize()
    n0 = qllm_b200.lib.b200q_launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        out = step.run(torch.cuda.current_stream().cuda_stream)
    launches_per_step = qllm_b200.lib.b200q_launch_count() - n0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        # W warm-up steps, then keep the GPU under the same load until nvidia-smi's first sample (~0.2 s) has been taken.
        # A FIXED count: every rank must replay the same number of steps (the fused hand-off pairs them up).
        for _ in range(warm + 300):
            graph.replay()
        barrier()
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        # ---- e2e: host pinned x -> device -> 224 C-ABI calls (graph) -> host, every step ----
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step.h.copy_(h0_pinned, non_blocking=True)
            graph.replay()
            y_pinned.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()
    ms_per_step = ms / steps
    tok_s = 1e3 / ms_per_step
    finite = bool(torch.isfinite(out.float()).all().item())
    replicas_equal, timeouts = None, None
    matches_plain = None
    if tagged_n1:                                         # the tagged chain against the kernel-boundary chain on the same weights
        plain = DecodeStep(blocks, dev, M, rank, world)
        plain.h.copy_(step.h)
        with torch.cuda.stream(s):
            ref_out = plain.run(s.cuda_stream)
        s.synchronize()
        matches_plain = bool(torch.equal(ref_out, out))
        timeouts = bool(step.arena.poisoned())
    if world > 1:
        allout = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(allout, out.contiguous())
        replicas_equal = all(bool(torch.equal(allout[0], t)) for t in allout[1:])
        if fused_sharded:
            timeouts = bool(step.arena.poisoned())

    if rank != 0:
        return
    total_bytes = BLOCKS * sum(alg_bytes(K, N, M) + (K * 4 if LAYOUT == "GPTQ_ACT" else 0) for _, K, N in SHAPES)
    n_layers = BLOCKS * len(SHAPES)
    achieved = total_bytes / (ms_per_step * 1e-3) / 1e9                      # whole-job algorithmic GB/s
    roofline = {"bound": "hbm", "achieved": achieved / world, "peak": P["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / world / P["hbm_gbs"], "traffic": None, "peak_source": P["source"],
                "kernel": ("decode_chain_kernel (persistent, TMA ring + IMMA.16832)" if chain_span else "gemv_imma_kernel (decode, IMMA.16832 on the K-packed re-layout)"), "bytes_per_launch": total_bytes / launches_per_step,
                "us_per_launch": ms_per_step * 1e3 / launches_per_step, "per_gpu": True}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        roofline["traffic"] = tr.get("decode_awq_dram_bytes_per_launch")
    except Exception:
        pass
    result = {
        "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": ("Llama-2-7B int4 g128 AWQ pack_mode=GEMM" if LAYOUT == "GEMM" else "Llama-2-13B int4 g128 GPTQ act_order (g_idx gather)")
                               + f", batch=1 decode: the {n_layers} QuantLinear layers of one token "
                               "in LlamaDecoderLayer dependency order (q|k|v -> o -> gate|up -> down)",
                   "launches_per_step": int(launches_per_step), "chain_blocks_per_launch": chain_span, "sibling_groups": bool(step.fuse and world == 1),
                   "M": M, "layers": n_layers, "parallelism": (f"column-shard x{world}, all-gather + hand-off fused into the decode kernels (NVLink peer stores, "
                                    + ("tagged activations" if getattr(step, "tagged", False) else "counter post/wait") + ")" if fused_sharded
                                   else f"column-shard x{world} + NCCL all-gather per layer") if world > 1 else "single GPU",
                   "handoff": ("tagged activations between the QuantLinears (fp16 | step tag << 16): a consumer's load stage follows the data, "
                               "no kernel-boundary wait ahead of x" if (tagged_n1 or (fused_sharded and getattr(step, "tagged", False)))
                               else "kernel boundary (programmatic dependent launch)"),
                   "matches_kernel_boundary_path": matches_plain,
                   "fused_sharded_error": fused_err, "replicas_equal": replicas_equal, "peer_wait_timeouts": timeouts,
                   "l2_policy": "inputs (3.4 GB packed weights) larger than L2", "cuda_graph": True, "pdl": True,
                   "outputs_finite": finite},
        "clocks": clk.summary(),
        "e2e": {"value": steps / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": M * HIDDEN * 2, "d2h_bytes_per_step": M * HIDDEN * 2},
        "gpu_launches": int(launches_per_step * steps),
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu:
        from oracle import cpu_baseline as C
        torch.set_num_threads(os.cpu_count() or 1)
        t_block = C.time_block(M=1, repeats=3, hidden=HIDDEN, inter=INTER)
        result["cpu_baseline"] = {"value": 1.0 / (t_block * BLOCKS), "unit": "tokens/s", "cores": torch.get_num_threads(),
                                  "kind": "port", "sample": f"1 of {BLOCKS} decoder blocks (7 QuantLinears, M=1, fp16 torch-CPU dequant+matmul), best of 3, x{BLOCKS}"}
    if world == 1 and not tagged_n1 and chain_span == 0 and args.config == "decode7b" and not args.no_prefill:
        # beside the headline (plain fp16 between the layers, what an unmodified model sees): the same chain with the N > 1 path's
        # tagged hand-off on one GPU (--handoff tagged) -- applies wherever one QuantLinear feeds the next directly
        try:
            result["tagged_handoff"] = tagged_chain_extra(blocks, dev, M, s, steps, warm, out, step.h, total_bytes, P)
        except Exception as e:
            result["tagged_handoff"] = {"error": str(e)[:200]}
    if world == 1 and not args.no_prefill and args.config == "decode7b":
        try:
            result["prefill"] = prefill_tflops(dev, P)
        except Exception as e:                       # the decode metric stands on its own
            result["prefill"] = {"error": str(e)[:200]}
    print(json.dumps(result), flush=True)


def tagged_chain_extra(blocks, dev, M, s, steps, warm, ref_out, h, total_bytes, P):
    """The decode step with tagged activations between the QuantLinears on one GPU (FusedShardedStep, world 1), timed like
    the headline; must reproduce the headline path's hidden state bit for bit."""
    import torch
    step = FusedShardedStep(blocks, dev, M, 0, 1)
    step.h.copy_(h)
    with torch.cuda.stream(s):
        step.run(s.cuda_stream)
        s.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            out = step.run(torch.cuda.current_stream().cuda_stream)
        ms = _timed(graph.replay, steps, warm + 50, torch.cuda.synchronize) 
    equal = bool(torch.equal(out, ref_out))
    return {"value": 1e3 / ms, "unit": "tokens/s", "ms_per_step": ms, "roofline_frac": total_bytes / (ms * 1e-3) / 1e9 / P["hbm_gbs"],
            "matches_headline_path": equal, "peer_wait_timeouts": bool(step.arena.poisoned()),
            "what": "bench.py --handoff tagged: activations between the QuantLinears as fp16 | step tag << 16 words, a consumer's load stage "
                    "follows the data instead of the kernel boundary (B200Q_PEER_NODE_EPOCH)"}


def prefill_tflops(dev, P, M=512, iters=8):
    """configs[2]: GPTQ layout, M=512 rows per call, the 7 shapes of one decoder block (tcgen05 GEMM)."""
    import torch
    import qllm_b200
    from tools.microbench import rand_layer
    copies = 6                                         # 6 blocks x 101 MB of packed weights > L2
    layers = [[rand_layer("GPTQ", BITS, GROUP, K, N, dev, 100 * c + i) for i, (_, K, N) in enumerate(SHAPES)] for c in range(copies)]
    xs = {K: torch.randn(M, K, dtype=torch.float16, device=dev) for K in (HIDDEN, INTER)}
    group = os.environ.get("B200Q_BENCH_NO_GROUP") is None

    def block(ls):
        """q|k|v and gate|up share their input: one b200q_linear_group call each (the later siblings' GEMMs are released by a
        flag instead of the kernel boundary); o and down are single calls.  7 GEMM launches either way."""
        if not group:
            for l in ls:
                l(xs[l.infeatures])
            return
        qllm_b200.linear_group(ls[0:3], xs[HIDDEN])
        ls[3](xs[HIDDEN])
        qllm_b200.linear_group(ls[4:6], xs[HIDDEN])
        ls[6](xs[INTER])

    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for c in range(copies):
            block(layers[c])
        s.synchronize()
        graph = torch.cuda.CUDAGraph()               # the six blocks as one graph (programmatic-launch edges), like the decode step
        with torch.cuda.graph(graph, stream=s):
            for c in range(copies):
                block(layers[c])
        for _ in range(3):
            graph.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(iters):
            graph.replay()
        e1.record()
        s.synchronize()
    flops = iters * copies * sum(2.0 * M * K * N for _, K, N in SHAPES)
    tf = flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return {"workload": "Llama-2-7B block, GPTQ int4 g128, M=512 per call (7 GEMMs)", "tflops": tf, "sibling_groups": group,
            "frac_of_bf16_peak": tf / P["bf16_tflops"], "peak": P["bf16_tflops"], "kernel": "gemm_tc_gptq4_kernel (tcgen05)"}


def _timed(fn, steps, warm, barrier):
    """W warm-up + K timed replays of fn() on the current stream, CUDA events; returns ms per step."""
    import torch
    for _ in range(warm):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def run_prefill(args, rank, world, local_rank):
    """configs[2]: Llama-2-7B GPTQ int4 g128, 512 rows per QuantLinear call (the graded prefill tile), the 224 layers of
    the model chained as in LlamaDecoderLayer, one CUDA graph; single GPU (N > 1: independent replicas, weak scaling)."""
    import torch
    import torch.distributed as dist
    import qllm_b200
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    P = peaks()
    steps, warm, M = max(1, args.steps), max(3, args.warmup), args.m if args.m > 0 else 512
    blocks = build_model(dev, 0, 1, "GPTQ")
    h = torch.zeros(M, HIDDEN, dtype=torch.float16, device=dev)
    h0 = torch.randn(M, HIDDEN, generator=torch.Generator().manual_seed(7)).to(torch.float16).pin_memory()
    y_pinned = torch.empty(M, HIDDEN, dtype=torch.float16).pin_memory()
    h.copy_(h0)

    group = os.environ.get("B200Q_BENCH_NO_GROUP") is None

    def forward():
        x = h
        for b in blocks:
            if group:          # q|k|v and gate|up share their input: b200q_linear_group (still one GEMM launch per layer)
                q, k, v = qllm_b200.linear_group([b["q"], b["k"], b["v"]], x)
                o = b["o"](v)
                gt, up = qllm_b200.linear_group([b["gate"], b["up"]], o)
            else:
                q, k, v = b["q"](x), b["k"](x), b["v"](x)
                o = b["o"](v)
                gt, up = b["gate"](o), b["up"](o)
            x = b["down"](gt)
        return x

    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        out = forward()
    s.synchronize()
    n0 = qllm_b200.lib.b200q_launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        out = forward()
    launches = qllm_b200.lib.b200q_launch_count() - n0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with ClockSampler(local_rank) as clk:
        ms = _timed(graph.replay, steps, warm, barrier)
        t0 = time.perf_counter()
        for _ in range(steps):
            h.copy_(h0, non_blocking=True)
            graph.replay()
            y_pinned.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()
    if rank != 0:
        return
    flops = 2.0 * M * BLOCKS * sum(K * N for _, K, N in SHAPES)
    tf = flops / (ms * 1e-3) / 1e12
    peak = P.get("bf16_tflops_sustained", P["bf16_tflops"])
    print(json.dumps({
        "metric": METRIC, "value": tf * world, "unit": "TFLOP/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"Llama-2-7B int4 g128 GPTQ, prefill tile M={M} rows per QuantLinear call: the 224 layers of the model "
                               "(q,k,v -> o -> gate,up -> down), one CUDA graph", "M": M, "layers": BLOCKS * len(SHAPES),
                   "parallelism": "single GPU" if world == 1 else f"{world} independent replicas",
                   "l2_policy": "inputs (3.4 GB packed weights) larger than L2", "cuda_graph": True, "sibling_groups": group,
                   "outputs_finite": bool(torch.isfinite(out.float()).all().item())},
        "clocks": clk.summary(),
        "e2e": {"value": flops / e2e_s / 1e12 * world, "unit": "TFLOP/s", "h2d_bytes_per_step": M * HIDDEN * 2, "d2h_bytes_per_step": M * HIDDEN * 2},
        "gpu_launches": int(launches * steps),
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                     "peak_source": P["source"] + " (sustained: the kernel is timed inside a long step)",
                     "kernel": "gemm_tc_gptq_kernel (tcgen05.mma kind::f16, W dequantised into TMEM, X by TMA)",
                     "flops_per_launch": flops / launches, "us_per_launch": ms * 1e3 / launches, "frac_of_burst_peak": tf / P["bf16_tflops"]},
    }), flush=True)


def run_mixtral(args, rank, world, local_rank):
    """configs[4]: Mixtral-8x7B expert FFN (w1, w3: 4096 -> 14336; w2: 14336 -> 4096), HQQ layouts.  Headline: 4-bit g64, one
    token per expert (M = 1): (token, expert) FFN evaluations per second over the 8 experts of one layer; experts are
    split across the ranks (expert parallel, no collective inside the path).  `sweep`: the other (bits, group, M)."""
    import torch
    import torch.distributed as dist
    import qllm_b200
    from tools.microbench import rand_layer
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    P = peaks()
    steps, warm = max(1, args.steps), max(3, args.warmup)
    experts = [e for e in range(8) if e % world == rank]

    def build(bits, gs):
        return [dict(w1=rand_layer("HQQ", bits, gs, HIDDEN, INTER, dev, 10 * e + 1), w3=rand_layer("HQQ", bits, gs, HIDDEN, INTER, dev, 10 * e + 2),
                     w2=rand_layer("HQQ", bits, gs, INTER, HIDDEN, dev, 10 * e + 3)) for e in experts]

    def make_graph(ex, M):
        x = torch.randn(M, HIDDEN, dtype=torch.float16, device=dev)

        def fwd():
            outs = []
            for e in ex:
                a, b = qllm_b200.linear_group([e["w1"], e["w3"]], x) if M <= 8 else (e["w1"](x), e["w3"](x))
                outs.append(e["w2"](a))                                       # silu(a) * b is outside the QuantLinear path
            return outs

        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fwd()
        s.synchronize()
        n0 = qllm_b200.lib.b200q_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            outs = fwd()
        return g, x, outs, qllm_b200.lib.b200q_launch_count() - n0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def bytes_per_expert(bits, gs, M):
        tot = 0
        for K, N in ((HIDDEN, INTER), (HIDDEN, INTER), (INTER, HIDDEN)):
            G = K // gs
            tot += K * N * bits // 8 + G * N * 2 + G * N * 2 + M * K * 2 + M * N * 2
        return tot

    ex = build(4, 64)
    g, x, outs, launches = make_graph(ex, 1)
    x_pinned = torch.randn(1, HIDDEN).to(torch.float16).pin_memory()
    y_pinned = torch.empty(1, HIDDEN, dtype=torch.float16).pin_memory()
    with ClockSampler(local_rank) as clk:
        ms = _timed(g.replay, steps, warm, barrier)
        t0 = time.perf_counter()
        for _ in range(steps):
            x.copy_(x_pinned, non_blocking=True)
            g.replay()
            y_pinned.copy_(outs[-1], non_blocking=True)
            torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / steps
    sweep = []
    if world == 1 and not args.no_sweep:
        for bits, gs in ((4, 64), (2, 64), (3, 64), (8, 128)):
            if (bits, gs) != (4, 64):
                del ex
                torch.cuda.empty_cache()
                ex = build(bits, gs)
            for M in (1, 16, 512):
                g2, _, _, _ = make_graph(ex, M)
                t = _timed(g2.replay, max(5, steps // 4), 3, barrier)
                b = bytes_per_expert(bits, gs, M) * len(ex)
                fl = 2.0 * M * 3 * HIDDEN * INTER * len(ex)
                sweep.append({"bits": bits, "group": gs, "M": M, "ms": round(t, 4), "GBps": round(b / t / 1e6, 1),
                              "hbm_frac": round(b / t / 1e6 / P["hbm_gbs"], 3), "TFLOPs": round(fl / t / 1e9, 1)})
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()
    if rank != 0:
        return
    b = bytes_per_expert(4, 64, 1) * len(ex)
    print(json.dumps({
        "metric": METRIC, "value": 8 / (ms * 1e-3), "unit": "expert-tokens/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": "Mixtral-8x7B expert FFN QuantLinears (w1|w3 4096->14336, w2 14336->4096), HQQ 4-bit g64, M=1 per expert, "
                               "the 8 experts of one layer", "experts_per_rank": len(ex), "parallelism": "single GPU" if world == 1 else f"expert parallel x{world} (no collective)",
                   "l2_policy": "inputs (8 x 99 MB packed weights) larger than L2", "cuda_graph": True},
        "clocks": clk.summary(),
        "e2e": {"value": 8 / e2e_s, "unit": "expert-tokens/s", "h2d_bytes_per_step": HIDDEN * 2, "d2h_bytes_per_step": HIDDEN * 2},
        "gpu_launches": int(launches * steps),
        "roofline": {"bound": "hbm", "achieved": b / ms / 1e6, "peak": P["hbm_gbs"], "unit": "GB/s", "frac": b / ms / 1e6 / P["hbm_gbs"],
                     "traffic": None, "peak_source": P["source"], "kernel": "gemv_imma_kernel (HQQ fp16 zeros)", "per_gpu": True,
                     "bytes_per_launch": b / launches, "us_per_launch": ms * 1e3 / launches},
        "sweep": sweep,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200q", choices=["b200q", "reference"])
    ap.add_argument("--handoff", default="kernel", choices=["kernel", "tagged"],
                    help="N = 1 decode: how consecutive QuantLinears hand activations over (kernel = plain fp16 + kernel-boundary order; "
                         "tagged = tagged words, data-flow per lane)")
    ap.add_argument("--chain", type=int, default=0,
                    help="decoder blocks per decode-chain launch (b200q_chain_run); 0 = one launch per sibling group (b200q_linear_group)")
    ap.add_argument("--config", default="decode7b", choices=sorted(CONFIGS),
                    help="BASELINE.json workload: decode7b = configs[1] (default, the configuration the metric is quoted on), "
                         "prefill7b = configs[2], act13b = configs[3], mixtral = configs[4]")
    ap.add_argument("--m", type=int, default=0, help="prefill7b: rows per QuantLinear call (default 512)")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-prefill", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    select_config(args.config)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.config == "prefill7b":
            run_prefill(args, rank, world, local_rank)
        elif args.config == "mixtral":
            run_mixtral(args, rank, world, local_rank)
        else:
            run_b200q(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
