/*
 * b200q.h -- C ABI of the B200-native fused dequant-matmul engine (libb200q.so).
 *
 * This is the drop-in boundary behind QLLM's QuantLinear.forward.  Every entry point below
 * replaces one (or several) of the reference's native entry points for the hot path; the
 * reference interface each one stands in for is cited as file:line relative to the reference
 * tree (wejoncy/QLLM @ df20c15).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions (all entry points):
 *   - plain C types only; every pointer is a CUDA *device* pointer unless stated otherwise;
 *   - the engine never allocates, frees or synchronises; work is enqueued on `stream`
 *     (the reference enqueues on at::cuda::getCurrentCUDAStream(): gemm_cuda_gen.cu:1126);
 *   - return value: B200Q_OK (0) or a negative b200q_status code, never abort()
 *     (the reference aborts on unsupported bits / launch failure: dq_gemv.cu:156-159,:172-176);
 *   - all calls are CUDA-graph capturable (no host-side data-dependent control flow);
 *   - activations and outputs are IEEE fp16, row-major, with explicit row strides.
 *
 * Logical math for every layout:
 *     W[k,n] = scales[g(k),n] * (q[k,n] - z[g(k),n]),   y[m,:] = x[m,:] @ W (+ bias)
 * with g(k) = g_idx[k] when g_idx != NULL, else k / group_size.
 */
#ifndef B200Q_H_
#define B200Q_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200Q_VERSION 100 /* 0.1.0 */

/* Opaque CUDA stream handle (cudaStream_t) passed as a pointer-sized integer. */
typedef void* b200q_stream_t;

typedef enum b200q_status {
  B200Q_OK = 0,
  B200Q_ERR_NULL = -1,        /* required pointer is NULL */
  B200Q_ERR_SHAPE = -2,       /* K/N/M/group violate the layout's constraints (cf. ERR_PROB_SHAPE, marlin_cuda_kernel.cu:803) */
  B200Q_ERR_UNSUPPORTED = -3, /* (layout, bits, group, g_idx) combination has no kernel */
  B200Q_ERR_ALIGNMENT = -4,   /* pointer / stride not 16-byte aligned where required */
  B200Q_ERR_WORKSPACE = -5,   /* workspace smaller than b200q_workspace_bytes() */
  B200Q_ERR_CUDA = -6,        /* a CUDA runtime call failed; see b200q_last_cuda_error() */
  B200Q_ERR_ARCH = -7         /* device is not sm_100 (tcgen05/TMA kernels cannot run) */
} b200q_status;

/* Packed-weight layouts ("pack modes", qllm/run.py:58-70; classes in qllm/modeling/q_layers/). */
typedef enum b200q_layout {
  B200Q_LAYOUT_GPTQ = 0,     /* QuantLinearGPTQ   quant_linear_gptq.py:92-143  qweight i32 [K*b/32, N], qzeros i32 [G, N*b/32] */
  B200Q_LAYOUT_AWQ_GEMM = 1, /* WQLinear_GEMM     quant_linear_awq.py:38-153   qweight i32 [K, N/8],   qzeros i32 [G, N/8], nibble order 0,2,4,6,1,3,5,7 */
  B200Q_LAYOUT_MARLIN = 2,   /* QuantLinearMarlin quant_linear_marlin.py:60-146 qweight i32 [K/16, 2N], scales permuted, z == 8 */
  B200Q_LAYOUT_HQQ = 3,      /* QuantLinearHQQ    quant_linear_hqq.py:48-80    qweight as GPTQ, qzeros fp16 [G, N] */
  B200Q_LAYOUT_ORT = 5,      /* QuantLinearORT    quant_linear_onnxruntime.py:85-174 (com.microsoft::MatMulNBits blobs, 4-bit): qweight u8
                                [N, G, group/2] (byte b of block (n, g) = q[g group + 2b] | q[.. + 1] << 4), qzeros u8 [N, ceil(G/2)]
                                (byte j = z[2j] | z[2j+1] << 4), scales fp16 [N, G]; g_idx allowed (groups of scales / zeros only) */
  B200Q_LAYOUT_AWQ_GEMV = 4  /* WQLinear_GEMV     quant_linear_awq.py:156-265  qweight i32 [N, K/8] (nibble i = k 8w+i), qzeros i32 [N, ZW]
                                (nibble i of word c = group 8c+i), scales fp16 [N, 8 ZW], ZW = calculate_zeros_width (:15-27);
                                kernels: gemv_cuda.cu:60-186, gemmv2 gemm_cuda_gen.cu:683-1094.  Unpack / dequant / the generic
                                kernel read it in place; the fast kernels run on its exact K-packed re-layout (b200q_repack_gptq4) */
} b200q_layout;

/*
 * One quantised linear layer, exactly as its buffers sit in a QLLM/AutoGPTQ/AutoAWQ checkpoint
 * after load_state_dict (no repacking).  Mirrors the module attributes of the q_layers classes.
 */
typedef struct b200q_layer {
  int32_t layout;      /* b200q_layout */
  int32_t bits;        /* 2,3,4,8 fast paths; 5,6,7 generic path (GPTQ/HQQ only); AWQ/MARLIN: 4 */
  int32_t group_size;  /* >0; per-channel (-1 in the reference) must be passed as K */
  int32_t K;           /* in_features  */
  int32_t N;           /* out_features held by this rank (column shard width when sharded) */
  int32_t zero_bias;   /* added to unpacked integer zeros then masked: the reference's
                          COMPATIBLE_WITH_AUTOGPTQ env / add_zero_bias arg (ort_ops.cc:63). 0 or 1. */
  const void* qweight; /* int32 (uint8 blobs for ORT), shape per layout */
  const void* qzeros;  /* int32 packed (GPTQ, AWQ_GEMM, AWQ_GEMV) | fp16 [G,N] (HQQ) | NULL (MARLIN) */
  const void* scales;  /* fp16 [G,N] (MARLIN: in the reference's permuted order) */
  const int32_t* g_idx;/* int32 [K] act-order group map, or NULL for k / group_size (GPTQ only) */
  const void* bias;    /* fp16 [N] or NULL */
  const int32_t* x_perm; /* NULL, or int32 [K]: packed row j multiplies x[:, x_perm[j]] -- the run-time half of the
                          act-order re-layout (b200q_repack_actorder); GPTQ/HQQ layouts with g_idx == NULL only */
} b200q_layer;

/*
 * y[M, N] = x[M, K] @ dequant(layer) (+ bias).  The single forward entry point: chooses the
 * decode kernel (M <= b200q_gemv_max_m()) or the tcgen05 tensor-core GEMM.
 * Replaces, depending on layer->layout:
 *   awq_inference_engine.gemm_forward_cuda   csrc/awq_cuda/pybind_awq.cpp:16, quantization/gemm_cuda_gen.cu:1102-1161
 *   awq_inference_engine.mul (Marlin)        csrc/awq_cuda/pybind_awq.cpp:20, quantization/marlin_cuda.cpp:29-74
 *   ort_ops.gemv                             csrc/ort_cuda/ort_ops.cc:94-140
 *   ort_ops.dequant + torch.matmul           csrc/ort_cuda/ort_ops.cc:58-92, quant_linear_gptq.py:81-85
 *   DequantAndUnpack + torch.matmul (HQQ)    qllm/modeling/q_layers/quant_linear_hqq.py:8-38
 * ldx / ldy are row strides in elements (ldy lets a column shard write into a wider output).
 * workspace: >= b200q_workspace_bytes(layer, M) bytes, 16-byte aligned, ZERO on first use; the
 * engine leaves it zeroed again on completion (same contract as Marlin's workspace,
 * marlin_cuda_kernel.cu:203-210), so one buffer may be reused by successive calls on a stream.
 */
int b200q_linear(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, void* y, int64_t ldy,
                 void* workspace, size_t workspace_bytes, b200q_stream_t stream);

/*
 * b200q_linear with the element-wise neighbours of the Linear fused in.  In an HF decoder
 * block the QuantLinear calls are surrounded by  down_proj(act_fn(gate_proj(x)) * up_proj(x))  and  residual + o_proj(..) /
 * residual + down_proj(..)  -- separate element-wise kernels in the reference's model code, each a launch plus an HBM
 * round trip of the activations:
 *   x_mul    != NULL: the layer's input is  silu(x) * x_mul  (x = gate_proj output, x_mul = up_proj output, both fp16
 *                     [M, ldx]); rounded exactly as the two fp16 ops round (silu to fp16, product to fp16);
 *   residual != NULL: y = fp16(fp16(x @ W + bias) + residual), residual fp16 [M, ldres] -- bit-identical to the separate add.
 * The integer decode kernel folds x_mul into its x-load stage, and it and the tcgen05 GEMM add the residual in their
 * epilogue (separate template instantiations: the plain paths carry none of it); other kernels get a small element-wise
 * pass ahead / behind (same results).
 *   act_dtype = B200Q_ACT_BF16: x, x_mul, residual and y are bfloat16 (same strides, in elements).  The arithmetic is the
 *                     reference's for a bf16 model: the input is rounded to fp16 (auto_cast, quant_linear_awq.py:29-36), the
 *                     Linear runs in fp16, the result is rounded to bf16 (out.to(x.dtype), :146); the fused neighbours round
 *                     as the model's bf16 torch ops do (silu to bf16, product to bf16, residual sum to bf16).  The integer
 *                     decode kernel reads and writes bf16 directly; the other kernels get a conversion pass through the
 *                     workspace.  No cast kernels on the host side.
 * Workspace: b200q_workspace_bytes_ex().
 */
enum b200q_act_dtype { B200Q_ACT_F16 = 0, B200Q_ACT_BF16 = 1 };
typedef struct b200q_fusion {
  const void* x_mul;
  const void* residual;
  int64_t ldres;
  int32_t act_dtype; /* b200q_act_dtype of x, x_mul, residual and y (bf16 callers without a cast on the host side) */
  int32_t reserved;  /* 0 */
} b200q_fusion;
int b200q_linear_ex(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, void* y, int64_t ldy, const b200q_fusion* fusion,
                    void* workspace, size_t workspace_bytes, b200q_stream_t stream);
size_t b200q_workspace_bytes_ex(const b200q_layer* layer, int64_t M, const b200q_fusion* fusion);

/* Same contract, forcing the decode (GEMV-class) kernel; M must be <= b200q_gemv_max_m(). */
int b200q_gemv(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, void* y, int64_t ldy,
               void* workspace, size_t workspace_bytes, b200q_stream_t stream);

/* Same contract, forcing the tcgen05 GEMM kernel (any M >= 1). */
int b200q_gemm(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, void* y, int64_t ldy,
               void* workspace, size_t workspace_bytes, b200q_stream_t stream);

/*
 * Sibling layers that share their input -- q_proj/k_proj/v_proj of LlamaAttention, gate_proj/up_proj of
 * LlamaMLP -- as ONE launch: y[i][M, N_i] = x[M, K] @ dequant(layers[i]) (+ bias_i), i < n_layers <= 3.
 * In the reference these are separate, back-to-back QuantLinear.forward calls on the same tensor
 * (quant_linear_awq.py:142-148 et al., called from the HF attention / MLP modules); at decode sizes each
 * call is launch- and latency-bound, so the host shim (qllm_b200.fuse_siblings) routes them here.  All
 * layers must agree on (layout, bits, group_size, K, zero_bias); otherwise, or for M above
 * b200q_gemv_max_m(), the call is one kernel per layer with the results of b200q_linear.
 * M > 64 with every layer on the tcgen05 GEMM (no g_idx / x_perm) and a workspace of at least 4096 bytes: the GEMMs are
 * launched back to back and only the first orders itself behind the stream; the later siblings' activation loads wait for a
 * flag the first one raises (they read the same x), so their CTAs run beside the previous sibling's last wave instead of
 * waiting for it to end.  The flag and its counter live in workspace words 1008..1013 and are zero again when the call's
 * kernels have finished (CUDA-graph replays need no reset).  Option "gemm_siblings" = 0 turns this off.
 */
int b200q_linear_group(const b200q_layer* const* layers, int32_t n_layers, const void* x, int64_t M, int64_t ldx,
                       void* const* y, const int64_t* ldy, void* workspace, size_t workspace_bytes,
                       b200q_stream_t stream);

/*
 * Multi-GPU column-sharded forward with the all-gather fused into the epilogue: this rank computes
 * y[:, n_offset : n_offset + layer->N] and stores the tile into `n_peers` output buffers
 * (peer_y[r] = rank r's full [M, ldy] output, mapped through NVLink peer access / symmetric memory;
 * peer_y[self] is the local one).  No reference equivalent (the reference is single-GPU).
 */
int b200q_linear_sharded(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx,
                         void* const* peer_y, int32_t n_peers, int64_t ldy, int64_t n_offset,
                         void* workspace, size_t workspace_bytes, b200q_stream_t stream);

/*
 * Column-sharded decode with the all-gather AND the cross-GPU hand-off fused into the kernel (M <= b200q_gemv_max_m()):
 * sibling layers that share x (n_layers <= 3) run as one launch; this rank computes columns
 * [n_offset[i], n_offset[i] + layers[i]->N) of output i and stores them into all n_peers replicas
 * peer_y[i * n_peers + r] (rank r's full [M, ldy[i]] buffer, mapped into this process through NVLink peer access /
 * symmetric memory); no collective kernel and no host involvement between layers.
 *   - before reading x the kernel waits until the LOCAL counter word counters[self][wait_slot] has reached
 *     (*epoch) * wait_count, i.e. every peer's storing CTAs of the layer(s) that produced x have posted
 *     (wait_slot < 0: x is local, no wait);
 *   - after the stores of all its CTAs are performed the launch adds b200q_sharded_posts() (currently 1: the storing
 *     CTAs arrive on a local counter in `workspace`, the last one posts) to counters[r][post_slot] on every peer
 *     r != self (release, system scope; post_slot < 0: nobody consumes this output remotely), so wait_count of the
 *     consumer = sum over the (n_peers - 1) peers of their producer's posts.
 * Counters (u64, symmetric memory, zero at start, slot 0 reserved: set to ~0 when a wait timed out after 2 s) only
 * ever grow and *epoch is the step number (>= 1, local memory, b200q_peer_epoch_advance once per token), so the whole
 * token, advance included, is CUDA-graph capturable.  The caller keeps ranks in lock-step by construction: a rank
 * cannot run ahead of a peer by more than the layers between two waits.
 * Returns B200Q_ERR_UNSUPPORTED where the integer-path decode kernel does not cover the layers (4-bit GPTQ / HQQ
 * K-packed buffers -- AWQ / Marlin through b200q_repack_gptq4 -- no act-order, M <= 2): use b200q_linear_sharded
 * plus a collective there.  No reference equivalent (the reference is single-GPU).
 */
typedef struct b200q_peer_sync {
  int32_t n_peers;            /* ranks, 1..8 (1: plain local group call) */
  int32_t self;               /* this rank */
  uint64_t* const* counters;  /* [n_peers] device pointers: rank r's counter array as mapped in this process */
  const uint64_t* epoch;      /* local device word: current step number */
  int32_t wait_slot;          /* counter index awaited before x is read, or -1 */
  uint32_t wait_count;        /* posts per step that complete wait_slot */
  int32_t post_slot;          /* counter index posted on every peer after the stores, or -1 */
  uint32_t flags;             /* B200Q_PEER_Y_TAGGED | B200Q_PEER_X_TAGGED, see below */
  uint32_t tag_stride;        /* tagged mode: calls per step (any value > every seq) */
  uint32_t y_seq;             /* tagged mode: index of THIS call inside the step (tags its outputs) */
  uint32_t x_seq;             /* tagged mode: index of the call that produced x (b200q_peer_untag: that produced `tagged`) */
} b200q_peer_sync;
/*
 * Tagged activations (flag-in-data, the low-latency form of the hand-off; measured on 2 x B200 the counter form costs
 * ~4-10 us per launch in system-scope fences after the NVLink stores plus ~3-6 us of polling, the tagged form one
 * NVLink write latency): an element is a 32-bit word  fp16 | tag << 16,  tag = ((*epoch) * tag_stride + seq) & 0xffff
 * with seq = the index of the writing call inside the step, so a buffer that is rewritten several times per step (once
 * per decoder block) never shows a stale word with the awaited tag.  Naturally aligned 4-byte stores are single-copy
 * atomic, so value and tag arrive together and neither side needs a fence, an atomic or a counter.
 *   B200Q_PEER_Y_TAGGED: peer_y[] are uint32 [M, ldy] buffers; the epilogue writes tagged words (post_slot is ignored).
 *   B200Q_PEER_X_TAGGED: x is a uint32 [M, ldx] tagged buffer (16-byte aligned); every lane spins on exactly the words
 *     it needs until they carry this step's tag (wait_slot is ignored).
 * A buffer must not already hold the awaited tag (zero-initialise, epoch starts at 1, tag_stride > 0; consecutive writes
 * of one buffer differ in seq).  b200q_peer_untag() converts a tagged buffer to plain fp16 for consumers outside the engine.
 */
#define B200Q_PEER_Y_TAGGED 1u
#define B200Q_PEER_X_TAGGED 2u
/*
 * B200Q_PEER_NODE_EPOCH (with B200Q_PEER_Y_TAGGED): `epoch` points at 1 + tag_stride words; word 0 is the step number as
 * before, word 1 + y_seq counts the completed executions of call y_seq and is maintained by that call's own kernel (its last
 * storing CTA).  The kernel derives the step from ITS OWN word (a call runs once per step, and a step starts after the
 * previous one has ended, so the word is stable and current when the kernel starts), which removes the one thing its
 * x-load stage needed the kernel boundary for: with tagged x it no longer executes griddepcontrol.wait before reading
 * activations (the tags order the data) -- only ahead of its stores, where the wait has long been satisfied.  Every call of
 * the step must run exactly once per step (a replayed CUDA graph), words start at zero.
 */
#define B200Q_PEER_NODE_EPOCH 4u
int b200q_peer_untag(const void* tagged, int64_t ld_tagged, void* y, int64_t ldy, int64_t M, int64_t N,
                     const b200q_peer_sync* sync, b200q_stream_t stream);
int b200q_linear_group_sharded(const b200q_layer* const* layers, int32_t n_layers, const void* x, int64_t M, int64_t ldx,
                               void* const* peer_y, const int64_t* ldy, const int64_t* n_offset, const b200q_peer_sync* sync,
                               void* workspace, size_t workspace_bytes, b200q_stream_t stream);
/* Posts per peer of one b200q_linear_group_sharded call on these layers at batch M; < 0: status. */
int b200q_sharded_posts(const b200q_layer* const* layers, int32_t n_layers, int64_t M);
/* The wait half alone, for consumers outside the engine (attention, lm_head, a device-to-host copy): enqueues a
 * one-thread kernel that returns once counters[self][wait_slot] >= (*epoch) * wait_count (same 2 s bound). */
int b200q_peer_wait(const b200q_peer_sync* sync, b200q_stream_t stream);
/* *epoch += 1 on `stream` (one tiny kernel; capture it at the head of the token's CUDA graph). */
int b200q_peer_epoch_advance(uint64_t* epoch, b200q_stream_t stream);

/*
 * Decode chain (M <= 2): a recorded run of b200q_linear_group calls executed by ONE persistent launch.
 * In the reference every QuantLinear.forward of a decode step is its own kernel launch (quant_linear_awq.py:142-148,
 * quant_linear_gptq.py:55-85); at batch 1 a Llama-sized layer is 1-7 us of HBM streaming and a launch boundary costs as
 * much again.  A chain keeps one CTA per SM alive across the layers: the packed weights of ALL steps stream through one
 * shared-memory ring (TMA) without draining at layer boundaries, layers are separated by a grid-wide barrier, and a
 * step whose x points into an output of the PREVIOUS step reads it without a round trip through a finished y.
 * Each step has exactly the semantics of b200q_linear_group(layers, n_layers, x, M, ldx, y, ldy): every y[i] is
 * written as fp16 (bias added) and is bit-identical to what a consumer inside the chain read.
 *   plan:  host-side, once per (steps, M): fills `plan_host` (b200q_chain_plan_bytes() bytes: tensor maps, schedule)
 *          and reports the workspace size; the caller copies the blob to device memory (64-byte aligned).
 *   run:   one kernel launch on `stream`; `workspace` zero on first use and left with its first 4 KB zeroed.
 * Supported: 4-bit K-packed layers (GPTQ / HQQ layouts with g_idx == NULL, optionally x_perm; AWQ-GEMM / Marlin through
 * b200q_repack_gptq4), group 64, 128 or a multiple of 256, K % 256 == 0, N % 32 == 0.  b200q_chain_plan returns
 * B200Q_ERR_UNSUPPORTED otherwise (run the steps through b200q_linear_group instead).
 */
typedef struct b200q_chain_step {
  const b200q_layer* const* layers; /* [n_layers] sibling layers sharing x (same K, group, layout) */
  int32_t n_layers;                 /* 1..3 */
  const void* x;                    /* fp16 [M, ldx] */
  int64_t ldx;
  void* const* y;                   /* [n_layers] fp16 [M, ldy[i]] */
  const int64_t* ldy;
} b200q_chain_step;
size_t b200q_chain_plan_bytes(const b200q_chain_step* steps, int32_t n_steps, int64_t M); /* 0: unsupported / invalid */
int b200q_chain_plan(const b200q_chain_step* steps, int32_t n_steps, int64_t M, void* plan_host, size_t plan_bytes,
                     size_t* workspace_bytes);
int b200q_chain_run(const void* plan_host, const void* plan_device, void* workspace, size_t workspace_bytes,
                    b200q_stream_t stream);

/*
 * W_out[K, N] (fp16, row-major, ld = N) = dequant(layer), rounded once: fp16((q - z) * s).
 * Replaces ort_ops.dequant (csrc/ort_cuda/ort_ops.cc:58-92 -> dq_gemv.cu:696-725) for every
 * layout and bit width, with or without g_idx.
 */
int b200q_dequant(const b200q_layer* layer, void* w_out, b200q_stream_t stream);

/*
 * Integer unpack (bit-exact gate): q_out int32 [K, N]; z_out int32 [G, N] (may be NULL; for HQQ
 * zeros are floats and z_out is ignored).  Device restatement of general_unpack_on_row
 * (compress_weight.py:87-92), WQLinear_GEMM.unpack_qweight/unpack_qzeros (quant_linear_awq.py:76-93)
 * and of the Marlin inverse permutation the reference lacks (quant_linear_marlin.py:139-140).
 */
int b200q_unpack(const b200q_layer* layer, int32_t* q_out, int32_t* z_out, b200q_stream_t stream);

/*
 * One-time exact integer re-layout of a 4-bit AWQ-GEMM, AWQ-GEMV or Marlin layer into the K-packed GPTQ layout
 * (qweight i32 [K/8, N], qzeros i32 [G, N/8] holding z with bias 0, scales fp16 [G, N] natural order).
 * The host shim re-lays an AWQ-GEMM / Marlin layer out ONCE (first forward), releases the checkpoint-format buffers and
 * runs every kernel on the K-packed copy (zero extra weight memory; b200q_repack_from_gptq4 restores the checkpoint
 * format for state_dict()).  GPTQ / HQQ checkpoints are consumed in place.  Device restatement of
 * the unpack -> pack round trip of repack_to_new_mode (qllm/auto_model_quantization.py:115-147) without
 * the fp16 dequant/requant in the middle.
 */
int b200q_repack_gptq4(const b200q_layer* layer, void* qweight_out, void* qzeros_out, void* scales_out, b200q_stream_t stream);

/*
 * The inverse of b200q_repack_gptq4 (exact): a K-packed 4-bit layer (layout GPTQ, g_idx == NULL) -> the buffers of
 * target_layout = B200Q_LAYOUT_AWQ_GEMM (qweight [K, N/8], qzeros [G, N/8], scales [G, N]), B200Q_LAYOUT_AWQ_GEMV
 * (qweight [N, K/8], qzeros [N, ZW], scales [N, 8 ZW], padding zero), B200Q_LAYOUT_ORT (MatMulNBits blobs) or B200Q_LAYOUT_MARLIN
 * (qweight [K/16, 2N], scales [G, N] in Marlin's permuted order; qzeros_out ignored -- the caller guarantees z == 8),
 * bit-identical to what WQLinear_GEMM.pack / QuantLinearMarlin.pack produce from the same integers
 * (quant_linear_awq.py:95-140, quant_linear_marlin.py:95-137).  Together with b200q_repack_gptq4 this is the integer
 * core of `--pack_mode` conversion (repack_to_new_mode, qllm/auto_model_quantization.py:115-147) without the fp16
 * dequantise / re-quantise round trip, and it lets the host shim keep ONE packed copy of an AWQ / Marlin layer in HBM
 * (the K-packed one the kernels read) while state_dict() still returns the checkpoint's own format.
 */
int b200q_repack_from_gptq4(const b200q_layer* layer, int32_t target_layout, void* qweight_out, void* qzeros_out, void* scales_out,
                            b200q_stream_t stream);

/*
 * Act-order (desc_act) checkpoints: qweight_out[K*b/32, N] = the layer's packed rows re-ordered so that packed row j
 * holds original row perm[j] (exact integer re-layout, any bit width 2..8).  With perm = a stable argsort of g_idx the groups
 * become contiguous, so {qweight_out, the ORIGINAL qzeros and scales, g_idx = NULL, x_perm = perm} is an ordinary
 * grouped layer whose activations are gathered through x_perm: inside the x-load stage of the integer-path decode
 * kernel (M <= 2), by a small gather pass into the workspace ahead of every other kernel (b200q_workspace_bytes accounts
 * for it).  Replaces the per-element g_idx look-ups of Gemv_g / DequantizeAndUnpackWeight*_g
 * (csrc/ort_cuda/dq_gemv.cu:459-541, :189-273).  Requires every group to own exactly group_size rows (true for GPTQ
 * act-order, static or not: gptq.py:230-237); other g_idx maps stay on the generic kernel.
 */
int b200q_repack_actorder(const b200q_layer* layer, const int32_t* perm, void* qweight_out, b200q_stream_t stream);

/* Bytes of zero-initialised workspace b200q_linear/gemv/gemm need for this layer at batch M. */
size_t b200q_workspace_bytes(const b200q_layer* layer, int64_t M);

/* Largest M routed to the decode kernel by b200q_linear. */
int b200q_gemv_max_m(void);

/* Which kernel b200q_linear would pick: 1 = decode kernel, 2 = tcgen05 GEMM, <0 = status. */
int b200q_select_kernel(const b200q_layer* layer, int64_t M);

/* Number of kernel launches issued by this process through the library (bench accounting). */
uint64_t b200q_launch_count(void);

/* Diagnostic only: when device_buf != NULL every decode-kernel CTA records 8 x u64 %globaltimer phase
 * stamps (start, prefetch issued, upstream done, operands in smem, math done, cluster reduced, stored, -)
 * appended launch after launch until `bytes` are used.  NULL (default) disables. */
void b200q_debug_set_timeline(void* device_buf, size_t bytes);

/* Diagnostic only: decode-chain phase stamps, 16 x u64 per (step, CTA): consumer warp 0 {step start, x ready, digits done,
 * own units done, all warps done, partial sums stored}, sync warp {grid barrier passed, y written}, producer {first slab
 * issued, last slab issued, ns stalled on a full ring}.  device_buf: n_steps * SMs * 128 bytes, or NULL to disable. */
void b200q_debug_set_chain_timeline(void* device_buf);

/* Diagnostic only: the decode kernel's launch plan for this layer at batch M (host-side, no CUDA call):
 * out = {cluster size (CTAs that split K), column tiles, dynamic shared memory bytes per CTA, k-steps}. */
int b200q_debug_decode_plan(const b200q_layer* layer, int64_t M, int32_t out[4]);

/* Diagnostic / tuning only: override a dispatch or planner switch at run time (same switches as the B200Q_*
 * environment variables): "stream" (0 = bulk-copy decode kernel only), "st_cluster", "st_depth", "st_tpc", "st_target",
 * "st_ring_kb", "st_lean", "imma", "im_cluster", "im_depth", "im_tpc", "im_target", "tt256_min_m", "gemm_pdl",
 * "gemm_splitk", "sync_flags", "chain_ctas", "chain_slots", "chain_window".  Returns B200Q_ERR_UNSUPPORTED for an unknown name. */
int b200q_debug_set_option(const char* name, double value);

const char* b200q_strerror(int status);
int b200q_last_cuda_error(void); /* cudaError_t of the most recent failing runtime call */
int b200q_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200Q_H_ */
