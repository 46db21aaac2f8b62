// Internal launch interface between api.cu and the kernel translation units.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace b200q {

void count_launch(uint64_t n = 1);
bool decode_carveout_max();     // B200Q_CARVEOUT=0 leaves the driver's default L1/shared split

struct LinearArgs {
  LayerView L;
  const __half* x;
  int64_t ldx;
  int M;
  // output: either a single y, or n_peers replicated destinations (fused all-gather epilogue)
  __half* y;
  int64_t ldy;
  int64_t n_offset;     // column offset of this shard inside the (possibly wider) output rows
  void* workspace;
  size_t workspace_bytes;
  cudaStream_t stream;
  // fused neighbours of the Linear (b200q_linear_ex; NULL = none)
  const __half* x_mul;       // the layer's input is fp16(fp16(silu(x)) * x_mul), same strides as x (LlamaMLP.down_proj(act(gate) * up))
  const __half* residual;    // y = fp16(fp16(x @ W + bias) + residual): the decoder block's skip connection
  int64_t ldres;
  // sibling GEMMs of one b200q_linear_group call at M > 64 (same x, launched back to back): 1 = first sibling, raises the
  // later ones' flags once its own stream dependency is met; 2 = later sibling number sib_index (0-based), whose
  // activation loads wait for that flag instead of for the kernel in front of it
  int sib_role, sib_index, sib_count;
  int act_bf16;              // b200q_fusion.act_dtype == bf16: integer-path decode kernel: x, x_mul, residual, y are bfloat16; tcgen05 GEMM:
                             // y and residual are (x arrives as an fp16 copy)
};

static constexpr int kMaxPeers = 8;
struct PeerOut {
  __half* y[kMaxPeers];
  int n;
};

// Cross-GPU hand-off fused into the decode kernels (b200q_linear_group_sharded): counters live in symmetric memory,
// counters[r] = rank r's array mapped into this process.  Counters only ever grow: a consumer waits for
// counter >= epoch * wait_count, a producer adds 1 per storing CTA on every peer, `epoch` is the step number kept
// in local device memory (b200q_peer_epoch_advance), so nothing is ever reset and a CUDA graph can bake all of it.
struct PeerSync {
  unsigned long long* counters[kMaxPeers];
  const unsigned long long* epoch;
  int n_peers, self, wait_slot, post_slot;
  unsigned int wait_count;                 // posts (summed over all peers) that complete one step of the awaited slot
  // Tagged mode (flag-in-data, no fences / atomics / counters): an activation element is a 32-bit word
  // fp16 | (epoch & 0xffff) << 16; 4-byte stores are single-copy atomic, so a reader that sees the step's tag sees the value.
  // The tag of a buffer write is ((*epoch) * tag_stride + seq) & 0xffff with seq = the writer's call index inside the
  // step: a buffer rewritten several times per step (once per decoder block) never shows a stale word with the awaited tag.
  int y_tagged, x_tagged;
  unsigned int tag_stride, y_seq, x_seq;
  int node_epoch;                          // B200Q_PEER_NODE_EPOCH: the step comes from epoch[1 + y_seq] + 1, no kernel-boundary wait ahead of tagged x
};

int decode_sync_flags();        // diagnostic switch (B200Q_SYNC_FLAGS / "sync_flags")

// unpack.cu
cudaError_t launch_unpack(const LayerView& L, int32_t* q_out, int32_t* z_out, cudaStream_t st);
cudaError_t launch_dequant(const LayerView& L, __half* w_out, cudaStream_t st);
cudaError_t launch_repack_actorder(const LayerView& L, const int* perm, uint32_t* qw_out, cudaStream_t st);
cudaError_t launch_gather_x(const __half* x, int64_t ldx, const int* perm, __half* out, int M, int K, cudaStream_t st);
cudaError_t launch_silu_mul(const __half* x, const __half* x_mul, int64_t ldx, __half* out, int64_t M, int K, cudaStream_t st);
cudaError_t launch_residual_add(__half* y, int64_t ldy, const __half* res, int64_t ldres, int64_t M, int N, cudaStream_t st);
// bf16 callers of the kernels without native bf16 I/O: x (and silu(x) * x_mul, rounded as bf16 ops round) -> fp16 copy;
// fp16 result (+ bf16 residual, rounded as the bf16 add rounds) -> bf16 y
cudaError_t launch_bf16_in(const void* x, const void* x_mul, int64_t ldx, __half* out, int64_t M, int K, cudaStream_t st);
cudaError_t launch_bf16_out(const __half* y16, void* y, int64_t ldy, const void* res, int64_t ldres, int64_t M, int N, cudaStream_t st);
cudaError_t launch_repack_gptq4(const LayerView& L, uint32_t* qw_out, uint32_t* qz_out, __half* s_out, cudaStream_t st);
cudaError_t launch_repack_from_gptq4(const LayerView& L, int target, uint32_t* qw_out, uint32_t* qz_out, __half* s_out, cudaStream_t st);

// gemv_generic.cu : any layout / bits / group / g_idx, M <= 16, CUDA cores
size_t gemv_generic_workspace(const LayerView& L, int M);
cudaError_t launch_gemv_generic(const LinearArgs& a, const PeerOut* peers);

// gemv_mma.cu : bulk-copy (cp.async.bulk + mbarrier) pipeline + mma.sync, M <= 8: fallback for the group sizes the streaming
// kernels do not tile, and the whole decode path when the streaming kernels are switched off (B200Q_GEMV=v1)
bool gemv_mma_supported(const LayerView& L, int M, const __half* x, int64_t ldx);
size_t gemv_mma_workspace(const LayerView& L, int M);
cudaError_t launch_gemv_mma(const LinearArgs& a, const PeerOut* peers);

// gemv_stream.cu : streaming decode path (per-warp cp.async rings, sibling layers fused in one launch), M <= 8
static constexpr int kMaxGroupLayers = 3;
bool gemv_stream_supported(const LinearArgs* a, int n);
cudaError_t launch_gemv_stream(const LinearArgs* a, int n, const PeerOut* peers);
bool gemv_stream_describe(const LinearArgs* a, int n, int out[6]);
void gemv_stream_set_option(int which, int value);
void gemv_stream_set_debug(unsigned long long* buf, size_t cap_entries);

// gemv_imma.cu : integer-tensor-path decode kernel for K-packed 4-bit layers, M <= 2
bool gemv_imma_supported(const LinearArgs* a, int n);
cudaError_t launch_gemv_imma(const LinearArgs* a, int n, const PeerOut* peers, const PeerSync* sync = nullptr);     // peers: NULL or one per layer
int gemv_imma_posts(const LinearArgs* a, int n);
bool gemv_imma_describe(const LinearArgs* a, int n, int out[6]);
void gemv_imma_set_option(int which, int value);
void gemv_imma_set_debug(unsigned long long* buf, size_t cap_entries);

// decode_chain.cu : a recorded run of b200q_linear_group calls as ONE persistent launch (M <= 2, K-packed 4-bit)
int decode_chain_plan(const LinearArgs* const* groups, const int* n_layers, int n_groups, int M, void* plan_out, size_t plan_cap,
                      size_t* plan_bytes, size_t* ws_bytes);
cudaError_t launch_decode_chain(const void* plan_host, const void* plan_dev, void* ws, size_t ws_bytes, cudaStream_t st);
void decode_chain_set_option(int which, int value);
void decode_chain_set_debug(unsigned long long* buf);   // diagnostic: 16 x u64 per (group, CTA)

// first kCounterBytes of every workspace are arrival counters that must stay zero between calls
static constexpr size_t kCounterBytes = 4096;

// gemm_tcgen05.cu : tensor-core GEMM (TMA + tcgen05 + TMEM), any M
bool gemm_tc_supported(const LayerView& L, int64_t M, const __half* x, int64_t ldx);
size_t gemm_tc_workspace(const LayerView& L, int64_t M);
cudaError_t launch_gemm_tc(const LinearArgs& a, const PeerOut* peers);
void gemm_tc_set_tt256_min_m(int m);
void gemm_tc_set_force(int which, int v);   // diagnostic: 0 = token-tile width, 1 = K splits
void gemm_tc_set_pdl(int on);
void gemm_tc_set_splitk(int on);
void gemm_tc_set_debug(unsigned long long* buf);
void gemm_tc_set_error_flag(int* device_flag);   // bring-up builds (B200Q_BOUNDED_WAITS) only

}  // namespace b200q
