// extern "C" surface of libb200q.so: argument validation, kernel selection, status codes.
// No torch, no allocation, no synchronisation (include/b200q.h states the contract).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace b200q {
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_last_cuda{0};
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool decode_carveout_max() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200Q_CARVEOUT"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

static int g_sync_flags = -1;
int decode_sync_flags() {
  if (g_sync_flags < 0) { const char* e = getenv("B200Q_SYNC_FLAGS"); g_sync_flags = e ? atoi(e) : 0; }
  return g_sync_flags;
}
static int cuda_status(cudaError_t e) {
  if (e == cudaSuccess) return B200Q_OK;
  g_last_cuda.store((int)e);
  (void)cudaGetLastError();
  return B200Q_ERR_CUDA;
}

// tcgen05 / TMA / cluster kernels exist for sm_100 only: checked once per device, B200Q_ERR_ARCH otherwise
static int check_arch() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cuda_status(cudaGetLastError());
  int c = cached[dev & 63].load(std::memory_order_relaxed);
  if (c == 0) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return cuda_status(cudaGetLastError());
    c = (major == 10) ? 1 : 2;
    cached[dev & 63].store(c, std::memory_order_relaxed);
  }
  return c == 1 ? B200Q_OK : B200Q_ERR_ARCH;
}

static int validate(const b200q_layer* L) {
  if (!L || !L->qweight || !L->scales) return B200Q_ERR_NULL;
  if (L->layout != B200Q_LAYOUT_MARLIN && !L->qzeros) return B200Q_ERR_NULL;
  if (L->K <= 0 || L->N <= 0 || L->group_size <= 0) return B200Q_ERR_SHAPE;
  if (L->bits < 2 || L->bits > 8) return B200Q_ERR_UNSUPPORTED;
  if (L->zero_bias != 0 && L->zero_bias != 1) return B200Q_ERR_UNSUPPORTED;
  if (L->K % L->group_size != 0 && L->g_idx == nullptr) return B200Q_ERR_SHAPE;
  if (L->N % 8 != 0) return B200Q_ERR_SHAPE;
  switch (L->layout) {
    case B200Q_LAYOUT_GPTQ:
    case B200Q_LAYOUT_HQQ:
      // bit-stream along K must fill whole words (compress_weight.py:178), qzeros along N likewise
      if (((int64_t)L->K * L->bits) % 32 != 0) return B200Q_ERR_SHAPE;
      if (L->layout == B200Q_LAYOUT_GPTQ && ((int64_t)L->N * L->bits) % 32 != 0) return B200Q_ERR_SHAPE;
      if (L->layout == B200Q_LAYOUT_HQQ && L->g_idx) return B200Q_ERR_UNSUPPORTED;
      break;
    case B200Q_LAYOUT_AWQ_GEMM:
      if (L->bits != 4) return B200Q_ERR_UNSUPPORTED;       // quant_linear_awq.py:42-43
      if (L->g_idx) return B200Q_ERR_UNSUPPORTED;           // quant_linear_awq.py:96-103
      break;
    case B200Q_LAYOUT_ORT:
      if (L->bits != 4) return B200Q_ERR_UNSUPPORTED;               // quant_linear_onnxruntime.py:114 (pack), :46 (forward)
      if (L->K % L->group_size != 0 || L->group_size % 2 != 0) return B200Q_ERR_SHAPE;
      break;
    case B200Q_LAYOUT_AWQ_GEMV:
      if (L->bits != 4 || L->g_idx) return B200Q_ERR_UNSUPPORTED;   // quant_linear_awq.py:161-162
      if (L->K % 8 != 0) return B200Q_ERR_SHAPE;
      if (L->group_size < 128 && L->group_size != 64 && L->group_size != 32) return B200Q_ERR_UNSUPPORTED;   // calculate_zeros_width :15-27
      break;
    case B200Q_LAYOUT_MARLIN:
      if (L->bits != 4 || L->g_idx) return B200Q_ERR_UNSUPPORTED;   // quant_linear_marlin.py:96-99
      if (L->K % 16 != 0 || L->N % 64 != 0) return B200Q_ERR_SHAPE; // tile permutation granularity
      break;
    default:
      return B200Q_ERR_UNSUPPORTED;
  }
  if (((uintptr_t)L->qweight & 3) || ((uintptr_t)L->scales & 1)) return B200Q_ERR_ALIGNMENT;
  if (L->x_perm) {
    if ((L->layout != B200Q_LAYOUT_GPTQ && L->layout != B200Q_LAYOUT_HQQ) || L->g_idx) return B200Q_ERR_UNSUPPORTED;
    if (((uintptr_t)L->x_perm & 15) || (L->K & 1)) return B200Q_ERR_ALIGNMENT;
  }
  return B200Q_OK;
}

static constexpr int kGemvMaxM = 8;
static constexpr int kGenericMaxM = 16;

enum { KERNEL_GEMV_MMA = 1, KERNEL_GEMM_TC = 2, KERNEL_GENERIC = 3 };

// Tuning/diagnostic switches, read once.  B200Q_GEMV=v1 (or the "stream" option = 0) takes the streaming decode kernels
// out of the dispatch: what remains is the bulk-copy / mbarrier kernel of gemv_mma.cu, which is also the fallback for the
// group sizes the streaming kernels do not tile (below 32 at 4-bit, 64 at 2-bit, 16 at 8-bit).
static bool g_use_stream = true;
static bool g_gemm_siblings = true;    // option "gemm_siblings" = 0: b200q_linear_group at M > 64 is a plain loop of launches
static int gemv_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200Q_GEMV");
    v = 1;
    if (e && (e[0] == 'v' || e[0] == 'r')) g_use_stream = false;
    const char* so[6] = {"B200Q_ST_CLUSTER", "B200Q_ST_DEPTH", "B200Q_ST_TPC", "B200Q_ST_TARGET", "B200Q_ST_RING_KB", "B200Q_ST_LEAN"};
    for (int i = 0; i < 6; ++i) {
      const char* sv = getenv(so[i]);
      if (sv) gemv_stream_set_option(i, atoi(sv));
    }
    const char* io[5] = {"B200Q_IMMA", "B200Q_IM_CLUSTER", "B200Q_IM_DEPTH", "B200Q_IM_TPC", "B200Q_IM_TARGET"};
    for (int i = 0; i < 5; ++i) {
      const char* sv = getenv(io[i]);
      if (sv) gemv_imma_set_option(i, atoi(sv));
    }
    const char* t = getenv("B200Q_TT256_MIN_M");
    if (t) gemm_tc_set_tt256_min_m(atoi(t));
    const char* gp = getenv("B200Q_GEMM_PDL");
    if (gp) gemm_tc_set_pdl(atoi(gp));
    const char* gk = getenv("B200Q_GEMM_SPLITK");
    if (gk) gemm_tc_set_splitk(atoi(gk));
  }
  return v;
}
static LinearArgs probe_args(const LayerView& V, int M, const __half* x, int64_t ldx) {
  LinearArgs a = {};
  a.L = V; a.x = x; a.ldx = ldx; a.M = M;
  return a;
}
static bool decode_supported(const LayerView& V, int M, const __half* x, int64_t ldx) {
  gemv_variant();
  if (g_use_stream) {
    const LinearArgs a = probe_args(V, M, x, ldx);
    if (gemv_imma_supported(&a, 1) || gemv_stream_supported(&a, 1)) return true;
  }
  return gemv_mma_supported(V, M, x, ldx);
}

static int select(const LayerView& V, int64_t M, const __half* x, int64_t ldx, int force) {
  if (force == KERNEL_GEMM_TC) return gemm_tc_supported(V, M, x, ldx) ? KERNEL_GEMM_TC : B200Q_ERR_UNSUPPORTED;
  if (force == KERNEL_GEMV_MMA) {
    if (M > kGenericMaxM) return B200Q_ERR_SHAPE;
    return (M <= kGemvMaxM && decode_supported(V, (int)M, x, ldx)) ? KERNEL_GEMV_MMA : KERNEL_GENERIC;
  }
  if (M <= kGemvMaxM && decode_supported(V, (int)M, x, ldx)) return KERNEL_GEMV_MMA;
  if (M > kGemvMaxM && gemm_tc_supported(V, M, x, ldx)) return KERNEL_GEMM_TC;
  // no decode kernel for this (layout, bits, group): the split-K GEMM streams the weights with every SM, the generic
  // kernel is the per-element fallback (3-bit, M = 1: 542 us generic vs the GEMM's tens of us)
  if (M <= kGemvMaxM) return (force == 0 && gemm_tc_supported(V, M, x, ldx)) ? KERNEL_GEMM_TC : KERNEL_GENERIC;
  if (gemm_tc_supported(V, M, x, ldx)) return KERNEL_GEMM_TC;
  return KERNEL_GENERIC;   // chunked over 16-row slabs: slow, but no configuration is refused
}

static size_t workspace_for(const LayerView& V, int64_t M) {
  size_t w = 0;
  const int mg = (int)(M < kGenericMaxM ? M : kGenericMaxM);
  w = gemv_generic_workspace(V, mg);
  if (M <= kGemvMaxM) { const size_t a = gemv_mma_workspace(V, (int)M); if (a > w) w = a; }
  const size_t g = gemm_tc_workspace(V, M);
  if (g > w) w = g;
  return (w + 255) & ~(size_t)255;
}
// act-order re-layout: the gathered activations x[:, x_perm] sit behind the kernels' own scratch (no zero contract there)
static size_t gather_bytes(const LayerView& V, int64_t M) {
  return V.x_perm ? (((size_t)M * (size_t)V.K * 2 + 255) & ~(size_t)255) : 0;
}

static bool act_bf16(const b200q_fusion* fu) { return fu && fu->act_dtype == B200Q_ACT_BF16; }
// scratch behind the kernel's own workspace: a fused / converted fp16 copy of x, and (bf16 callers of a kernel without
// native bf16 I/O) the fp16 result ahead of its conversion
static size_t silu_bytes(const LayerView& V, int64_t M, const b200q_fusion* fu) {
  return (fu && (fu->x_mul || act_bf16(fu))) ? (((size_t)M * (size_t)V.K * 2 + 255) & ~(size_t)255) : 0;
}
static size_t y16_bytes(const LayerView& V, int64_t M, const b200q_fusion* fu) {
  return act_bf16(fu) ? (((size_t)M * (size_t)V.N * 2 + 255) & ~(size_t)255) : 0;
}

static int run(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, const PeerOut* peers, void* y,
               int64_t ldy, int64_t n_offset, void* ws, size_t ws_bytes, b200q_stream_t stream, int force,
               const b200q_fusion* fu = nullptr) {
  const int v = validate(layer);
  if (v != B200Q_OK) return v;
  if (!x || (!y && !peers)) return B200Q_ERR_NULL;
  const int arch = check_arch();
  if (arch != B200Q_OK) return arch;
  if (M < 1 || ldx < layer->K || ldy < n_offset + layer->N) return B200Q_ERR_SHAPE;
  gemv_variant();                          // one-time parse of the B200Q_* switches
  LayerView V = make_view(layer);
  const size_t base = workspace_for(V, M), need = base + gather_bytes(V, M) + silu_bytes(V, M, fu) + y16_bytes(V, M, fu);
  if (need > 0 && (!ws || ws_bytes < need)) return B200Q_ERR_WORKSPACE;
  if (ws && ((uintptr_t)ws & 15)) return B200Q_ERR_ALIGNMENT;
  const __half* x_mul = fu ? (const __half*)fu->x_mul : nullptr;
  const __half* residual = fu ? (const __half*)fu->residual : nullptr;
  const int64_t ldres = fu ? fu->ldres : 0;
  if (residual && (ldres < layer->N || peers)) return B200Q_ERR_SHAPE;
  if (fu && fu->act_dtype != B200Q_ACT_F16 && fu->act_dtype != B200Q_ACT_BF16) return B200Q_ERR_UNSUPPORTED;
  bool bf16 = act_bf16(fu);
  if (bf16 && peers) return B200Q_ERR_UNSUPPORTED;
  void* y_user = y;
  int64_t ldy_user = ldy;
  bool bf16_post = false;                   // the kernel writes fp16 into the workspace, a pass converts (and adds the residual)
  bool gemm_bf16_out = false;               // ... or the tcgen05 GEMM's epilogue does
  if (bf16) {
    // bfloat16 activations: the integer-path decode kernel converts in its load stage and epilogue; every other kernel
    // works on an fp16 copy of x and leaves an fp16 result for the conversion pass
    LinearArgs probe = {};
    probe.L = V; probe.x = (const __half*)x; probe.ldx = ldx; probe.M = (int)M; probe.x_mul = x_mul;
    const bool native = force != KERNEL_GEMM_TC && M <= kGemvMaxM && g_use_stream && gemv_imma_supported(&probe, 1);
    if (!native) {
      __half* xs = (__half*)((char*)ws + base + gather_bytes(V, M));
      const cudaError_t e = launch_bf16_in(x, x_mul, ldx, xs, M, V.K, (cudaStream_t)stream);
      if (e != cudaSuccess) return cuda_status(e);
      x = xs; ldx = V.K; x_mul = nullptr;
      // the tcgen05 GEMM rounds to bf16 (and adds a bf16 residual) in its own epilogue; the rest leave fp16 for a pass
      LayerView Vg = V;
      Vg.x_perm = nullptr;
      gemm_bf16_out = select(Vg, M, (const __half*)x, ldx, force) == KERNEL_GEMM_TC &&
                      (!residual || (((uintptr_t)residual & 15) == 0 && (ldres % 8) == 0));
      bf16 = false;
      if (!gemm_bf16_out) {
        y = (char*)ws + base + gather_bytes(V, M) + silu_bytes(V, M, fu);
        ldy = V.N;
        bf16_post = true;
      }
    }
  }
  const __half* residual_user = residual;
  if (bf16_post) residual = nullptr;        // added by the conversion pass, in bf16
  if (x_mul && !bf16) {
    // act(gate) * up: folded into the x stage of the integer-path decode kernel; every other kernel reads a fused copy
    LinearArgs probe = {};
    probe.L = V; probe.x = (const __half*)x; probe.ldx = ldx; probe.M = (int)M; probe.x_mul = x_mul;
    const bool folded = force != KERNEL_GEMM_TC && M <= kGemvMaxM && g_use_stream && gemv_imma_supported(&probe, 1);
    if (!folded) {
      __half* xs = (__half*)((char*)ws + base + gather_bytes(V, M));
      const cudaError_t e = launch_silu_mul((const __half*)x, x_mul, ldx, xs, M, V.K, (cudaStream_t)stream);
      if (e != cudaSuccess) return cuda_status(e);
      x = xs; ldx = V.K; x_mul = nullptr;
    }
  }
  if (V.x_perm) {
    // the integer-path decode kernel gathers x through x_perm in its own load stage; everything else reads a gathered copy
    LinearArgs probe = {};
    probe.L = V; probe.x = (const __half*)x; probe.ldx = ldx; probe.M = (int)M; probe.x_mul = x_mul;
    const bool folded = force != KERNEL_GEMM_TC && M <= kGemvMaxM && g_use_stream && (gemv_variant(), true) && gemv_imma_supported(&probe, 1);
    if (!folded) {
      if (M > 0x7fffffff / 2) return B200Q_ERR_SHAPE;
      __half* xg = (__half*)((char*)ws + base);
      const cudaError_t e = launch_gather_x((const __half*)x, ldx, V.x_perm, xg, (int)M, V.K, (cudaStream_t)stream);
      if (e != cudaSuccess) return cuda_status(e);
      x = xg; ldx = V.K; V.x_perm = nullptr;
    }
  }
  const int kern = select(V, M, (const __half*)x, ldx, force);
  if (kern < 0) return kern;
  LinearArgs a = {};
  a.L = V; a.x = (const __half*)x; a.ldx = ldx; a.M = (int)M; a.y = (__half*)y; a.ldy = ldy; a.n_offset = n_offset;
  a.workspace = ws; a.workspace_bytes = ws_bytes; a.stream = (cudaStream_t)stream;
  a.x_mul = x_mul; a.residual = residual; a.ldres = ldres; a.act_bf16 = (bf16 || gemm_bf16_out) ? 1 : 0;
  // bf16 callers of a kernel without native bf16 I/O: fp16 result in the workspace -> bf16 y (+ residual, as the bf16 add rounds)
  auto finish = [&](int st) {
    if (st != B200Q_OK || !bf16_post) return st;
    return cuda_status(launch_bf16_out((const __half*)y, y_user, ldy_user, residual_user, ldres, M, V.N, (cudaStream_t)stream));
  };
  // kernels that do not carry the residual epilogue get it as a small pass behind them
  auto with_residual = [&](cudaError_t e) {
    if (e != cudaSuccess || !residual) return finish(cuda_status(e));
    return finish(cuda_status(launch_residual_add((__half*)y, ldy, residual, ldres, M, V.N, (cudaStream_t)stream)));
  };
  if (kern == KERNEL_GEMV_MMA) {
    if (g_use_stream && gemv_imma_supported(&a, 1)) return cuda_status(launch_gemv_imma(&a, 1, peers));
    a.x_mul = nullptr;                                      // (never set here: only the integer-path kernel folds it)
    a.residual = nullptr;                                   // ... and only it and the tcgen05 GEMM carry the residual epilogue
    if (g_use_stream && gemv_stream_supported(&a, 1)) return with_residual(launch_gemv_stream(&a, 1, peers));
    return with_residual(launch_gemv_mma(a, peers));
  }
  if (kern == KERNEL_GEMM_TC) {
    if (residual && (((uintptr_t)residual & 15) != 0 || (ldres % 8) != 0)) { a.residual = nullptr; return with_residual(launch_gemm_tc(a, peers)); }
    return finish(cuda_status(launch_gemm_tc(a, peers)));
  }
  a.residual = nullptr;
  // generic: slabs of <= 16 activation rows
  for (int64_t m0 = 0; m0 < M; m0 += kGenericMaxM) {
    LinearArgs s = a;
    s.M = (int)((M - m0) < kGenericMaxM ? (M - m0) : kGenericMaxM);
    s.x = a.x + m0 * ldx;
    PeerOut po;
    if (peers) { po = *peers; for (int i = 0; i < po.n; ++i) po.y[i] += m0 * ldy; }
    else { po.n = 1; po.y[0] = a.y + m0 * ldy; }
    const cudaError_t e = launch_gemv_generic(s, &po);
    if (e != cudaSuccess) return cuda_status(e);
  }
  return with_residual(cudaSuccess);
}
}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_linear(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, void* y, int64_t ldy, void* workspace,
                 size_t workspace_bytes, b200q_stream_t stream) {
  return run(layer, x, M, ldx, nullptr, y, ldy, 0, workspace, workspace_bytes, stream, 0);
}

int b200q_linear_ex(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, void* y, int64_t ldy, const b200q_fusion* fusion,
                    void* workspace, size_t workspace_bytes, b200q_stream_t stream) {
  return run(layer, x, M, ldx, nullptr, y, ldy, 0, workspace, workspace_bytes, stream, 0, fusion);
}

size_t b200q_workspace_bytes_ex(const b200q_layer* layer, int64_t M, const b200q_fusion* fusion) {
  if (validate(layer) != B200Q_OK || M < 1) return 0;
  gemv_variant();
  const LayerView V = make_view(layer);
  return workspace_for(V, M) + gather_bytes(V, M) + silu_bytes(V, M, fusion) + y16_bytes(V, M, fusion);
}

int b200q_gemv(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, void* y, int64_t ldy, void* workspace,
               size_t workspace_bytes, b200q_stream_t stream) {
  return run(layer, x, M, ldx, nullptr, y, ldy, 0, workspace, workspace_bytes, stream, KERNEL_GEMV_MMA);
}

int b200q_gemm(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, void* y, int64_t ldy, void* workspace,
               size_t workspace_bytes, b200q_stream_t stream) {
  return run(layer, x, M, ldx, nullptr, y, ldy, 0, workspace, workspace_bytes, stream, KERNEL_GEMM_TC);
}

int b200q_linear_sharded(const b200q_layer* layer, const void* x, int64_t M, int64_t ldx, void* const* peer_y,
                         int32_t n_peers, int64_t ldy, int64_t n_offset, void* workspace, size_t workspace_bytes,
                         b200q_stream_t stream) {
  if (!peer_y) return B200Q_ERR_NULL;
  if (n_peers < 1 || n_peers > kMaxPeers || n_offset < 0) return B200Q_ERR_SHAPE;
  PeerOut po;
  po.n = n_peers;
  for (int i = 0; i < n_peers; ++i) {
    if (!peer_y[i]) return B200Q_ERR_NULL;
    po.y[i] = (__half*)peer_y[i];
  }
  return run(layer, x, M, ldx, &po, nullptr, ldy, n_offset, workspace, workspace_bytes, stream, 0);
}

int b200q_linear_group(const b200q_layer* const* layers, int32_t n_layers, const void* x, int64_t M, int64_t ldx,
                       void* const* y, const int64_t* ldy, void* workspace, size_t workspace_bytes, b200q_stream_t stream) {
  if (!layers || !y || !ldy) return B200Q_ERR_NULL;
  if (n_layers < 1 || n_layers > kMaxGroupLayers) return B200Q_ERR_SHAPE;
  gemv_variant();
  LinearArgs a[kMaxGroupLayers];
  bool fused = g_use_stream && M >= 1 && M <= kGemvMaxM;
  for (int i = 0; i < n_layers; ++i) {
    const int v = validate(layers[i]);
    if (v != B200Q_OK) return v;
    if (!x || !y[i]) return B200Q_ERR_NULL;
    if (M < 1 || ldx < layers[i]->K || ldy[i] < layers[i]->N) return B200Q_ERR_SHAPE;
    a[i] = {};
    a[i].L = make_view(layers[i]); a[i].x = (const __half*)x; a[i].ldx = ldx; a[i].M = (int)(fused ? M : 1);
    a[i].y = (__half*)y[i]; a[i].ldy = ldy[i]; a[i].n_offset = 0;
    a[i].workspace = workspace; a[i].workspace_bytes = workspace_bytes; a[i].stream = (cudaStream_t)stream;
  }
  if (fused && gemv_imma_supported(a, n_layers)) return cuda_status(launch_gemv_imma(a, n_layers, nullptr));
  if (fused && gemv_stream_supported(a, n_layers)) return cuda_status(launch_gemv_stream(a, n_layers, nullptr));
  // M > 64 on the tcgen05 GEMM: one launch per sibling, the later ones released by the first sibling's flag instead of by
  // the kernel boundary (gemm_tcgen05.cu, "sibling release"), so their CTAs overlap the previous sibling's last wave
  bool sib = g_gemm_siblings && n_layers >= 2 && M > 64 && workspace && workspace_bytes >= kCounterBytes &&
             !((uintptr_t)workspace & 15) && check_arch() == B200Q_OK;
  for (int i = 0; sib && i < n_layers; ++i)
    sib = !a[i].L.x_perm && !a[i].L.g_idx && select(a[i].L, M, (const __half*)x, ldx, 0) == KERNEL_GEMM_TC && workspace_for(a[i].L, M) == 0;
  if (sib) {
    for (int i = 0; i < n_layers; ++i) {
      a[i].M = (int)M;
      a[i].sib_role = i == 0 ? 1 : 2; a[i].sib_index = i - 1; a[i].sib_count = n_layers - 1;
      const cudaError_t e = launch_gemm_tc(a[i], nullptr);
      if (e != cudaSuccess) return cuda_status(e);
    }
    return B200Q_OK;
  }
  for (int i = 0; i < n_layers; ++i) {      // not fusable (shape / layout mix / M): same result, one launch per layer
    const int st = run(layers[i], x, M, ldx, nullptr, y[i], ldy[i], 0, workspace, workspace_bytes, stream, 0);
    if (st != B200Q_OK) return st;
  }
  return B200Q_OK;
}

__global__ void peer_epoch_advance_kernel(unsigned long long* epoch) { *epoch += 1ull; }

int b200q_peer_epoch_advance(uint64_t* epoch, b200q_stream_t stream) {
  if (!epoch) return B200Q_ERR_NULL;
  if ((uintptr_t)epoch & 7) return B200Q_ERR_ALIGNMENT;
  count_launch();
  peer_epoch_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)epoch);
  return cuda_status(cudaGetLastError());
}

// x of a non-b200q consumer (attention, lm_head, the host): one thread waits as the decode kernels do
__global__ void peer_wait_kernel(const unsigned long long* counter, const unsigned long long* epoch, unsigned int count,
                                 unsigned long long* poison) {
  const unsigned long long target = *reinterpret_cast<const volatile unsigned long long*>(epoch) * count;
  unsigned long long v, t0 = 0;
  unsigned spins = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
    if (v >= target) break;
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { *poison = ~0ull; break; }
    }
  }
}

int b200q_peer_wait(const b200q_peer_sync* sync, b200q_stream_t stream) {
  if (!sync) return B200Q_ERR_NULL;
  if (sync->n_peers < 1 || sync->n_peers > kMaxPeers || sync->self < 0 || sync->self >= sync->n_peers) return B200Q_ERR_SHAPE;
  if (sync->n_peers == 1 || sync->wait_slot < 0) return B200Q_OK;
  if (!sync->counters || !sync->counters[sync->self] || !sync->epoch) return B200Q_ERR_NULL;
  unsigned long long* c = (unsigned long long*)sync->counters[sync->self];
  count_launch();
  peer_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(c + sync->wait_slot, (const unsigned long long*)sync->epoch, sync->wait_count, c);
  return cuda_status(cudaGetLastError());
}

// tagged words -> plain fp16, waiting for this step's tag on every word (consumers outside the engine)
__global__ void peer_untag_kernel(const uint32_t* src, int64_t lds, __half* dst, int64_t ldd, int M, int N,
                                  const unsigned long long* epoch, unsigned int stride, unsigned int seq, unsigned long long* poison) {
  const uint32_t tag = ((uint32_t)(*reinterpret_cast<const volatile unsigned long long*>(epoch)) * stride + seq) & 0xffffu;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)M * N; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i - (int64_t)m * N);
    const volatile uint32_t* w = src + (size_t)m * lds + n;
    uint32_t v, spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
      v = *w;
      if ((v >> 16) == tag) break;
      if ((++spins & 1023u) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ull) { *poison = ~0ull; break; }
      }
    }
    dst[(size_t)m * ldd + n] = __ushort_as_half((unsigned short)(v & 0xffffu));
  }
}

int b200q_peer_untag(const void* tagged, int64_t ld_tagged, void* y, int64_t ldy, int64_t M, int64_t N,
                     const b200q_peer_sync* sync, b200q_stream_t stream) {
  if (!tagged || !y || !sync || !sync->epoch || !sync->counters) return B200Q_ERR_NULL;
  if (sync->self < 0 || sync->self >= sync->n_peers || sync->n_peers > kMaxPeers || !sync->counters[sync->self]) return B200Q_ERR_SHAPE;
  if (M < 1 || N < 1 || ld_tagged < N || ldy < N) return B200Q_ERR_SHAPE;
  const int64_t total = M * N;
  const int blocks = (int)((total + 255) / 256 < 148 ? (total + 255) / 256 : 148);
  count_launch();
  peer_untag_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)tagged, ld_tagged, (__half*)y, ldy, (int)M, (int)N,
                                                             (const unsigned long long*)sync->epoch, sync->tag_stride, sync->x_seq,
                                                             (unsigned long long*)sync->counters[sync->self]);
  return cuda_status(cudaGetLastError());
}

static int group_args(const b200q_layer* const* layers, int32_t n_layers, const void* x, int64_t M, int64_t ldx, LinearArgs* a) {
  if (!layers) return B200Q_ERR_NULL;
  if (n_layers < 1 || n_layers > kMaxGroupLayers) return B200Q_ERR_SHAPE;
  if (M < 1 || M > kGemvMaxM) return B200Q_ERR_SHAPE;
  gemv_variant();
  for (int i = 0; i < n_layers; ++i) {
    const int v = validate(layers[i]);
    if (v != B200Q_OK) return v;
    if (ldx < layers[i]->K) return B200Q_ERR_SHAPE;
    a[i] = {};
    a[i].L = make_view(layers[i]); a[i].x = (const __half*)x; a[i].ldx = ldx; a[i].M = (int)M;
  }
  return B200Q_OK;
}

int b200q_sharded_posts(const b200q_layer* const* layers, int32_t n_layers, int64_t M) {
  LinearArgs a[kMaxGroupLayers];
  const int v = group_args(layers, n_layers, (const void*)16, M, 1 << 30, a);
  if (v != B200Q_OK) return v;
  int n = gemv_imma_posts(a, n_layers);
  return n < 0 ? B200Q_ERR_UNSUPPORTED : n;
}

int b200q_linear_group_sharded(const b200q_layer* const* layers, int32_t n_layers, const void* x, int64_t M, int64_t ldx,
                               void* const* peer_y, const int64_t* ldy, const int64_t* n_offset, const b200q_peer_sync* sync,
                               void* workspace, size_t workspace_bytes, b200q_stream_t stream) {
  if (!x || !peer_y || !ldy || !n_offset || !sync) return B200Q_ERR_NULL;
  if (sync->n_peers < 1 || sync->n_peers > kMaxPeers || sync->self < 0 || sync->self >= sync->n_peers) return B200Q_ERR_SHAPE;
  if (sync->n_peers > 1 && (!sync->counters || !sync->epoch)) return B200Q_ERR_NULL;
  LinearArgs a[kMaxGroupLayers];
  const int v = group_args(layers, n_layers, x, M, ldx, a);
  if (v != B200Q_OK) return v;
  PeerOut po[kMaxGroupLayers];
  for (int i = 0; i < n_layers; ++i) {
    if (n_offset[i] < 0 || ldy[i] < n_offset[i] + layers[i]->N) return B200Q_ERR_SHAPE;
    po[i].n = sync->n_peers;
    for (int r = 0; r < sync->n_peers; ++r) {
      void* y = peer_y[(size_t)i * sync->n_peers + r];
      if (!y) return B200Q_ERR_NULL;
      po[i].y[r] = (__half*)y;
    }
    a[i].y = po[i].y[sync->self]; a[i].ldy = ldy[i]; a[i].n_offset = n_offset[i];
    a[i].workspace = workspace; a[i].workspace_bytes = workspace_bytes; a[i].stream = (cudaStream_t)stream;
  }
  PeerSync ps = {};
  ps.n_peers = sync->n_peers; ps.self = sync->self; ps.wait_slot = sync->wait_slot; ps.post_slot = sync->post_slot;
  ps.wait_count = sync->wait_count; ps.epoch = (const unsigned long long*)sync->epoch;
  ps.y_tagged = (sync->flags & B200Q_PEER_Y_TAGGED) ? 1 : 0;
  ps.x_tagged = (sync->flags & B200Q_PEER_X_TAGGED) ? 1 : 0;
  ps.tag_stride = sync->tag_stride; ps.y_seq = sync->y_seq; ps.x_seq = sync->x_seq;
  ps.node_epoch = (sync->flags & B200Q_PEER_NODE_EPOCH) ? 1 : 0;
  if (ps.node_epoch && (!ps.y_tagged || sync->y_seq >= sync->tag_stride)) return B200Q_ERR_UNSUPPORTED;
  if ((ps.y_tagged || ps.x_tagged) && !sync->epoch) return B200Q_ERR_NULL;
  if (ps.x_tagged && ((uintptr_t)x & 15)) return B200Q_ERR_ALIGNMENT;
  if ((ps.y_tagged || ps.x_tagged) && (!sync->counters || !sync->counters[sync->self])) return B200Q_ERR_NULL;   // slot 0: time-out poison
  for (int r = 0; r < sync->n_peers && (sync->n_peers > 1 || ps.x_tagged || ps.y_tagged); ++r) {
    if (!sync->counters[r]) return B200Q_ERR_NULL;
    ps.counters[r] = (unsigned long long*)sync->counters[r];
  }
  // the hand-off lives in the streaming decode kernels; other kernels would need a separate barrier
  if (g_use_stream && gemv_imma_supported(a, n_layers)) return cuda_status(launch_gemv_imma(a, n_layers, po, &ps));
  return B200Q_ERR_UNSUPPORTED;                             // only the integer-path decode kernel carries the hand-off
}

// ---- decode chain: a recorded run of b200q_linear_group calls as one persistent launch (decode_chain.cu) ----
static int chain_collect(const b200q_chain_step* steps, int32_t n_steps, int64_t M, std::vector<LinearArgs>& flat,
                         std::vector<const LinearArgs*>& ptrs, std::vector<int>& counts) {
  if (!steps) return B200Q_ERR_NULL;
  if (n_steps < 1 || n_steps > 4096 || M < 1 || M > 2) return B200Q_ERR_SHAPE;
  flat.resize((size_t)n_steps * kMaxGroupLayers);
  ptrs.resize(n_steps);
  counts.resize(n_steps);
  for (int g = 0; g < n_steps; ++g) {
    const b200q_chain_step& S = steps[g];
    if (!S.layers || !S.x || !S.y || !S.ldy) return B200Q_ERR_NULL;
    if (S.n_layers < 1 || S.n_layers > kMaxGroupLayers) return B200Q_ERR_SHAPE;
    for (int j = 0; j < S.n_layers; ++j) {
      const int v = validate(S.layers[j]);
      if (v != B200Q_OK) return v;
      if (!S.y[j]) return B200Q_ERR_NULL;
      if (S.ldx < S.layers[j]->K || S.ldy[j] < S.layers[j]->N) return B200Q_ERR_SHAPE;
      LinearArgs& a = flat[(size_t)g * kMaxGroupLayers + j];
      a = {};
      a.L = make_view(S.layers[j]); a.x = (const __half*)S.x; a.ldx = S.ldx; a.M = (int)M; a.y = (__half*)S.y[j]; a.ldy = S.ldy[j];
    }
    ptrs[g] = &flat[(size_t)g * kMaxGroupLayers];
    counts[g] = S.n_layers;
  }
  return B200Q_OK;
}

size_t b200q_chain_plan_bytes(const b200q_chain_step* steps, int32_t n_steps, int64_t M) {
  std::vector<LinearArgs> flat;
  std::vector<const LinearArgs*> ptrs;
  std::vector<int> counts;
  if (chain_collect(steps, n_steps, M, flat, ptrs, counts) != B200Q_OK) return 0;
  size_t bytes = 0;
  if (decode_chain_plan(ptrs.data(), counts.data(), n_steps, (int)M, nullptr, 0, &bytes, nullptr) != B200Q_OK) return 0;
  return bytes;
}

int b200q_chain_plan(const b200q_chain_step* steps, int32_t n_steps, int64_t M, void* plan_host, size_t plan_bytes,
                     size_t* workspace_bytes) {
  if (!plan_host) return B200Q_ERR_NULL;
  const int arch = check_arch();
  if (arch != B200Q_OK) return arch;
  std::vector<LinearArgs> flat;
  std::vector<const LinearArgs*> ptrs;
  std::vector<int> counts;
  const int v = chain_collect(steps, n_steps, M, flat, ptrs, counts);
  if (v != B200Q_OK) return v;
  return decode_chain_plan(ptrs.data(), counts.data(), n_steps, (int)M, plan_host, plan_bytes, nullptr, workspace_bytes);
}

int b200q_chain_run(const void* plan_host, const void* plan_device, void* workspace, size_t workspace_bytes, b200q_stream_t stream) {
  if (!plan_host || !plan_device || !workspace) return B200Q_ERR_NULL;
  if (((uintptr_t)plan_device & 63) || ((uintptr_t)workspace & 15)) return B200Q_ERR_ALIGNMENT;
  const cudaError_t e = launch_decode_chain(plan_host, plan_device, workspace, workspace_bytes, (cudaStream_t)stream);
  if (e == cudaErrorInvalidValue) return B200Q_ERR_WORKSPACE;
  return cuda_status(e);
}

int b200q_repack_actorder(const b200q_layer* layer, const int32_t* perm, void* qweight_out, b200q_stream_t stream) {
  const int v = validate(layer);
  if (v != B200Q_OK) return v;
  if (!perm || !qweight_out) return B200Q_ERR_NULL;
  if (layer->layout != B200Q_LAYOUT_GPTQ && layer->layout != B200Q_LAYOUT_HQQ) return B200Q_ERR_UNSUPPORTED;
  if (32 % layer->bits != 0 && layer->K % 32 != 0) return B200Q_ERR_SHAPE;      // 3/5/6/7-bit: 32-row packs
  return cuda_status(launch_repack_actorder(make_view(layer), perm, (uint32_t*)qweight_out, (cudaStream_t)stream));
}

int b200q_dequant(const b200q_layer* layer, void* w_out, b200q_stream_t stream) {
  const int v = validate(layer);
  if (v != B200Q_OK) return v;
  if (!w_out) return B200Q_ERR_NULL;
  if ((uintptr_t)w_out & 15) return B200Q_ERR_ALIGNMENT;
  return cuda_status(launch_dequant(make_view(layer), (__half*)w_out, (cudaStream_t)stream));
}

int b200q_unpack(const b200q_layer* layer, int32_t* q_out, int32_t* z_out, b200q_stream_t stream) {
  const int v = validate(layer);
  if (v != B200Q_OK) return v;
  if (!q_out) return B200Q_ERR_NULL;
  if ((uintptr_t)q_out & 15) return B200Q_ERR_ALIGNMENT;
  return cuda_status(launch_unpack(make_view(layer), q_out, z_out, (cudaStream_t)stream));
}

int b200q_repack_gptq4(const b200q_layer* layer, void* qweight_out, void* qzeros_out, void* scales_out, b200q_stream_t stream) {
  const int v = validate(layer);
  if (v != B200Q_OK) return v;
  if (!qweight_out || !qzeros_out || !scales_out) return B200Q_ERR_NULL;
  if (layer->bits != 4 || layer->g_idx || layer->layout == B200Q_LAYOUT_HQQ || layer->K % 8 != 0) return B200Q_ERR_UNSUPPORTED;
  return cuda_status(launch_repack_gptq4(make_view(layer), (uint32_t*)qweight_out, (uint32_t*)qzeros_out, (__half*)scales_out,
                                         (cudaStream_t)stream));
}

int b200q_repack_from_gptq4(const b200q_layer* layer, int32_t target_layout, void* qweight_out, void* qzeros_out, void* scales_out,
                            b200q_stream_t stream) {
  const int v = validate(layer);
  if (v != B200Q_OK) return v;
  if (!qweight_out || !scales_out) return B200Q_ERR_NULL;
  if (layer->layout != B200Q_LAYOUT_GPTQ || layer->bits != 4 || layer->g_idx || layer->x_perm || layer->K % 8 != 0) return B200Q_ERR_UNSUPPORTED;
  if (target_layout == B200Q_LAYOUT_AWQ_GEMM || target_layout == B200Q_LAYOUT_AWQ_GEMV || target_layout == B200Q_LAYOUT_ORT) {
    if (!qzeros_out) return B200Q_ERR_NULL;
    if (target_layout == B200Q_LAYOUT_AWQ_GEMV && layer->group_size < 128 && layer->group_size != 64 && layer->group_size != 32)
      return B200Q_ERR_UNSUPPORTED;
  } else if (target_layout == B200Q_LAYOUT_MARLIN) {
    if (layer->K % 16 != 0 || layer->N % 64 != 0) return B200Q_ERR_SHAPE;          // tile permutation granularity
    if (layer->group_size != 128 && layer->group_size != layer->K) return B200Q_ERR_UNSUPPORTED;   // quant_linear_marlin.py:78-80
  } else {
    return B200Q_ERR_UNSUPPORTED;
  }
  return cuda_status(launch_repack_from_gptq4(make_view(layer), target_layout, (uint32_t*)qweight_out, (uint32_t*)qzeros_out,
                                              (__half*)scales_out, (cudaStream_t)stream));
}

size_t b200q_workspace_bytes(const b200q_layer* layer, int64_t M) {
  if (validate(layer) != B200Q_OK || M < 1) return 0;
  gemv_variant();
  const LayerView V = make_view(layer);
  return workspace_for(V, M) + gather_bytes(V, M);
}

int b200q_gemv_max_m(void) { return kGemvMaxM; }

int b200q_select_kernel(const b200q_layer* layer, int64_t M) {
  const int v = validate(layer);
  if (v != B200Q_OK) return v;
  // alignment-independent answer: assume a 16-byte aligned, densely strided x
  LayerView V = make_view(layer);
  if (V.x_perm) {                          // as in run(): only the integer-path decode kernel reads through x_perm itself
    LinearArgs probe = {};
    probe.L = V; probe.ldx = layer->K; probe.M = (int)(M < kGemvMaxM ? M : kGemvMaxM);
    gemv_variant();
    if (!(M <= kGemvMaxM && g_use_stream && gemv_imma_supported(&probe, 1))) V.x_perm = nullptr;
  }
  return select(V, M, (const __half*)nullptr, layer->K, 0);
}

uint64_t b200q_launch_count(void) { return g_launches.load(); }

const char* b200q_strerror(int status) {
  switch (status) {
    case B200Q_OK: return "ok";
    case B200Q_ERR_NULL: return "required pointer is NULL";
    case B200Q_ERR_SHAPE: return "shape violates the layout's constraints";
    case B200Q_ERR_UNSUPPORTED: return "unsupported (layout, bits, group, g_idx) combination";
    case B200Q_ERR_ALIGNMENT: return "pointer or stride misaligned";
    case B200Q_ERR_WORKSPACE: return "workspace too small (see b200q_workspace_bytes)";
    case B200Q_ERR_CUDA: return "CUDA runtime error (see b200q_last_cuda_error)";
    case B200Q_ERR_ARCH: return "device is not sm_100";
    default: return "unknown status";
  }
}

int b200q_last_cuda_error(void) { return g_last_cuda.load(); }
/* diagnostic: per-CTA phase timestamps of the decode kernel (8 x u64 per CTA); NULL disables */
int b200q_debug_set_option(const char* name, double value) {
  if (!name) return B200Q_ERR_NULL;
  gemv_variant();                                    // environment defaults first, then the override
  const std::string n(name);
  if (n == "tt256_min_m") gemm_tc_set_tt256_min_m((int)value);
  else if (n == "stream") g_use_stream = value != 0;
  else if (n == "st_cluster") gemv_stream_set_option(0, (int)value);
  else if (n == "st_depth") gemv_stream_set_option(1, (int)value);
  else if (n == "st_tpc") gemv_stream_set_option(2, (int)value);
  else if (n == "st_target") gemv_stream_set_option(3, (int)value);
  else if (n == "st_ring_kb") gemv_stream_set_option(4, (int)value);
  else if (n == "st_lean") gemv_stream_set_option(5, (int)value);
  else if (n == "imma") gemv_imma_set_option(0, (int)value);
  else if (n == "im_cluster") gemv_imma_set_option(1, (int)value);
  else if (n == "im_depth") gemv_imma_set_option(2, (int)value);
  else if (n == "im_tpc") gemv_imma_set_option(3, (int)value);
  else if (n == "im_target") gemv_imma_set_option(4, (int)value);
  else if (n == "sync_flags") g_sync_flags = (int)value;
  else if (n == "gemm_pdl") gemm_tc_set_pdl((int)value);
  else if (n == "gemm_splitk") gemm_tc_set_splitk((int)value);
  else if (n == "gemm_siblings") g_gemm_siblings = value != 0;
  else if (n == "gemm_force_tt") gemm_tc_set_force(0, (int)value);
  else if (n == "gemm_force_ksplit") gemm_tc_set_force(1, (int)value);
  else if (n == "chain_ctas") decode_chain_set_option(0, (int)value);
  else if (n == "chain_slots") decode_chain_set_option(1, (int)value);
  else if (n == "chain_barrier") decode_chain_set_option(2, (int)value);
  else if (n == "chain_window") decode_chain_set_option(3, (int)value);
  else return B200Q_ERR_UNSUPPORTED;
  return B200Q_OK;
}

int b200q_debug_decode_plan(const b200q_layer* layer, int64_t M, int32_t out[4]) {
  const int v = validate(layer);
  if (v != B200Q_OK) return v;
  if (!out || M < 1 || M > kGemvMaxM) return B200Q_ERR_SHAPE;
  gemv_variant();
  const LayerView V = make_view(layer);
  int o[6];
  const LinearArgs pa = probe_args(V, (int)M, nullptr, V.K);
  if (g_use_stream && (gemv_imma_describe(&pa, 1, o) || gemv_stream_describe(&pa, 1, o))) {
    for (int i = 0; i < 4; ++i) out[i] = o[i];
    return B200Q_OK;
  }
  return B200Q_ERR_UNSUPPORTED;
}

void b200q_debug_set_timeline(void* device_buf, size_t bytes) {
  gemv_variant();
  gemv_stream_set_debug((unsigned long long*)device_buf, bytes / 8);
  gemv_imma_set_debug((unsigned long long*)device_buf, bytes / 8);
  gemm_tc_set_debug(bytes >= 64 * 1024 ? (unsigned long long*)device_buf : nullptr);
}
/* diagnostic: per-(group, CTA) phase stamps of the decode chain kernel: 16 x u64 each, caller sizes the buffer (n_steps * SMs * 128 B) */
void b200q_debug_set_chain_timeline(void* device_buf) {
  decode_chain_set_debug((unsigned long long*)device_buf);   // GEMM: CTA(0,0), 8 stamps per k-block
}
int b200q_version(void) { return B200Q_VERSION; }

}  // extern "C"
