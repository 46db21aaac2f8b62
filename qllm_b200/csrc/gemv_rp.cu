// Decode kernel (M <= 8): latency-optimised, "integer-in-subnormal" tensor-core formulation.
//
// A decode-sized layer streams 8-23 MB in ~1.3-3.5 us at HBM speed, so both the dependent chain
// (wait for the previous layer's y -> fetch x -> compute -> reduce -> store) and the unpack ALU work
// matter as much as the streaming itself.  Design:
//
//   1. prefetch: the CTA's whole packed-weight slice goes into shared memory with cp.async (or, variant
//      REGS, into registers with 128/64-bit loads), and the per-group (scale, zero) pairs of its columns
//      into an fp32 table, all BEFORE griddepcontrol.wait -- weights do not depend on the previous
//      kernel, so with programmatic dependent launch the next layers' weights stream in while this
//      layer computes.
//   2. math: y[m,n] = sum_g s[g,n] * ( sum_{k in g} q[k,n] x[m,k]  -  z[g,n] * sum_{k in g} x[m,k] ).
//      The packed nibbles are fed to mma.sync.m16n8k16 *as they are*: (word & 0x000f000f) is a pair of
//      fp16 SUBNORMALS with value q * 2^-24, exact for q < 1024, so unpacking costs one AND per two
//      weights (no magic-number subtract, no per-weight zero/scale arithmetic).  Nibbles sitting 4 bits
//      higher ((word & 0x00f000f0) = 16 q * 2^-24) go to their own accumulator and are folded in with
//      an exact 1/16 at group end.  sum_k x[m,k] comes from one extra MMA against an all-ones A
//      fragment.  Per group the fp32 fix-up  s * (2^24 * acc - z * S)  is 3 FMAs per output.
//      Everything after the integer product is fp32, so this is *more* accurate than the reference's
//      fp16((q-z)*s) weights (gemm_cuda_gen.cu:153-176) while agreeing with it to ~1e-4 of max|y|.
//   3. reduce: the 8 warps of a CTA split K and reduce through shared memory; CTAs that split K further
//      form a thread-block cluster and reduce through distributed shared memory (no global scratch, no
//      atomics, fixed summation order -> deterministic).
//
// The weights are the 16-row A operand of the MMA, the <= 8 activation rows the n8 B operand; the
// k-slot order inside an MMA is whatever the unpack produces and x is permuted to match.  No repacking
// of checkpoint bytes: GPTQ/HQQ (2/4/8-bit), AWQ-GEMM and Marlin layouts are consumed in place.
// Replaces ort_ops.gemv (dq_gemv.cu:40-177), gemm_forward_cuda at M<=8 (gemm_cuda_gen.cu:31-353),
// Marlin at M<=8 (marlin_cuda_kernel.cu:222-733) and the torch HQQ path (quant_linear_hqq.py:8-38).
#include "common.cuh"
#include "common.cuh"
#include "kernels.h"
#include "rp_layouts.cuh"

namespace b200q {

struct RpParams {
  LayerView L;
  const __half* x;
  int64_t ldx;
  int M;
  PeerOut out;
  int64_t ldy, n_offset;
  int n_tiles, cluster, steps_total, group_shift;
  int x_stride;                   // bytes
  int off_x, off_tab, off_red, off_rbuf, off_rbar, off_w;
  unsigned long long* dbg;        // optional per-CTA phase timestamps (diagnostic; nullptr in production)
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define RP_STAMP(i) do { if (p.dbg && tid == 0) p.dbg[(size_t)blockIdx.x * 8 + (i)] = gtime(); } while (0)

// ------------------------------------------------------------------------------------------------
// SM == true : the CTA's packed slice is prefetched into shared memory with cp.async (default).
// SM == false: prefetched into registers (first MAXSTEPS steps per warp).
// MC = 1 for M == 1 (only batch column 0 is live), 2 otherwise.
template <class T, bool SM, int MC>
__global__ void __launch_bounds__(kRpThreads, SM ? T::SM_MIN_BLOCKS : 2) gemv_rp_kernel(const RpParams p) {
  extern __shared__ __align__(128) char smem[];
  char* xs = smem + p.off_x;
  float2* tab = reinterpret_cast<float2*>(smem + p.off_tab);
  float* red = reinterpret_cast<float*>(smem + p.off_red);
  float* rbuf = reinterpret_cast<float*>(smem + p.off_rbuf);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = p.cluster;
  const int rank = (int)cluster_ctarank();
  const int n_tile = blockIdx.x / cs;
  const int n0 = n_tile * T::NT;
  const int ncols = min(T::NT, p.L.N - n0);             // last tile may be partial (multiple of N_GRAN)
  const int U = cs * kWarps;
  const int S = p.steps_total;
  const int unit = rank * kWarps + warp;
  const int s_begin = (int)(((long long)unit * S) / U), s_end = (int)(((long long)(unit + 1) * S) / U);
  const int cta_s0 = (int)(((long long)rank * kWarps * S) / U), cta_s1 = (int)(((long long)(rank + 1) * kWarps * S) / U);
  const int k_cta0 = cta_s0 * T::KSTEP, k_cta1 = cta_s1 * T::KSTEP;
  const int kslice = k_cta1 - k_cta0;
  RpCtx cx;
  cx.M = p.M; cx.k_cta0 = k_cta0; cx.x_stride = p.x_stride; cx.group = p.L.group; cx.gshift = p.group_shift;
  cx.tab = tab; cx.xs = xs;
  cx.g_first = group_of_k(cx, k_cta0);
  const int g_count = kslice > 0 ? group_of_k(cx, k_cta1 - 1) - cx.g_first + 1 : 0;

  RP_STAMP(0);
  pdl_launch_dependents();
  // cluster reduce rendezvous: rank 0 arms an mbarrier for the (cs-1) partial vectors it will receive through st.async;
  // the cluster-wide arrive is issued now and only waited on right before the remote stores, far off the critical path.
  uint64_t* rbar = reinterpret_cast<uint64_t*>(smem + p.off_rbar);
  if (cs > 1) {
    if (rank == 0 && tid == 0) {
      mbar_init(rbar, 1);
      mbar_expect_tx(rbar, (uint32_t)(cs - 1) * (uint32_t)(T::NT * p.M) * 4u);   // every rank sends NT*M floats
      fence_mbar_init();
    }
    cluster_arrive_relaxed();
  }

  // ---- 1. weights -> shared / registers, group table -> shared (independent of the upstream kernel) ----
  typename T::Step w[SM ? 1 : T::MAXSTEPS];
  const bool lane_ok = (lane >> 2) * T::LANE_COLS < ncols;      // REGS variant: this lane's columns exist
  if (!SM) {
#pragma unroll
    for (int i = 0; i < (SM ? 1 : T::MAXSTEPS); ++i) w[i] = typename T::Step{};
  }
  const uint32_t* wtile = reinterpret_cast<const uint32_t*>(smem + p.off_w);
  if (SM) {
    constexpr int CPR = T::ROW_WORDS / 4;                       // 16-byte chunks per packed row
    const int row0 = cta_s0 * T::ROWS_PER_STEP, nrows = (cta_s1 - cta_s0) * T::ROWS_PER_STEP;
    char* wt = smem + p.off_w;
    static_assert(kRpThreads % CPR == 0, "chunk column is fixed per thread");
    constexpr int RSTEP = kRpThreads / CPR;
    static_assert(RSTEP % T::SWZ_ROWS == 0, "a thread's swizzle term is the same for every row it copies");
    const int cc = tid % CPR, r0 = tid / CPR;
    if (cc * T::COLS_PER_CHUNK < ncols) {
      const uint32_t* src = p.L.qw + T::src_word(p.L, row0 + r0, n0) + 4 * cc;
      const size_t sstep = (T::src_word(p.L, 1, 0) - T::src_word(p.L, 0, 0)) * RSTEP;
      char* dst = wt + T::smem_chunk_byte(r0, cc);
      for (int r = r0; r < nrows; r += RSTEP) {
        cp_async16(dst, src);
        src += sstep;
        dst += (size_t)RSTEP * T::RS_WORDS * 4;
      }
    }
    cp_async_commit();
  } else {
#pragma unroll
    for (int i = 0; i < (SM ? 1 : T::MAXSTEPS); ++i)
      if (s_begin + i < s_end && lane_ok) T::load(w[i], p.L, s_begin + i, n0, lane);
  }
  // columns past N get scale 0 (whatever bits sit in their shared-memory slots then contribute 0)
  for (int item = tid; item < g_count * (T::NT / 8); item += kRpThreads) {
    const int gl = item / (T::NT / 8), c8 = (item % (T::NT / 8)) * 8;
    float2 e[8];
    if (c8 < ncols) T::table_entries8(p.L, cx.g_first + gl, n0 + c8, e);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = make_float2(0.f, 0.f);
    }
    float4* dst = reinterpret_cast<float4*>(tab + (size_t)gl * T::NT + c8);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_float4(e[2 * i].x, e[2 * i].y, e[2 * i + 1].x, e[2 * i + 1].y);
  }
  RP_STAMP(1);

  // ---- 2. activations (produced by the upstream kernel) ----
  pdl_wait();
  RP_STAMP(2);
  {
    const int vec_per_row = kslice >> 3;
    for (int idx = tid; idx < p.M * vec_per_row; idx += kRpThreads) {
      const int m = idx / vec_per_row, v = idx % vec_per_row;
      *reinterpret_cast<uint4*>(xs + (size_t)m * p.x_stride + 16 * v) =
          *reinterpret_cast<const uint4*>(p.x + (size_t)m * p.ldx + k_cta0 + 8 * v);
    }
  }
  if (SM) cp_async_wait_all();
  __syncthreads();
  RP_STAMP(3);

  float tot[T::NTOT][4], acc[T::NACC][4], accS[4] = {0.f, 0.f, 0.f, 0.f};
  zero4(tot);
  zero4(acc);
  int gcur = (s_begin < s_end) ? group_of_k(cx, T::step_k(s_begin, lane)) : 0;

  auto do_step = [&](const typename T::Step& ws, int s) {
    const int gi = group_of_k(cx, T::step_k(s, lane));
    if (gi != gcur) {                       // warp-uniform
      T::template group_end<MC>(tot, acc, accS, cx, gcur - cx.g_first, lane);
      zero4(acc);
      accS[0] = accS[1] = accS[2] = accS[3] = 0.f;
      gcur = gi;
    }
    T::compute(ws, s, cx, acc, accS, lane);
  };

  if (SM) {
    // software pipeline: the next step's packed words are read from shared memory before this step's math
    typename T::Step wn;
    if (s_begin < s_end) T::load_smem(w[0], wtile, s_begin - cta_s0, lane);
#pragma unroll 2
    for (int s = s_begin; s < s_end; ++s) {
      wn = w[0];
      if (s + 1 < s_end) T::load_smem(wn, wtile, s + 1 - cta_s0, lane);
      do_step(w[0], s);
      w[0] = wn;
    }
  } else {
#pragma unroll
    for (int i = 0; i < (SM ? 1 : T::MAXSTEPS); ++i)
      if (s_begin + i < s_end) do_step(w[i], s_begin + i);
    for (int sb = s_begin + T::MAXSTEPS; sb < s_end; sb += T::MAXSTEPS) {
#pragma unroll
      for (int i = 0; i < (SM ? 1 : T::MAXSTEPS); ++i)
        if (sb + i < s_end && lane_ok) T::load(w[i], p.L, sb + i, n0, lane);
#pragma unroll
      for (int i = 0; i < (SM ? 1 : T::MAXSTEPS); ++i)
        if (sb + i < s_end) do_step(w[i], sb + i);
    }
  }
  if (s_begin < s_end) T::template group_end<MC>(tot, acc, accS, cx, gcur - cx.g_first, lane);
  RP_STAMP(4);

  // ---- 3. reduce: warps -> CTA (shared), CTAs of the cluster -> rank 0 (distributed shared) ----
  const int ms = p.M;
  T::store_tot(red + (size_t)warp * T::NT * ms, ms, tot, lane, p.M);
  __syncthreads();
  const int total = T::NT * p.M;          // idx = n * M + m
  constexpr int NV = (T::NT * (MC == 1 ? 1 : kMB) + kRpThreads - 1) / kRpThreads;   // MC == 1 <=> M == 1
  float v[NV];
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    const int idx = tid + r * kRpThreads;
    float sum = 0.f;
    if (idx < total) {
#pragma unroll
      for (int wq = 0; wq < kWarps; ++wq) sum += red[(size_t)wq * T::NT * ms + idx];
    }
    v[r] = sum;
  }
  if (cs > 1) {
    cluster_wait();                          // completes the arrive issued at kernel start: rank 0's mbarrier is armed
    if (rank != 0) {
#pragma unroll
      for (int r = 0; r < NV; ++r) {
        const int idx = tid + r * kRpThreads;
        if (idx < total) st_async_f32(rbuf + (size_t)(rank - 1) * total + idx, rbar, 0u, v[r]);
      }
      RP_STAMP(5);
      return;
    }
    mbar_wait(rbar, 0);
    RP_STAMP(5);
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      const int idx = tid + r * kRpThreads;
      if (idx < total)
        for (int q = 0; q < cs - 1; ++q) v[r] += rbuf[(size_t)q * total + idx];
    }
  }
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    const int idx = tid + r * kRpThreads;
    if (idx < total && idx / p.M < ncols) {
      const int n = idx / p.M, m = idx % p.M;
      float o = v[r];
      if (p.L.bias) o += __half2float(__ldg(p.L.bias + n0 + n));
      const __half h = __float2half_rn(o);
      for (int q = 0; q < p.out.n; ++q) p.out.y[q][(size_t)m * p.ldy + p.n_offset + n0 + n] = h;
    }
  }
  RP_STAMP(6);
}

// ------------------------------------------------------------------------------------------------
struct RpPlan {
  int kind, NT, KSTEP, MAXSTEPS, n_gran, RPS, RSW, min_blocks;
  bool sm;
  int n_tiles, cluster, steps_total, x_stride, group_shift;
  int off_x, off_tab, off_red, off_rbuf, off_rbar, off_w, smem_bytes;
};

template <class T>
static void rp_fill(RpPlan& pl) {
  pl.min_blocks = T::SM_MIN_BLOCKS;
  pl.NT = T::NT; pl.KSTEP = T::KSTEP; pl.MAXSTEPS = T::MAXSTEPS; pl.n_gran = T::N_GRAN; pl.RPS = T::ROWS_PER_STEP;
  pl.RSW = T::RS_WORDS;
}

static int g_rp_max_cluster = 8;
static bool g_rp_smem = true;
static int g_rp_slice_kb = 40;
static int g_rp_min_steps = 8;
static int g_rp_planner = 1;
static int g_rp_force_cluster = 0;      // tuning: B200Q_FORCE_CLUSTER
static double g_rp_fill_cap = 0.0;       // 0 = per-kernel default
static unsigned long long* g_rp_dbg = nullptr;
static size_t g_rp_dbg_cap = 0, g_rp_dbg_pos = 0;   // in u64 entries; consecutive launches append

static bool rp_plan(const LayerView& L, int M, RpPlan& pl) {
  pl.kind = 0;
  pl.sm = g_rp_smem;
  if (M < 1 || M > kMB || L.g_idx != nullptr || L.x_perm != nullptr) return false;
  if (L.layout == B200Q_LAYOUT_GPTQ || L.layout == B200Q_LAYOUT_HQQ) {
    if (L.bits == 2) { pl.kind = 1; rp_fill<RpGptq<2>>(pl); }
    else if (L.bits == 4) { pl.kind = 2; rp_fill<RpGptq<4>>(pl); }
    else if (L.bits == 8) { pl.kind = 3; rp_fill<RpGptq<8>>(pl); }
    else return false;
    if (L.group % (4 * (32 / L.bits)) != 0) return false;  // a step (4 packed rows) must lie inside one group
  } else if (L.layout == B200Q_LAYOUT_AWQ_GEMM) {
    if (L.bits != 4 || L.group % 16 != 0) return false;
    pl.kind = 4; rp_fill<RpAwq>(pl);
  } else if (L.layout == B200Q_LAYOUT_MARLIN) {
    if (L.bits != 4 || L.group % 16 != 0) return false;
    pl.kind = 5; rp_fill<RpMarlin>(pl);
  } else return false;
  if (L.K % pl.KSTEP != 0 || L.N % pl.n_gran != 0 || L.K % L.group != 0) { pl.kind = 0; return false; }
  pl.group_shift = -1;
  if ((L.group & (L.group - 1)) == 0) { int s = 0; while ((1 << s) < L.group) ++s; pl.group_shift = s; }
  pl.n_tiles = (L.N + pl.NT - 1) / pl.NT;
  pl.steps_total = L.K / pl.KSTEP;
  // shared-memory footprint of a cluster size (a CTA's k-slice is at most ceil(steps / cs) steps)
  auto layout = [&](int cs, RpPlan& q) {
    const int slice_steps = (pl.steps_total + cs - 1) / cs;
    const int kslice = slice_steps * pl.KSTEP;
    q.x_stride = kslice * 2;
    q.x_stride += (64 - (q.x_stride % 128) + 128) % 128;
    const int gcap = kslice / L.group + 2;
    int off = 0;
    q.off_x = off; off += M * q.x_stride;
    off = (off + 15) & ~15;
    q.off_tab = off; off += gcap * pl.NT * 8;
    off = (off + 15) & ~15;
    q.off_red = off; off += kWarps * pl.NT * M * 4;
    q.off_rbuf = off; off += (cs - 1) * pl.NT * M * 4;
    off = (off + 7) & ~7;
    q.off_rbar = off; off += 8;
    off = (off + 127) & ~127;
    q.off_w = off;
    if (pl.sm) off += slice_steps * pl.RPS * pl.RSW * 4;
    q.smem_bytes = off;
  };
  int cs = 1;
  if (g_rp_planner == 0) {                 // legacy power-of-two rule (B200Q_PLANNER=0)
    if (pl.sm) {
      const long long bytes = (long long)pl.steps_total * pl.RPS * pl.RSW * 4;
      while (cs < g_rp_max_cluster && bytes / cs > (long long)g_rp_slice_kb * 1024) cs *= 2;
    } else {
      while (cs < g_rp_max_cluster && (pl.steps_total + cs * kWarps - 1) / (cs * kWarps) > pl.MAXSTEPS) cs *= 2;
    }
    while (cs < g_rp_max_cluster && pl.n_tiles * cs < 148 && pl.steps_total / (2 * cs * kWarps) >= 2) cs *= 2;
    while (cs > 1 && pl.steps_total / (cs * kWarps) < g_rp_min_steps && pl.n_tiles * (cs / 2) >= 96) cs /= 2;
  } else {
    // wave-aware (rule fitted to the B200 sweep in profiles/r1_decode_cluster_sweep.jsonl): keep the launch to the
    // fewest waves of co-resident CTAs; inside one wave take the largest K split whose CTAs fill at most `cap` of the
    // SM slots -- the free slots are where the next layer's CTAs prefetch their weights under programmatic dependent
    // launch -- else the largest split that still fits one wave.
    const int minb = pl.min_blocks;
    const double cap = g_rp_fill_cap > 0 ? g_rp_fill_cap : (minb <= 2 ? 0.70 : 0.90);
    int best_waves = 1 << 30, best_c = 1;
    bool best_under = false;
    for (int c = 1; c <= g_rp_max_cluster; ++c) {
      if (c > 1 && pl.steps_total / c < kWarps / 2) break;              // at least half the warps get a step
      RpPlan q;
      layout(c, q);
      if (q.smem_bytes > 200 * 1024) continue;
      int per_sm = (226 * 1024) / (q.smem_bytes + 1024);
      if (per_sm > minb) per_sm = minb;
      if (per_sm < 1) per_sm = 1;
      const int ctas = pl.n_tiles * c, slots = 148 * per_sm;
      const int waves = (ctas + slots - 1) / slots;
      const bool under = (double)ctas <= cap * waves * slots;
      bool take = false;
      if (waves < best_waves) take = true;
      else if (waves == best_waves) take = under || !best_under;         // larger c wins unless it loses the free slots
      if (take) { best_waves = waves; best_c = c; best_under = under; }
    }
    cs = best_c;
  }
  if (g_rp_force_cluster > 0 && g_rp_force_cluster <= 8 && pl.steps_total / g_rp_force_cluster >= 1) cs = g_rp_force_cluster;
  pl.cluster = cs;
  layout(cs, pl);
  if (pl.smem_bytes > 200 * 1024) { pl.kind = 0; return false; }
  return true;
}

bool gemv_rp_supported(const LayerView& L, int M, const __half* x, int64_t ldx) {
  RpPlan pl;
  if (!rp_plan(L, M, pl)) return false;
  if (((uintptr_t)x & 15) != 0 || (ldx % 8) != 0) return false;
  if (((uintptr_t)L.qw & 15) != 0) return false;
  return true;
}

// plan introspection (tests / tuning): {cluster, n_tiles, smem_bytes, steps_total}
bool gemv_rp_describe(const LayerView& L, int M, int out[4]) {
  RpPlan pl;
  if (!rp_plan(L, M, pl)) return false;
  out[0] = pl.cluster; out[1] = pl.n_tiles; out[2] = pl.smem_bytes; out[3] = pl.steps_total;
  return true;
}

// shared-memory footprint of the plan (0 if unsupported): lets the dispatcher avoid >1-wave configurations
int gemv_rp_smem_bytes(const LayerView& L, int M) {
  RpPlan pl;
  return rp_plan(L, M, pl) ? pl.smem_bytes : 0;
}

void gemv_rp_set_max_cluster(int c) { g_rp_max_cluster = c < 1 ? 1 : (c > 8 ? 8 : c); }
void gemv_rp_set_smem(bool on, int slice_kb) { g_rp_smem = on; if (slice_kb > 0) g_rp_slice_kb = slice_kb; }
void gemv_rp_set_min_steps(int n) { g_rp_min_steps = n; }
void gemv_rp_set_force_cluster(int c) { g_rp_force_cluster = c; }
void gemv_rp_set_planner(int mode, double fill_cap) {
  g_rp_planner = mode;
  if (fill_cap > 0) g_rp_fill_cap = fill_cap;
}
void gemv_rp_set_debug(unsigned long long* buf, size_t cap_entries) { g_rp_dbg = buf; g_rp_dbg_cap = cap_entries; g_rp_dbg_pos = 0; }

template <class T, bool SM, int MC>
static cudaError_t rp_launch_k(const RpParams& p, const RpPlan& pl, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemv_rp_kernel<T, SM, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    // all of the SM's unified L1/shared array as shared memory, so that consecutive layers' CTAs can be co-resident
    if (decode_carveout_max()) cudaFuncSetAttribute(gemv_rp_kernel<T, SM, MC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    attr_done[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.n_tiles * pl.cluster);
  cfg.blockDim = dim3(kRpThreads);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = pl.cluster;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  count_launch();
  return cudaLaunchKernelEx(&cfg, gemv_rp_kernel<T, SM, MC>, p);
}

template <class T>
static cudaError_t rp_launch_t(const RpParams& p, const RpPlan& pl, cudaStream_t st) {
  if (pl.sm) return p.M == 1 ? rp_launch_k<T, true, 1>(p, pl, st) : rp_launch_k<T, true, 2>(p, pl, st);
  return p.M == 1 ? rp_launch_k<T, false, 1>(p, pl, st) : rp_launch_k<T, false, 2>(p, pl, st);
}

cudaError_t launch_gemv_rp(const LinearArgs& a, const PeerOut* peers) {
  RpPlan pl;
  if (!rp_plan(a.L, a.M, pl)) return cudaErrorInvalidValue;
  RpParams p;
  p.L = a.L; p.x = a.x; p.ldx = a.ldx; p.M = a.M;
  if (peers) p.out = *peers; else { p.out.n = 1; p.out.y[0] = a.y; }
  p.ldy = a.ldy; p.n_offset = a.n_offset;
  p.n_tiles = pl.n_tiles; p.cluster = pl.cluster; p.steps_total = pl.steps_total; p.x_stride = pl.x_stride;
  p.group_shift = pl.group_shift;
  p.off_x = pl.off_x; p.off_tab = pl.off_tab; p.off_red = pl.off_red; p.off_rbuf = pl.off_rbuf; p.off_rbar = pl.off_rbar; p.off_w = pl.off_w;
  p.dbg = nullptr;
  if (g_rp_dbg) {
    const size_t need = (size_t)pl.n_tiles * pl.cluster * 8;
    if (g_rp_dbg_pos + need <= g_rp_dbg_cap) { p.dbg = g_rp_dbg + g_rp_dbg_pos; g_rp_dbg_pos += need; }
  }
  switch (pl.kind) {
    case 1: return rp_launch_t<RpGptq<2>>(p, pl, a.stream);
    case 2: return rp_launch_t<RpGptq<4>>(p, pl, a.stream);
    case 3: return rp_launch_t<RpGptq<8>>(p, pl, a.stream);
    case 4: return rp_launch_t<RpAwq>(p, pl, a.stream);
    case 5: return rp_launch_t<RpMarlin>(p, pl, a.stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace b200q
