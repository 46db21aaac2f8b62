// Decode kernel (M <= 8), streaming form: every warp runs its own cp.async ring over its share of K.
//
// A chain of decode-sized layers is bound by how continuously HBM is kept busy across kernel boundaries,
// not by any single kernel.  This kernel is built for that:
//
//   * bounded shared memory: a warp owns a ring of D one-step slots (0.5-1 KB each) instead of the CTA
//     holding its whole packed slice, so a launch takes ~1/3 of an SM and two or three consecutive layers
//     are co-resident under programmatic dependent launch -- while layer i computes, layers i+1 and i+2
//     already have their rings filled and are parked in griddepcontrol.wait;
//   * no CTA-wide synchronisation on the critical path: ring fills, the activation slice a warp needs and
//     the step loop are warp-private (cp.async.wait_group + __syncwarp); the CTA meets once, for the
//     final reduction;
//   * sibling layers that share their input (q/k/v, gate/up) run as ONE launch (b200q_linear_group):
//     the n-tiles of up to three layers are laid side by side in one grid;
//   * a CTA may walk several adjacent n-tiles (narrow-tile layouts), keeping the grid at one co-resident
//     wave for any N.
//
// Math (shared with gemv_rp.cu through rp_layouts.cuh): packed nibbles go to mma.sync as fp16 subnormals,
// y = sum_g s_g (sum_k q_k x_k - z_g sum_k x_k), fp32 from the integer product on; split-K across the 8
// warps (shared memory) and across a thread-block cluster (st.async into rank 0), fixed order.
// Replaces ort_ops.gemv (dq_gemv.cu:40-177), gemm_forward_cuda at M<=8 (gemm_cuda_gen.cu:31-353) and
// Marlin at M<=8 (marlin_cuda_kernel.cu:222-733); checkpoint bytes are consumed in place.
#include "gemv_stream.cuh"

namespace b200q {

// one k-step of packed words of tile column n0 -> ring slot (layout's swizzled placement)
template <class T>
__device__ __forceinline__ void st_issue_step(char* slot, const LayerView& L, int s, int n0, int ncols, int lane) {
  constexpr int CPR = T::ROW_WORDS / 4;                       // 16-byte chunks per packed row
  constexpr int CHUNKS = T::ROWS_PER_STEP * CPR;
  static_assert(CHUNKS % 32 == 0, "a step is a whole number of warp-wide copies");
#pragma unroll
  for (int i = 0; i < CHUNKS / 32; ++i) {
    const int idx = lane + 32 * i, r = idx / CPR, cc = idx % CPR;
    if (cc * T::COLS_PER_CHUNK < ncols)
      cp_async16(slot + T::smem_chunk_byte(r, cc), L.qw + T::src_word(L, s * T::ROWS_PER_STEP + r, n0) + 4 * cc);
  }
}

template <class T, int MC>
__global__ void __launch_bounds__(kRpThreads, 2) gemv_stream_kernel(const __grid_constant__ StParams p) {
  extern __shared__ __align__(128) char smem[];
  constexpr int NT = T::NT;
  constexpr int STEP_BYTES = T::ROWS_PER_STEP * T::RS_WORDS * 4;
  char* xs = smem + p.off_x;
  float2* tab = reinterpret_cast<float2*>(smem + p.off_tab);
  float* red = reinterpret_cast<float*>(smem + p.off_red);
  float* rbuf = reinterpret_cast<float*>(smem + p.off_rbuf);
  uint64_t* rbar = reinterpret_cast<uint64_t*>(smem + p.off_rbar);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = p.cluster, D = p.depth;
  const int rank = (int)cluster_ctarank();
  const int cta = blockIdx.x / cs;
  int j = 0;
  if (p.n_layers > 1 && cta >= p.layer[1].cta0) j = 1;
  if (p.n_layers > 2 && cta >= p.layer[2].cta0) j = 2;
  const StLayer& SL = p.layer[j];
  LayerView L;
  L.layout = p.layout; L.bits = p.bits; L.group = p.group; L.K = p.K; L.N = SL.N; L.G = p.G; L.zero_bias = p.zero_bias;
  L.qw = SL.qw; L.qz = SL.qz; L.s = SL.s; L.g_idx = nullptr; L.bias = SL.bias;
  const int tiles_l = (L.N + NT - 1) / NT;
  const int t0 = (cta - SL.cta0) * p.tpc;
  const int nt = min(p.tpc, tiles_l - t0);                 // n-tiles this CTA walks
  const int n0 = t0 * NT;
  const int ncols_cta = min(nt * NT, L.N - n0);

  // K split: unit u (= cluster rank * 8 + warp) owns steps [u q + min(u, r), (u + 1) q + min(u + 1, r)), q = S / U, r = S % U
  const int unit = rank * kWarps + warp;
  const int s_begin = unit * p.split_q + min(unit, p.split_r), s_end = (unit + 1) * p.split_q + min(unit + 1, p.split_r);
  const int cta_s0 = rank * kWarps * p.split_q + min(rank * kWarps, p.split_r);
  const int cta_s1 = (rank + 1) * kWarps * p.split_q + min((rank + 1) * kWarps, p.split_r);
  const int k_cta0 = cta_s0 * T::KSTEP, k_cta1 = cta_s1 * T::KSTEP;
  RpCtx cx;
  cx.M = p.M; cx.k_cta0 = k_cta0; cx.x_stride = p.x_stride; cx.group = p.group; cx.gshift = p.group_shift;
  cx.tab = tab; cx.xs = xs;
  cx.g_first = group_of_k(cx, k_cta0);
  const int g_count = (k_cta1 > k_cta0) ? group_of_k(cx, k_cta1 - 1) - cx.g_first + 1 : 0;

  ST_STAMP(0);
  pdl_launch_dependents();
  if (cs > 1) {      // rank 0 arms the mbarrier that counts the partial vectors arriving through st.async
    if (rank == 0 && tid == 0) {
      mbar_init(rbar, 1);
      mbar_expect_tx(rbar, (uint32_t)(cs - 1) * (uint32_t)(nt * NT * p.M) * 4u);
      fence_mbar_init();
    }
    cluster_arrive_relaxed();
  }

  // ---- 1. fill this warp's ring (weights do not depend on the upstream kernel) ----
  const int nsw = s_end - s_begin, total = nsw * nt;       // this warp's step sequence: tile-major, then k
  char* ringw = smem + p.off_ring + (size_t)warp * D * STEP_BYTES;
  int ii = 0, t_i = 0, s_i = s_begin;                      // issue cursor
  auto issue_next = [&](int slot) {
    if (ii < total) {
      const int tn0 = n0 + t_i * NT;
      st_issue_step<T>(ringw + (size_t)slot * STEP_BYTES, L, s_i, tn0, min(NT, L.N - tn0), lane);
      ++ii;
      if (++s_i == s_end) { s_i = s_begin; ++t_i; }
    }
    cp_async_commit();                                      // always: keeps the group count uniform
  };
  for (int d = 0; d < D; ++d) issue_next(d);

  // per-(group, column) (scale, zero) pairs of the CTA's k-slice, fp32: [tile][group][NT]
  {
    constexpr int TX = NT / 8, TY = kRpThreads / TX;        // a thread owns one 8-column block; rows (tile, group) strided by TY
    const int c8 = (tid % TX) * 8;
    for (int t = 0; t < nt; ++t) {
      const int n = n0 + t * NT + c8;
      for (int gl = tid / TX; gl < g_count; gl += TY) {
        float2 e[8];
        if (n < L.N) T::table_entries8(L, cx.g_first + gl, n, e);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = make_float2(0.f, 0.f);
        }
        float4* dst = reinterpret_cast<float4*>(tab + ((size_t)t * p.gcap + gl) * NT + c8);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(e[2 * i].x, e[2 * i].y, e[2 * i + 1].x, e[2 * i + 1].y);
      }
    }
  }
  __syncthreads();                                          // table visible; still ahead of the dependency
  ST_STAMP(1);

  // ---- 2. activations of this warp's k-range (produced by the upstream kernel) ----
  pdl_wait();
  ST_STAMP(2);
  {
    const int kw0 = s_begin * T::KSTEP, nv = (nsw * T::KSTEP) >> 3;
    for (int m = 0; m < (MC == 1 ? 1 : p.M); ++m)
      for (int v = lane; v < nv; v += 32)
        cp_async16(xs + (size_t)m * p.x_stride + (size_t)(kw0 - k_cta0) * 2 + 16 * v, p.x + (size_t)m * p.ldx + kw0 + 8 * v);
    cp_async_commit();
  }

  float tot[T::NTOT][4], acc[T::NACC][4], accS[4] = {0.f, 0.f, 0.f, 0.f};
  zero4(tot);
  zero4(acc);
  int gcur = (nsw > 0) ? group_of_k(cx, T::step_k(s_begin, lane)) : 0;
  int t_c = 0, s_c = s_begin;
  const int ms = p.M;
  float* redw = red + (size_t)warp * p.red_stride;

  for (int ci = 0; ci < total; ++ci) {
    if (ci == 0) cp_async_wait<0>(); else cp_async_wait_ring(D);
    __syncwarp();
    const int slot = ci & (D - 1);
    typename T::Step w;
    T::load_smem(w, reinterpret_cast<const uint32_t*>(ringw), slot, lane);
    const int gi = group_of_k(cx, T::step_k(s_c, lane));
    if (gi != gcur) {                                       // warp-uniform
      T::template group_end<MC>(tot, acc, accS, cx, gcur - cx.g_first, lane);
      zero4(acc);
      accS[0] = accS[1] = accS[2] = accS[3] = 0.f;
      gcur = gi;
    }
    T::compute(w, s_c, cx, acc, accS, lane);
    __syncwarp();                                           // every lane has read the slot
    issue_next(slot);
    if (++s_c == s_end) {                                   // tile finished: park its partial sums
      T::template group_end<MC>(tot, acc, accS, cx, gcur - cx.g_first, lane);
      T::store_tot(redw + (size_t)t_c * NT * ms, ms, tot, lane, p.M);
      zero4(tot);
      zero4(acc);
      accS[0] = accS[1] = accS[2] = accS[3] = 0.f;
      s_c = s_begin;
      ++t_c;
      cx.tab += (size_t)p.gcap * NT;
      gcur = group_of_k(cx, T::step_k(s_begin, lane));
    }
  }
  if (total == 0) {                                         // more warps than steps: contribute zeros
    cp_async_wait<0>();
    for (int t = 0; t < nt; ++t) T::store_tot(redw + (size_t)t * NT * ms, ms, tot, lane, p.M);
  }
  ST_STAMP(4);

  // ---- 3. reduce: warps -> CTA (shared), CTAs of the cluster -> rank 0 (st.async), store ----
  st_reduce_store<MC, 128>(p, SL, red, rbuf, rbar, nt * NT, ncols_cta, n0, cs, rank, tid);
}

// ------------------------------------------------------------------------------------------------
// Lean AWQ-GEMM instantiation of the same design.  At B200's HBM-to-issue ratio the decode path is bound by
// instructions per weight (an SM must consume ~46 int4 weights per cycle to keep up with HBM), so this kernel
// strips the step loop to the arithmetic the layout needs (per 16 k x 128 columns and lane: 4 LDS.64 + 2 LDS.32,
// 8 PRMT, 8 SHF, 32 LOP3, 9 HMMA, 2 LDGSTS): ring depth is a template constant and the loop is unrolled over the
// ring, so every shared-memory address is base + immediate; source pointers advance by a constant; group and
// tile boundaries are count-downs; lanes with no activation row read a zeroed pad instead of being predicated.
// A CTA walks up to two adjacent 128-column tiles (keeps big-N launches at one co-resident wave).
// ------------------------------------------------------------------------------------------------
template <int MC, int D>
__global__ void __launch_bounds__(kRpThreads, 2) gemv_awq_lean_kernel(const __grid_constant__ StParams p) {
  extern __shared__ __align__(128) char smem[];
  using T = RpAwq;
  constexpr int NT = 128, STEP_BYTES = 1024, KSTEP = 16;
  char* xs = smem + p.off_x;
  float2* tab = reinterpret_cast<float2*>(smem + p.off_tab);
  float* red = reinterpret_cast<float*>(smem + p.off_red);
  float* rbuf = reinterpret_cast<float*>(smem + p.off_rbuf);
  uint64_t* rbar = reinterpret_cast<uint64_t*>(smem + p.off_rbar);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int cs = p.cluster;
  const int rank = (int)cluster_ctarank();
  const int cta = blockIdx.x / cs;
  int j = 0;
  if (p.n_layers > 1 && cta >= p.layer[1].cta0) j = 1;
  if (p.n_layers > 2 && cta >= p.layer[2].cta0) j = 2;
  const StLayer& SL = p.layer[j];
  LayerView L;
  L.layout = p.layout; L.bits = p.bits; L.group = p.group; L.K = p.K; L.N = SL.N; L.G = p.G; L.zero_bias = p.zero_bias;
  L.qw = SL.qw; L.qz = SL.qz; L.s = SL.s; L.g_idx = nullptr; L.bias = SL.bias;
  const int tiles_l = (L.N + NT - 1) / NT;
  const int t0 = (cta - SL.cta0) * p.tpc;
  const int nt = min(p.tpc, tiles_l - t0);
  const int n0 = t0 * NT;
  const int ncols_cta = min(nt * NT, L.N - n0);

  // K split: unit u (= cluster rank * 8 + warp) owns steps [u q + min(u, r), (u + 1) q + min(u + 1, r)), q = S / U, r = S % U
  const int unit = rank * kWarps + warp;
  const int s_begin = unit * p.split_q + min(unit, p.split_r), s_end = (unit + 1) * p.split_q + min(unit + 1, p.split_r);
  const int cta_s0 = rank * kWarps * p.split_q + min(rank * kWarps, p.split_r);
  const int cta_s1 = (rank + 1) * kWarps * p.split_q + min((rank + 1) * kWarps, p.split_r);
  const int k_cta0 = cta_s0 * KSTEP, k_cta1 = cta_s1 * KSTEP;
  const int gsh = p.group_shift;
  const int g_first = k_cta0 >> gsh;
  const int g_count = (k_cta1 > k_cta0) ? ((k_cta1 - 1) >> gsh) - g_first + 1 : 0;
  const int nsw = s_end - s_begin, total = nsw * nt;

  ST_STAMP(0);
  pdl_launch_dependents();
  if (cs > 1) {
    if (rank == 0 && tid == 0) {
      mbar_init(rbar, 1);
      mbar_expect_tx(rbar, (uint32_t)(cs - 1) * (uint32_t)(nt * NT * p.M) * 4u);
      fence_mbar_init();
    }
    cluster_arrive_relaxed();
  }

  // ---- per-lane constants ----
  const uint32_t ring = smem_u32(smem + p.off_ring) + (uint32_t)warp * (D * STEP_BYTES);
  // reads (RpAwq::load_smem): rows 2t, 2t+1 are one 128-byte unit; 16-byte chunk l of unit u sits at l ^ 2(u & 3)
  const int cq = g >> 1;
  const uint32_t rdA = ring + (uint32_t)(((t * 32 + 2 * (g & 1)) + 4 * (cq ^ (2 * t))) * 4);
  const uint32_t rdB = ring + (uint32_t)(((t * 32 + 2 * (g & 1)) + 4 * ((4 + cq) ^ (2 * t))) * 4);
  // copies: chunk `lane` = (row lane / 4, 16-byte column lane % 4) and the same 8 rows further down (+512 B)
  const int r0 = lane >> 2, cc = lane & 3;
  const uint32_t wr = ring + (uint32_t)T::smem_chunk_byte(r0, cc);
  const size_t pitch = (size_t)(L.N >> 3);                               // words per k row
  const size_t src_hi = 8 * pitch, src_step = 16 * pitch;
  const uint32_t* src = L.qw + (size_t)(s_begin * KSTEP + r0) * pitch + (size_t)(n0 >> 3) + 4 * cc;
  bool pc = cc * 32 < min(NT, L.N - n0);
  int irem = nsw, itiles = nt;                                           // issue cursor: steps left in its tile, tiles left
  auto issue = [&](uint32_t dst) {
    if (itiles > 0) {
      if (pc) {
        cp_async16_s(dst, src);
        cp_async16_s(dst + 512, src + src_hi);
      }
      src += src_step;
      if (--irem == 0) {                                                 // next 128-column tile, back to this warp's first k
        irem = nsw;
        --itiles;
        src += 16 - (ptrdiff_t)((size_t)nsw * src_step);
        pc = cc * 32 < L.N - (n0 + (nt - itiles) * NT);
      }
    }
    cp_async_commit();
  };
  if (nsw == 0) itiles = 0;
#pragma unroll 1
  for (int d = 0; d < D; ++d) issue(wr + d * STEP_BYTES);

  // (scale, zero) table of the CTA's k-slice: [tile][group][NT] float2; zero pad for lanes without an activation row
  {
    const int c8 = (tid & 15) * 8;                         // a thread owns one 8-column block; groups strided by 16
    for (int tt = 0; tt < nt; ++tt) {
      const int n = n0 + tt * NT + c8;
      for (int gl = tid >> 4; gl < g_count; gl += 16) {
        float2 e[8];
        if (n < L.N) T::table_entries8(L, g_first + gl, n, e);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = make_float2(0.f, 0.f);
        }
        float4* dst = reinterpret_cast<float4*>(tab + ((size_t)tt * p.gcap + gl) * NT + c8);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(e[2 * i].x, e[2 * i].y, e[2 * i + 1].x, e[2 * i + 1].y);
      }
    }
  }
  uint32_t* zpad = reinterpret_cast<uint32_t*>(smem + p.off_zpad);
  if (tid < 16) zpad[tid] = 0u;
  __syncthreads();
  ST_STAMP(1);

  // ---- activations of this warp's k-range ----
  pdl_wait();
  ST_STAMP(2);
  {
    const int kw0 = s_begin * KSTEP, nv = (nsw * KSTEP) >> 3;
    for (int m = 0; m < (MC == 1 ? 1 : p.M); ++m)
      for (int v = lane; v < nv; v += 32)
        cp_async16(xs + (size_t)m * p.x_stride + (size_t)(kw0 - k_cta0) * 2 + 16 * v, p.x + (size_t)m * p.ldx + kw0 + 8 * v);
    cp_async_commit();
  }
  // B fragment source: row g of the staged activations (k-slots 2t, 2t+1 | +8), or the zero pad
  uint32_t xp = (g < p.M) ? smem_u32(xs) + (uint32_t)(g * p.x_stride + (s_begin * KSTEP - k_cta0 + 2 * t) * 2) : smem_u32(zpad);
  const uint32_t xstep = (g < p.M) ? 32u : 0u, xrewind = (g < p.M) ? (uint32_t)(nsw * 32) : 0u;

  float tot[8][4], acc[8][4], accS[4] = {0.f, 0.f, 0.f, 0.f};
  zero4(tot);
  zero4(acc);
  const int spg = p.group >> 4;                                          // steps per group
  const int gl0 = ((s_begin * KSTEP) >> gsh) - g_first;
  const int gleft0 = spg - (s_begin & (spg - 1));
  int gleft = gleft0, crem = nsw, tcur = 0;
  const float2* tabrow = tab + (size_t)gl0 * NT + 16 * g;
  float* redw = red + (size_t)warp * p.red_stride;
  const int ms = p.M;

  auto group_close = [&]() {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int n = (q >> 2) * 8 + 2 * (q & 3);
      fixup<MC>(tot[q], acc[q], accS, tabrow[n], tabrow[n + 1], (q & 1) ? (kTwo24 / 16.f) : kTwo24);
    }
    zero4(acc);
    accS[0] = accS[1] = accS[2] = accS[3] = 0.f;
  };

  cp_async_wait<0>();
  __syncwarp();
  // One rolled loop (the body must stay resident in the instruction cache: an unrolled-by-D version measured most of
  // its stall cycles on instruction fetch); the slot's shared-memory addresses are base + slot * 1 KB.
#pragma unroll 1
  for (int i = 0; i < total; ++i) {
    const uint32_t so = (uint32_t)(i & (D - 1)) * STEP_BYTES;
    cp_async_wait<D - 1>();
    __syncwarp();
    const uint2 wr0 = lds64_s(rdA + so), wr1 = lds64_s(rdB + so);
    const uint2 wr2 = lds64_s(rdA + so + 512), wr3 = lds64_s(rdB + so + 512);
    const uint32_t b0 = lds32_s(xp), b1 = lds32_s(xp + 16);
    xp += xstep;
#pragma unroll
    for (int wc = 0; wc < 2; ++wc) {
      const uint32_t wa = wc ? wr0.y : wr0.x, wb = wc ? wr1.y : wr1.x;
      const uint32_t wcw = wc ? wr2.y : wr2.x, wd = wc ? wr3.y : wr3.x;
      const uint32_t u01 = prmt(wa, wb, 0x5410), v01 = prmt(wa, wb, 0x7632);
      const uint32_t u23 = prmt(wcw, wd, 0x5410), v23 = prmt(wcw, wd, 0x7632);
      const uint32_t u01h = u01 >> 8, v01h = v01 >> 8, u23h = u23 >> 8, v23h = v23 >> 8;
      mma_16816(acc[wc * 4 + 0], u01 & LO4, v01 & LO4, u23 & LO4, v23 & LO4, b0, b1);
      mma_16816(acc[wc * 4 + 1], u01 & HI4, v01 & HI4, u23 & HI4, v23 & HI4, b0, b1);
      mma_16816(acc[wc * 4 + 2], u01h & LO4, v01h & LO4, u23h & LO4, v23h & LO4, b0, b1);
      mma_16816(acc[wc * 4 + 3], u01h & HI4, v01h & HI4, u23h & HI4, v23h & HI4, b0, b1);
    }
    mma_16816(accS, ONES, ONES, ONES, ONES, b0, b1);
    __syncwarp();                                                        // every lane has read the slot
    issue(wr + so);
    --crem;
    if (--gleft == 0 || crem == 0) {                                     // group and / or tile boundary (warp-uniform)
      group_close();
      tabrow += NT;
      gleft = spg;
      if (crem == 0) {                                                   // park the tile's partial sums, restart at this warp's first k
        T::store_tot(redw + (size_t)tcur * NT * ms, ms, tot, lane, p.M);
        zero4(tot);
        ++tcur;
        crem = nsw;
        gleft = gleft0;
        tabrow = tab + ((size_t)tcur * p.gcap + gl0) * NT + 16 * g;
        xp -= xrewind;
      }
    }
  }
  if (total == 0) {
    cp_async_wait<0>();
    for (int tt = 0; tt < nt; ++tt) T::store_tot(redw + (size_t)tt * NT * ms, ms, tot, lane, p.M);
  }
  ST_STAMP(4);
  st_reduce_store<MC, 256>(p, SL, red, rbuf, rbar, nt * NT, ncols_cta, n0, cs, rank, tid);
}

// ------------------------------------------------------------------------------------------------
struct StPlan {
  int kind, NT, KSTEP, step_bytes, n_gran;
  int ctas, cluster, tpc, depth, steps_total, group_shift, gcap, x_stride;
  int cta0[kMaxGroupLayers];
  int off_x, off_tab, off_red, off_rbuf, off_rbar, off_ring, off_zpad, smem_bytes;
  bool lean;
};

template <class T>
static void st_fill(StPlan& pl) {
  pl.NT = T::NT; pl.KSTEP = T::KSTEP; pl.n_gran = T::N_GRAN;
  pl.step_bytes = T::ROWS_PER_STEP * T::RS_WORDS * 4;
}

static int g_st_cluster = 0, g_st_depth = 0, g_st_tpc = 0, g_st_target = 120, g_st_ring_kb = 64, g_st_lean = 1;
static unsigned long long* g_st_dbg = nullptr;
static size_t g_st_dbg_cap = 0, g_st_dbg_pos = 0;
void gemv_stream_set_option(int which, int value) {
  if (which == 0) g_st_cluster = value;
  else if (which == 1) g_st_depth = value;
  else if (which == 2) g_st_tpc = value;
  else if (which == 3) g_st_target = value;
  else if (which == 4) g_st_ring_kb = value;
  else if (which == 5) g_st_lean = value;
}
void gemv_stream_set_debug(unsigned long long* buf, size_t cap_entries) { g_st_dbg = buf; g_st_dbg_cap = cap_entries; g_st_dbg_pos = 0; }

static bool st_plan(const LinearArgs* a, int n, StPlan& pl) {
  pl.kind = 0;
  if (n < 1 || n > kMaxGroupLayers) return false;
  const LayerView& L = a[0].L;
  const int M = a[0].M;
  if (M < 1 || M > kMB || L.g_idx != nullptr || L.x_perm != nullptr) return false;
  for (int i = 1; i < n; ++i) {
    const LayerView& B = a[i].L;
    if (B.layout != L.layout || B.bits != L.bits || B.group != L.group || B.K != L.K || B.zero_bias != L.zero_bias ||
        B.g_idx != nullptr || B.x_perm != nullptr || a[i].M != M || a[i].x != a[0].x || a[i].ldx != a[0].ldx)
      return false;
  }
  if (L.layout == B200Q_LAYOUT_GPTQ || L.layout == B200Q_LAYOUT_HQQ) {
    if (L.bits == 2) { pl.kind = 1; st_fill<RpGptq<2>>(pl); }
    else if (L.bits == 4) { pl.kind = 2; st_fill<RpGptq<4>>(pl); }
    else if (L.bits == 8) { pl.kind = 3; st_fill<RpGptq<8>>(pl); }
    else return false;
    if (L.group % (4 * (32 / L.bits)) != 0) return false;   // a step (4 packed rows) lies inside one group
  } else if (L.layout == B200Q_LAYOUT_AWQ_GEMM) {
    if (L.bits != 4 || L.group % 16 != 0) return false;
    pl.kind = 4; st_fill<RpAwq>(pl);
  } else if (L.layout == B200Q_LAYOUT_MARLIN) {
    if (L.bits != 4 || L.group % 16 != 0) return false;
    pl.kind = 5; st_fill<RpMarlin>(pl);
  } else return false;
  if (L.K % pl.KSTEP != 0 || L.K % L.group != 0) { pl.kind = 0; return false; }
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    if (a[i].L.N % pl.n_gran != 0) { pl.kind = 0; return false; }
    tiles += (a[i].L.N + pl.NT - 1) / pl.NT;
  }
  pl.group_shift = -1;
  if ((L.group & (L.group - 1)) == 0) { int s = 0; while ((1 << s) < L.group) ++s; pl.group_shift = s; }
  pl.steps_total = L.K / pl.KSTEP;

  // lean instantiation: AWQ, power-of-two group >= one step
  pl.lean = g_st_lean && pl.kind == 4 && pl.group_shift >= 4;
  // tiles per CTA: widest walk (<= 128 columns; 256 for the lean AWQ kernel at M <= 2) that still leaves >= target CTA groups
  const int max_cols = (pl.lean && M <= 2) ? 256 : 128;
  int tpc = 1;
  for (int c = 2; c * pl.NT <= max_cols; c *= 2)
    if ((tiles + c - 1) / c >= g_st_target) tpc = c;
  if (g_st_tpc > 0 && g_st_tpc * pl.NT <= max_cols) tpc = g_st_tpc;
  int groups = 0;
  for (int i = 0; i < n; ++i) {
    pl.cta0[i] = groups;
    groups += ((a[i].L.N + pl.NT - 1) / pl.NT + tpc - 1) / tpc;
  }
  // cluster (K split): smallest that brings the grid to >= target CTAs while a warp keeps >= 2 steps
  int cs = 1;
  while (cs < 8 && groups * cs < g_st_target && pl.steps_total / ((cs + 1) * kWarps) >= 2) ++cs;
  if (g_st_cluster > 0 && g_st_cluster <= 8 && pl.steps_total / g_st_cluster >= 1) cs = g_st_cluster;
  pl.tpc = tpc; pl.cluster = cs; pl.ctas = groups * cs;
  // ring depth: whole per-warp sequence if it fits the per-CTA ring budget, else the budget
  const int per_sm = (pl.ctas + 147) / 148;
  const int seq = ((pl.steps_total + cs * kWarps - 1) / (cs * kWarps)) * tpc;
  int budget = (g_st_ring_kb * 1024) / (per_sm > 2 ? 2 : per_sm);          // bytes of ring per CTA
  int depth = 2;
  while (depth < 16 && depth < seq && 2 * depth * pl.step_bytes * kWarps <= budget) depth *= 2;
  if (g_st_depth == 2 || g_st_depth == 4 || g_st_depth == 8 || g_st_depth == 16) depth = g_st_depth;
  if (pl.lean) depth = depth < 4 ? 4 : (depth > 8 ? 8 : depth);            // instantiated ring depths
  pl.depth = depth;

  // unit u of the K split owns q (+1 for the first r units) steps, so cluster rank 0 holds the longest slice
  const int split_q = pl.steps_total / (cs * kWarps), split_r = pl.steps_total % (cs * kWarps);
  const int slice_steps = kWarps * split_q + (split_r < kWarps ? split_r : kWarps);
  const int kslice = slice_steps * pl.KSTEP;
  pl.x_stride = kslice * 2;
  pl.x_stride += (64 - (pl.x_stride % 128) + 128) % 128;
  pl.gcap = kslice / L.group + 2;
  int off = 0;
  pl.off_x = off; off += M * pl.x_stride;
  off = (off + 15) & ~15;
  pl.off_tab = off; off += tpc * pl.gcap * pl.NT * 8;
  off = (off + 15) & ~15;
  pl.off_red = off; off += kWarps * tpc * pl.NT * M * 4;
  pl.off_rbuf = off; off += (cs - 1) * tpc * pl.NT * M * 4;
  off = (off + 7) & ~7;
  pl.off_rbar = off; off += 8;
  off = (off + 127) & ~127;
  pl.off_ring = off; off += kWarps * depth * pl.step_bytes;
  pl.off_zpad = off; off += depth * 32 + 64;
  pl.smem_bytes = off;
  if (pl.smem_bytes > 100 * 1024) { pl.kind = 0; return false; }           // two CTAs per SM must fit
  if (pl.ctas > 148 * 2) { pl.kind = 0; return false; }                     // one co-resident wave
  return true;
}

bool gemv_stream_supported(const LinearArgs* a, int n) {
  StPlan pl;
  if (!st_plan(a, n, pl)) return false;
  if (((uintptr_t)a[0].x & 15) != 0 || (a[0].ldx % 8) != 0) return false;
  for (int i = 0; i < n; ++i)
    if (((uintptr_t)a[i].L.qw & 15) != 0 || ((uintptr_t)a[i].L.s & 15) != 0) return false;
  return true;
}

bool gemv_stream_describe(const LinearArgs* a, int n, int out[6]) {
  StPlan pl;
  if (!st_plan(a, n, pl)) return false;
  out[0] = pl.cluster; out[1] = pl.ctas; out[2] = pl.smem_bytes; out[3] = pl.steps_total; out[4] = pl.tpc; out[5] = pl.depth;
  return true;
}

template <class T, int MC>
static cudaError_t st_launch_k(const StParams& p, const StPlan& pl, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemv_stream_kernel<T, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    // all of the SM's unified L1/shared array as shared memory, so that consecutive layers' CTAs can be co-resident
    if (decode_carveout_max()) cudaFuncSetAttribute(gemv_stream_kernel<T, MC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    attr_done[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.ctas);
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
  return cudaLaunchKernelEx(&cfg, gemv_stream_kernel<T, MC>, p);
}

template <int MC, int D>
static cudaError_t st_launch_lean(const StParams& p, const StPlan& pl, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemv_awq_lean_kernel<MC, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    if (decode_carveout_max()) cudaFuncSetAttribute(gemv_awq_lean_kernel<MC, D>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    attr_done[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.ctas);
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
  return cudaLaunchKernelEx(&cfg, gemv_awq_lean_kernel<MC, D>, p);
}

template <class T>
static cudaError_t st_launch_t(const StParams& p, const StPlan& pl, cudaStream_t st) {
  return p.M == 1 ? st_launch_k<T, 1>(p, pl, st) : st_launch_k<T, 2>(p, pl, st);
}

cudaError_t launch_gemv_stream(const LinearArgs* a, int n, const PeerOut* peers) {
  StPlan pl;
  if (!st_plan(a, n, pl)) return cudaErrorInvalidValue;
  const LayerView& L = a[0].L;
  StParams p = {};
  p.n_layers = n;
  for (int i = 0; i < kMaxGroupLayers; ++i) {
    const int k = i < n ? i : n - 1;                        // unused slots mirror the last layer (never selected)
    StLayer& d = p.layer[i];
    d.qw = a[k].L.qw; d.qz = a[k].L.qz; d.s = a[k].L.s; d.bias = a[k].L.bias; d.N = a[k].L.N;
    d.cta0 = i < n ? pl.cta0[i] : (1 << 30);
    if (peers && n == 1) d.out = *peers; else { d.out.n = 1; d.out.y[0] = a[k].y; }
    d.ldy = a[k].ldy; d.n_offset = a[k].n_offset;
  }
  p.layout = L.layout; p.bits = L.bits; p.group = L.group; p.K = L.K; p.G = L.G; p.zero_bias = L.zero_bias;
  p.x = a[0].x; p.ldx = a[0].ldx; p.M = a[0].M;
  p.cluster = pl.cluster; p.tpc = pl.tpc; p.depth = pl.depth; p.steps_total = pl.steps_total; p.group_shift = pl.group_shift;
  p.split_q = pl.steps_total / (pl.cluster * kWarps); p.split_r = pl.steps_total % (pl.cluster * kWarps);
  p.gcap = pl.gcap; p.x_stride = pl.x_stride; p.red_stride = pl.tpc * pl.NT * a[0].M;
  p.off_x = pl.off_x; p.off_tab = pl.off_tab; p.off_red = pl.off_red; p.off_rbuf = pl.off_rbuf; p.off_rbar = pl.off_rbar;
  p.off_ring = pl.off_ring; p.off_zpad = pl.off_zpad;
  p.dbg = nullptr;
  if (g_st_dbg) {
    const size_t need = (size_t)pl.ctas * 8;
    if (g_st_dbg_pos + need <= g_st_dbg_cap) { p.dbg = g_st_dbg + g_st_dbg_pos; g_st_dbg_pos += need; }
  }
  switch (pl.kind) {
    case 1: return st_launch_t<RpGptq<2>>(p, pl, a[0].stream);
    case 2: return st_launch_t<RpGptq<4>>(p, pl, a[0].stream);
    case 3: return st_launch_t<RpGptq<8>>(p, pl, a[0].stream);
    case 4:
      if (pl.lean) {
        if (pl.depth == 4) return p.M == 1 ? st_launch_lean<1, 4>(p, pl, a[0].stream) : st_launch_lean<2, 4>(p, pl, a[0].stream);
        return p.M == 1 ? st_launch_lean<1, 8>(p, pl, a[0].stream) : st_launch_lean<2, 8>(p, pl, a[0].stream);
      }
      return st_launch_t<RpAwq>(p, pl, a[0].stream);
    case 5: return st_launch_t<RpMarlin>(p, pl, a[0].stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace b200q
