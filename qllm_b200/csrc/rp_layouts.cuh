// Layout functors of the decode kernels ("integer-in-subnormal" mma.sync formulation): how a warp reads one
// k-step of packed words (from global or from its shared-memory ring), feeds them to mma.sync and folds the
// per-group (scale, zero).  Shared by gemv_rp.cu (whole-slice prefetch) and gemv_stream.cu (per-warp rings).
// Formats: SURVEY.md Appendix A (compress_weight.py:10-92, quant_linear_awq.py:95-140,
// quant_linear_marlin.py:18-42).
#pragma once
#include "common.cuh"

namespace b200q {

static constexpr int kWarps = 8;
static constexpr int kRpThreads = kWarps * 32;
static constexpr int kMB = 8;
static constexpr uint32_t LO4 = 0x000f000fu, HI4 = 0x00f000f0u, ONES = 0x3C003C00u;
static constexpr float kTwo24 = 16777216.f;

__device__ __forceinline__ uint4 ldg128_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg64_stream(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(const float* local_smem, uint32_t rank, float v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
// remote store that also signals: value -> rank's shared memory, 4 bytes of complete_tx on rank's mbarrier
__device__ __forceinline__ void st_async_f32(const float* local_smem, const uint64_t* local_bar, uint32_t rank, float v) {
  uint32_t ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem)), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(__float_as_uint(v)), "r"(rb)
               : "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t r_lds32(const void* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint2 r_lds64(const void* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ uint4 r_lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }

// D(16x8,f32) += A(16x8,f16,row) * B(8x8,f16,col)
__device__ __forceinline__ void mma_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}

struct RpCtx {
  int M, k_cta0, g_first, x_stride, group, gshift;
  const float2* tab;      // [g_local][NT] (scale, zero) as fp32
  const char* xs;
};
__device__ __forceinline__ int group_of_k(const RpCtx& cx, int k) { return cx.gshift >= 0 ? (k >> cx.gshift) : (k / cx.group); }

template <int N>
__device__ __forceinline__ void zero4(float (&a)[N][4]) {
#pragma unroll
  for (int i = 0; i < N; ++i) { a[i][0] = 0.f; a[i][1] = 0.f; a[i][2] = 0.f; a[i][3] = 0.f; }
}

// tot += s * (mul * acc - z * S) for the c-fragment entries of one accumulator (rows g / g+8 = columns
// n_lo / n_hi; batch columns 2t [, 2t+1 when MC == 2])
template <int MC>
__device__ __forceinline__ void fixup(float (&tot)[4], const float (&acc)[4], const float (&accS)[4], float2 lo, float2 hi,
                                      float mul) {
  tot[0] = fmaf(lo.x, fmaf(acc[0], mul, -lo.y * accS[0]), tot[0]);
  tot[2] = fmaf(hi.x, fmaf(acc[2], mul, -hi.y * accS[0]), tot[2]);
  if (MC == 2) {
    tot[1] = fmaf(lo.x, fmaf(acc[1], mul, -lo.y * accS[1]), tot[1]);
    tot[3] = fmaf(hi.x, fmaf(acc[3], mul, -hi.y * accS[1]), tot[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// GPTQ / HQQ, BITS in {2,4,8}: CTA tile 32 columns; step = 4 packed rows (lane t -> row t), lane g
// holds the words of columns 4g..4g+3.  Sets: {cols 0,1} and {cols 2,3} (MMA rows g and g+8).
// ------------------------------------------------------------------------------------------------
template <int BITS>
struct RpGptq {
  static constexpr int P = 32 / BITS;
  static constexpr int NT = 32, KSTEP = 4 * P, MAXSTEPS = 16, N_GRAN = 32;
  static constexpr int ROWS_PER_STEP = 4, ROW_WORDS = 32, RS_WORDS = 32, SM_MIN_BLOCKS = 3;   // rows unpadded, chunks XOR-swizzled
  static constexpr int COLS_PER_CHUNK = 4, LANE_COLS = 4;       // columns per 16-byte chunk / per lane-g
  static constexpr int NTOT = 2;                               // output accumulators (sets)
  static constexpr int NACC = (BITS == 4) ? 4 : 2;             // 4-bit: {set} x {LO, HI}
  using Step = uint4;

  __device__ static void load(Step& w, const LayerView& L, int s, int n0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    w = ldg128_stream(L.qw + (size_t)(s * 4 + t) * L.N + n0 + 4 * g);
  }
  __device__ static size_t src_word(const LayerView& L, int row, int n0) { return (size_t)row * L.N + n0; }
  __device__ static void load_smem(Step& w, const uint32_t* tile, int ls, int lane) {
    const int g = lane >> 2, t = lane & 3;
    w = r_lds128(tile + (size_t)(ls * 4 + t) * RS_WORDS + 4 * (g ^ (2 * t)));
  }
  // shared-memory byte offset of 16-byte chunk cc of packed row `row`: the four rows a quarter-warp reads together
  // start at the same bank, so chunk c of row r sits at c ^ 2(r & 3) (conflict-free LDS.128 without padding)
  __device__ static int smem_chunk_byte(int row, int cc) { return row * (RS_WORDS * 4) + ((cc ^ (2 * (row & 3))) << 4); }
  static constexpr int SWZ_ROWS = 4;
  // An MMA mixes the k-slots of all four t-lanes, so a step (4 packed rows) must lie inside one group.
  __device__ static int step_k(int s, int) { return s * KSTEP; }
  // (scale, zero) of 8 adjacent columns n..n+7 of group g, as fp32 (n % 8 == 0)
  __device__ static void table_entries8(const LayerView& L, int g, int n, float2 (&e)[8]) {
    const uint4 sv = __ldg(reinterpret_cast<const uint4*>(L.s + (size_t)g * L.N + n));
    const __half* sh = reinterpret_cast<const __half*>(&sv);
    if (L.layout == B200Q_LAYOUT_HQQ) {
      const uint4 zv = __ldg(reinterpret_cast<const uint4*>((const __half*)L.qz + (size_t)g * L.N + n));
      const __half* zh = reinterpret_cast<const __half*>(&zv);
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = make_float2(__half2float(sh[i]), __half2float(zh[i]));
      return;
    }
    const uint32_t* zrow = (const uint32_t*)L.qz + (size_t)g * (((size_t)L.N * BITS) >> 5);
    const int bit0 = n * BITS;                                  // 8 columns = 8*BITS bits: one word (two for 8-bit)
    const uint32_t w0 = __ldg(zrow + (bit0 >> 5));
    const uint32_t w1 = (BITS == 8) ? __ldg(zrow + (bit0 >> 5) + 1) : 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int b = (bit0 & 31) + i * BITS;
      const uint32_t raw = (b < 32) ? (w0 >> b) : (w1 >> (b - 32));
      const uint32_t z = ((raw & ((1u << BITS) - 1u)) + (uint32_t)L.zero_bias) & ((1u << BITS) - 1u);
      e[i] = make_float2(__half2float(sh[i]), (float)z);
    }
  }
  __device__ static float2 table_entry(const LayerView& L, int g, int n) {
    const float sv = __half2float(__ldg(L.s + (size_t)g * L.N + n));
    if (L.layout == B200Q_LAYOUT_HQQ) return make_float2(sv, __half2float(__ldg((const __half*)L.qz + (size_t)g * L.N + n)));
    const int bit = n * BITS;
    const uint32_t zw = __ldg((const uint32_t*)L.qz + (size_t)g * (((size_t)L.N * BITS) >> 5) + (bit >> 5));
    const uint32_t z = (((zw >> (bit & 31)) & ((1u << BITS) - 1u)) + (uint32_t)L.zero_bias) & ((1u << BITS) - 1u);
    return make_float2(sv, (float)z);
  }

  __device__ static void compute(const Step& w, int s, const RpCtx& cx, float (&acc)[NACC][4], float (&accS)[4], int lane) {
    const int g = lane >> 2;
    const int koff = s * KSTEP + P * (lane & 3) - cx.k_cta0;      // first k of this lane's word
    const char* xr = cx.xs + (size_t)g * cx.x_stride + koff * 2;
    const bool act = g < cx.M;
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    if (BITS == 4) {
      uint32_t bl0 = 0, bl1 = 0, bh0 = 0, bh1 = 0;
      if (act) {
        const uint4 v = r_lds128(xr);
        bl0 = prmt(v.x, v.z, 0x5410); bl1 = prmt(v.y, v.w, 0x5410);     // (x0,x4) (x2,x6)  <- LO nibbles k0,k4 | k2,k6
        bh0 = prmt(v.x, v.z, 0x7632); bh1 = prmt(v.y, v.w, 0x7632);     // (x1,x5) (x3,x7)  <- HI nibbles k1,k5 | k3,k7
      }
      uint32_t l0[4], l1[4], h0[4], h1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t hi = ww[i] >> 8;
        l0[i] = ww[i] & LO4; h0[i] = ww[i] & HI4; l1[i] = hi & LO4; h1[i] = hi & HI4;
      }
      mma_16816(acc[0], l0[0], l0[1], l1[0], l1[1], bl0, bl1);
      mma_16816(acc[1], h0[0], h0[1], h1[0], h1[1], bh0, bh1);
      mma_16816(acc[2], l0[2], l0[3], l1[2], l1[3], bl0, bl1);
      mma_16816(acc[3], h0[2], h0[3], h1[2], h1[3], bh0, bh1);
      mma_16816(accS, ONES, ONES, ONES, ONES, bl0, bl1);
      mma_16816(accS, ONES, ONES, ONES, ONES, bh0, bh1);
    } else if (BITS == 8) {
      uint32_t b0 = 0, b1 = 0;
      if (act) { const uint2 v = r_lds64(xr); b0 = v.x; b1 = v.y; }
      uint32_t p0[4], p1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { p0[i] = prmt(ww[i], 0u, 0x4140); p1[i] = prmt(ww[i], 0u, 0x4342); }   // (k0,k1) (k2,k3)
      mma_16816(acc[0], p0[0], p0[1], p1[0], p1[1], b0, b1);
      mma_16816(acc[1], p0[2], p0[3], p1[2], p1[3], b0, b1);
      mma_16816(accS, ONES, ONES, ONES, ONES, b0, b1);
    } else {   // 2-bit: pair i = (k_i, k_{i+8}), all brought to bit 0 by shifts
      uint32_t xb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xb[j] = 0u;
      if (act) {
        const uint4 a = r_lds128(xr), b = r_lds128(xr + 16);
        xb[0] = prmt(a.x, b.x, 0x5410); xb[1] = prmt(a.x, b.x, 0x7632);
        xb[2] = prmt(a.y, b.y, 0x5410); xb[3] = prmt(a.y, b.y, 0x7632);
        xb[4] = prmt(a.z, b.z, 0x5410); xb[5] = prmt(a.z, b.z, 0x7632);
        xb[6] = prmt(a.w, b.w, 0x5410); xb[7] = prmt(a.w, b.w, 0x7632);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        uint32_t e0[4], e1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          e0[i] = (ww[i] >> (4 * m)) & 0x00030003u;
          e1[i] = (ww[i] >> (4 * m + 2)) & 0x00030003u;
        }
        mma_16816(acc[0], e0[0], e0[1], e1[0], e1[1], xb[2 * m], xb[2 * m + 1]);
        mma_16816(acc[1], e0[2], e0[3], e1[2], e1[3], xb[2 * m], xb[2 * m + 1]);
        mma_16816(accS, ONES, ONES, ONES, ONES, xb[2 * m], xb[2 * m + 1]);
      }
    }
  }

  template <int MC>
  __device__ static void group_end(float (&tot)[NTOT][4], float (&acc)[NACC][4], float (&accS)[4], const RpCtx& cx, int gl,
                                   int lane) {
    const int g = lane >> 2;
    const float2* tb = cx.tab + (size_t)gl * NT + 4 * g;
    const float2 c0 = tb[0], c1 = tb[1], c2 = tb[2], c3 = tb[3];
    if (BITS == 4) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = fmaf(acc[1][i], 0.0625f, acc[0][i]); b[i] = fmaf(acc[3][i], 0.0625f, acc[2][i]); }
      fixup<MC>(tot[0], a, accS, c0, c1, kTwo24);
      fixup<MC>(tot[1], b, accS, c2, c3, kTwo24);
    } else {
      fixup<MC>(tot[0], acc[0], accS, c0, c1, kTwo24);
      fixup<MC>(tot[1], acc[1], accS, c2, c3, kTwo24);
    }
  }

  __device__ static void store_tot(float* r, int mstride, const float (&tot)[NTOT][4], int lane, int M) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int s = 0; s < NTOT; ++s) {
      const int n = 4 * g + 2 * s;
      if (2 * t < M) { r[n * mstride + 2 * t] = tot[s][0]; r[(n + 1) * mstride + 2 * t] = tot[s][2]; }
      if (2 * t + 1 < M) { r[n * mstride + 2 * t + 1] = tot[s][1]; r[(n + 1) * mstride + 2 * t + 1] = tot[s][3]; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// AWQ GEMM layout: CTA tile 128 columns (16 words = 64 B per k row); step = 16 k rows; lane (g,t)
// holds word columns 2g,2g+1 at rows 2t, 2t+1, 8+2t, 9+2t.  prmt pairs two rows so that one AND yields
// (n@k, n@k+1).  MMA j of a word covers columns 2j (rows g) and 2j+1 (rows g+8); odd j carry 16*q.
// ------------------------------------------------------------------------------------------------
struct RpAwq {
  static constexpr int NT = 128, KSTEP = 16, MAXSTEPS = 4, N_GRAN = 32;
  static constexpr int ROWS_PER_STEP = 16, ROW_WORDS = 16, RS_WORDS = 16, SM_MIN_BLOCKS = 2;   // unpadded, XOR-swizzled
  static constexpr int COLS_PER_CHUNK = 32, LANE_COLS = 16;
  static constexpr int NTOT = 8, NACC = 8;
  struct Step { uint2 r[4]; };

  __device__ static void load(Step& w, const LayerView& L, int s, int n0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const size_t rs = (size_t)(L.N >> 3);
    const uint32_t* base = L.qw + (size_t)(s * 16 + 2 * t) * rs + (n0 >> 3) + 2 * g;
    w.r[0] = ldg64_stream(base);
    w.r[1] = ldg64_stream(base + rs);
    w.r[2] = ldg64_stream(base + 8 * rs);
    w.r[3] = ldg64_stream(base + 9 * rs);
  }
  __device__ static size_t src_word(const LayerView& L, int row, int n0) { return (size_t)row * (L.N >> 3) + (n0 >> 3); }
  __device__ static void load_smem(Step& w, const uint32_t* tile, int ls, int lane) {
    const int g = lane >> 2, t = lane & 3;
    // rows 2t, 2t+1 form one 128-byte unit (8 chunks of 16 B); chunk l of unit u sits at l ^ 2(u & 3)
    const uint32_t* unit = tile + (size_t)(ls * 8 + t) * 32 + 2 * (g & 1);
    const int c = g >> 1, x = 2 * t;
    w.r[0] = r_lds64(unit + 4 * (c ^ x));
    w.r[1] = r_lds64(unit + 4 * ((4 + c) ^ x));
    w.r[2] = r_lds64(unit + 4 * 32 + 4 * (c ^ x));            // rows +8 / +9: unit + 4 (same u & 3)
    w.r[3] = r_lds64(unit + 4 * 32 + 4 * ((4 + c) ^ x));
  }
  __device__ static int smem_chunk_byte(int row, int cc) {
    const int u = row >> 1, l = ((row & 1) << 2) + cc;
    return u * 128 + ((l ^ (2 * (u & 3))) << 4);
  }
  static constexpr int SWZ_ROWS = 8;
  __device__ static int step_k(int s, int) { return s * KSTEP; }
  __device__ static void table_entries8(const LayerView& L, int g, int n, float2 (&e)[8]) {
    const uint4 sv = __ldg(reinterpret_cast<const uint4*>(L.s + (size_t)g * L.N + n));
    const __half* sh = reinterpret_cast<const __half*>(&sv);
    const uint32_t zw = __ldg((const uint32_t*)L.qz + (size_t)g * (L.N >> 3) + (n >> 3));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int nib = (i >> 1) + ((i & 1) << 2);
      e[i] = make_float2(__half2float(sh[i]), (float)((((zw >> (4 * nib)) & 0xFu) + (uint32_t)L.zero_bias) & 0xFu));
    }
  }
  __device__ static float2 table_entry(const LayerView& L, int g, int n) {
    const float sv = __half2float(__ldg(L.s + (size_t)g * L.N + n));
    const uint32_t zw = __ldg((const uint32_t*)L.qz + (size_t)g * (L.N >> 3) + (n >> 3));
    const int i = n & 7, nib = (i >> 1) + ((i & 1) << 2);
    const uint32_t z = (((zw >> (4 * nib)) & 0xFu) + (uint32_t)L.zero_bias) & 0xFu;
    return make_float2(sv, (float)z);
  }

  __device__ static void compute(const Step& w, int s, const RpCtx& cx, float (&acc)[NACC][4], float (&accS)[4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    uint32_t b0 = 0u, b1 = 0u;
    if (g < cx.M) {
      const char* xr = cx.xs + (size_t)g * cx.x_stride + (size_t)(s * KSTEP - cx.k_cta0 + 2 * t) * 2;
      b0 = r_lds32(xr);
      b1 = r_lds32(xr + 16);
    }
#pragma unroll
    for (int wc = 0; wc < 2; ++wc) {
      const uint32_t wa = wc ? w.r[0].y : w.r[0].x, wb = wc ? w.r[1].y : w.r[1].x;
      const uint32_t wcw = wc ? w.r[2].y : w.r[2].x, wd = wc ? w.r[3].y : w.r[3].x;
      const uint32_t u01 = prmt(wa, wb, 0x5410), v01 = prmt(wa, wb, 0x7632);
      const uint32_t u23 = prmt(wcw, wd, 0x5410), v23 = prmt(wcw, wd, 0x7632);
      const uint32_t u01h = u01 >> 8, v01h = v01 >> 8, u23h = u23 >> 8, v23h = v23 >> 8;
      mma_16816(acc[wc * 4 + 0], u01 & LO4, v01 & LO4, u23 & LO4, v23 & LO4, b0, b1);       // cols 0,1
      mma_16816(acc[wc * 4 + 1], u01 & HI4, v01 & HI4, u23 & HI4, v23 & HI4, b0, b1);       // cols 2,3 (x16)
      mma_16816(acc[wc * 4 + 2], u01h & LO4, v01h & LO4, u23h & LO4, v23h & LO4, b0, b1);   // cols 4,5
      mma_16816(acc[wc * 4 + 3], u01h & HI4, v01h & HI4, u23h & HI4, v23h & HI4, b0, b1);   // cols 6,7 (x16)
    }
    mma_16816(accS, ONES, ONES, ONES, ONES, b0, b1);
  }

  template <int MC>
  __device__ static void group_end(float (&tot)[NTOT][4], float (&acc)[NACC][4], float (&accS)[4], const RpCtx& cx, int gl,
                                   int lane) {
    const int g = lane >> 2;
    const float2* tb = cx.tab + (size_t)gl * NT + 16 * g;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int n = (q >> 2) * 8 + 2 * (q & 3);
      fixup<MC>(tot[q], acc[q], accS, tb[n], tb[n + 1], (q & 1) ? (kTwo24 / 16.f) : kTwo24);
    }
  }

  __device__ static void store_tot(float* r, int mstride, const float (&tot)[NTOT][4], int lane, int M) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int q = 0; q < NTOT; ++q) {
      const int n = 16 * g + (q >> 2) * 8 + 2 * (q & 3);
      if (2 * t < M) { r[n * mstride + 2 * t] = tot[q][0]; r[(n + 1) * mstride + 2 * t] = tot[q][2]; }
      if (2 * t + 1 < M) { r[n * mstride + 2 * t + 1] = tot[q][1]; r[(n + 1) * mstride + 2 * t + 1] = tot[q][3]; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Marlin layout: CTA tile 64 columns (one 128-word block per k16 row); step = one k16 row; a lane's
// word j is its A fragment of the 16-column tile j: low nibbles = k-slots 2t,2t+1, x16 nibbles =
// k-slots 2t+8,2t+9 -> two m16n8k8 MMAs with separate accumulators.
// ------------------------------------------------------------------------------------------------
struct RpMarlin {
  static constexpr int NT = 64, KSTEP = 16, MAXSTEPS = 16, N_GRAN = 64;
  static constexpr int ROWS_PER_STEP = 1, ROW_WORDS = 128, RS_WORDS = 128, SM_MIN_BLOCKS = 3;
  static constexpr int COLS_PER_CHUNK = 2, LANE_COLS = 0;       // 4 words = 2 columns per chunk; tiles are always whole
  static constexpr int NTOT = 4, NACC = 8;
  using Step = uint4;

  __device__ static void load(Step& w, const LayerView& L, int s, int n0, int lane) {
    w = ldg128_stream(L.qw + (size_t)s * (2 * (size_t)L.N) + 2 * n0 + 4 * lane);
  }
  __device__ static size_t src_word(const LayerView& L, int row, int n0) { return (size_t)row * (2 * (size_t)L.N) + 2 * n0; }
  __device__ static void load_smem(Step& w, const uint32_t* tile, int ls, int lane) { w = r_lds128(tile + (size_t)ls * RS_WORDS + 4 * lane); }
  __device__ static int smem_chunk_byte(int row, int cc) { return row * (RS_WORDS * 4) + (cc << 4); }
  static constexpr int SWZ_ROWS = 1;
  __device__ static int step_k(int s, int) { return s * KSTEP; }
  __device__ static float2 table_entry(const LayerView& L, int g, int n) {
    return make_float2(__half2float(__ldg(L.s + (size_t)g * L.N + marlin_scale_index(n, L.group == L.K))), 8.0f);
  }
  __device__ static void table_entries8(const LayerView& L, int g, int n, float2 (&e)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = table_entry(L, g, n + i);
  }

  __device__ static void compute(const Step& w, int s, const RpCtx& cx, float (&acc)[NACC][4], float (&accS)[4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    uint32_t b0 = 0u, b1 = 0u;
    if (g < cx.M) {
      const char* xr = cx.xs + (size_t)g * cx.x_stride + (size_t)(s * KSTEP - cx.k_cta0 + 2 * t) * 2;
      b0 = r_lds32(xr);
      b1 = r_lds32(xr + 16);
    }
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t lo = ww[j], hi = ww[j] >> 8;
      mma_1688(acc[2 * j], lo & LO4, hi & LO4, b0);          // rows g | g+8, k-slots 2t,2t+1
      mma_1688(acc[2 * j + 1], lo & HI4, hi & HI4, b1);      // k-slots 2t+8,2t+9 (x16)
    }
    mma_16816(accS, ONES, ONES, ONES, ONES, b0, b1);
  }

  template <int MC>
  __device__ static void group_end(float (&tot)[NTOT][4], float (&acc)[NACC][4], float (&accS)[4], const RpCtx& cx, int gl,
                                   int lane) {
    const int g = lane >> 2;
    const float2* tb = cx.tab + (size_t)gl * NT;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = fmaf(acc[2 * j + 1][i], 0.0625f, acc[2 * j][i]);
      fixup<MC>(tot[j], a, accS, tb[16 * j + g], tb[16 * j + g + 8], kTwo24);
    }
  }

  __device__ static void store_tot(float* r, int mstride, const float (&tot)[NTOT][4], int lane, int M) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = 16 * j + g;
      if (2 * t < M) { r[n * mstride + 2 * t] = tot[j][0]; r[(n + 8) * mstride + 2 * t] = tot[j][2]; }
      if (2 * t + 1 < M) { r[n * mstride + 2 * t + 1] = tot[j][1]; r[(n + 8) * mstride + 2 * t + 1] = tot[j][3]; }
    }
  }
};

}  // namespace b200q
