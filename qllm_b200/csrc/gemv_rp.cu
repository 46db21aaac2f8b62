// Decode kernel v2 ("register prefetch"): latency-optimised M<=8 path.
//
// A decode-sized layer streams 8-23 MB in ~1.3-3.5 us at HBM speed, so the per-kernel dependent
// chain (wait for the previous layer's y -> fetch x -> compute -> reduce -> store) costs as much as
// the streaming itself.  This kernel keeps that chain short and takes the weight traffic off it:
//
//   1. every thread issues ALL of its packed-weight loads (<=16 x 128/64-bit, coalesced along N,
//      L1 no-allocate) into registers, and the CTA stages its group scales/zeros in shared memory,
//      BEFORE griddepcontrol.wait -- weights do not depend on the previous kernel, so with
//      programmatic dependent launch the next layers' weights stream in while this layer computes;
//   2. after the wait: x slice -> shared (one sync), unpack in registers (lop3/prmt), (q-z)*s in
//      fp16x2, mma.sync.m16n8k16 with the weights as the 16-row A operand (fp32 accumulate);
//   3. the 8 warps of a CTA split K and reduce through shared memory; CTAs that split K further
//      form a thread-block cluster and reduce through distributed shared memory (no global scratch,
//      no atomics, fixed summation order -> deterministic).
//
// Layout handling (no repacking of checkpoint bytes) is as in gemv_mma.cu; see the notes there.
#include "common.cuh"
#include "kernels.h"

namespace b200q {

static constexpr int kWarps = 8;
static constexpr int kRpThreads = kWarps * 32;
static constexpr int kMB = 8;
static constexpr uint32_t MAGIC = 0x64006400u;
static constexpr uint32_t LO4 = 0x000f000fu, HI4 = 0x00f000f0u;
static constexpr uint32_t H_1_4 = 0x34003400u, H_1_16 = 0x2c002c00u, H_1_64 = 0x24002400u;

struct RpParams {
  LayerView L;
  const __half* x;
  int64_t ldx;
  int M;
  PeerOut out;
  int64_t ldy, n_offset;
  int n_tiles, cluster, steps_total;
  int x_stride;                   // bytes
  int off_x, off_sc, off_zq, off_red, off_rbuf, off_w;
};

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
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t r_lds32(const void* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint2 r_lds64(const void* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ uint4 r_lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }

struct RpCtx {
  const LayerView* L;
  int M, k_cta0, g_first, x_stride;
  const char *sc, *zq, *xs;
};

// ------------------------------------------------------------------------------------------------
// GPTQ / HQQ, BITS in {2,4,8}: CTA tile 32 columns; step = 4 packed rows (lane t -> row t), lane g
// loads the words of columns 4g..4g+3 with one LDG.128.
// ------------------------------------------------------------------------------------------------
template <int BITS, bool FLOATZ>
struct RpGptq {
  static constexpr int P = 32 / BITS;
  static constexpr int NT = 32, KSTEP = 4 * P, MAXSTEPS = 16, NSETS = 2, N_GRAN = 32;
  static constexpr int ROWS_PER_STEP = 4, ROW_WORDS = 32, RS_WORDS = 40, SM_MIN_BLOCKS = 3;   // smem-staged variant geometry
  static constexpr int SC_ROW_BYTES = NT * 2;
  static constexpr int ZQ_ROW_BYTES = FLOATZ ? NT * 2 : NT * BITS / 8;
  static constexpr int NC = (BITS == 2) ? 4 : (BITS == 4 ? 2 : 1);
  static constexpr bool kHasZq = true;
  using Step = uint4;
  struct Consts {
    int gcur;
    uint32_t s2[4], c[4][NC], z2[4];
  };

  __device__ static void load(Step& w, const LayerView& L, int s, int n0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    w = ldg128_stream(L.qw + (size_t)(s * 4 + t) * L.N + n0 + 4 * g);
  }
  __device__ static size_t src_word(const LayerView& L, int row, int n0) { return (size_t)row * L.N + n0; }
  __device__ static void load_smem(Step& w, const uint32_t* tile, int ls, int lane) {
    const int g = lane >> 2, t = lane & 3;
    w = r_lds128(tile + (size_t)(ls * 4 + t) * RS_WORDS + 4 * g);
  }

  __device__ static void reload(Consts& c, const RpCtx& cx, int gl, int ncol0) {
    const uint2 sv = r_lds64(cx.sc + (size_t)gl * SC_ROW_BYTES + ncol0 * 2);
    c.s2[0] = prmt(sv.x, sv.x, 0x1010); c.s2[1] = prmt(sv.x, sv.x, 0x3232);
    c.s2[2] = prmt(sv.y, sv.y, 0x1010); c.s2[3] = prmt(sv.y, sv.y, 0x3232);
    if (FLOATZ) {
      const uint2 zv = r_lds64(cx.zq + (size_t)gl * ZQ_ROW_BYTES + ncol0 * 2);
      c.z2[0] = prmt(zv.x, zv.x, 0x1010); c.z2[1] = prmt(zv.x, zv.x, 0x3232);
      c.z2[2] = prmt(zv.y, zv.y, 0x1010); c.z2[3] = prmt(zv.y, zv.y, 0x3232);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        c.c[i][0] = MAGIC;
        if (BITS == 4) c.c[i][1] = 0xD400D400u;
        if (BITS == 2) { c.c[i][1] = 0xDC00DC00u; c.c[i][2] = 0xD400D400u; c.c[i][3] = 0xCC00CC00u; }
      }
    } else {
      const int bitpos = ncol0 * BITS;
      const uint32_t word = r_lds32(cx.zq + (size_t)gl * ZQ_ROW_BYTES + (bitpos >> 5) * 4);
      const uint32_t zs = word >> (bitpos & 31);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t z = (((zs >> (BITS * i)) & ((1u << BITS) - 1u)) + (uint32_t)cx.L->zero_bias) & ((1u << BITS) - 1u);
        c.c[i][0] = (0x6400u | z) * 0x00010001u;
        if (BITS == 4) c.c[i][1] = (0xD400u + (z << 4)) * 0x00010001u;
        if (BITS == 2) {
          c.c[i][1] = (0xDC00u + (z << 2)) * 0x00010001u;
          c.c[i][2] = (0xD400u + (z << 4)) * 0x00010001u;
          c.c[i][3] = (0xCC00u + (z << 6)) * 0x00010001u;
        }
      }
    }
  }

  template <int TYPE>
  __device__ static uint32_t finish(uint32_t h, const Consts& c, int i) {
    uint32_t d;
    if (BITS == 8) d = hsub2_u(h, c.c[i][0]);
    else if (BITS == 4) d = (TYPE == 0) ? hsub2_u(h, c.c[i][0]) : hfma2_u(h, H_1_16, c.c[i][1]);
    else d = (TYPE == 0) ? hsub2_u(h, c.c[i][0])
           : (TYPE == 1) ? hfma2_u(h, H_1_4, c.c[i][1])
           : (TYPE == 2) ? hfma2_u(h, H_1_16, c.c[i][2]) : hfma2_u(h, H_1_64, c.c[i][3]);
    if (FLOATZ) d = hsub2_u(d, c.z2[i]);
    return hmul2_u(d, c.s2[i]);
  }

  __device__ static void unpack_word(uint32_t w, const Consts& c, int i, uint32_t (&o)[P / 2]) {
    if (BITS == 4) {
      const uint32_t hi = w >> 8;
      o[0] = finish<0>(and_or(w, LO4, MAGIC), c, i);
      o[1] = finish<1>(and_or(w, HI4, MAGIC), c, i);
      o[2] = finish<0>(and_or(hi, LO4, MAGIC), c, i);
      o[3] = finish<1>(and_or(hi, HI4, MAGIC), c, i);
    } else if (BITS == 8) {
      o[0] = finish<0>(prmt(w, MAGIC, 0x5150), c, i);
      o[1] = finish<0>(prmt(w, MAGIC, 0x5352), c, i);
    } else {
      const uint32_t hi = w >> 8;
      o[0] = finish<0>(and_or(w, 0x00030003u, MAGIC), c, i);
      o[1] = finish<1>(and_or(w, 0x000C000Cu, MAGIC), c, i);
      o[2] = finish<2>(and_or(w, 0x00300030u, MAGIC), c, i);
      o[3] = finish<3>(and_or(w, 0x00C000C0u, MAGIC), c, i);
      o[4] = finish<0>(and_or(hi, 0x00030003u, MAGIC), c, i);
      o[5] = finish<1>(and_or(hi, 0x000C000Cu, MAGIC), c, i);
      o[6] = finish<2>(and_or(hi, 0x00300030u, MAGIC), c, i);
      o[7] = finish<3>(and_or(hi, 0x00C000C0u, MAGIC), c, i);
    }
  }

  __device__ static void load_x(const char* xrow, int koff, bool active, uint32_t (&xb)[P / 2]) {
#pragma unroll
    for (int j = 0; j < P / 2; ++j) xb[j] = 0u;
    if (!active) return;
    if (BITS == 8) {
      const uint2 v = r_lds64(xrow + koff * 2);
      xb[0] = v.x; xb[1] = v.y;
    } else if (BITS == 4) {
      const uint4 v = r_lds128(xrow + koff * 2);
      xb[0] = prmt(v.x, v.z, 0x5410); xb[1] = prmt(v.x, v.z, 0x7632);
      xb[2] = prmt(v.y, v.w, 0x5410); xb[3] = prmt(v.y, v.w, 0x7632);
    } else {
      const uint4 a = r_lds128(xrow + koff * 2), b = r_lds128(xrow + koff * 2 + 16);
      xb[0] = prmt(a.x, b.x, 0x5410); xb[1] = prmt(a.x, b.x, 0x7632);
      xb[2] = prmt(a.y, b.y, 0x5410); xb[3] = prmt(a.y, b.y, 0x7632);
      xb[4] = prmt(a.z, b.z, 0x5410); xb[5] = prmt(a.z, b.z, 0x7632);
      xb[6] = prmt(a.w, b.w, 0x5410); xb[7] = prmt(a.w, b.w, 0x7632);
    }
  }

  __device__ static void compute(const Step& w, int s, const RpCtx& cx, Consts& c, float (&acc)[NSETS][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int kthr = s * KSTEP + P * t;
    const int gi = kthr / cx.L->group;
    if (gi != c.gcur) { c.gcur = gi; reload(c, cx, gi - cx.g_first, 4 * g); }
    uint32_t xb[P / 2];
    load_x(cx.xs + (size_t)g * cx.x_stride, kthr - cx.k_cta0, g < cx.M, xb);
    uint32_t a[4][P / 2];
    unpack_word(w.x, c, 0, a[0]);
    unpack_word(w.y, c, 1, a[1]);
    unpack_word(w.z, c, 2, a[2]);
    unpack_word(w.w, c, 3, a[3]);
#pragma unroll
    for (int m = 0; m < P / 4; ++m) {
      mma_16816(acc[0], a[0][2 * m], a[1][2 * m], a[0][2 * m + 1], a[1][2 * m + 1], xb[2 * m], xb[2 * m + 1]);
      mma_16816(acc[1], a[2][2 * m], a[3][2 * m], a[2][2 * m + 1], a[3][2 * m + 1], xb[2 * m], xb[2 * m + 1]);
    }
  }

  __device__ static void store_acc(float* r, int mstride, const float (&acc)[NSETS][4], int lane, int M) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int s = 0; s < NSETS; ++s) {
      const int n = 4 * g + 2 * s;
      if (2 * t < M) { r[n * mstride + 2 * t] = acc[s][0]; r[(n + 1) * mstride + 2 * t] = acc[s][2]; }
      if (2 * t + 1 < M) { r[n * mstride + 2 * t + 1] = acc[s][1]; r[(n + 1) * mstride + 2 * t + 1] = acc[s][3]; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// AWQ GEMM layout: CTA tile 128 columns (16 words = 64 B per k row); step = 16 k rows; lane (g,t)
// loads word columns 2g,2g+1 (LDG.64) at rows 2t, 2t+1, 8+2t, 9+2t.
// ------------------------------------------------------------------------------------------------
struct RpAwq {
  static constexpr int NT = 128, KSTEP = 16, MAXSTEPS = 4, NSETS = 8, N_GRAN = 128;
  static constexpr int ROWS_PER_STEP = 16, ROW_WORDS = 16, RS_WORDS = 20, SM_MIN_BLOCKS = 2;
  static constexpr int SC_ROW_BYTES = NT * 2, ZQ_ROW_BYTES = NT / 2;
  static constexpr bool kHasZq = true;
  struct Step { uint2 r[4]; };
  struct Consts {
    int gcur;
    uint32_t s2[16], c[16];
  };

  __device__ static void load(Step& w, const LayerView& L, int s, int n0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t* base = L.qw + (size_t)(s * 16 + 2 * t) * (L.N >> 3) + (n0 >> 3) + 2 * g;
    const size_t rs = (size_t)(L.N >> 3);
    w.r[0] = ldg64_stream(base);
    w.r[1] = ldg64_stream(base + rs);
    w.r[2] = ldg64_stream(base + 8 * rs);
    w.r[3] = ldg64_stream(base + 9 * rs);
  }
  __device__ static size_t src_word(const LayerView& L, int row, int n0) { return (size_t)row * (L.N >> 3) + (n0 >> 3); }
  __device__ static void load_smem(Step& w, const uint32_t* tile, int ls, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t* base = tile + (size_t)(ls * 16 + 2 * t) * RS_WORDS + 2 * g;
    w.r[0] = r_lds64(base);
    w.r[1] = r_lds64(base + RS_WORDS);
    w.r[2] = r_lds64(base + 8 * RS_WORDS);
    w.r[3] = r_lds64(base + 9 * RS_WORDS);
  }

  __device__ static void reload(Consts& c, const RpCtx& cx, int gl, int g) {
    const uint4 s0 = r_lds128(cx.sc + (size_t)gl * SC_ROW_BYTES + g * 32);
    const uint4 s1 = r_lds128(cx.sc + (size_t)gl * SC_ROW_BYTES + g * 32 + 16);
    const uint32_t sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) { c.s2[2 * i] = prmt(sv[i], sv[i], 0x1010); c.s2[2 * i + 1] = prmt(sv[i], sv[i], 0x3232); }
    const uint2 zw = r_lds64(cx.zq + (size_t)gl * ZQ_ROW_BYTES + g * 8);
#pragma unroll
    for (int wc = 0; wc < 2; ++wc) {
      const uint32_t zz = wc ? zw.y : zw.x;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int nib = (n >> 1) + ((n & 1) << 2);
        const int j = n >> 1;
        const uint32_t z = (((zz >> (4 * nib)) & 0xFu) + (uint32_t)cx.L->zero_bias) & 0xFu;
        c.c[wc * 8 + n] = (j & 1) ? (0xD400u + (z << 4)) * 0x00010001u : (0x6400u | z) * 0x00010001u;
      }
    }
  }

  __device__ static void compute(const Step& w, int s, const RpCtx& cx, Consts& c, float (&acc)[NSETS][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int k0 = s * KSTEP;
    const int gi = k0 / cx.L->group;
    if (gi != c.gcur) { c.gcur = gi; reload(c, cx, gi - cx.g_first, g); }
    uint32_t b0 = 0u, b1 = 0u;
    if (g < cx.M) {
      const char* xr = cx.xs + (size_t)g * cx.x_stride + (size_t)(k0 - cx.k_cta0 + 2 * t) * 2;
      b0 = r_lds32(xr);
      b1 = r_lds32(xr + 16);
    }
#pragma unroll
    for (int wc = 0; wc < 2; ++wc) {
      const uint32_t wa = wc ? w.r[0].y : w.r[0].x, wb = wc ? w.r[1].y : w.r[1].x;
      const uint32_t wcw = wc ? w.r[2].y : w.r[2].x, wd = wc ? w.r[3].y : w.r[3].x;
      const uint32_t u01 = prmt(wa, wb, 0x5410), v01 = prmt(wa, wb, 0x7632);
      const uint32_t u23 = prmt(wcw, wd, 0x5410), v23 = prmt(wcw, wd, 0x7632);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int sh = (j >> 1) * 8;
        const uint32_t msk = (j & 1) ? HI4 : LO4;
        uint32_t a0 = and_or(u01 >> sh, msk, MAGIC), a1 = and_or(v01 >> sh, msk, MAGIC);
        uint32_t a2 = and_or(u23 >> sh, msk, MAGIC), a3 = and_or(v23 >> sh, msk, MAGIC);
        const int na = wc * 8 + 2 * j, nb = na + 1;
        if (j & 1) {
          a0 = hfma2_u(a0, H_1_16, c.c[na]); a1 = hfma2_u(a1, H_1_16, c.c[nb]);
          a2 = hfma2_u(a2, H_1_16, c.c[na]); a3 = hfma2_u(a3, H_1_16, c.c[nb]);
        } else {
          a0 = hsub2_u(a0, c.c[na]); a1 = hsub2_u(a1, c.c[nb]);
          a2 = hsub2_u(a2, c.c[na]); a3 = hsub2_u(a3, c.c[nb]);
        }
        a0 = hmul2_u(a0, c.s2[na]); a1 = hmul2_u(a1, c.s2[nb]);
        a2 = hmul2_u(a2, c.s2[na]); a3 = hmul2_u(a3, c.s2[nb]);
        mma_16816(acc[wc * 4 + j], a0, a1, a2, a3, b0, b1);
      }
    }
  }

  __device__ static void store_acc(float* r, int mstride, const float (&acc)[NSETS][4], int lane, int M) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int q = 0; q < NSETS; ++q) {
      const int n = 16 * g + (q >> 2) * 8 + 2 * (q & 3);
      if (2 * t < M) { r[n * mstride + 2 * t] = acc[q][0]; r[(n + 1) * mstride + 2 * t] = acc[q][2]; }
      if (2 * t + 1 < M) { r[n * mstride + 2 * t + 1] = acc[q][1]; r[(n + 1) * mstride + 2 * t + 1] = acc[q][3]; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Marlin layout: CTA tile 64 columns (one 128-word block per k16 row); step = one k16 row; each
// lane's LDG.128 is its A fragment for 4 MMAs (16-column tiles j=0..3).
// ------------------------------------------------------------------------------------------------
struct RpMarlin {
  static constexpr int NT = 64, KSTEP = 16, MAXSTEPS = 16, NSETS = 4, N_GRAN = 64;
  static constexpr int ROWS_PER_STEP = 1, ROW_WORDS = 128, RS_WORDS = 128, SM_MIN_BLOCKS = 3;
  static constexpr int SC_ROW_BYTES = NT * 2, ZQ_ROW_BYTES = 0;
  static constexpr bool kHasZq = false;
  using Step = uint4;
  struct Consts {
    int gcur;
    uint32_t s2[8];
  };

  __device__ static void load(Step& w, const LayerView& L, int s, int n0, int lane) {
    w = ldg128_stream(L.qw + (size_t)s * (2 * (size_t)L.N) + 2 * n0 + 4 * lane);
  }
  __device__ static size_t src_word(const LayerView& L, int row, int n0) { return (size_t)row * (2 * (size_t)L.N) + 2 * n0; }
  __device__ static void load_smem(Step& w, const uint32_t* tile, int ls, int lane) {
    w = r_lds128(tile + (size_t)ls * RS_WORDS + 4 * lane);
  }

  __device__ static void reload(Consts& c, const RpCtx& cx, int gl, int g) {
    if (cx.L->group != cx.L->K) {
      const uint4 sv = r_lds128(cx.sc + (size_t)gl * SC_ROW_BYTES + (8 * g) * 2);
      c.s2[0] = prmt(sv.x, sv.x, 0x1010); c.s2[1] = prmt(sv.x, sv.x, 0x3232);
      c.s2[2] = prmt(sv.y, sv.y, 0x1010); c.s2[3] = prmt(sv.y, sv.y, 0x3232);
      c.s2[4] = prmt(sv.z, sv.z, 0x1010); c.s2[5] = prmt(sv.z, sv.z, 0x3232);
      c.s2[6] = prmt(sv.w, sv.w, 0x1010); c.s2[7] = prmt(sv.w, sv.w, 0x3232);
    } else {
      const __half* row = reinterpret_cast<const __half*>(cx.sc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c.s2[2 * j] = dup_half(row[marlin_scale_index(16 * j + g, true)]);
        c.s2[2 * j + 1] = dup_half(row[marlin_scale_index(16 * j + g + 8, true)]);
      }
    }
  }

  __device__ static void compute(const Step& w, int s, const RpCtx& cx, Consts& c, float (&acc)[NSETS][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int k0 = s * KSTEP;
    const int gi = k0 / cx.L->group;
    if (gi != c.gcur) { c.gcur = gi; reload(c, cx, gi - cx.g_first, g); }
    uint32_t b0 = 0u, b1 = 0u;
    if (g < cx.M) {
      const char* xr = cx.xs + (size_t)g * cx.x_stride + (size_t)(k0 - cx.k_cta0 + 2 * t) * 2;
      b0 = r_lds32(xr);
      b1 = r_lds32(xr + 16);
    }
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t lo = ww[j], hi = ww[j] >> 8;
      uint32_t a0 = hsub2_u(and_or(lo, LO4, MAGIC), 0x64086408u);
      uint32_t a2 = hfma2_u(and_or(lo, HI4, MAGIC), H_1_16, 0xD480D480u);
      uint32_t a1 = hsub2_u(and_or(hi, LO4, MAGIC), 0x64086408u);
      uint32_t a3 = hfma2_u(and_or(hi, HI4, MAGIC), H_1_16, 0xD480D480u);
      a0 = hmul2_u(a0, c.s2[2 * j]); a2 = hmul2_u(a2, c.s2[2 * j]);
      a1 = hmul2_u(a1, c.s2[2 * j + 1]); a3 = hmul2_u(a3, c.s2[2 * j + 1]);
      mma_16816(acc[j], a0, a1, a2, a3, b0, b1);
    }
  }

  __device__ static void store_acc(float* r, int mstride, const float (&acc)[NSETS][4], int lane, int M) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = 16 * j + g;
      if (2 * t < M) { r[n * mstride + 2 * t] = acc[j][0]; r[(n + 8) * mstride + 2 * t] = acc[j][2]; }
      if (2 * t + 1 < M) { r[n * mstride + 2 * t + 1] = acc[j][1]; r[(n + 8) * mstride + 2 * t + 1] = acc[j][3]; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// SM == false: weights prefetched into registers (v2).  SM == true: the CTA's whole packed slice is
// prefetched into shared memory with cp.async (v3) -- no register cost, so more CTAs (and therefore
// more layers' worth of weights) are in flight per SM.
template <class T, bool SM>
__global__ void __launch_bounds__(kRpThreads, SM ? T::SM_MIN_BLOCKS : 2) gemv_rp_kernel(const RpParams p) {
  extern __shared__ __align__(128) char smem[];
  char* xs = smem + p.off_x;
  char* sc = smem + p.off_sc;
  char* zq = smem + p.off_zq;
  float* red = reinterpret_cast<float*>(smem + p.off_red);
  float* rbuf = reinterpret_cast<float*>(smem + p.off_rbuf);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = p.cluster;
  const int rank = (int)cluster_ctarank();
  const int n_tile = blockIdx.x / cs;
  const int n0 = n_tile * T::NT;
  const int U = cs * kWarps;
  const int S = p.steps_total;
  const int unit = rank * kWarps + warp;
  const int s_begin = (int)(((long long)unit * S) / U), s_end = (int)(((long long)(unit + 1) * S) / U);
  const int cta_s0 = (int)(((long long)rank * kWarps * S) / U), cta_s1 = (int)(((long long)(rank + 1) * kWarps * S) / U);
  const int k_cta0 = cta_s0 * T::KSTEP, k_cta1 = cta_s1 * T::KSTEP;
  const int kslice = k_cta1 - k_cta0;
  const int g_first = k_cta0 / p.L.group;
  const int g_count = kslice > 0 ? (k_cta1 - 1) / p.L.group - g_first + 1 : 0;

  pdl_launch_dependents();

  // ---- 1. weights -> registers / shared (independent of the upstream kernel) ----
  typename T::Step w[SM ? 1 : T::MAXSTEPS];
  const uint32_t* wtile = reinterpret_cast<const uint32_t*>(smem + p.off_w);
  if (SM) {
    constexpr int CPR = T::ROW_WORDS / 4;                       // 16-byte chunks per packed row
    const int row0 = cta_s0 * T::ROWS_PER_STEP, nrows = (cta_s1 - cta_s0) * T::ROWS_PER_STEP;
    char* wt = smem + p.off_w;
    for (int idx = tid; idx < nrows * CPR; idx += kRpThreads) {
      const int r = idx / CPR, cc = idx % CPR;
      cp_async16(wt + ((size_t)r * T::RS_WORDS + 4 * cc) * 4, p.L.qw + T::src_word(p.L, row0 + r, n0) + 4 * cc);
    }
    cp_async_commit();
  } else {
#pragma unroll
    for (int i = 0; i < (SM ? 1 : T::MAXSTEPS); ++i)
      if (s_begin + i < s_end) T::load(w[i], p.L, s_begin + i, n0, lane);
  }

  // group constants of this CTA's k-range -> shared
  for (int idx = tid; idx < g_count * T::NT; idx += kRpThreads) {
    const int gl = idx / T::NT, n = idx % T::NT;
    reinterpret_cast<__half*>(sc + (size_t)gl * T::SC_ROW_BYTES)[n] = __ldg(p.L.s + (size_t)(g_first + gl) * p.L.N + n0 + n);
  }
  if (T::kHasZq) {
    if (p.L.layout == B200Q_LAYOUT_HQQ) {
      for (int idx = tid; idx < g_count * T::NT; idx += kRpThreads) {
        const int gl = idx / T::NT, n = idx % T::NT;
        reinterpret_cast<__half*>(zq + (size_t)gl * T::ZQ_ROW_BYTES)[n] =
            __ldg(reinterpret_cast<const __half*>(p.L.qz) + (size_t)(g_first + gl) * p.L.N + n0 + n);
      }
    } else {
      const int zwords = (T::NT * p.L.bits) >> 5;
      const size_t zrow = ((size_t)p.L.N * p.L.bits) >> 5;
      const size_t zoff = ((size_t)n0 * p.L.bits) >> 5;
      for (int idx = tid; idx < g_count * zwords; idx += kRpThreads) {
        const int gl = idx / zwords, wv = idx % zwords;
        reinterpret_cast<uint32_t*>(zq + (size_t)gl * T::ZQ_ROW_BYTES)[wv] =
            __ldg(reinterpret_cast<const uint32_t*>(p.L.qz) + (size_t)(g_first + gl) * zrow + zoff + wv);
      }
    }
  }

  // ---- 2. activations (produced by the upstream kernel) ----
  pdl_wait();
  {
    const int vec_per_row = kslice >> 3;
    for (int idx = tid; idx < p.M * vec_per_row; idx += kRpThreads) {
      const int m = idx / vec_per_row, v = idx % vec_per_row;
      const uint4 val = *reinterpret_cast<const uint4*>(p.x + (size_t)m * p.ldx + k_cta0 + 8 * v);
      *reinterpret_cast<uint4*>(xs + (size_t)m * p.x_stride + 16 * v) = val;
    }
  }
  if (SM) cp_async_wait_all();
  __syncthreads();

  RpCtx cx;
  cx.L = &p.L; cx.M = p.M; cx.k_cta0 = k_cta0; cx.g_first = g_first; cx.x_stride = p.x_stride;
  cx.sc = sc; cx.zq = zq; cx.xs = xs;
  float acc[T::NSETS][4];
#pragma unroll
  for (int s = 0; s < T::NSETS; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[s][i] = 0.f;
  typename T::Consts c;
  c.gcur = -1;

  if (SM) {
#pragma unroll 2
    for (int s = s_begin; s < s_end; ++s) {
      T::load_smem(w[0], wtile, s - cta_s0, lane);
      T::compute(w[0], s, cx, c, acc, lane);
    }
  } else {
#pragma unroll
    for (int i = 0; i < (SM ? 1 : T::MAXSTEPS); ++i)
      if (s_begin + i < s_end) T::compute(w[i], s_begin + i, cx, c, acc, lane);
    // further rounds (only when one warp owns more than MAXSTEPS steps)
    for (int sb = s_begin + T::MAXSTEPS; sb < s_end; sb += T::MAXSTEPS) {
#pragma unroll
      for (int i = 0; i < (SM ? 1 : T::MAXSTEPS); ++i)
        if (sb + i < s_end) T::load(w[i], p.L, sb + i, n0, lane);
#pragma unroll
      for (int i = 0; i < (SM ? 1 : T::MAXSTEPS); ++i)
        if (sb + i < s_end) T::compute(w[i], sb + i, cx, c, acc, lane);
    }
  }

  // ---- 3. reduce: warps -> CTA (shared), CTAs of the cluster -> rank 0 (distributed shared) ----
  const int ms = p.M;
  T::store_acc(red + (size_t)warp * T::NT * ms, ms, acc, lane, p.M);
  __syncthreads();
  const int total = T::NT * p.M;          // idx = n * M + m
  float v[(T::NT * kMB + kRpThreads - 1) / kRpThreads];
#pragma unroll
  for (int r = 0; r < (T::NT * kMB + kRpThreads - 1) / kRpThreads; ++r) {
    const int idx = tid + r * kRpThreads;
    float sum = 0.f;
    if (idx < total) {
#pragma unroll
      for (int wq = 0; wq < kWarps; ++wq) sum += red[(size_t)wq * T::NT * ms + idx];
    }
    v[r] = sum;
  }
  if (cs > 1) {
    if (rank != 0) {
#pragma unroll
      for (int r = 0; r < (T::NT * kMB + kRpThreads - 1) / kRpThreads; ++r) {
        const int idx = tid + r * kRpThreads;
        if (idx < total) st_cluster_f32(rbuf + (size_t)(rank - 1) * total + idx, 0u, v[r]);
      }
    }
    cluster_sync_all();
    if (rank != 0) return;
#pragma unroll
    for (int r = 0; r < (T::NT * kMB + kRpThreads - 1) / kRpThreads; ++r) {
      const int idx = tid + r * kRpThreads;
      if (idx < total)
        for (int q = 0; q < cs - 1; ++q) v[r] += rbuf[(size_t)q * total + idx];
    }
  }
#pragma unroll
  for (int r = 0; r < (T::NT * kMB + kRpThreads - 1) / kRpThreads; ++r) {
    const int idx = tid + r * kRpThreads;
    if (idx < total) {
      const int n = idx / p.M, m = idx % p.M;
      float o = v[r];
      if (p.L.bias) o += __half2float(__ldg(p.L.bias + n0 + n));
      const __half h = __float2half_rn(o);
      for (int q = 0; q < p.out.n; ++q) p.out.y[q][(size_t)m * p.ldy + p.n_offset + n0 + n] = h;
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct RpPlan {
  int kind, NT, KSTEP, MAXSTEPS, SC_ROW, ZQ_ROW, n_gran, RPS, RSW;
  bool sm;
  int off_w;
  int n_tiles, cluster, steps_total, x_stride;
  int off_x, off_sc, off_zq, off_red, off_rbuf, smem_bytes;
};

template <class T>
static void rp_fill(RpPlan& pl) {
  pl.NT = T::NT; pl.KSTEP = T::KSTEP; pl.MAXSTEPS = T::MAXSTEPS; pl.SC_ROW = T::SC_ROW_BYTES; pl.ZQ_ROW = T::ZQ_ROW_BYTES;
  pl.n_gran = T::N_GRAN; pl.RPS = T::ROWS_PER_STEP; pl.RSW = T::RS_WORDS;
}

static int g_rp_max_cluster = 8;
static bool g_rp_smem = false;
static int g_rp_slice_kb = 40;

static bool rp_plan(const LayerView& L, int M, RpPlan& pl) {
  pl.kind = 0;
  pl.sm = g_rp_smem;
  if (M < 1 || M > kMB || L.g_idx != nullptr) return false;
  const bool fz = (L.layout == B200Q_LAYOUT_HQQ);
  if (L.layout == B200Q_LAYOUT_GPTQ || L.layout == B200Q_LAYOUT_HQQ) {
    if (L.bits == 2) { pl.kind = 1; rp_fill<RpGptq<2, false>>(pl); }
    else if (L.bits == 4) { pl.kind = 2; rp_fill<RpGptq<4, false>>(pl); }
    else if (L.bits == 8) { pl.kind = 3; rp_fill<RpGptq<8, false>>(pl); }
    else return false;
    if (fz) { pl.kind += 16; pl.ZQ_ROW = pl.NT * 2; }
    if (L.group % (32 / L.bits) != 0) return false;
  } else if (L.layout == B200Q_LAYOUT_AWQ_GEMM) {
    if (L.bits != 4 || L.group % 16 != 0) return false;
    pl.kind = 4; rp_fill<RpAwq>(pl);
  } else if (L.layout == B200Q_LAYOUT_MARLIN) {
    if (L.bits != 4 || L.group % 16 != 0) return false;
    pl.kind = 5; rp_fill<RpMarlin>(pl);
  } else return false;
  if (L.K % pl.KSTEP != 0 || L.N % pl.n_gran != 0 || L.K % L.group != 0) { pl.kind = 0; return false; }
  pl.n_tiles = L.N / pl.NT;
  pl.steps_total = L.K / pl.KSTEP;
  // cluster size: enough k-split that one round of register prefetch covers a warp's steps
  int cs = 1;
  if (pl.sm) {
    const long long bytes = (long long)pl.steps_total * pl.RPS * pl.RSW * 4;
    while (cs < g_rp_max_cluster && bytes / cs > (long long)g_rp_slice_kb * 1024) cs *= 2;
  } else {
    while (cs < g_rp_max_cluster && (pl.steps_total + cs * kWarps - 1) / (cs * kWarps) > pl.MAXSTEPS) cs *= 2;
  }
  // and enough CTAs to cover the machine
  while (cs < g_rp_max_cluster && pl.n_tiles * cs < 148 && pl.steps_total / (2 * cs * kWarps) >= 2) cs *= 2;
  pl.cluster = cs;
  const int kslice = ((pl.steps_total + cs - 1) / cs + kWarps) * pl.KSTEP;   // upper bound of a CTA's k-range
  pl.x_stride = kslice * 2;
  pl.x_stride += (64 - (pl.x_stride % 128) + 128) % 128;
  const int gcap = kslice / L.group + 2;
  int off = 0;
  pl.off_x = off; off += M * pl.x_stride;
  off = (off + 15) & ~15;
  pl.off_sc = off; off += gcap * pl.SC_ROW;
  off = (off + 15) & ~15;
  pl.off_zq = off; off += gcap * pl.ZQ_ROW;
  off = (off + 15) & ~15;
  pl.off_red = off; off += kWarps * pl.NT * M * 4;
  pl.off_rbuf = off; off += (cs - 1) * pl.NT * M * 4;
  off = (off + 127) & ~127;
  pl.off_w = off;
  if (pl.sm) off += ((pl.steps_total + cs - 1) / cs + kWarps) * pl.RPS * pl.RSW * 4;
  pl.smem_bytes = off;
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

void gemv_rp_set_max_cluster(int c) { g_rp_max_cluster = c < 1 ? 1 : (c > 8 ? 8 : c); }
void gemv_rp_set_smem(bool on, int slice_kb) { g_rp_smem = on; if (slice_kb > 0) g_rp_slice_kb = slice_kb; }

template <class T, bool SM>
static cudaError_t rp_launch_s(const RpParams& p, const RpPlan& pl, cudaStream_t st);

template <class T>
static cudaError_t rp_launch_t(const RpParams& p, const RpPlan& pl, cudaStream_t st) {
  return pl.sm ? rp_launch_s<T, true>(p, pl, st) : rp_launch_s<T, false>(p, pl, st);
}

template <class T, bool SM>
static cudaError_t rp_launch_s(const RpParams& p, const RpPlan& pl, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemv_rp_kernel<T, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
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
  return cudaLaunchKernelEx(&cfg, gemv_rp_kernel<T, SM>, p);
}

cudaError_t launch_gemv_rp(const LinearArgs& a, const PeerOut* peers) {
  RpPlan pl;
  if (!rp_plan(a.L, a.M, pl)) return cudaErrorInvalidValue;
  RpParams p;
  p.L = a.L; p.x = a.x; p.ldx = a.ldx; p.M = a.M;
  if (peers) p.out = *peers; else { p.out.n = 1; p.out.y[0] = a.y; }
  p.ldy = a.ldy; p.n_offset = a.n_offset;
  p.n_tiles = pl.n_tiles; p.cluster = pl.cluster; p.steps_total = pl.steps_total; p.x_stride = pl.x_stride;
  p.off_x = pl.off_x; p.off_sc = pl.off_sc; p.off_zq = pl.off_zq; p.off_red = pl.off_red; p.off_rbuf = pl.off_rbuf; p.off_w = pl.off_w;
  switch (pl.kind) {
    case 1: return rp_launch_t<RpGptq<2, false>>(p, pl, a.stream);
    case 2: return rp_launch_t<RpGptq<4, false>>(p, pl, a.stream);
    case 3: return rp_launch_t<RpGptq<8, false>>(p, pl, a.stream);
    case 17: return rp_launch_t<RpGptq<2, true>>(p, pl, a.stream);
    case 18: return rp_launch_t<RpGptq<4, true>>(p, pl, a.stream);
    case 19: return rp_launch_t<RpGptq<8, true>>(p, pl, a.stream);
    case 4: return rp_launch_t<RpAwq>(p, pl, a.stream);
    case 5: return rp_launch_t<RpMarlin>(p, pl, a.stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace b200q
