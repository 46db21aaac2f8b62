// Decode (M <= 8) kernel family: HBM-bound streaming of the packed weights exactly as they sit in
// the checkpoint, no repacking.
//
//   producer warp : cp.async.bulk (TMA engine, 1-D form) global -> shared, one copy per packed row
//                   segment, S-stage mbarrier ring; weights/scales do not depend on the previous
//                   kernel so they are issued before griddepcontrol.wait (PDL), x after it.
//   8 consumer warps: read packed words from shared memory, unpack in registers with lop3/prmt
//                   into fp16 pairs, apply (q - z) * s, and feed mma.sync.m16n8k16 with the
//                   weights as the 16-row A operand and the <=8 activation rows as the n8 B operand
//                   (fp32 accumulation).  The k-slot order inside an MMA is permuted to whatever
//                   order the unpack produces; x is permuted to match, so no data shuffles.
//   reduction      : k-split warps reduce through shared memory; k-split CTAs through an fp32
//                   scratch + arrival counter, summed in fixed order by the last CTA (deterministic),
//                   which also adds bias, converts to fp16 and stores to every peer output.
//
// Replaces (hot cases of): ort_ops.gemv (dq_gemv.cu:40-177), gemm_forward_cuda at M<=8
// (gemm_cuda_gen.cu:31-353), Marlin at M<=8 (marlin_cuda_kernel.cu:222-733), torch HQQ path.
// Dequant arithmetic: fp16((q - z) * s), identical to gemm_cuda_gen.cu:153-176.
#include "common.cuh"
#include "kernels.h"

namespace b200q {

static constexpr int kConsumerWarps = 8;
static constexpr int kThreads = (kConsumerWarps + 1) * 32;
static constexpr int kStages = 4;
static constexpr int kMB = 8;  // activation rows per MMA (n8)

static constexpr uint32_t MAGIC = 0x64006400u;   // half2(1024, 1024)
static constexpr uint32_t LO4 = 0x000f000fu, HI4 = 0x00f000f0u;
static constexpr uint32_t H_1_4 = 0x34003400u, H_1_16 = 0x2c002c00u, H_1_64 = 0x24002400u;

struct GemvParams {
  LayerView L;
  const __half* x;
  int64_t ldx;
  int M;
  PeerOut out;
  int64_t ldy, n_offset;
  int n_tiles, ksplit, steps_total, steps_per_cta;
  unsigned* counters;
  float* partial;
  int off_stage, off_x, x_stride, off_sc, off_zq, off_red, off_flag;  // byte offsets in dynamic smem
  int gcap;                                                           // groups the const slices can hold
};

__device__ __forceinline__ uint32_t lds32(const void* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint2 lds64(const void* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }

// ------------------------------------------------------------------------------------------------
// GPTQ / HQQ layout, BITS in {2,4,8}.  Tile: 64 columns x 32 packed rows per stage.
// warp = (nsub = w & 1 -> 32 columns, ksub = w >> 1); a step is 4 packed rows (thread t takes row t),
// each thread LDS.128s 4 adjacent columns' words -> 2 "sets" of 2 columns (MMA rows g and g+8).
// ------------------------------------------------------------------------------------------------
template <int BITS, bool FLOATZ>
struct GptqTraits {
  static constexpr int P = 32 / BITS;               // k values per word
  static constexpr int NT = 64, NW = 2, KW = 4;
  static constexpr int ROWS_PER_STEP = 4, KSTEP = 4 * P, SPS = 8;
  static constexpr int ROW_WORDS = 64, RS_WORDS = 72;
  static constexpr int NSETS = 2;
  static constexpr int N_GRAN = 32;                  // valid tile width granularity
  static constexpr int SC_ROW_BYTES = NT * 2;
  static constexpr int ZQ_ROW_BYTES = FLOATZ ? NT * 2 : NT * BITS / 8;
  static constexpr int NC = (BITS == 2) ? 4 : (BITS == 4 ? 2 : 1);   // zero constants per column

  struct Consts {
    int gcur;
    uint32_t s2[4];
    uint32_t c[4][NC];
    uint32_t z2[4];      // FLOATZ only
  };

  __device__ static size_t tile_src_word(const LayerView& L, int row, int n0) { return (size_t)row * L.N + n0; }
  __device__ static int valid_row_bytes(int ncols) { return ncols * 4; }

  __device__ static void reload(Consts& c, const LayerView& L, const char* sc, const char* zq, int gl, int ncol0) {
    const uint2 sv = lds64(sc + (size_t)gl * SC_ROW_BYTES + ncol0 * 2);
    c.s2[0] = prmt(sv.x, sv.x, 0x1010); c.s2[1] = prmt(sv.x, sv.x, 0x3232);
    c.s2[2] = prmt(sv.y, sv.y, 0x1010); c.s2[3] = prmt(sv.y, sv.y, 0x3232);
    if (FLOATZ) {
      const uint2 zv = lds64(zq + (size_t)gl * ZQ_ROW_BYTES + ncol0 * 2);
      c.z2[0] = prmt(zv.x, zv.x, 0x1010); c.z2[1] = prmt(zv.x, zv.x, 0x3232);
      c.z2[2] = prmt(zv.y, zv.y, 0x1010); c.z2[3] = prmt(zv.y, zv.y, 0x3232);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        c.c[i][0] = MAGIC;
        if (BITS == 4) c.c[i][1] = 0xD400D400u;
        if (BITS == 2) { c.c[i][1] = 0xDC00DC00u; c.c[i][2] = 0xD400D400u; c.c[i][3] = 0xCC00CC00u; }
      }
    } else {
      // 4 adjacent columns' zeros: 4*BITS consecutive bits of the packed zero row of this tile
      const int bitpos = ncol0 * BITS;
      const uint32_t word = lds32(zq + (size_t)gl * ZQ_ROW_BYTES + (bitpos >> 5) * 4);
      const uint32_t zs = word >> (bitpos & 31);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t z = (((zs >> (BITS * i)) & ((1u << BITS) - 1u)) + (uint32_t)L.zero_bias) & ((1u << BITS) - 1u);
        c.c[i][0] = (0x6400u | z) * 0x00010001u;                                   // 1024 + z
        if (BITS == 4) c.c[i][1] = (0xD400u + (z << 4)) * 0x00010001u;             // -(64 + z)
        if (BITS == 2) {
          c.c[i][1] = (0xDC00u + (z << 2)) * 0x00010001u;                          // -(256 + z)
          c.c[i][2] = (0xD400u + (z << 4)) * 0x00010001u;                          // -(64 + z)
          c.c[i][3] = (0xCC00u + (z << 6)) * 0x00010001u;                          // -(16 + z)
        }
      }
    }
  }

  // (q - z) * s for one extracted pair; TYPE selects the bias form of the pair
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

  // Unpack one word into P/2 fp16 pairs (already dequantised). Pair order = k-slot order used for x.
  __device__ static void unpack_word(uint32_t w, const Consts& c, int i, uint32_t (&o)[P / 2]) {
    if (BITS == 4) {
      const uint32_t hi = w >> 8;
      o[0] = finish<0>(and_or(w, LO4, MAGIC), c, i);    // (k0,k4)
      o[1] = finish<1>(and_or(w, HI4, MAGIC), c, i);    // (k1,k5)
      o[2] = finish<0>(and_or(hi, LO4, MAGIC), c, i);   // (k2,k6)
      o[3] = finish<1>(and_or(hi, HI4, MAGIC), c, i);   // (k3,k7)
    } else if (BITS == 8) {
      o[0] = finish<0>(prmt(w, MAGIC, 0x5150), c, i);   // (k0,k1): bytes [w.b0, 0x64, w.b1, 0x64]
      o[1] = finish<0>(prmt(w, MAGIC, 0x5352), c, i);   // (k2,k3)
    } else {                                             // 2-bit: pairs (i, i+8)
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

  // x pairs in the same order as unpack_word's pairs
  __device__ static void load_x(const char* xrow, int koff, bool active, uint32_t (&xb)[P / 2]) {
#pragma unroll
    for (int j = 0; j < P / 2; ++j) xb[j] = 0u;
    if (!active) return;
    if (BITS == 8) {
      const uint2 v = lds64(xrow + koff * 2);
      xb[0] = v.x; xb[1] = v.y;
    } else if (BITS == 4) {
      const uint4 v = lds128(xrow + koff * 2);
      xb[0] = prmt(v.x, v.z, 0x5410); xb[1] = prmt(v.x, v.z, 0x7632);
      xb[2] = prmt(v.y, v.w, 0x5410); xb[3] = prmt(v.y, v.w, 0x7632);
    } else {
      const uint4 a = lds128(xrow + koff * 2), b = lds128(xrow + koff * 2 + 16);
      xb[0] = prmt(a.x, b.x, 0x5410); xb[1] = prmt(a.x, b.x, 0x7632);
      xb[2] = prmt(a.y, b.y, 0x5410); xb[3] = prmt(a.y, b.y, 0x7632);
      xb[4] = prmt(a.z, b.z, 0x5410); xb[5] = prmt(a.z, b.z, 0x7632);
      xb[6] = prmt(a.w, b.w, 0x5410); xb[7] = prmt(a.w, b.w, 0x7632);
    }
  }

  __device__ static void step(const GemvParams& p, const uint32_t* sw, int st, int kabs, int k_cta0, int g_first,
                              const char* sc, const char* zq, const char* xs, Consts& c, float (&acc)[NSETS][4],
                              int warp, int lane, int ncols) {
    const int g = lane >> 2, t = lane & 3, nsub = warp & 1;
    if (nsub * 32 >= ncols) return;
    const int ncol0 = nsub * 32 + 4 * g;
    const int kthr = kabs + P * t;
    const int gi = kthr / p.L.group;
    if (gi != c.gcur) { c.gcur = gi; reload(c, p.L, sc, zq, gi - g_first, ncol0); }
    const uint4 w = lds128(sw + (size_t)(st * 4 + t) * RS_WORDS + ncol0);
    uint32_t xb[P / 2];
    load_x(xs + (size_t)g * p.x_stride, kthr - k_cta0, g < p.M, xb);
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

  __device__ static int ksub_of(int warp) { return warp >> 1; }

  // write accumulators into red[ksub][n_local][m]
  __device__ static void store_acc(float* red, const float (&acc)[NSETS][4], int warp, int lane, int M) {
    const int g = lane >> 2, t = lane & 3, nsub = warp & 1, ksub = warp >> 1;
    float* r = red + (size_t)ksub * NT * kMB;
#pragma unroll
    for (int s = 0; s < NSETS; ++s) {
      const int n = nsub * 32 + 4 * g + 2 * s;
      if (2 * t < M) { r[n * kMB + 2 * t] = acc[s][0]; r[(n + 1) * kMB + 2 * t] = acc[s][2]; }
      if (2 * t + 1 < M) { r[n * kMB + 2 * t + 1] = acc[s][1]; r[(n + 1) * kMB + 2 * t + 1] = acc[s][3]; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// AWQ "GEMM" layout (4-bit, qweight [K, N/8], nibble order 0,2,4,6,1,3,5,7).
// Tile: 512 columns (64 words) x 32 k rows per stage; warp w owns words 8w..8w+7; a step is 16 rows.
// Thread (g,t) reads its word column at rows 2t,2t+1 and 8+2t,9+2t; prmt pairs the two rows so that
// one lop3 yields (n@k, n@k+1) -- a k-pair for one column, which is what the MMA A fragment wants.
// ------------------------------------------------------------------------------------------------
struct AwqTraits {
  static constexpr int NT = 512, NW = 8, KW = 1;
  static constexpr int ROWS_PER_STEP = 16, KSTEP = 16, SPS = 2;
  static constexpr int ROW_WORDS = 64, RS_WORDS = 68;
  static constexpr int NSETS = 4;
  static constexpr int N_GRAN = 64;
  static constexpr int SC_ROW_BYTES = NT * 2;
  static constexpr int ZQ_ROW_BYTES = NT / 2;

  struct Consts {
    int gcur;
    uint32_t s2[8];
    uint32_t c[8];
  };

  __device__ static size_t tile_src_word(const LayerView& L, int row, int n0) { return (size_t)row * (L.N >> 3) + (n0 >> 3); }
  __device__ static int valid_row_bytes(int ncols) { return ncols / 2; }

  __device__ static void reload(Consts& c, const LayerView& L, const char* sc, const char* zq, int gl, int cw) {
    const uint4 sv = lds128(sc + (size_t)gl * SC_ROW_BYTES + cw * 16);
    c.s2[0] = prmt(sv.x, sv.x, 0x1010); c.s2[1] = prmt(sv.x, sv.x, 0x3232);
    c.s2[2] = prmt(sv.y, sv.y, 0x1010); c.s2[3] = prmt(sv.y, sv.y, 0x3232);
    c.s2[4] = prmt(sv.z, sv.z, 0x1010); c.s2[5] = prmt(sv.z, sv.z, 0x3232);
    c.s2[6] = prmt(sv.w, sv.w, 0x1010); c.s2[7] = prmt(sv.w, sv.w, 0x3232);
    const uint32_t zw = lds32(zq + (size_t)gl * ZQ_ROW_BYTES + cw * 4);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int nib = (n >> 1) + ((n & 1) << 2);      // column n sits in nibble inv[n]
      const int j = n >> 1;                           // MMA index; odd j uses the x16 form
      const uint32_t z = (((zw >> (4 * nib)) & 0xFu) + (uint32_t)L.zero_bias) & 0xFu;
      c.c[n] = (j & 1) ? (0xD400u + (z << 4)) * 0x00010001u : (0x6400u | z) * 0x00010001u;
    }
  }

  __device__ static void step(const GemvParams& p, const uint32_t* sw, int st, int kabs, int k_cta0, int g_first,
                              const char* sc, const char* zq, const char* xs, Consts& c, float (&acc)[NSETS][4],
                              int warp, int lane, int ncols) {
    const int g = lane >> 2, t = lane & 3;
    if (warp * 64 >= ncols) return;
    const int cw = warp * 8 + g;
    const int gi = kabs / p.L.group;
    if (gi != c.gcur) { c.gcur = gi; reload(c, p.L, sc, zq, gi - g_first, cw); }
    const uint32_t* base = sw + (size_t)(st * 16 + 2 * t) * RS_WORDS + cw;
    const uint32_t wa = base[0], wb = base[RS_WORDS], wc = base[8 * RS_WORDS], wd = base[9 * RS_WORDS];
    uint32_t b0 = 0u, b1 = 0u;
    if (g < p.M) {
      const char* xr = xs + (size_t)g * p.x_stride + (size_t)(kabs - k_cta0 + 2 * t) * 2;
      b0 = lds32(xr);
      b1 = lds32(xr + 16);
    }
    const uint32_t u01 = prmt(wa, wb, 0x5410), v01 = prmt(wa, wb, 0x7632);
    const uint32_t u23 = prmt(wc, wd, 0x5410), v23 = prmt(wc, wd, 0x7632);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int sh = (j >> 1) * 8;
      const uint32_t msk = (j & 1) ? HI4 : LO4;
      uint32_t a0 = and_or(u01 >> sh, msk, MAGIC), a1 = and_or(v01 >> sh, msk, MAGIC);
      uint32_t a2 = and_or(u23 >> sh, msk, MAGIC), a3 = and_or(v23 >> sh, msk, MAGIC);
      const int na = 2 * j, nb = 2 * j + 1;
      if (j & 1) {
        a0 = hfma2_u(a0, H_1_16, c.c[na]); a1 = hfma2_u(a1, H_1_16, c.c[nb]);
        a2 = hfma2_u(a2, H_1_16, c.c[na]); a3 = hfma2_u(a3, H_1_16, c.c[nb]);
      } else {
        a0 = hsub2_u(a0, c.c[na]); a1 = hsub2_u(a1, c.c[nb]);
        a2 = hsub2_u(a2, c.c[na]); a3 = hsub2_u(a3, c.c[nb]);
      }
      a0 = hmul2_u(a0, c.s2[na]); a1 = hmul2_u(a1, c.s2[nb]);
      a2 = hmul2_u(a2, c.s2[na]); a3 = hmul2_u(a3, c.s2[nb]);
      mma_16816(acc[j], a0, a1, a2, a3, b0, b1);
    }
  }

  __device__ static int ksub_of(int) { return 0; }

  __device__ static void store_acc(float* red, const float (&acc)[NSETS][4], int warp, int lane, int M) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = (warp * 8 + g) * 8 + 2 * j;
      if (2 * t < M) { red[n * kMB + 2 * t] = acc[j][0]; red[(n + 1) * kMB + 2 * t] = acc[j][2]; }
      if (2 * t + 1 < M) { red[n * kMB + 2 * t + 1] = acc[j][1]; red[(n + 1) * kMB + 2 * t + 1] = acc[j][3]; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Marlin layout (symmetric 4-bit, qweight [K/16, 2N]): a word already IS one lane's A fragment of a
// 16(n) x 16(k) tile (natural k order).  Tile: 256 columns (512 words) x 4 k16-rows per stage.
// warp = (blk = w & 3 -> 64 columns, ksub = w >> 2); one LDS.128 = 4 MMAs.
// ------------------------------------------------------------------------------------------------
struct MarlinTraits {
  static constexpr int NT = 256, NW = 4, KW = 2;
  static constexpr int ROWS_PER_STEP = 1, KSTEP = 16, SPS = 4;
  static constexpr int ROW_WORDS = 512, RS_WORDS = 512;
  static constexpr int NSETS = 4;
  static constexpr int N_GRAN = 64;
  static constexpr int SC_ROW_BYTES = NT * 2;
  static constexpr int ZQ_ROW_BYTES = 0;

  struct Consts {
    int gcur;
    uint32_t s2[8];   // [2j] = column 16j+g, [2j+1] = column 16j+g+8
  };

  __device__ static size_t tile_src_word(const LayerView& L, int row, int n0) { return (size_t)row * (2 * (size_t)L.N) + 2 * n0; }
  __device__ static int valid_row_bytes(int ncols) { return ncols * 8; }

  __device__ static void reload(Consts& c, const LayerView& L, const char* sc, int gl, int blk, int g) {
    if (L.group != L.K) {       // grouped: stored[64m + 8a + b] = s[64m + 8b + a]
      const uint4 sv = lds128(sc + (size_t)gl * SC_ROW_BYTES + (blk * 64 + 8 * g) * 2);
      c.s2[0] = prmt(sv.x, sv.x, 0x1010); c.s2[1] = prmt(sv.x, sv.x, 0x3232);
      c.s2[2] = prmt(sv.y, sv.y, 0x1010); c.s2[3] = prmt(sv.y, sv.y, 0x3232);
      c.s2[4] = prmt(sv.z, sv.z, 0x1010); c.s2[5] = prmt(sv.z, sv.z, 0x3232);
      c.s2[6] = prmt(sv.w, sv.w, 0x1010); c.s2[7] = prmt(sv.w, sv.w, 0x3232);
    } else {                    // per-channel permutation
      const __half* row = reinterpret_cast<const __half*>(sc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = blk * 64 + 16 * j + g;
        c.s2[2 * j] = dup_half(row[marlin_scale_index(n, true)]);
        c.s2[2 * j + 1] = dup_half(row[marlin_scale_index(n + 8, true)]);
      }
    }
  }

  __device__ static void step(const GemvParams& p, const uint32_t* sw, int st, int kabs, int k_cta0, int g_first,
                              const char* sc, const char*, const char* xs, Consts& c, float (&acc)[NSETS][4],
                              int warp, int lane, int ncols) {
    const int g = lane >> 2, t = lane & 3, blk = warp & 3;
    if (blk * 64 >= ncols) return;
    const int gi = kabs / p.L.group;
    if (gi != c.gcur) { c.gcur = gi; reload(c, p.L, sc, gi - g_first, blk, g); }
    const uint4 w = lds128(sw + (size_t)st * RS_WORDS + blk * 128 + 4 * lane);
    uint32_t b0 = 0u, b1 = 0u;
    if (g < p.M) {
      const char* xr = xs + (size_t)g * p.x_stride + (size_t)(kabs - k_cta0 + 2 * t) * 2;
      b0 = lds32(xr);
      b1 = lds32(xr + 16);
    }
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t lo = ww[j], hi = ww[j] >> 8;
      uint32_t a0 = hsub2_u(and_or(lo, LO4, MAGIC), 0x64086408u);               // q - 8
      uint32_t a2 = hfma2_u(and_or(lo, HI4, MAGIC), H_1_16, 0xD480D480u);       // q - 8 via /16 - 72
      uint32_t a1 = hsub2_u(and_or(hi, LO4, MAGIC), 0x64086408u);
      uint32_t a3 = hfma2_u(and_or(hi, HI4, MAGIC), H_1_16, 0xD480D480u);
      a0 = hmul2_u(a0, c.s2[2 * j]); a2 = hmul2_u(a2, c.s2[2 * j]);
      a1 = hmul2_u(a1, c.s2[2 * j + 1]); a3 = hmul2_u(a3, c.s2[2 * j + 1]);
      mma_16816(acc[j], a0, a1, a2, a3, b0, b1);
    }
  }

  __device__ static int ksub_of(int warp) { return warp >> 2; }

  __device__ static void store_acc(float* red, const float (&acc)[NSETS][4], int warp, int lane, int M) {
    const int g = lane >> 2, t = lane & 3, blk = warp & 3, ksub = warp >> 2;
    float* r = red + (size_t)ksub * NT * kMB;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = blk * 64 + 16 * j + g;
      if (2 * t < M) { r[n * kMB + 2 * t] = acc[j][0]; r[(n + 8) * kMB + 2 * t] = acc[j][2]; }
      if (2 * t + 1 < M) { r[n * kMB + 2 * t + 1] = acc[j][1]; r[(n + 8) * kMB + 2 * t + 1] = acc[j][3]; }
    }
  }
};

template <class T>
struct IsMarlin { static constexpr bool v = false; };
template <>
struct IsMarlin<MarlinTraits> { static constexpr bool v = true; };

// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kThreads) gemv_mma_kernel(const GemvParams p) {
  extern __shared__ __align__(128) char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kStages;
  uint64_t* xbar = empty + kStages;
  char* stage0 = smem + p.off_stage;
  char* xs = smem + p.off_x;
  char* sc = smem + p.off_sc;
  char* zq = smem + p.off_zq;
  float* red = reinterpret_cast<float*>(smem + p.off_red);
  int* flag = reinterpret_cast<int*>(smem + p.off_flag);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.x % p.n_tiles, ks = blockIdx.x / p.n_tiles;
  const int n0 = n_tile * T::NT;
  const int ncols = min(T::NT, p.L.N - n0);
  const int step0 = ks * p.steps_per_cta;
  const int nsteps = min(p.steps_per_cta, p.steps_total - step0);
  const int k_cta0 = step0 * T::KSTEP;
  const int k_cta1 = k_cta0 + nsteps * T::KSTEP;
  const int g_first = k_cta0 / p.L.group;
  const int g_count = (k_cta1 - 1) / p.L.group - g_first + 1;
  const int row0 = step0 * T::ROWS_PER_STEP;
  const int nrows = nsteps * T::ROWS_PER_STEP;
  constexpr int SR = T::SPS * T::ROWS_PER_STEP;           // packed rows per stage
  const int nstages = (nrows + SR - 1) / SR;
  constexpr int STAGE_BYTES = SR * T::RS_WORDS * 4;

  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    mbar_init(xbar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ===== producer warp =====
    const int row_bytes = T::valid_row_bytes(ncols);
    auto issue_stage = [&](int i) {
      const int slot = i % kStages;
      const int r0 = i * SR;
      const int rows = min(SR, nrows - r0);
      if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)(rows * row_bytes));
      __syncwarp();
      for (int r = lane; r < rows; r += 32)
        bulk_g2s(stage0 + (size_t)slot * STAGE_BYTES + (size_t)r * T::RS_WORDS * 4,
                 p.L.qw + T::tile_src_word(p.L, row0 + r0 + r, n0), (uint32_t)row_bytes, &full[slot]);
    };
    const int pre = min(nstages, kStages);
    for (int i = 0; i < pre; ++i) issue_stage(i);
    // activations depend on the upstream kernel
    pdl_wait();
    if (lane == 0) mbar_expect_tx(xbar, (uint32_t)(p.M * (k_cta1 - k_cta0) * 2));
    __syncwarp();
    for (int m = lane; m < p.M; m += 32)
      bulk_g2s(xs + (size_t)m * p.x_stride, p.x + (size_t)m * p.ldx + k_cta0, (uint32_t)((k_cta1 - k_cta0) * 2), xbar);
    for (int i = pre; i < nstages; ++i) {
      const int slot = i % kStages;
      mbar_wait(&empty[slot], ((i / kStages) + 1) & 1);
      issue_stage(i);
    }
    return;
  }

  // ===== consumer warps =====
  // group constants of this CTA's k range: plain cooperative loads (tiny, arbitrary alignment)
  {
    const int ctid = tid;  // 0..255
    const int sc_halfs = ncols;
    for (int idx = ctid; idx < g_count * sc_halfs; idx += kConsumerWarps * 32) {
      const int gl = idx / sc_halfs, n = idx % sc_halfs;
      reinterpret_cast<__half*>(sc + (size_t)gl * T::SC_ROW_BYTES)[n] = __ldg(p.L.s + (size_t)(g_first + gl) * p.L.N + n0 + n);
    }
    if (!IsMarlin<T>::v) {
      if (p.L.layout == B200Q_LAYOUT_HQQ) {
        for (int idx = ctid; idx < g_count * ncols; idx += kConsumerWarps * 32) {
          const int gl = idx / ncols, n = idx % ncols;
          reinterpret_cast<__half*>(zq + (size_t)gl * T::ZQ_ROW_BYTES)[n] =
              __ldg(reinterpret_cast<const __half*>(p.L.qz) + (size_t)(g_first + gl) * p.L.N + n0 + n);
        }
      } else {
        const int zwords = (ncols * p.L.bits) >> 5;                 // packed zero words of this tile per group
        const size_t zrow = ((size_t)p.L.N * p.L.bits) >> 5;        // words per qzeros row
        const size_t zoff = ((size_t)n0 * p.L.bits) >> 5;
        for (int idx = ctid; idx < g_count * zwords; idx += kConsumerWarps * 32) {
          const int gl = idx / zwords, wv = idx % zwords;
          reinterpret_cast<uint32_t*>(zq + (size_t)gl * T::ZQ_ROW_BYTES)[wv] =
              __ldg(reinterpret_cast<const uint32_t*>(p.L.qz) + (size_t)(g_first + gl) * zrow + zoff + wv);
        }
      }
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");

  float acc[T::NSETS][4];
#pragma unroll
  for (int s = 0; s < T::NSETS; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[s][i] = 0.f;
  typename T::Consts c;
  c.gcur = -1;

  mbar_wait(xbar, 0);
  const int ksub = T::ksub_of(warp);
  for (int i = 0; i < nstages; ++i) {
    const int slot = i % kStages;
    mbar_wait(&full[slot], (i / kStages) & 1);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(stage0 + (size_t)slot * STAGE_BYTES);
    const int steps_here = min(T::SPS, nsteps - i * T::SPS);
    for (int st = ksub; st < steps_here; st += T::KW)
      T::step(p, sw, st, k_cta0 + (i * T::SPS + st) * T::KSTEP, k_cta0, g_first, sc, zq, xs, c, acc, warp, lane, ncols);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
  }

  // ---- reduce over k-split warps, then over k-split CTAs ----
  T::store_acc(red, acc, warp, lane, p.M);
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
  const int total = ncols * p.M;
  if (p.ksplit == 1) {
    for (int idx = tid; idx < total; idx += kConsumerWarps * 32) {
      const int m = idx / ncols, n = idx % ncols;
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < T::KW; ++k) v += red[((size_t)k * T::NT + n) * kMB + m];
      if (p.L.bias) v += __half2float(__ldg(p.L.bias + n0 + n));
      const __half h = __float2half_rn(v);
      for (int q = 0; q < p.out.n; ++q) p.out.y[q][(size_t)m * p.ldy + p.n_offset + n0 + n] = h;
    }
    return;
  }
  for (int idx = tid; idx < total; idx += kConsumerWarps * 32) {
    const int m = idx / ncols, n = idx % ncols;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < T::KW; ++k) v += red[((size_t)k * T::NT + n) * kMB + m];
    p.partial[((size_t)ks * p.M + m) * p.L.N + n0 + n] = v;
  }
  __threadfence();
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
  if (tid == 0) {
    const unsigned old = atomicAdd(&p.counters[n_tile], 1u);
    *flag = (old == (unsigned)(p.ksplit - 1));
    if (*flag) p.counters[n_tile] = 0u;       // leave the workspace zeroed for the next call
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
  if (!*flag) return;
  __threadfence();
  for (int idx = tid; idx < total; idx += kConsumerWarps * 32) {
    const int m = idx / ncols, n = idx % ncols;
    float v = 0.f;
    for (int s = 0; s < p.ksplit; ++s) v += __ldcg(&p.partial[((size_t)s * p.M + m) * p.L.N + n0 + n]);
    if (p.L.bias) v += __half2float(__ldg(p.L.bias + n0 + n));
    const __half h = __float2half_rn(v);
    for (int q = 0; q < p.out.n; ++q) p.out.y[q][(size_t)m * p.ldy + p.n_offset + n0 + n] = h;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Plan {
  int kind;  // 0 none, 1 gptq2, 2 gptq4, 3 gptq8, 4 awq, 5 marlin ; +16 for float zeros
  int NT, KSTEP, SPS, ROWS_PER_STEP, RS_WORDS, KW, SC_ROW, ZQ_ROW, n_gran;
  int n_tiles, ksplit, steps_total, steps_per_cta;
  int gcap, x_stride;
  int off_stage, off_x, off_sc, off_zq, off_red, off_flag, smem_bytes;
};

template <class T>
static void fill_traits(Plan& pl) {
  pl.NT = T::NT; pl.KSTEP = T::KSTEP; pl.SPS = T::SPS; pl.ROWS_PER_STEP = T::ROWS_PER_STEP; pl.RS_WORDS = T::RS_WORDS;
  pl.KW = T::KW; pl.SC_ROW = T::SC_ROW_BYTES; pl.ZQ_ROW = T::ZQ_ROW_BYTES; pl.n_gran = T::N_GRAN;
}

static bool make_plan(const LayerView& L, int M, Plan& pl) {
  pl.kind = 0;
  if (M < 1 || M > kMB || L.g_idx != nullptr || L.x_perm != nullptr) return false;
  const bool fz = (L.layout == B200Q_LAYOUT_HQQ);
  if (L.layout == B200Q_LAYOUT_GPTQ || L.layout == B200Q_LAYOUT_HQQ) {
    if (L.bits == 2) { pl.kind = 1; fz ? fill_traits<GptqTraits<2, true>>(pl) : fill_traits<GptqTraits<2, false>>(pl); }
    else if (L.bits == 4) { pl.kind = 2; fz ? fill_traits<GptqTraits<4, true>>(pl) : fill_traits<GptqTraits<4, false>>(pl); }
    else if (L.bits == 8) { pl.kind = 3; fz ? fill_traits<GptqTraits<8, true>>(pl) : fill_traits<GptqTraits<8, false>>(pl); }
    else return false;
    if (fz) pl.kind += 16;
    if (L.group % (32 / L.bits) != 0) return false;
  } else if (L.layout == B200Q_LAYOUT_AWQ_GEMM) {
    if (L.bits != 4 || L.group % 16 != 0) return false;
    pl.kind = 4; fill_traits<AwqTraits>(pl);
  } else if (L.layout == B200Q_LAYOUT_MARLIN) {
    if (L.bits != 4 || L.group % 16 != 0) return false;
    pl.kind = 5; fill_traits<MarlinTraits>(pl);
  } else return false;
  if (L.K % pl.KSTEP != 0 || L.N % pl.n_gran != 0 || L.K % L.group != 0) { pl.kind = 0; return false; }
  pl.n_tiles = (L.N + pl.NT - 1) / pl.NT;
  pl.steps_total = L.K / pl.KSTEP;
  // split K so that ~2 CTAs per SM are resident, each CTA owning >= 2 stages when possible
  int ks = (2 * 148 + pl.n_tiles - 1) / pl.n_tiles;
  const int max_ks = (pl.steps_total + pl.SPS - 1) / pl.SPS;       // at least one stage per CTA
  if (ks > max_ks) ks = max_ks;
  if (ks < 1) ks = 1;
  int spc = (pl.steps_total + ks - 1) / ks;
  spc = (spc + pl.SPS - 1) / pl.SPS * pl.SPS;                      // whole stages per CTA
  // bound the activation slice held in shared memory (M rows)
  const int max_k_cta = (96 * 1024) / (2 * M);
  while (spc * pl.KSTEP > max_k_cta && spc > pl.SPS) spc -= pl.SPS;
  // keep every CTA's k-range aligned to whole groups' constant slices cheaply: no constraint needed
  pl.steps_per_cta = spc;
  pl.ksplit = (pl.steps_total + spc - 1) / spc;
  const int k_cta = spc * pl.KSTEP;
  pl.gcap = k_cta / L.group + 2;
  pl.x_stride = k_cta * 2;
  pl.x_stride += (64 - (pl.x_stride % 128) + 128) % 128;            // row stride == 64 (mod 128) bytes
  int off = 128;                                                    // mbarriers
  pl.off_stage = off; off += kStages * pl.SPS * pl.ROWS_PER_STEP * pl.RS_WORDS * 4;
  pl.off_x = off; off += M * pl.x_stride;
  off = (off + 15) & ~15;
  pl.off_sc = off; off += pl.gcap * pl.SC_ROW;
  off = (off + 15) & ~15;
  pl.off_zq = off; off += pl.gcap * pl.ZQ_ROW;
  off = (off + 15) & ~15;
  pl.off_red = off; off += pl.KW * pl.NT * kMB * 4;
  pl.off_flag = off; off += 16;
  pl.smem_bytes = off;
  if (pl.smem_bytes > 200 * 1024) { pl.kind = 0; return false; }
  return true;
}

bool gemv_mma_supported(const LayerView& L, int M, const __half* x, int64_t ldx) {
  Plan pl;
  if (!make_plan(L, M, pl)) return false;
  if (((uintptr_t)x & 15) != 0 || (ldx % 8) != 0) return false;
  if (((uintptr_t)L.qw & 15) != 0) return false;
  return true;
}


size_t gemv_mma_workspace(const LayerView& L, int M) {
  Plan pl;
  if (!make_plan(L, M, pl)) return 0;
  return kCounterBytes + (size_t)pl.ksplit * M * L.N * sizeof(float);
}

template <class T>
static cudaError_t launch_t(const GemvParams& p, const Plan& pl, cudaStream_t st) {
  static bool attr_done[64] = {};   // per instantiation and device; idempotent
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemv_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.n_tiles * pl.ksplit);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  count_launch();
  return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<T>, p);
}

cudaError_t launch_gemv_mma(const LinearArgs& a, const PeerOut* peers) {
  Plan pl;
  if (!make_plan(a.L, a.M, pl)) return cudaErrorInvalidValue;
  if (pl.n_tiles * 4 > (int)kCounterBytes) return cudaErrorInvalidValue;
  GemvParams p = {};
  p.L = a.L; p.x = a.x; p.ldx = a.ldx; p.M = a.M;
  if (peers) p.out = *peers; else { p.out.n = 1; p.out.y[0] = a.y; }
  p.ldy = a.ldy; p.n_offset = a.n_offset;
  p.n_tiles = pl.n_tiles; p.ksplit = pl.ksplit; p.steps_total = pl.steps_total; p.steps_per_cta = pl.steps_per_cta;
  p.counters = (unsigned*)a.workspace;
  p.partial = (float*)((char*)a.workspace + kCounterBytes);
  p.off_stage = pl.off_stage; p.off_x = pl.off_x; p.x_stride = pl.x_stride; p.off_sc = pl.off_sc; p.off_zq = pl.off_zq;
  p.off_red = pl.off_red; p.off_flag = pl.off_flag; p.gcap = pl.gcap;
  switch (pl.kind) {
    case 1: return launch_t<GptqTraits<2, false>>(p, pl, a.stream);
    case 2: return launch_t<GptqTraits<4, false>>(p, pl, a.stream);
    case 3: return launch_t<GptqTraits<8, false>>(p, pl, a.stream);
    case 17: return launch_t<GptqTraits<2, true>>(p, pl, a.stream);
    case 18: return launch_t<GptqTraits<4, true>>(p, pl, a.stream);
    case 19: return launch_t<GptqTraits<8, true>>(p, pl, a.stream);
    case 4: return launch_t<AwqTraits>(p, pl, a.stream);
    case 5: return launch_t<MarlinTraits>(p, pl, a.stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace b200q
