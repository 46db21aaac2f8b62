// Decode kernel for M <= 2: CUDA-core fp16x2 FMA formulation (the "warp GEMV").
//
// At one or two activation rows an mma.sync tile is >= 7/8 padding, and on sm_100 the legacy HMMA path
// is slow enough that the padded MMAs, not HBM, bound the decode kernel (profiles/README.md).  Here the
// unpacked pairs are multiplied by the matching activation pair with HFMA2 instead:
//
//   * every lane owns a fixed set of output columns and walks down K, so partial sums never cross lanes
//     (GPTQ: 4 columns per lane, one 128-bit load = 4 columns x 8 k; AWQ: 8 columns per lane, one 32-bit
//     word = 8 columns x 1 k; a warp-wide load is 512 B / 128 B of contiguous packed weights);
//   * (q - z) is formed exactly in fp16 with the lop3 magic-number trick (per-group constants), products
//     are accumulated in fp16x2 for at most 8 terms, then flushed into fp32 with the group scale
//     (tot += s * acc), so the error budget stays ~1e-4 of max|y| (tests: 1e-3);
//   * all packed-weight loads of a warp (<= kMaxLoads per lane) and its group constants are issued
//     BEFORE griddepcontrol.wait -- with programmatic dependent launch the next layer's weights are
//     already in registers when its activations arrive; x is staged in shared memory (pre-permuted /
//     duplicated into the pair order the unpack produces) after the wait;
//   * 8 warps split K (shared-memory reduce), clusters of <= 8 CTAs split K further (distributed shared
//     memory reduce), fixed summation order -> deterministic.
// No repacking of checkpoint bytes.  Replaces ort_ops.gemv (dq_gemv.cu:40-150) and gemm_forward_cuda at
// M <= 2 (gemm_cuda_gen.cu:31-353).
#include "common.cuh"
#include "kernels.h"

namespace b200q {

static constexpr int kFWarps = 8;
static constexpr int kFThreads = kFWarps * 32;
static constexpr uint32_t MAGIC = 0x64006400u, LO4 = 0x000f000fu, HI4 = 0x00f000f0u;
static constexpr uint32_t H_1_16 = 0x2c002c00u, H_M1_16 = 0xac00ac00u;

struct FmaParams {
  LayerView L;
  const __half* x;
  int64_t ldx;
  int M;
  PeerOut out;
  int64_t ldy, n_offset;
  int n_tiles, cluster, rows_total, group_shift;
  int x_stride;     // bytes of one staged x row
  int off_x, off_red, off_rbuf;
};

__device__ __forceinline__ uint4 f_ldg128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t f_ldg32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t f_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void f_st_cluster(const float* local_smem, uint32_t rank, float v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
__device__ __forceinline__ void f_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float h2_lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xFFFFu))); }
__device__ __forceinline__ float h2_hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }

// ------------------------------------------------------------------------------------------------
// GPTQ 4-bit: rows are packed words of 8 k; lane l owns columns n0 + 4l .. 4l+3 (CTA tile 128 columns).
// Staged x order inside each 8-block: (x0,x4,x1,x5,x2,x6,x3,x7) = the lop3 pair order.
// ------------------------------------------------------------------------------------------------
struct FGptq4 {
  static constexpr int NT = 128, COLS = 4, KROW = 8, MAXLD = 16, FLUSH_ROWS = 2, XDUP = 1;
  static constexpr int LPR = 32, RPL = 1;        // lanes per packed row, rows per warp-wide load
  using Load = uint4;
  struct GC {           // per-group constants of this lane's 4 columns
    uint32_t clo[4], chi[4];
    float s[4];
  };
  struct Raw { uint32_t z; uint2 s; };

  __device__ static void load(Load& w, const LayerView& L, int row, int n0, int lane) {
    w = f_ldg128(L.qw + (size_t)row * L.N + n0 + 4 * lane);
  }
  __device__ static void load_raw(Raw& r, const LayerView& L, int g, int n0, int lane) {
    const int n = n0 + 4 * lane;
    const uint32_t zw = __ldg((const uint32_t*)L.qz + (size_t)g * (L.N >> 3) + (n >> 3));
    r.z = (zw >> (4 * (n & 7))) & 0xFFFFu;
    r.s = __ldg(reinterpret_cast<const uint2*>(L.s + (size_t)g * L.N + n));
  }
  __device__ static void make_gc(GC& c, const Raw& r, int zero_bias) {
    const uint32_t sv[2] = {r.s.x, r.s.y};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t z = (((r.z >> (4 * i)) & 0xFu) + (uint32_t)zero_bias) & 0xFu;
      c.clo[i] = (0x6400u | z) * 0x00010001u;
      c.chi[i] = (0xD400u + (z << 4)) * 0x00010001u;
      c.s[i] = (i & 1) ? h2_hi(sv[i >> 1]) : h2_lo(sv[i >> 1]);
    }
  }
  // x staging: element k of the slice goes to position perm(k) so that an LDS.128 yields the 4 pairs
  __device__ static int x_pos(int k) { const int i = k & 7; return (k & ~7) + ((i & 3) << 1) + (i >> 2); }

  template <int MB>
  __device__ static void row(const Load& w, const char* xs, int x_stride, int krel, const GC& c, uint32_t (&acc)[MB][4]) {
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    uint4 xp[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) xp[m] = *reinterpret_cast<const uint4*>(xs + (size_t)m * x_stride + krel * 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t lo = ww[i], hi = ww[i] >> 8;
      const uint32_t d0 = hsub2_u(and_or(lo, LO4, MAGIC), c.clo[i]);
      const uint32_t d1 = hfma2_u(and_or(lo, HI4, MAGIC), H_1_16, c.chi[i]);
      const uint32_t d2 = hsub2_u(and_or(hi, LO4, MAGIC), c.clo[i]);
      const uint32_t d3 = hfma2_u(and_or(hi, HI4, MAGIC), H_1_16, c.chi[i]);
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        uint32_t a = acc[m][i];
        a = hfma2_u(d0, xp[m].x, a);
        a = hfma2_u(d1, xp[m].y, a);
        a = hfma2_u(d2, xp[m].z, a);
        a = hfma2_u(d3, xp[m].w, a);
        acc[m][i] = a;
      }
    }
  }
  template <int MB>
  __device__ static void flush(float (&tot)[MB][4], uint32_t (&acc)[MB][4], const GC& c) {
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        tot[m][i] = fmaf(c.s[i], h2_lo(acc[m][i]) + h2_hi(acc[m][i]), tot[m][i]);
        acc[m][i] = 0u;
      }
  }
  __device__ static int col_of(int lane, int i) { return 4 * lane + i; }
};

// ------------------------------------------------------------------------------------------------
// AWQ GEMM 4-bit: rows are k; a half-warp covers one row: lane l owns word column n0/8 + (l & 15), i.e.
// columns n0 + 8(l&15) .. +7 (CTA tile 128 columns), and the two half-warps take alternate rows (their
// partial sums meet in the final reduction).  lop3 pairs are adjacent columns: (0,1) LO, (2,3) x16,
// (4,5) LO, (6,7) x16.  x is staged duplicated (x_k, x_k).
// ------------------------------------------------------------------------------------------------
struct FAwq4 {
  static constexpr int NT = 128, COLS = 8, KROW = 1, MAXLD = 32, FLUSH_ROWS = 8, XDUP = 2;
  static constexpr int LPR = 16, RPL = 2;
  using Load = uint32_t;
  struct GC {
    uint32_t c[4];
    float s[8];
  };
  struct Raw { uint32_t z; uint4 s; };

  __device__ static void load(Load& w, const LayerView& L, int row, int n0, int lane) {
    w = f_ldg32(L.qw + (size_t)row * (L.N >> 3) + (n0 >> 3) + (lane & 15));
  }
  __device__ static void load_raw(Raw& r, const LayerView& L, int g, int n0, int lane) {
    r.z = __ldg((const uint32_t*)L.qz + (size_t)g * (L.N >> 3) + (n0 >> 3) + (lane & 15));
    r.s = __ldg(reinterpret_cast<const uint4*>(L.s + (size_t)g * L.N + n0 + 8 * (lane & 15)));
  }
  __device__ static void make_gc(GC& c, const Raw& r, int zero_bias) {
    uint32_t zw = r.z;
    if (zero_bias) zw = ((zw & 0x77777777u) + 0x11111111u) ^ (zw & 0x88888888u);      // nibble-wise (z+1)&15
    const uint32_t zh = zw >> 8;
    c.c[0] = and_or(zw, LO4, MAGIC);                                                   // (1024+z0, 1024+z1)
    c.c[1] = hmul2_u(and_or(zw, HI4, MAGIC), H_M1_16);                                 // -(64+z2), -(64+z3)
    c.c[2] = and_or(zh, LO4, MAGIC);
    c.c[3] = hmul2_u(and_or(zh, HI4, MAGIC), H_M1_16);
    const uint32_t sv[4] = {r.s.x, r.s.y, r.s.z, r.s.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { c.s[2 * i] = h2_lo(sv[i]); c.s[2 * i + 1] = h2_hi(sv[i]); }
  }
  __device__ static int x_pos(int k) { return k; }          // duplicated layout handled by XDUP

  template <int MB>
  __device__ static void row(const Load& w, const char* xs, int x_stride, int krel, const GC& c, uint32_t (&acc)[MB][4]) {
    const uint32_t lo = w, hi = w >> 8;
    const uint32_t d0 = hsub2_u(and_or(lo, LO4, MAGIC), c.c[0]);
    const uint32_t d1 = hfma2_u(and_or(lo, HI4, MAGIC), H_1_16, c.c[1]);
    const uint32_t d2 = hsub2_u(and_or(hi, LO4, MAGIC), c.c[2]);
    const uint32_t d3 = hfma2_u(and_or(hi, HI4, MAGIC), H_1_16, c.c[3]);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const uint32_t xd = *reinterpret_cast<const uint32_t*>(xs + (size_t)m * x_stride + krel * 4);
      acc[m][0] = hfma2_u(d0, xd, acc[m][0]);
      acc[m][1] = hfma2_u(d1, xd, acc[m][1]);
      acc[m][2] = hfma2_u(d2, xd, acc[m][2]);
      acc[m][3] = hfma2_u(d3, xd, acc[m][3]);
    }
  }
  template <int MB>
  __device__ static void flush(float (&tot)[MB][8], uint32_t (&acc)[MB][4], const GC& c) {
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        tot[m][2 * j] = fmaf(c.s[2 * j], h2_lo(acc[m][j]), tot[m][2 * j]);
        tot[m][2 * j + 1] = fmaf(c.s[2 * j + 1], h2_hi(acc[m][j]), tot[m][2 * j + 1]);
        acc[m][j] = 0u;
      }
  }
  __device__ static int col_of(int lane, int i) { return 8 * (lane & 15) + i; }
};

// ------------------------------------------------------------------------------------------------
template <class T, int MB>
__global__ void __launch_bounds__(kFThreads, 2) gemv_fma_kernel(const FmaParams p) {
  extern __shared__ __align__(16) char smem[];
  char* xs = smem + p.off_x;
  float* red = reinterpret_cast<float*>(smem + p.off_red);
  float* rbuf = reinterpret_cast<float*>(smem + p.off_rbuf);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = p.cluster;
  const int rank = (int)f_ctarank();
  const int n_tile = blockIdx.x / cs;
  const int n0 = n_tile * T::NT;
  const int ncols = min(T::NT, p.L.N - n0);
  const bool lane_ok = (lane % T::LPR) * T::COLS < ncols;
  const int sub = lane / T::LPR;                         // which of the RPL rows of a warp-wide load
  const int U = cs * kFWarps, R = p.rows_total;
  const int unit = rank * kFWarps + warp;
  // rows are split in multiples of one flush block (FLUSH_ROWS loads x RPL rows) so that a block never
  // straddles two warps or two groups
  constexpr int BLK = T::FLUSH_ROWS * T::RPL;
  const int B = R / BLK;
  const int r_begin = (int)(((long long)unit * B) / U) * BLK, r_end = (int)(((long long)(unit + 1) * B) / U) * BLK;
  const int cta_r0 = (int)(((long long)rank * kFWarps * B) / U) * BLK;
  const int cta_r1 = (int)(((long long)(rank + 1) * kFWarps * B) / U) * BLK;
  const int nld = (r_end - r_begin) / T::RPL;            // loads per lane
  const int k_cta0 = cta_r0 * T::KROW, kslice = (cta_r1 - cta_r0) * T::KROW;
  auto group_of = [&](int k) { return p.group_shift >= 0 ? (k >> p.group_shift) : (k / p.L.group); };

  pdl_launch_dependents();

  // ---- 1. this warp's packed weights + group constants -> registers (independent of upstream) ----
  typename T::Load w[T::MAXLD];
#pragma unroll
  for (int i = 0; i < T::MAXLD; ++i) {
    w[i] = typename T::Load{};
    if (i < nld && lane_ok) T::load(w[i], p.L, r_begin + i * T::RPL + sub, n0, lane);
  }
  constexpr int NG = 3;                                  // groups whose constants are pre-loaded
  const int g_first = (r_begin < r_end) ? group_of(r_begin * T::KROW) : 0;
  typename T::Raw raw[NG];
#pragma unroll
  for (int j = 0; j < NG; ++j) {
    raw[j] = typename T::Raw{};
    if (lane_ok && g_first + j < p.L.G) T::load_raw(raw[j], p.L, g_first + j, n0, lane);
  }

  // ---- 2. activations: wait for the upstream kernel, stage x (pair order / duplicated) ----
  pdl_wait();
  for (int idx = tid; idx < p.M * kslice; idx += kFThreads) {
    const int m = idx / kslice, k = idx % kslice;
    const __half v = p.x[(size_t)m * p.ldx + k_cta0 + k];
    __half* dst = reinterpret_cast<__half*>(xs + (size_t)m * p.x_stride);
    if (T::XDUP == 2) { dst[2 * k] = v; dst[2 * k + 1] = v; }
    else dst[T::x_pos(k)] = v;
  }
  __syncthreads();

  // ---- 3. math ----
  float tot[MB][T::COLS];
  uint32_t acc[MB][4];
#pragma unroll
  for (int m = 0; m < MB; ++m) {
#pragma unroll
    for (int i = 0; i < T::COLS; ++i) tot[m][i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[m][i] = 0u;
  }
  typename T::GC gc;
  int gcur = -1;
  auto set_group = [&](int g) {
    if (g == gcur) return;
    gcur = g;
    typename T::Raw rr = raw[0];
    const int j = g - g_first;
    if (j == 1) rr = raw[1];
    else if (j == 2) rr = raw[2];
    else if (j > 2 && lane_ok) T::load_raw(rr, p.L, g, n0, lane);         // long k-ranges: fetch on demand
    T::make_gc(gc, rr, p.L.zero_bias);
  };
  const char* xw = xs - (size_t)k_cta0 * 2 * T::XDUP;                     // so that x is indexed by absolute k
  for (int base = 0; base < nld; base += T::MAXLD) {                       // one iteration unless the k-range is long
    if (base > 0) {
#pragma unroll
      for (int i = 0; i < T::MAXLD; ++i)
        if (base + i < nld && lane_ok) T::load(w[i], p.L, r_begin + (base + i) * T::RPL + sub, n0, lane);
    }
#pragma unroll
    for (int i = 0; i < T::MAXLD; ++i) {
      if (base + i < nld) {
        const int r = r_begin + (base + i) * T::RPL + sub;
        if (i % T::FLUSH_ROWS == 0) set_group(group_of((r - sub) * T::KROW));
        T::template row<MB>(w[i], xw, p.x_stride, r * T::KROW, gc, acc);
        if (i % T::FLUSH_ROWS == T::FLUSH_ROWS - 1) T::template flush<MB>(tot, acc, gc);
      }
    }
  }

  // ---- 4. reduce: warps -> CTA (shared), CTAs of the cluster -> rank 0 (distributed shared) ----
#pragma unroll
  for (int m = 0; m < MB; ++m)
#pragma unroll
    for (int i = 0; i < T::COLS; ++i) red[((size_t)(warp * T::RPL + sub) * MB + m) * T::NT + T::col_of(lane, i)] = tot[m][i];
  __syncthreads();
  constexpr int NV = (T::NT * MB + kFThreads - 1) / kFThreads;
  float v[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const int idx = tid + q * kFThreads;                 // idx = m * NT + n
    float sum = 0.f;
    if (idx < T::NT * MB) {
#pragma unroll
      for (int wq = 0; wq < kFWarps * T::RPL; ++wq) sum += red[(size_t)wq * MB * T::NT + idx];
    }
    v[q] = sum;
  }
  if (cs > 1) {
    if (rank != 0) {
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const int idx = tid + q * kFThreads;
        if (idx < T::NT * MB) f_st_cluster(rbuf + (size_t)(rank - 1) * T::NT * MB + idx, 0u, v[q]);
      }
    }
    f_cluster_sync();
    if (rank != 0) return;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int idx = tid + q * kFThreads;
      if (idx < T::NT * MB)
        for (int c = 0; c < cs - 1; ++c) v[q] += rbuf[(size_t)c * T::NT * MB + idx];
    }
  }
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const int idx = tid + q * kFThreads;
    const int m = idx / T::NT, n = idx % T::NT;
    if (idx < T::NT * MB && m < p.M && n < ncols) {
      float o = v[q];
      if (p.L.bias) o += __half2float(__ldg(p.L.bias + n0 + n));
      const __half h = __float2half_rn(o);
      for (int c = 0; c < p.out.n; ++c) p.out.y[c][(size_t)m * p.ldy + p.n_offset + n0 + n] = h;
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct FmaPlan {
  int kind, NT, KROW, MAXLD, FLUSH, XDUP, RPL;
  int n_tiles, cluster, rows_total, group_shift, x_stride;
  int off_x, off_red, off_rbuf, smem_bytes;
};

static int g_fma_max_m = 2;
void gemv_fma_set_max_m(int m) { g_fma_max_m = m; }

static bool fma_plan(const LayerView& L, int M, FmaPlan& pl) {
  pl.kind = 0;
  if (M < 1 || M > 2 || M > g_fma_max_m || L.g_idx != nullptr || L.x_perm != nullptr || L.bits != 4) return false;
  if (L.layout == B200Q_LAYOUT_GPTQ) { pl.kind = 1; pl.NT = FGptq4::NT; pl.KROW = 8; pl.MAXLD = FGptq4::MAXLD; pl.FLUSH = 2; pl.XDUP = 1; pl.RPL = 1; }
  else if (L.layout == B200Q_LAYOUT_AWQ_GEMM) { pl.kind = 2; pl.NT = FAwq4::NT; pl.KROW = 1; pl.MAXLD = FAwq4::MAXLD; pl.FLUSH = 8; pl.XDUP = 2; pl.RPL = 2; }
  else return false;
  const int flush_k = pl.FLUSH * pl.RPL * pl.KROW;                  // a flush block must lie inside one group
  if (L.group % flush_k != 0 || L.K % L.group != 0 || L.N % 32 != 0) { pl.kind = 0; return false; }
  pl.group_shift = -1;
  if ((L.group & (L.group - 1)) == 0) { int s = 0; while ((1 << s) < L.group) ++s; pl.group_shift = s; }
  pl.n_tiles = (L.N + pl.NT - 1) / pl.NT;
  pl.rows_total = L.K / pl.KROW;
  // k-split: enough units that one prefetch round (MAXLD rows per warp) covers the matrix, and >= ~1 CTA per SM
  int cs = 1;
  while (cs < 8 && (pl.rows_total + cs * kFWarps - 1) / (cs * kFWarps) > pl.MAXLD * pl.RPL) cs *= 2;
  while (cs < 8 && pl.n_tiles * cs < 120) cs *= 2;
  pl.cluster = cs;
  const int kslice = ((pl.rows_total + cs - 1) / cs + pl.FLUSH * pl.RPL * kFWarps) * pl.KROW;
  pl.x_stride = kslice * 2 * pl.XDUP + 16;
  int off = 0;
  pl.off_x = off; off += M * pl.x_stride;
  off = (off + 15) & ~15;
  pl.off_red = off; off += kFWarps * pl.RPL * M * pl.NT * 4;
  pl.off_rbuf = off; off += (cs - 1) * M * pl.NT * 4;
  pl.smem_bytes = off;
  if (pl.smem_bytes > 96 * 1024) { pl.kind = 0; return false; }
  return true;
}

bool gemv_fma_supported(const LayerView& L, int M, const __half* x, int64_t ldx) {
  FmaPlan pl;
  if (!fma_plan(L, M, pl)) return false;
  if (((uintptr_t)L.qw & 15) != 0 || ((uintptr_t)L.s & 15) != 0 || (L.N % 8) != 0) return false;
  (void)x; (void)ldx;
  return true;
}

template <class T, int MB>
static cudaError_t fma_launch_k(const FmaParams& p, const FmaPlan& pl, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemv_fma_kernel<T, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    // all of the SM's unified L1/shared array as shared memory, so that consecutive layers' CTAs can be co-resident
    if (decode_carveout_max()) cudaFuncSetAttribute(gemv_fma_kernel<T, MB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    attr_done[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.n_tiles * pl.cluster);
  cfg.blockDim = dim3(kFThreads);
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
  return cudaLaunchKernelEx(&cfg, gemv_fma_kernel<T, MB>, p);
}

cudaError_t launch_gemv_fma(const LinearArgs& a, const PeerOut* peers) {
  FmaPlan pl;
  if (!fma_plan(a.L, a.M, pl)) return cudaErrorInvalidValue;
  FmaParams p;
  p.L = a.L; p.x = a.x; p.ldx = a.ldx; p.M = a.M;
  if (peers) p.out = *peers; else { p.out.n = 1; p.out.y[0] = a.y; }
  p.ldy = a.ldy; p.n_offset = a.n_offset;
  p.n_tiles = pl.n_tiles; p.cluster = pl.cluster; p.rows_total = pl.rows_total; p.group_shift = pl.group_shift;
  p.x_stride = pl.x_stride; p.off_x = pl.off_x; p.off_red = pl.off_red; p.off_rbuf = pl.off_rbuf;
  if (pl.kind == 1) return a.M == 1 ? fma_launch_k<FGptq4, 1>(p, pl, a.stream) : fma_launch_k<FGptq4, 2>(p, pl, a.stream);
  return a.M == 1 ? fma_launch_k<FAwq4, 1>(p, pl, a.stream) : fma_launch_k<FAwq4, 2>(p, pl, a.stream);
}

}  // namespace b200q
