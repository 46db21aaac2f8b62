// Decode CHAIN kernel (M <= 2): a recorded sequence of b200q_linear_group calls executed by ONE persistent launch.
//
// Why: at batch 1 a Llama-sized QuantLinear is 1.3-7 us of HBM streaming, and the round-1 launch-per-layer kernels lose
// as much again per launch to the serial chain "previous layer stored -> kernel boundary -> x loaded and split into
// digits -> first weights consumed -> reduced -> stored" during which HBM idles (profiles/r1_decode_timeline_bench.txt),
// plus 2-3 CTAs per SM of uneven work.  Here one CTA per SM lives for the whole chain:
//
//   * ONE producer thread per SM streams the packed weights of every layer of the chain, in order, through ONE
//     shared-memory ring with TMA (3-D boxes: 64 columns x 256 k = 8 KB, 128-byte swizzle; the slab's scales and packed
//     zeros ride behind it as two small boxes).  The ring does not drain at layer boundaries: while the consumers wait
//     for a layer's activations, the next layer's weights keep arriving (~150 KB per SM, ~3.5 us of HBM time in flight);
//   * the work of a group (sibling layers sharing x) is cut into slabs and every CTA takes an equal CONTIGUOUS range
//     of (tile, k) slabs -- balanced to one slab (1-2 %) for any N, no cluster, no tail wave;
//   * 16 consumer warps run the integer tensor path of gemv_imma.cu (IMMA.16832 on nibbles, activations as three
//     base-128 digits, exact int32 accumulation per <= 128-k part, fp32 scale/zero fix-up per group);
//   * split-K partial sums go to a small fp32 scratch (L2 resident); layers are separated by ONE grid-wide barrier
//     (red.release.gpu + ld.acquire.gpu poll, ~1 us); the NEXT layer's x-load stage sums the partials, adds the bias and
//     rounds to fp16 itself (same fixed order everywhere -> deterministic), so no reduction pass sits on the critical path.
//     A finalizer warp writes the fp16 outputs y (what QuantLinear.forward returns) off the critical path.
//
// Replaces a run of ort_ops.gemv / gemm_forward_cuda calls (dq_gemv.cu:40-177, gemm_cuda_gen.cu:1102-1161) at M <= 2.
#include <cuda.h>

#include <cstring>
#include <vector>

#include "gemv_stream.cuh"

namespace b200q {

static constexpr int kChWarps = 16;                       // consumer warps
static constexpr int kChThreads = 32 * (kChWarps + 4);    // + warpgroup 0: producer warp, finalizer warp, two idle warps
static constexpr int kSlabK = 256, kTileN = 64, kSlabBytes = 8192;
static constexpr uint32_t kChMagic = 0xB2C4A118u;
static constexpr uint32_t NIBM = 0x0f0f0f0fu;
static constexpr int kCtrWord = 512, kExitWord = 513, kErrWord = 514;   // u32 words inside the 4 KB counter region
static constexpr size_t kChWsHdr = 256;                   // after the counter region: word 0 = launch epoch (never reset)

struct alignas(64) ChLayer {
  CUtensorMap wmap;
  const __half* s;
  const void* qz;
  const __half* bias;
  __half* y;
  int64_t ldy;
  int32_t N, tile0, ntiles, pad_;
};
struct alignas(64) ChGroup {
  ChLayer layer[kMaxGroupLayers];
  int32_t n_layers, K, tiles, kc, U, group, pk, pps, zfp16, zero_bias;
  int32_t ncta;                                // CTAs that share this group's slabs: min(grid, U)
  int32_t region, smax, ncols;                 // this group's partial sums: P[region][smax][M][ncols] x {fp32, tag}
  int32_t xmode;                               // 0: plain fp16 x; 1: partial sums of the previous group
  const __half* x;
  int64_t ldx;
  const int32_t* xperm;
  int32_t src_smax, src_ncols, src_pcol0, src_region;
  const __half* src_bias;                      // bias of the source columns (already offset), or NULL
  uint32_t src_tab_off, tab_off;               // byte offsets (from the plan base) of u32 [tiles] tables: c0 | cnt << 16
};
// the scalars of a step the consumers and the finalizer need, copied into shared memory at kernel start (a global load
// costs ~1 us under the weight stream; the step descriptors are on every step's critical path)
struct alignas(16) ChGS {
  const __half* s[kMaxGroupLayers];
  const void* qz[kMaxGroupLayers];
  const __half* bias[kMaxGroupLayers];
  __half* y[kMaxGroupLayers];
  int64_t ldy[kMaxGroupLayers];
  const __half* x;
  const int32_t* xperm;
  const __half* src_bias;
  int64_t ldx;
  int32_t N[kMaxGroupLayers], tile0[kMaxGroupLayers];
  int32_t n_layers, K, tiles, kc, U, group, pk, pps, zfp16, zero_bias, ncta, region, smax, ncols, xmode;
  int32_t src_smax, src_ncols, src_pcol0, src_region, tab_off, gshift, pad_[1];
};
static_assert(sizeof(ChGS) % 16 == 0, "ChGS is copied as uint4");
struct ChHeader {
  uint32_t magic, n_groups, M, n_cta, slots, max_tiles, smem_bytes, total_bytes;
  uint32_t off_bars, off_digits, off_parts, off_red, off_tab, off_zpad, off_ring, use_barrier;
  uint32_t groups_off, gs_off, off_gs, n_cached, window, pad_;
  uint64_t region_elems;                       // 8-byte elements per partial-sum region
  uint64_t ws_bytes;
};
struct ChParams {
  const char* plan;       // device copy of the plan blob
  char* ws;
  ChHeader h;
  unsigned long long* dbg;
};

__device__ __forceinline__ void tma_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// every wait in this kernel is bounded (4 s): a wedged chain traps instead of hanging the GPU
__device__ __noinline__ void ch_fail(char* ws, uint32_t code) {
  reinterpret_cast<volatile uint32_t*>(ws)[kErrWord] = code;
  __threadfence_system();
  __trap();
}
struct ChSpin {
  unsigned long long t0 = 0;
  uint32_t spins = 0;
  __device__ __forceinline__ void tick(char* ws, uint32_t code) {
    if ((++spins & 1023u) == 0) {
      const unsigned long long now = st_gtime();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) ch_fail(ws, code);
    }
  }
};
__device__ __forceinline__ void mbar_wait_b(uint32_t bar, uint32_t parity, char* ws, uint32_t code) {
  if (mbar_try_a(bar, parity)) return;
  ChSpin sp;
  while (!mbar_try_a(bar, parity)) sp.tick(ws, code);
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// a partial sum travels as one 64-bit word {fp32 bits, tag}: naturally aligned 8-byte accesses are single-copy atomic,
// so a reader that sees the step's tag sees the value -- no fence, no flag, the load is the hand-off
__device__ __forceinline__ unsigned long long ld_tagged(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_tagged(unsigned long long* p, float v, uint32_t tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned short ldcg_u16(const void* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kChWarps * 32) : "memory"); }

__device__ __forceinline__ void imma_acc(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// sum over the contributing CTAs' slots (in slot order) of NE consecutive tagged partial sums; spins until every word
// carries `tag`.  All loads are issued before the first check.
template <int NE>
__device__ __forceinline__ void ch_gather(const unsigned long long* base, size_t slot_stride, int cnt, uint32_t tag, float (&o)[NE],
                                          char* ws, uint32_t code) {
#pragma unroll
  for (int e = 0; e < NE; ++e) o[e] = 0.f;
  for (int s0 = 0; s0 < cnt; s0 += 4) {
    unsigned long long r[4][NE];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (s0 + i < cnt) {
#pragma unroll
        for (int e = 0; e < NE; ++e) r[i][e] = ld_tagged(base + (size_t)(s0 + i) * slot_stride + e);
      }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (s0 + i < cnt) {
        ChSpin sp;
        for (;;) {
          bool ok = true;
#pragma unroll
          for (int e = 0; e < NE; ++e) ok = ok && ((uint32_t)(r[i][e] >> 32) == tag);
          if (ok) break;
          sp.tick(ws, code);
#pragma unroll
          for (int e = 0; e < NE; ++e) r[i][e] = ld_tagged(base + (size_t)(s0 + i) * slot_stride + e);
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) o[e] += __uint_as_float((uint32_t)r[i][e]);
      }
  }
}

// diagnostic: 16 x u64 per (step, CTA): consumer warp 0 phases 0..5, finalizer warp 6..7, producer 8..10, consumer warp 0
// cycle counters 11..13 (waiting for weights, units, unit loop)
#define CH_STAMP(g, i) do { if (p.dbg && lane == 0) p.dbg[((size_t)(g) * gridDim.x + blockIdx.x) * 16 + (i)] = st_gtime(); } while (0)

#define CH_PROLOGUE()                                                                                                   \
  /* step scalars -> shared memory (finalizer warp + consumers copy, then meet on named barrier 2) */                  \
  for (int i = (warp == 1) ? lane : tid - 96; i < ncached * (int)(sizeof(ChGS) / 16); i += kChWarps * 32 + 32)        \
    reinterpret_cast<uint4*>(smem + H.off_gs)[i] = __ldg(reinterpret_cast<const uint4*>(gs_global) + i);               \
  pdl_wait(); /* upstream results (x, y buffers, workspace) are complete */                                            \
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(p.ws + kCounterBytes);                            \
  const uint32_t tag0 = epoch * (uint32_t)(NG + 1) + 1u; /* tag of step g: tag0 + g (never 0) */                       \
  asm volatile("bar.sync 2, %0;" ::"n"(kChWarps * 32 + 32) : "memory");

template <int MTOK>
__global__ void __launch_bounds__(kChThreads, 1) decode_chain_kernel(const __grid_constant__ ChParams p) {
  extern __shared__ __align__(1024) char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const ChHeader& H = p.h;
  const int NS = (int)H.slots;
  const uint32_t bars = smem_u32(smem + H.off_bars);
  const uint32_t bar_full = bars, bar_empty = bars + 8u * 32u, bar_xready = bars + 8u * 64u, bar_cdone = bars + 8u * 65u;
  volatile int* fin_count = reinterpret_cast<volatile int*>(smem + H.off_bars + 8 * 66);   // steps whose y this CTA has written
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x, ncta = gridDim.x;
  const int NG = (int)H.n_groups;
  const bool use_barrier = H.use_barrier != 0;
  const ChGroup* groups = reinterpret_cast<const ChGroup*>(p.plan + H.groups_off);
  const ChGS* gs_global = reinterpret_cast<const ChGS*>(p.plan + H.gs_off);
  const ChGS* gs_cache = reinterpret_cast<const ChGS*>(smem + H.off_gs);
  const int ncached = (int)H.n_cached;
  uint32_t* ctr = reinterpret_cast<uint32_t*>(p.ws);
  unsigned long long* Pbase = reinterpret_cast<unsigned long long*>(p.ws + kCounterBytes + kChWsHdr);

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(smem + H.off_bars) + i, 1);
      mbar_init(reinterpret_cast<uint64_t*>(smem + H.off_bars) + 32 + i, 4);
    }
    mbar_init(reinterpret_cast<uint64_t*>(smem + H.off_bars) + 64, 1);
    mbar_init(reinterpret_cast<uint64_t*>(smem + H.off_bars) + 65, kChWarps);
    *fin_count = 0;
    fence_mbar_init();
  }
  if (tid < 16) reinterpret_cast<uint32_t*>(smem + H.off_zpad)[tid] = 0u;
  __syncthreads();
  pdl_launch_dependents();

  // register budget: the control warpgroup gives registers back, the four consumer warpgroups take them
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 2 || warp == 3) return;
    // =========================================== producer ===========================================
    // packed weights are constants: the stream starts before the upstream kernel has finished (no griddepcontrol.wait).
    // At most `window` slabs are in flight per SM: more only lengthens every queue between the SM and HBM (a demand
    // load behind 128 KB of bulk reads waits ~3 us) without adding bandwidth; landed slabs may fill the whole ring.
    if (warp == 0) {
      if (lane != 0) return;
      const uint32_t ring = smem_u32(smem + H.off_ring);
      const int W = (int)H.window;
      int slot = 0, wslot = 0, issued = 0;
      uint32_t round = 0, wround = 0;
      for (int g = 0; g < NG; ++g) {
        const ChGroup* G = groups + g;
        const int U = G->U, KC = G->kc, nl = G->n_layers, ng = G->ncta;
        const int a = c < ng ? (int)((long long)c * U / ng) : 0, b = c < ng ? (int)((long long)(c + 1) * U / ng) : 0;
        if (a >= b) continue;
        CH_STAMP(g, 8);
        unsigned long long stall = 0;
        int tile = a / KC, kk = a - tile * KC, j = 0;
        while (j + 1 < nl && tile >= G->layer[j + 1].tile0) ++j;
        int tile0 = G->layer[j].tile0, tile_end = tile0 + G->layer[j].ntiles;
        const void* wmap = &G->layer[j].wmap;
        asm volatile("prefetch.tensormap [%0];" ::"l"(wmap) : "memory");
        for (int i = a; i < b; ++i) {
          const unsigned long long t0 = p.dbg ? st_gtime() : 0ull;
          if (round > 0) mbar_wait_b(bar_empty + 8u * slot, (round - 1) & 1u, p.ws, 0x100u + g);
          if (issued >= W) {                                               // the slab issued W slabs ago has landed
            mbar_wait_b(bar_full + 8u * wslot, wround & 1u, p.ws, 0x180u + g);
            if (++wslot == NS) { wslot = 0; ++wround; }
          }
          if (p.dbg) stall += st_gtime() - t0;
          const uint32_t fb = bar_full + 8u * slot;
          mbar_expect_tx_a(fb, kSlabBytes);
          tma_3d(ring + (uint32_t)slot * kSlabBytes, wmap, 0, 2 * (tile - tile0), 32 * kk, fb);
          ++issued;
          if (++slot == NS) { slot = 0; ++round; }
          if (++kk == KC) {
            kk = 0;
            if (++tile == tile_end && j + 1 < nl) {
              ++j;
              tile0 = tile_end; tile_end = tile0 + G->layer[j].ntiles;
              wmap = &G->layer[j].wmap;
              asm volatile("prefetch.tensormap [%0];" ::"l"(wmap) : "memory");
            }
          }
        }
        CH_STAMP(g, 9);
        if (p.dbg) p.dbg[((size_t)g * gridDim.x + blockIdx.x) * 16 + 10] = stall;
      }
      return;
    }
  // ============================== finalizer warp (and, in barrier mode, the grid barrier) ======================
  {
    CH_PROLOGUE();
    for (int g = 0; g < NG; ++g) {
      const ChGS* G = g < ncached ? gs_cache + g : gs_global + g;
      if (use_barrier) {
        mbar_wait_b(bar_cdone, (uint32_t)g & 1u, p.ws, 0x200u + g);        // this CTA's partial sums of step g are stored
        if (lane == 0) {
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr + kCtrWord) : "memory");
          const uint32_t target = (uint32_t)(g + 1) * (uint32_t)ncta;
          ChSpin sp;
          while (ld_acquire_gpu(ctr + kCtrWord) < target) sp.tick(p.ws, 0x300u + g);
          if (g + 1 < NG) mbar_arrive_n(bar_xready, 1);
        }
        __syncwarp();
      }
      CH_STAMP(g, 6);
      // y = fp16(sum of partials + bias) for the tiles whose FIRST slab this CTA owns (2 columns per lane)
      const int U = G->U, KC = G->kc, ng = G->ncta, smax = G->smax, ncols = G->ncols, nl = G->n_layers;
      const int a = c < ng ? (int)((long long)c * U / ng) : 0, b = c < ng ? (int)((long long)(c + 1) * U / ng) : 0;
      const unsigned long long* P = Pbase + (size_t)G->region * H.region_elems;
      const uint32_t tag = tag0 + (uint32_t)g;
      for (int tile = (a + KC - 1) / KC; tile * KC < b; ++tile) {
        int j = 0;
        while (j + 1 < nl && tile >= G->tile0[j + 1]) ++j;
        const int col = (tile - G->tile0[j]) * kTileN + 2 * lane;
#pragma unroll
        for (int m = 0; m < MTOK; ++m) {
          float v[2];
          ch_gather<2>(P + (size_t)m * ncols + (size_t)tile * kTileN + 2 * lane, (size_t)MTOK * ncols, smax, tag, v, p.ws, 0x600u + g);
          if (col < G->N[j]) {
            const __half* bias = G->bias[j];
            if (bias) { v[0] += __half2float(__ldg(bias + col)); v[1] += __half2float(__ldg(bias + col + 1)); }
            *reinterpret_cast<__half2*>(G->y[j] + (size_t)m * G->ldy[j] + col) = __floats2half2_rn(v[0], v[1]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) { __threadfence_block(); *fin_count = g + 1; }
      CH_STAMP(g, 7);
    }
    if (lane == 0) {                                                       // the last CTA out advances the epoch, re-zeroes the counters
      __threadfence();
      const uint32_t old = atomicAdd(ctr + kExitWord, 1u);
      if (old == (uint32_t)ncta - 1u) {
        *reinterpret_cast<volatile uint32_t*>(p.ws + kCounterBytes) = epoch + 1u;
        ctr[kCtrWord] = 0u;
        ctr[kExitWord] = 0u;
        __threadfence();
      }
    }
    return;
  }
  }

  // =========================================== consumers ============================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  CH_PROLOGUE();
  const int w = warp - 4, ctid = tid - 128;
  const int g8 = lane >> 2, t = lane & 3;
  char* xdig = smem + H.off_digits;
  float2* parts = reinterpret_cast<float2*>(smem + H.off_parts);
  float* red = reinterpret_cast<float*>(smem + H.off_red);
  float2* tabw = reinterpret_cast<float2*>(smem + H.off_tab) + w * kTileN;
  const int MT = (int)H.max_tiles;
  float* redw = red + (size_t)w * MT * kTileN * MTOK;
  const uint32_t ring = smem_u32(smem + H.off_ring);
  // weights: sub-step ss of a slab, half h: ring + slot * 8192 + ss * 1024 + off_h
  const uint32_t offh0 = (uint32_t)((2 * t) * 128 + ((g8 ^ ((2 * t) & 7)) << 4));
  const uint32_t offh1 = (uint32_t)((2 * t + 1) * 128 + ((g8 ^ ((2 * t + 1) & 7)) << 4));
  // digits: B column g8 -> token g8 >> 2, digit g8 & 3 (3 = unused -> zero pad)
  const bool xreal = (g8 & 3) < 3 && (g8 >> 2) < MTOK;
  const uint32_t xlane = xreal ? smem_u32(xdig) + (uint32_t)(((g8 >> 2) * 3 + (g8 & 3)) * 32 + t * 8) : smem_u32(smem + H.off_zpad);
  const uint32_t xsub = xreal ? (uint32_t)(96 * MTOK) : 0u;
  const int mytok = t >> 1;
  const bool fx_on = mytok < MTOK;
  const float dscale = (t & 1) ? (1.0f / 128.0f) : 128.0f, zflag = (t & 1) ? 0.f : 1.f;
  const int nsmask = NS - 1;                                               // NS is 8 or 16

  uint32_t q0 = 0;                                                         // CTA-local slab sequence number at step start
  for (int g = 0; g < NG; ++g) {
    const ChGS* G = g < ncached ? gs_cache + g : gs_global + g;
    const int K = G->K, U = G->U, KC = G->kc, PK = G->pk, PPS = G->pps, ng = G->ncta;
    const int zf = G->zfp16, zbias = G->zero_bias, gshift = G->gshift, nl = G->n_layers;
    const int a = c < ng ? (int)((long long)c * U / ng) : 0, b = c < ng ? (int)((long long)(c + 1) * U / ng) : 0;
    const int nparts = K / PK;
    const int pshift = (PPS == 4) ? 2 : 1;
    const int dslab = kChWarps >> pshift;                                  // slabs between two consecutive units of a warp
    const int tile_first = (a < b) ? a / KC : 0;
    const int ntl = (a < b) ? ((b - 1) / KC - tile_first + 1) : 0;
    if (w == 0) CH_STAMP(g, 0);
    // table entry (first contributor, contributors) of the tile this thread will store a partial sum of
    uint32_t my_tab = 0;
    if (ctid < ntl * kTileN * MTOK)
      my_tab = __ldg(reinterpret_cast<const uint32_t*>(p.plan + G->tab_off) + tile_first + ctid / (kTileN * MTOK));

    // ---- this warp's first unit; its (scale, zero) words are fetched now, one unit ahead of their use, from then on ----
    const uint32_t ubase = q0 << pshift, uend = (q0 + (uint32_t)(b - a)) << pshift;
    uint32_t u = ubase + (((uint32_t)w + kChWarps - (ubase & (kChWarps - 1))) & (kChWarps - 1));
    int tile = 0, kk = 0, lj = 0, ltile0 = 0, ltile_end = 0, lN = 0;
    const __half* ls = nullptr;
    const char* lqz = nullptr;
    uint32_t pf_s = 0, pf_z = 0;
    auto load_layer = [&]() {
      while (lj + 1 < nl && tile >= G->tile0[lj + 1]) ++lj;
      ltile0 = G->tile0[lj]; ltile_end = (lj + 1 < nl) ? G->tile0[lj + 1] : G->tiles; lN = G->N[lj];
      ls = G->s[lj]; lqz = reinterpret_cast<const char*>(G->qz[lj]);
    };
    auto prefetch_sz = [&](int tile_, int kk_, int pp_) {                  // scale / zero words of columns 2 lane, 2 lane + 1 of the unit
      const int col = (tile_ - ltile0) * kTileN + 2 * lane;
      const int k = kk_ * kSlabK + pp_ * PK;
      const int grow = gshift >= 0 ? (k >> gshift) : k / G->group;
      pf_s = 0; pf_z = 0;
      if (col < lN) {
        pf_s = __ldg(reinterpret_cast<const uint32_t*>(ls + (size_t)grow * lN + col));
        pf_z = zf ? __ldg(reinterpret_cast<const uint32_t*>(lqz + ((size_t)grow * lN + col) * 2))
                  : __ldg(reinterpret_cast<const uint32_t*>(lqz + ((size_t)grow * (lN >> 3) + (col >> 3)) * 4));
      }
    };
    if (u < uend) {
      const int i0 = a + (int)((u >> pshift) - q0);
      tile = i0 / KC; kk = i0 - tile * KC;
      load_layer();
      prefetch_sz(tile, kk, (int)(u & (uint32_t)(PPS - 1)));
    }

    if (use_barrier && g > 0) mbar_wait_b(bar_xready, (uint32_t)(g - 1) & 1u, p.ws, 0x400u + g);
    if (w == 0) CH_STAMP(g, 1);

    // ---- x -> three base-128 digits per element, power-of-two scale per part (PK k) -----------------------------
    {
      const unsigned long long* Psrc = Pbase + (size_t)G->src_region * H.region_elems;
      const uint32_t xtag = tag0 + (uint32_t)(g - 1);
      const int EL = PK >> 5;                                              // elements per lane: 4 (PK = 128) or 2 (PK = 64)
      const int xmode = G->xmode, ssmax = G->src_smax, sncols = G->src_ncols, spc0 = G->src_pcol0;
      const int* xperm = G->xperm;
      const __half* sbias = G->src_bias;
      for (int pr = w; pr < nparts; pr += kChWarps) {
        const int k0 = pr * PK + EL * lane;
#pragma unroll
        for (int m = 0; m < MTOK; ++m) {
          float xv[4] = {0.f, 0.f, 0.f, 0.f};
          if (xmode == 1) {
            if (!xperm && EL == 4) {
              ch_gather<4>(Psrc + (size_t)m * sncols + spc0 + k0, (size_t)MTOK * sncols, ssmax, xtag, xv, p.ws, 0x700u + g);
              if (sbias) {
                const uint2 bb = __ldg(reinterpret_cast<const uint2*>(sbias + k0));
                const __half2 b01 = *reinterpret_cast<const __half2*>(&bb.x), b23 = *reinterpret_cast<const __half2*>(&bb.y);
                xv[0] += __low2float(b01); xv[1] += __high2float(b01); xv[2] += __low2float(b23); xv[3] += __high2float(b23);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (e < EL) {
                  const int kx = xperm ? __ldg(xperm + k0 + e) : k0 + e;
                  float v1[1];
                  ch_gather<1>(Psrc + (size_t)m * sncols + spc0 + kx, (size_t)MTOK * sncols, ssmax, xtag, v1, p.ws, 0x700u + g);
                  xv[e] = v1[0];
                  if (sbias) xv[e] += __half2float(__ldg(sbias + kx));
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) xv[e] = __half2float(__float2half_rn(xv[e]));     // y is fp16 (QuantLinear.forward's output)
          } else {
            const __half* xr = G->x + (size_t)m * G->ldx;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < EL) {
                const int kx = xperm ? __ldg(xperm + k0 + e) : k0 + e;
                xv[e] = __half2float(__ushort_as_half(ldcg_u16(xr + kx)));
              }
          }
          // non-finite activations poison the part (the fp16 kernels propagate NaN / Inf through their FMAs)
          uint32_t bad = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) bad |= ((__float_as_uint(xv[e]) & 0x7f800000u) == 0x7f800000u) ? 1u : 0u;
          float mx = fmaxf(fmaxf(fabsf(xv[0]), fabsf(xv[1])), fmaxf(fabsf(xv[2]), fabsf(xv[3])));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          bad = __any_sync(0xffffffffu, bad);
          if (bad) { mx = 0.f; xv[0] = xv[1] = xv[2] = xv[3] = 0.f; }
          const int ex = (int)((__float_as_uint(mx) >> 23) & 0xffu);
          const int E = (mx > 0.f) ? (19 + 127 - ex) : 0;
          const float sc = __uint_as_float((uint32_t)(E + 127) << 23), isc = __uint_as_float((uint32_t)(127 - E) << 23);
          int tsum = 0;
          uint32_t dig[3] = {0u, 0u, 0u};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int tv = __float2int_rn(xv[e] * sc);                     // |tv| <= 2^20
            tsum += tv;
            const int d2 = ((tv + 64) & 127) - 64;
            const int t1 = (tv - d2) >> 7;
            const int d1 = ((t1 + 64) & 127) - 64;
            const int d0 = (t1 - d1) >> 7;
            dig[0] |= (uint32_t)(d0 & 0xff) << (8 * e);
            dig[1] |= (uint32_t)(d1 & 0xff) << (8 * e);
            dig[2] |= (uint32_t)(d2 & 0xff) << (8 * e);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
          if (lane == 0) parts[pr * MTOK + m] = bad ? make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000))
                                                    : make_float2(isc, (float)tsum * isc);
          // element e of this lane sits at k' = (EL lane + e) % 32 of 32-k sub-step sub: prow k' / 8, kk = k' % 8;
          // even kk -> byte kk / 2 of the first word, odd kk -> byte kk / 2 of the second word
          const int sub = (pr * PK + EL * lane) >> 5, kq = (EL * lane) & 31;
          char* dbase = xdig + (size_t)sub * (96 * MTOK) + (size_t)(3 * m) * 32 + (kq >> 3) * 8 + ((kq & 7) >> 1);
          if (EL == 4) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const uint32_t v = dig[d];
              *reinterpret_cast<uint16_t*>(dbase + d * 32) = (uint16_t)((v & 0xffu) | ((v >> 8) & 0xff00u));             // e = 0, 2
              *reinterpret_cast<uint16_t*>(dbase + d * 32 + 4) = (uint16_t)(((v >> 8) & 0xffu) | ((v >> 16) & 0xff00u)); // e = 1, 3
            }
          } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              dbase[d * 32] = (char)(dig[d] & 0xffu);
              dbase[d * 32 + 4] = (char)((dig[d] >> 8) & 0xffu);
            }
          }
        }
      }
    }
    consumer_bar();
    if (w == 0) CH_STAMP(g, 2);

    // ---- this warp's units: (slab, part) pairs numbered u = q * PPS + part, unit u -> warp u % 16 ------------------
    int acc[2][2][4];
    float tot[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) tot[i][0] = tot[i][1] = 0.f;
    uint32_t touched = 0;
    int cur_tl = -1;                                                       // local tile (tile - first tile of this CTA) of `tot`
    long long cyc_wait = 0, cyc_loop = 0;
    int n_units = 0;
    auto flush = [&]() {
      if (cur_tl >= 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a0 = tot[q][0] + __shfl_xor_sync(0xffffffffu, tot[q][0], 1);
          const float a1 = tot[q][1] + __shfl_xor_sync(0xffffffffu, tot[q][1], 1);
          if (fx_on && !(t & 1)) {
            const int n = 32 * (q >> 1) + 4 * g8 + 2 * (q & 1);            // MMA row g8 -> column n, row g8 + 8 -> n + 1
            float* r = redw + ((size_t)cur_tl * kTileN + n) * MTOK + mytok;
            r[0] = a0;
            r[MTOK] = a1;
          }
          tot[q][0] = tot[q][1] = 0.f;
        }
        touched |= 1u << cur_tl;
      }
    };
    if (u < uend) {
      const long long tl0 = p.dbg ? clock64() : 0;
      const int steps_sub = PK >> 5;                                       // 32-k sub-steps per part (4 or 2)
      int slot = (int)((u >> pshift) & (uint32_t)nsmask);
      uint32_t par = ((u >> pshift) / (uint32_t)NS) & 1u;
      const int pp = (int)(u & (uint32_t)(PPS - 1));                       // the part index of a warp's units never changes (16 % PPS == 0)
      const int ss0 = pp * steps_sub;
      for (; u < uend; u += kChWarps) {
        const int tl = tile - tile_first;
        if (tl != cur_tl) { flush(); cur_tl = tl; }
        // (scale, zero) of the part's group for the 64 columns of the tile -> this warp's table (words fetched one unit ago)
        {
          const __half2 s2 = *reinterpret_cast<const __half2*>(&pf_s);
          float z0, z1;
          if (zf) {
            const __half2 z2 = *reinterpret_cast<const __half2*>(&pf_z);
            z0 = __low2float(z2); z1 = __high2float(z2);
          } else {
            const uint32_t zz = pf_z >> (8 * (lane & 3));
            z0 = (float)((zz + (uint32_t)zbias) & 15u);
            z1 = (float)(((zz >> 4) + (uint32_t)zbias) & 15u);
          }
          __syncwarp();                                                    // previous unit's table reads are done
          *reinterpret_cast<float4*>(tabw + 2 * lane) = make_float4(__low2float(s2), z0, __high2float(s2), z1);
        }
        const int kk_cur = kk;
        // next unit of this warp: advance (tile, kk), fetch its scale / zero words now
        {
          kk += dslab;
          while (kk >= KC) { kk -= KC; ++tile; }
          if (u + kChWarps < uend) {
            if (tile >= ltile_end) load_layer();
            prefetch_sz(tile, kk, pp);
          }
        }
        if (p.dbg) {
          const long long tw = clock64();
          mbar_wait_b(bar_full + 8u * slot, par, p.ws, 0x500u + g);
          cyc_wait += clock64() - tw;
          ++n_units;
        } else {
          mbar_wait_b(bar_full + 8u * slot, par, p.ws, 0x500u + g);
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int cp = 0; cp < 2; ++cp) acc[h][cp][0] = acc[h][cp][1] = acc[h][cp][2] = acc[h][cp][3] = 0;
        const uint32_t wb = ring + (uint32_t)slot * kSlabBytes + (uint32_t)ss0 * 1024u;
        uint32_t xp = xlane + (uint32_t)(kk_cur * 8 + ss0) * xsub;
#pragma unroll 2
        for (int s = 0; s < steps_sub; ++s) {
          const uint4 wa = lds128_s(wb + (uint32_t)s * 1024u + offh0), wc = lds128_s(wb + (uint32_t)s * 1024u + offh1);
          const uint2 xb = lds64_s(xp);
          xp += xsub;
          imma_acc(acc[0][0], wa.x & NIBM, wa.y & NIBM, (wa.x >> 4) & NIBM, (wa.y >> 4) & NIBM, xb.x, xb.y);
          imma_acc(acc[0][1], wa.z & NIBM, wa.w & NIBM, (wa.z >> 4) & NIBM, (wa.w >> 4) & NIBM, xb.x, xb.y);
          imma_acc(acc[1][0], wc.x & NIBM, wc.y & NIBM, (wc.x >> 4) & NIBM, (wc.y >> 4) & NIBM, xb.x, xb.y);
          imma_acc(acc[1][1], wc.z & NIBM, wc.w & NIBM, (wc.z >> 4) & NIBM, (wc.w >> 4) & NIBM, xb.x, xb.y);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_n(bar_empty + 8u * slot, (uint32_t)(4 >> (pshift - 1)) >> 1);   // 4 / PPS: this warp is done with the slot
        // fix-up: y += s * (2^-E * dscale * (d_a 2^7 + d_b) - zflag * z * sum(x)) for this lane's 8 (column, token) outputs
        if (fx_on) {
          const float2 pt = parts[(kk_cur * PPS + pp) * MTOK + mytok];
          const float xs = pt.x * dscale, sxz = pt.y * zflag;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 e01 = *reinterpret_cast<const float4*>(tabw + 32 * h + 4 * g8);
            const float4 e23 = *reinterpret_cast<const float4*>(tabw + 32 * h + 4 * g8 + 2);
            const float v00 = (float)((acc[h][0][0] << 7) + acc[h][0][1]), v01 = (float)((acc[h][0][2] << 7) + acc[h][0][3]);
            const float v10 = (float)((acc[h][1][0] << 7) + acc[h][1][1]), v11 = (float)((acc[h][1][2] << 7) + acc[h][1][3]);
            tot[2 * h][0] = fmaf(e01.x, fmaf(v00, xs, -e01.y * sxz), tot[2 * h][0]);
            tot[2 * h][1] = fmaf(e01.z, fmaf(v01, xs, -e01.w * sxz), tot[2 * h][1]);
            tot[2 * h + 1][0] = fmaf(e23.x, fmaf(v10, xs, -e23.y * sxz), tot[2 * h + 1][0]);
            tot[2 * h + 1][1] = fmaf(e23.z, fmaf(v11, xs, -e23.w * sxz), tot[2 * h + 1][1]);
          }
        }
        slot += dslab;
        if (slot >= NS) { slot -= NS; par ^= 1u; }
      }
      flush();
      if (p.dbg) cyc_loop = clock64() - tl0;
    }
    // tiles of this CTA the warp never touched contribute zeros
    for (int tl = 0; tl < ntl; ++tl)
      if (!((touched >> tl) & 1u))
        for (int idx = lane; idx < kTileN * MTOK; idx += 32) redw[(size_t)tl * kTileN * MTOK + idx] = 0.f;
    if (w == 0) CH_STAMP(g, 3);
    consumer_bar();
    if (w == 0) CH_STAMP(g, 4);
    // the partial sums of step g - 2 live where step g's go: this CTA's finalizer must have read what it needed of them
    if (!use_barrier && g >= 2) {
      ChSpin sp;
      while (*fin_count < g - 1) sp.tick(p.ws, 0x800u + g);
    }
    // ---- CTA-level reduction over the 16 warps (fixed order) -> this CTA's slot of the tile's partial sums; the LAST
    //      contributor of a tile also fills the tile's unused slots with tagged zeros, so readers never need the table ----
    {
      unsigned long long* P = Pbase + (size_t)G->region * H.region_elems;
      const uint32_t tag = tag0 + (uint32_t)g;
      const int smax = G->smax, ncols = G->ncols;
      for (int idx = ctid; idx < ntl * kTileN * MTOK; idx += kChWarps * 32) {
        float sum = 0.f;
#pragma unroll
        for (int wq = 0; wq < kChWarps; ++wq) sum += red[(size_t)wq * MT * kTileN * MTOK + idx];
        const int tl = idx / (kTileN * MTOK), rem = idx - tl * (kTileN * MTOK);
        const int n = rem / MTOK, m = rem - n * MTOK;
        const int tile_o = tile_first + tl;
        const uint32_t e = (idx == ctid) ? my_tab : __ldg(reinterpret_cast<const uint32_t*>(p.plan + G->tab_off) + tile_o);
        const int c0 = (int)(e & 0xffffu), cnt = (int)(e >> 16);
        unsigned long long* dst = P + (size_t)m * ncols + (size_t)tile_o * kTileN + n;
        st_tagged(dst + (size_t)(c - c0) * MTOK * ncols, sum, tag);
        if (c == c0 + cnt - 1)
          for (int sl = cnt; sl < smax; ++sl) st_tagged(dst + (size_t)sl * MTOK * ncols, 0.f, tag);
      }
    }
    if (w == 0) {
      CH_STAMP(g, 5);
      if (p.dbg && lane == 0) {
        unsigned long long* d = p.dbg + ((size_t)g * gridDim.x + blockIdx.x) * 16;
        d[11] = (unsigned long long)cyc_wait; d[12] = (unsigned long long)n_units; d[13] = (unsigned long long)cyc_loop;
      }
    }
    if (use_barrier) {
      __syncwarp();
      if (lane == 0) mbar_arrive_n(bar_cdone, 1);
    }
    q0 += (uint32_t)(b - a);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: plan + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn ch_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int g_ch_ctas = 0, g_ch_slots = 0, g_ch_barrier = 0, g_ch_window = 0;
static unsigned long long* g_ch_dbg = nullptr;
void decode_chain_set_debug(unsigned long long* buf) { g_ch_dbg = buf; }
void decode_chain_set_option(int which, int value) {
  if (which == 0) g_ch_ctas = value;
  else if (which == 1) g_ch_slots = value;
  else if (which == 2) g_ch_barrier = value;
  else if (which == 3) g_ch_window = value;
}

size_t decode_chain_plan_bytes(int n_groups, const int* tiles_per_group) {
  size_t b = sizeof(ChHeader);
  b = (b + 63) & ~(size_t)63;
  b += (size_t)n_groups * sizeof(ChGroup);
  b += ((size_t)n_groups * sizeof(ChGS) + 63) & ~(size_t)63;
  for (int i = 0; i < n_groups; ++i) b += ((size_t)tiles_per_group[i] * 4 + 63) & ~(size_t)63;
  return b;
}

static bool ch_layer_ok(const LayerView& L) {
  if (L.layout != B200Q_LAYOUT_GPTQ && L.layout != B200Q_LAYOUT_HQQ) return false;
  if (L.bits != 4 || L.g_idx) return false;
  if (!(L.group == 64 || L.group == 128 || L.group % kSlabK == 0)) return false;
  if (L.K % kSlabK != 0 || L.K % L.group != 0 || L.N % 32 != 0) return false;
  if (((uintptr_t)L.qw & 15) || ((uintptr_t)L.s & 15) || ((uintptr_t)L.qz & 15)) return false;
  return true;
}

// groups[i]: n layers sharing x (LinearArgs as in b200q_linear_group).  Returns 0, or a negative b200q status.
int decode_chain_plan(const LinearArgs* const* groups, const int* n_layers, int n_groups, int M, void* plan_out, size_t plan_cap,
                      size_t* plan_bytes, size_t* ws_bytes) {
  if (n_groups < 1 || M < 1 || M > 2) return B200Q_ERR_SHAPE;
  EncodeTiledFn enc = ch_encode();
  if (!enc) return B200Q_ERR_CUDA;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return B200Q_ERR_CUDA;
  const int ncta = (g_ch_ctas > 0 && g_ch_ctas <= sms) ? g_ch_ctas : sms;
  std::vector<int> tiles(n_groups);
  int kmax = 0, pkmin = 128;
  for (int g = 0; g < n_groups; ++g) {
    if (n_layers[g] < 1 || n_layers[g] > kMaxGroupLayers) return B200Q_ERR_SHAPE;
    const LayerView& A = groups[g][0].L;
    int T = 0;
    for (int j = 0; j < n_layers[g]; ++j) {
      const LinearArgs& a = groups[g][j];
      if (!ch_layer_ok(a.L)) return B200Q_ERR_UNSUPPORTED;
      if (a.L.K != A.K || a.L.group != A.group || a.L.layout != A.layout || a.L.zero_bias != A.zero_bias || a.L.x_perm != A.x_perm ||
          a.x != groups[g][0].x || a.ldx != groups[g][0].ldx || a.M != M)
        return B200Q_ERR_UNSUPPORTED;
      if (!a.y || a.ldy < a.L.N || ((uintptr_t)a.y & 3) || (a.ldy & 1)) return B200Q_ERR_ALIGNMENT;
      T += (a.L.N + kTileN - 1) / kTileN;
    }
    if (!groups[g][0].x || ((uintptr_t)groups[g][0].x & 1)) return B200Q_ERR_NULL;
    tiles[g] = T;
    if (A.K > kmax) kmax = A.K;
    if (A.group == 64) pkmin = 64;
  }
  const size_t need = decode_chain_plan_bytes(n_groups, tiles.data());
  if (plan_bytes) *plan_bytes = need;
  if (!plan_out) return B200Q_OK;                                         // size query
  if (plan_cap < need) return B200Q_ERR_WORKSPACE;
  char* blob = (char*)plan_out;
  memset(blob, 0, need);
  ChHeader* H = (ChHeader*)blob;
  size_t off = (sizeof(ChHeader) + 63) & ~(size_t)63;
  H->groups_off = (uint32_t)off;
  ChGroup* GG = (ChGroup*)(blob + off);
  off += (size_t)n_groups * sizeof(ChGroup);
  H->gs_off = (uint32_t)off;
  ChGS* GS = (ChGS*)(blob + off);
  off += ((size_t)n_groups * sizeof(ChGS) + 63) & ~(size_t)63;
  int max_tiles = 1, smax_all = 1, ncols_all = 0;
  bool small_group = false;
  for (int g = 0; g < n_groups; ++g) {
    ChGroup& G = GG[g];
    const LayerView& A = groups[g][0].L;
    G.n_layers = n_layers[g]; G.K = A.K; G.tiles = tiles[g]; G.kc = A.K / kSlabK; G.U = G.tiles * G.kc; G.group = A.group;
    G.pk = (A.group == 64) ? 64 : 128; G.pps = kSlabK / G.pk;
    G.zfp16 = (A.layout == B200Q_LAYOUT_HQQ) ? 1 : 0; G.zero_bias = A.zero_bias;
    G.region = g & 1; G.ncols = G.tiles * kTileN;
    // every CTA gets >= kc / 3 slabs (a tile then has <= 4 contributors: readers sum `smax` slots unconditionally)
    G.ncta = 3 * G.tiles < ncta ? 3 * G.tiles : ncta;
    if (G.ncta < ncta) small_group = true;            // idle CTAs: the tag hand-off alone does not order their finalizers
    G.x = groups[g][0].x; G.ldx = groups[g][0].ldx; G.xperm = A.x_perm; G.xmode = 0;
    int t0 = 0;
    for (int j = 0; j < n_layers[g]; ++j) {
      const LinearArgs& a = groups[g][j];
      ChLayer& L = G.layer[j];
      L.s = a.L.s; L.qz = a.L.qz; L.bias = a.L.bias; L.y = a.y; L.ldy = a.ldy; L.N = a.L.N; L.tile0 = t0; L.ntiles = (a.L.N + kTileN - 1) / kTileN;
      t0 += L.ntiles;
      {
        cuuint64_t dims[3] = {32, (cuuint64_t)a.L.N / 32, (cuuint64_t)a.L.K / 8};
        cuuint64_t strides[2] = {128, (cuuint64_t)a.L.N * 4};
        cuuint32_t box[3] = {32, 2, 32}, es[3] = {1, 1, 1};
        if (enc(&L.wmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)a.L.qw, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return B200Q_ERR_CUDA;
      }
    }
    // tile -> (first contributing CTA, number of contributing CTAs): CTA c owns slabs [c U / n, (c + 1) U / n)
    G.tab_off = (uint32_t)off;
    uint32_t* tab = (uint32_t*)(blob + off);
    off += ((size_t)G.tiles * 4 + 63) & ~(size_t)63;
    auto owner = [&](long long i) { return (int)((((i + 1) * G.ncta) + G.U - 1) / G.U - 1); };
    int smax = 1;
    for (int tl = 0; tl < G.tiles; ++tl) {
      const int c0 = owner((long long)tl * G.kc), c1 = owner((long long)(tl + 1) * G.kc - 1);
      tab[tl] = (uint32_t)c0 | ((uint32_t)(c1 - c0 + 1) << 16);
      if (c1 - c0 + 1 > smax) smax = c1 - c0 + 1;
    }
    G.smax = smax;
    if (smax > smax_all) smax_all = smax;
    if (G.ncols > ncols_all) ncols_all = G.ncols;
    for (int c = 0; c < G.ncta; ++c) {
      const long long a = (long long)c * G.U / G.ncta, b = (long long)(c + 1) * G.U / G.ncta;
      if (b > a) { const int nt = (int)((b - 1) / G.kc - a / G.kc + 1); if (nt > max_tiles) max_tiles = nt; }
    }
    // x produced by the previous group of the chain?  (x points into one of its outputs, same row stride)
    if (g > 0) {
      const ChGroup& S = GG[g - 1];
      for (int j = 0; j < S.n_layers; ++j) {
        const ChLayer& L = S.layer[j];
        const __half* x = groups[g][0].x;
        const __half* xe = x + (size_t)(M - 1) * groups[g][0].ldx + A.K;
        const __half* ye = L.y + (size_t)(M - 1) * L.ldy + L.N;
        const bool overlap = x < ye && L.y < xe;
        const bool lazy = x >= L.y && x + A.K <= L.y + L.N && groups[g][0].ldx == L.ldy && ((x - L.y) % 4) == 0;
        if (overlap && !lazy) return B200Q_ERR_UNSUPPORTED;   // y of the previous group is only complete one barrier later
        if (lazy) {
          G.xmode = 1; G.src_region = S.region; G.src_smax = S.smax; G.src_ncols = S.ncols;
          G.src_pcol0 = L.tile0 * kTileN + (int)(x - L.y); G.src_tab_off = S.tab_off;
          G.src_bias = L.bias ? L.bias + (x - L.y) : nullptr;
        }
      }
    }
  }
  if (max_tiles > 30) return B200Q_ERR_UNSUPPORTED;
  for (int g = 0; g < n_groups; ++g) {
    const ChGroup& G = GG[g];
    ChGS& S = GS[g];
    for (int j = 0; j < G.n_layers; ++j) {
      const ChLayer& L = G.layer[j];
      S.s[j] = L.s; S.qz[j] = L.qz; S.bias[j] = L.bias; S.y[j] = L.y; S.ldy[j] = L.ldy; S.N[j] = L.N; S.tile0[j] = L.tile0;
    }
    S.x = G.x; S.xperm = G.xperm; S.src_bias = G.src_bias; S.ldx = G.ldx;
    S.n_layers = G.n_layers; S.K = G.K; S.tiles = G.tiles; S.kc = G.kc; S.U = G.U; S.group = G.group; S.pk = G.pk; S.pps = G.pps;
    S.zfp16 = G.zfp16; S.zero_bias = G.zero_bias; S.ncta = G.ncta; S.region = G.region; S.smax = G.smax; S.ncols = G.ncols;
    S.xmode = G.xmode; S.src_smax = G.src_smax; S.src_ncols = G.src_ncols; S.src_pcol0 = G.src_pcol0; S.src_region = G.src_region;
    S.tab_off = (int32_t)G.tab_off;
    S.gshift = -1;
    for (int sh = 0; sh < 31; ++sh)
      if ((1 << sh) == G.group) S.gshift = sh;
  }
  H->magic = kChMagic; H->n_groups = (uint32_t)n_groups; H->M = (uint32_t)M; H->n_cta = (uint32_t)ncta; H->max_tiles = (uint32_t)max_tiles;
  H->region_elems = (uint64_t)smax_all * M * ncols_all;
  H->ws_bytes = kCounterBytes + kChWsHdr + 2ull * H->region_elems * 8ull;
  H->use_barrier = (g_ch_barrier || small_group) ? 1u : 0u;
  // shared memory
  uint32_t so = 0;
  H->off_bars = so; so += 1024;
  H->off_digits = so; so += (uint32_t)(kmax / 32) * 96u * (uint32_t)M; so = (so + 15u) & ~15u;
  H->off_parts = so; so += (uint32_t)(kmax / pkmin) * 8u * (uint32_t)M; so = (so + 15u) & ~15u;
  H->off_red = so; so += (uint32_t)kChWarps * (uint32_t)max_tiles * kTileN * 4u * (uint32_t)M;
  H->off_tab = so; so += (uint32_t)kChWarps * kTileN * 8u;
  H->off_zpad = so; so += 64;
  so = (so + 1023u) & ~1023u;
  const uint32_t fixed = so;
  const uint32_t budget = 226u * 1024u - 1024u;                            // 1 KB slack for the 1024-byte alignment of the base
  int slots = (int)((budget - fixed) / kSlabBytes);
  if (g_ch_slots > 0 && g_ch_slots < slots) slots = g_ch_slots;
  // a slot must always be consumed by the same warps (a waiter may be at most one mbarrier phase ahead): a warp's
  // consecutive units are 16 / pps = 8 or 4 slabs apart, so the ring holds 8 or 16 slabs
  slots = slots >= 16 ? 16 : (slots >= 8 ? 8 : 0);
  if (slots < 8) return B200Q_ERR_UNSUPPORTED;
  H->slots = (uint32_t)slots;
  H->off_ring = so; so += (uint32_t)slots * kSlabBytes;
  H->off_gs = so;
  int ncached = (int)((budget - so) / sizeof(ChGS));
  if (ncached > n_groups) ncached = n_groups;
  if (ncached < 0) ncached = 0;
  H->n_cached = (uint32_t)ncached;
  so += (uint32_t)(ncached * sizeof(ChGS));
  H->window = (uint32_t)((g_ch_window > 0 && g_ch_window <= slots) ? g_ch_window : (slots < 8 ? slots : 8));
  H->smem_bytes = so + 1024u;
  H->total_bytes = (uint32_t)need;
  if (ws_bytes) *ws_bytes = (size_t)H->ws_bytes;
  return B200Q_OK;
}

cudaError_t launch_decode_chain(const void* plan_host, const void* plan_dev, void* ws, size_t ws_bytes, cudaStream_t st) {
  const ChHeader* H = (const ChHeader*)plan_host;
  if (!H || H->magic != kChMagic || !plan_dev || !ws || ws_bytes < H->ws_bytes) return cudaErrorInvalidValue;
  static bool attr_done[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int mi = H->M == 1 ? 0 : 1;
  if (!attr_done[dev & 63][mi]) {
    cudaError_t e = mi == 0 ? cudaFuncSetAttribute(decode_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                            : cudaFuncSetAttribute(decode_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63][mi] = true;
  }
  ChParams p;
  p.plan = (const char*)plan_dev; p.ws = (char*)ws; p.h = *H; p.dbg = g_ch_dbg;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(H->n_cta);
  cfg.blockDim = dim3(kChThreads);
  cfg.dynamicSmemBytes = H->smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  count_launch();
  return mi == 0 ? cudaLaunchKernelEx(&cfg, decode_chain_kernel<1>, p) : cudaLaunchKernelEx(&cfg, decode_chain_kernel<2>, p);
}

}  // namespace b200q
