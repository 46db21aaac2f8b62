// Decode CHAIN kernel (M <= 2): a recorded sequence of b200q_linear_group calls executed by ONE persistent launch.
//
// Why: at batch 1 a Llama-sized QuantLinear is 1.3-7 us of HBM streaming, and the launch-per-layer kernels lose as much
// again per launch to the serial chain "previous layer stored -> kernel boundary -> x loaded and split into digits ->
// first weights consumed -> reduced -> stored" during which HBM idles (profiles/r1_decode_timeline_bench.txt), plus
// 2-3 CTAs per SM of uneven work.  Here one CTA per SM lives for the whole chain:
//
//   * ONE producer thread per SM streams the packed weights of every step of the chain, in order, through ONE
//     shared-memory ring with TMA (2-D boxes: 32 columns x 256 k = 4 KB, 128-byte swizzle).  The ring does not drain
//     at layer boundaries: while the consumers wait for a step's activations, the next step's weights keep arriving
//     (128 KB per SM, ~19 MB on the chip);
//   * every CTA OWNS whole 32-column tiles of a step (a contiguous range, all of K): the 16 consumer warps split K
//     among themselves and reduce through shared memory, so no partial sum ever crosses an SM;
//   * the consumers run the integer tensor path of gemv_imma.cu (IMMA.16832 on nibbles, activations as three
//     base-128 digits, exact int32 accumulation per <= 128-k part, fp32 scale/zero fix-up per group);
//   * the hand-off between steps is data-flow, not a barrier: a finished output element is stored as ONE 32-bit word
//     fp16 | tag << 16 (tag = launch epoch and step number); 4-byte stores are single-copy atomic, so the next step's
//     x stage simply re-reads the words it needs until they carry the awaited tag -- no fence, no atomic, no flag;
//     the plain fp16 y (what QuantLinear.forward returns) is written beside it.
//
// Replaces a run of ort_ops.gemv / gemm_forward_cuda calls (dq_gemv.cu:40-177, gemm_cuda_gen.cu:1102-1161) at M <= 2.
#include <cuda.h>

#include <cstring>
#include <type_traits>
#include <vector>

#include "gemv_stream.cuh"

namespace b200q {

static constexpr int kChWarps = 16;                       // consumer warps
static constexpr int kChThreads = 32 * (kChWarps + 4);    // + warpgroup 0: producer warp and three idle warps
static constexpr int kSlabK = 256, kTileN = 32, kSlabBytes = 4096;
static constexpr uint32_t kChMagic = 0xB2C4A119u;
static constexpr uint32_t NIBM = 0x0f0f0f0fu;
static constexpr int kExitWord = 513, kErrWord = 514;     // u32 words inside the 4 KB counter region
static constexpr int kAuxD = 4;                           // slabs a warp's (scale, zero) words are fetched ahead of their use
static constexpr size_t kChWsHdr = 256;                   // after the counter region: word 0 = launch epoch (never reset)

// One step of the chain.  Copied into shared memory at kernel start (a global load costs ~1 us under the weight stream
// and the step descriptors are on every step's critical path).
struct alignas(16) ChGS {
  const __half* s[kMaxGroupLayers];
  const void* qz[kMaxGroupLayers];
  const __half* bias[kMaxGroupLayers];
  __half* y[kMaxGroupLayers];
  int64_t ldy[kMaxGroupLayers];
  const __half* x;
  const int32_t* xperm;
  int64_t ldx;
  int32_t N[kMaxGroupLayers], tile0[kMaxGroupLayers];
  int32_t n_layers, K, tiles, kc, group, pk, pps, zfp16, zero_bias, region, ncols, xmode;
  int32_t src_region, src_ncols, src_pcol0, gshift, pad_[2];
};
static_assert(sizeof(ChGS) % 16 == 0, "ChGS is copied as uint4");
struct ChHeader {
  uint32_t magic, n_groups, M, n_cta, slots, max_tiles, smem_bytes, total_bytes;
  uint32_t off_bars, off_digits, off_parts, off_red, off_zpad, off_ring, off_gs, n_cached;
  uint32_t off_aux, pad2_[3];
  uint32_t gs_off, maps_off, window, pad_;
  uint64_t region_words;                       // 32-bit words per tagged-activation region
  uint64_t ws_bytes;
};
struct ChParams {
  const char* plan;       // device copy of the plan blob
  char* ws;
  ChHeader h;
  unsigned long long* dbg;
};

__device__ __forceinline__ void tma_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// tagged activations: fp16 | tag << 16 in one naturally aligned 32-bit word
__device__ __forceinline__ uint4 ld_tagged4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_tagged1(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_tagged(uint32_t* p, uint32_t w) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(w) : "memory");
}
__device__ __forceinline__ uint2 ldcg_u2(const void* p) {
  uint2 v;
  asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned short ldcg_u16(const void* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kChWarps * 32) : "memory"); }

// not volatile: the accumulators are consumed, so the instruction cannot be dropped, and the scheduler may interleave
// it with the unpack ALU work and the loads of later sub-steps
__device__ __forceinline__ void imma_acc(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// diagnostic: 16 x u64 per (step, CTA): consumer warp 0 {0 step start, 1 x words loaded, 2 digits done, 3 own slabs done,
// 4 all warps done, 5 outputs stored}, producer {8 first slab issued, 9 last slab issued, 10 ns stalled}, consumer warp 0
// cycle counters {11 waiting for weights, 12 slabs, 13 slab loop}
#define CH_STAMP(g, i) do { if (p.dbg && lane == 0) p.dbg[((size_t)(g) * gridDim.x + blockIdx.x) * 16 + (i)] = st_gtime(); } while (0)

template <int MTOK>
__global__ void __launch_bounds__(kChThreads, 1) decode_chain_kernel(const __grid_constant__ ChParams p) {
  extern __shared__ __align__(1024) char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const ChHeader& H = p.h;
  const int NS = (int)H.slots;
  const uint32_t bars = smem_u32(smem + H.off_bars);
  const uint32_t bar_full = bars, bar_empty = bars + 8u * 32u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x, ncta = gridDim.x;
  const int NG = (int)H.n_groups;
  const ChGS* gs_global = reinterpret_cast<const ChGS*>(p.plan + H.gs_off);

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(smem + H.off_bars) + i, 1);
      mbar_init(reinterpret_cast<uint64_t*>(smem + H.off_bars) + 32 + i, 1);
    }
    fence_mbar_init();
  }
  if (tid < 16) reinterpret_cast<uint32_t*>(smem + H.off_zpad)[tid] = 0u;
  __syncthreads();
  pdl_launch_dependents();

  // register budget: the control warpgroup gives registers back, the four consumer warpgroups take them
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp > 1 || lane != 0) return;                                     // two producer threads: even / odd slabs
    // =========================================== producer ===========================================
    // packed weights are constants: the stream starts before the upstream kernel has finished (no griddepcontrol.wait).
    // At most `window` slabs are in flight per SM: more only lengthens every queue between the SM and HBM (a demand
    // load behind 128 KB of bulk reads waits microseconds) without adding bandwidth; landed slabs may fill the whole ring.
    const uint32_t ring = smem_u32(smem + H.off_ring);
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(p.plan + H.maps_off);
    const int W = (int)H.window;
    int slot = 0, wslot = 0, issued = 0;
    uint32_t round = 0, wround = 0;
    for (int g = 0; g < NG; ++g) {
      const ChGS* G = gs_global + g;
      const int T = G->tiles, KC = G->kc, nl = G->n_layers;
      const int ta = (int)((long long)c * T / ncta), tb = (int)((long long)(c + 1) * T / ncta);
      if (ta >= tb) continue;
      if (warp == 0) CH_STAMP(g, 8);
      unsigned long long stall = 0;
      int j = 0;
      while (j + 1 < nl && ta >= G->tile0[j + 1]) ++j;
      int tile0 = G->tile0[j], tile_end = (j + 1 < nl) ? G->tile0[j + 1] : T;
      const void* wmap = maps + (size_t)g * kMaxGroupLayers + j;
      asm volatile("prefetch.tensormap [%0];" ::"l"(wmap) : "memory");
      for (int tile = ta; tile < tb; ++tile) {
        if (tile == tile_end) {
          ++j;
          tile0 = tile_end; tile_end = (j + 1 < nl) ? G->tile0[j + 1] : T;
          wmap = maps + (size_t)g * kMaxGroupLayers + j;
          asm volatile("prefetch.tensormap [%0];" ::"l"(wmap) : "memory");
        }
        for (int kk = 0; kk < KC; ++kk) {
          if ((issued & 1) == warp) {                                      // this thread's slab
            const unsigned long long t0 = p.dbg ? st_gtime() : 0ull;
            if (round > 0) mbar_wait_b(bar_empty + 8u * slot, (round - 1) & 1u, p.ws, 0x100u + g);
            if (issued >= W) mbar_wait_b(bar_full + 8u * wslot, wround & 1u, p.ws, 0x180u + g);   // the slab issued W slabs ago has landed
            if (p.dbg) stall += st_gtime() - t0;
            const uint32_t fb = bar_full + 8u * slot;
            mbar_expect_tx_a(fb, kSlabBytes);
            tma_2d(ring + (uint32_t)slot * kSlabBytes, wmap, kTileN * (tile - tile0), 32 * kk, fb);
          }
          if (issued >= W && ++wslot == NS) { wslot = 0; ++wround; }
          ++issued;
          if (++slot == NS) { slot = 0; ++round; }
        }
      }
      if (warp == 0) CH_STAMP(g, 9);
      if (p.dbg && warp == 0) p.dbg[((size_t)g * gridDim.x + blockIdx.x) * 16 + 10] = stall;
    }
    return;
  }

  // =========================================== consumers ============================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  const int w = warp - 4, ctid = tid - 128;
  const int ncached = (int)H.n_cached;
  for (int i = ctid; i < ncached * (int)(sizeof(ChGS) / 16); i += kChWarps * 32)
    reinterpret_cast<uint4*>(smem + H.off_gs)[i] = __ldg(reinterpret_cast<const uint4*>(gs_global) + i);
  pdl_wait();                                                              // upstream results (x, y buffers, workspace) are complete
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(p.ws + kCounterBytes);
  const uint32_t tag0 = epoch * (uint32_t)(NG + 1) + 1u;                   // tag of step g: (tag0 + g) & 0xffff
  uint32_t* Ybase = reinterpret_cast<uint32_t*>(p.ws + kCounterBytes + kChWsHdr);
  consumer_bar();
  const ChGS* gs_cache = reinterpret_cast<const ChGS*>(smem + H.off_gs);

  const int g8 = lane >> 2, t = lane & 3;
  char* xdig = smem + H.off_digits;
  float2* parts = reinterpret_cast<float2*>(smem + H.off_parts);
  float* red = reinterpret_cast<float*>(smem + H.off_red);
  const int MT = (int)H.max_tiles;
  float* redw = red + (size_t)w * MT * kTileN * MTOK;
  const uint32_t ring = smem_u32(smem + H.off_ring);
  // weights: a slab is 32 packed rows of 128 bytes (32 columns), 16-byte chunk index XOR (row & 7) (TMA 128-byte swizzle).
  // Lane (g8, t) reads packed row 4 ss + t, chunk C = (g8 >> 1) + 4 (g8 & 1): a quarter-warp then touches 8 distinct chunks.
  const int C = (g8 >> 1) + 4 * (g8 & 1);
  const uint32_t woff = (uint32_t)(t * 128 + ((C ^ t) << 4));             // + ss * 512; odd ss: rows 4..7 -> chunk ^ 4
  // digits: B column g8 -> token g8 >> 2, digit g8 & 3 (3 = unused -> zero pad)
  const bool xreal = (g8 & 3) < 3 && (g8 >> 2) < MTOK;
  const uint32_t xlane = xreal ? smem_u32(xdig) + (uint32_t)(((g8 >> 2) * 3 + (g8 & 3)) * 32 + t * 8) : smem_u32(smem + H.off_zpad);
  const uint32_t xsub = xreal ? (uint32_t)(96 * MTOK) : 0u;
  const int mytok = t >> 1;
  const bool fx_on = mytok < MTOK;
  const float dscale = (t & 1) ? (1.0f / 128.0f) : 128.0f, zflag = (t & 1) ? 0.f : 1.f;
  const int nsmask = NS - 1;                                               // NS is 16 or 32

  uint32_t q0 = 0;                                                         // CTA-local slab sequence number at step start
  for (int g = 0; g < NG; ++g) {
    const ChGS* G = g < ncached ? gs_cache + g : gs_global + g;
    const int K = G->K, T = G->tiles, KC = G->kc, PK = G->pk, PPS = G->pps;
    const int zf = G->zfp16, zbias = G->zero_bias, gshift = G->gshift, nl = G->n_layers;
    const int ta = (int)((long long)c * T / ncta), tb = (int)((long long)(c + 1) * T / ncta);
    const int nt = tb - ta, nslab = nt * KC;
    const int nparts = K / PK;
    if (w == 0) CH_STAMP(g, 0);

    // ---- this warp's first slab (slab q -> warp q % 16); the (scale, zero) words of a slab's first two parts are
    //      fetched one slab ahead of their use ----
    const uint32_t qfirst = q0 + (((uint32_t)w + kChWarps - (q0 & (kChWarps - 1))) & (kChWarps - 1));
    int left = (int)(q0 + (uint32_t)nslab) - (int)qfirst;                  // > 0: this warp has slabs in this step
    int tile = 0, kk = 0;
    // (scale, zero) words: a global load takes 1-2 us under the weight stream, longer than a slab's arithmetic, so the
    // words of a slab's first two parts travel through a per-warp cp.async ring, kAuxD slabs ahead of their use
    // (quad leaders copy: the four lanes of a quad own the same four columns)
    int ptile = 0, pkk = 0, plj = 0, pltile0 = 0, pltile_end = 0, plN = 0, pleft = 0;
    const __half* pls = nullptr;
    const char* plqz = nullptr;
    const uint32_t auxw = smem_u32(smem + H.off_aux) + (uint32_t)w * (kAuxD * 256) + (uint32_t)g8 * 32;
    auto pf_layer = [&]() {
      while (plj + 1 < nl && ptile >= G->tile0[plj + 1]) ++plj;
      pltile0 = G->tile0[plj]; pltile_end = (plj + 1 < nl) ? G->tile0[plj + 1] : T; plN = G->N[plj];
      pls = G->s[plj]; plqz = reinterpret_cast<const char*>(G->qz[plj]);
    };
    auto pf_issue = [&](int stage) {                                       // slab (ptile, pkk) -> ring stage; then the cursor moves on
      if (pleft > 0) {
        if (ptile >= pltile_end) pf_layer();
        if (t == 0) {
          const int col = (ptile - pltile0) * kTileN + 4 * C;
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const int k = pkk * kSlabK + pp * PK;
            const int grow = gshift >= 0 ? (k >> gshift) : k / G->group;
            const uint32_t dst = auxw + (uint32_t)stage * 256u + (uint32_t)pp * 16u;
            cp_async8(dst, pls + (size_t)grow * plN + col);
            if (zf) cp_async8(dst + 8u, plqz + ((size_t)grow * plN + col) * 2);
            else cp_async4(dst + 8u, plqz + ((size_t)grow * (plN >> 3) + (col >> 3)) * 4);
          }
        }
        pleft -= kChWarps;
        pkk += kChWarps;
        while (pkk >= KC) { pkk -= KC; ++ptile; }
      }
      cp_async_commit();
    };
    // parts 2, 3 of a group-64 slab (rare path): plain loads at the point of use
    auto fetch_sz = [&](int tile_, int kk_, int pp_, uint2& so, uint2& zo) {
      int j = 0;
      while (j + 1 < nl && tile_ >= G->tile0[j + 1]) ++j;
      const int lN = G->N[j];
      const int col = (tile_ - G->tile0[j]) * kTileN + 4 * C;
      const int k = kk_ * kSlabK + pp_ * PK;
      const int grow = gshift >= 0 ? (k >> gshift) : k / G->group;
      const char* lqz = reinterpret_cast<const char*>(G->qz[j]);
      so = __ldg(reinterpret_cast<const uint2*>(G->s[j] + (size_t)grow * lN + col));
      if (zf) zo = __ldg(reinterpret_cast<const uint2*>(lqz + ((size_t)grow * lN + col) * 2));
      else zo.x = __ldg(reinterpret_cast<const uint32_t*>(lqz + ((size_t)grow * (lN >> 3) + (col >> 3)) * 4));
    };
    if (left > 0) {
      const int i0 = (int)(qfirst - q0);
      const int tl = i0 / KC;
      tile = ta + tl; kk = i0 - tl * KC;
      ptile = tile; pkk = kk; pleft = left;
      pf_layer();
#pragma unroll
      for (int d = 0; d < kAuxD; ++d) pf_issue(d);
    }

    // ---- x -> three base-128 digits per element, power-of-two scale per part (PK k).  All of this warp's words are
    //      requested first (one L2 round trip), then polled until they carry the producing step's tag. -------------------
    if (nslab > 0) {
      const int EL = PK >> 5;                                              // elements per lane: 4 (PK = 128) or 2 (PK = 64)
      const int xmode = G->xmode;
      const int* xperm = G->xperm;
      const uint32_t* Ysrc = Ybase + (size_t)G->src_region * H.region_words + G->src_pcol0;
      const int sncols = G->src_ncols;
      const uint32_t xtag = (tag0 + (uint32_t)(g - 1)) & 0xffffu;
      const bool fast = (EL == 4) && !xperm;
      for (int pr0 = w; pr0 < nparts; pr0 += 4 * kChWarps) {
#pragma unroll
        for (int m = 0; m < MTOK; ++m) {
          uint4 raw[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) raw[it] = make_uint4(0u, 0u, 0u, 0u);
          if (fast) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int pr = pr0 + it * kChWarps;
              if (pr < nparts) {
                const int k0 = pr * PK + 4 * lane;
                if (xmode == 1) raw[it] = ld_tagged4(Ysrc + (size_t)m * sncols + k0);
                else { const uint2 v = ldcg_u2(G->x + (size_t)m * G->ldx + k0); raw[it] = make_uint4(v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16); }
              }
            }
            if (xmode == 1) {
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int pr = pr0 + it * kChWarps;
                if (pr < nparts) {
                  ChSpin sp;
                  while (((raw[it].x >> 16) != xtag) | ((raw[it].y >> 16) != xtag) | ((raw[it].z >> 16) != xtag) | ((raw[it].w >> 16) != xtag)) {
                    sp.tick(p.ws, 0x700u + g);
                    raw[it] = ld_tagged4(Ysrc + (size_t)m * sncols + pr * PK + 4 * lane);
                  }
                }
              }
            }
          }
          if (w == 0 && m == 0 && pr0 == w) CH_STAMP(g, 1);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int pr = pr0 + it * kChWarps;
            if (pr < nparts) {
              float xv[4] = {0.f, 0.f, 0.f, 0.f};
              if (fast) {
                xv[0] = __half2float(__ushort_as_half((unsigned short)raw[it].x)); xv[1] = __half2float(__ushort_as_half((unsigned short)raw[it].y));
                xv[2] = __half2float(__ushort_as_half((unsigned short)raw[it].z)); xv[3] = __half2float(__ushort_as_half((unsigned short)raw[it].w));
              } else {                                                     // group 64 and / or act-order gather: element-wise
                const int k0 = pr * PK + EL * lane;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (e < EL) {
                    const int kx = xperm ? __ldg(xperm + k0 + e) : k0 + e;
                    if (xmode == 1) {
                      uint32_t v = ld_tagged1(Ysrc + (size_t)m * sncols + kx);
                      ChSpin sp;
                      while ((v >> 16) != xtag) { sp.tick(p.ws, 0x700u + g); v = ld_tagged1(Ysrc + (size_t)m * sncols + kx); }
                      xv[e] = __half2float(__ushort_as_half((unsigned short)v));
                    } else {
                      xv[e] = __half2float(__ushort_as_half(ldcg_u16(G->x + (size_t)m * G->ldx + kx)));
                    }
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
                const int tv = __float2int_rn(xv[e] * sc);                 // |tv| <= 2^20
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
      }
    }
    consumer_bar();
    if (w == 0) CH_STAMP(g, 2);

    // ---- this warp's slabs ---------------------------------------------------------------------------------------
    float tot[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    uint32_t touched = 0;
    int cur_tl = -1;                                                       // local tile (tile - ta) of `tot`
    long long cyc_wait = 0, cyc_loop = 0;
    int n_units = 0;
    auto flush = [&]() {
      if (cur_tl >= 0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float a0 = tot[q][0] + __shfl_xor_sync(0xffffffffu, tot[q][0], 1);
          const float a1 = tot[q][1] + __shfl_xor_sync(0xffffffffu, tot[q][1], 1);
          if (fx_on && !(t & 1)) {
            const int n = 4 * C + 2 * q;                                   // MMA q: row g8 -> column n, row g8 + 8 -> n + 1
            float* r = redw + ((size_t)cur_tl * kTileN + n) * MTOK + mytok;
            r[0] = a0;
            r[MTOK] = a1;
          }
          tot[q][0] = tot[q][1] = 0.f;
        }
        touched |= 1u << cur_tl;
      }
    };
    // one part (PK k) of the slab in `slot`: int32 accumulation over its sub-steps, then the (scale, zero) fix-up
    auto do_part = [&](uint32_t wb, uint32_t xp, const float2 pt, const uint2 sw, const uint2 zw, auto nsub_c) {
      constexpr int NSUB = decltype(nsub_c)::value;
      int acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      // all shared-memory loads of the part first (independent), then the unpack + IMMA chain
      uint4 wv[NSUB];
      uint2 xb[NSUB];
#pragma unroll
      for (int s = 0; s < NSUB; ++s) {
        // packed rows 4 s + t of the part; rows 4..7 of an 8-row swizzle period flip chunk bit 2
        wv[s] = lds128_s(wb + (uint32_t)s * 512u + (woff ^ (((uint32_t)s & 1u) << 6)));
        xb[s] = lds64_s(xp + (uint32_t)s * xsub);
      }
#pragma unroll
      for (int s = 0; s < NSUB; ++s) {
        const uint4 wa = wv[s];
        imma_acc(acc[0], wa.x & NIBM, wa.y & NIBM, (wa.x >> 4) & NIBM, (wa.y >> 4) & NIBM, xb[s].x, xb[s].y);
        imma_acc(acc[1], wa.z & NIBM, wa.w & NIBM, (wa.z >> 4) & NIBM, (wa.w >> 4) & NIBM, xb[s].x, xb[s].y);
      }
      if (fx_on) {
        // y += s * (2^-E * dscale * (d_a 2^7 + d_b) - zflag * z * sum(x)) for this lane's 4 (column, token) outputs
        const float xs = pt.x * dscale, sxz = pt.y * zflag;
        const __half2 s01 = *reinterpret_cast<const __half2*>(&sw.x), s23 = *reinterpret_cast<const __half2*>(&sw.y);
        float z0, z1, z2, z3;
        if (zf) {
          const __half2 z01 = *reinterpret_cast<const __half2*>(&zw.x), z23 = *reinterpret_cast<const __half2*>(&zw.y);
          z0 = __low2float(z01); z1 = __high2float(z01); z2 = __low2float(z23); z3 = __high2float(z23);
        } else {
          const uint32_t zz = zw.x >> (16 * (C & 1));                      // columns 4 C .. 4 C + 3: nibbles 4 (C & 1) .. + 3 of the word
          z0 = (float)((zz + (uint32_t)zbias) & 15u); z1 = (float)(((zz >> 4) + (uint32_t)zbias) & 15u);
          z2 = (float)(((zz >> 8) + (uint32_t)zbias) & 15u); z3 = (float)(((zz >> 12) + (uint32_t)zbias) & 15u);
        }
        const float v00 = (float)((acc[0][0] << 7) + acc[0][1]), v01 = (float)((acc[0][2] << 7) + acc[0][3]);
        const float v10 = (float)((acc[1][0] << 7) + acc[1][1]), v11 = (float)((acc[1][2] << 7) + acc[1][3]);
        tot[0][0] = fmaf(__low2float(s01), fmaf(v00, xs, -z0 * sxz), tot[0][0]);
        tot[0][1] = fmaf(__high2float(s01), fmaf(v01, xs, -z1 * sxz), tot[0][1]);
        tot[1][0] = fmaf(__low2float(s23), fmaf(v10, xs, -z2 * sxz), tot[1][0]);
        tot[1][1] = fmaf(__high2float(s23), fmaf(v11, xs, -z3 * sxz), tot[1][1]);
      }
    };
    if (left > 0) {
      const long long tl0 = p.dbg ? clock64() : 0;
      int slot = (int)(qfirst & (uint32_t)nsmask);
      uint32_t par = (qfirst / (uint32_t)NS) & 1u;
      int stage = 0;
      for (; left > 0; left -= kChWarps) {
        const int tl = tile - ta;
        if (tl != cur_tl) { flush(); cur_tl = tl; }
        const int kk_cur = kk, tile_cur = tile;
        cp_async_wait<kAuxD - 1>();                                        // this slab's (scale, zero) words have landed
        __syncwarp();
        const uint4 a0 = *reinterpret_cast<const uint4*>(smem + H.off_aux + (size_t)w * (kAuxD * 256) + (size_t)stage * 256 + g8 * 32);
        const uint4 a1 = *reinterpret_cast<const uint4*>(smem + H.off_aux + (size_t)w * (kAuxD * 256) + (size_t)stage * 256 + g8 * 32 + 16);
        const uint2 s0 = make_uint2(a0.x, a0.y), z0 = make_uint2(a0.z, a0.w), s1 = make_uint2(a1.x, a1.y), z1 = make_uint2(a1.z, a1.w);
        __syncwarp();                                                      // every lane has read the stage before it is refilled
        pf_issue(stage);
        stage = (stage + 1) & (kAuxD - 1);
        uint2 s2 = make_uint2(0u, 0u), s3 = s2, z2 = s2, z3 = s2;
        if (PPS == 4) {                                                    // group 64: parts 2, 3 of this slab (rare path, not prefetched)
          fetch_sz(tile_cur, kk_cur, 2, s2, z2);
          fetch_sz(tile_cur, kk_cur, 3, s3, z3);
        }
        kk += kChWarps;
        while (kk >= KC) { kk -= KC; ++tile; }
        if (p.dbg) {
          const long long tw = clock64();
          mbar_wait_b(bar_full + 8u * slot, par, p.ws, 0x500u + g);
          cyc_wait += clock64() - tw;
          ++n_units;
        } else {
          mbar_wait_b(bar_full + 8u * slot, par, p.ws, 0x500u + g);
        }
        const uint32_t wb = ring + (uint32_t)slot * kSlabBytes;
        const uint32_t xp = xlane + (uint32_t)(kk_cur * 8) * xsub;
        const float2* pp = parts + (size_t)(kk_cur * PPS) * MTOK + (fx_on ? mytok : 0);
        if (PPS == 2) {
          do_part(wb, xp, pp[0], s0, z0, std::integral_constant<int, 4>());
          do_part(wb + 2048u, xp + 4u * xsub, pp[MTOK], s1, z1, std::integral_constant<int, 4>());
        } else {
          do_part(wb, xp, pp[0], s0, z0, std::integral_constant<int, 2>());
          do_part(wb + 1024u, xp + 2u * xsub, pp[MTOK], s1, z1, std::integral_constant<int, 2>());
          do_part(wb + 2048u, xp + 4u * xsub, pp[2 * MTOK], s2, z2, std::integral_constant<int, 2>());
          do_part(wb + 3072u, xp + 6u * xsub, pp[3 * MTOK], s3, z3, std::integral_constant<int, 2>());
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bar_empty + 8u * slot);               // this warp is done with the slot
        slot += kChWarps;
        if (slot >= NS) { slot -= NS; par ^= 1u; }
      }
      cp_async_wait<0>();
      flush();
      if (p.dbg) cyc_loop = clock64() - tl0;
    }
    // tiles of this CTA the warp never touched contribute zeros
    for (int tl = 0; tl < nt; ++tl)
      if (!((touched >> tl) & 1u))
        for (int idx = lane; idx < kTileN * MTOK; idx += 32) redw[(size_t)tl * kTileN * MTOK + idx] = 0.f;
    if (w == 0) CH_STAMP(g, 3);
    consumer_bar();
    if (w == 0) CH_STAMP(g, 4);
    // ---- reduction over the 16 warps (fixed order), bias, fp16: tagged word for the next step + plain y ------------
    {
      uint32_t* Y = Ybase + (size_t)G->region * H.region_words;
      const uint32_t tag = ((tag0 + (uint32_t)g) & 0xffffu) << 16;
      const int ncols = G->ncols;
      for (int idx = ctid; idx < nt * kTileN * MTOK; idx += kChWarps * 32) {
        float sum = 0.f;
#pragma unroll
        for (int wq = 0; wq < kChWarps; ++wq) sum += red[(size_t)wq * MT * kTileN * MTOK + idx];
        const int tl = idx / (kTileN * MTOK), rem = idx - tl * (kTileN * MTOK);
        const int n = rem / MTOK, m = rem - n * MTOK;
        const int tile_o = ta + tl;
        int j = 0;
        while (j + 1 < nl && tile_o >= G->tile0[j + 1]) ++j;
        const int col = (tile_o - G->tile0[j]) * kTileN + n;
        const __half* bias = G->bias[j];
        if (bias) sum += __half2float(__ldg(bias + col));
        const __half h = __float2half_rn(sum);
        st_tagged(Y + (size_t)m * ncols + (size_t)tile_o * kTileN + n, tag | (uint32_t)__half_as_ushort(h));
        G->y[j][(size_t)m * G->ldy[j] + col] = h;
      }
    }
    if (w == 0) {
      CH_STAMP(g, 5);
      if (p.dbg && lane == 0) {
        unsigned long long* d = p.dbg + ((size_t)g * gridDim.x + blockIdx.x) * 16;
        d[11] = (unsigned long long)cyc_wait; d[12] = (unsigned long long)n_units; d[13] = (unsigned long long)cyc_loop;
      }
    }
    q0 += (uint32_t)nslab;
  }
  // the last CTA out advances the launch epoch and re-zeroes the exit counter
  consumer_bar();
  if (ctid == 0) {
    uint32_t* ctr = reinterpret_cast<uint32_t*>(p.ws);
    __threadfence();
    const uint32_t old = atomicAdd(ctr + kExitWord, 1u);
    if (old == (uint32_t)ncta - 1u) {
      *reinterpret_cast<volatile uint32_t*>(p.ws + kCounterBytes) = epoch + 1u;
      ctr[kExitWord] = 0u;
      __threadfence();
    }
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

static int g_ch_ctas = 0, g_ch_slots = 0, g_ch_window = 0;
static unsigned long long* g_ch_dbg = nullptr;
void decode_chain_set_debug(unsigned long long* buf) { g_ch_dbg = buf; }
void decode_chain_set_option(int which, int value) {
  if (which == 0) g_ch_ctas = value;
  else if (which == 1) g_ch_slots = value;
  else if (which == 3) g_ch_window = value;
}

static size_t ch_plan_bytes(int n_groups) {
  size_t b = (sizeof(ChHeader) + 63) & ~(size_t)63;
  b += ((size_t)n_groups * sizeof(ChGS) + 63) & ~(size_t)63;
  b += (size_t)n_groups * kMaxGroupLayers * sizeof(CUtensorMap);
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
  int kmax = 0, pkmin = 128;
  for (int g = 0; g < n_groups; ++g) {
    if (n_layers[g] < 1 || n_layers[g] > kMaxGroupLayers) return B200Q_ERR_SHAPE;
    const LayerView& A = groups[g][0].L;
    for (int j = 0; j < n_layers[g]; ++j) {
      const LinearArgs& a = groups[g][j];
      if (!ch_layer_ok(a.L)) return B200Q_ERR_UNSUPPORTED;
      if (a.L.K != A.K || a.L.group != A.group || a.L.layout != A.layout || a.L.zero_bias != A.zero_bias || a.L.x_perm != A.x_perm ||
          a.x != groups[g][0].x || a.ldx != groups[g][0].ldx || a.M != M)
        return B200Q_ERR_UNSUPPORTED;
      if (!a.y || a.ldy < a.L.N || ((uintptr_t)a.y & 1)) return B200Q_ERR_ALIGNMENT;
    }
    if (!groups[g][0].x || ((uintptr_t)groups[g][0].x & 7) || (groups[g][0].ldx & 3)) return B200Q_ERR_ALIGNMENT;
    if (A.K > kmax) kmax = A.K;
    if (A.group == 64) pkmin = 64;
  }
  const size_t need = ch_plan_bytes(n_groups);
  if (plan_bytes) *plan_bytes = need;
  if (!plan_out) return B200Q_OK;                                         // size query
  if (plan_cap < need) return B200Q_ERR_WORKSPACE;
  char* blob = (char*)plan_out;
  memset(blob, 0, need);
  ChHeader* H = (ChHeader*)blob;
  size_t off = (sizeof(ChHeader) + 63) & ~(size_t)63;
  H->gs_off = (uint32_t)off;
  ChGS* GS = (ChGS*)(blob + off);
  off += ((size_t)n_groups * sizeof(ChGS) + 63) & ~(size_t)63;
  H->maps_off = (uint32_t)off;
  CUtensorMap* maps = (CUtensorMap*)(blob + off);
  int max_tiles = 1, ncols_all = 0;
  for (int g = 0; g < n_groups; ++g) {
    ChGS& G = GS[g];
    const LayerView& A = groups[g][0].L;
    G.n_layers = n_layers[g]; G.K = A.K; G.kc = A.K / kSlabK; G.group = A.group;
    G.pk = (A.group == 64) ? 64 : 128; G.pps = kSlabK / G.pk;
    G.zfp16 = (A.layout == B200Q_LAYOUT_HQQ) ? 1 : 0; G.zero_bias = A.zero_bias;
    G.region = g & 1;
    G.x = groups[g][0].x; G.ldx = groups[g][0].ldx; G.xperm = A.x_perm; G.xmode = 0;
    G.gshift = -1;
    for (int sh = 0; sh < 31; ++sh)
      if ((1 << sh) == A.group) G.gshift = sh;
    int t0 = 0;
    for (int j = 0; j < n_layers[g]; ++j) {
      const LinearArgs& a = groups[g][j];
      G.s[j] = a.L.s; G.qz[j] = a.L.qz; G.bias[j] = a.L.bias; G.y[j] = a.y; G.ldy[j] = a.ldy; G.N[j] = a.L.N; G.tile0[j] = t0;
      t0 += a.L.N / kTileN;
      cuuint64_t dims[2] = {(cuuint64_t)a.L.N, (cuuint64_t)a.L.K / 8};
      cuuint64_t strides[1] = {(cuuint64_t)a.L.N * 4};
      cuuint32_t box[2] = {(cuuint32_t)kTileN, 32}, es[2] = {1, 1};
      if (enc(&maps[(size_t)g * kMaxGroupLayers + j], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)a.L.qw, dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return B200Q_ERR_CUDA;
    }
    G.tiles = t0; G.ncols = t0 * kTileN;
    if (G.ncols > ncols_all) ncols_all = G.ncols;
    const int mt = (G.tiles + ncta - 1) / ncta;
    if (mt > max_tiles) max_tiles = mt;
    // x produced by the previous step of the chain?  (x points into one of its outputs, same row stride)
    if (g > 0) {
      const ChGS& S = GS[g - 1];
      for (int j = 0; j < S.n_layers; ++j) {
        const __half* x = groups[g][0].x;
        const __half* xe = x + (size_t)(M - 1) * groups[g][0].ldx + A.K;
        const __half* ye = S.y[j] + (size_t)(M - 1) * S.ldy[j] + S.N[j];
        const bool overlap = x < ye && S.y[j] < xe;
        const bool lazy = x >= S.y[j] && x + A.K <= S.y[j] + S.N[j] && groups[g][0].ldx == S.ldy[j] && ((x - S.y[j]) % 4) == 0;
        if (overlap && !lazy) return B200Q_ERR_UNSUPPORTED;   // the plain y of the previous step has no completion flag
        if (lazy) {
          G.xmode = 1; G.src_region = S.region; G.src_ncols = S.ncols;
          G.src_pcol0 = S.tile0[j] * kTileN + (int)(x - S.y[j]);
        }
      }
    }
    // an x that points into an OLDER output of this chain would be read without any ordering: refuse
    for (int g2 = 0; g2 + 1 < g; ++g2)
      for (int j = 0; j < GS[g2].n_layers; ++j) {
        const __half* x = groups[g][0].x;
        const __half* xe = x + (size_t)(M - 1) * groups[g][0].ldx + A.K;
        const __half* ye = GS[g2].y[j] + (size_t)(M - 1) * GS[g2].ldy[j] + GS[g2].N[j];
        if (x < ye && GS[g2].y[j] < xe) return B200Q_ERR_UNSUPPORTED;
      }
  }
  if (max_tiles > 30) return B200Q_ERR_UNSUPPORTED;
  H->magic = kChMagic; H->n_groups = (uint32_t)n_groups; H->M = (uint32_t)M; H->n_cta = (uint32_t)ncta; H->max_tiles = (uint32_t)max_tiles;
  H->region_words = (uint64_t)M * ncols_all;
  H->ws_bytes = kCounterBytes + kChWsHdr + 2ull * H->region_words * 4ull;
  // shared memory
  uint32_t so = 0;
  H->off_bars = so; so += 1024;
  H->off_digits = so; so += (uint32_t)(kmax / 32) * 96u * (uint32_t)M; so = (so + 15u) & ~15u;
  H->off_parts = so; so += (uint32_t)(kmax / pkmin) * 8u * (uint32_t)M; so = (so + 15u) & ~15u;
  H->off_red = so; so += (uint32_t)kChWarps * (uint32_t)max_tiles * kTileN * 4u * (uint32_t)M;
  H->off_zpad = so; so += 64;
  so = (so + 255u) & ~255u;
  H->off_aux = so; so += (uint32_t)kChWarps * kAuxD * 256u;
  so = (so + 1023u) & ~1023u;
  const uint32_t budget = 226u * 1024u - 1024u;                            // 1 KB slack for the 1024-byte alignment of the base
  if (so + 16u * kSlabBytes > budget) return B200Q_ERR_UNSUPPORTED;
  int slots = (int)((budget - so) / kSlabBytes);
  if (g_ch_slots > 0 && g_ch_slots < slots) slots = g_ch_slots;
  // slab q is consumed by warp q % 16 and lives in slot q % slots: with 16 or 32 slots a slot always belongs to the same
  // warp, so no waiter is ever more than one mbarrier phase ahead
  slots = slots >= 32 ? 32 : (slots >= 16 ? 16 : 0);
  if (slots < 16) return B200Q_ERR_UNSUPPORTED;
  H->slots = (uint32_t)slots;
  H->off_ring = so; so += (uint32_t)slots * kSlabBytes;
  H->off_gs = so;
  int ncached = (int)((budget - so) / sizeof(ChGS));
  if (ncached > n_groups) ncached = n_groups;
  if (ncached < 0) ncached = 0;
  H->n_cached = (uint32_t)ncached;
  so += (uint32_t)(ncached * sizeof(ChGS));
  H->window = (uint32_t)((g_ch_window > 0 && g_ch_window <= slots) ? g_ch_window : 16);
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
  ChParams p = {};
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
