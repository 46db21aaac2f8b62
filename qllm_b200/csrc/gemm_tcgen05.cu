// Prefill / batched path: fused dequant + tcgen05 tensor-core GEMM for sm_100a.
//
// Orientation: D[n, tok] = W^T[n, k] . X^T[k, tok]  (one CTA = 128 output columns x TT tokens)
//   * the dequantised weights are the UMMA *A* operand and live in TENSOR MEMORY: the 128 dequant
//     threads map 1:1 to the 128 TMEM lanes (= output columns n).  A thread reads its column's packed
//     words from shared memory, unpacks 64 k-values in registers (lop3 + exact fp16 (q-z), one
//     rounding in *s), and writes them with one tcgen05.st.32x32b.x32 -- no shared-memory round trip
//     for the fp16 weights, which halves shared-memory traffic vs. staging them for an SS-mode MMA;
//   * X^T is the B operand, K-major, loaded by TMA with 128-byte swizzle straight from the caller's
//     [M, K] activation matrix (rows beyond M are zero-filled by TMA);
//   * accumulators (fp32) stay in TMEM; the epilogue reads them with tcgen05.ld, adds bias, converts
//     to fp16 and stores y[tok, n] (coalesced along n across the warp).
// Warp roles (896 threads): w0-23 dequant + epilogue (three teams of eight) | w24 X TMA producer | w25 packed-W TMA
// producer | w26 TMEM allocator | w27 MMA issuer (one thread).  The control warps carry the HIGHEST warp ids: the
// SM's issue arbiter favours high warp ids, and the single MMA-issuing thread is the serial resource of the CTA
// (measured: as warp 1 it was starved to ~560 cycles per k-block against a 256-cycle MMA floor).
// mbarrier rings: X stages (TMA -> MMA), packed-W stages (TMA -> dequant), A stages in TMEM (dequant -> MMA; their
// release, one tcgen05.commit per k-block, also frees the X stage), accumulator-ready.
//
// Replaces: ort_ops.dequant + cuBLAS (quant_linear_gptq.py:81-85), gemm_forward_cuda
// (gemm_cuda_gen.cu:1102-1161), marlin mul (marlin_cuda.cpp:29-74) at M > 8.
#include <cuda.h>

#include "common.cuh"
#include <mutex>

#include "kernels.h"

namespace b200q {

static constexpr int kDqWarps = 8;      // dequant warps that cooperate on one k-block (two per TMEM lane quadrant)
static constexpr int kDqPar = 3;        // k-blocks dequantised concurrently: team p takes k-blocks p, p+kDqPar, ...
static constexpr int kTcThreads = (4 + kDqWarps * kDqPar) * 32;   // 896
static constexpr int kWXProd = kDqWarps * kDqPar, kWWProd = kWXProd + 1, kWAlloc = kWXProd + 2, kWMma = kWXProd + 3;
static constexpr int kNSXMax = 8;     // X-tile stages (TMA -> MMA):      8 for TT <= 128, 5 for TT = 256
static constexpr int kNSWMax = 16;    // packed-W stages (TMA -> dequant): sized to what shared memory leaves
__host__ __device__ constexpr int tc_stages(int tt) { return tt <= 128 ? 8 : 5; }
static constexpr int kNA = 8;         // A stages in TMEM (64 k = 32 columns each)
static constexpr int kBK = 64;        // k per stage
static constexpr int kBN = 128;       // output columns per CTA (= UMMA M)
static constexpr uint32_t kSpinLimit = 4u << 20;

static constexpr uint32_t MAGIC = 0x64006400u, LO4 = 0x000f000fu, HI4 = 0x00f000f0u, H_1_16 = 0x2c002c00u;

struct TcParams {
  LayerView L;
  int M;
  PeerOut out;
  int64_t ldy, n_offset;
  const __half* residual;                   // NULL, or [M, ldres]: added to the rounded output in the epilogue (b200q_linear_ex)
  int64_t ldres;
  int out_bf16;                             // y (and residual) are bfloat16: the fp16 result is rounded to bf16 in the epilogue
  int kblocks, gshift, group32, nsx, nsw;   // nsx / nsw: X and W ring depths actually used
  int ksplit, kb_per;                       // split-K over gridDim.z (small M): k-blocks per split; partial tiles + last-arriver reduction
  int csplit, off_part;                     // cluster split-K (M > 64): 2 = the CTA pair (cluster dims 1,1,2) halves K and swaps half-tiles
                                            // of fp32 partial sums through distributed shared memory; off_part: [TT/2][128] fp32
  // sibling release (b200q_linear_group, M > 64): q/k/v or gate/up read the same x, so only the first sibling orders itself
  // behind the stream (griddepcontrol.wait) and then raises dep_flag[2 i] for sibling i; a later sibling's X producer polls
  // its flag instead, so its CTAs fill the SMs the previous sibling's last wave leaves idle.  dep_flag[1] counts the pollers;
  // the last one clears the flag (both words are back at zero when the kernel ends: CUDA-graph replays need no reset).
  unsigned int* dep_flag;
  int dep_role, dep_n;
  float* partial;                           // [ksplit][M][N] fp32 (workspace, behind the counter region)
  unsigned int* counters;                   // arrival counter per output tile (workspace head, left zeroed)
  int kbc, nchunks, gcap;                   // group tables are staged per chunk of kbc k-blocks (<= gcap groups): long-K / small-group layers
  float inv_group;   // ns: input stages actually used (<= kNSMax, sized to fit shared memory)
  int off_x, off_w, off_sc, off_zq, off_bar;
  int* err;
  unsigned long long* dbg;   // diagnostic: CTA (0,0) records per-k-block phase stamps (nullptr in production)
};

__device__ __forceinline__ unsigned long long tc_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));      // SM clock: every stamp comes from CTA (0,0), i.e. one SM
  return t;
}
// thread-block cluster pieces of the split-K pair (non-.aligned forms: the role branches reach them warp by warp)
__device__ __forceinline__ void tc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t tc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t tc_mapa(uint32_t local_smem, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem), "r"(rank));
  return r;
}
// one fp32 into the peer's shared memory, counted in bytes on the peer's mbarrier (the receiver waits on its own barrier:
// the pattern the decode kernels' cluster reduction uses)
__device__ __forceinline__ void tc_st_async_f32(uint32_t remote_addr, uint32_t remote_bar, float v) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(__float_as_uint(v)),
               "r"(remote_bar)
               : "memory");
}
// stamps (SM cycles): [kb][0..3] lead warp of the dequant team that owns kb (packed words landed, ALU done, A stage
// free, TMEM store retired + signalled), [kb][4..7] MMA thread (X landed, A landed, MMAs issued, commits issued)
#define TC_STAMP(kb, i) do { if (DBG && dbg_on) p.dbg[(size_t)(kb) * 8 + (i)] = tc_gtime(); } while (0)

// ---- small PTX wrappers -----------------------------------------------------------------------
#ifdef B200Q_BOUNDED_WAITS     // bring-up aid: every wait gives up after kSpinLimit polls and records a code
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int* err, int code) {
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++n > kSpinLimit) {
      if (err) atomicExch(err, code);
      return false;
    }
  }
  return true;
}
#else
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int*, int) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"      // %2: suspend-time hint (ns)
      "@p bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return true;
}
#endif
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// one arrival per warp, issued by lane 0 under a predicate (no divergent branch)
__device__ __forceinline__ void mbar_arrive_lane0(uint64_t* bar, int lane) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, 0;\n\t"
      "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}"
      ::"r"(smem_u32(bar)), "r"(lane)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart (version 1)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------------
// GPTQ / HQQ 4-bit producer of A: lane n of quadrant q owns column n; word r of a stage holds
// k = 8r..8r+7.  Two dequant warps per TMEM lane quadrant split the 8 words of a stage.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <int TT, int BITS, bool FZ, bool DBG>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_gptq_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap, const TcParams p) {
  extern __shared__ __align__(1024) char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SW128 tiles need 1024-byte alignment
  char* xst = smem + p.off_x;                                      // kNS x [TT][64] fp16, SW128
  uint32_t* wst = reinterpret_cast<uint32_t*>(smem + p.off_w);     // kNS x [64*BITS/32][128] words
  __half* sc = reinterpret_cast<__half*>(smem + p.off_sc);         // [G][128]
  char* zq = smem + p.off_zq;                                      // [G][128] nibbles (64 B) | [G][128] fp16
  uint64_t* full_x = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* empty_x = full_x + kNSXMax;
  uint64_t* full_w = empty_x + kNSXMax;
  uint64_t* empty_w = full_w + kNSWMax;
  uint64_t* a_full = empty_w + kNSWMax;
  uint64_t* a_empty = a_full + kNA;
  uint64_t* acc_full = a_empty + kNA;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  uint64_t* part_bar = acc_full + 2;                                // cluster split-K: counts the bytes of the peer's half-tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * kBN;
  const int tok0 = blockIdx.y * TT;
  const int kbz0 = blockIdx.z * p.kb_per;                              // this CTA's k-blocks: kbz0 + [0, nkb)
  const int nkb = min(p.kblocks, kbz0 + p.kb_per) - kbz0;
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;

  if (tid == 0) TC_STAMP(p.kblocks, 0);                                                        // kernel entry
  // Programmatic dependent launch: the next kernel of the stream may start its prologue (barriers, TMEM, group tables,
  // packed-weight TMA + dequant -- none of which depends on this kernel) on every SM this grid has left; only its
  // activation loads and its stores wait for this grid to finish (griddepcontrol.wait below).
  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < kNSXMax; ++s) { mbar_init(&full_x[s], 1); mbar_init(&empty_x[s], 1); }   // empty_x: unused
    for (int s = 0; s < kNSWMax; ++s) { mbar_init(&full_w[s], 1); mbar_init(&empty_w[s], kDqWarps); }
    for (int s = 0; s < kNA; ++s) { mbar_init(&a_full[s], kDqWarps); mbar_init(&a_empty[s], 1); }
    mbar_init(acc_full, TT <= 128 ? 2 : 1);
    mbar_init(part_bar, 1);
    fence_mbar_init();
  }
  if (warp == kWAlloc) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();                         // barriers initialised, TMEM allocated: the TMA producers start right away
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (TT >= 128 && p.csplit == 2 && tid == 0) mbar_expect_tx(part_bar, (uint32_t)((TT / 2) * kBN * 4));   // armed long before the peer sends
  if (tid == 0) TC_STAMP(p.kblocks, 1);                                                        // barriers + TMEM ready
  constexpr int P = 32 / BITS;                     // k values per packed word (3-bit: 32 values straddle three words)
  constexpr int RS = kBK * BITS / 32;              // packed rows per stage
  constexpr int WH = RS / 2;                       // words per column per half-stage (32 k)
  constexpr uint32_t X_BYTES = TT * kBK * 2, W_BYTES = RS * kBN * 4;
  constexpr int NI = (TT <= 128) ? 2 : 1;           // MMA-issuing threads (k-blocks i, i + NI, ...), one accumulator each
  constexpr int kACol = NI * TT;                   // A stages sit right after the accumulator columns

  if (warp == kWXProd) {                   // X producer: the ring only spans TMA latency + MMA lag
    if (lane == 0) {
      if (p.dep_role == 2) {               // x is not the previous kernel's output: that kernel is a sibling reading the same x
        unsigned int v;
        unsigned long long t0 = 0;
        unsigned spins = 0;
        for (;;) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.dep_flag) : "memory");
          if (v) break;
          __nanosleep(64);
          if ((++spins & 1023u) == 0) {     // bounded (2 s): a missing first sibling must not hang the GPU
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) break;
          }
        }
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicInc(p.dep_flag + 1, total - 1u) == total - 1u)
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.dep_flag), "r"(0u) : "memory");
        asm volatile("fence.proxy.async;" ::: "memory");      // the TMA loads below follow the acquire
      } else {
        pdl_wait();                        // x may be the previous kernel's output
        if (p.dep_role == 1 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
          for (int i = 0; i < p.dep_n; ++i)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.dep_flag + 2 * i), "r"(1u) : "memory");
      }
      int s = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        // X stage s was last read by the MMAs of k-block kb - nsx, whose completion is signalled on that k-block's
        // A-stage barrier (one commit per k-block releases both)
        if (kb >= p.nsx) {
          const int kp = kb - p.nsx;
          if (!mbar_wait_bounded(&a_empty[kp % kNA], (uint32_t)((kp / kNA) & 1), p.err, 1)) break;
        }
        mbar_expect_tx(&full_x[s], X_BYTES);
        tma_load_2d(xst + (size_t)s * X_BYTES, &xmap, &full_x[s], (kbz0 + kb) * kBK, tok0);
        if (++s == p.nsx) s = 0;
      }
    }
  } else if (warp == kWWProd) {                  // packed-W producer: runs far ahead (HBM latency + dequant + A ring)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        if (!mbar_wait_bounded(&empty_w[s], ph ^ 1u, p.err, 7)) break;
        mbar_expect_tx(&full_w[s], W_BYTES);
        tma_load_2d(wst + (size_t)s * (W_BYTES / 4), &wmap, &full_w[s], n0, (kbz0 + kb) * RS);
        if (++s == p.nsw) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kWMma || (NI == 2 && warp == kWAlloc)) {
    // MMA issuer(s).  The tensor pipe accepts roughly one tcgen05.mma at a time, so a single issuing thread leaves it
    // idle for the ~300 cycles it spends on the two mbarrier waits and the commit of each k-block (measured: 504 cycles
    // per k-block against the 256-cycle MMA time at TT = 128).  Two threads in two warps alternate k-blocks, each into
    // its own accumulator (summed in the epilogue), so one thread's waits hide under the other's MMAs.
    if (lane == 0) {
      const int me = (NI == 2 && warp == kWAlloc) ? 1 : 0;
      // instruction descriptor: D=f32, A=B=f16, both K-major, N=TT, M=128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(TT >> 3) << 17) | ((uint32_t)(kBN >> 4) << 24);
      const uint64_t bdesc0 = umma_desc_k_sw128(smem_u32(xst));
      const uint32_t d_tmem = tmem + me * TT;
      bool ok = true;
      // ring cursors advance NI stages per iteration (no runtime division on the issuing thread)
      int s = me % p.nsx, sa = me;
      uint32_t ph = (uint32_t)((me / p.nsx) & 1), pha = 0;
      for (int kb = me; kb < nkb && ok; kb += NI) {
        ok = mbar_wait_bounded(&full_x[s], ph, p.err, 2);
        TC_STAMP(kb, 4);
        ok = ok && mbar_wait_bounded(&a_full[sa], pha, p.err, 3);
        TC_STAMP(kb, 5);
        tc_fence_after();
        const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)s * (X_BYTES >> 4));
        const uint32_t a_src = tmem + kACol + sa * 32;
#pragma unroll
        for (int j = 0; j < kBK / 16; ++j)
          tc_mma_ts(d_tmem, a_src + j * 8, bdesc + (uint64_t)(2 * j), idesc, (kb >= NI || j != 0) ? 1u : 0u);
        TC_STAMP(kb, 6);
        tc_commit(&a_empty[sa]);                 // frees A stage sa and X stage s
        TC_STAMP(kb, 7);
        s += NI;
        if (s >= p.nsx) { s -= p.nsx; ph ^= 1u; }
        sa += NI;
        if (sa >= kNA) { sa -= kNA; pha ^= 1u; }
      }
      tc_commit(acc_full);
    }
  } else if (warp < kDqWarps * kDqPar) {
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access (warp id % 4)
    const int half = (warp >> 2) & 1;           // which half (32 k) of a stage
    const int par = warp >> 3;                  // team: k-blocks par, par + kDqPar, ...
    const int eidx = warp >> 2;                 // epilogue slot 0 .. 2*kDqPar-1 (token chunks eidx, eidx + 2*kDqPar, ...)
    const int n = q * 32 + lane;                // column within the CTA tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int gcur = -1;
    uint32_t c_lo = 0, c_hi = 0, s2 = 0, z2 = 0;
    bool ok = true;
    const uint32_t* ws_lane = wst + (WH * half) * kBN + n;
    const uint32_t a_dst = tmem + lane_addr + kACol + half * 16;
    constexpr uint32_t ZMASK = (1u << BITS) - 1u;
    const int zbit = n * BITS;
    const uint32_t* zq_lane = reinterpret_cast<const uint32_t*>(zq) + (zbit >> 5);
    // group constants of this lane's column: s2 = (s,s); integer zeros folded into the magic constants,
    // (h - (1024+z)) * s  [h = 1024+q];  float zeros (HQQ): ((h - 1024) - z) * s
    int gbase = 0;                              // first group held by the shared-memory tables (current chunk)
    auto load_group = [&](int gabs) {
      const int gi = gabs - gbase;
      s2 = dup_half(sc[gi * kBN + n]);
      if (FZ) {
        z2 = dup_half(reinterpret_cast<const __half*>(zq)[gi * kBN + n]);
        c_lo = MAGIC; c_hi = 0xD400D400u;
      } else {
        uint32_t zraw = zq_lane[gi * (kBN * BITS / 32)] >> (zbit & 31);
        if (BITS == 3 && (zbit & 31) > 29) zraw |= zq_lane[gi * (kBN * BITS / 32) + 1] << (32 - (zbit & 31));   // field straddles two words
        const uint32_t z = ((zraw & ZMASK) + (uint32_t)p.L.zero_bias) & ZMASK;
        c_lo = (0x6400u | z) * 0x00010001u;
        c_hi = (0xD400u + (z << 4)) * 0x00010001u;
      }
    };
    auto fin_lo = [&](uint32_t h) { uint32_t d = hsub2_u(h, c_lo); if (FZ) d = hsub2_u(d, z2); return hmul2_u(d, s2); };
    auto fin_hi = [&](uint32_t h) { uint32_t d = hfma2_u(h, H_1_16, c_hi); if (FZ) d = hsub2_u(d, z2); return hmul2_u(d, s2); };
    // one packed word -> 16/BITS... fp16 pairs in MMA k order, written to a[r * (P/2) ...]
    auto dq_word = [&](uint32_t wv, uint32_t* o) {
      if (BITS == 4) {
        const uint32_t hi = wv >> 8;
        const uint32_t p0 = fin_lo(and_or(wv, LO4, MAGIC));             // (k0,k4)
        const uint32_t p1 = fin_hi(and_or(wv, HI4, MAGIC));             // (k1,k5)
        const uint32_t p2 = fin_lo(and_or(hi, LO4, MAGIC));             // (k2,k6)
        const uint32_t p3 = fin_hi(and_or(hi, HI4, MAGIC));             // (k3,k7)
        o[0] = prmt(p0, p1, 0x5410);    // (k0,k1)
        o[1] = prmt(p2, p3, 0x5410);    // (k2,k3)
        o[2] = prmt(p0, p1, 0x7632);    // (k4,k5)
        o[3] = prmt(p2, p3, 0x7632);    // (k6,k7)
      } else if (BITS == 8) {
        o[0] = fin_lo(prmt(wv, MAGIC, 0x5150));                         // (k0,k1): bytes -> 1024+q
        o[1] = fin_lo(prmt(wv, MAGIC, 0x5352));                         // (k2,k3)
      } else {                                                           // 2-bit: E_i = (k_i, k_{i+8})
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = fin_lo(and_or(wv >> (2 * i), 0x00030003u, MAGIC));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = prmt(e[2 * j], e[2 * j + 1], 0x5410);                   // (k_2j, k_2j+1)
          o[4 + j] = prmt(e[2 * j], e[2 * j + 1], 0x7632);               // (k_8+2j, k_9+2j)
        }
      }
    };
    auto group_at = [&](int k) { return p.gshift >= 0 ? (k >> p.gshift) : (int)(((float)k + 0.5f) * p.inv_group); };
    const int kNS = p.nsw;
    int s = par % kNS;
    uint32_t ph = (uint32_t)((par / kNS) & 1);
    constexpr int kDqThreads = kDqWarps * kDqPar * 32;
    const int nchunks = (nkb + p.kbc - 1) / p.kbc;
    for (int c = 0; c < nchunks; ++c) {
    // ---- group tables of this chunk's k-blocks for the CTA's 128 columns: loaded by the dequant warps only, so the first
    // chunk's load latency overlaps the first TMA round trips; later chunks cost one pipeline bubble each ----
    const int kb0 = c * p.kbc, kb1 = min(nkb, kb0 + p.kbc);
    {
      gbase = ((kbz0 + kb0) * kBK) / p.L.group;
      const int gcount = (((kbz0 + kb1) * kBK - 1) / p.L.group) - gbase + 1;
      if (c > 0) asm volatile("bar.sync 8, %0;" ::"n"(kDqThreads) : "memory");       // every warp is done with the previous tables
      for (int idx = tid; idx < gcount * kBN; idx += kDqThreads) {
        const int g = gbase + idx / kBN, nn = idx % kBN;
        sc[idx] = (n0 + nn < p.L.N) ? __ldg(p.L.s + (size_t)g * p.L.N + n0 + nn) : __float2half(0.f);
      }
      if (FZ) {
        for (int idx = tid; idx < gcount * kBN; idx += kDqThreads) {
          const int g = gbase + idx / kBN, nn = idx % kBN;
          reinterpret_cast<__half*>(zq)[idx] =
              (n0 + nn < p.L.N) ? __ldg(reinterpret_cast<const __half*>(p.L.qz) + (size_t)g * p.L.N + n0 + nn) : __float2half(0.f);
        }
      } else {
        constexpr int ZW = kBN * BITS / 32;                                 // packed zero words of this tile per group
        const size_t zrow = ((size_t)p.L.N * BITS) >> 5;
        for (int idx = tid; idx < gcount * ZW; idx += kDqThreads) {
          const int g = gbase + idx / ZW, wv = idx % ZW;
          reinterpret_cast<uint32_t*>(zq)[idx] =
              ((n0 * BITS) / 32 + wv < (int)zrow) ? __ldg(reinterpret_cast<const uint32_t*>(p.L.qz) + (size_t)g * zrow + (n0 * BITS) / 32 + wv) : 0u;
        }
      }
      asm volatile("bar.sync 8, %0;" ::"n"(kDqThreads) : "memory");
      if (tid == 0 && c == 0) TC_STAMP(p.kblocks, 2);                                  // first group tables in shared memory
      gcur = -1;
    }
    for (int kb = kb0 + ((par - kb0 % kDqPar) + kDqPar) % kDqPar; kb < kb1; kb += kDqPar) {
      const int sa = kb % kNA;
      mbar_wait_bounded(&full_w[s], ph, p.err, 4);
      if ((warp & 7) == 0 && lane == 0) TC_STAMP(kb, 0);
      uint32_t w[WH];
      const uint32_t* ws = ws_lane + (size_t)s * (W_BYTES / 4);
#pragma unroll
      for (int r = 0; r < WH; ++r) w[r] = ws[r * kBN];
      __syncwarp();
      mbar_arrive_lane0(&empty_w[s], lane);
      uint32_t a[16];
      const int k0 = (kbz0 + kb) * kBK + 32 * half;
      if (BITS == 3) {
        // 32 values in three words, LSB-first bit-stream (compress_weight.py:27-43): pair j = (k_2j, k_2j+1) starts at bit 6 j;
        // (v & 7) | ((v << 13) & 0x70000) | 0x64006400 = the fp16 pair (1024 + q_2j, 1024 + q_2j+1); group % 32 == 0
        const int gi = group_at(k0);
        if (gi != gcur) { gcur = gi; load_group(gi); }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int b = 6 * j, wi = b >> 5, o = b & 31;
          const uint32_t v = (o > 26) ? __funnelshift_r(w[wi], w[wi < WH - 1 ? wi + 1 : wi], o) : (w[wi] >> o);
          a[j] = fin_lo((v & 7u) | ((v << 13) & 0x00070000u) | MAGIC);
        }
      } else if (p.group32) {                            // the 32 k of this half-stage share one group (warp-uniform)
        const int gi = group_at(k0);
        if (gi != gcur) { gcur = gi; load_group(gi); }
#pragma unroll
        for (int r = 0; r < WH; ++r) dq_word(w[r], a + r * (P / 2));
      } else {                                           // small groups: resolve per packed word
#pragma unroll
        for (int r = 0; r < WH; ++r) {
          const int gi = group_at(k0 + P * r);
          if (gi != gcur) { gcur = gi; load_group(gi); }
          dq_word(w[r], a + r * (P / 2));
        }
      }
      if ((warp & 7) == 0 && lane == 0) TC_STAMP(kb, 1);
      mbar_wait_bounded(&a_empty[sa], ((kb / kNA) + 1) & 1, p.err, 5);
      if ((warp & 7) == 0 && lane == 0) TC_STAMP(kb, 2);
      tc_fence_after();
      tc_st16(a_dst + sa * 32, a);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      mbar_arrive_lane0(&a_full[sa], lane);
      if ((warp & 7) == 0 && lane == 0) TC_STAMP(kb, 3);
      s += kDqPar;
      while (s >= kNS) { s -= kNS; ph ^= 1u; }
    }
    }
    // ---- epilogue: TMEM -> registers -> fp16 -> shared (transpose) -> 16-byte coalesced stores ----
    ok = __all_sync(0xffffffffu, ok && mbar_wait_bounded(acc_full, 0, p.err, 6));
    if (tid == 0) TC_STAMP(p.kblocks, 3);                                                      // accumulators complete
    pdl_wait();                                     // stores follow the previous kernel's last reads (returns at once by now)
    tc_fence_after();
    const float bias = (p.L.bias && (n0 + n) < p.L.N) ? __half2float(__ldg(p.L.bias + n0 + n)) : 0.f;
    __half* stg = reinterpret_cast<__half*>(xst) + (size_t)eidx * 32 * (kBN + 8);     // [32 tok][128+8 n], X stages are free now
    const int tih = q * 32 + lane;                                                     // thread index within this slot's 4 warps
    // accumulator chunk c0 (32 tokens x this lane's column), both issuers' accumulators summed
    auto load_acc = [&](int c0, uint32_t (&v)[32]) {
      tc_ld32(tmem + lane_addr + c0, v);
      if (NI == 2 && nkb > 1) {                   // the second issuer's accumulator exists only if it had a k-block
        uint32_t v2[32];
        tc_ld32(tmem + lane_addr + TT + c0, v2);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
      } else {
        tc_wait_ld();
      }
    };
    // Cluster split-K: the two CTAs of a pair hold partial sums of the same 128 x TT tile over the two halves of K.  Each
    // finalises one half of the tokens: it sends the other half of its partial tile into the peer's (now idle) stage memory
    // through distributed shared memory and adds what the peer sent to its own accumulators in the loop below.
    int c_begin = 0, c_end = TT;
    const float* part = reinterpret_cast<const float*>(smem + p.off_part);             // [TT/2][128] fp32, written by the peer
    if (TT >= 128 && p.csplit == 2) {
      const uint32_t crank = tc_cluster_rank();
      tc_cluster_sync();                          // both CTAs' MMAs have completed: stage memory is free on either side
      const uint32_t remote = tc_mapa(smem_u32(part), crank ^ 1u), remote_bar = tc_mapa(smem_u32(part_bar), crank ^ 1u);
      const int pbase = (int)(crank ^ 1u) * (TT / 2);
#pragma unroll 1
      for (int j = eidx; j < TT / 64; j += 2 * kDqPar) {
        uint32_t v[32];
        load_acc(pbase + 32 * j, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) tc_st_async_f32(remote + (uint32_t)(((32 * j + i) * kBN + n) * 4), remote_bar, __uint_as_float(v[i]));
      }
      c_begin = (int)crank * (TT / 2);
      c_end = c_begin + TT / 2;
      // the peer's half-tile has landed when its bytes are counted on this CTA's barrier.  Every epilogue thread waits (not
      // only the finalising ones): the CTA must not leave while stores into its shared memory are in flight
      mbar_wait_bounded(part_bar, 0, p.err, 8);
    }
#pragma unroll 1
    for (int c0 = c_begin + eidx * 32; c0 < c_end && ok; c0 += 64 * kDqPar) {
      uint32_t v[32];
      load_acc(c0, v);
      if (TT >= 128 && p.csplit == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + part[(c0 - c_begin + i) * kBN + n]);
      }
      if (p.ksplit > 1) {                         // split-K: fp32 partial tile of this split, coalesced along n
        float* slab = p.partial + (size_t)blockIdx.z * (size_t)p.M * p.L.N;
        if (n0 + n < p.L.N) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (tok0 + c0 + i < p.M) slab[(size_t)(tok0 + c0 + i) * p.L.N + n0 + n] = __uint_as_float(v[i]);
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        __half h = __float2half_rn(__uint_as_float(v[i]) + bias);
        if (p.out_bf16) h = __ushort_as_half((unsigned short)float_to_bf16_bits(__half2float(h)));   // bf16 caller: out.to(x.dtype)
        stg[(size_t)i * (kBN + 8) + n] = h;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + eidx) : "memory");
      // 32 token rows x 256 B: 16 chunks of 16 B per row, 128 threads -> 4 chunks each
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = tih + 128 * j, row = idx >> 4, ch = idx & 15;
        const int tok = tok0 + c0 + row;
        if (tok < p.M && n0 + 8 * ch < p.L.N) {
          uint4 val = *reinterpret_cast<const uint4*>(stg + (size_t)row * (kBN + 8) + 8 * ch);
          if (p.residual) {                         // y = fp16(fp16(acc + bias) + residual), as the unfused fp16 add rounds
            const uint4 rv = __ldg(reinterpret_cast<const uint4*>(p.residual + (size_t)tok * p.ldres + n0 + 8 * ch));
            if (p.out_bf16) {                       // bf16 model: the skip connection is a bf16 add
              const uint32_t* a1 = reinterpret_cast<const uint32_t*>(&val);
              const uint32_t* r1 = reinterpret_cast<const uint32_t*>(&rv);
              uint32_t o1[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t lo = float_to_bf16_bits(bf16_bits_to_float(a1[e] & 0xffffu) + bf16_bits_to_float(r1[e] & 0xffffu));
                const uint32_t hi = float_to_bf16_bits(bf16_bits_to_float(a1[e] >> 16) + bf16_bits_to_float(r1[e] >> 16));
                o1[e] = lo | (hi << 16);
              }
              val = *reinterpret_cast<const uint4*>(o1);
            } else {
            const __half2* a2 = reinterpret_cast<const __half2*>(&val);
            const __half2* r2 = reinterpret_cast<const __half2*>(&rv);
            __half2 o2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fa = __half22float2(a2[e]), fr = __half22float2(r2[e]);
              o2[e] = __floats2half2_rn(fa.x + fr.x, fa.y + fr.y);
            }
            val = *reinterpret_cast<const uint4*>(o2);
            }
          }
          for (int qd = 0; qd < p.out.n; ++qd)
            *reinterpret_cast<uint4*>(p.out.y[qd] + (size_t)tok * p.ldy + p.n_offset + n0 + 8 * ch) = val;
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + eidx) : "memory");
    }
    if (p.ksplit > 1) {
      // last-arriving split of this output tile sums the slabs in split order (deterministic), adds bias, stores fp16
      constexpr int kDqT = kDqWarps * kDqPar * 32;
      __threadfence();
      asm volatile("bar.sync 8, %0;" ::"n"(kDqT) : "memory");
      unsigned int* cnt = p.counters + (blockIdx.y * gridDim.x + blockIdx.x);
      if (tid == 0) {
        const unsigned int old = atomicAdd(cnt, 1u);
        tmem_slot[1] = (old == (unsigned int)p.ksplit - 1u) ? 1u : 0u;
        if (old == (unsigned int)p.ksplit - 1u) *cnt = 0u;
      }
      asm volatile("bar.sync 8, %0;" ::"n"(kDqT) : "memory");
      if (tmem_slot[1]) {
        __threadfence();
        const int rows = min(TT, p.M - tok0);
        // nvcc hoists a __ldg above its null check (seen as an unpredicated LDG.CONSTANT): read through a pointer that is
        // always valid (row 0 of the scales when there is no bias) and select afterwards
        const bool has_bias = p.L.bias != nullptr;
        const __half* bsrc = has_bias ? p.L.bias : p.L.s;
        // four columns per thread, every split's 16-byte load in flight before the first add (the slabs come from L2; a
        // dependent scalar loop here cost more than the main loop at small M); summed in split order -> deterministic
        for (int idx = tid; idx < rows * (kBN / 4); idx += kDqT) {
          const int tok = tok0 + idx / (kBN / 4), nn = n0 + 4 * (idx % (kBN / 4));
          if (nn < p.L.N) {                       // N % 8 == 0: the four columns are in range together
            float4 v[8];
#pragma unroll
            for (int z = 0; z < 8; ++z)
              if (z < p.ksplit) v[z] = __ldcg(reinterpret_cast<const float4*>(p.partial + ((size_t)z * p.M + tok) * p.L.N + nn));
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int z = 0; z < 8; ++z)
              if (z < p.ksplit) { acc[0] += v[z].x; acc[1] += v[z].y; acc[2] += v[z].z; acc[3] += v[z].w; }
            __half h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float bv = __half2float(bsrc[nn + e]);
              h[e] = __float2half_rn(has_bias ? acc[e] + bv : acc[e]);
              if (p.out_bf16) {
                float v = round_bf16(__half2float(h[e]));
                if (p.residual) v = round_bf16(v + bf16_bits_to_float(reinterpret_cast<const unsigned short*>(p.residual)[(size_t)tok * p.ldres + nn + e]));
                h[e] = __ushort_as_half((unsigned short)float_to_bf16_bits(v));
              } else if (p.residual) {
                h[e] = __float2half_rn(__half2float(h[e]) + __half2float(p.residual[(size_t)tok * p.ldres + nn + e]));
              }
            }
            const uint2 pk = make_uint2((uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
                                        (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
            for (int qd = 0; qd < p.out.n; ++qd)
              *reinterpret_cast<uint2*>(p.out.y[qd] + (size_t)tok * p.ldy + p.n_offset + nn) = pk;
          }
        }
      }
    }
    tc_fence_before();
  }
  if (TT >= 128 && p.csplit == 2 && warp >= kDqWarps * kDqPar) tc_cluster_sync();   // producer / issuer warps: the pair's barrier counts every thread
  __syncthreads();
  if (tid == 0) TC_STAMP(p.kblocks, 4);                                                        // epilogue stored
  if (warp == kWAlloc) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static unsigned long long* g_tc_dbg = nullptr;
void gemm_tc_set_debug(unsigned long long* buf) { g_tc_dbg = buf; }
// bring-up aid (B200Q_BOUNDED_WAITS builds only): a CALLER-provided device int that receives the code of a wait that
// gave up.  The library itself never allocates (b200q.h): without a flag the kernels run with err == nullptr.
static int* g_err_flag = nullptr;
void gemm_tc_set_error_flag(int* device_flag) { g_err_flag = device_flag; }

// TMA descriptors are a pure function of (base pointer, shape, strides, box, swizzle): encoded once, then served from a
// small direct-mapped cache (cuTensorMapEncodeTiled costs ~1-2 us of host time per call, twice per GEMM launch).
struct TmapKey {
  const void* base;
  uint64_t d0, d1, stride;
  uint32_t b0, b1, dtype, swizzle;
  bool operator==(const TmapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && stride == o.stride && b0 == o.b0 && b1 == o.b1 && dtype == o.dtype && swizzle == o.swizzle;
  }
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool valid; };
static constexpr int kTmapSlots = 2048;
static TmapSlot g_tmaps[kTmapSlots];
static std::mutex g_tmap_mu;
static bool cached_tmap_2d(EncodeTiledFn enc, CUtensorMap* out, CUtensorMapDataType dtype, const void* base, uint64_t d0, uint64_t d1,
                           uint64_t stride_bytes, uint32_t b0, uint32_t b1, CUtensorMapSwizzle swz) {
  const TmapKey key{base, d0, d1, stride_bytes, b0, b1, (uint32_t)dtype, (uint32_t)swz};
  uint64_t h = (uint64_t)(uintptr_t)base * 0x9E3779B97F4A7C15ull;
  h ^= (d0 * 0xC2B2AE3D27D4EB4Full) ^ (d1 * 0x165667B19E3779F9ull) ^ (stride_bytes << 7) ^ ((uint64_t)b0 << 40) ^ ((uint64_t)b1 << 20) ^ swz;
  TmapSlot& sl = g_tmaps[(h >> 17) % kTmapSlots];
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (sl.valid && sl.key == key) { *out = sl.map; return true; }
  }
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
  cuuint32_t es[2] = {1, 1};
  if (enc(out, dtype, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  sl.key = key; sl.map = *out; sl.valid = true;
  return true;
}

static int g_tc_pdl = 1;                  // B200Q_GEMM_PDL=0 / option "gemm_pdl": plain stream-ordered launches
void gemm_tc_set_pdl(int on) { g_tc_pdl = on; }
static int g_tc_splitk = 1;               // B200Q_GEMM_SPLITK=0 / option "gemm_splitk": no K split (workspace form at M <= 64, CTA pairs above)
static int g_tc_tt256_min_m = 0;           // option "tt256_min_m" / B200Q_TT256_MIN_M: > 0 forces 256-token tiles from that M on
void gemm_tc_set_tt256_min_m(int m) { g_tc_tt256_min_m = m; }
static int g_tc_force_tt = 0, g_tc_force_ks = 0;      // diagnostic: options "gemm_force_tt" / "gemm_force_ksplit" (0 = automatic)
void gemm_tc_set_force(int which, int v) { (which == 0 ? g_tc_force_tt : g_tc_force_ks) = v; }

// Tile choice for M > 64: token-tile width TT (each packed weight is dequantised once per token tile, so 256 halves the
// dequant work that bounds the 128-wide tile) and cs = 2 to halve K across a CTA pair when the grid would leave SMs idle.
// Modelled time = waves x (k-blocks per CTA x time per k-block + fixed); the constants are fitted to the B200 measurements
// of tools/sweep_gemm_tiles.py (profiles/r2_sweep_gemm_tiles.txt): 0.27 us per k-block at TT = 128 (dequant issue-bound),
// 0.37 us at TT = 256 with the chip full (tensor pipe at power-capped clocks; 0.31 us when at most half the SMs run),
// 5 us fixed per launch (prologue + epilogue as they overlap in a chain), 3.3 us more for the pair's exchange.
struct TcTile { int tt, cs; };
static TcTile tc_pick(const LayerView& L, int64_t M) {
  TcTile best = {M <= 32 ? 32 : (M <= 64 ? 64 : 128), 1};
  if (g_tc_force_tt == 32 || g_tc_force_tt == 64 || g_tc_force_tt == 128 || g_tc_force_tt == 256) best.tt = g_tc_force_tt;
  const int kblocks = L.K / kBK;
  const int64_t tiles = (L.N + kBN - 1) / kBN;
  if (M <= 64) return best;
  if (g_tc_force_tt || g_tc_force_ks) {
    if (!g_tc_force_tt) best.tt = (g_tc_tt256_min_m > 0 && M >= g_tc_tt256_min_m) ? 256 : 128;
    best.cs = (g_tc_force_ks == 2 && best.tt >= 128 && kblocks >= 16) ? 2 : 1;
    return best;
  }
  double best_t = 1e30;
  for (int tt = 128; tt <= 256; tt *= 2) {
    if (g_tc_tt256_min_m > 0 && (tt == 256) != (M >= g_tc_tt256_min_m)) continue;
    for (int cs = 1; cs <= 2; ++cs) {
      if (cs == 2 && (!g_tc_splitk || kblocks < 16)) continue;
      const int64_t ctas = tiles * ((M + tt - 1) / tt) * cs;
      const int slots = cs == 2 ? 144 : 148;                        // CTA pairs may not cover every GPC's last SM
      // (whole waves also for the siblings of a b200q_linear_group call: choosing by SM time there, on the idea that the
      // neighbours fill every partial wave, measured 894 against 958 TFLOP/s on the 7B block -- a dependent launch only
      // starts once the previous grid's last wave is running)
      const double waves = (double)((ctas + slots - 1) / slots);
      const double per_kb = tt == 128 ? 0.27 : (ctas <= 74 ? 0.31 : 0.37);
      const double t = waves * ((double)((kblocks + cs - 1) / cs) * per_kb + 5.0 + (cs == 2 ? 3.3 : 0.0));
      if (t < best_t * 0.97) { best_t = t; best.tt = tt; best.cs = cs; }   // ties go to the simpler configuration (tried first)
    }
  }
  return best;
}
static int pick_tt(const LayerView& L, int64_t M) { return tc_pick(L, M).tt; }

bool gemm_tc_supported(const LayerView& L, int64_t M, const __half* x, int64_t ldx) {
  if (!(L.layout == B200Q_LAYOUT_GPTQ || L.layout == B200Q_LAYOUT_HQQ) || L.g_idx || L.x_perm) return false;
  if (L.bits != 2 && L.bits != 3 && L.bits != 4 && L.bits != 8) return false;
  if (L.bits == 3 ? (L.group % 32 != 0 || L.N % 32 != 0) : (L.group % (32 / L.bits) != 0)) return false;   // a packed word (3-bit: a 32-value pack) never straddles two groups
  if (L.K % kBK != 0 || L.N % 8 != 0 || L.group % 8 != 0 || L.K % L.group != 0) return false;
  if (((uintptr_t)x & 15) != 0 || (ldx % 8) != 0 || ((uintptr_t)L.qw & 15) != 0 || (L.N % 8) != 0) return false;
  return get_encode() != nullptr && M >= 1;
}

// Split-K for small M (one token tile): a 128-column CTA is bound by its own dequant rate (~8192 weights per ~400
// cycles), so N / 128 CTAs leave most SMs idle; splitting K fills them (K = 14336, N = 4096, M = 16: 67 -> ~17 us).
void gemm_tc_set_splitk(int on) { g_tc_splitk = on; }
static int tc_ksplit(const LayerView& L, int64_t M) {
  const int tiles = (L.N + kBN - 1) / kBN, kblocks = L.K / kBK;
  if (!g_tc_splitk || M > 64) return 1;
  int ks = 148 / tiles;
  if (ks > 8) ks = 8;
  if (ks > kblocks / 8) ks = kblocks / 8;
  while (ks > 1 && (ks - 1) * ((kblocks + ks - 1) / ks) >= kblocks) --ks;
  if (tiles > (int)(kCounterBytes / 4)) return 1;
  return ks < 2 ? 1 : ks;
}
size_t gemm_tc_workspace(const LayerView& L, int64_t M) {
  const int ks = tc_ksplit(L, M);
  return ks > 1 ? kCounterBytes + (size_t)ks * (size_t)M * (size_t)L.N * sizeof(float) : 0;
}

template <int TT, int BITS, bool FZ, bool DBG = false>
static cudaError_t tc_launch(const LinearArgs& a, const PeerOut* peers) {
  const LayerView& L = a.L;
  EncodeTiledFn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  CUtensorMap xmap, wmap;
  if (!cached_tmap_2d(enc, &xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a.x, (uint64_t)L.K, (uint64_t)a.M, (uint64_t)a.ldx * 2, (uint32_t)kBK,
                      (uint32_t)TT, CU_TENSOR_MAP_SWIZZLE_128B))
    return cudaErrorInvalidValue;
  if (!cached_tmap_2d(enc, &wmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, L.qw, (uint64_t)L.N, (uint64_t)(L.K * BITS / 32), (uint64_t)L.N * 4,
                      (uint32_t)kBN, (uint32_t)(kBK * BITS / 32), CU_TENSOR_MAP_SWIZZLE_NONE))
    return cudaErrorInvalidValue;
  TcParams p = {};
  p.L = L; p.M = a.M;
  if (peers) p.out = *peers; else { p.out.n = 1; p.out.y[0] = a.y; }
  p.ldy = a.ldy; p.n_offset = a.n_offset;
  p.residual = a.residual; p.ldres = a.ldres; p.out_bf16 = a.act_bf16;
  if (a.residual && (((uintptr_t)a.residual & 15) != 0 || (a.ldres % 8) != 0)) return cudaErrorInvalidValue;
  p.kblocks = L.K / kBK;
  p.ksplit = tc_ksplit(L, a.M);
  p.csplit = (TT >= 128 && !peers) ? tc_pick(L, a.M).cs : 1;
  p.kb_per = (p.kblocks + p.ksplit * p.csplit - 1) / (p.ksplit * p.csplit);
  p.partial = nullptr; p.counters = nullptr;
  if (p.ksplit > 1) {
    if (!a.workspace || a.workspace_bytes < gemm_tc_workspace(L, a.M)) return cudaErrorInvalidValue;
    p.counters = reinterpret_cast<unsigned int*>(a.workspace);
    p.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(a.workspace) + kCounterBytes);
  }
  p.dep_role = 0; p.dep_n = 0; p.dep_flag = nullptr;
  if (a.sib_role && p.ksplit == 1 && a.workspace && a.workspace_bytes >= kCounterBytes && a.sib_count >= 1 && a.sib_count <= 2) {
    p.dep_role = a.sib_role; p.dep_n = a.sib_count;
    p.dep_flag = reinterpret_cast<unsigned int*>(a.workspace) + 1008 + (a.sib_role == 2 ? 2 * a.sib_index : 0);
  } else if (a.sib_role) {
    return cudaErrorInvalidValue;          // the caller checked all of this: a half-armed sibling chain must not launch
  }
  p.gshift = -1;
  p.group32 = (L.group % 32 == 0) ? 1 : 0;
  if ((L.group & (L.group - 1)) == 0) { int sh = 0; while ((1 << sh) < L.group) ++sh; p.gshift = sh; }
  p.inv_group = 1.0f / (float)L.group;
  p.err = g_err_flag;
  p.dbg = g_tc_dbg;
  const int zq_row = (L.layout == B200Q_LAYOUT_HQQ) ? kBN * 2 : kBN * BITS / 8;
  // group tables: all groups when they fit in 56 KB, else per chunk of kbc k-blocks
  const int max_groups = (56 * 1024) / (kBN * 2 + zq_row);
  if (L.G <= max_groups) { p.kbc = p.kblocks; p.nchunks = 1; p.gcap = L.G; }
  else {
    p.kbc = (int)(((long long)(max_groups - 1) * L.group) / kBK);
    if (p.kbc < 1) return cudaErrorInvalidValue;
    p.nchunks = (p.kblocks + p.kbc - 1) / p.kbc;
    p.gcap = (p.kbc * kBK + L.group - 1) / L.group + 1;
  }
  int off = 0;
  const int stage_bytes = TT * kBK * 2 + (kBK * BITS / 32) * kBN * 4;
  const int fixed_bytes = p.gcap * (kBN * 2 + zq_row) + 64 + 1024 + 1024;
  const int xb = TT * kBK * 2, wb = (kBK * BITS / 32) * kBN * 4, budget = 220 * 1024 - fixed_bytes;
  int nsx = tc_stages(TT), nsw = kNSWMax;
  while (nsw > 6 && nsx * xb + nsw * wb > budget) --nsw;
  while (nsx > 2 && nsx * xb + nsw * wb > budget) --nsx;
  if (nsx * xb + nsw * wb > budget) return cudaErrorInvalidValue;
  (void)stage_bytes;
  p.nsx = nsx; p.nsw = nsw;
  p.off_x = off; off += nsx * xb;
  p.off_w = off; off += nsw * wb;
  p.off_sc = off; off += p.gcap * kBN * 2;
  off = (off + 15) & ~15;
  p.off_zq = off; off += p.gcap * zq_row;
  off = (off + 15) & ~15;
  p.off_bar = off; off += 1024;
  // cluster split-K: the received half-tile of partial sums sits behind the epilogue's six transpose slots, inside the (by
  // then idle) X / W stage memory
  p.off_part = p.off_x + ((2 * kDqPar * 32 * (kBN + 8) * 2 + 1023) & ~1023);
  if (p.csplit == 2 && p.off_part + (TT / 2) * kBN * 4 > p.off_sc) { p.csplit = 1; p.kb_per = p.kblocks; }
  const int smem_bytes = off + 1024;   // slack for 1024-byte alignment of the dynamic window
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_gptq_kernel<TT, BITS, FZ, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  for (int i = 0; i < p.out.n; ++i)
    if (((uintptr_t)p.out.y[i] & 15) != 0) return cudaErrorInvalidValue;
  if ((a.ldy % 8) != 0 || (a.n_offset % 8) != 0) return cudaErrorInvalidValue;     // 16-byte epilogue stores
  dim3 grid((L.N + kBN - 1) / kBN, (a.M + TT - 1) / TT, p.ksplit * p.csplit);
  count_launch();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = a.stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = g_tc_pdl ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = 1; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 2;
  cfg.attrs = at;
  cfg.numAttrs = p.csplit == 2 ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_tc_gptq_kernel<TT, BITS, FZ, DBG>, xmap, wmap, p);
}

cudaError_t launch_gemm_tc(const LinearArgs& a, const PeerOut* peers) {
  // M is chunked so that a.M fits int and grid.y <= 65535
  const bool fz = a.L.layout == B200Q_LAYOUT_HQQ;
  // diagnostic timeline build: only the TT=128 4-bit integer-zero instantiation carries the stamps
  if (g_tc_dbg && !fz && a.L.bits == 4 && pick_tt(a.L, a.M) == 128) return tc_launch<128, 4, false, true>(a, peers);
#define B200Q_TC_DISPATCH(BITS)                                  \
  switch (pick_tt(a.L, a.M)) {                                       \
    case 32: return (fz ? tc_launch<32, BITS, true>(a, peers) : tc_launch<32, BITS, false>(a, peers));              \
    case 64: return (fz ? tc_launch<64, BITS, true>(a, peers) : tc_launch<64, BITS, false>(a, peers));              \
    case 256: return (fz ? tc_launch<256, BITS, true>(a, peers) : tc_launch<256, BITS, false>(a, peers));            \
    default: return (fz ? tc_launch<128, BITS, true>(a, peers) : tc_launch<128, BITS, false>(a, peers));             \
  }
  if (a.L.bits == 2) { B200Q_TC_DISPATCH(2) }
  if (a.L.bits == 3) { B200Q_TC_DISPATCH(3) }
  if (a.L.bits == 8) { B200Q_TC_DISPATCH(8) }
  B200Q_TC_DISPATCH(4)
#undef B200Q_TC_DISPATCH
}

}  // namespace b200q
