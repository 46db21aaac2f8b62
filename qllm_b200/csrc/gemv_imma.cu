// Decode kernel for K-packed 4-bit layers (GPTQ / HQQ checkpoints in place; AWQ-GEMM / Marlin through their exact
// K-packed re-layout) at M <= 2: the INTEGER tensor path.
//
// On B200 a decode-sized matrix arrives from HBM at ~46 int4 weights per SM-cycle, so the kernel is bound by
// instructions per weight, not by arithmetic.  The fp16 formulation (gemv_stream.cu) spends one LOP3 per two weights
// to cut nibbles into fp16 sub-normals and one HMMA.16816 per 256 weights.  Here:
//
//   * a packed word (8 consecutive k of one column) becomes two MMA operand registers with three ALU ops:
//     (w & 0x0f0f0f0f) = the even k as four u8, ((w >> 4) & 0x0f0f0f0f) = the odd k -- no magic numbers, no pairing;
//   * the activations are split ONCE per launch into three balanced base-128 digits of a 21-bit fixed-point value
//     (power-of-two scale per <= 128-k part: block fixed point, an element below 2^-9 of its part's maximum keeps fewer
//     than fp16's 11 bits -- an absolute error below 2^-21 of that maximum); the three digits ride in three
//     of the eight B columns of the same mma.sync.m16n8k32.s32.u8.s8 (SASS IMMA.16832.U8.S8, 512 weights per
//     instruction), whose int32 accumulation is exact; a second token uses columns 4..6;
//   * per part: y += s * (2^-E * (d0 2^14 + d1 2^7 + d2) - z * sum(x)), fp32 -- the same algebra as the fp16 kernels.
//
// Everything around the math is the streaming design of gemv_stream.cu: per-warp cp.async rings filled before
// griddepcontrol.wait, no CTA-wide synchronisation until the final reduction, sibling layers in one launch, split-K
// across the warps and a thread-block cluster with a fixed-order reduction.
// Replaces ort_ops.gemv (dq_gemv.cu:40-177) and, through the re-layout, gemm_forward_cuda / Marlin at M <= 2.
#include "gemv_stream.cuh"

namespace b200q {

static constexpr int kINT = 64;             // columns per tile
static constexpr int kIStep = 64;           // k per step (two m16n8k32 deep)
static constexpr int kIRow = 288;           // shared-memory pitch of a packed row: 256 B + 32 B pad (conflict-free LDS.128)
static constexpr int kIStepBytes = 8 * kIRow;   // one ring slot: 8 packed rows (64 k) x 64 columns
static constexpr uint32_t NIB = 0x0f0f0f0fu;

__device__ __forceinline__ void imma_16832(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Tile = 64 columns, step = 64 k (8 packed rows of 256 B).  The tile is kept narrow so that a thread needs ~70
// registers (16 int32 accumulators, 8 fp32 partial sums) and three CTAs -- 24 warps -- share an SM: at decode sizes every
// warp runs a mostly serial instruction stream, and the time of a layer is that stream's length.
// Shared-memory placement of one step: packed rows 288 B apart, so the quarter-warps of an LDS.128 (rows t = 0..3 of a
// k-half, 16-byte chunks 8i + g) fall into distinct banks without any swizzle; every address is base + immediate.
// Column map of a tile: MMA j = 2i + h (i, h = 0..1), row g -> column 32i + 4g + 2h, row g + 8 -> that + 1
// (a lane's LDS.128 #i = its four words 32i + 4g .. + 3 of packed row 4 kh + t).
// Activation digits in shared memory, per 32-k sub-step: [column c < 4 MTOK][t = 0..3][8 B] -- the B fragment (b0, b1) of
// lane (g = c, t): bytes 0..3 = digit at k = 8t + {0,2,4,6}, bytes 4..7 = k = 8t + {1,3,5,7} (the order the unpack yields).
// Column 4m + d holds digit d (most significant first) of token m; column 4m + 3 is absent (lanes read a zero pad).
// PEER: the cross-GPU hand-off (counter wait / post, tagged activations) is compiled in; the single-GPU instantiation carries none of it.
template <int MTOK, int D, bool PEER, bool FUSED>
__global__ void __launch_bounds__(kRpThreads, 3) gemv_imma_kernel(const __grid_constant__ StParams p) {
  extern __shared__ __align__(128) char smem[];
  using T = RpGptq<4>;                                                   // table_entries8: (scale, zero) decode of the K-packed layout
  constexpr int NT = kINT, MC = (MTOK == 1) ? 1 : 2;
  constexpr int XQ_SUB = 32 * 4 * MTOK;                                  // digit bytes per 32-k sub-step
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

  const int unit = rank * kWarps + warp;
  const int s_begin = unit * p.split_q + min(unit, p.split_r), s_end = (unit + 1) * p.split_q + min(unit + 1, p.split_r);
  const int cta_s0 = rank * kWarps * p.split_q + min(rank * kWarps, p.split_r);
  const int cta_s1 = (rank + 1) * kWarps * p.split_q + min((rank + 1) * kWarps, p.split_r);
  const int k_cta0 = cta_s0 * kIStep, k_cta1 = cta_s1 * kIStep;
  const int gsh = p.group_shift;
  const int g_first = k_cta0 >> gsh;
  const int g_count = (k_cta1 > k_cta0) ? ((k_cta1 - 1) >> gsh) - g_first + 1 : 0;
  const int nsw = s_end - s_begin, total = nsw * nt;

  ST_STAMP(0);
  pdl_launch_dependents();
  // node-epoch mode: this call's step word is stable from the moment the kernel starts -- read it now, behind the weight
  // prefetch, instead of as an L2 round trip in front of the tagged x loads and another in front of the stores
  uint32_t step_early = 0;
  if (PEER && p.sync.node_epoch) step_early = st_step(p);
  if (cs > 1) {
    if (rank == 0 && tid == 0) {
      mbar_init(rbar, 1);
      mbar_expect_tx(rbar, (uint32_t)(cs - 1) * (uint32_t)(nt * NT * p.M) * 4u);
      fence_mbar_init();
    }
    cluster_arrive_relaxed();
  }

  // ---- per-warp ring of D slots filled with cp.async: lane -> packed rows lane / 16 + {0, 2, 4, 6}, chunk lane % 16 ----
  const uint32_t ring = smem_u32(smem + p.off_ring) + (uint32_t)warp * (D * kIStepBytes);
  const uint32_t rd = ring + (uint32_t)(t * kIRow + g * 16);                                 // + 128 i, + 4 kIRow kh, + slot
  const uint32_t wr = ring + (uint32_t)((lane >> 4) * kIRow + (lane & 15) * 16);             // + 2 kIRow r, + slot
  int pitch = L.N;                                                                           // words per packed row
  uint32_t wr_ = wr, rd_ = rd;
  asm volatile("" : "+r"(wr_), "+r"(rd_), "+r"(pitch));                                      // keep in registers: ptxas otherwise re-derives them per step
  const uint32_t* src = L.qw + ((size_t)(s_begin * 8) + (lane >> 4)) * (size_t)pitch + n0 + 4 * (lane & 15);
  bool pc = 4 * (lane & 15) < min(NT, L.N - n0);
  int irem = nsw, itiles = (nsw > 0) ? nt : 0;
  auto issue = [&](uint32_t slot_wr) {
    if (itiles > 0) {
      if (pc) {
        const uint32_t* s1 = src + 2 * pitch;
        const uint32_t* s2 = s1 + 2 * pitch;
        cp_async16_s(slot_wr, src);
        cp_async16_s(slot_wr + 2 * kIRow, s1);
        cp_async16_s(slot_wr + 4 * kIRow, s2);
        cp_async16_s(slot_wr + 6 * kIRow, s2 + 2 * pitch);
      }
      src += 8 * pitch;
      if (--irem == 0) {                                                                     // next tile, back to this warp's first k
        irem = nsw;
        --itiles;
        src += NT - (ptrdiff_t)((size_t)nsw * 8 * (size_t)pitch);
        pc = 4 * (lane & 15) < L.N - (n0 + (nt - itiles) * NT);
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int d = 0; d < D; ++d) issue(wr_ + d * kIStepBytes);

  // (scale, zero) table of the CTA's k-slice: [tile][group][NT] float2
  {
    const int c8 = (tid & 7) * 8;
    for (int tt = 0; tt < nt; ++tt) {
      const int n = n0 + tt * NT + c8;
      for (int gl = tid >> 3; gl < g_count; gl += 32) {
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

  // ---- activations: wait for the upstream kernel, then split this warp's k-range into digits ----
  // parts: <= 128 k (4 sub-steps of 32 k), never across a group boundary; part table (per warp): {2^-E, sum(x)} per token
  // tagged x in node-epoch mode: the tags order the data and the step comes from this call's own word, so the load stage
  // does not wait for the kernel in front of it (st_reduce_store waits ahead of the stores instead)
  if (!(PEER && p.sync.node_epoch && p.sync.x_tagged)) pdl_wait();
  if (PEER) {
    ST_STAMP(3);                                                         // local upstream done; stamp 2 follows the peers' posts
    st_sync_wait(p, lane);
  }
  ST_STAMP(2);
  const uint32_t step_now = (PEER && (p.sync.x_tagged || p.sync.y_tagged)) ? (p.sync.node_epoch ? step_early : st_step(p)) : 0u;
  const uint32_t xtag = (step_now * p.sync.tag_stride + p.sync.x_seq) & 0xffffu;
  char* xq = smem + p.off_x + (size_t)(s_begin - cta_s0) * (2 * XQ_SUB);   // this warp's digit sub-steps
  float2* part = reinterpret_cast<float2*>(smem + p.off_part) + (size_t)warp * (p.part_cap * MTOK);
  const int part_sub = min(4, p.group >> 5);                             // sub-steps per full part (2 or 4: group >= 64)
  {
    int ks = 2 * s_begin, pi = 0;                                        // in 32-k sub-steps
    const int ke = 2 * s_end;
    while (ks < ke) {
      const int lim = min(ke, (ks / part_sub + 1) * part_sub);
      const int len = (lim - ks) * 32;                                   // k in this part (<= 128)
#pragma unroll
      for (int m = 0; m < MTOK; ++m) {
        float xv[4] = {0.f, 0.f, 0.f, 0.f};
        if (4 * lane < len) {
          const size_t xo = (size_t)m * p.ldx + (size_t)ks * 32 + 4 * lane;
          uint2 raw;
          if (PEER && p.sync.x_tagged && p.xperm) {                      // act-order layer behind a tagged producer: gather the four words
            const int4 pi = *reinterpret_cast<const int4*>(p.xperm + (size_t)ks * 32 + 4 * lane);
            raw = st_gather_tagged4(p, reinterpret_cast<const uint32_t*>(p.x) + (size_t)m * p.ldx, pi, xtag);
          } else if (PEER && p.sync.x_tagged) {
            raw = st_load_tagged4(p, reinterpret_cast<const uint32_t*>(p.x) + xo, xtag);
          } else if (p.xperm) {                                          // act-order re-layout: packed row j multiplies x[x_perm[j]]
            const int4 pi = *reinterpret_cast<const int4*>(p.xperm + (size_t)ks * 32 + 4 * lane);
            const unsigned short* xr = reinterpret_cast<const unsigned short*>(p.x + (size_t)m * p.ldx);
            raw = make_uint2((uint32_t)xr[pi.x] | ((uint32_t)xr[pi.y] << 16), (uint32_t)xr[pi.z] | ((uint32_t)xr[pi.w] << 16));
          } else {
            raw = *reinterpret_cast<const uint2*>(p.x + xo);
          }
          const __half2 h01 = *reinterpret_cast<const __half2*>(&raw.x), h23 = *reinterpret_cast<const __half2*>(&raw.y);
          xv[0] = __low2float(h01); xv[1] = __high2float(h01); xv[2] = __low2float(h23); xv[3] = __high2float(h23);
          const bool bf16 = FUSED && st_bf16(p.layer[0]);
          if (bf16) {                                                    // bf16 activations: the reference's cast to fp16, in the load
            xv[0] = bf16_bits_to_float(raw.x & 0xffffu); xv[1] = bf16_bits_to_float(raw.x >> 16);
            xv[2] = bf16_bits_to_float(raw.y & 0xffffu); xv[3] = bf16_bits_to_float(raw.y >> 16);
          }
          const __half* const xmul = FUSED ? (const __half*)p.layer[0].out.y[3] : nullptr;
          if (FUSED && xmul) {                                         // fused act(gate) * up: x = silu(x) * x_mul, rounded as the fp16 ops round
            uint2 rm;
            if (p.xperm) {
              const int4 pi = *reinterpret_cast<const int4*>(p.xperm + (size_t)ks * 32 + 4 * lane);
              const unsigned short* xr = reinterpret_cast<const unsigned short*>(xmul + (size_t)m * p.ldx);
              rm = make_uint2((uint32_t)xr[pi.x] | ((uint32_t)xr[pi.y] << 16), (uint32_t)xr[pi.z] | ((uint32_t)xr[pi.w] << 16));
            } else {
              rm = *reinterpret_cast<const uint2*>(xmul + xo);
            }
            if (bf16) {
              xv[0] = silu_mul_bf16_to_f16(xv[0], bf16_bits_to_float(rm.x & 0xffffu)); xv[1] = silu_mul_bf16_to_f16(xv[1], bf16_bits_to_float(rm.x >> 16));
              xv[2] = silu_mul_bf16_to_f16(xv[2], bf16_bits_to_float(rm.y & 0xffffu)); xv[3] = silu_mul_bf16_to_f16(xv[3], bf16_bits_to_float(rm.y >> 16));
            } else {
              const __half2 u01 = *reinterpret_cast<const __half2*>(&rm.x), u23 = *reinterpret_cast<const __half2*>(&rm.y);
              xv[0] = silu_mul_f16(xv[0], __low2float(u01)); xv[1] = silu_mul_f16(xv[1], __high2float(u01));
              xv[2] = silu_mul_f16(xv[2], __low2float(u23)); xv[3] = silu_mul_f16(xv[3], __high2float(u23));
            }
          } else if (bf16) {
#pragma unroll
            for (int e = 0; e < 4; ++e) xv[e] = __half2float(__float2half_rn(xv[e]));
          }
        }
        // a NaN / Inf activation must poison its part (the fixed-point conversion would turn it into finite garbage; the
        // reference's fp16 FMA chains propagate both)
        // fp16-ranged values: the fp32 sum of four is finite unless one of them is not -- one vote beside the maximum's
        // shuffle chain.  (Measured on one box, tok/s of the 7B decode bench: no check 932, this vote 920, a NaN-propagating
        // max chain 917, integer max of the bit patterns 914, a CTA flag + poisoned partial sums 908.)
        const float fsum = (xv[0] + xv[1]) + (xv[2] + xv[3]);
        const bool bad = __any_sync(0xffffffffu, !(fabsf(fsum) <= 3.0e38f));
        float mx = fmaxf(fmaxf(fabsf(xv[0]), fabsf(xv[1])), fmaxf(fabsf(xv[2]), fabsf(xv[3])));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const int ex = (int)((__float_as_uint(mx) >> 23) & 0xffu);       // biased exponent of the maximum
        // scale 2^E with max * 2^E in [2^19, 2^20): exponent arithmetic on the float's bit pattern (mx == 0 -> E = 0)
        const int E = (mx > 0.f) ? (19 + 127 - ex) : 0;
        const float sc = __uint_as_float((uint32_t)(E + 127) << 23), isc = __uint_as_float((uint32_t)(127 - E) << 23);
        int tsum = 0;
        uint32_t dig[3] = {0u, 0u, 0u};                                  // per digit: 4 bytes (k = 4 lane .. + 3)
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
        if (lane == 0) part[pi * MTOK + m] = bad ? make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000))
                                                 : make_float2(isc, (float)tsum * isc);
        if (4 * lane < len) {
          // element e of this lane: k' = 4 (lane % 8) + e inside its sub-step -> word t' = (lane % 8) / 2, kk = 4 (lane & 1) + e:
          // even kk -> byte kk / 2 of the first half, odd kk -> byte kk / 2 of the second half
          const int st = (ks - 2 * s_begin) + (lane >> 3), tq = (lane & 7) >> 1, bo = 2 * (lane & 1);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            char* dst = xq + (size_t)st * XQ_SUB + (size_t)(4 * m + d) * 32 + tq * 8 + bo;
            const uint32_t v = dig[d];
            *reinterpret_cast<uint16_t*>(dst) = (uint16_t)((v & 0xffu) | ((v >> 8) & 0xff00u));              // e = 0, 2
            *reinterpret_cast<uint16_t*>(dst + 4) = (uint16_t)(((v >> 8) & 0xffu) | ((v >> 16) & 0xff00u));  // e = 1, 3
          }
        }
      }
      ks = lim;
      ++pi;
    }
  }
  // B fragment source: column g of the digit block (digit columns 4m + {0,1,2}), else the zero pad
  const bool xreal = (g & 3) < 3 && (g >> 2) < MTOK;
  uint32_t xp = xreal ? smem_u32(xq) + (uint32_t)(g * 32 + t * 8) : smem_u32(zpad);
  const uint32_t xsub = xreal ? (uint32_t)XQ_SUB : 0u, xrewind = xreal ? (uint32_t)(nsw * 2 * XQ_SUB) : 0u;
  // fix-up role of this lane: token t >> 1; even t holds digits 0, 1 (v = d0 2^7 + d1, worth 2^7), odd t digit 2 (v = d2 2^7, worth 2^-7)
  const int mytok = t >> 1;
  const bool fx_on = mytok < MTOK;
  const float dscale = (t & 1) ? (1.0f / 128.0f) : 128.0f, zflag = (t & 1) ? 0.f : 1.f;

  int acc[4][4];
  float tot[4][2];
#pragma unroll
  for (int q = 0; q < 4; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0; tot[q][0] = tot[q][1] = 0.f; }
  // part / tile cursors (in steps; a part is part_sub / 2 = 1 or 2 steps)
  const int psteps = part_sub >> 1;
  const int pleft0 = psteps - (s_begin & (psteps - 1));
  int pleft = min(pleft0, nsw), crem = nsw, tcur = 0, pi = 0, kcur = s_begin;
  float* redw = red + (size_t)warp * p.red_stride;
  const int ms = p.M;

  auto part_close = [&]() {
    // y += s * (2^-E * dscale * (d_a 2^7 + d_b) - zflag * z * sum(x)) for this lane's 8 (column, token) outputs
    if (fx_on) {
      const float2 pt = part[pi * MTOK + mytok];
      const float xs = pt.x * dscale, sxz = pt.y * zflag;
      const int gl = (((kcur - 1) * kIStep) >> gsh) - g_first;           // group of the part just finished
      const float2* row = tab + ((size_t)tcur * p.gcap + gl) * NT + 4 * g;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float4 e01 = *reinterpret_cast<const float4*>(row + 32 * i), e23 = *reinterpret_cast<const float4*>(row + 32 * i + 2);
        // MMA 2i: rows g / g+8 = columns +0 / +1; MMA 2i+1: columns +2 / +3
        const float v00 = (float)((acc[2 * i][0] << 7) + acc[2 * i][1]), v01 = (float)((acc[2 * i][2] << 7) + acc[2 * i][3]);
        const float v10 = (float)((acc[2 * i + 1][0] << 7) + acc[2 * i + 1][1]), v11 = (float)((acc[2 * i + 1][2] << 7) + acc[2 * i + 1][3]);
        tot[2 * i][0] = fmaf(e01.x, fmaf(v00, xs, -e01.y * sxz), tot[2 * i][0]);
        tot[2 * i][1] = fmaf(e01.z, fmaf(v01, xs, -e01.w * sxz), tot[2 * i][1]);
        tot[2 * i + 1][0] = fmaf(e23.x, fmaf(v10, xs, -e23.y * sxz), tot[2 * i + 1][0]);
        tot[2 * i + 1][1] = fmaf(e23.z, fmaf(v11, xs, -e23.w * sxz), tot[2 * i + 1][1]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0;
    ++pi;
  };

  cp_async_wait<0>();
  __syncwarp();                                                          // ring, digits and part table visible to the warp
#pragma unroll 1
  for (int i = 0; i < total; ++i) {
    const uint32_t so = (uint32_t)(i & (D - 1)) * kIStepBytes;
    cp_async_wait<D - 1>();
    __syncwarp();
#pragma unroll
    for (int kh = 0; kh < 2; ++kh) {
      const uint4 wa = lds128_s(rd_ + so + kh * 4 * kIRow), wb = lds128_s(rd_ + so + kh * 4 * kIRow + 128);
      const uint2 xb = lds64_s(xp);
      xp += xsub;
      imma_16832(acc[0], wa.x & NIB, wa.y & NIB, (wa.x >> 4) & NIB, (wa.y >> 4) & NIB, xb.x, xb.y);
      imma_16832(acc[1], wa.z & NIB, wa.w & NIB, (wa.z >> 4) & NIB, (wa.w >> 4) & NIB, xb.x, xb.y);
      imma_16832(acc[2], wb.x & NIB, wb.y & NIB, (wb.x >> 4) & NIB, (wb.y >> 4) & NIB, xb.x, xb.y);
      imma_16832(acc[3], wb.z & NIB, wb.w & NIB, (wb.z >> 4) & NIB, (wb.w >> 4) & NIB, xb.x, xb.y);
    }
    __syncwarp();                                                        // every lane has read the slot
    issue(wr_ + so);
    ++kcur;
    --crem;
    if (--pleft == 0 || crem == 0) {                                     // part and / or tile boundary (warp-uniform)
      part_close();
      pleft = min(psteps, crem);
      if (crem == 0) {                                                   // tile done: combine the digit lanes, park the sums
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a0 = tot[q][0] + __shfl_xor_sync(0xffffffffu, tot[q][0], 1);
          const float a1 = tot[q][1] + __shfl_xor_sync(0xffffffffu, tot[q][1], 1);
          if (fx_on && !(t & 1)) {
            const int n = 32 * (q >> 1) + 4 * g + 2 * (q & 1);           // row g -> n, row g + 8 -> n + 1
            float* r = redw + (size_t)tcur * NT * ms;
            r[n * ms + mytok] = a0;
            r[(n + 1) * ms + mytok] = a1;
          }
          tot[q][0] = tot[q][1] = 0.f;
        }
        ++tcur;
        crem = nsw;
        kcur = s_begin;
        pi = 0;
        pleft = min(pleft0, nsw);
        xp -= xrewind;
      }
    }
  }
  if (total == 0) {
    cp_async_wait<0>();
    for (int idx = lane; idx < nt * NT * ms; idx += 32) redw[idx] = 0.f;
  }
  ST_STAMP(4);
  st_reduce_store<MC, 256, PEER, FUSED>(p, SL, red, rbuf, rbar, nt * NT, ncols_cta, n0, cs, rank, tid, step_now);
}

// ------------------------------------------------------------------------------------------------
struct ImPlan {
  int ctas, cluster, tpc, depth, steps_total, group_shift, gcap, part_cap, split_q, split_r;
  int cta0[kMaxGroupLayers];
  int off_x, off_tab, off_red, off_rbuf, off_rbar, off_ring, off_zpad, off_part, off_mbar, smem_bytes;
};

static int g_im_on = 1, g_im_cluster = 0, g_im_depth = 0, g_im_tpc = 0, g_im_target = 296;
static unsigned long long* g_im_dbg = nullptr;
static size_t g_im_dbg_cap = 0, g_im_dbg_pos = 0;
void gemv_imma_set_option(int which, int value) {
  if (which == 0) g_im_on = value;
  else if (which == 1) g_im_cluster = value;
  else if (which == 2) g_im_depth = value;
  else if (which == 3) g_im_tpc = value;
  else if (which == 4) g_im_target = value;
}
void gemv_imma_set_debug(unsigned long long* buf, size_t cap_entries) { g_im_dbg = buf; g_im_dbg_cap = cap_entries; g_im_dbg_pos = 0; }

static bool im_plan(const LinearArgs* a, int n, ImPlan& pl) {
  if (!g_im_on || n < 1 || n > kMaxGroupLayers) return false;
  const LayerView& L = a[0].L;
  const int M = a[0].M;
  if (M < 1 || M > 2 || L.g_idx != nullptr || L.bits != 4) return false;
  if (L.layout != B200Q_LAYOUT_GPTQ && L.layout != B200Q_LAYOUT_HQQ) return false;
  for (int i = 1; i < n; ++i) {
    const LayerView& B = a[i].L;
    if (B.layout != L.layout || B.bits != L.bits || B.group != L.group || B.K != L.K || B.zero_bias != L.zero_bias ||
        B.g_idx != nullptr || B.x_perm != L.x_perm || a[i].M != M || a[i].x != a[0].x || a[i].ldx != a[0].ldx)
      return false;
  }
  // a step (64 k) lies inside one group; groups are powers of two (count-down bookkeeping by shifts)
  if (L.group < kIStep || (L.group & (L.group - 1)) != 0 || L.K % L.group != 0 || L.K % kIStep != 0) return false;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    if (a[i].L.N % 32 != 0) return false;
    tiles += (a[i].L.N + kINT - 1) / kINT;
  }
  pl.group_shift = 0;
  while ((1 << pl.group_shift) < L.group) ++pl.group_shift;
  pl.steps_total = L.K / kIStep;
  // tiles per CTA: widest walk (<= 256 columns) that still leaves >= target CTA groups
  int tpc = 1;
  for (int c = 2; c <= 4; c *= 2)
    if ((tiles + c - 1) / c >= g_im_target) tpc = c;
  if (g_im_tpc == 1 || g_im_tpc == 2 || g_im_tpc == 4) tpc = g_im_tpc;
  int groups = 0;
  for (int i = 0; i < n; ++i) {
    pl.cta0[i] = groups;
    groups += ((a[i].L.N + kINT - 1) / kINT + tpc - 1) / tpc;
  }
  int cs = 1;
  while (cs < 8 && groups * cs < g_im_target && pl.steps_total / ((cs + 1) * kWarps) >= 2) ++cs;
  if (g_im_cluster > 0 && g_im_cluster <= 8 && pl.steps_total / (g_im_cluster * kWarps) >= 1) cs = g_im_cluster;
  pl.tpc = tpc; pl.cluster = cs; pl.ctas = groups * cs;
  if (pl.ctas > 148 * 3) return false;                                   // three CTAs per SM (80 registers per thread)
  pl.split_q = pl.steps_total / (cs * kWarps);
  pl.split_r = pl.steps_total % (cs * kWarps);
  const int seq = (pl.split_q + (pl.split_r ? 1 : 0)) * tpc;             // longest per-warp step sequence
  const int per_sm = (pl.ctas + 147) / 148;
  const int slice_steps = kWarps * pl.split_q + (pl.split_r < kWarps ? pl.split_r : kWarps);
  const int kslice = slice_steps * kIStep;
  pl.gcap = kslice / L.group + 2;
  const int psteps = (L.group >> 6) < 2 ? 1 : 2;                         // steps per part (<= 128 k)
  pl.part_cap = (pl.split_q + 1 + psteps - 1) / psteps + 2;              // parts per warp
  auto layout = [&](int depth) {
    int off = 0;
    pl.off_x = off; off += slice_steps * 2 * 32 * 4 * M;                 // digit sub-steps of the CTA's slice
    off = (off + 15) & ~15;
    pl.off_tab = off; off += tpc * pl.gcap * kINT * 8;
    off = (off + 15) & ~15;
    pl.off_red = off; off += kWarps * tpc * kINT * M * 4;
    pl.off_rbuf = off; off += (cs - 1) * tpc * kINT * M * 4;
    off = (off + 7) & ~7;
    pl.off_rbar = off; off += 8;
    pl.off_part = off; off += kWarps * pl.part_cap * M * 8;
    pl.off_mbar = off; off += kWarps * depth * 8;
    off = (off + 127) & ~127;
    pl.off_zpad = off; off += 128;
    pl.off_ring = off; off += kWarps * depth * kIStepBytes;
    pl.smem_bytes = off;
    pl.depth = depth;
  };
  // ring depth: four slots per warp (74 KB per CTA) only when the launch leaves one CTA per SM, else two (37 KB)
  int depth = (per_sm >= 2 || seq <= 2) ? 2 : 4;
  if (g_im_depth == 2 || g_im_depth == 4) depth = g_im_depth;
  layout(depth);
  if (pl.smem_bytes > 72 * 1024 && depth == 4) layout(2);
  (void)seq;
  return pl.smem_bytes <= 72 * 1024;
}

bool gemv_imma_supported(const LinearArgs* a, int n) {
  ImPlan pl;
  if (!im_plan(a, n, pl)) return false;
  if (((uintptr_t)a[0].x & 7) != 0 || (a[0].ldx % 4) != 0 || ((uintptr_t)a[0].x_mul & 7) != 0) return false;
  for (int i = 0; i < n; ++i)
    if (((uintptr_t)a[i].L.qw & 15) != 0 || ((uintptr_t)a[i].L.s & 15) != 0) return false;
  return true;
}

bool gemv_imma_describe(const LinearArgs* a, int n, int out[6]) {
  ImPlan pl;
  if (!im_plan(a, n, pl)) return false;
  out[0] = pl.cluster; out[1] = pl.ctas; out[2] = pl.smem_bytes; out[3] = pl.steps_total; out[4] = pl.tpc; out[5] = pl.depth;
  return true;
}

template <int MTOK, int D, bool PEER, bool FUSED = false>
static cudaError_t im_launch_k(const StParams& p, const ImPlan& pl, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemv_imma_kernel<MTOK, D, PEER, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    if (e != cudaSuccess) return e;
    if (decode_carveout_max()) cudaFuncSetAttribute(gemv_imma_kernel<MTOK, D, PEER, FUSED>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
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
  return cudaLaunchKernelEx(&cfg, gemv_imma_kernel<MTOK, D, PEER, FUSED>, p);
}

int gemv_imma_posts(const LinearArgs* a, int n) {
  ImPlan pl;
  return im_plan(a, n, pl) ? 1 : -1;                       // the last storing CTA posts for the whole launch
}

cudaError_t launch_gemv_imma(const LinearArgs* a, int n, const PeerOut* peers, const PeerSync* sync) {
  ImPlan pl;
  if (!im_plan(a, n, pl)) return cudaErrorInvalidValue;
  const LayerView& L = a[0].L;
  StParams p = {};
  p.n_layers = n;
  for (int i = 0; i < kMaxGroupLayers; ++i) {
    const int k = i < n ? i : n - 1;
    StLayer& d = p.layer[i];
    d.qw = a[k].L.qw; d.qz = a[k].L.qz; d.s = a[k].L.s; d.bias = a[k].L.bias; d.N = a[k].L.N;
    d.cta0 = i < n ? pl.cta0[i] : (1 << 30);
    if (peers) d.out = peers[k]; else { d.out.n = 1; d.out.y[0] = a[k].y; }
    d.ldy = a[k].ldy; d.n_offset = a[k].n_offset;
  }
  if (sync) {
    p.sync = *sync;
    p.arrive = reinterpret_cast<unsigned int*>(a[0].workspace) + 1000;   // inside the 4 KB counter region every kernel leaves zeroed
    p.store_ctas = pl.ctas / pl.cluster;
    p.sync_flags = decode_sync_flags();
    if (sync->n_peers > 1 && sync->post_slot >= 0 && (!a[0].workspace || a[0].workspace_bytes < kCounterBytes)) return cudaErrorInvalidValue;
    if (sync->node_epoch && (!a[0].workspace || a[0].workspace_bytes < kCounterBytes)) return cudaErrorInvalidValue;
  }
  p.layout = L.layout; p.bits = L.bits; p.group = L.group; p.K = L.K; p.G = L.G; p.zero_bias = L.zero_bias;
  p.x = a[0].x; p.ldx = a[0].ldx; p.M = a[0].M; p.xperm = L.x_perm;
  if ((a[0].x_mul || a[0].act_bf16) && sync) return cudaErrorInvalidValue;     // tagged activations carry no second operand
  p.cluster = pl.cluster; p.tpc = pl.tpc; p.depth = pl.depth; p.steps_total = pl.steps_total; p.group_shift = pl.group_shift;
  p.gcap = pl.gcap; p.split_q = pl.split_q; p.split_r = pl.split_r; p.part_cap = pl.part_cap;
  p.x_stride = 0; p.red_stride = pl.tpc * kINT * a[0].M;
  p.off_x = pl.off_x; p.off_tab = pl.off_tab; p.off_red = pl.off_red; p.off_rbuf = pl.off_rbuf; p.off_rbar = pl.off_rbar;
  p.off_ring = pl.off_ring; p.off_zpad = pl.off_zpad; p.off_part = pl.off_part; p.off_mbar = pl.off_mbar;
  p.dbg = nullptr;
  if (g_im_dbg) {
    const size_t need = (size_t)pl.ctas * 8;
    if (g_im_dbg_pos + need <= g_im_dbg_cap) { p.dbg = g_im_dbg + g_im_dbg_pos; g_im_dbg_pos += need; }
  }
  if (sync && (sync->n_peers > 1 || sync->x_tagged || sync->y_tagged)) {
    if (p.M == 1) return pl.depth == 2 ? im_launch_k<1, 2, true>(p, pl, a[0].stream) : im_launch_k<1, 4, true>(p, pl, a[0].stream);
    return pl.depth == 2 ? im_launch_k<2, 2, true>(p, pl, a[0].stream) : im_launch_k<2, 4, true>(p, pl, a[0].stream);
  }
  bool fused = a[0].x_mul != nullptr || a[0].act_bf16 != 0;
  for (int i = 0; i < n; ++i) fused = fused || a[i].residual != nullptr;
  if (fused) {
    if (peers) return cudaErrorInvalidValue;                            // the fusion operands live in the peer slots
    for (int i = 0; i < kMaxGroupLayers; ++i) {
      const int k = i < n ? i : n - 1;
      st_set_fusion(p.layer[i], a[k].residual, a[k].ldres, a[0].x_mul, a[0].act_bf16 != 0);
    }                                                           // b200q_linear_ex: prologue / epilogue compiled in
    if (p.M == 1) return pl.depth == 2 ? im_launch_k<1, 2, false, true>(p, pl, a[0].stream) : im_launch_k<1, 4, false, true>(p, pl, a[0].stream);
    return pl.depth == 2 ? im_launch_k<2, 2, false, true>(p, pl, a[0].stream) : im_launch_k<2, 4, false, true>(p, pl, a[0].stream);
  }
  if (p.M == 1) return pl.depth == 2 ? im_launch_k<1, 2, false>(p, pl, a[0].stream) : im_launch_k<1, 4, false>(p, pl, a[0].stream);
  return pl.depth == 2 ? im_launch_k<2, 2, false>(p, pl, a[0].stream) : im_launch_k<2, 4, false>(p, pl, a[0].stream);
}

}  // namespace b200q
