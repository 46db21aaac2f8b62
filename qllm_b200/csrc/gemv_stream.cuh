// Shared pieces of the streaming decode kernels (gemv_stream.cu, gemv_imma.cu): launch parameters, cp.async helpers,
// the final split-K reduction + store.
#pragma once
#include "common.cuh"
#include "kernels.h"
#include "rp_layouts.cuh"

namespace b200q {

struct StLayer {
  const uint32_t* qw;
  const void* qz;
  const __half* s;
  const __half* bias;
  int N, cta0;              // output columns; first CTA-group index of this layer
  PeerOut out;
  int64_t ldy, n_offset;
};
// b200q_linear_ex operands ride in peer slots the single-GPU (FUSED) instantiations never use -- the parameter block stays at
// its round-1 size (56 more bytes measured 0.4 % of the decode bench): out.y[1] = residual ([M, ldres] fp16 or NULL),
// out.y[2] = ldres, layer[0].out.y[3] = x_mul, layer[0].out.y[4] = flags (bit 0: activations are bfloat16).  out.n stays 1.
__host__ __device__ __forceinline__ const __half* st_residual(const StLayer& SL) { return SL.out.y[1]; }
__host__ __device__ __forceinline__ int64_t st_ldres(const StLayer& SL) { return (int64_t)(uintptr_t)SL.out.y[2]; }
__host__ __device__ __forceinline__ bool st_bf16(const StLayer& SL0) { return ((uintptr_t)SL0.out.y[4] & 1u) != 0; }
__host__ inline void st_set_fusion(StLayer& SL, const __half* residual, int64_t ldres, const __half* x_mul, bool bf16) {
  SL.out.y[1] = const_cast<__half*>(residual);
  SL.out.y[2] = reinterpret_cast<__half*>((uintptr_t)ldres);
  SL.out.y[3] = const_cast<__half*>(x_mul);
  SL.out.y[4] = reinterpret_cast<__half*>((uintptr_t)(bf16 ? 1 : 0));
}

struct StParams {
  StLayer layer[kMaxGroupLayers];
  int n_layers;
  int layout, bits, group, K, G, zero_bias;      // shared by the layers of a group
  const __half* x;
  const int* xperm;                              // act-order re-layout: x is read through this map (integer-path kernel only)
  int64_t ldx;
  int M;
  int cluster, tpc, depth, steps_total, group_shift, gcap, split_q, split_r, part_cap;
  int x_stride;                                  // bytes of one staged activation row
  int red_stride;                                // floats between two warps' partial-sum vectors
  int off_x, off_tab, off_red, off_rbuf, off_rbar, off_ring, off_zpad, off_part, off_mbar;
  unsigned long long* dbg;                       // optional per-CTA phase stamps (diagnostic)
  // cross-GPU hand-off (PEER instantiations only; n_peers == 0: none)
  PeerSync sync;
  unsigned int* arrive;                          // local arrival counter of the storing CTAs (workspace, left zeroed)
  int store_ctas;
  int sync_flags;                                // diagnostic (B200Q_SYNC_FLAGS): 2 no wait, 4 no post, 8 back-off polling, 16 one poller per CTA
};

__device__ __forceinline__ unsigned long long st_gtime();
// Consumer side of the cross-GPU hand-off: x is complete once every peer's storing CTAs of this step have posted.
// One lane per warp polls the local counter (acquire, system scope); the spin is bounded (2 s) so a dead peer
// cannot hang the GPU -- the step's numbers are then garbage, and counter slot 0 of this rank is poisoned.
__device__ __forceinline__ void st_sync_wait(const StParams& p, int lane) {
  if (p.sync.n_peers > 1 && p.sync.wait_slot >= 0 && !p.sync.x_tagged && !(p.sync_flags & 2)) {
    if (p.sync_flags & 16) {
      if (threadIdx.x == 0) {
        const unsigned long long target = *reinterpret_cast<const volatile unsigned long long*>(p.sync.epoch) * p.sync.wait_count;
        const unsigned long long* c = p.sync.counters[p.sync.self] + p.sync.wait_slot;
        unsigned long long v;
        for (;;) {
          asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(c) : "memory");
          if (v >= target) break;
          __nanosleep(64);
        }
        asm volatile("fence.acq_rel.sys;" ::: "memory");
      }
      __syncthreads();
      return;
    }
    if (lane == 0) {
      const unsigned long long target = *reinterpret_cast<const volatile unsigned long long*>(p.sync.epoch) * p.sync.wait_count;
      const unsigned long long* c = p.sync.counters[p.sync.self] + p.sync.wait_slot;
      unsigned long long v, t0 = 0;
      unsigned spins = 0;
      for (;;) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(c) : "memory");
        if (v >= target) break;
        if (p.sync_flags & 8) __nanosleep(64);
        if ((++spins & 1023u) == 0) {
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          else if (now - t0 > 2000000000ull) { p.sync.counters[p.sync.self][0] = ~0ull; break; }
        }
      }
      asm volatile("fence.acq_rel.sys;" ::: "memory");      // the peers' stores that preceded their posts are visible from here on
    }
    __syncwarp();
  }
}
// Tagged activations: four consecutive elements (words) of x, spun on until every word carries this step's tag.
// Each lane polls only the words it needs -- the hand-off is data-flow, not a barrier.  Same 2 s bound as st_sync_wait.
__device__ __forceinline__ uint2 st_load_tagged4(const StParams& p, const uint32_t* w, uint32_t tag) {
  uint32_t a, b, c, d, spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(w) : "memory");
    if (((a >> 16) == tag) & ((b >> 16) == tag) & ((c >> 16) == tag) & ((d >> 16) == tag)) break;
    if (p.sync.node_epoch && !(p.sync_flags & 32)) __nanosleep((p.sync_flags & 64) ? 32 : 96);   // this launch may have started long before its producer ends: poll gently
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { p.sync.counters[p.sync.self][0] = ~0ull; break; }
    }
  }
  return make_uint2((a & 0xffffu) | (b << 16), (c & 0xffffu) | (d << 16));
}
// step number of this launch: the step word, or (node-epoch mode) one more than the completed executions of this call
// ... the same for an act-order layer: the four words sit at x_perm[j .. j + 3]
__device__ __forceinline__ uint2 st_gather_tagged4(const StParams& p, const uint32_t* row, int4 idx, uint32_t tag) {
  uint32_t a, b, c, d, spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(a) : "l"(row + idx.x) : "memory");
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(b) : "l"(row + idx.y) : "memory");
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(c) : "l"(row + idx.z) : "memory");
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(d) : "l"(row + idx.w) : "memory");
    if (((a >> 16) == tag) & ((b >> 16) == tag) & ((c >> 16) == tag) & ((d >> 16) == tag)) break;
    if (p.sync.node_epoch && !(p.sync_flags & 32)) __nanosleep((p.sync_flags & 64) ? 32 : 96);
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { p.sync.counters[p.sync.self][0] = ~0ull; break; }
    }
  }
  return make_uint2((a & 0xffffu) | (b << 16), (c & 0xffffu) | (d << 16));
}
__device__ __forceinline__ uint32_t st_step(const StParams& p) {
  const volatile unsigned long long* e = reinterpret_cast<const volatile unsigned long long*>(p.sync.epoch);
  return p.sync.node_epoch ? (uint32_t)e[1 + p.sync.y_seq] + 1u : (uint32_t)e[0];
}
__device__ __forceinline__ uint32_t st_step_tag(const StParams& p, uint32_t seq) {
  return (st_step(p) * p.sync.tag_stride + seq) & 0xffffu;
}

// Producer side: called by every thread of a storing CTA after its stores to the peers' buffers.  The storing CTAs
// arrive on a LOCAL counter; the last one posts once per peer (many system-scope atomics on one remote word
// serialise at ~100 ns each over NVLink -- measured: 250 posts per call cost 24 us).
__device__ __forceinline__ void st_sync_post(const StParams& p, int tid) {
  if (p.sync.n_peers > 1 && p.sync.post_slot >= 0 && !p.sync.y_tagged && !(p.sync_flags & 4)) {
    __syncthreads();                                        // all of the CTA's peer stores precede the fence below
    if (tid == 0) {
      __threadfence_system();                               // ... and are performed at the peers before the arrival
      const unsigned int old = atomicAdd(p.arrive, 1u);
      if (old == (unsigned int)p.store_ctas - 1u) {
        *p.arrive = 0u;                                     // self-cleaning: the next call starts from zero
        __threadfence_system();
        for (int r = 0; r < p.sync.n_peers; ++r)
          if (r != p.sync.self)
            asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p.sync.counters[r] + p.sync.post_slot), "l"(1ull) : "memory");
      }
    }
  }
}

__device__ __forceinline__ unsigned long long st_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// slot 7 of a CTA's row: (grid size << 32 | CTA index), written with the first stamp (lets a reader split launches)
#define ST_STAMP(i) do { if (p.dbg && tid == 0) { p.dbg[(size_t)blockIdx.x * 8 + (i)] = st_gtime(); \
    if ((i) == 0) p.dbg[(size_t)blockIdx.x * 8 + 7] = ((unsigned long long)gridDim.x << 32) | blockIdx.x; } } while (0)

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// ring depth is a launch parameter (2, 4, 8 or 16): all but the newest depth-1 groups must have landed
__device__ __forceinline__ void cp_async_wait_ring(int depth) {
  if (depth == 8) cp_async_wait<7>();
  else if (depth == 16) cp_async_wait<15>();
  else if (depth == 4) cp_async_wait<3>();
  else cp_async_wait<1>();
}

// Final reduction shared by the stream kernels: the 8 warps' partial sums (red[warp][ncols_alloc * M], idx = n * M + m)
// -> one vector per CTA; CTAs of a cluster send theirs to rank 0 through st.async (fixed order); rank 0 adds bias,
// rounds to fp16 and stores (to every peer buffer when sharded).  MAXCOLS: columns a CTA may own.
// FUSED: the residual epilogue of b200q_linear_ex is compiled in (the plain instantiations carry none of it)
template <int MC, int MAXCOLS, bool PEER = false, bool FUSED = false>
__device__ __forceinline__ void st_reduce_store(const StParams& p, const StLayer& SL, const float* red, float* rbuf, uint64_t* rbar,
                                                int ncols_alloc, int ncols_cta, int n0, int cs, int rank, int tid,
                                                uint32_t step_now = 0) {
  __syncthreads();
  const int totalv = ncols_alloc * p.M;
  const int wstride = p.red_stride;                         // floats between two warps' partial vectors
  constexpr int NV = (MAXCOLS * (MC == 1 ? 1 : kMB) + kRpThreads - 1) / kRpThreads;
  float v[NV];
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    const int idx = tid + r * kRpThreads;
    float sum = 0.f;
    if (idx < totalv) {
#pragma unroll
      for (int wq = 0; wq < kWarps; ++wq) sum += red[(size_t)wq * wstride + idx];
    }
    v[r] = sum;
  }
  if (cs > 1) {
    cluster_wait();                                         // rank 0's mbarrier is armed
    if (rank != 0) {
#pragma unroll
      for (int r = 0; r < NV; ++r) {
        const int idx = tid + r * kRpThreads;
        if (idx < totalv) st_async_f32(rbuf + (size_t)(rank - 1) * totalv + idx, rbar, 0u, v[r]);
      }
      ST_STAMP(5);
      return;
    }
    mbar_wait(rbar, 0);
    ST_STAMP(5);
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      const int idx = tid + r * kRpThreads;
      if (idx < totalv)
        for (int q = 0; q < cs - 1; ++q) v[r] += rbuf[(size_t)q * totalv + idx];
    }
  }
  const uint32_t ytag = (PEER && p.sync.y_tagged) ? ((step_now * p.sync.tag_stride + p.sync.y_seq) & 0xffffu) << 16 : 0u;
  // node-epoch mode skipped the kernel-boundary wait ahead of the tagged x loads: the stores still follow the previous
  // kernel's last reads (and keep "kernel i ends after kernel i - 1" for everything downstream)
  if (PEER && p.sync.node_epoch && p.sync.x_tagged) pdl_wait();
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    const int idx = tid + r * kRpThreads;
    const int n = (MC == 1) ? idx : idx / p.M, m = (MC == 1) ? 0 : idx - n * p.M;
    if (idx < totalv && n < ncols_cta) {
      float o = v[r];
      if (SL.bias) o += __half2float(__ldg(SL.bias + n0 + n));
      __half h = __float2half_rn(o);
      if (FUSED && st_bf16(p.layer[0])) {                    // bf16 caller: fp16 result -> bf16, residual summed as the bf16 add rounds
        float v = round_bf16(__half2float(h));
        if (st_residual(SL))
          v = round_bf16(v + bf16_bits_to_float(__ldg(reinterpret_cast<const unsigned short*>(st_residual(SL)) + (size_t)m * st_ldres(SL) + n0 + n)));
        h = __ushort_as_half((unsigned short)float_to_bf16_bits(v));
      } else if (FUSED && st_residual(SL)) {
        h = __float2half_rn(__half2float(h) + __half2float(__ldg(st_residual(SL) + (size_t)m * st_ldres(SL) + n0 + n)));
      }
      if (PEER && p.sync.y_tagged) {                        // one 4-byte store per element and replica: value and tag land together
        const uint32_t w = ytag | (uint32_t)__half_as_ushort(h);
        for (int q = 0; q < SL.out.n; ++q)
          reinterpret_cast<uint32_t*>(SL.out.y[q])[(size_t)m * SL.ldy + SL.n_offset + n0 + n] = w;
        continue;
      }
      for (int q = 0; q < SL.out.n; ++q) SL.out.y[q][(size_t)m * SL.ldy + SL.n_offset + n0 + n] = h;
    }
  }
  if (PEER) st_sync_post(p, tid);
  if (PEER && p.sync.node_epoch) {                          // the last storing CTA records that this call has run once more
    __syncthreads();
    if (tid == 0) {
      unsigned int* arr = p.arrive + 2 + (p.sync.y_seq & 1u);   // neighbouring calls overlap: alternate words (self-cleaning)
      if (atomicAdd(arr, 1u) == (unsigned int)p.store_ctas - 1u) {
        *arr = 0u;
        const_cast<unsigned long long*>(p.sync.epoch)[1 + p.sync.y_seq] = (unsigned long long)step_now;
      }
    }
  }
  ST_STAMP(6);
}


__device__ __forceinline__ void cp_async16_s(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ uint2 lds64_s(uint32_t a) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t lds32_s(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}

__device__ __forceinline__ uint4 lds128_s(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
}  // namespace b200q
