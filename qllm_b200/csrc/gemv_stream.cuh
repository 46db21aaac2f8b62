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

struct StParams {
  StLayer layer[kMaxGroupLayers];
  int n_layers;
  int layout, bits, group, K, G, zero_bias;      // shared by the layers of a group
  const __half* x;
  int64_t ldx;
  int M;
  int cluster, tpc, depth, steps_total, group_shift, gcap, split_q, split_r, part_cap;
  int x_stride;                                  // bytes of one staged activation row
  int red_stride;                                // floats between two warps' partial-sum vectors
  int off_x, off_tab, off_red, off_rbuf, off_rbar, off_ring, off_zpad, off_part, off_mbar;
  unsigned long long* dbg;                       // optional per-CTA phase stamps (diagnostic)
};

__device__ __forceinline__ unsigned long long st_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ST_STAMP(i) do { if (p.dbg && tid == 0) p.dbg[(size_t)blockIdx.x * 8 + (i)] = st_gtime(); } while (0)

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
template <int MC, int MAXCOLS>
__device__ __forceinline__ void st_reduce_store(const StParams& p, const StLayer& SL, const float* red, float* rbuf, uint64_t* rbar,
                                                int ncols_alloc, int ncols_cta, int n0, int cs, int rank, int tid) {
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
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    const int idx = tid + r * kRpThreads;
    const int n = (MC == 1) ? idx : idx / p.M, m = (MC == 1) ? 0 : idx - n * p.M;
    if (idx < totalv && n < ncols_cta) {
      float o = v[r];
      if (SL.bias) o += __half2float(__ldg(SL.bias + n0 + n));
      const __half h = __float2half_rn(o);
      for (int q = 0; q < SL.out.n; ++q) SL.out.y[q][(size_t)m * SL.ldy + SL.n_offset + n0 + n] = h;
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
