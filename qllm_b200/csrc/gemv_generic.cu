// Generic decode kernel: any layout, any bit width 2..8, any group size, optional g_idx, M <= 16.
// CUDA cores, one thread per output column, split-K across blockIdx.y with a deterministic
// two-pass reduction.  This is the coverage path for configurations the mma decode kernel and
// the tcgen05 GEMM do not take (3/5/6/7-bit, odd shapes); hot configurations never come here.
// Replaces the reference's gemv<half>/Gemv_g (dq_gemv.cu:40-150, :459-541) for those cases.
#include "common.cuh"
#include "kernels.h"

namespace b200q {

static constexpr int kGenThreads = 128;

static int generic_ksplit(const LayerView& L) {
  const int nblk = (L.N + kGenThreads - 1) / kGenThreads;
  int ks = (2 * 148 + nblk - 1) / nblk;
  const int max_ks = (L.K + 255) / 256;
  if (ks > max_ks) ks = max_ks;
  if (ks > 32) ks = 32;
  return ks < 1 ? 1 : ks;
}

size_t gemv_generic_workspace(const LayerView& L, int M) {
  return kCounterBytes + (size_t)generic_ksplit(L) * (size_t)M * (size_t)L.N * sizeof(float);
}

template <int MT>
__global__ void __launch_bounds__(kGenThreads) gemv_generic_kernel(LayerView L, const __half* __restrict__ x, int64_t ldx,
                                                                    int M, int k_per_split, float* __restrict__ partial) {
  const int n = blockIdx.x * kGenThreads + threadIdx.x;
  const int k0 = blockIdx.y * k_per_split;
  const int k1 = min(L.K, k0 + k_per_split);
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.f;
  if (n < L.N) {
    int g_cur = -1;
    float z = 0.f, s = 0.f;
    for (int k = k0; k < k1; ++k) {
      const int g = group_of(L, k);
      if (g != g_cur) {
        g_cur = g;
        z = load_z(L, g, n);
        s = load_s(L, g, n);
      }
      const float w = __half2float(dequant_one(L, load_q(L, k, n), z, s));
#pragma unroll
      for (int m = 0; m < MT; ++m)
        if (m < M) acc[m] = fmaf(__half2float(__ldg(x + (size_t)m * ldx + k)), w, acc[m]);
    }
#pragma unroll
    for (int m = 0; m < MT; ++m)
      if (m < M) partial[((size_t)blockIdx.y * M + m) * L.N + n] = acc[m];
  }
}

__global__ void __launch_bounds__(256) gemv_generic_finalize(const float* __restrict__ partial, int ks, int M, int N,
                                                             const __half* __restrict__ bias, PeerOut out, int64_t ldy,
                                                             int64_t n_offset) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * N) return;
  const int m = (int)(idx / N), n = (int)(idx % N);
  float v = 0.f;
  for (int s = 0; s < ks; ++s) v += partial[((size_t)s * M + m) * N + n];
  if (bias) v += __half2float(__ldg(bias + n));
  const __half h = __float2half_rn(v);
  for (int p = 0; p < out.n; ++p) out.y[p][(size_t)m * ldy + n_offset + n] = h;
}

cudaError_t launch_gemv_generic(const LinearArgs& a, const PeerOut* peers) {
  const LayerView& L = a.L;
  const int ks = generic_ksplit(L);
  int kps = (L.K + ks - 1) / ks;
  kps = (kps + 31) / 32 * 32;
  dim3 grid((L.N + kGenThreads - 1) / kGenThreads, (L.K + kps - 1) / kps);
  float* partial = (float*)((char*)a.workspace + kCounterBytes);   // never touch the counter region
  PeerOut out;
  if (peers) out = *peers; else { out.n = 1; out.y[0] = a.y; }
#define B200Q_GEN(MT) gemv_generic_kernel<MT><<<grid, kGenThreads, 0, a.stream>>>(L, a.x, a.ldx, a.M, kps, partial)
  if (a.M <= 1) B200Q_GEN(1);
  else if (a.M <= 2) B200Q_GEN(2);
  else if (a.M <= 4) B200Q_GEN(4);
  else if (a.M <= 8) B200Q_GEN(8);
  else B200Q_GEN(16);
#undef B200Q_GEN
  const size_t total = (size_t)a.M * L.N;
  gemv_generic_finalize<<<(unsigned)((total + 255) / 256), 256, 0, a.stream>>>(partial, (int)grid.y, a.M, L.N, L.bias, out,
                                                                             a.ldy, a.n_offset);
  count_launch(2);
  return cudaGetLastError();
}

}  // namespace b200q
