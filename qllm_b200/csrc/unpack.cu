// Element-wise unpack / dequant kernels for every layout and bit width (HBM-bound streaming).
//   b200q_unpack  : packed -> int32 q[K,N], z[G,N]   (bit-exact gate vs compress_weight.py:87-92,
//                   quant_linear_awq.py:76-93; adds the Marlin inverse the reference lacks)
//   b200q_dequant : packed -> fp16 W[K,N]            (replaces ort_ops.dequant, ort_ops.cc:58-92)
// One thread produces 8 consecutive columns of one row so that global stores are 16-byte
// (fp16) / 32-byte (int32) vectors and packed loads along N are coalesced.
#include "common.cuh"
#include "kernels.h"

namespace b200q {

template <bool kDequant>
__global__ void __launch_bounds__(256) unpack_kernel(LayerView L, int32_t* __restrict__ q_out, __half* __restrict__ w_out) {
  const int n8 = L.N >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)L.K * n8) return;
  const int k = (int)(idx / n8);
  const int n0 = (int)(idx % n8) << 3;
  const int g = group_of(L, k);
  if (kDequant) {
    __align__(16) __half w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = dequant_one(L, load_q(L, k, n0 + i), load_z(L, g, n0 + i), load_s(L, g, n0 + i));
    *reinterpret_cast<uint4*>(w_out + (size_t)k * L.N + n0) = *reinterpret_cast<const uint4*>(w);
  } else {
    __align__(16) int32_t q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = (int32_t)load_q(L, k, n0 + i);
    int4* dst = reinterpret_cast<int4*>(q_out + (size_t)k * L.N + n0);
    dst[0] = *reinterpret_cast<const int4*>(q);
    dst[1] = *reinterpret_cast<const int4*>(q + 4);
  }
}

__global__ void __launch_bounds__(256) unpack_zeros_kernel(LayerView L, int32_t* __restrict__ z_out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)L.G * L.N) return;
  const int g = (int)(idx / L.N), n = (int)(idx % L.N);
  z_out[idx] = (int32_t)load_z(L, g, n);
}

// One-time integer re-layout (exact): any 4-bit layout -> K-packed GPTQ words (8 k per word, N-contiguous),
// GPTQ qzeros (8 columns per word) and natural-order fp16 scales.  One thread per output word.
__global__ void __launch_bounds__(256) repack_gptq4_kernel(LayerView L, uint32_t* __restrict__ qw_out, uint32_t* __restrict__ qz_out,
                                                           __half* __restrict__ s_out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nw = (size_t)(L.K >> 3) * L.N;
  if (idx < nw) {
    const int kw = (int)(idx / L.N), n = (int)(idx % L.N);
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) w |= (load_q(L, 8 * kw + i, n) & 0xFu) << (4 * i);
    qw_out[idx] = w;
  }
  const size_t nz = (size_t)L.G * (L.N >> 3);
  if (idx < nz) {
    const int g = (int)(idx / (L.N >> 3)), c = (int)(idx % (L.N >> 3));
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) w |= ((uint32_t)load_z(L, g, 8 * c + i) & 0xFu) << (4 * i);
    qz_out[idx] = w;
  }
  const size_t ns = (size_t)L.G * L.N;
  if (idx < ns) s_out[idx] = __float2half_rn(load_s(L, (int)(idx / L.N), (int)(idx % L.N)));
}

cudaError_t launch_repack_gptq4(const LayerView& L, uint32_t* qw_out, uint32_t* qz_out, __half* s_out, cudaStream_t st) {
  const size_t total = (size_t)(L.K >> 3) * L.N;      // >= G*N/8; scales G*N <= total when group >= 8
  const size_t ns = (size_t)L.G * L.N;
  const size_t m = total > ns ? total : ns;
  repack_gptq4_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(L, qw_out, qz_out, s_out);
  count_launch();
  return cudaGetLastError();
}

// The inverse re-layout (exact): K-packed GPTQ words -> AWQ-GEMM (target 1) or Marlin (target 2) buffers, i.e. what
// WQLinear_GEMM.pack / QuantLinearMarlin.pack would have produced from the same integers (quant_linear_awq.py:95-140,
// quant_linear_marlin.py:95-137).  One thread per element; nibbles are OR-ed into zero-initialised outputs.
__global__ void __launch_bounds__(256) repack_from_gptq4_kernel(LayerView L, int target, uint32_t* __restrict__ qw_out,
                                                                uint32_t* __restrict__ qz_out, __half* __restrict__ s_out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (size_t)L.K * L.N) {
    const int k = (int)(idx / L.N), n = (int)(idx % L.N);
    const uint32_t q = load_q(L, k, n) & 0xFu;
    if (target == B200Q_LAYOUT_AWQ_GEMM) {
      atomicOr(qw_out + (size_t)k * (L.N >> 3) + (n >> 3), q << (4 * awq_nibble_of_col(n & 7)));
    } else if (target == B200Q_LAYOUT_AWQ_GEMV) {
      atomicOr(qw_out + (size_t)n * (L.K >> 3) + (k >> 3), q << (4 * (k & 7)));
    } else if (target == B200Q_LAYOUT_ORT) {                              // byte n K/2 + k/2 -> word (little endian), nibble k & 7
      atomicOr(qw_out + (size_t)n * (L.K >> 3) + (k >> 3), q << (4 * (k & 7)));
    } else {
      size_t word; int nib;
      marlin_locate(k, n, L.N, word, nib);
      atomicOr(qw_out + word, q << (4 * nib));
    }
  }
  if (idx < (size_t)L.G * L.N) {
    const int g = (int)(idx / L.N), n = (int)(idx % L.N);
    if (target == B200Q_LAYOUT_AWQ_GEMM) {
      atomicOr(qz_out + (size_t)g * (L.N >> 3) + (n >> 3), ((uint32_t)load_z(L, g, n) & 0xFu) << (4 * awq_nibble_of_col(n & 7)));
      s_out[idx] = L.s[idx];
    } else if (target == B200Q_LAYOUT_AWQ_GEMV) {
      atomicOr(qz_out + (size_t)n * L.zw + (g >> 3), ((uint32_t)load_z(L, g, n) & 0xFu) << (4 * (g & 7)));
      s_out[(size_t)n * (8 * L.zw) + g] = L.s[idx];
    } else if (target == B200Q_LAYOUT_ORT) {                              // nibble index n * 2 ceil(G/2) + g of a byte stream
      const size_t nib = (size_t)n * (2 * ((L.G + 1) >> 1)) + g;
      atomicOr(qz_out + (nib >> 3), ((uint32_t)load_z(L, g, n) & 0xFu) << (4 * (nib & 7)));
      s_out[(size_t)n * L.G + g] = L.s[idx];
    } else {
      s_out[(size_t)g * L.N + marlin_scale_index(n, L.group == L.K)] = L.s[idx];
    }
  }
}

cudaError_t launch_repack_from_gptq4(const LayerView& L, int target, uint32_t* qw_out, uint32_t* qz_out, __half* s_out, cudaStream_t st) {
  const size_t words = (size_t)(L.K >> 3) * L.N;                          // same word count in all three 4-bit layouts
  cudaError_t e = cudaMemsetAsync(qw_out, 0, words * 4, st);
  if (e != cudaSuccess) return e;
  if (target == B200Q_LAYOUT_AWQ_GEMM) {
    e = cudaMemsetAsync(qz_out, 0, (size_t)L.G * (L.N >> 3) * 4, st);
    if (e != cudaSuccess) return e;
  } else if (target == B200Q_LAYOUT_ORT) {                                // N ceil(G/2) bytes, rounded up to whole words
    e = cudaMemsetAsync(qz_out, 0, (((size_t)L.N * ((L.G + 1) >> 1)) + 3) / 4 * 4, st);
    if (e != cudaSuccess) return e;
  } else if (target == B200Q_LAYOUT_AWQ_GEMV) {                           // rows are padded to ZW words / 8 ZW scales
    e = cudaMemsetAsync(qz_out, 0, (size_t)L.N * L.zw * 4, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(s_out, 0, (size_t)L.N * L.zw * 8 * 2, st);
    if (e != cudaSuccess) return e;
  }
  const size_t total = (size_t)L.K * L.N;
  repack_from_gptq4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L, target, qw_out, qz_out, s_out);
  count_launch();
  return cudaGetLastError();
}

// Act-order re-layout: packed row-block kw of the output holds original rows perm[P kw .. P kw + P - 1] (P = 32 / bits).
__global__ void __launch_bounds__(256) repack_actorder_kernel(LayerView L, const int* __restrict__ perm, uint32_t* __restrict__ qw_out) {
  const int P = 32 / L.bits;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)(L.K / P) * L.N) return;
  const int kw = (int)(idx / L.N), n = (int)(idx % L.N);
  const uint32_t mask = (1u << L.bits) - 1u;
  uint32_t w = 0;
  for (int i = 0; i < P; ++i) w |= (load_q(L, perm[P * kw + i], n) & mask) << (L.bits * i);
  qw_out[idx] = w;
}

// 3 / 5 / 6 / 7-bit streams: 32 consecutive rows fill exactly `bits` words (compress_weight.py:27-43); one thread re-packs one
// such block of one column, LSB first.
__global__ void __launch_bounds__(256) repack_actorder_anybit_kernel(LayerView L, const int* __restrict__ perm, uint32_t* __restrict__ qw_out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)(L.K / 32) * L.N) return;
  const int kb = (int)(idx / L.N), n = (int)(idx % L.N);
  const uint32_t mask = (1u << L.bits) - 1u;
  unsigned long long acc = 0;
  int have = 0, w = 0;
  for (int i = 0; i < 32; ++i) {
    acc |= (unsigned long long)(load_q(L, perm[32 * kb + i], n) & mask) << have;
    have += L.bits;
    if (have >= 32) {
      qw_out[((size_t)kb * L.bits + w) * L.N + n] = (uint32_t)acc;
      acc >>= 32; have -= 32; ++w;
    }
  }
}

cudaError_t launch_repack_actorder(const LayerView& L, const int* perm, uint32_t* qw_out, cudaStream_t st) {
  if (32 % L.bits != 0) {
    const size_t blocks = (size_t)(L.K / 32) * L.N;
    repack_actorder_anybit_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, st>>>(L, perm, qw_out);
    count_launch();
    return cudaGetLastError();
  }
  const size_t total = (size_t)(L.K * L.bits / 32) * L.N;
  repack_actorder_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L, perm, qw_out);
  count_launch();
  return cudaGetLastError();
}

// x_out[m, j] = x[m, perm[j]]: the activation half of the act-order re-layout for the kernels that stage x with bulk copies
__global__ void __launch_bounds__(256) gather_x_kernel(const __half* __restrict__ x, int64_t ldx, const int* __restrict__ perm,
                                                       __half* __restrict__ out, int M, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * (K >> 1)) return;
  const int m = (int)(idx / (K >> 1)), j = (int)(idx % (K >> 1)) << 1;
  const int2 pi = *reinterpret_cast<const int2*>(perm + j);
  const __half* row = x + (size_t)m * ldx;
  *reinterpret_cast<__half2*>(out + (size_t)m * K + j) = __halves2half2(row[pi.x], row[pi.y]);
}

cudaError_t launch_gather_x(const __half* x, int64_t ldx, const int* perm, __half* out, int M, int K, cudaStream_t st) {
  const size_t total = (size_t)M * (K >> 1);
  gather_x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, ldx, perm, out, M, K);
  count_launch();
  return cudaGetLastError();
}

// Fallback halves of the fused neighbours for kernels that do not carry them: out = silu(x) * x_mul ahead of the kernel,
// y += residual behind it (the integer decode kernel and the tcgen05 GEMM fuse them instead).
__global__ void __launch_bounds__(256) silu_mul_kernel(const __half* __restrict__ x, const __half* __restrict__ xm, int64_t ldx,
                                                       __half* __restrict__ out, int64_t M, int K) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * K) return;
  const int64_t m = idx / K;
  const int k = (int)(idx - m * K);
  out[idx] = __float2half_rn(silu_mul_f16(__half2float(x[m * ldx + k]), __half2float(xm[m * ldx + k])));
}
cudaError_t launch_silu_mul(const __half* x, const __half* x_mul, int64_t ldx, __half* out, int64_t M, int K, cudaStream_t st) {
  const int64_t total = M * K;
  silu_mul_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, x_mul, ldx, out, M, K);
  count_launch();
  return cudaGetLastError();
}
__global__ void __launch_bounds__(256) residual_add_kernel(__half* __restrict__ y, int64_t ldy, const __half* __restrict__ res, int64_t ldres,
                                                           int64_t M, int N) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int64_t m = idx / N;
  const int n = (int)(idx - m * N);
  y[m * ldy + n] = __float2half_rn(__half2float(y[m * ldy + n]) + __half2float(res[m * ldres + n]));
}
cudaError_t launch_residual_add(__half* y, int64_t ldy, const __half* res, int64_t ldres, int64_t M, int N, cudaStream_t st) {
  const int64_t total = M * N;
  residual_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(y, ldy, res, ldres, M, N);
  count_launch();
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) bf16_in_kernel(const unsigned short* __restrict__ x, const unsigned short* __restrict__ xm, int64_t ldx,
                                                      __half* __restrict__ out, int64_t M, int K) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * K) return;
  const int64_t m = idx / K;
  const int k = (int)(idx - m * K);
  const float g = bf16_bits_to_float(x[m * ldx + k]);
  out[idx] = xm ? __float2half_rn(silu_mul_bf16_to_f16(g, bf16_bits_to_float(xm[m * ldx + k]))) : __float2half_rn(g);
}
cudaError_t launch_bf16_in(const void* x, const void* x_mul, int64_t ldx, __half* out, int64_t M, int K, cudaStream_t st) {
  const int64_t total = M * K;
  bf16_in_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const unsigned short*)x, (const unsigned short*)x_mul, ldx, out, M, K);
  count_launch();
  return cudaGetLastError();
}
__global__ void __launch_bounds__(256) bf16_out_kernel(const __half* __restrict__ y16, unsigned short* __restrict__ y, int64_t ldy,
                                                       const unsigned short* __restrict__ res, int64_t ldres, int64_t M, int N) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int64_t m = idx / N;
  const int n = (int)(idx - m * N);
  float v = round_bf16(__half2float(y16[idx]));
  if (res) v = round_bf16(v + bf16_bits_to_float(res[m * ldres + n]));
  y[m * ldy + n] = (unsigned short)float_to_bf16_bits(v);
}
cudaError_t launch_bf16_out(const __half* y16, void* y, int64_t ldy, const void* res, int64_t ldres, int64_t M, int N, cudaStream_t st) {
  const int64_t total = M * N;
  bf16_out_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(y16, (unsigned short*)y, ldy, (const unsigned short*)res, ldres, M, N);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_unpack(const LayerView& L, int32_t* q_out, int32_t* z_out, cudaStream_t st) {
  const size_t total = (size_t)L.K * (L.N >> 3);
  unpack_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L, q_out, nullptr);
  count_launch();
  if (z_out && L.layout != B200Q_LAYOUT_HQQ) {
    const size_t tz = (size_t)L.G * L.N;
    unpack_zeros_kernel<<<(unsigned)((tz + 255) / 256), 256, 0, st>>>(L, z_out);
    count_launch();
  }
  return cudaGetLastError();
}

cudaError_t launch_dequant(const LayerView& L, __half* w_out, cudaStream_t st) {
  const size_t total = (size_t)L.K * (L.N >> 3);
  unpack_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L, nullptr, w_out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace b200q
