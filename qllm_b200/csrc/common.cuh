// Shared device helpers for libb200q: layer view, element-wise format decoding, PTX wrappers.
// Format definitions follow SURVEY.md Appendix A (derived from the reference's codec:
// compress_weight.py:10-92, quant_linear_awq.py:95-140, quant_linear_marlin.py:18-42).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200q.h"

namespace b200q {

struct LayerView {
  int layout, bits, group, K, N, G, zero_bias;
  int zw;                            // AWQ-GEMV: words per qzeros row (calculate_zeros_width, quant_linear_awq.py:15-27)
  const uint32_t* __restrict__ qw;
  const void* __restrict__ qz;       // uint32 packed | half [G,N] (HQQ) | nullptr (Marlin)
  const __half* __restrict__ s;
  const int* __restrict__ g_idx;     // nullptr -> k / group
  const __half* __restrict__ bias;
  const int* __restrict__ x_perm;    // nullptr, or [K]: packed row j multiplies x[:, x_perm[j]] (act-order re-layout)
};

__host__ inline LayerView make_view(const b200q_layer* L) {
  LayerView v;
  v.layout = L->layout; v.bits = L->bits; v.group = L->group_size; v.K = L->K; v.N = L->N;
  v.G = (L->K + L->group_size - 1) / L->group_size; v.zero_bias = L->zero_bias;
  v.qw = (const uint32_t*)L->qweight; v.qz = L->qzeros; v.s = (const __half*)L->scales;
  v.g_idx = L->g_idx; v.bias = (const __half*)L->bias; v.x_perm = L->x_perm;
  const int mult = L->group_size >= 128 ? 1 : (L->group_size == 64 ? 2 : 4);
  v.zw = ((((v.G + 7) / 8) + mult - 1) / mult) * mult;
  return v;
}

// ---- element-wise decoders (generic, any bit width; used by unpack/dequant/generic GEMV) ----

// value r of an LSB-first bit-stream whose consecutive 32-bit words are `stride` words apart
__device__ __forceinline__ uint32_t bitstream_get(const uint32_t* __restrict__ base, size_t stride, int r, int bits) {
  const int pos = r * bits, w = pos >> 5, o = pos & 31;
  uint32_t v = __ldg(base + (size_t)w * stride) >> o;
  if (o + bits > 32) v |= __ldg(base + (size_t)(w + 1) * stride) << (32 - o);
  return v & ((1u << bits) - 1u);
}

__device__ __forceinline__ int awq_nibble_of_col(int i) {   // column 8c+i sits in nibble inv[i]
  return (i >> 1) + ((i & 1) << 2);                          // inv = {0,4,1,5,2,6,3,7}
}

// Marlin: word index and nibble holding element (k, n)  (closed form)
__device__ __forceinline__ void marlin_locate(int k, int n, int N, size_t& word, int& nib) {
  const int r = k >> 4, row = k & 15;
  const int blk = n >> 6, j = (n & 63) >> 4, colp = n & 15;
  const int col = colp & 7, block = colp >> 3;
  const int q = (row & 7) >> 1;
  const int ri = (row & 1) + ((row >> 3) << 1);
  const int vidx = ri + 4 * block;
  nib = awq_nibble_of_col(vidx);                             // nibble i holds v[{0,2,4,6,1,3,5,7}[i]]
  const int t = 4 * col + q;
  word = (size_t)r * (2 * (size_t)N) + 128 * blk + 4 * t + j;
}

__device__ __forceinline__ int marlin_scale_index(int n, bool per_channel) {
  if (!per_channel) return (n & ~63) + 8 * (n & 7) + ((n & 63) >> 3);
  const int n32 = n & 31, h = n32 >> 3, r = n32 & 7;
  return (n & ~31) + 8 * (r >> 1) + 2 * h + (r & 1);
}

__device__ __forceinline__ uint32_t load_q(const LayerView& L, int k, int n) {
  switch (L.layout) {
    case B200Q_LAYOUT_GPTQ:
    case B200Q_LAYOUT_HQQ:
      return bitstream_get(L.qw + n, (size_t)L.N, k, L.bits);
    case B200Q_LAYOUT_AWQ_GEMM: {
      const uint32_t w = __ldg(L.qw + (size_t)k * (L.N >> 3) + (n >> 3));
      return (w >> (4 * awq_nibble_of_col(n & 7))) & 0xFu;
    }
    case B200Q_LAYOUT_AWQ_GEMV:
      return (__ldg(L.qw + (size_t)n * (L.K >> 3) + (k >> 3)) >> (4 * (k & 7))) & 0xFu;
    case B200Q_LAYOUT_ORT: {   // blobs [N, K / group, group / 2] bytes == [N, K / 2] bytes (K % group == 0): byte k / 2, nibble k & 1
      const uint8_t b = __ldg(reinterpret_cast<const uint8_t*>(L.qw) + (size_t)n * (L.K >> 1) + (k >> 1));
      return (uint32_t)(b >> (4 * (k & 1))) & 0xFu;
    }
    default: {  // MARLIN
      size_t word; int nib;
      marlin_locate(k, n, L.N, word, nib);
      return (__ldg(L.qw + word) >> (4 * nib)) & 0xFu;
    }
  }
}

// zero point of (group g, column n) as float (integer-valued except HQQ float zeros)
__device__ __forceinline__ float load_z(const LayerView& L, int g, int n) {
  switch (L.layout) {
    case B200Q_LAYOUT_GPTQ: {
      const size_t zw = ((size_t)L.N * L.bits + 31) / 32;   // words per qzeros row
      const uint32_t z = bitstream_get((const uint32_t*)L.qz + (size_t)g * zw, 1, n, L.bits);
      return (float)((z + (uint32_t)L.zero_bias) & ((1u << L.bits) - 1u));
    }
    case B200Q_LAYOUT_HQQ:
      return __half2float(__ldg((const __half*)L.qz + (size_t)g * L.N + n));
    case B200Q_LAYOUT_AWQ_GEMM: {
      const uint32_t w = __ldg((const uint32_t*)L.qz + (size_t)g * (L.N >> 3) + (n >> 3));
      return (float)((w >> (4 * awq_nibble_of_col(n & 7))) & 0xFu);
    }
    case B200Q_LAYOUT_AWQ_GEMV:
      return (float)((__ldg((const uint32_t*)L.qz + (size_t)n * L.zw + (g >> 3)) >> (4 * (g & 7))) & 0xFu);
    case B200Q_LAYOUT_ORT: {
      const uint8_t b = __ldg(reinterpret_cast<const uint8_t*>(L.qz) + (size_t)n * ((L.G + 1) >> 1) + (g >> 1));
      return (float)((b >> (4 * (g & 1))) & 0xFu);
    }
    default:
      return 8.0f;
  }
}

__device__ __forceinline__ float load_s(const LayerView& L, int g, int n) {
  if (L.layout == B200Q_LAYOUT_MARLIN)
    return __half2float(__ldg(L.s + (size_t)g * L.N + marlin_scale_index(n, L.group == L.K)));
  if (L.layout == B200Q_LAYOUT_AWQ_GEMV) return __half2float(__ldg(L.s + (size_t)n * (8 * L.zw) + g));
  if (L.layout == B200Q_LAYOUT_ORT) return __half2float(__ldg(L.s + (size_t)n * L.G + g));
  return __half2float(__ldg(L.s + (size_t)g * L.N + n));
}

__device__ __forceinline__ int group_of(const LayerView& L, int k) {
  return L.g_idx ? __ldg(L.g_idx + k) : k / L.group;
}

// The engine's dequant rule: one rounding, fp16((q - z) * s).  (q - z) is exact for integer
// zeros; for HQQ float zeros it is first rounded to fp16, as the half2 fast paths do.
__device__ __forceinline__ __half dequant_one(const LayerView& L, uint32_t q, float z, float s) {
  float d = (float)q - z;
  if (L.layout == B200Q_LAYOUT_HQQ) d = __half2float(__float2half_rn(d));
  return __float2half_rn(d * s);
}

// act(gate) * up as the unfused torch ops compute it in fp16: silu evaluated in fp32 and rounded, the product rounded again
__device__ __forceinline__ float silu_mul_f16(float g, float u) {
  const float s = __half2float(__float2half_rn(g / (1.0f + expf(-g))));
  return __half2float(__float2half_rn(s * u));
}

// bfloat16 <-> float on raw bits (round to nearest even; NaN stays NaN)
__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }
__device__ __forceinline__ uint32_t float_to_bf16_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (u >> 16) | 0x40u;       // quiet NaN
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ float round_bf16(float f) { return bf16_bits_to_float(float_to_bf16_bits(f)); }
// a bf16 model's act(gate) * up feeding a QuantLinear: silu and the product rounded to bf16 (the torch ops), then the
// reference's cast of the layer input to fp16 (auto_cast, quant_linear_awq.py:29-36)
__device__ __forceinline__ float silu_mul_bf16_to_f16(float g, float u) {
  const float s = round_bf16(g / (1.0f + expf(-g)));
  return __half2float(__float2half_rn(round_bf16(s * u)));
}

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return r;
}
// (a & b) | c
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {
  return lop3<(0xF0 & 0xCC) | 0xAA>(a, b, c);
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ uint32_t hsub2_u(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("sub.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t hmul2_u(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t dup_half(__half h) {
  const uint32_t u = (uint32_t)__half_as_ushort(h);
  return u | (u << 16);
}

// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- mbarrier / bulk-copy (TMA engine, non-tensor form) ------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy, completion counted in bytes on `bar`. size % 16 == 0, 16B-aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Programmatic dependent launch: wait for the upstream kernel's memory to be visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace b200q
