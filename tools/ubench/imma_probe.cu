// Micro-probe: legacy integer tensor path on sm_100a, mma.sync.m16n8k32.s32.u8.s8 (SASS IMMA.16832.U8.S8): issue rate vs
// the fp16 m16n8k16 path, with 1 / 4 / 8 independent accumulators and 4..32 warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_probe imma_probe.cu && ./imma_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void hmma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NACC>
__global__ void rate_i(uint32_t aval, uint32_t bval, int iters, long long* out, int* sink) {
  int acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0;
  uint32_t a = aval + (threadIdx.x & 1), b = bval;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) { imma(acc[i], a, a ^ 0x01010101u, a, a ^ 0x01010101u, b, b); a += 0x00010001u & (uint32_t)it; }
  }
  long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 123456789) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
template <int NACC>
__global__ void rate_h(uint32_t aval, uint32_t bval, int iters, long long* out, float* sink) {
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  uint32_t a = aval + (threadIdx.x & 1), b = bval;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) { hmma(acc[i], a, a ^ 0x00010001u, a, a ^ 0x00010001u, b, b); a += 0x00010001u & (uint32_t)it; }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int NACC>
void run(int threads) {
  long long* d; int* sink; long long hi, hh;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) rate_i<NACC><<<148, threads>>>(0x03070b0fu, 0x01020304u, iters, d, sink);
  cudaDeviceSynchronize();
  cudaMemcpy(&hi, d, 8, cudaMemcpyDeviceToHost);
  for (int rep = 0; rep < 2; ++rep) rate_h<NACC><<<148, threads>>>(0x00070003u, 0x3C003C00u, iters, d, (float*)sink);
  cudaDeviceSynchronize();
  cudaMemcpy(&hh, d, 8, cudaMemcpyDeviceToHost);
  const double pi = (double)hi / (iters * NACC), ph = (double)hh / (iters * NACC);
  const int wps = (threads / 32 + 3) / 4;
  printf("acc=%d warps/SM=%2d : IMMA.16832 %6.2f cyc/warp (%5.2f per SMSP, %6.1f weights/clk/SM) | HMMA.16816 %6.2f cyc/warp (%5.2f per SMSP, %6.1f weights/clk/SM)\n",
         NACC, threads / 32, pi, pi / wps, 512.0 * (threads / 32) / pi, ph, ph / wps, 256.0 * (threads / 32) / ph);
  cudaFree(d); cudaFree(sink);
}

int main() {
  for (int warps : {4, 8, 16, 32}) { run<1>(warps * 32); run<4>(warps * 32); run<8>(warps * 32); }
  // exactness of one IMMA: A = bytes 0..15, B = signed digits
  return 0;
}
