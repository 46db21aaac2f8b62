// Micro-probe: mma.sync.m16n8k16 latency / throughput on this GPU, normal vs subnormal fp16 inputs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NACC>
__global__ void probe(uint32_t aval, uint32_t bval, int iters, long long* out, float* sink) {
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  uint32_t a = aval + (threadIdx.x & 1), b = bval;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) mma(acc[i], a, a, a, a, b, b);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

__global__ void hfma_probe(int iters, long long* out, float* sink) {
  uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t a = 0x3C003C00u + threadIdx.x, b = 0x38003800u;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= acc[i];
  if (s == 0x12345u) sink[0] = 1.f;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int NACC>
void run(const char* name, uint32_t aval, int threads, int blocks) {
  long long* d; float* sink; long long h;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  probe<NACC><<<blocks, threads>>>(aval, 0x3C003C00u, iters, d, sink);
  probe<NACC><<<blocks, threads>>>(aval, 0x3C003C00u, iters, d, sink);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s acc=%d warps/SM=%2d : %.2f cycles per MMA per warp  (%.2f cycles per MMA per SMSP)\n", name, NACC, threads / 32,
         (double)h / (iters * NACC), (double)h / (iters * NACC) / ((threads / 32 + 3) / 4));
  cudaFree(d); cudaFree(sink);
}

int main() {
  const uint32_t NORMAL = 0x3C003C00u, SUB = 0x00070003u;
  for (int warps : {1, 4, 8, 16}) {
    run<1>("normal dependent", NORMAL, warps * 32, 148);
    run<8>("normal 8 independent", NORMAL, warps * 32, 148);
    run<1>("subnormal dependent", SUB, warps * 32, 148);
    run<8>("subnormal 8 independent", SUB, warps * 32, 148);
  }
  long long* d; float* sink; long long h;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  for (int warps : {1, 4, 8, 16}) {
    hfma_probe<<<148, warps * 32>>>(4000, d, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("HFMA2 8 independent chains   warps/SM=%2d : %.2f cycles per HFMA2 per warp\n", warps, (double)h / (4000 * 8));
  }
  return 0;
}
