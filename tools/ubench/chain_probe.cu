// Probes behind the round-2 decode design: a persistent one-CTA-per-SM kernel that streams the
// packed weights of a CHAIN of layers through one shared-memory ring with TMA, crossing layer boundaries without
// draining, with a grid-wide barrier per layer.
//   (1) grid barrier latency: red.release.gpu + ld.acquire.gpu poll, 148 CTAs, back to back;
//   (2) weight-stream rate: 3-D TMA boxes (32 words x 2 halves x 32 packed rows = 8 KB, 128-byte swizzle) of
//       [K/8, N] int32 matrices, contiguous slab ranges per CTA, ring of S slots, consumers that only read the slot
//       (LDS.128 over all of it) -- with and without a grid barrier between layers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../qllm_b200/csrc -I../../include chain_probe.cu -o chain_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

using namespace b200q;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void grid_arrive(unsigned int* ctr) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void barrier_probe(unsigned int* ctr, int iters, long long* out) {
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    __syncthreads();
    if (threadIdx.x == 0) {
      grid_arrive(ctr);
      const unsigned int target = (unsigned int)(i + 1) * gridDim.x;
      while (ld_acquire(ctr) < target) {}
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

struct StreamParams {
  const CUtensorMap* maps;     // one per layer (device memory)
  int layers, tiles, kc;       // slabs per layer = tiles * kc
  int slots, barrier, nwarps;
  unsigned int* ctr;
  unsigned long long* sink;
};

// warp 0: producer; warps 1..nwarps: consumers (slab q -> warp q % nwarps)
__global__ void __launch_bounds__(576, 1) stream_probe(const __grid_constant__ StreamParams p) {
  extern __shared__ __align__(1024) char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 32;
  char* ring = smem + 1024;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < p.slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int U = p.tiles * p.kc, n = gridDim.x, c = blockIdx.x;
  const int a = (int)((long long)c * U / n), b = (int)((long long)(c + 1) * U / n);
  if (warp == 0) {
    if (lane == 0) {
      int q = 0;
      for (int l = 0; l < p.layers; ++l) {
        const CUtensorMap* m = p.maps + l;
        for (int i = a; i < b; ++i, ++q) {
          const int slot = q % p.slots, round = q / p.slots;
          if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
          const int tile = i / p.kc, kc = i - tile * p.kc;
          mbar_expect_tx(&full[slot], 8192);
          tma_load_3d(ring + slot * 8192, m, 0, 2 * tile, 32 * kc, &full[slot]);
        }
      }
    }
  } else {
    const int w = warp - 1;
    uint32_t acc = 0;
    int q = 0;
    for (int l = 0; l < p.layers; ++l) {
      if (p.barrier && l > 0) {
        if (lane == 0) { const unsigned int target = (unsigned int)l * n; while (ld_acquire(p.ctr) < target) {} }
        __syncwarp();
      }
      for (int i = a; i < b; ++i, ++q) {
        if (q % p.nwarps != w) continue;
        const int slot = q % p.slots, round = q / p.slots;
        mbar_wait(&full[slot], round & 1);
        const uint4* s = reinterpret_cast<const uint4*>(ring + slot * 8192);
#pragma unroll
        for (int j = 0; j < 16; ++j) { const uint4 v = s[lane + 32 * j]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
      if (p.barrier) {
        asm volatile("bar.sync 1, %0;" ::"r"(p.nwarps * 32) : "memory");
        if (w == 0 && lane == 0) { __threadfence(); grid_arrive(p.ctr); }
      }
    }
    if (acc == 0x12345678u) p.sink[0] = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned int* ctr;
  long long* out;
  CK(cudaMalloc(&ctr, 256));
  CK(cudaMalloc(&out, sms * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  // ---- (1) grid barrier
  for (int threads : {32, 512}) {
    CK(cudaMemset(ctr, 0, 256));
    const int iters = 2000;
    barrier_probe<<<sms, threads>>>(ctr, 10, out);
    CK(cudaMemset(ctr, 0, 256));
    cudaEventRecord(e0);
    barrier_probe<<<sms, threads>>>(ctr, iters, out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("grid barrier, %d CTAs x %d threads: %.3f us per barrier\n", sms, threads, ms * 1e3 / iters);
  }
  // ---- (2) stream
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  struct Shape { int K, N; const char* name; } shapes[] = {{4096, 4096, "4096x4096"}, {4096, 12288, "4096x12288 (q|k|v)"},
                                                            {4096, 22016, "4096x22016 (gate|up)"}, {11008, 4096, "11008x4096"}};
  const size_t total_target = (size_t)1500 << 20;
  for (auto& sh : shapes) {
    const size_t bytes = (size_t)sh.K / 8 * sh.N * 4;
    const int layers = (int)(total_target / bytes);
    char* w;
    CK(cudaMalloc(&w, bytes * layers));
    CK(cudaMemset(w, 1, bytes * layers));
    std::vector<CUtensorMap> maps(layers);
    for (int l = 0; l < layers; ++l) {
      cuuint64_t dims[3] = {32, (cuuint64_t)sh.N / 32, (cuuint64_t)sh.K / 8};
      cuuint64_t strides[2] = {128, (cuuint64_t)sh.N * 4};
      cuuint32_t box[3] = {32, 2, 32};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, w + bytes * l, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    CUtensorMap* dmaps;
    CK(cudaMalloc(&dmaps, sizeof(CUtensorMap) * layers));
    CK(cudaMemcpy(dmaps, maps.data(), sizeof(CUtensorMap) * layers, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(stream_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int slots : {8, 16, 22}) {
      for (int barrier : {0, 1}) {
        for (int nw : {8, 16}) {
          StreamParams p = {dmaps, layers, sh.N / 64, sh.K / 256, slots, barrier, nw, ctr, (unsigned long long*)out};
          CK(cudaMemset(ctr, 0, 256));
          stream_probe<<<sms, 32 * (nw + 1), 1024 + 1024 + slots * 8192>>>(p);     // warm-up
          CK(cudaDeviceSynchronize());
          CK(cudaMemset(ctr, 0, 256));
          cudaEventRecord(e0);
          stream_probe<<<sms, 32 * (nw + 1), 1024 + 1024 + slots * 8192>>>(p);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          printf("stream %-22s layers=%3d slots=%2d barrier=%d warps=%2d: %.3f ms  %.0f GB/s  %.2f us/layer\n", sh.name, layers, slots,
                 barrier, nw, ms, bytes * layers / (ms * 1e-3) / 1e9, ms * 1e3 / layers);
        }
      }
    }
    CK(cudaFree(w));
    CK(cudaFree(dmaps));
  }
  return 0;
}
