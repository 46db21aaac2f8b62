// tcgen05.mma issue-rate probe.  Question: the prefill GEMM retires one 128 x 128 x 16 kind::f16 MMA
// per ~110 cycles with its operands ready -- is that the TS form (A in tensor memory), the second issuer, or contention
// from what the rest of the CTA does (tcgen05.st of the next A stages, TMA-like writes into shared memory, LDS)?
// Round-1 answer (profiles/r1_ubench_umma_rate_probe.txt): none of those -- TS = SS = 67.8 cycles (floor 64), two
// issuers 64.1, and tcgen05.st / LDS / cp.async / 16 KB bulk-copy traffic move it by at most one cycle.  Still open
// (noise 6, not yet run): whether 24 ALU-saturated warps starve the single issuing thread of issue slots.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../qllm_b200/csrc -I../../include \
//        umma_rate_probe.cu -o umma_rate_probe && ./umma_rate_probe
//
// Each CTA (one per SM) issues `reps` k-blocks of four MMAs (K = 64) from one or two threads into one or two
// accumulators, commits every k-block to an mbarrier (as the GEMM does) and reports cycles per MMA.  Operands are
// zero-filled (the timing does not depend on the values).  Floor: M = 128, N columns -> N / 2 cycles per MMA.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "common.cuh"

using namespace b200q;

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
               "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 B apart (same descriptor as gemm_tcgen05.cu)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// noise: 0 none | 1 tcgen05.st into spare TMEM columns | 2 LDS+STS over 32 KB | 3 cp.async 16 B global -> shared
//        4 two threads keep 16 KB bulk copies (cp.async.bulk, the TMA data path) in flight into shared memory
//        5 = 4 plus four warps of LDS.128 (what the GEMM's X/W TMA traffic and dequant reads look like together)
//        6 every non-issuing warp of a 28-warp CTA spins on lop3 / hfma2 chains (the dequant teams' issue pressure)
template <int N, bool TS>
__global__ void __launch_bounds__(896, 1) probe(unsigned long long* out, int reps, int issuers, int noise, const uint4* gsrc) {
  extern __shared__ __align__(1024) char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  char* btile = smem;                       // N rows x 64 k fp16 = N * 128 B (<= 32 KB)
  char* atile = smem + 32 * 1024;           // 128 rows x 64 k fp16 = 16 KB (SS form)
  char* scratch = smem + 48 * 1024;         // 32 KB for the LDS / cp.async noise
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 80 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  volatile int* done = reinterpret_cast<volatile int*>(slot + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 80 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar[0], (uint32_t)reps); mbar_init(&bar[1], (uint32_t)reps); mbar_init(&bar[2], 1); mbar_init(&bar[3], 1); *done = 0; fence_mbar_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();                      // zero fill (generic proxy) visible to the tensor core's shared-memory reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;

  // issuers sit in warps 4, 5 -- or, with 28 warps (noise 6), in the two highest warps as in the GEMM
  const int w_iss = (noise == 6) ? (int)(blockDim.x >> 5) - issuers : 4;
  if (warp >= w_iss && warp < w_iss + issuers && lane == 0) {
    const int me = warp - w_iss;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(btile)), adesc = umma_desc_k_sw128(smem_u32(atile));
    const uint32_t d = tmem + me * N;       // accumulators: columns [0, 2 N) (N <= 128 with two issuers)
    const uint32_t a_src = tmem + 384;      // A stage in TMEM: 32 columns = 64 k
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (TS) tc_mma_ts(d, a_src + j * 8, bdesc + (uint64_t)(2 * j), idesc, 1u);
        else tc_mma_ss(d, adesc + (uint64_t)(2 * j), bdesc + (uint64_t)(2 * j), idesc, 1u);
      }
      tc_commit(&bar[me]);
    }
    mbar_wait(&bar[me], 0);
    const long long t1 = clock64();
    out[(size_t)blockIdx.x * 2 + me] = (unsigned long long)(t1 - t0);
    __threadfence_block();
    atomicAdd(const_cast<int*>(done), 1);
  } else if (noise == 6) {
    if (warp < w_iss) {                                   // ALU pressure: dependent lop3 / hfma2 chains, four per iteration
      uint32_t a = tid, b = 0x3c003c00u;
      while (*done < issuers) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          a = (a & 0x000f000fu) | 0x64006400u;
          asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(b) : "r"(a));
          a ^= b >> 3;
        }
      }
      if (a == 0x12345678u) out[0] = b;
    }
  } else if (warp >= 6 && warp < 8 && noise >= 4) {
    if (lane == 0) {                                      // bulk-copy traffic: 16 KB per copy, one copy in flight per thread
      uint64_t* b = &bar[warp - 4];
      char* dst = scratch + (warp - 6) * 16384;
      uint32_t ph = 0;
      while (*done < issuers) {
        mbar_expect_tx(b, 16384u);
        bulk_g2s(dst, reinterpret_cast<const char*>(gsrc) + (warp - 6) * 65536, 16384u, b);
        mbar_wait(b, ph);
        ph ^= 1u;
      }
    }
  } else if (warp < 4 && noise == 5) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    int it = 0;
    while (*done < issuers) {
      const uint4* p = reinterpret_cast<const uint4*>(atile) + ((tid + 128 * it) & 1023);
      const uint4 v0 = p[0], v1 = p[128 & 1023 ? 128 : 0];
      acc.x ^= v0.x ^ v1.y;
      ++it;
    }
    if (acc.x == 0x12345678u) out[0] = 0;
  } else if (warp < 4 && noise != 0 && noise < 4) {
    uint32_t regs[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) regs[i] = (uint32_t)(tid + i);
    uint4 acc = make_uint4(0, 0, 0, 0);
    int it = 0;
    while (*done < issuers) {
      if (noise == 1) {
        tc_st16(tmem + ((uint32_t)(warp * 32) << 16) + 416 + 16 * (it & 1), regs);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else if (noise == 2) {
        uint4* p = reinterpret_cast<uint4*>(scratch) + ((tid + 128 * it) & 2047);
        const uint4 v = *p;
        acc.x ^= v.x;
        *p = acc;
      } else {
        const uint32_t dst = smem_u32(scratch) + (uint32_t)(((tid + 128 * it) & 2047) * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gsrc + ((tid + 128 * it) & 65535)) : "memory");
        if ((it & 7) == 7) asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 4;" ::: "memory");
      }
      ++it;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (acc.x == 0x12345678u) out[0] = 0;   // keep the loads
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

template <int N, bool TS>
static void run(const char* name, int grid, int reps, int issuers, int noise, unsigned long long* dout, const uint4* gsrc) {
  const int smem = 82 * 1024 + 1024;
  cudaFuncSetAttribute(probe<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(dout, 0, 148 * 2 * sizeof(unsigned long long));
  probe<N, TS><<<grid, noise == 6 ? 896 : 256, smem>>>(dout, reps, issuers, noise, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
  unsigned long long h[148 * 2];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int i = 0; i < grid; ++i)
    for (int m = 0; m < issuers; ++m) worst = worst > (double)h[i * 2 + m] ? worst : (double)h[i * 2 + m];
  // with two issuers both accumulators share the pipe: cycles per MMA of the CTA = span / (4 reps issuers)
  printf("%-44s grid %3d  N %3d  issuers %d  noise %d : %7.1f cycles / MMA (floor %d)\n", name, grid, N, issuers, noise,
         worst / (4.0 * reps * issuers), N / 2);
}

int main() {
  unsigned long long* dout;
  uint4* gsrc;
  cudaMalloc(&dout, 148 * 2 * sizeof(unsigned long long));
  cudaMalloc(&gsrc, 65536 * sizeof(uint4));
  cudaMemset(gsrc, 0, 65536 * sizeof(uint4));
  const int reps = 512;
  if (getenv("UMMA_PROBE_ALU")) {                         // round 2: issue-slot starvation by 26-27 ALU-bound warps
    run<128, true>("TS + 27 ALU-bound warps", 148, reps, 1, 6, dout, gsrc);
    run<128, true>("TS, two issuers + 26 ALU-bound warps", 148, reps, 2, 6, dout, gsrc);
    run<256, true>("TS, N = 256 + 27 ALU-bound warps", 148, reps, 1, 6, dout, gsrc);
    return 0;
  }
  if (getenv("UMMA_PROBE_HEAVY")) {                       // second call: the bulk-copy / LDS contention cases only
    run<128, true>("TS + bulk copies", 148, reps, 1, 4, dout, gsrc);
    run<128, true>("TS, two issuers + bulk copies", 148, reps, 2, 4, dout, gsrc);
    run<128, true>("TS + bulk copies + LDS", 148, reps, 1, 5, dout, gsrc);
    run<128, true>("TS, two issuers + bulk copies + LDS", 148, reps, 2, 5, dout, gsrc);
    run<128, false>("SS + bulk copies + LDS", 148, reps, 1, 5, dout, gsrc);
    return 0;
  }
  for (int grid : {1, 148}) {
    run<128, true>("TS (A in TMEM)", grid, reps, 1, 0, dout, gsrc);
    run<128, false>("SS (A in shared memory)", grid, reps, 1, 0, dout, gsrc);
    run<128, true>("TS, two issuers / two accumulators", grid, reps, 2, 0, dout, gsrc);
    run<64, true>("TS, N = 64", grid, reps, 1, 0, dout, gsrc);
    run<256, true>("TS, N = 256", grid, reps, 1, 0, dout, gsrc);
    run<128, true>("TS + tcgen05.st noise", grid, reps, 1, 1, dout, gsrc);
    run<128, true>("TS + LDS/STS noise", grid, reps, 1, 2, dout, gsrc);
    run<128, true>("TS + cp.async global->shared noise", grid, reps, 1, 3, dout, gsrc);
    run<128, true>("TS, two issuers + tcgen05.st noise", grid, reps, 2, 1, dout, gsrc);
  }
  return 0;
}
