// Micro-probe: mma.sync.m16n8k32 e4m3 x e4m3 -> f32 on sm_100a.
//   (1) issue rate vs the fp16 m16n8k16 path (is the legacy fp8 MMA full rate on Blackwell?);
//   (2) exactness: A = int4 nibbles q read as e4m3 bytes 0x0q (= q * 2^-9, sub-normal below 8), B = a 3-term e4m3 split of
//       fp16 activations in three B columns; the f32 accumulators must equal the integer dot products exactly.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f8_probe f8_probe.cu && ./f8_probe
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_f8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NACC, bool F8>
__global__ void rate(uint32_t aval, uint32_t bval, int iters, long long* out, float* sink) {
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  uint32_t a = aval + (threadIdx.x & 1), b = bval;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) { if (F8) mma_f8(acc[i], a, a, a, a, b, b); else mma_f16(acc[i], a, a, a, a, b, b); }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// exactness: one warp, D[16 x 8] += A[16 x 32] * B[32 x 8] over `chunks` k-chunks of 32.
// A[r][k] = q (0..15) given as bytes; B[k][c] = e4m3 bytes.  Fragment layouts (PTX ISA, m16n8k32 8-bit):
//   a0: row g,   k 4t..4t+3 | a1: row g+8, k 4t..4t+3 | a2: row g, k 16+4t.. | a3: row g+8, k 16+4t..
//   b0: k 4t..4t+3, col g   | b1: k 16+4t.., col g ;  d0,d1: row g, cols 2t,2t+1 | d2,d3: row g+8
__global__ void exact(const uint8_t* A, const uint8_t* B, int chunks, float* D) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < chunks; ++c) {
    const uint8_t* a = A + (size_t)c * 16 * 32;      // [16][32]
    const uint8_t* b = B + (size_t)c * 32 * 8;       // [32][8]
    auto pa = [&](int r, int k0) { return (uint32_t)a[r * 32 + k0] | ((uint32_t)a[r * 32 + k0 + 1] << 8) | ((uint32_t)a[r * 32 + k0 + 2] << 16) | ((uint32_t)a[r * 32 + k0 + 3] << 24); };
    auto pb = [&](int k0, int col) { return (uint32_t)b[k0 * 8 + col] | ((uint32_t)b[(k0 + 1) * 8 + col] << 8) | ((uint32_t)b[(k0 + 2) * 8 + col] << 16) | ((uint32_t)b[(k0 + 3) * 8 + col] << 24); };
    mma_f8(d, pa(g, 4 * t), pa(g + 8, 4 * t), pa(g, 16 + 4 * t), pa(g + 8, 16 + 4 * t), pb(4 * t, g), pb(16 + 4 * t, g));
  }
  D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1]; D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}

static float e4m3_to_float(uint8_t v) {
  const int s = v >> 7, e = (v >> 3) & 15, m = v & 7;
  float f = e == 0 ? ldexpf((float)m, -9) : ldexpf(1.f + m / 8.f, e - 7);
  return s ? -f : f;
}
static uint8_t float_to_e4m3(float x) {      // round to nearest even, saturate to 448
  __nv_fp8_e4m3 v(x);
  return *reinterpret_cast<uint8_t*>(&v);
}

template <int NACC, bool F8>
void run_rate(const char* name, int threads) {
  long long* d; float* sink; long long h;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) rate<NACC, F8><<<148, threads>>>(F8 ? 0x03070b0fu : 0x00070003u, F8 ? 0x38383838u : 0x3C003C00u, iters, d, sink);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * NACC);
  printf("%-26s acc=%d warps/SM=%2d : %6.2f cycles per MMA per warp (%5.2f per SMSP)  -> %6.1f weight-MACs/clk/SM at M=1\n", name, NACC, threads / 32,
         per, per / ((threads / 32 + 3) / 4), (F8 ? 512.0 : 256.0) * (threads / 32) / per);
  cudaFree(d); cudaFree(sink);
}

int main() {
  for (int warps : {4, 8, 16, 32}) {
    run_rate<8, false>("f16 m16n8k16 8 indep", warps * 32);
    run_rate<8, true>("e4m3 m16n8k32 8 indep", warps * 32);
    run_rate<1, true>("e4m3 m16n8k32 dependent", warps * 32);
  }
  // ---- exactness ----
  const int chunks = 128;      // K = 4096
  uint8_t* hA = (uint8_t*)malloc(chunks * 16 * 32), *hB = (uint8_t*)malloc(chunks * 32 * 8);
  double ref[16][8] = {};
  srand(1);
  float xs[4096];
  for (int k = 0; k < chunks * 32; ++k) {
    float u = (rand() / (float)RAND_MAX + rand() / (float)RAND_MAX + rand() / (float)RAND_MAX - 1.5f) * 2.f;   // ~N(0,1)-ish
    if (k % 97 == 0) u *= 40.f;                                                                                 // outliers
    xs[k] = __half2float(__float2half(u));
  }
  double maxabs_err = 0, maxabs_ref = 0;
  for (int c = 0; c < chunks; ++c) {
    // per-chunk power-of-two scale so that max |x| lands in [128, 256)
    float mx = 0.f;
    for (int k = 0; k < 32; ++k) mx = fmaxf(mx, fabsf(xs[c * 32 + k]));
    int E = 0;
    if (mx > 0.f) { int e; frexpf(mx, &e); E = 8 - e; }
    for (int k = 0; k < 32; ++k) {
      const float x = ldexpf(xs[c * 32 + k], E);
      const uint8_t t0 = float_to_e4m3(x);
      const float r1 = x - e4m3_to_float(t0);
      const uint8_t t1 = float_to_e4m3(r1);
      const float r2 = r1 - e4m3_to_float(t1);
      const uint8_t t2 = float_to_e4m3(r2);
      for (int col = 0; col < 8; ++col) hB[(c * 32 + k) * 8 + col] = 0;
      hB[(c * 32 + k) * 8 + 0] = t0; hB[(c * 32 + k) * 8 + 1] = t1; hB[(c * 32 + k) * 8 + 2] = t2;
      // column 3: the activation's own scale carrier is not needed; columns 3..7 stay zero
    }
    for (int r = 0; r < 16; ++r)
      for (int k = 0; k < 32; ++k) {
        const int q = rand() & 15;
        hA[(c * 16 + r) * 32 + k] = (uint8_t)q;
      }
  }
  uint8_t *dA, *dB; float* dD;
  cudaMalloc(&dA, chunks * 16 * 32); cudaMalloc(&dB, chunks * 32 * 8); cudaMalloc(&dD, 16 * 8 * 4);
  cudaMemcpy(dA, hA, chunks * 16 * 32, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, chunks * 32 * 8, cudaMemcpyHostToDevice);
  // (a) one chunk at a time: exact products, per-chunk scale undone on the host
  double tot[16] = {};
  double worst_chunk = 0;
  for (int c = 0; c < chunks; ++c) {
    exact<<<1, 32>>>(dA + c * 16 * 32, dB + c * 32 * 8, 1, dD);
    float hD[16 * 8];
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    float mx = 0.f;
    for (int k = 0; k < 32; ++k) mx = fmaxf(mx, fabsf(xs[c * 32 + k]));
    int E = 0;
    if (mx > 0.f) { int e; frexpf(mx, &e); E = 8 - e; }
    for (int r = 0; r < 16; ++r) {
      double want = 0, want_terms = 0;
      for (int k = 0; k < 32; ++k) {
        want += (double)hA[(c * 16 + r) * 32 + k] * xs[c * 32 + k];
        for (int j = 0; j < 3; ++j) want_terms += (double)hA[(c * 16 + r) * 32 + k] * e4m3_to_float(hB[(c * 32 + k) * 8 + j]);
      }
      const double got_scaled = ((double)hD[r * 8 + 0] + hD[r * 8 + 1] + hD[r * 8 + 2]) * 512.0;      // undo q * 2^-9
      worst_chunk = fmax(worst_chunk, fabs(got_scaled - want_terms));
      const double got = ldexp(got_scaled, -E);
      tot[r] += got;
      ref[r][0] += want;
    }
  }
  for (int r = 0; r < 16; ++r) { maxabs_err = fmax(maxabs_err, fabs(tot[r] - ref[r][0])); maxabs_ref = fmax(maxabs_ref, fabs(ref[r][0])); }
  printf("exactness, per-chunk MMAs : max |MMA - exact sum of the e4m3 terms| = %.3g (0 = products and f32 accumulation exact)\n", worst_chunk);
  printf("3-term e4m3 split, K=4096  : max |y - fp64 ref| = %.3g, max |ref| = %.3g -> rel %.2e\n", maxabs_err, maxabs_ref, maxabs_err / maxabs_ref);
  // (b) all 128 chunks accumulated inside the MMA accumulator (one common scale): checks long f32 accumulation chains
  {
    float mx = 0.f;
    for (int k = 0; k < chunks * 32; ++k) mx = fmaxf(mx, fabsf(xs[k]));
    int e; frexpf(mx, &e); const int E = 8 - e;
    for (int k = 0; k < chunks * 32; ++k) {
      const float x = ldexpf(xs[k], E);
      const uint8_t t0 = float_to_e4m3(x); const float r1 = x - e4m3_to_float(t0);
      const uint8_t t1 = float_to_e4m3(r1); const float r2 = r1 - e4m3_to_float(t1);
      hB[k * 8 + 0] = t0; hB[k * 8 + 1] = t1; hB[k * 8 + 2] = float_to_e4m3(r2);
    }
    cudaMemcpy(dB, hB, chunks * 32 * 8, cudaMemcpyHostToDevice);
    exact<<<1, 32>>>(dA, dB, chunks, dD);
    float hD[16 * 8];
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double worst = 0, mref = 0;
    for (int r = 0; r < 16; ++r) {
      double want = 0;
      for (int c = 0; c < chunks; ++c) for (int k = 0; k < 32; ++k) want += (double)hA[(c * 16 + r) * 32 + k] * xs[c * 32 + k];
      const double got = ldexp(((double)hD[r * 8] + hD[r * 8 + 1] + hD[r * 8 + 2]) * 512.0, -E);
      worst = fmax(worst, fabs(got - want)); mref = fmax(mref, fabs(want));
    }
    printf("one accumulator over K=4096, common scale: max |y - ref| = %.3g, max |ref| = %.3g -> rel %.2e\n", worst, mref, worst / mref);
  }
  return 0;
}
