// Launch-chain floor: per-launch time of trivial kernels replayed from a CUDA graph (200 launches),
// plain / with programmatic dependent launch / with thread-block clusters / both.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_plain(float* p) { if (p && threadIdx.x == 9999) p[0] = 1.f; }
__global__ void k_pdl(float* p) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p && threadIdx.x == 9999) p[0] = 1.f;
}
__global__ void k_cluster_sync(float* p) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (p && threadIdx.x == 9999) p[0] = 1.f;
}

template <class K>
float run(K kern, int grid, int block, int cluster, bool pdl, size_t smem) {
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  auto launch = [&]() {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2]; int n = 0;
    if (pdl) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    if (cluster > 0) { at[n].id = cudaLaunchAttributeClusterDimension; at[n].val.clusterDim.x = cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n; }
    cfg.attrs = at; cfg.numAttrs = n;
    cudaLaunchKernelEx(&cfg, kern, (float*)nullptr);
  };
  for (int i = 0; i < 10; ++i) launch();
  cudaStreamSynchronize(st);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < 200; ++i) launch();
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / (5 * 200);
}

int main() {
  for (int grid : {128, 256, 512}) {
    printf("grid=%3d x256thr: plain %.2f us | pdl %.2f us | cluster8 %.2f us | pdl+cluster8 %.2f us | pdl+cluster4 %.2f us | pdl+cluster8+sync %.2f us | pdl+64KB smem %.2f us\n", grid,
           run(k_plain, grid, 256, 0, false, 0), run(k_pdl, grid, 256, 0, true, 0), run(k_plain, grid, 256, 8, false, 0),
           run(k_pdl, grid, 256, 8, true, 0), run(k_pdl, grid, 256, 4, true, 0), run(k_cluster_sync, grid, 256, 8, true, 0),
           run(k_pdl, grid, 256, 0, true, 64 * 1024));
  }
  return 0;
}
