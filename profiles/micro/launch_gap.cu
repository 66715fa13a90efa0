// Micro-benchmark: per-node cost of dependent kernel chains inside a CUDA graph on this GPU.
//   nvcc -arch=sm_100a -O3 -o launch_gap launch_gap.cu && ./launch_gap
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
__global__ void k_empty(float* p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f;
}
__global__ void k_chain(float* p, int n) {  // one dependent global load->store chain per CTA
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float v = p[blockIdx.x * 32 + (threadIdx.x & 31)];
  for (int i = 0; i < n; ++i) v = p[((int)v & 1023) + (threadIdx.x & 31)] + 1.f;
  if (threadIdx.x < 32) p[blockIdx.x * 32 + threadIdx.x] = v * 0.f;
}
static float run(const char* name, int nodes, bool pdl, int cluster, size_t smem, int grid, int mode, float* buf) {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < nodes; ++i) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2]; int na = 0;
    if (cluster > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
    if (pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    cfg.attrs = at; cfg.numAttrs = na;
    if (mode == 0) cudaLaunchKernelEx(&cfg, k_empty, buf);
    else cudaLaunchKernelEx(&cfg, k_chain, buf, mode);
  }
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(a, st);
  for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(b, st); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  float us = 1e3f * ms / (5 * nodes);
  printf("%-58s %7.2f us/node  (%s)\n", name, us, cudaGetErrorString(cudaGetLastError()));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
  return us;
}
int main() {
  float* buf; cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 0, 1 << 20);
  const int N = 1000;
  run("empty  grid=1            no PDL", N, false, 1, 0, 1, 0, buf);
  run("empty  grid=1            PDL", N, true, 1, 0, 1, 0, buf);
  run("empty  grid=128          no PDL", N, false, 1, 0, 128, 0, buf);
  run("empty  grid=128          PDL", N, true, 1, 0, 128, 0, buf);
  run("empty  grid=128 smem=75K PDL", N, true, 1, 75 * 1024, 128, 0, buf);
  run("empty  grid=128 cluster=4 smem=75K PDL", N, true, 4, 75 * 1024, 128, 0, buf);
  run("empty  grid=384 cluster=3 smem=75K PDL", N, true, 3, 75 * 1024, 384, 0, buf);
  run("empty  grid=128 smem=200K PDL", N, true, 1, 200 * 1024, 128, 0, buf);
  run("chain1 grid=128          no PDL (1 dependent L2 load)", N, false, 1, 0, 128, 1, buf);
  run("chain1 grid=128          PDL", N, true, 1, 0, 128, 1, buf);
  run("chain4 grid=128          PDL (4 dependent L2 loads)", N, true, 1, 0, 128, 4, buf);
  run("chain4 grid=16           PDL", N, true, 1, 0, 16, 4, buf);
  return 0;
}
