// umma_tf32_check.cu -- does tcgen05.mma kind::tf32 accept MN-major (transposed) shared-memory operands?
// One MMA, M = 128, N = 16, K = 8: D[m][n] = sum_k A[m][k] B[n][k] with A[m][k] = (m % 7) + k, B[n][k] = (n + 1) * (k == n % 8).
// Expected D[m][n] = ((m % 7) + n % 8) * (n + 1).  Operands are staged no-swizzle in both majors:
//   MN-major: [chunk of 4 MN elements][8 k rows][16 B]   (the layout csrc/sd_wgrad_tc.cuh wanted to use)
//   K-major : [chunk of 4 K elements][rows][16 B]
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I safe_dreamer_b200/csrc -o profiles/micro/umma_tf32_check profiles/micro/umma_tf32_check.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sd_tc.cuh"
using namespace sd::tc;
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__global__ void check(int mn_major, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  float* A = reinterpret_cast<float*>(g);             // 4 KB
  float* B = reinterpret_cast<float*>(g + 8192);
  const uint32_t bar = base + 16384, slot = bar + 8;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(g + 16384 + 8);
  const int M = 128, N = 16, K = 8;
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    const float v = (float)(m % 7) + (float)k;
    if (mn_major) A[(m / 4) * (K * 4) + k * 4 + (m % 4)] = v;          // chunk m/4: [k][4 m]
    else A[(k / 4) * (M * 4) + m * 4 + (k % 4)] = v;                   // chunk k/4: [m][4 k]
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const float v = (k == n % 8) ? (float)(n + 1) : 0.f;
    if (mn_major) B[(n / 4) * (K * 4) + k * 4 + (n % 4)] = v;
    else B[(k / 4) * (N * 4) + n * 4 + (k % 4)] = v;
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint64_t dA, dB;
    if (mn_major) { idesc |= (1u << 15) | (1u << 16); dA = desc_nosw(base, 128, K * 16); dB = desc_nosw(base + 8192, 128, K * 16); }
    else { dA = desc_nosw(base, M * 16, 128); dB = desc_nosw(base + 8192, N * 16, 128); }
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(dA), "l"(dB), "r"(idesc), "r"(0u) : "memory");
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  if (threadIdx.x < 128) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(threadIdx.x & ~31) << 16), v);
    for (int n = 0; n < N; ++n) out[threadIdx.x * N + n] = v[n];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32));
}
int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  cudaFuncSetAttribute(check, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
  for (int mn = 0; mn < 2; ++mn) {
    cudaMemset(d, 0, 128 * 16 * 4);
    check<<<1, 128, 20 * 1024>>>(mn, d);
    float h[128 * 16];
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", mn ? "MN-major" : "K-major", cudaGetErrorString(e)); cudaDeviceReset(); continue; }
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) { const float want = ((m % 7) + n % 8) * (float)(n + 1); if (h[m * 16 + n] != want) ++bad; }
    printf("kind::tf32 %s operands: %d / 2048 entries wrong; D[5][3] = %g (want %g), D[100][9] = %g (want %g)\n", mn ? "MN-major" : "K-major ", bad,
           h[5 * 16 + 3], (5 % 7 + 3) * 4.f, h[100 * 16 + 9], (100 % 7 + 1) * 10.f);
  }
  return 0;
}
