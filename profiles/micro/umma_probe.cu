// umma_probe.cu -- how long does one tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) take as a function of N and of the
// shared-memory operand layout?  One CTA per SM-sized grid of 1; thread 0 of warp 1 issues `reps` accumulating MMAs
// back to back on the same operands, commits, and waits; cycles / reps is printed.
//   layouts: 0 = K-major SWIZZLE_128B (rows of 128 B, 8-row atoms, SBO 1024)
//            1 = K-major no swizzle, chunk-major [K/8][rows][16 B] (LBO = rows*16, SBO = 128)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I safe_dreamer_b200/csrc -o profiles/micro/umma_probe profiles/micro/umma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sd_tc.cuh"
using namespace sd::tc;

__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(256, 1) probe(int N, int layout, int reps, int ksteps, int nacc, long long* out, int lboB = 0, int noise = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < 196 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0u;
  const uint32_t bar = base + 196 * 1024, slot = bar + 8;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(g + 196 * 1024 + 8);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  volatile int* stop = reinterpret_cast<volatile int*>(g + 196 * 1024 + 64);
  if (threadIdx.x == 0) *stop = 0;
  __syncthreads();
  if (noise && threadIdx.x >= 128) {
    uint4* dst = reinterpret_cast<uint4*>(g + 190 * 1024) + (threadIdx.x - 128);
    int n = 0;
    while (!*stop && n < 2000000) { for (int j = 0; j < 2; ++j) dst[j * 128] = make_uint4(n, n, n, n); ++n; }
  }
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc(128, N) | (layout == 2 ? (1u << 15) | (1u << 16) : 0u);
    const uint32_t a = base, b = base + 32 * 1024;   // A: 128 x 64 bf16 (16 KB), B: up to 256 x 64 (32 KB)
    uint64_t dA, dB; uint32_t ka, kb;
    if (layout == 0) { dA = make_desc_sw128(a); dB = make_desc_sw128(b); ka = kb = 32 >> 4; }
    else if (layout == 2) { dA = desc_nosw(a, 128, 2048); dB = desc_nosw(b, 128, 2048); ka = kb = 256 >> 4; }   // MN-major: [chunk][128 px][16 B], K = pixels
    else { const int lb = lboB ? lboB : N * 16; dA = desc_nosw(a, 128 * 16, 128); dB = desc_nosw(b, lb, 128); ka = (2 * 128 * 16) >> 4; kb = (2 * lb) >> 4; }
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      // four k-steps per trip, descriptors precomputed, accumulator chosen without a division (nacc is a power of two)
      const uint64_t a0 = dA, a1 = dA + ka, a2 = dA + 2 * ka, a3 = dA + 3 * ka, b0 = dB, b1 = dB + kb, b2 = dB + 2 * kb, b3 = dB + 3 * kb;
#pragma unroll 1
      for (int i = 0; i < reps; i += 4) {
        const uint32_t t = tmem + (uint32_t)(((i >> 2) & (nacc - 1)) * N);
        tc_mma_f16(t, a0, b0, idesc, 1u);
        tc_mma_f16(t, a1, b1, idesc, 1u);
        tc_mma_f16(t, a2, b2, idesc, 1u);
        tc_mma_f16(t, a3, b3, idesc, 1u);
      }
      const long long t1 = clock64();
      tc_commit(bar);
      mbar_wait(bar, phase); phase ^= 1;
      const long long t2 = clock64();
      out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0;
    }
    *stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512)); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 256;
  for (int layout = 0; layout < 2; ++layout)
    for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
      probe<<<1, 256, 200 * 1024>>>(N, layout, reps, 4, 1, d);
      long long h[6]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("layout %s N=%3d: issue %.1f cycles/MMA, complete %.1f cycles/MMA  (%.0f flop/clk)\n", layout ? "no-swizzle" : "SW128     ", N,
             (double)h[4] / reps, (double)h[5] / reps, 2.0 * 128 * N * 16 / ((double)h[5] / reps));
    }
  // independent accumulators (round robin over `nacc` TMEM column ranges): is the ~160-cycle floor a dependency latency?
  for (int nacc : {1, 4})
    for (int N : {32, 64}) {
      probe<<<1, 256, 200 * 1024>>>(N, 1, reps, 4, nacc, d);
      long long h[6]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
      printf("no-swizzle N=%3d, %d accumulators round robin: issue %.1f, complete %.1f cycles/MMA (%.0f flop/clk)\n", N, nacc,
             (double)h[4] / reps, (double)h[5] / reps, 2.0 * 128 * N * 16 / ((double)h[5] / reps));
    }
  for (int N : {32, 48, 64, 128}) {
    probe<<<1, 256, 200 * 1024>>>(N, 2, reps, 4, 1, d);
    long long h[6]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    printf("MN-major (both operands) N=%3d: issue %.1f, complete %.1f cycles/MMA\n", N, (double)h[4] / reps, (double)h[5] / reps);
  }
  struct V { int N, lbo, nacc, noise; const char* what; };
  const V vs[] = {{96, 0, 1, 0, "N=96 dense B"}, {96, 19200, 1, 0, "N=96 B rows 19200 B apart per K chunk"}, {96, 19200, 2, 0, "same, two accumulators"},
                  {96, 19200, 2, 1, "same + 4 warps storing to shared memory"}, {256, 0, 1, 1, "N=256 + store noise"}, {64, 0, 1, 1, "N=64 + store noise"}};
  for (const V& v : vs) {
    probe<<<1, 256, 200 * 1024>>>(v.N, 1, reps, 4, v.nacc, d, v.lbo, v.noise);
    long long h[6]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    printf("%-45s: %.1f cycles/MMA\n", v.what, (double)h[5] / reps);
  }
  return 0;
}
