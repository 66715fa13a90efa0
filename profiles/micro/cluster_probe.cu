// cluster_probe.cu -- measurements that size the persistent imagination kernel (csrc/sd_pimg.cuh):
//   * can 8 clusters of 16 CTAs (1 CTA/SM, ~200 KB smem) be co-resident on this B200?
//   * cost of one cluster barrier vs an all-to-all remote mbarrier arrive
//   * L2 -> SM ingest per CTA with cp.async.bulk: 128 CTAs distinct data / the 16 CTAs of a cluster reading the SAME data /
//     one multicast load delivered to 16 CTAs
//   * DSMEM push bandwidth: st.shared::cluster.v4 and cp.async.bulk smem -> remote smem
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_probe cluster_probe.cu ; prints one line per test.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 24); ++i) if (mbar_try(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}

// ---- 1. cluster barrier latency
__global__ void k_cluster_barrier(long long* out, int iters) {
  cluster_sync_all();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) cluster_sync_all();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
}

// ---- 2. all-to-all remote mbarrier arrive (one thread per CTA arrives on all peers, one thread waits)
__global__ void k_mbar_a2a(long long* out, int iters, int nrank) {
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), nrank); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  cluster_sync_all();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int i = 0; i < iters; ++i) {
      for (int r = 0; r < nrank; ++r) remote_arrive(mapa(smem_u32(&bar), (uint32_t)r));
      mbar_wait(smem_u32(&bar), (uint32_t)(i & 1));
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
  cluster_sync_all();
}

// ---- 3. L2 -> SM ingest with cp.async.bulk (1D), ring of `depth` chunks
// mode 0: every CTA walks its own slice of the buffer; 1: all CTAs of a cluster read the same addresses;
// 2: rank 0 issues ONE multicast load per chunk delivered to all CTAs of the cluster
constexpr int CHUNK = 16384;
__global__ void k_bulk_ingest(const uint8_t* __restrict__ src, size_t src_bytes, int nchunks, int depth, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t full[8];
  const uint32_t rank = cta_rank();
  uint32_t csize; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const int cluster_id = blockIdx.x / csize;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    const size_t per = (size_t)nchunks * CHUNK;
    size_t base = (mode == 0 ? (size_t)blockIdx.x : (size_t)cluster_id) * per % (src_bytes - per);
    base &= ~(size_t)1023;
    for (int i = 0; i < nchunks + depth; ++i) {
      if (i >= depth) mbar_wait(smem_u32(&full[(i - depth) % depth]), (uint32_t)(((i - depth) / depth) & 1));
      if (mode == 2 && i >= depth) {
        // consumer-release handshake for the multicast ring: every CTA tells rank 0 its slot is free (cluster barrier by t0 only)
      }
      if (i < nchunks) {
        const int s = i % depth;
        mbar_expect_tx(smem_u32(&full[s]), CHUNK);
        if (mode == 2) {
          if (rank == 0) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(sm + s * CHUNK)), "l"(src + base + (size_t)i * CHUNK), "r"(CHUNK), "r"(smem_u32(&full[s])), "h"((uint16_t)((1u << csize) - 1)) : "memory");
          }
        } else {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(sm + s * CHUNK)), "l"(src + base + (size_t)i * CHUNK), "r"(CHUNK), "r"(smem_u32(&full[s])) : "memory");
        }
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  cluster_sync_all();
}

// ---- 4. DSMEM push with st.shared::cluster.v4 (every thread pushes 16 B per iteration to the next rank)
__global__ void k_dsmem_st(long long* out, int bytes_per_cta, int iters) {
  extern __shared__ __align__(16) uint8_t sm[];
  const uint32_t rank = cta_rank();
  uint32_t csize; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  cluster_sync_all();
  long long t0 = clock64();
  const uint32_t peer_base = mapa(smem_u32(sm), (rank + 1) % csize);
  for (int it = 0; it < iters; ++it) {
    for (int off = threadIdx.x * 16; off < bytes_per_cta; off += blockDim.x * 16)
      asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(peer_base + off), "f"(1.f), "f"(2.f), "f"(3.f), "f"((float)it) : "memory");
  }
  cluster_sync_all();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// ---- 5. DSMEM push with cp.async.bulk smem -> remote smem (16 KB chunks, remote mbarrier complete_tx)
__global__ void k_dsmem_bulk(long long* out, int nchunks) {
  extern __shared__ __align__(1024) uint8_t sm[];   // [0, 64K): source, [64K, 128K): destination written by the previous rank
  __shared__ __align__(8) uint64_t bar;
  const uint32_t rank = cta_rank();
  uint32_t csize; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  cluster_sync_all();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar), (uint32_t)nchunks * CHUNK);
    const uint32_t peer = (rank + 1) % csize;
    for (int i = 0; i < nchunks; ++i) {
      const uint32_t dst = mapa(smem_u32(sm + 65536 + (i % 4) * CHUNK), peer);
      const uint32_t rbar = mapa(smem_u32(&bar), peer);
      asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "r"(smem_u32(sm + (i % 4) * CHUNK)), "r"(CHUNK), "r"(rbar) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  cluster_sync_all();
}


// ---- 6. 2-D tiled TMA ingest (the GEMM operand path): boxes of 128 rows x 64 bf16 (16 KB, SWIZZLE_128B) from a
// [rows][2048] bf16 matrix; `streams` producer threads (in different warps) each run their own ring of `depth` boxes.
// same=1: the 16 CTAs of a team read the same boxes.
#include <cuda.h>
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
struct Maps4 { CUtensorMap m[4]; };
__global__ void k_tiled_ingest(const __grid_constant__ Maps4 maps, int nrows, int nbox, int depth, int streams, int same, int per_bar,
                               int box_bytes, int distinct_maps, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t full[4][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&full[s][i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp < streams && lane == 0) {
    const CUtensorMap* map = &maps.m[distinct_maps ? warp : 0];
    const int who = same ? (int)(blockIdx.x / 16) : (int)blockIdx.x;
    const int row_tiles = nrows / 128;
    const int ngroups = nbox / per_bar;
    long long c_wait = 0, c_exp = 0, c_tma = 0;
    for (int i = 0; i < ngroups + depth; ++i) {
      long long a0 = clock64();
      if (i >= depth) mbar_wait(smem_u32(&full[warp][(i - depth) % depth]), (uint32_t)(((i - depth) / depth) & 1));
      long long a1 = clock64();
      c_wait += a1 - a0;
      if (i < ngroups) {
        const int s = i % depth;
        mbar_expect_tx(smem_u32(&full[warp][s]), (uint32_t)(box_bytes * per_bar));
        long long a2 = clock64();
        c_exp += a2 - a1;
        for (int j = 0; j < per_bar; ++j) {
          const int lin = who * 977 + warp * 331 + i * per_bar + j;
          const int rt = (lin / 32) % row_tiles, kb = lin % 32;
          tma2d(smem_u32(sm + ((warp * depth + s) * per_bar + j) * 16384), map, kb * 64, rt * 128, smem_u32(&full[warp][s]));
        }
        c_tma += clock64() - a2;
      }
    }
    if (blockIdx.x == 0 && warp == 0) { out[200] = c_wait / ngroups; out[201] = c_exp / ngroups; out[202] = c_tma / ngroups; }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// ---- 7. generic loads: every thread streams uint4 from L2 into shared memory (ILP 8)
__global__ void k_ldg_ingest(const uint4* __restrict__ src, size_t n16, int iters, long long* out) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* s4 = reinterpret_cast<uint4*>(sm);
  long long t0 = clock64();
  size_t base = ((size_t)blockIdx.x * 131071) % (n16 - (size_t)iters * 8 * blockDim.x - 1);
  for (int it = 0; it < iters; ++it) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(src + base + ((size_t)it * 8 + j) * blockDim.x + threadIdx.x));
#pragma unroll
    for (int j = 0; j < 8; ++j) s4[(j * blockDim.x + threadIdx.x) % 4096] = v[j];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// ---- 8. team sync through L2: 16 CTAs per team, red.release.gpu + relaxed poll + acquire fence, one line per team
__global__ void k_flag_sync(unsigned int* flags, int iters, long long* out) {
  unsigned int* f = flags + (blockIdx.x / 16) * 32;
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int i = 1; i <= iters; ++i) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
      unsigned int v = 0;
      for (int spin = 0; spin < (1 << 22); ++spin) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v >= (unsigned)i * 16u) break;
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}

template <class K, class... Args>
static cudaError_t launch_cluster(K kernel, int grid, int block, size_t smem, int cluster, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s SMs=%d smem/block optin=%zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  long long* out;
  CK(cudaMalloc(&out, 1024 * sizeof(long long)));
  long long h[1024];
  const size_t SRC = 64ull << 20;
  uint8_t* src;
  CK(cudaMalloc(&src, SRC));
  CK(cudaMemset(src, 1, SRC));

  // occupancy of 16-CTA clusters with a large smem footprint
  for (int cs : {8, 16}) {
    for (size_t smem : {(size_t)100 * 1024, (size_t)200 * 1024, (size_t)220 * 1024}) {
      CK(cudaFuncSetAttribute(k_bulk_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaFuncSetAttribute(k_bulk_ingest, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(cs * 8); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int ncl = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, k_bulk_ingest, &cfg);
      printf("occupancy cluster=%d smem=%zuKB threads=384: max active clusters=%d (%s)\n", cs, smem / 1024, ncl, cudaGetErrorString(e));
    }
  }
  CK(cudaFuncSetAttribute(k_cluster_barrier, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(k_mbar_a2a, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(k_dsmem_st, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(k_dsmem_bulk, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(k_dsmem_st, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  CK(cudaFuncSetAttribute(k_dsmem_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CK(cudaFuncSetAttribute(k_bulk_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));

  for (int cs : {2, 4, 8, 16}) {
    for (int threads : {128, 384}) {
      CK(launch_cluster(k_cluster_barrier, cs * 8, threads, 0, cs, out, 2000));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
      printf("cluster barrier: cluster=%d threads=%d grid=%d  %lld cycles\n", cs, threads, cs * 8, h[0]);
    }
    CK(launch_cluster(k_mbar_a2a, cs * 8, 128, 0, cs, out, 2000, cs));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
    printf("all-to-all remote mbarrier arrive + wait: cluster=%d  %lld cycles\n", cs, h[0]);
  }
  // warm the source into L2
  CK(launch_cluster(k_bulk_ingest, 128, 128, (size_t)8 * CHUNK, 16, (const uint8_t*)src, SRC, 256, 4, 0, out));
  CK(cudaDeviceSynchronize());
  for (int grid : {16, 128}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int depth : {2, 4, 8}) {
        if (mode == 2 && depth != 8) continue;
        const int nch = mode == 2 ? 8 : 128;   // multicast variant: one ring fill only (no release handshake): 8 chunks in flight
        for (int rep = 0; rep < 2; ++rep) {
          CK(launch_cluster(k_bulk_ingest, grid, 128, (size_t)8 * CHUNK, 16, (const uint8_t*)src, SRC, nch, depth, mode, out));
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost));
        long long mx = 0, sum = 0;
        for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
        printf("bulk ingest mode=%d (0 distinct,1 cluster-same,2 multicast) grid=%d depth=%d: %d x 16KB per CTA, max %lld cycles -> %.1f B/clk per CTA (mean %.1f)\n",
               mode, grid, depth, nch, mx, (double)nch * CHUNK / mx, (double)nch * CHUNK * grid / sum);
      }
    }
  }
  for (int cs : {2, 16}) {
    for (int threads : {128, 256}) {
      CK(launch_cluster(k_dsmem_st, cs * 8, threads, 65536, cs, out, 65536, 16));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
      printf("DSMEM st.v4 push: cluster=%d threads=%d 16 x 64KB per CTA: %lld cycles -> %.1f B/clk per CTA\n", cs, threads, h[0], 16.0 * 65536 / h[0]);
    }
    CK(launch_cluster(k_dsmem_bulk, cs * 8, 128, 140 * 1024, cs, out, 32));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
    printf("DSMEM cp.async.bulk push: cluster=%d 32 x 16KB per CTA: %lld cycles -> %.1f B/clk per CTA\n", cs, h[0], 32.0 * CHUNK / h[0]);
  }

  // ---- 2-D tiled TMA
  {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fp;
    const int nrows = 8192;                       // 8192 x 2048 bf16 = 32 MB: L2 resident
    Maps4 maps;
    cuuint64_t dims[2] = {2048, (cuuint64_t)nrows}; cuuint64_t strides[1] = {4096}; cuuint32_t es[2] = {1, 1};
    auto mk = [&](CUtensorMap* m, int box_rows, void* base) {
      cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
      return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CK(cudaFuncSetAttribute(k_tiled_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int box_rows : {128, 16}) {
      for (int i = 0; i < 4; ++i) { CUresult r = mk(&maps.m[i], box_rows, src); if (r) printf("encode failed %d\n", (int)r); }
      for (int grid : {128}) for (int streams : {1, 2, 4}) for (int per_bar : {1, 2}) for (int dm : {0}) {
        const int depth = 3;
        if (box_rows != 128 && (per_bar == 2 || dm == 1 || grid == 16)) continue;
        if ((size_t)streams * depth * per_bar * 16384 + 1024 > 200 * 1024) continue;
        const int nbox = 240;
        for (int rep = 0; rep < 2; ++rep) {
          k_tiled_ingest<<<grid, 128, (size_t)streams * depth * per_bar * 16384 + 1024>>>(maps, nrows, nbox, depth, streams, 1, per_bar, box_rows * 128, dm, out);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, out, 203 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("[per iteration: wait %lld expect_tx %lld tma issue %lld cycles] ", h[200], h[201], h[202]);
        printf("tiled TMA ingest box=%dx64 grid=%d streams=%d boxes/barrier=%d distinct_maps=%d depth=%d: %d boxes per stream, max %lld cycles -> %.1f B/clk per CTA, %.0f cycles per box per stream\n",
               box_rows, grid, streams, per_bar, dm, depth, nbox, mx, (double)streams * nbox * box_rows * 128 / mx, (double)mx / nbox);
      }
    }
  }
  // ---- generic loads
  CK(cudaFuncSetAttribute(k_ldg_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int grid : {16, 128}) for (int threads : {256, 512}) {
    const int iters = 64;
    for (int rep = 0; rep < 2; ++rep) { k_ldg_ingest<<<grid, threads, 65536>>>((const uint4*)src, SRC / 16, iters, out); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("ld.global.cg.v4 ingest grid=%d threads=%d: %.0f KB per CTA, max %lld cycles -> %.1f B/clk per CTA\n", grid, threads,
           iters * 8.0 * threads * 16 / 1024, mx, iters * 8.0 * threads * 16 / mx);
  }
  // ---- team sync through L2
  {
    unsigned int* flags; CK(cudaMalloc(&flags, 4096));
    for (int grid : {16, 128}) {
      CK(cudaMemset(flags, 0, 4096));
      k_flag_sync<<<grid, 128>>>(flags, 1000, out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("team sync through L2 (16 CTAs per team, red.release + poll + fence): grid=%d  %lld cycles per sync\n", grid, mx);
    }
  }
  printf("done\n");
  return 0;
}
