// Development probe: shared-memory fill bandwidth of TMA tile loads when several CTAs need the SAME tile.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ace_b200/csrc tools/tma_bw.cu -o gpurun_out/tma_bw
//
// Every CTA (one per SM, no compute) streams 32 KB stages through a 6-deep ring.  Modes:
//   distinct   every CTA loads its own tiles (L2-resident source): the L2 -> SM fabric limit
//   shared     groups of G neighbouring CTAs load the same tile, each with its own unicast TMA (what the three row tiles of a
//              384-row convolution do today with the activation tile)
//   multicast  clusters of G CTAs: each loads 1/G of the tile and multicasts it into all G CTAs; a stage is refilled once all G
//              consumers have released it (remote mbarrier arrives from a relay thread, not from the consumer's critical path)
// Prints one JSON line per case: bytes landed in shared memory per second (all CTAs), and the same per SM clock.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace ace;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));                          \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

constexpr int kStages = 6, kRows = 256, kStageBytes = kRows * 128;  // box = [kRows][64 bf16]

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_2d_mc(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// mode 0 distinct, 1 shared (unicast), 2 multicast
__global__ void __launch_bounds__(128, 1) tma_bw_kernel(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_slice,
                                                        int mode, int G, int iters, int ntiles, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bars[2 * kStages];
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = (mode == 2) ? ptx::cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), mode == 2 ? G : 1);
    }
    ptx::fence_barrier_init();
  }
  if (mode == 2) ptx::cluster_sync();
  else __syncthreads();
  const int group = blockIdx.x / G, ngroups = gridDim.x / G;
  const long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t dst = sbase + stage * kStageBytes;
      ptx::mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
      if (mode == 0) {
        const int tile = (int)(((long long)i * gridDim.x + blockIdx.x) % ntiles);
        tma_2d(dst, &tm_full, full_bar(stage), 0, tile * kRows);
      } else if (mode == 1) {
        const int tile = (int)(((long long)i * ngroups + group) % ntiles);
        tma_2d(dst, &tm_full, full_bar(stage), 0, tile * kRows);
      } else {
        const int tile = (int)(((long long)i * ngroups + group) % ntiles);
        const int rows = kRows / G;  // G divides kRows (host-checked)
        tma_2d_mc(dst + crank * rows * 128, &tm_slice, full_bar(stage), 0, tile * kRows + (int)crank * rows, (uint16_t)((1u << G) - 1u));
      }
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1 && lane == 0) {
    // consumer: the data has landed -> release the stage (locally, or through the relay in multicast mode)
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      ptx::mbar_wait(full_bar(stage), phase);
      if (mode == 2) {
        for (int r = 0; r < G; ++r) remote_arrive(mapa(empty_bar(stage), (uint32_t)r));
      } else {
        ptx::mbar_arrive(empty_bar(stage));
      }
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (mode == 2) ptx::cluster_sync();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)f;
  const int ntiles = 2048;  // 64 MB source: L2-resident after the warm-up pass
  const size_t bytes = (size_t)ntiles * kStageBytes;
  void* src = nullptr;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMemset(src, 0, bytes));
  long long* cycles = nullptr;
  CK(cudaMalloc(&cycles, 256 * sizeof(long long)));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  auto make_map = [&](int box_rows) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {64, (cuuint64_t)ntiles * kRows}, strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("encode failed %d\n", (int)r);
      exit(1);
    }
    return tm;
  };
  const int smem = kStages * kStageBytes + 1024;
  CK(cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int iters = 600;
  struct Case { int mode, G; };
  const Case cases[] = {{0, 1}, {1, 2}, {1, 3}, {1, 4}, {2, 2}, {2, 4}, {2, 8}, {1, 8}};
  for (const Case& c : cases) {
    if (c.mode == 2 && kRows % c.G) continue;
    const int grid = sms / c.G * c.G;
    CUtensorMap tm_full = make_map(kRows), tm_slice = make_map(kRows / c.G);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c.mode == 2 ? c.G : 1;
    attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0));
      CK(cudaLaunchKernelEx(&cfg, tma_bw_kernel, tm_full, tm_slice, c.mode, c.G, iters, ntiles, cycles));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long cmax = 0;
    for (long long v : h) cmax = v > cmax ? v : cmax;
    const double landed = (double)grid * iters * kStageBytes;
    printf("{\"mode\": \"%s\", \"G\": %d, \"grid\": %d, \"ms\": %.4f, \"smem_fill_TBps\": %.2f, \"smem_fill_B_per_clk_per_sm\": %.1f, \"l2_read_TBps_if_no_dedup\": %.2f}\n",
           c.mode == 0 ? "distinct" : c.mode == 1 ? "shared-unicast" : "multicast", c.G, grid, best, landed / best / 1e9, (double)iters * kStageBytes / (double)cmax,
           (c.mode == 2 ? landed / c.G : landed) / best / 1e9);
  }
  return 0;
}
