// Development probe: issue cost of tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, SS mode) as a function of the N extent,
// the operand layout and the operand re-use pattern, measured per SM with clock64 (no TMA, operands = whatever is in smem).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ace_b200/csrc tools/mma_rate.cu -o gpurun_out/mma_rate
//
// Prints one JSON line per case: cycles per MMA (median over CTAs), the tcgen05 floor N/2, and the smem bytes one MMA reads.
// Used to choose the dhconv operand orientation and the tile widths of the small-N SHT stages.
#include <algorithm>
#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace ace;

struct Case {
  int n;          // MMA N
  int mn_major;   // operands MN-major (128B swizzle atoms) instead of K-major
  int bk;         // K extent of a stage: 32 (64B swizzle) or 64 (128B swizzle); ignored for MN-major
  int pattern;    // 0: one A, one B descriptor, rotating over 3 stages; 1: complex pattern (4 A planes, 4 B planes, 2 accumulators)
                  // 2: three-term pattern (a_hi b_hi, a_hi b_lo, a_lo b_hi)
  int pair;       // cta_group::2
};

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Case c, int reps, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t crank = PAIR ? ptx::cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_barrier_init();
  }
  // zero the operand area (values are irrelevant for timing; avoid NaN power artefacts)
  for (uint32_t i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (sbase - raw))[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (warp == 0) {
    if constexpr (PAIR) {
      ptx::tmem_alloc_2sm(ptx::smem_u32(&tmem_slot), 512);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512);
      ptx::tmem_relinquish();
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync();
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  long long cycles = 0;
  if (warp == 1 && crank == 0) {
    const int BK = c.bk;
    const uint32_t klayout = (BK == 64) ? 2u : 4u;
    const uint64_t d0 = c.mn_major ? ptx::smem_desc(0, BK * 128, 1024, 2u) : ptx::smem_desc(0, 0, 8 * BK * 2, klayout);
    const uint32_t kstep = (c.mn_major ? 2048 : 32) >> 4;
    const uint32_t idesc = ptx::instr_desc_bf16(PAIR ? 256 : 128, c.n, c.mn_major, c.mn_major);
    const uint32_t a_plane = 128 * BK * 2, b_plane = (PAIR ? 128 : 256) * BK * 2;  // worst-case plane sizes
    const uint32_t nplanes = c.pattern == 1 ? 4 : 2;
    const uint32_t stage_bytes = nplanes * (a_plane + b_plane);
    const int nstages = (int)min(3u, (200u * 1024u) / stage_bytes);
    const uint32_t bar_a = ptx::smem_u32(&bar);
    __syncwarp();
    const long long t0 = clock64();
    if (ptx::elect_one()) {
      int stage = 0;
      for (int r = 0; r < reps; ++r) {
        const uint32_t sA = sbase + stage * stage_bytes;
        const uint64_t a0 = d0 + (uint64_t)(sA >> 4), b0 = d0 + (uint64_t)((sA + nplanes * a_plane) >> 4);
        const uint64_t AP = a_plane >> 4, BP = b_plane >> 4;
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t a = a0 + kk * kstep, b = b0 + kk * kstep;
          if (c.pattern == 0) {
            if constexpr (PAIR) ptx::umma_bf16_2sm(tmem_base, a, b, idesc, 1u);
            else ptx::umma_bf16(tmem_base, a, b, idesc, 1u);
          } else if (c.pattern == 2) {
            if constexpr (PAIR) {
              ptx::umma_bf16_2sm(tmem_base, a, b, idesc, 1u);
              ptx::umma_bf16_2sm(tmem_base, a, b + BP, idesc, 1u);
              ptx::umma_bf16_2sm(tmem_base, a + AP, b, idesc, 1u);
            } else {
              ptx::umma_bf16(tmem_base, a, b, idesc, 1u);
              ptx::umma_bf16(tmem_base, a, b + BP, idesc, 1u);
              ptx::umma_bf16(tmem_base, a + AP, b, idesc, 1u);
            }
          } else {
            const uint32_t tr = tmem_base, ti = tmem_base + 256;
            const uint32_t ineg = idesc | (1u << 13);
            if constexpr (!PAIR) {
              ptx::umma_bf16(tr, a, b, idesc, 1u);
              ptx::umma_bf16(tr, a, b + BP, idesc, 1u);
              ptx::umma_bf16(tr, a + AP, b, idesc, 1u);
              ptx::umma_bf16(tr, a + 2 * AP, b + 2 * BP, ineg, 1u);
              ptx::umma_bf16(tr, a + 2 * AP, b + 3 * BP, ineg, 1u);
              ptx::umma_bf16(tr, a + 3 * AP, b + 2 * BP, ineg, 1u);
              ptx::umma_bf16(ti, a + 2 * AP, b, idesc, 1u);
              ptx::umma_bf16(ti, a + 2 * AP, b + BP, idesc, 1u);
              ptx::umma_bf16(ti, a + 3 * AP, b, idesc, 1u);
              ptx::umma_bf16(ti, a, b + 2 * BP, idesc, 1u);
              ptx::umma_bf16(ti, a, b + 3 * BP, idesc, 1u);
              ptx::umma_bf16(ti, a + AP, b + 2 * BP, idesc, 1u);
            }
          }
        }
        if (++stage == nstages) stage = 0;
      }
      if constexpr (PAIR) ptx::umma_commit_2sm(bar_a);
      else ptx::umma_commit(bar_a);
    }
    __syncwarp();
    ptx::mbar_wait(bar_a, 0);
    cycles = clock64() - t0;
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x] = cycles;
  } else if (warp == 1) {
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);  // peer: the multicast commit arrives here too
  }
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync();
  else __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_2sm(tmem_base, 512);
    else ptx::tmem_dealloc(tmem_base, 512);
  }
}

static void run_case(const Case& c, int sms, long long* d_out) {
  const int reps = 2000;
  const size_t smem = 220 * 1024;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(c.pair ? (sms / 2) * 2 : sms);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = c.pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  for (int it = 0; it < 2; ++it) {  // second run is the measurement
    if (c.pair) e = cudaLaunchKernelEx(&cfg, mma_rate_kernel<true>, c, reps, d_out);
    else e = cudaLaunchKernelEx(&cfg, mma_rate_kernel<false>, c, reps, d_out);
    if (e != cudaSuccess) {
      printf("{\"error\": \"launch: %s\"}\n", cudaGetErrorString(e));
      return;
    }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("{\"error\": \"sync: %s\"}\n", cudaGetErrorString(e));
      return;
    }
  }
  std::vector<long long> h(cfg.gridDim.x);
  cudaMemcpy(h.data(), d_out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  std::vector<long long> v;
  for (size_t i = 0; i < h.size(); i += (c.pair ? 2 : 1)) v.push_back(h[i]);
  std::sort(v.begin(), v.end());
  const int per_rep = (c.bk / 16) * (c.pattern == 0 ? 1 : c.pattern == 2 ? 3 : 12);
  const double n_mma = (double)reps * per_rep;
  const int brows = c.pair ? c.n / 2 : c.n;
  printf("{\"n\": %d, \"mn_major\": %d, \"bk\": %d, \"pattern\": %d, \"pair\": %d, \"cycles_per_mma_med\": %.1f, \"min\": %.1f, \"max\": %.1f, "
         "\"floor_n_over_2\": %.1f, \"smem_bytes_per_mma_per_cta\": %d}\n",
         c.n, c.mn_major, c.bk, c.pattern, c.pair, v[v.size() / 2] / n_mma, v.front() / n_mma, v.back() / n_mma, c.n / 2.0,
         128 * 32 + brows * 32);
  fflush(stdout);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  long long* d_out;
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  const int ns[] = {16, 32, 48, 64, 96, 128, 160, 192, 256};
  for (int n : ns) run_case({n, 0, 32, 0, 0}, sms, d_out);
  for (int n : ns) run_case({n, 0, 64, 0, 0}, sms, d_out);
  for (int n : ns) run_case({n, 0, 32, 2, 0}, sms, d_out);
  for (int n : {16, 32, 64, 96, 128}) run_case({n, 0, 32, 1, 0}, sms, d_out);
  for (int n : {64, 128, 192, 256}) run_case({n, 1, 32, 2, 0}, sms, d_out);
  for (int n : {32, 64, 96, 128, 192, 256}) run_case({n, 0, 32, 2, 1}, sms, d_out);
  for (int n : {32, 64, 96, 128, 192, 256}) run_case({n, 0, 64, 2, 1}, sms, d_out);
  cudaFree(d_out);
  return 0;
}
