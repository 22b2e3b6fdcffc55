// Shared host/device helpers for the ace_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/ace_b200.h"

namespace ace {

typedef __nv_bfloat16 bf16;

// ---- errors: internal code throws, the extern "C" layer converts to code + message ----
struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline std::string strprintf(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return std::string(buf);
}

#define ACE_CHECK_CUDA(expr)                                                                       \
  do {                                                                                             \
    cudaError_t err__ = (expr);                                                                    \
    if (err__ != cudaSuccess)                                                                      \
      throw ::ace::Error(ACE_ERR_CUDA, ::ace::strprintf("%s failed: %s (%s:%d)", #expr,           \
                                                        cudaGetErrorString(err__), __FILE__, __LINE__)); \
  } while (0)

#define ACE_REQUIRE(cond, ...)                                                           \
  do {                                                                                   \
    if (!(cond)) throw ::ace::Error(ACE_ERR_INVALID, ::ace::strprintf(__VA_ARGS__));    \
  } while (0)

// extern "C" wrappers: convert exceptions to (code, thread-local message)
void set_last_error(const char* msg);
#define ACE_API_BEGIN try {
#define ACE_API_END                         \
  }                                         \
  catch (const ::ace::Error& e) {           \
    ::ace::set_last_error(e.what());        \
    return e.code;                          \
  }                                         \
  catch (const std::exception& e) {         \
    ::ace::set_last_error(e.what());        \
    return ACE_ERR_INVALID;                 \
  }                                         \
  return ACE_OK;

// launch bookkeeping (ace_launch_count) + launch error check
extern std::atomic<long long> g_launch_count;
extern std::atomic<long long> g_umma_count, g_simt_count;  // GEMMs routed to each kernel
inline void after_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error(ACE_ERR_CUDA, strprintf("launch of %s failed: %s", what, cudaGetErrorString(e)));
  }
}

// Optional per-kernel timing (option "profile" = 1): CUDA events around every launch on its own stream,
// aggregated by name in ace_profile_report().  Off by default; cannot be used under graph capture.
struct ProfileScope {
  const char* name;
  cudaStream_t stream;
  cudaEvent_t start = nullptr, stop = nullptr;
  bool ranged = false;  // an NVTX range is open (option "nvtx")
  bool hooked = false;  // the host's scope callback was told about the begin (ace_set_scope_callback)
  ProfileScope(const char* name, cudaStream_t stream);
  ~ProfileScope();
};

// runtime options
struct Options {
  int profile = 0;
  int nvtx = 0;      // NVTX range per operator launch, named like the profile report (nsys / `ncu --nvtx --nvtx-include "dhconv/"`); env ACE_B200_NVTX
  int force_simt = 0;
  int split_terms = 3;
  int pair = -1;     // CTA-pair (cta_group::2, 256 x 256 tiles) variants: -1 = per-op choice (convolutions with M % 256 == 0), 0 = never, 1 = wherever compiled
  int conv_bn = 0;   // N tile of the 1x1-convolution GEMMs: 0 = per-op choice (wave quantisation), 192, 256
  int pdl = 0;       // programmatic dependent launch of the tcgen05 GEMM / prep kernels (prologue overlaps the previous kernel's tail)
  int dbg = 0;       // development switches of the tcgen05 kernel (results are wrong when non-zero)
  int umma_bk = 0;   // K extent per pipeline stage of the ROWC (forward SHT / dhconv) variants: 0 = per-op hint, 32, 64
  int tile_list = 1; // triangular GEMMs walk a host-built list of their non-empty tiles, heaviest first (0: implicit round-robin walk of the full tile box)
  int l2_persist = 0;  // experiment: L2 persisting access-policy window over the split-plane output of every tcgen05 GEMM
  int inv2 = 1;      // networks use the padded / parity-split inverse Legendre stage + butterfly inverse DFT (sht.cuh); 0: first-generation pair
  int dhconv_t = 0;  // dhconv orientation: 0 = weights on the rows, the l + 1 orders on the columns (fewest multiplications: measured 83 vs 89 us);
                     // 1 = orders on the rows / output channels on the columns (128 x 128 MMAs, NC epilogue)
  int group_order = 1;  // dense ops with several N tiles: a worker takes all N tiles of its M tile back to back (0: implicit round-robin walk)
  int mma_batch = 1;  // tcgen05 kernel: MMAs of two ring slots per barrier round when both have landed (0: one slot per round)
  int sp = 1;        // 1x1 convolutions with >= 128 output channels run with the spatial positions on the accumulator rows (gemm_umma.cu, Cfg::SP)
  int sp_tma = 1;    // SP variants store their split-plane output by TMA from a shared-memory tile (0: element-wise stores from registers)
  int bfly_pair = 1; // butterfly inverse DFT on CTA pairs when the row count is a multiple of 256
  int tile_serpentine = 1;  // tile lists of the triangular GEMMs: odd strata reversed so every worker's tile costs sum to about the same
  int sp_tmx = 1;    // SP variants fetch the epilogue's addend / residual tile by TMA (0: loads from the epilogue threads)
  int cln_gemm = 1;  // ConditionalLayerNorm of the noise-conditioned SFNO on the tcgen05 kernel (GemmOp::cln: statistics pass + one GEMM whose
                     // epilogue normalises and modulates) where eligible (>= 128 channels, pixel count % 4 == 0, context padded to a multiple
                     // of 32); 0: the streaming kernel up to 64 context channels.  A/B on the ERA5-baseline network: 10.54 vs 10.71 ms per
                     // forward (profiles/r02c_cln_ab.json); env ACE_B200_CLN_GEMM
  int trace = 0;     // development: per-tile clock samples of the tcgen05 kernel's roles appended to $ACE_B200_TRACE_FILE (tools/trace_report.py)
  int umma_bn = 0;   // 0 = choose per op; otherwise force the N tile of the K-major x K-major variants (192 / 256)
};
Options& options();

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

// ---- device buffer with RAII (allocations happen at create/first-use time, never per step) ----
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) {
    o.p = nullptr;
    o.bytes = 0;
  }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p;
      bytes = o.bytes;
      o.p = nullptr;
      o.bytes = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  // (re)allocate zero-filled storage of at least n bytes; returns true if a new allocation was made
  bool ensure(size_t n) {
    if (n <= bytes && p) return false;
    release();
    ACE_CHECK_CUDA(cudaMalloc(&p, n));
    ACE_CHECK_CUDA(cudaMemset(p, 0, n));
    ACE_CHECK_CUDA(cudaDeviceSynchronize());  // allocation time only: make the zero fill visible to every stream
    bytes = n;
    return true;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- split-bf16 representation: v ~= hi + lo, |v - (hi+lo)| <= 2^-17 |v| ----
__host__ __device__ inline void split_bf16(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__host__ __device__ inline float join_bf16(bf16 hi, bf16 lo) { return __bfloat162float(hi) + __bfloat162float(lo); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Exact-erf GELU for the tensor-core epilogues: gelu(x) = relu(x) - 0.5 |x| erfc(|x| / sqrt 2) with
// erfc(t) = 2^(-t Q(t)), Q a degree-9 minimax fit on [0, 4.6] (beyond it erfc < 1e-10).  Max abs error vs the
// fp64 definition 2.4e-7 on [-8, 8] (= the fp32 rounding floor; nn.GELU's own erff path has the same size),
// one branch-free chain of 10 FMAs + one MUFU.EX2 instead of erff's two divergent branches.
__device__ __forceinline__ float gelu_fast(float x) {
  const float ax = fabsf(x);
  const float t = fminf(ax * 0.70710678118654752440f, 4.6f);
  float q = 4.565055586e-07f;
  q = fmaf(q, t, -1.097697806e-05f);
  q = fmaf(q, t, 1.118192688e-04f);
  q = fmaf(q, t, -6.078477993e-04f);
  q = fmaf(q, t, 1.635471654e-03f);
  q = fmaf(q, t, 8.210374325e-04f);
  q = fmaf(q, t, -2.841062484e-02f);
  q = fmaf(q, t, 1.485603089e-01f);
  q = fmaf(q, t, 9.184083273e-01f);
  q = fmaf(q, t, 1.627907927e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-q * t));
  return fmaf(-0.5f * ax, e, fmaxf(x, 0.f));
}

}  // namespace ace
