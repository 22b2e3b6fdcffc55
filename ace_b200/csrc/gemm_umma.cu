// tcgen05 / TMA implementation of the GemmOp contract (gemm.cuh) for sm_100a.
//
// One persistent CTA per SM, 256 threads, warp-specialised:
//   warp 0   TMA producer   -- cp.async.bulk.tensor loads of the hi and lo bf16 planes of the A and B
//                              tiles into a multi-stage shared-memory ring (128B/64B hardware swizzle)
//   warp 1   MMA issuer     -- one thread issues tcgen05.mma (M=128, N<=256, K=16, bf16 x bf16 -> fp32
//                              in TMEM); per K-step THREE products: Ahi*Bhi + Ahi*Blo + Alo*Bhi, which
//                              reproduces an fp32 product to ~2^-17 relative (error budget)
//   warp 2   TMEM allocator -- 2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 4-7 epilogue      -- tcgen05.ld TMEM -> registers, bias / addend / residual / GELU, InstanceNorm
//                              statistics via warp-shuffle transpose-reduce + double atomics, split-bf16
//                              and/or fp32 stores straight from registers
//
// A may be K-major (rows of K contiguous) or MN-major (the M index contiguous: activations in NCHW seen
// as [channel][space]); B is K-major.  Out-of-range parts of any tile are zero-filled by TMA, so M, N, K
// need no padding; only 16-byte alignment of the strides is required (umma_eligible()).
//
// Triangular ops (gemm.cuh): the N extent of the MMA shrinks to the non-zero column range of the tile
// (UMMA N is a runtime field of the instruction descriptor) and K-chunks below k_lo are skipped.
#include <cuda.h>

#include <mutex>

#include "gemm.cuh"
#include "ptx.cuh"

namespace ace {

namespace {

struct UmmaParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  GemmOp op;
  int tiles_m, tiles_n;
  int nterms;
  int a_z1_on, a_z2_on, b_z1_on, b_z2_on;  // 0 when the operand does not vary along that batch axis
};

template <int BN_, int BK_, bool A_MN_>
struct Cfg {
  static constexpr int BM = 128, BN = BN_, BK = BK_;
  static constexpr bool A_MN = A_MN_;
  static constexpr int UMMA_K = 16;
  static constexpr int A_PLANE = BM * BK * 2;  // bytes
  static constexpr int B_PLANE = BN * BK * 2;
  static constexpr int STAGE = 2 * (A_PLANE + B_PLANE);
  static constexpr int STAGES = (200 * 1024) / STAGE;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int K_SWZ_BYTES = BK * 2;                        // swizzle span of K-major tiles
  static constexpr uint32_t K_LAYOUT = (K_SWZ_BYTES == 128) ? 2u : 4u;  // UMMA layout code
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE + BAR_BYTES + 1024;  // + alignment slack
  static_assert(BK == 64 || BK == 32, "BK");
  static_assert(BN % 32 == 0 && BN <= 256, "BN");
  static_assert(STAGES >= 2, "pipeline too shallow");
  static_assert(A_PLANE % 1024 == 0 && B_PLANE % 1024 == 0, "tiles must keep 1024B alignment");
};

struct Tile {
  int m0, n_begin, n_count, n_lo, n_end, z1, z2, k_begin, num_kc;
};

template <int BN, int BK>
__device__ __forceinline__ bool decode_tile(const UmmaParams& p, long long t, Tile& ti) {
  const GemmOp& op = p.op;
  int tn = (int)(t % p.tiles_n);
  long long r = t / p.tiles_n;
  int tm = (int)(r % p.tiles_m);
  r /= p.tiles_m;
  ti.z2 = (int)(r % op.Z2);
  ti.z1 = (int)(r / op.Z2);
  const int n_lo = op.n_lo_z1 ? ti.z1 : 0;
  const int n_hi = op.n_hi_z1 ? min(op.N, ti.z1 + 1) : op.N;
  const int k_lo = op.k_lo_z1 ? ti.z1 : 0;
  ti.m0 = tm * 128;
  ti.n_begin = max(tn * BN, (n_lo / 16) * 16);
  ti.n_end = min(tn * BN + BN, n_hi);
  ti.n_lo = n_lo;
  ti.n_count = ti.n_end - ti.n_begin;
  ti.k_begin = (k_lo / BK) * BK;
  ti.num_kc = (op.K - ti.k_begin + BK - 1) / BK;
  return ti.n_count > 0 && ti.n_end > n_lo && ti.num_kc > 0;
}

template <class C>
__global__ void __launch_bounds__(256, 1) gemm_umma_kernel(const __grid_constant__ UmmaParams p) {
  constexpr int BN = C::BN, BK = C::BK, STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;  // 128B swizzle atoms need 1024B alignment
  uint8_t* smem = smem_raw + (sbase - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE);
  const uint32_t bar0 = sbase + STAGES * C::STAGE;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmOp& op = p.op;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmA);
    ptx::prefetch_tensormap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32((const void*)tmem_slot), C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long total = (long long)p.tiles_m * p.tiles_n * op.Z1 * op.Z2;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    Tile ti;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      if (!decode_tile<BN, BK>(p, t, ti)) continue;
      const int az1 = ti.z1 * p.a_z1_on, az2 = ti.z2 * p.a_z2_on, bz1 = ti.z1 * p.b_z1_on, bz2 = ti.z2 * p.b_z2_on;
      for (int kc = 0; kc < ti.num_kc; ++kc) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t fb = full_bar(stage);
        ptx::mbar_arrive_expect_tx(fb, (uint32_t)C::STAGE);
        const uint32_t sA = sbase + stage * C::STAGE;
        const uint32_t sB = sA + 2 * C::A_PLANE;
        const int k0 = ti.k_begin + kc * BK;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          if (!C::A_MN) {
            ptx::tma_load_5d(sA + pl * C::A_PLANE, &p.tmA, fb, k0, ti.m0, az1, az2, pl);
          } else {
            // two 64-wide M atoms, each [BK rows][128 B]
            ptx::tma_load_5d(sA + pl * C::A_PLANE, &p.tmA, fb, ti.m0, k0, az1, az2, pl);
            ptx::tma_load_5d(sA + pl * C::A_PLANE + BK * 128, &p.tmA, fb, ti.m0 + 64, k0, az1, az2, pl);
          }
          ptx::tma_load_5d(sB + pl * C::B_PLANE, &p.tmB, fb, k0, ti.n_begin, bz1, bz2, pl);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    uint32_t it = 0;
    Tile ti;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      if (!decode_tile<BN, BK>(p, t, ti)) continue;
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      ptx::mbar_wait(tempty_bar(as), aph ^ 1u);
      ptx::tc_fence_after();
      const int n_eff = (ti.n_count + 15) & ~15;
      const uint32_t idesc = ptx::instr_desc_bf16(128, n_eff, C::A_MN ? 1 : 0, 0);
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kc = 0; kc < ti.num_kc; ++kc) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sA = sbase + stage * C::STAGE;
        const uint32_t sB = sA + 2 * C::A_PLANE;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          uint64_t a_hi, a_lo;
          if (!C::A_MN) {
            // K-major: rows of BK*2 bytes, 8-row groups SBO apart; K-step = 32 bytes inside the swizzled row
            a_hi = ptx::smem_desc(sA + kk * 32, 0, 8 * C::K_SWZ_BYTES, C::K_LAYOUT);
            a_lo = ptx::smem_desc(sA + C::A_PLANE + kk * 32, 0, 8 * C::K_SWZ_BYTES, C::K_LAYOUT);
          } else {
            // MN-major, 128B swizzle: [k][64 m] atoms; 8-k groups SBO = 1024 B apart, the second 64-m atom
            // LBO = BK*128 B away; one K-step = 16 k-rows = 2048 B
            a_hi = ptx::smem_desc(sA + kk * 2048, BK * 128, 1024, 2u);
            a_lo = ptx::smem_desc(sA + C::A_PLANE + kk * 2048, BK * 128, 1024, 2u);
          }
          const uint64_t b_hi = ptx::smem_desc(sB + kk * 32, 0, 8 * C::K_SWZ_BYTES, C::K_LAYOUT);
          const uint64_t b_lo = ptx::smem_desc(sB + C::B_PLANE + kk * 32, 0, 8 * C::K_SWZ_BYTES, C::K_LAYOUT);
          ptx::umma_bf16(tmem_d, a_hi, b_hi, idesc, (kc | kk) != 0 ? 1u : 0u);
          if (p.nterms == 3) {
            ptx::umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
            ptx::umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
          }
        }
        ptx::umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      ptx::umma_commit(tfull_bar(as));  // accumulator ready for the epilogue
      ++it;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may access
    const EpiParams& e = op.epi;
    uint32_t it = 0;
    Tile ti;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      if (!decode_tile<BN, BK>(p, t, ti)) continue;
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      ptx::mbar_wait(tfull_bar(as), aph);
      ptx::tc_fence_after();
      const int row = ti.m0 + 32 * q + lane;
      const bool row_ok = row < op.M;
      const int m1 = row / e.mdiv, mr = row % e.mdiv;
      const long long off_add = (long long)ti.z2 * e.add_z2 + (long long)m1 * e.add_m1 + (long long)mr * e.add_m0;
      const long long off_res = (long long)ti.z2 * e.res_z2 + (long long)m1 * e.res_m1 + (long long)mr * e.res_m0;
      const long long off_out = (long long)ti.z1 * e.o_z1 + (long long)ti.z2 * e.o_z2 + (long long)m1 * e.o_m1 + (long long)mr * e.o_m0;
      const long long off_f = (long long)ti.z1 * e.f_z1 + (long long)ti.z2 * e.f_z2 + (long long)m1 * e.f_m1 + (long long)mr * e.f_m0;
      const float* bias = e.col_bias + (long long)ti.z2 * e.cb_z2;
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(32 * q) << 16);
      for (int c0 = 0; c0 < ti.n_count; c0 += 32) {
        float v[32];
        ptx::tmem_ld_32x32(taddr + c0, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = ti.n_begin + c0 + j;
          const bool ok = row_ok && n >= ti.n_lo && n < ti.n_end;
          float x = v[j];
          if (ok) {
            if (e.flags & EPI_COL_BIAS) x += __ldg(bias + n);
            if (e.flags & EPI_ADD_F32) x += __ldg(e.add + off_add + (long long)n * e.add_n);
            if (e.flags & EPI_RES_PLANES) {
              const bf16* r = e.res + off_res + (long long)n * e.res_n;
              x += __bfloat162float(r[0]) + __bfloat162float(r[e.res_plane]);
            }
            if (e.flags & EPI_GELU) x = gelu_erf(x);
            if (e.flags & EPI_OUT_PLANES) {
              bf16 hi, lo;
              split_bf16(x, hi, lo);
              bf16* o = e.out + off_out + (long long)n * e.o_n;
              o[0] = hi;
              o[e.out_plane] = lo;
            }
            if (e.flags & EPI_OUT_F32) e.outf[off_f + (long long)n * e.f_n] = x;
          } else {
            x = 0.f;
          }
          v[j] = x;
        }
        if (e.flags & EPI_STATS) {
          // per-column sums over this warp's 32 rows: butterfly transpose-reduce, 31 shuffles per quantity;
          // afterwards lane j holds the total of column c0 + j
          float s[32], sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { s[j] = v[j]; sq[j] = v[j] * v[j]; }
#pragma unroll
          for (int w = 16; w >= 1; w >>= 1) {
            const bool up = (lane & w) != 0;
#pragma unroll
            for (int j = 0; j < w; ++j) {
              float keep = up ? s[j + w] : s[j], send = up ? s[j] : s[j + w];
              s[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
              float keepq = up ? sq[j + w] : sq[j], sendq = up ? sq[j] : sq[j + w];
              sq[j] = keepq + __shfl_xor_sync(0xffffffffu, sendq, w);
            }
          }
          const int n = ti.n_begin + c0 + lane;
          if (n >= ti.n_lo && n < ti.n_end) {
            double* st = e.stats + ((long long)ti.z2 * e.stats_z2 + n) * 2;
            atomicAdd(st, (double)s[0]);
            atomicAdd(st + 1, (double)sq[0]);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
      ++it;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  if (!fn) throw Error(ACE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    ACE_CHECK_CUDA(cudaGetDevice(&dev));
    ACE_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

// 5-D map: (inner, outer, z1, z2, plane).  Batch axes the operand does not vary along get extent 1.
void make_tmap(CUtensorMap* tm, const Operand& o, bool mn_major, long long rows, long long kext, int Z1, int Z2,
               int box_inner, int box_outer, CUtensorMapSwizzle swz, int* z1_on, int* z2_on, const char* name) {
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  if (!mn_major) {
    dims[0] = (cuuint64_t)kext;
    dims[1] = (cuuint64_t)rows;
    strides[0] = (cuuint64_t)o.s_row * 2;
  } else {
    dims[0] = (cuuint64_t)rows;
    dims[1] = (cuuint64_t)kext;
    strides[0] = (cuuint64_t)o.s_k * 2;
  }
  *z1_on = (Z1 > 1 && o.s_z1 != 0) ? 1 : 0;
  *z2_on = (Z2 > 1 && o.s_z2 != 0) ? 1 : 0;
  dims[2] = *z1_on ? (cuuint64_t)Z1 : 1;
  strides[1] = *z1_on ? (cuuint64_t)o.s_z1 * 2 : strides[0] * dims[1];
  dims[3] = *z2_on ? (cuuint64_t)Z2 : 1;
  strides[2] = *z2_on ? (cuuint64_t)o.s_z2 * 2 : strides[1] * dims[2];
  dims[4] = 2;
  strides[3] = (cuuint64_t)o.plane * 2;
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)o.ptr, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(ACE_ERR_CUDA,
                strprintf("gemm %s: cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu,%llu) box=(%u,%u)",
                          name, (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                          (unsigned long long)dims[3], (unsigned long long)dims[4], (unsigned long long)strides[0],
                          (unsigned long long)strides[1], (unsigned long long)strides[2], (unsigned long long)strides[3],
                          box[0], box[1]));
}

template <class C>
void launch(const GemmOp& op, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    ACE_CHECK_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  UmmaParams p;
  p.op = op;
  p.tiles_m = (op.M + 127) / 128;
  p.tiles_n = (op.N + C::BN - 1) / C::BN;
  p.nterms = options().split_terms;
  const bool a_mn = C::A_MN;
  const CUtensorMapSwizzle kswz = (C::K_SWZ_BYTES == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (!a_mn)
    make_tmap(&p.tmA, op.A, false, op.M, op.K, op.Z1, op.Z2, C::BK, 128, kswz, &p.a_z1_on, &p.a_z2_on, op.name);
  else
    make_tmap(&p.tmA, op.A, true, op.M, op.K, op.Z1, op.Z2, 64, C::BK, CU_TENSOR_MAP_SWIZZLE_128B, &p.a_z1_on, &p.a_z2_on, op.name);
  make_tmap(&p.tmB, op.B, false, op.N, op.K, op.Z1, op.Z2, C::BK, C::BN, kswz, &p.b_z1_on, &p.b_z2_on, op.name);
  long long total = (long long)p.tiles_m * p.tiles_n * op.Z1 * op.Z2;
  int grid = (int)std::min<long long>(total, sm_count());
  gemm_umma_kernel<C><<<grid, 256, C::SMEM_BYTES, stream>>>(p);
  after_launch(op.name);
  g_umma_count.fetch_add(1, std::memory_order_relaxed);
}

template <int BK, bool A_MN>
void launch_bn(const GemmOp& op, int bn, cudaStream_t stream) {
  switch (bn) {
    case 128: launch<Cfg<128, BK, A_MN>>(op, stream); break;
    case 192: launch<Cfg<192, BK, A_MN>>(op, stream); break;
    default: launch<Cfg<256, BK, A_MN>>(op, stream); break;
  }
}

bool aligned8(long long v) { return (v & 7) == 0; }

}  // namespace

bool umma_eligible(const GemmOp& op, const char** why) {
  auto fail = [&](const char* w) {
    if (why) *why = w;
    return false;
  };
  if (op.B.s_k != 1) return fail("B is not K-major");
  const bool a_k = op.A.s_k == 1, a_mn = op.A.s_row == 1;
  if (!a_k && !a_mn) return fail("A is neither K-major nor MN-major");
  if (a_k && !aligned8(op.A.s_row)) return fail("A row stride not 16B aligned");
  if (!a_k && !aligned8(op.A.s_k)) return fail("A k stride not 16B aligned");
  if (!aligned8(op.B.s_row)) return fail("B row stride not 16B aligned");
  if (!aligned8(op.A.plane) || !aligned8(op.B.plane)) return fail("plane offset not 16B aligned");
  if (!aligned8(op.A.s_z1) || !aligned8(op.A.s_z2) || !aligned8(op.B.s_z1) || !aligned8(op.B.s_z2)) return fail("batch stride not 16B aligned");
  if ((((uintptr_t)op.A.ptr) & 15) || (((uintptr_t)op.B.ptr) & 15)) return fail("base pointer not 16B aligned");
  if (op.M < 1 || op.N < 1 || op.K < 1) return fail("empty");
  if (op.A.plane <= 0 || op.B.plane <= 0) return fail("plane offset must be positive");
  return true;
}

void run_gemm_umma(const GemmOp& op, cudaStream_t stream) {
  const char* why = nullptr;
  ACE_REQUIRE(umma_eligible(op, &why), "gemm %s: not eligible for the tcgen05 kernel: %s", op.name, why);
  // N tile: least padded columns, ties to the larger tile (fewer re-reads of A)
  int best = 128;
  long long best_waste = -1;
  for (int bn : {128, 192, 256}) {
    long long tiles = (op.N + bn - 1) / bn;
    long long waste = tiles * bn - op.N;
    if (best_waste < 0 || waste < best_waste || (waste == best_waste && bn > best)) {
      best = bn;
      best_waste = waste;
    }
  }
  if (options().umma_bn) best = options().umma_bn;
  const bool a_mn = !(op.A.s_k == 1);
  const int bk = options().umma_bk;
  if (bk == 32) {
    if (a_mn) launch_bn<32, true>(op, best, stream); else launch_bn<32, false>(op, best, stream);
  } else {
    if (a_mn) launch_bn<64, true>(op, best, stream); else launch_bn<64, false>(op, best, stream);
  }
}

}  // namespace ace
