// Generic SIMT implementation of the GemmOp contract (gemm.cuh).
//
// Serves (a) shapes the tcgen05 kernel cannot take (strides not 16-byte aligned: the 9x18
// golden grids), (b) the on-device cross-check of the tcgen05 kernel in the GPU tests.  It
// reconstructs fp32 values from the split planes and accumulates with FFMA, i.e. it computes
// the full (a_hi+a_lo)*(b_hi+b_lo) product -- a superset of the 3-term tensor-core product.
#include "gemm.cuh"

namespace ace {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, NT = 256;

__global__ void __launch_bounds__(NT) gemm_simt_kernel(const __grid_constant__ GemmOp op) {
  const int z1 = blockIdx.y, z2 = blockIdx.z;
  const int tiles_m = (op.M + BM - 1) / BM;
  const int tile_m = blockIdx.x % tiles_m, tile_n = blockIdx.x / tiles_m;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  const int n_lo = op.n_lo_z1 ? z1 : 0;
  const int n_hi = op.n_hi_z1 ? min(op.N, z1 + 1) : op.N;
  const int k_lo = op.k_lo_z1 ? z1 : 0;
  const int m_hi = op.m_hi_z1 ? min(op.M, z1 + 1) : op.M;
  if (n0 >= n_hi || n0 + BN <= n_lo || m0 >= m_hi) return;

  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const bf16* A = op.A.ptr + (long long)z1 * op.A.s_z1 + (long long)z2 * op.A.s_z2;
  const bf16* B = op.B.ptr + (long long)z1 * op.B.s_z1 + (long long)z2 * op.B.s_z2;
  const bool a_kmajor = (op.A.s_k == 1);
  const bool b_kmajor = (op.B.s_k == 1);

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // tx -> n, ty -> m
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int k_begin = (k_lo / BK) * BK;
  for (int k0 = k_begin; k0 < op.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < BM * BK / NT; ++i) {
      int idx = tid + i * NT;
      int mm, kk;
      if (a_kmajor) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
      int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < op.M && k < op.K && k >= k_lo) {
        if (op.cplx == 2 && op.group_n) {
          // grouped: this tile's columns belong to group g, which contracts k range [g * group_n, + Kh) of each part
          const int Nh = op.N >> 1, Kh = op.K >> 1;
          const int g = (n0 % Nh) / op.group_n, ri = k >= Kh, i2 = k - ri * Kh;
          const bf16* p = A + (long long)m * op.A.s_row + ((long long)ri * op.a_part_k + (long long)g * op.group_n + i2) * op.A.s_k;
          v = __bfloat162float(p[0]) + __bfloat162float(p[op.A.plane]);
        } else if (op.cplx != 1) {
          // (cplx == 2: A is [Ar | Ai] along k, i.e. a plain matrix of the doubled K extent)
          const bf16* p = A + (long long)m * op.A.s_row + (long long)k * op.A.s_k;
          v = __bfloat162float(p[0]) + __bfloat162float(p[op.A.plane]);
        } else {
          // op.M / op.K are the real-ified extents here (run_gemm_simt doubled them): A'[(ro,o)][(ri,i)]
          const int Mh = op.M >> 1, Kh = op.K >> 1;
          const int ro = m >= Mh, o = m - ro * Mh, ri = k >= Kh, i2 = k - ri * Kh;
          const bf16* p = A + (ro != ri ? op.a_part : 0) + (long long)o * op.A.s_row + (long long)i2 * op.A.s_k;
          v = __bfloat162float(p[0]) + __bfloat162float(p[op.A.plane]);
          if (ro == 0 && ri == 1) v = -v;
        }
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < BN * BK / NT; ++i) {
      int idx = tid + i * NT;
      int kk, nn;
      if (b_kmajor) { kk = idx % BK; nn = idx / BK; } else { nn = idx % BN; kk = idx / BN; }
      int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < op.N && k < op.K && k >= k_lo) {
        if (op.cplx != 2) {
          const bf16* p = B + (long long)n * op.B.s_row + (long long)k * op.B.s_k;
          v = __bfloat162float(p[0]) + __bfloat162float(p[op.B.plane]);
        } else {
          // op.N / op.K are the real-ified extents (run_gemm_simt doubled them): B'[(ro,o)][(ri,i)] = [[Br, -Bi], [Bi, Br]]
          const int Nh = op.N >> 1, Kh = op.K >> 1;
          const int ro = n >= Nh, o = n - ro * Nh, ri = k >= Kh, i2 = k - ri * Kh;
          const bf16* p = B + (ro != ri ? op.b_part : 0) + (long long)o * op.B.s_row + (long long)i2 * op.B.s_k;
          v = __bfloat162float(p[0]) + __bfloat162float(p[op.B.plane]);
          if (ro == 0 && ri == 1) v = -v;
        }
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const EpiParams& e = op.epi;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= m_hi) continue;
    int m1 = m / e.mdiv, mr = m % e.mdiv;
    float rsum = 0.f, rsq = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < n_lo || n >= n_hi) continue;
      float v = epi_value(e, acc[i][j], m1, mr, n, z2);
      rsum += v;
      rsq += v * v;
      epi_store(e, v, m1, mr, n, z1, z2);
    }
    if (e.flags & EPI_ROW_STATS) {
      double* st = e.stats + ((long long)z2 * e.stats_z2 + m) * 2;
      atomicAdd(st, (double)rsum);
      atomicAdd(st + 1, (double)rsq);
    }
  }
}

}  // namespace

void run_gemm_simt(const GemmOp& op_in, cudaStream_t stream) {
  GemmOp op = op_in.bfly ? bfly_dense(op_in) : op_in;
  if (op.cplx) {  // the SIMT kernel walks the real-ified problem
    ACE_REQUIRE(!op.k_lo_z1, "gemm %s: complex mode with a triangular K range is not supported", op.name);
    ACE_REQUIRE(op.cplx == 1 || op.cplx == 2, "gemm %s: bad complex mode %d", op.name, op.cplx);
    if (op.cplx == 1) {
      ACE_REQUIRE(!op.m_hi_z1, "gemm %s: complex mode 1 takes no triangular M range", op.name);
      op.M *= 2;
    } else {
      ACE_REQUIRE(!op.n_lo_z1 && !op.n_hi_z1, "gemm %s: complex mode 2 takes no triangular N range", op.name);
      ACE_REQUIRE(op.group_n == 0 || (op.group_n % BN == 0 && op.N % op.group_n == 0 && op.K <= op.group_n),
                  "gemm %s: grouped complex mode needs groups of a multiple of %d columns", op.name, BN);
      op.N *= 2;
    }
    op.K *= 2;
  }
  ACE_REQUIRE(op.M > 0 && op.N > 0 && op.K > 0 && op.Z1 > 0 && op.Z2 > 0, "gemm %s: empty problem", op.name);
  ACE_REQUIRE(op.Z1 <= 65535 && op.Z2 <= 65535, "gemm %s: batch extent too large", op.name);
  int tiles_m = (op.M + BM - 1) / BM, tiles_n = (op.N + BN - 1) / BN;
  dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)op.Z1, (unsigned)op.Z2);
  gemm_simt_kernel<<<grid, NT, 0, stream>>>(op);
  after_launch(op.name);
  g_simt_count.fetch_add(1, std::memory_order_relaxed);
}

}  // namespace ace
