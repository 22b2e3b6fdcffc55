// GemmOp: the one descriptor every GEMM-shaped piece of the hot path is expressed in.
//
//   for z2 in [0,Z2), z1 in [0,Z1):
//     D[m][n] = sum_{k in [k_lo, K)} A_z[m][k] * B_z[n][k]      m in [0,M), n in [n_lo, n_hi)
//     epilogue(D[m][n]) -> row affine / bias / addend / residual / GELU / per-row statistics / stores
//
// Operands live in HBM as "split-bf16 planes": two bf16 arrays (hi at `ptr`, lo at
// `ptr + plane`), value = hi + lo.  Strides are in elements.  A and B may each be K-major
// (s_k == 1) or MN-major (s_row == 1).  The SIMT kernel (gemm_simt.cu) implements this
// contract for any strides and any epilogue flag combination; the tcgen05 kernel
// (gemm_umma.cu) implements the fast subsets the network uses (16-byte aligned strides and
// one of the epilogue shapes listed in umma_eligible()); run_gemm() routes between them.
//
// Triangular structure of the Legendre / dhconv stages is expressed through z1:
//   n_lo_z1: n_lo = z1      (forward Legendre: only degrees l >= m are non-zero)
//   n_hi_z1: n_hi = z1 + 1  (dhconv at degree l: only orders m <= l are non-zero)
//   k_lo_z1: k_lo = z1      (inverse Legendre: sum over l >= m)
//   m_hi_z1: m_hi = z1 + 1  (dhconv at degree l with the orders on the rows)
// Buffers written under these restrictions are zero-initialised once and the
// excluded region only ever receives exact zeros, so tile-granular kernels may
// read or write the excluded part freely.
#pragma once
#include "common.cuh"

namespace ace {

struct Operand {
  const bf16* ptr;     // hi plane
  long long plane;     // lo plane = ptr + plane
  long long s_row;     // stride of m (A) / n (B)
  long long s_k;       // stride of k
  long long s_z1, s_z2;
};

enum EpiFlags : uint32_t {
  EPI_ADD_F32 = 1u << 1,     // v += add[z2*add_z2 + m1*add_m1 + m0*add_m0 + n*add_n]
  EPI_RES_PLANES = 1u << 2,  // v += (res_hi + res_lo)[z2*res_z2 + m1*res_m1 + m0*res_m0 + n*res_n]
  EPI_GELU = 1u << 3,        // v = exact (erf) GELU of v
  EPI_OUT_PLANES = 1u << 5,  // split-bf16 store
  EPI_OUT_F32 = 1u << 6,     // fp32 store
  EPI_ROW_BIAS = 1u << 7,    // v += row_bias[z2*rb_z2 + m]
  EPI_ROW_STATS = 1u << 8,   // stats[(z2*stats_z2 + m)*2 + {0,1}] += {v, v*v}   (double atomics)
  // deferred InstanceNorm (sfno.cu): the operand was stored un-normalised, the affine map is applied here
  EPI_ROW_AFFINE = 1u << 9,  // first of all: v = ra_scale[z2*ra_z2 + m1] * v + (n == 0 ? ra_shift0[z2*ra_z2 + m1] : 0)
  EPI_RES_AFFINE = 1u << 10, // with EPI_RES_PLANES: residual r -> res_a[z2*rsa_z2 + m] * r + res_s[z2*rsa_z2 + m]
};

// Row index m is decomposed as m1 = m / mdiv, m0 = m % mdiv so that flattened
// (channel, latitude) rows can be scattered into padded layouts.
struct EpiParams {
  uint32_t flags;
  int mdiv;
  int mrows;  // 0 = every row exists; otherwise rows with m0 >= mrows are padding: computed, never stored (mdiv > mrows)
  const float* row_bias;
  long long rb_z2;
  const float* ra_scale;
  const float* ra_shift0;
  long long ra_z2;
  const float* res_a;
  const float* res_s;
  long long rsa_z2;
  const float* add;
  long long add_z2, add_m1, add_m0, add_n;
  const bf16* res;
  long long res_plane, res_z2, res_m1, res_m0, res_n;
  double* stats;
  long long stats_z2;
  bf16* out;
  long long out_plane, o_z1, o_z2, o_m1, o_m0, o_n;
  // o_z1b != 0: the z1 axis is stored split by parity -- offset = (z1 >> 1) * o_z1 + (z1 & 1) * o_z1b (inverse Legendre stage:
  // even orders m first, then the odd ones, which is the k order the butterfly inverse DFT wants)
  long long o_z1b;
  float* outf;
  long long f_z1, f_z2, f_m1, f_m0, f_n;
  // ConditionalLayerNorm mode (GemmOp::cln): per-column {mean, rstd} of the channel LayerNorm, per-row affine of that norm and
  // the per-(sample, row) {scale0, bias0} of the Linear context terms (null: 1, 0)
  const float2* cln_musr;
  long long cln_musr_z2;
  const float* cln_lnw;
  const float* cln_lnb;
  const float* cln_sb0;
  long long cln_sb0_z2;
};

struct GemmOp {
  int M, N, K;
  int Z1, Z2;
  Operand A, B;
  int n_lo_z1, n_hi_z1, k_lo_z1;
  int m_hi_z1;  // rows m > z1 are excluded (dhconv with the modes on the rows: only orders m <= degree l exist)
  // Complex mode (dhconv): A holds two real matrices Ar, Ai of M x K each (Ai = A + a_part elements), B holds
  // [Br | Bi] along its k axis (K + K columns), and the op computes the 2M x N real result
  //   D[0:M]  = Ar Br - Ai Bi,   D[M:2M] = Ai Br + Ar Bi        (rows M..2M-1 are stored at row index m + M)
  // i.e. exactly the real-ified GEMM with A' = [[Ar, -Ai], [Ai, Ar]] without materialising A'.
  // cplx == 2 is the same complex product with the operand roles exchanged (columns contiguous in the output):
  //   A holds [Ar | Ai] along its k axis (K + K columns), B holds two real matrices Br, Bi of N x K each (Bi = B + b_part
  //   elements), and the op computes the M x 2N real result   D[:, 0:N] = Ar Br - Ai Bi,   D[:, N:2N] = Ar Bi + Ai Br
  //   (column n of the imaginary part is stored at column index n + N).
  //   Grouped form (group_n > 0, dhconv with filter_num_groups > 1): the N columns are split into groups of group_n
  //   columns and group g only contracts the k range [g * group_n, g * group_n + K) of each part of A (K = inputs per
  //   group; the imaginary part of A starts a_part_k columns after the real part instead of K): a block-diagonal
  //   operator stored and multiplied as its diagonal blocks.
  int cplx;
  long long a_part, b_part;
  int group_n, a_part_k;
  // Butterfly mode (inverse longitude DFT of an even-length grid, sht.cu): the k axis is split at k_split,
  //   E[m][n] = sum_{k < k_split} A[m][k] B[n][k],   O[m][n] = sum_{k >= k_split} A[m][k] B[n][k],     n in [0, N)
  //   D[m][n] = E + O,   D[m][n + N] = E - O
  // B holds 2N rows; rows [N, 2N) equal rows [0, N) with the k >= k_split part negated, so the op is the plain GEMM with
  // 2N columns at half the multiplications (run_gemm falls back to exactly that when the butterfly kernel is not eligible).
  int bfly, k_split;
  // ConditionalLayerNorm mode (fme/core/models/conditional_sfno/layers.py:285-320; tcgen05 kernel only, csfno.cu falls back to the
  // streaming kernel of cln.cu otherwise): A = [W_scale | W_bias] along k (k_split = K / 2), B = the per-pixel context with its
  // channels twice along k, so the two accumulators are S = W_scale ctx and T = W_bias ctx; the epilogue reads the norm's input x
  // from epi.res (split planes) and stores  y = ((x - mean_n) rstd_n lnw_m + lnb_m) (scale0_m + S) + bias0_m + T  as split planes.
  int cln;
  int bk_hint;  // 0 = kernel default (32); 64 = K extent per pipeline stage for K-major x K-major ops whose A streams from HBM
  EpiParams epi;
  const char* name;  // for error messages / profiling
};

inline GemmOp make_gemm_op(const char* name) {
  GemmOp op;
  memset(&op, 0, sizeof(op));
  op.Z1 = op.Z2 = 1;
  op.epi.mdiv = 1 << 30;
  op.name = name;
  return op;
}

#ifdef __CUDACC__
// Epilogue value: everything up to (and including) GELU.  Shared by both kernels.
__device__ __forceinline__ float epi_value(const EpiParams& e, float acc, int m1, int m0, int n, int z2) {
  float v = acc;
  if (e.flags & EPI_ROW_AFFINE) {
    const long long i = (long long)z2 * e.ra_z2 + m1;
    v = __ldg(e.ra_scale + i) * v + (n == 0 ? __ldg(e.ra_shift0 + i) : 0.f);
  }
  if (e.flags & EPI_ROW_BIAS) v += __ldg(e.row_bias + (long long)z2 * e.rb_z2 + (long long)m1 * e.mdiv + m0);
  if (e.flags & EPI_ADD_F32) v += __ldg(e.add + (long long)z2 * e.add_z2 + (long long)m1 * e.add_m1 + (long long)m0 * e.add_m0 + (long long)n * e.add_n);
  if (e.flags & EPI_RES_PLANES) {
    const bf16* r = e.res + (long long)z2 * e.res_z2 + (long long)m1 * e.res_m1 + (long long)m0 * e.res_m0 + (long long)n * e.res_n;
    float rv = __bfloat162float(r[0]) + __bfloat162float(r[e.res_plane]);
    if (e.flags & EPI_RES_AFFINE) {
      const long long i = (long long)z2 * e.rsa_z2 + (long long)m1 * e.mdiv + m0;
      rv = fmaf(__ldg(e.res_a + i), rv, __ldg(e.res_s + i));
    }
    v += rv;
  }
  if (e.flags & EPI_GELU) v = gelu_erf(v);
  return v;
}

__device__ __forceinline__ long long epi_z1_off(const EpiParams& e, int z1) {
  return e.o_z1b ? (long long)(z1 >> 1) * e.o_z1 + (long long)(z1 & 1) * e.o_z1b : (long long)z1 * e.o_z1;
}

__device__ __forceinline__ void epi_store(const EpiParams& e, float v, int m1, int m0, int n, int z1, int z2) {
  if (e.mrows && m0 >= e.mrows) return;
  if (e.flags & EPI_OUT_PLANES) {
    bf16 hi, lo;
    split_bf16(v, hi, lo);
    bf16* o = e.out + epi_z1_off(e, z1) + (long long)z2 * e.o_z2 + (long long)m1 * e.o_m1 + (long long)m0 * e.o_m0 + (long long)n * e.o_n;
    o[0] = hi;
    o[e.out_plane] = lo;
  }
  if (e.flags & EPI_OUT_F32) {
    e.outf[(long long)z1 * e.f_z1 + (long long)z2 * e.f_z2 + (long long)m1 * e.f_m1 + (long long)m0 * e.f_m0 + (long long)n * e.f_n] = v;
  }
}
#endif

// the plain 2N-column GEMM a butterfly op is shorthand for
inline GemmOp bfly_dense(const GemmOp& op) {
  GemmOp d = op;
  d.N = 2 * op.N;
  d.bfly = 0;
  d.k_split = 0;
  return d;
}

// Dispatcher: tcgen05 kernel when the op is eligible (and not overridden), SIMT kernel otherwise.
void run_gemm(const GemmOp& op, cudaStream_t stream);
void run_gemm_simt(const GemmOp& op, cudaStream_t stream);
// returns false (with reason) if the op cannot run on the tcgen05 kernel
bool umma_eligible(const GemmOp& op, const char** why);
void run_gemm_umma(const GemmOp& op, cudaStream_t stream);
bool run_gemm_cln(const GemmOp& op, cudaStream_t stream);  // ConditionalLayerNorm mode (GemmOp::cln); false = not eligible

}  // namespace ace
