// SHT plan construction and the standalone transform API (ace_sht_*).
#include "sht.cuh"

#include <cmath>
#include <vector>

#include "kernels.cuh"

namespace ace {

namespace {

// host-side split of an fp64 value that has already been rounded to what the reference would hold
inline void host_split(double v, bf16& hi, bf16& lo) {
  float f = (float)v;
  split_bf16(f, hi, lo);
}

void upload_planes(DevBuf& buf, const std::vector<bf16>& hi, const std::vector<bf16>& lo, long long& plane) {
  plane = (long long)hi.size();
  buf.ensure(2 * hi.size() * sizeof(bf16));
  ACE_CHECK_CUDA(cudaMemcpy(buf.as<bf16>(), hi.data(), hi.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  ACE_CHECK_CUDA(cudaMemcpy(buf.as<bf16>() + plane, lo.data(), lo.size() * sizeof(bf16), cudaMemcpyHostToDevice));
}

}  // namespace

GemmOp sht_op_dft_fwd(const ace_sht_plan& p, const bf16* x, long long x_plane, long long x_batch_stride, int C, int B,
                      bf16* x1, long long x1_plane, int rows_per_channel) {
  GemmOp op = make_gemm_op("sht.dft_fwd");
  const int Kq = rows_per_channel > 0 ? rows_per_channel : p.K;
  op.bk_hint = 64;  // 128-byte TMA rows of the streaming operand (measured: 113 -> 91 us at 180x360x384)
  op.M = C * Kq;
  op.N = 2 * p.M;
  op.K = p.W;
  op.Z2 = B;
  op.A = {x, x_plane, (long long)p.W, 1, 0, x_batch_stride};
  op.B = {p.fdft.as<bf16>(), p.fdft_plane, (long long)p.Wp, 1, 0, 0};
  op.epi.flags = EPI_OUT_PLANES;
  op.epi.mdiv = Kq;
  op.epi.out = x1;
  op.epi.out_plane = x1_plane;
  op.epi.o_z2 = p.x1_elems(C);
  op.epi.o_n = (long long)C * p.Kp;
  op.epi.o_m1 = p.Kp;
  op.epi.o_m0 = 1;
  return op;
}

GemmOp sht_op_legendre_fwd(const ace_sht_plan& p, const bf16* x1, long long x1_plane, int C, int B, bf16* c1,
                           long long c1_plane) {
  GemmOp op = make_gemm_op("sht.legendre_fwd");
  op.bk_hint = 64;
  op.M = 2 * C;
  op.N = p.L;
  op.K = p.K;
  op.Z1 = p.M;
  op.Z2 = B;
  op.A = {x1, x1_plane, (long long)p.Kp, 1, 2LL * C * p.Kp, p.x1_elems(C)};
  op.B = {p.wt.as<bf16>(), p.wt_plane, (long long)p.Kt, 1, (long long)p.L * p.Kt, 0};
  op.n_lo_z1 = 1;  // P_l^m = 0 for l < m
  op.epi.flags = EPI_OUT_PLANES;
  op.epi.out = c1;
  op.epi.out_plane = c1_plane;
  op.epi.o_z2 = p.c1_elems(C);
  op.epi.o_n = (long long)p.M * 2 * C;
  op.epi.o_z1 = 2LL * C;
  op.epi.o_m0 = 1;
  return op;
}

GemmOp sht_op_legendre_inv(const ace_sht_plan& p, const bf16* c2, long long c2_plane, int C, int B, bf16* g,
                           long long g_plane) {
  GemmOp op = make_gemm_op("sht.legendre_inv");
  op.M = 2 * C;
  op.N = p.K;
  op.K = p.L;
  op.Z1 = p.M;
  op.Z2 = B;
  op.A = {c2, c2_plane, 1, 2LL * C, (long long)p.Lp * 2 * C, p.c2_elems(C)};  // MN-major
  op.B = {p.pinv.as<bf16>(), p.pinv_plane, (long long)p.Lt, 1, (long long)p.Kg * p.Lt, 0};
  op.k_lo_z1 = 1;  // coefficients with l < m are zero
  op.epi.flags = EPI_OUT_PLANES;
  op.epi.out = g;  // row = reim*C + c  ->  g[(2m + reim)][c][k]: affine in the row index
  op.epi.out_plane = g_plane;
  op.epi.o_z2 = p.g_elems(C);
  op.epi.o_z1 = 2LL * C * p.K;
  op.epi.o_m0 = p.K;
  op.epi.o_n = 1;
  return op;
}

GemmOp sht_op_dft_inv(const ace_sht_plan& p, const bf16* g, long long g_plane, int C, int B, float* y,
                      long long y_batch_stride) {
  GemmOp op = make_gemm_op("sht.dft_inv");
  op.M = C * p.K;
  op.N = p.W;
  op.K = 2 * p.M;
  op.Z2 = B;
  op.A = {g, g_plane, 1, (long long)C * p.K, 0, p.g_elems(C)};  // MN-major
  op.B = {p.idft.as<bf16>(), p.idft_plane, (long long)p.K2p, 1, 0, 0};
  op.epi.flags = EPI_OUT_F32;
  op.epi.outf = y;
  op.epi.f_z2 = y_batch_stride;
  op.epi.f_m0 = p.W;
  op.epi.f_n = 1;
  return op;
}

GemmOp sht_op_legendre_inv2(const ace_sht_plan& p, const bf16* c2, long long c2_plane, int C, int B, bf16* g2, long long g2_plane) {
  GemmOp op = sht_op_legendre_inv(p, c2, c2_plane, C, B, g2, g2_plane);
  op.N = p.Kg;  // the table's rows k >= K are zero: the pad columns of g2 receive exact zeros and every epilogue chunk is full
  op.epi.o_z2 = p.g2_elems(C);
  op.epi.o_z1 = 2LL * C * p.Kg;              // even orders: k index 2 (m / 2) + reim
  op.epi.o_z1b = (long long)p.Ke * C * p.Kg;  // odd orders start at k index Ke
  op.epi.o_m0 = p.Kg;
  return op;
}

GemmOp sht_op_dft_inv2(const ace_sht_plan& p, const bf16* g2, long long g2_plane, int C, int B, float* y, long long y_batch_stride) {
  GemmOp op = make_gemm_op("sht.dft_inv");
  op.M = C * p.Kg;
  op.K = p.K2;
  op.Z2 = B;
  op.A = {g2, g2_plane, 1, (long long)C * p.Kg, 0, p.g2_elems(C)};  // MN-major
  op.B = {p.idft2.as<bf16>(), p.idft2_plane, (long long)p.K2q, 1, 0, 0};
  if (p.W % 2 == 0 && p.M > 1) {
    op.N = p.W / 2;
    op.bfly = 1;
    op.k_split = p.Ke;
  } else {
    op.N = p.W;
  }
  op.epi.flags = EPI_OUT_F32;
  op.epi.mdiv = p.Kg;  // row = c * Kg + k; rows k >= K are padding
  op.epi.mrows = p.K;
  op.epi.outf = y;
  op.epi.f_z2 = y_batch_stride;
  op.epi.f_m1 = (long long)p.K * p.W;
  op.epi.f_m0 = p.W;
  op.epi.f_n = 1;
  return op;
}

GemmOp sht_op_legendre_inv_from_c1(const ace_sht_plan& p, const bf16* c1, long long c1_plane, int C, int B, bf16* g,
                                   long long g_plane) {
  GemmOp op = sht_op_legendre_inv(p, c1, c1_plane, C, B, g, g_plane);
  op.name = "sht.legendre_inv_res";
  // c1 is [B][L][M][2C]: degree l strides over M*2C, order m (z1) over 2C; entries with l < m are exact zeros
  op.A = {c1, c1_plane, 1, (long long)p.M * 2 * C, 2LL * C, p.c1_elems(C)};
  return op;
}

GemmOp sht_op_dft_inv_planes(const ace_sht_plan& p, const bf16* g, long long g_plane, int C, int B, bf16* y,
                             long long y_plane, long long y_batch_stride) {
  GemmOp op = sht_op_dft_inv(p, g, g_plane, C, B, nullptr, 0);
  op.name = "sht.dft_inv_res";
  op.epi.flags = EPI_OUT_PLANES;
  op.epi.outf = nullptr;
  op.epi.out = y;
  op.epi.out_plane = y_plane;
  op.epi.o_z2 = y_batch_stride;
  op.epi.o_m0 = p.W;
  op.epi.o_n = 1;
  return op;
}

}  // namespace ace

using namespace ace;

static unsigned long long fnv1a(const void* data, size_t n, unsigned long long h) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; ++i) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

static void plan_build(ace_sht_plan& p, const double* fwd, const double* inv) {
  const int K = p.K, W = p.W, L = p.L, M = p.M;
  p.table_id = fnv1a(fwd, sizeof(double) * (size_t)M * L * K, 1469598103934665603ull);
  p.table_id = fnv1a(inv, sizeof(double) * (size_t)M * L * K, p.table_id);
  // K-major rows that TMA fetches in 128-byte boxes (X1 and the forward table along k, the inverse table along l) get a pitch of
  // whole 128-byte lines; Lp only counts the (zero) pad rows of c2 and stays at the 16-byte granularity
  p.Kp = (int)round_up(K, 8);
  p.Lp = (int)round_up(L, 8);
  p.Kt = (int)round_up(K, 64);
  p.Lt = (int)round_up(L, 64);
  // row pitch of the DFT matrices: whole 128-byte lines, so that every 128-byte TMA box row (BK = 64) is one aligned L2 line
  // (a pitch of 8 elements leaves most rows straddling two lines: 5 sectors and 2 tag look-ups per box row instead of 4 / 1)
  p.Wp = (int)round_up(W, 64);
  p.K2p = (int)round_up(2 * M, 64);
  p.Kg = (int)round_up(K, 64);
  {
    const int Me = (M + 1) / 2, Mo = M / 2;
    p.Ke = (int)round_up(2 * Me, 64);
    p.K2 = p.Ke + 2 * Mo;
    p.K2a = (int)round_up(p.K2, 8);
    p.K2q = (int)round_up(p.K2, 64);
  }
  {
    std::vector<bf16> hi((size_t)M * L * p.Kt, __float2bfloat16(0.f)), lo(hi.size(), __float2bfloat16(0.f));
    for (int m = 0; m < M; ++m)
      for (int l = 0; l < L; ++l)
        for (int k = 0; k < K; ++k) {
          size_t d = ((size_t)m * L + l) * p.Kt + k;
          host_split(fwd[((size_t)m * L + l) * K + k], hi[d], lo[d]);
        }
    upload_planes(p.wt, hi, lo, p.wt_plane);
  }
  {
    std::vector<bf16> hi((size_t)M * p.Kg * p.Lt, __float2bfloat16(0.f)), lo(hi.size(), __float2bfloat16(0.f));
    for (int m = 0; m < M; ++m)
      for (int l = 0; l < L; ++l)
        for (int k = 0; k < K; ++k) {
          size_t d = ((size_t)m * p.Kg + k) * p.Lt + l;
          host_split(inv[((size_t)m * L + l) * K + k], hi[d], lo[d]);
        }
    upload_planes(p.pinv, hi, lo, p.pinv_plane);
  }
  const double two_pi = 6.283185307179586476925286766559;
  const int nyq = (W % 2 == 0) ? W / 2 : -1;
  {
    // forward rows n = 2m + reim:  X[k][m] = (2 pi / W) sum_j x[k][j] exp(-2 pi i j m / W); modes beyond
    // W/2 do not exist in rfft's output and are zero-padded by fme/fft.py:71-72.
    std::vector<bf16> hi((size_t)2 * M * p.Wp, __float2bfloat16(0.f)), lo(hi.size(), __float2bfloat16(0.f));
    for (int m = 0; m < M && m <= W / 2; ++m)
      for (int j = 0; j < W; ++j) {
        double ang = two_pi * (double)(((long long)j * m) % W) / (double)W;
        double re = std::cos(ang) * two_pi / W;
        double im = (m == 0 || m == nyq) ? 0.0 : -std::sin(ang) * two_pi / W;
        size_t d = (size_t)(2 * m) * p.Wp + j;
        // keep the extra bits of the fp64 value in lo: hi + lo approximates `re` itself
        bf16 h = __float2bfloat16_rn((float)re);
        bf16 l2 = __float2bfloat16_rn((float)(re - (double)__bfloat162float(h)));
        hi[d] = h;
        lo[d] = l2;
        h = __float2bfloat16_rn((float)im);
        l2 = __float2bfloat16_rn((float)(im - (double)__bfloat162float(h)));
        hi[d + p.Wp] = h;
        lo[d + p.Wp] = l2;
      }
    upload_planes(p.fdft, hi, lo, p.fdft_plane);
  }
  {
    // inverse rows j, columns kk = 2m + reim: y[j] = sum_m w_m (Gr cos - Gi sin); Im of m = 0 and of the
    // Nyquist mode are dropped (fme/fft.py:87-92); modes above W/2 are ignored by irfft(n = W).
    std::vector<bf16> hi((size_t)W * p.K2p, __float2bfloat16(0.f)), lo(hi.size(), __float2bfloat16(0.f));
    for (int j = 0; j < W; ++j)
      for (int m = 0; m < M && m <= W / 2; ++m) {
        double wm = (m == 0 || m == nyq) ? 1.0 : 2.0;
        double ang = two_pi * (double)(((long long)j * m) % W) / (double)W;
        double re = wm * std::cos(ang);
        double im = (m == 0 || m == nyq) ? 0.0 : -wm * std::sin(ang);
        size_t d = (size_t)j * p.K2p + 2 * m;
        bf16 h = __float2bfloat16_rn((float)re);
        hi[d] = h;
        lo[d] = __float2bfloat16_rn((float)(re - (double)__bfloat162float(h)));
        h = __float2bfloat16_rn((float)im);
        hi[d + 1] = h;
        lo[d + 1] = __float2bfloat16_rn((float)(im - (double)__bfloat162float(h)));
      }
    upload_planes(p.idft, hi, lo, p.idft_plane);
  }
  {
    // the same rows with the k axis split by the parity of m: k = 2 (m / 2) + reim for even m, Ke + 2 (m / 2) + reim for odd m
    std::vector<bf16> hi((size_t)W * p.K2q, __float2bfloat16(0.f)), lo(hi.size(), __float2bfloat16(0.f));
    for (int j = 0; j < W; ++j)
      for (int m = 0; m < M && m <= W / 2; ++m) {
        double wm = (m == 0 || m == nyq) ? 1.0 : 2.0;
        double ang = two_pi * (double)(((long long)j * m) % W) / (double)W;
        double re = wm * std::cos(ang);
        double im = (m == 0 || m == nyq) ? 0.0 : -wm * std::sin(ang);
        size_t d = (size_t)j * p.K2q + ((m & 1) ? p.Ke : 0) + 2 * (m >> 1);
        bf16 h = __float2bfloat16_rn((float)re);
        hi[d] = h;
        lo[d] = __float2bfloat16_rn((float)(re - (double)__bfloat162float(h)));
        h = __float2bfloat16_rn((float)im);
        hi[d + 1] = h;
        lo[d + 1] = __float2bfloat16_rn((float)(im - (double)__bfloat162float(h)));
      }
    upload_planes(p.idft2, hi, lo, p.idft2_plane);
  }
}

static void plan_ensure_ws(ace_sht_plan& p, long long nf, cudaStream_t s) {
    if (nf < p.ws_fields) {
    // same storage, different layout: restore the all-zero state the triangular ops rely on
    ACE_CHECK_CUDA(cudaMemsetAsync(p.ws_c1.p, 0, p.ws_c1.bytes, s));
    ACE_CHECK_CUDA(cudaMemsetAsync(p.ws_c2.p, 0, p.ws_c2.bytes, s));
    p.ws_fields_layout = nf;
    return;
  }
  int C = (int)nf;
  p.ws_x.ensure(2 * (size_t)nf * p.K * p.W * sizeof(bf16));
  p.ws_x1.ensure(2 * (size_t)p.x1_elems(C) * sizeof(bf16));
  // c1 / c2 are separate: each relies on its never-written region staying zero (see gemm.cuh)
  p.ws_c1.ensure(2 * (size_t)p.c1_elems(C) * sizeof(bf16));
  p.ws_c2.ensure(2 * (size_t)p.c2_elems(C) * sizeof(bf16));
  p.ws_g.ensure(2 * (size_t)p.g_elems(C) * sizeof(bf16));
  p.ws_fields = nf;
  p.ws_fields_layout = nf;
}

// ------------------------------------------------------------------------------------ C ABI

// (internal, used by healpix.cu) make the plan's standalone-transform workspaces valid for `nfields` fields
extern "C" int ace_sht_plan_reserve(ace_sht_plan* plan, long long nfields, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(plan != nullptr, "ace_sht_plan_reserve: null plan");
  if (nfields != plan->ws_fields_layout) plan_ensure_ws(*plan, nfields, (cudaStream_t)stream);
  ACE_API_END
}

extern "C" int ace_sht_plan_create(int nlat, int nlon, int lmax, int mmax, const double* legendre_fwd_host,
                                   const double* legendre_inv_host, ace_sht_plan** out) {
  ACE_API_BEGIN
  ACE_REQUIRE(out != nullptr, "ace_sht_plan_create: out is null");
  ACE_REQUIRE(nlat > 0 && nlon > 0 && lmax > 0 && mmax > 0, "ace_sht_plan_create: non-positive size");
  ACE_REQUIRE(legendre_fwd_host && legendre_inv_host, "ace_sht_plan_create: null table");
  ace_sht_plan* p = new ace_sht_plan();
  p->K = nlat;
  p->W = nlon;
  p->L = lmax;
  p->M = mmax;
  try {
    plan_build(*p, legendre_fwd_host, legendre_inv_host);
  } catch (...) {
    delete p;
    throw;
  }
  *out = p;
  ACE_API_END
}

extern "C" void ace_sht_plan_destroy(ace_sht_plan* plan) { delete plan; }

extern "C" int ace_sht_forward(ace_sht_plan* plan, const float* x_dev, float* coeffs_dev, long long nfields, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(plan && x_dev && coeffs_dev, "ace_sht_forward: null argument");
  ACE_REQUIRE(nfields > 0 && nfields < (1 << 20), "ace_sht_forward: bad nfields %lld", nfields);
  cudaStream_t s = (cudaStream_t)stream;
  ace_sht_plan& p = *plan;
  if (nfields != p.ws_fields_layout) plan_ensure_ws(p, nfields, s);
  const int C = (int)nfields;
  const long long HW = (long long)p.K * p.W;
  const long long xp = (long long)(p.ws_x.bytes / sizeof(bf16) / 2), x1p = (long long)(p.ws_x1.bytes / sizeof(bf16) / 2),
                  cp = (long long)(p.ws_c1.bytes / sizeof(bf16) / 2);
  launch_norm_split(x_dev, 1, C, HW, nullptr, nullptr, nullptr, 0.f, p.ws_x.as<bf16>(), xp, 0, HW, s);
  run_gemm(sht_op_dft_fwd(p, p.ws_x.as<bf16>(), xp, 0, C, 1, p.ws_x1.as<bf16>(), x1p), s);
  run_gemm(sht_op_legendre_fwd(p, p.ws_x1.as<bf16>(), x1p, C, 1, p.ws_c1.as<bf16>(), cp), s);
  launch_spec_planes_to_complex(p.ws_c1.as<bf16>(), cp, C, p.L, p.M, coeffs_dev, s);
  ACE_API_END
}

extern "C" int ace_sht_inverse(ace_sht_plan* plan, const float* coeffs_dev, float* x_dev, long long nfields, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(plan && x_dev && coeffs_dev, "ace_sht_inverse: null argument");
  ACE_REQUIRE(nfields > 0 && nfields < (1 << 20), "ace_sht_inverse: bad nfields %lld", nfields);
  cudaStream_t s = (cudaStream_t)stream;
  ace_sht_plan& p = *plan;
  if (nfields != p.ws_fields_layout) plan_ensure_ws(p, nfields, s);
  const int C = (int)nfields;
  const long long cp = (long long)(p.ws_c2.bytes / sizeof(bf16) / 2), gp = (long long)(p.ws_g.bytes / sizeof(bf16) / 2);
  // the pad rows l in [L, Lp) of c2 stay zero from allocation; the l < m region is overwritten with the
  // caller's values, which is what the reference contracts too (its table is zero there)
  launch_spec_complex_to_planes(coeffs_dev, C, p.L, p.M, p.Lp, p.ws_c2.as<bf16>(), cp, s);
  GemmOp inv = sht_op_legendre_inv(p, p.ws_c2.as<bf16>(), cp, C, 1, p.ws_g.as<bf16>(), gp);
  run_gemm(inv, s);
  run_gemm(sht_op_dft_inv(p, p.ws_g.as<bf16>(), gp, C, 1, x_dev, 0), s);
  ACE_API_END
}
