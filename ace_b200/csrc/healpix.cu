// HEALPix spherical harmonic transform: per-ring real DFT + phase shift, then the SAME
// Legendre contraction as the lat-lon transform (tcgen05 GEMMs of sht.cu) over the 4*nside - 1 iso-latitude rings.
//
// Replaces /root/reference/fme/core/cuhpx/sht.py:32-153 (SHT / iSHT) with tools.py:34-83 (healpix_rfft_torch /
// healpix_irfft_torch: a Python loop of 4*nside - 1 torch.fft calls) by ONE kernel per direction for the ring stage.
// Pixels are in RING order; the Legendre tables (no Condon-Shortley sign, ring quadrature weights folded in) come from
// the host exactly like the lat-lon plan's (ace_sht_plan_create with nlat = 4*nside - 1).
#include "kernels.cuh"
#include "sht.cuh"

namespace ace {
namespace {

constexpr int kFT = 8;      // fields per block
constexpr int kMaxPhi = 1024;  // 4*nside <= 1024

struct Ring {
  int nphi, start;
  double phi0;
};
__host__ __device__ inline Ring ring_of(int t, int nside) {
  Ring r;
  const int npc = 2 * nside * (nside - 1);  // pixels in the north polar cap (rings 0 .. nside-2)
  if (t < nside - 1) {
    r.nphi = 4 * (t + 1);
    r.start = 2 * t * (t + 1);
    r.phi0 = 3.14159265358979323846 / (2.0 * (t + 1)) * 0.5;
  } else if (t <= 3 * nside - 1) {
    r.nphi = 4 * nside;
    r.start = npc + (t - (nside - 1)) * 4 * nside;
    r.phi0 = 3.14159265358979323846 / (2.0 * nside) * 0.5 * (double)((t - nside + 2) % 2);
  } else {
    const int s = 4 * nside - t - 1;  // rings left to the south pole, incl. this one
    r.nphi = 4 * s;
    r.start = 12 * nside * nside - 2 * s * (s + 1);
    r.phi0 = 3.14159265358979323846 / (2.0 * s) * 0.5;
  }
  return r;
}

// x [C][npix] fp32 -> X1 planes [(2m + reim)][C][Kp] at column t:  X[m] = e^{-i m phi0} sum_j x[j] e^{-2 pi i j m / nphi}
// for m < min(nphi/2 + 1, L), zero beyond (tools.py:44-55).  grid (rings, ceil(C / kFT)), 256 threads.
__global__ void __launch_bounds__(256) hpx_ring_dft_fwd_kernel(const float* __restrict__ x, int C, int nside, int L, int Kp,
                                                              bf16* __restrict__ x1, long long plane) {
  __shared__ float sx[kFT][kMaxPhi];
  __shared__ float2 tw[kMaxPhi];
  const int t = blockIdx.x, c0 = blockIdx.y * kFT;
  const Ring r = ring_of(t, nside);
  const long long npix = 12LL * nside * nside;
  for (int i = threadIdx.x; i < r.nphi; i += blockDim.x) {
    double s, c;
    sincospi(2.0 * (double)i / (double)r.nphi, &s, &c);
    tw[i] = make_float2((float)c, (float)s);
  }
  for (int i = threadIdx.x; i < kFT * r.nphi; i += blockDim.x) {
    const int f = i / r.nphi, j = i - f * r.nphi;
    sx[f][j] = (c0 + f < C) ? x[(long long)(c0 + f) * npix + r.start + j] : 0.f;
  }
  __syncthreads();
  const int nm = min(r.nphi / 2 + 1, L);
  for (int o = threadIdx.x; o < kFT * L; o += blockDim.x) {
    const int f = o / L, m = o - f * L;
    if (c0 + f >= C) continue;
    float re = 0.f, im = 0.f;
    if (m < nm) {
      int q = 0;  // (j * m) mod nphi
      for (int j = 0; j < r.nphi; ++j) {
        const float2 w = tw[q];
        const float v = sx[f][j];
        re = fmaf(v, w.x, re);
        im = fmaf(-v, w.y, im);
        q += m;
        if (q >= r.nphi) q -= r.nphi;
      }
      double ps, pc;
      sincos(-(double)m * r.phi0, &ps, &pc);
      const float a = re * (float)pc - im * (float)ps, b = re * (float)ps + im * (float)pc;
      re = a;
      im = b;
    }
    bf16 h, l;
    bf16* d = x1 + ((long long)(2 * m) * C + (c0 + f)) * Kp + t;
    split_bf16(re, h, l);
    d[0] = h;
    d[plane] = l;
    d += (long long)C * Kp;
    split_bf16(im, h, l);
    d[0] = h;
    d[plane] = l;
  }
}

// g planes [(2m + reim)][C][K] at column t -> y [C][npix]:  irfft(n = nphi, norm = "forward") of G[m] e^{+i m phi0}
// (tools.py:58-76): modes above nphi/2 are dropped, the imaginary parts of m = 0 and of the Nyquist mode are ignored.
__global__ void __launch_bounds__(256) hpx_ring_dft_inv_kernel(const bf16* __restrict__ g, long long plane, int C, int nside,
                                                              int L, int K, float* __restrict__ y) {
  __shared__ float2 sg[kFT][kMaxPhi / 2 + 1];
  __shared__ float2 tw[kMaxPhi];
  const int t = blockIdx.x, c0 = blockIdx.y * kFT;
  const Ring r = ring_of(t, nside);
  const long long npix = 12LL * nside * nside;
  const int nyq = r.nphi / 2;
  const int nm = min(nyq + 1, L);
  for (int i = threadIdx.x; i < r.nphi; i += blockDim.x) {
    double s, c;
    sincospi(2.0 * (double)i / (double)r.nphi, &s, &c);
    tw[i] = make_float2((float)c, (float)s);
  }
  for (int i = threadIdx.x; i < kFT * nm; i += blockDim.x) {
    const int f = i / nm, m = i - f * nm;
    float2 v = make_float2(0.f, 0.f);
    if (c0 + f < C) {
      const bf16* p = g + ((long long)(2 * m) * C + (c0 + f)) * K + t;
      const float gr = __bfloat162float(p[0]) + __bfloat162float(p[plane]);
      p += (long long)C * K;
      const float gi = __bfloat162float(p[0]) + __bfloat162float(p[plane]);
      double ps, pc;
      sincos((double)m * r.phi0, &ps, &pc);
      v.x = gr * (float)pc - gi * (float)ps;
      v.y = gr * (float)ps + gi * (float)pc;
      // Hermitian weights of the c2r transform
      const float wgt = (m == 0 || m == nyq) ? 1.f : 2.f;
      v.x *= wgt;
      v.y = (m == 0 || m == nyq) ? 0.f : v.y * wgt;
    }
    sg[f][m] = v;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kFT * r.nphi; o += blockDim.x) {
    const int f = o / r.nphi, j = o - f * r.nphi;
    if (c0 + f >= C) continue;
    float acc = 0.f;
    int q = 0;  // (j * m) mod nphi
    for (int m = 0; m < nm; ++m) {
      const float2 w = tw[q], v = sg[f][m];
      acc = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc));
      q += j;
      if (q >= r.nphi) q -= r.nphi;
    }
    y[(long long)(c0 + f) * npix + r.start + j] = acc;
  }
}

}  // namespace
}  // namespace ace

using namespace ace;

// The Legendre stages run on an ace_sht_plan created with nlat = 4*nside - 1 (the rings), nlon = 4*nside; its lat-lon DFT
// matrices are simply unused.  Workspaces are the plan's own (same layouts as the lat-lon transform).
static void hpx_check(const ace_sht_plan* p, int nside, long long nfields) {
  ACE_REQUIRE(p != nullptr, "ace_hpx: null plan");
  ACE_REQUIRE(nside >= 1 && 4 * nside <= kMaxPhi, "ace_hpx: nside must be in [1, %d]", kMaxPhi / 4);
  ACE_REQUIRE(p->K == 4 * nside - 1, "ace_hpx: plan has %d rings, nside %d needs %d", p->K, nside, 4 * nside - 1);
  ACE_REQUIRE(nfields > 0 && nfields < (1 << 20), "ace_hpx: bad nfields %lld", nfields);
}

extern "C" int ace_sht_plan_reserve(ace_sht_plan* plan, long long nfields, void* stream);

extern "C" int ace_hpx_forward(ace_sht_plan* plan, int nside, const float* x_dev, float* coeffs_dev, long long nfields,
                               void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(x_dev && coeffs_dev, "ace_hpx_forward: null argument");
  hpx_check(plan, nside, nfields);
  cudaStream_t s = (cudaStream_t)stream;
  ace_sht_plan& p = *plan;
  ACE_REQUIRE(ace_sht_plan_reserve(plan, nfields, stream) == ACE_OK, "ace_hpx_forward: workspace");
  const int C = (int)nfields;
  const long long x1p = (long long)(p.ws_x1.bytes / sizeof(bf16) / 2), cp = (long long)(p.ws_c1.bytes / sizeof(bf16) / 2);
  {
    ProfileScope prof("hpx.ring_dft_fwd", s);
    dim3 grid(p.K, (C + kFT - 1) / kFT);
    hpx_ring_dft_fwd_kernel<<<grid, 256, 0, s>>>(x_dev, C, nside, p.M, p.Kp, p.ws_x1.as<bf16>(), x1p);
    after_launch("hpx_ring_dft_fwd");
  }
  run_gemm(sht_op_legendre_fwd(p, p.ws_x1.as<bf16>(), x1p, C, 1, p.ws_c1.as<bf16>(), cp), s);
  launch_spec_planes_to_complex(p.ws_c1.as<bf16>(), cp, C, p.L, p.M, coeffs_dev, s);
  ACE_API_END
}

extern "C" int ace_hpx_inverse(ace_sht_plan* plan, int nside, const float* coeffs_dev, float* x_dev, long long nfields,
                               void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(x_dev && coeffs_dev, "ace_hpx_inverse: null argument");
  hpx_check(plan, nside, nfields);
  cudaStream_t s = (cudaStream_t)stream;
  ace_sht_plan& p = *plan;
  ACE_REQUIRE(ace_sht_plan_reserve(plan, nfields, stream) == ACE_OK, "ace_hpx_inverse: workspace");
  const int C = (int)nfields;
  const long long cp = (long long)(p.ws_c2.bytes / sizeof(bf16) / 2), gp = (long long)(p.ws_g.bytes / sizeof(bf16) / 2);
  launch_spec_complex_to_planes(coeffs_dev, C, p.L, p.M, p.Lp, p.ws_c2.as<bf16>(), cp, s);
  run_gemm(sht_op_legendre_inv(p, p.ws_c2.as<bf16>(), cp, C, 1, p.ws_g.as<bf16>(), gp), s);
  {
    ProfileScope prof("hpx.ring_dft_inv", s);
    dim3 grid(p.K, (C + kFT - 1) / kFT);
    hpx_ring_dft_inv_kernel<<<grid, 256, 0, s>>>(p.ws_g.as<bf16>(), gp, C, nside, p.M, p.K, x_dev);
    after_launch("hpx_ring_dft_inv");
  }
  ACE_API_END
}
