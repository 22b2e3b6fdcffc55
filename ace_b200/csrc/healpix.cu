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

// Ring stage, forward:  X[m] = e^{-i m phi0} sum_j x[j] e^{-2 pi i j m / nphi}  for m < min(nphi/2 + 1, M), zero beyond
// (tools.py:44-55).  Two kernels:
//   hpx_ring_dft_fwd_kernel   x [C][npix] fp32 -> tmp [C][K rings][M] complex fp32.  A block owns one ring and kFT = 8 fields; the
//       ring's samples sit in shared memory as [j][8 fields] (two warp-uniform 16-byte reads per j), a thread owns one order m and
//       keeps the 8 fields' (re, im) sums in registers: 16 FMAs per twiddle read instead of 2.  Stores are contiguous along m.
//   hpx_tmp_to_x1_kernel      tmp -> X1 planes [(2m + reim)][C][Kp] (ring index contiguous, the Legendre GEMM's A operand):
//       32 x 32 tiles transposed through shared memory, so the 2-byte plane stores come in 64-byte runs instead of one sector per
//       element (a block of the one-kernel version owned ONE ring, i.e. one column of X1).
// grid (rings, ceil(C / kFT)), 128 threads.
__global__ void __launch_bounds__(128) hpx_ring_dft_fwd_kernel(const float* __restrict__ x, int C, int nside, int K, int M,
                                                              float2* __restrict__ tmp) {
  __shared__ float4 sx[kMaxPhi][kFT / 4];
  __shared__ float2 tw[kMaxPhi];
  const int t = blockIdx.x, c0 = blockIdx.y * kFT;
  const Ring r = ring_of(t, nside);
  const long long npix = 12LL * nside * nside;
  for (int i = threadIdx.x; i < r.nphi; i += blockDim.x) {
    double s, c;
    sincospi(2.0 * (double)i / (double)r.nphi, &s, &c);
    tw[i] = make_float2((float)c, (float)s);
  }
  float* sxf = reinterpret_cast<float*>(sx);
  for (int i = threadIdx.x; i < kFT * r.nphi; i += blockDim.x) {
    const int f = i / r.nphi, j = i - f * r.nphi;
    sxf[j * kFT + f] = (c0 + f < C) ? x[(long long)(c0 + f) * npix + r.start + j] : 0.f;
  }
  __syncthreads();
  // real input: fold x[j] and x[nphi - j] (same cosine, opposite sine) in place -- slot j keeps e = x[j] + x[n-j], slot n - j
  // o = x[j] - x[n-j] for 0 < j < n/2 -- which halves the multiplications: Re = sum_j e cos, Im = -sum_j o sin
  const int half = r.nphi / 2;  // nphi is a multiple of 4
  for (int i = threadIdx.x; i < kFT * (half - 1); i += blockDim.x) {
    const int f = i % kFT, j = 1 + i / kFT;
    const float a = sxf[j * kFT + f], b = sxf[(r.nphi - j) * kFT + f];
    sxf[j * kFT + f] = a + b;
    sxf[(r.nphi - j) * kFT + f] = a - b;
  }
  __syncthreads();
  const int nm = min(half + 1, M);
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    float re[kFT], im[kFT];
#pragma unroll
    for (int f = 0; f < kFT; ++f) re[f] = im[f] = 0.f;
    if (m < nm) {
      {  // j = 0 and j = n/2: cosine 1 and (-1)^m, no sine term
        const float4 a = sx[0][0], b = sx[0][1], c = sx[half][0], d = sx[half][1];
        const float sg = (m & 1) ? -1.f : 1.f;
        const float v0[kFT] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}, vh[kFT] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
        for (int f = 0; f < kFT; ++f) re[f] = fmaf(vh[f], sg, v0[f]);
      }
      int q = m;  // (j * m) mod nphi
#pragma unroll 2
      for (int j = 1; j < half; ++j) {
        const float2 w = tw[q];
        const float4 a = sx[j][0], b = sx[j][1], c = sx[r.nphi - j][0], d = sx[r.nphi - j][1];
        const float e[kFT] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}, o[kFT] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
        for (int f = 0; f < kFT; ++f) {
          re[f] = fmaf(e[f], w.x, re[f]);
          im[f] = fmaf(-o[f], w.y, im[f]);
        }
        q += m;
        if (q >= r.nphi) q -= r.nphi;
      }
      double ps, pc;
      sincos(-(double)m * r.phi0, &ps, &pc);
      const float fc = (float)pc, fs = (float)ps;
#pragma unroll
      for (int f = 0; f < kFT; ++f) {
        const float a = re[f] * fc - im[f] * fs, b = re[f] * fs + im[f] * fc;
        re[f] = a;
        im[f] = b;
      }
    }
#pragma unroll
    for (int f = 0; f < kFT; ++f)
      if (c0 + f < C) tmp[((long long)(c0 + f) * K + t) * M + m] = make_float2(re[f], im[f]);
  }
}

// tmp [C][K][2M] fp32 (2M = interleaved re / im of the orders) -> X1 planes [(2m + reim)][C][Kp]; grid (C, ceil(2M/32), ceil(K/32))
__global__ void __launch_bounds__(256) hpx_tmp_to_x1_kernel(const float* __restrict__ tmp, int C, int K, int M2, int Kp,
                                                           bf16* __restrict__ x1, long long plane) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.z * 32, r0 = blockIdx.y * 32, f = blockIdx.x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, mr = r0 + tx;
    tile[i][tx] = (t < K && mr < M2) ? tmp[((long long)f * K + t) * M2 + mr] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int mr = r0 + i, t = t0 + tx;
    if (mr < M2 && t < K) {
      bf16 h, l;
      split_bf16(tile[tx][i], h, l);
      bf16* d = x1 + ((long long)mr * C + f) * Kp + t;
      d[0] = h;
      d[plane] = l;
    }
  }
}

// Ring stage, inverse:  irfft(n = nphi, norm = "forward") of G[m] e^{+i m phi0} (tools.py:58-76): modes above nphi/2 are dropped,
// the imaginary parts of m = 0 and of the Nyquist mode are ignored.  Mirror image of the forward pair:
//   hpx_g_to_tmp_kernel       g2 planes [row(m, reim)][C][Kg] -> tmp [C][K][2M] fp32 (tile transposition); g2 is the padded,
//       parity-split layout of the lat-lon networks' inverse Legendre stage (sht.cuh: even orders first, odd orders from row Ke on,
//       Kg = rings padded to 64): with 4 nside - 1 rings the unpadded layout would put that GEMM on element-wise stores
//   hpx_ring_dft_inv_kernel   tmp -> y [C][npix]: a block owns one ring and 8 fields, the ring's phase-shifted, Hermitian-weighted
//       coefficients sit in shared memory as [m][8 fields], a thread owns one pixel and the 8 fields' sums.
__global__ void __launch_bounds__(256) hpx_g_to_tmp_kernel(const bf16* __restrict__ g, long long plane, int C, int K, int M2, int Kg, int Ke,
                                                          float* __restrict__ tmp) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.z * 32, r0 = blockIdx.y * 32, f = blockIdx.x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int mr = r0 + i, t = t0 + tx;
    float v = 0.f;
    if (mr < M2 && t < K) {
      const int row = (((mr >> 1) & 1) ? Ke : 0) + 2 * (mr >> 2) + (mr & 1);  // order m = mr / 2, reim = mr & 1
      const bf16* s = g + ((long long)row * C + f) * Kg + t;
      v = __bfloat162float(s[0]) + __bfloat162float(s[plane]);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, mr = r0 + tx;
    if (t < K && mr < M2) tmp[((long long)f * K + t) * M2 + mr] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(256) hpx_ring_dft_inv_kernel(const float2* __restrict__ tmp, int C, int nside, int K, int M,
                                                              float* __restrict__ y) {
  __shared__ float4 sg[kMaxPhi / 2 + 1][kFT / 2];  // [m][field pair] = (re, im, re, im)
  __shared__ float2 tw[kMaxPhi];
  const int t = blockIdx.x, c0 = blockIdx.y * kFT;
  const Ring r = ring_of(t, nside);
  const long long npix = 12LL * nside * nside;
  const int nyq = r.nphi / 2;
  const int nm = min(nyq + 1, M);
  for (int i = threadIdx.x; i < r.nphi; i += blockDim.x) {
    double s, c;
    sincospi(2.0 * (double)i / (double)r.nphi, &s, &c);
    tw[i] = make_float2((float)c, (float)s);
  }
  float2* sg2 = reinterpret_cast<float2*>(sg);
  for (int i = threadIdx.x; i < kFT * nm; i += blockDim.x) {
    const int f = i / nm, m = i - f * nm;
    float2 v = make_float2(0.f, 0.f);
    if (c0 + f < C) {
      const float2 gc = tmp[((long long)(c0 + f) * K + t) * M + m];
      double ps, pc;
      sincos((double)m * r.phi0, &ps, &pc);
      v.x = gc.x * (float)pc - gc.y * (float)ps;
      v.y = gc.x * (float)ps + gc.y * (float)pc;
      // Hermitian weights of the c2r transform
      const float wgt = (m == 0 || m == nyq) ? 1.f : 2.f;
      v.x *= wgt;
      v.y = (m == 0 || m == nyq) ? 0.f : v.y * wgt;
    }
    sg2[m * kFT + f] = v;
  }
  __syncthreads();
  // real output: y[j] and y[nphi - j] share sum_m vr cos (A) and differ in the sign of sum_m vi sin (S): a thread owns the pair
  for (int j = threadIdx.x; j <= nyq; j += blockDim.x) {
    float A[kFT], S[kFT];
#pragma unroll
    for (int f = 0; f < kFT; ++f) A[f] = S[f] = 0.f;
    int q = 0;  // (j * m) mod nphi
#pragma unroll 2
    for (int m = 0; m < nm; ++m) {
      const float2 w = tw[q];
#pragma unroll
      for (int h = 0; h < kFT / 2; ++h) {
        const float4 v = sg[m][h];
        A[2 * h] = fmaf(v.x, w.x, A[2 * h]);
        S[2 * h] = fmaf(v.y, w.y, S[2 * h]);
        A[2 * h + 1] = fmaf(v.z, w.x, A[2 * h + 1]);
        S[2 * h + 1] = fmaf(v.w, w.y, S[2 * h + 1]);
      }
      q += j;
      if (q >= r.nphi) q -= r.nphi;
    }
    const bool mirror = j > 0 && j < nyq;
#pragma unroll
    for (int f = 0; f < kFT; ++f)
      if (c0 + f < C) {
        float* yr = y + (long long)(c0 + f) * npix + r.start;
        yr[j] = A[f] - S[f];
        if (mirror) yr[r.nphi - j] = A[f] + S[f];
      }
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
  p.ws_hpx.ensure((size_t)C * p.K * p.M * sizeof(float2));
  {
    ProfileScope prof("hpx.ring_dft_fwd", s);
    dim3 grid(p.K, (C + kFT - 1) / kFT);
    hpx_ring_dft_fwd_kernel<<<grid, 128, 0, s>>>(x_dev, C, nside, p.K, p.M, p.ws_hpx.as<float2>());
    after_launch("hpx_ring_dft_fwd");
  }
  {
    ProfileScope prof("hpx.tmp_to_x1", s);
    dim3 grid(C, (2 * p.M + 31) / 32, (p.K + 31) / 32);
    hpx_tmp_to_x1_kernel<<<grid, 256, 0, s>>>(p.ws_hpx.as<float>(), C, p.K, 2 * p.M, p.Kp, p.ws_x1.as<bf16>(), x1p);
    after_launch("hpx_tmp_to_x1");
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
  p.ws_g2.ensure(2 * (size_t)p.g2_elems(C) * sizeof(bf16));
  const long long cp = (long long)(p.ws_c2.bytes / sizeof(bf16) / 2), gp = (long long)(p.ws_g2.bytes / sizeof(bf16) / 2);
  launch_spec_complex_to_planes(coeffs_dev, C, p.L, p.M, p.Lp, p.ws_c2.as<bf16>(), cp, s);
  run_gemm(sht_op_legendre_inv2(p, p.ws_c2.as<bf16>(), cp, C, 1, p.ws_g2.as<bf16>(), gp), s);
  p.ws_hpx.ensure((size_t)C * p.K * p.M * sizeof(float2));
  {
    ProfileScope prof("hpx.g_to_tmp", s);
    dim3 grid(C, (2 * p.M + 31) / 32, (p.K + 31) / 32);
    hpx_g_to_tmp_kernel<<<grid, 256, 0, s>>>(p.ws_g2.as<bf16>(), gp, C, p.K, 2 * p.M, p.Kg, p.Ke, p.ws_hpx.as<float>());
    after_launch("hpx_g_to_tmp");
  }
  {
    ProfileScope prof("hpx.ring_dft_inv", s);
    dim3 grid(p.K, (C + kFT - 1) / kFT);
    hpx_ring_dft_inv_kernel<<<grid, 160, 0, s>>>  // nphi/2 + 1 <= 129 pixel pairs per equatorial ring at nside 64
       (p.ws_hpx.as<float2>(), C, nside, p.K, p.M, x_dev);
    after_launch("hpx_ring_dft_inv");
  }
  ACE_API_END
}
