// HEALPix spherical harmonic transform: per-ring real DFT + phase shift, then the SAME
// Legendre contraction as the lat-lon transform (tcgen05 GEMMs of sht.cu) over the 4*nside - 1 iso-latitude rings.
//
// Replaces /root/reference/fme/core/cuhpx/sht.py:32-153 (SHT / iSHT) with tools.py:34-83 (healpix_rfft_torch /
// healpix_irfft_torch: a Python loop of 4*nside - 1 torch.fft calls) by a ring-DFT kernel + a tile transposition per direction.
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
//   hpx_ring_dft_fwd_kernel   x [C][npix] fp32 -> tmp [C][K rings][M] complex fp32.  A block is ONE warp that owns one ring and
//       kFT = 8 fields.  The ring's samples sit in shared memory as [j][8 fields], folded for the real-input symmetry (slot j keeps
//       e = x[j] + x[n-j], slot n - j keeps o = x[j] - x[n-j]: Re = sum_j e cos, Im = -sum_j o sin, half the multiplications).
//       A lane owns kMT = 4 orders m = lane + 32 i and keeps their 8 fields' (re, im) sums in registers: 64 FMAs per j against
//       64 bytes of samples and ONE per-lane twiddle read -- the other three twiddles are that one times the warp-uniform
//       e^{-2 pi i 32 j / n}.  (The shared-memory return path moves 128 B/clk per SM whether or not the lanes read the same
//       address: the version with one order per lane needed 4.5 B per FMA and ran at a fifth of the FMA rate.)
//   hpx_tmp_to_x1_kernel      tmp -> X1 planes [(2m + reim)][C][Kp] (ring index contiguous, the Legendre GEMM's A operand):
//       32 x 32 tiles transposed through shared memory, so the 2-byte plane stores come in 64-byte runs instead of one sector per
//       element (a block of the ring kernel owns ONE ring, i.e. one column of X1).
// grid (ceil(C / kFT), rings), 32 threads, dynamic shared memory hpx_fwd_smem(nside).
constexpr int kMT = 4;  // orders (forward) / pixel pairs (inverse) per lane at most
__host__ __device__ inline size_t hpx_fwd_smem(int nside) { return (size_t)4 * nside * (kFT * sizeof(float) + sizeof(float2)); }
__host__ __device__ inline size_t hpx_inv_smem(int nside) { return (size_t)(2 * nside + 1) * kFT * sizeof(float2) + (size_t)4 * nside * sizeof(float2); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// packed fp32 pairs (fma.rn.f32x2, SASS FFMA2): one issue slot per two FMAs -- the loops below are issue-bound otherwise
struct f2 {
  unsigned long long u;
};
__device__ __forceinline__ f2 mk2(float a, float b) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.u) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void un2(f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v.u)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.u) : "l"(a.u), "l"(b.u), "l"(c.u));
  return d;
}

// the orders m = m0 + lane + 32 i, i < MT, of one ring and 8 fields (MT by ring length: short polar rings have few orders)
template <int MT>
__device__ __forceinline__ void hpx_fwd_orders(const float4* sx, const float2* tw, int n, int nm, int m0, int lane, int M, int C, int c0, int K, int t,
                                               double phi0, float2* __restrict__ tmp) {
  const int half = n / 2;
  f2 re[MT][kFT / 2], im[MT][kFT / 2];
  {  // j = 0 and j = n/2: cosine 1 and (-1)^m, no sine term
    const float4 a = sx[0], b = sx[1], c = sx[half * 2], d = sx[half * 2 + 1];
    const f2 v0[kFT / 2] = {mk2(a.x, a.y), mk2(a.z, a.w), mk2(b.x, b.y), mk2(b.z, b.w)};
    const f2 vh[kFT / 2] = {mk2(c.x, c.y), mk2(c.z, c.w), mk2(d.x, d.y), mk2(d.z, d.w)};
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const float sg = ((m0 + lane + 32 * i) & 1) ? -1.f : 1.f;
      const f2 sg2 = mk2(sg, sg);
#pragma unroll
      for (int h = 0; h < kFT / 2; ++h) {
        re[i][h] = fma2(vh[h], sg2, v0[h]);
        im[i][h] = mk2(0.f, 0.f);
      }
    }
  }
  const int mstep = (m0 + lane) % n, ustep = 32 % n;
  int q = mstep, qu = ustep;  // (j * m) mod n for this lane's first order; (j * 32) mod n
  for (int j = 1; j < half; ++j) {
    float2 w[MT];
    w[0] = tw[q];
    if (MT > 1) {
      const float2 u = tw[qu];
#pragma unroll
      for (int i = 1; i < MT; ++i) w[i] = cmul(w[i - 1], u);
    }
    const float4 a = sx[2 * j], b = sx[2 * j + 1], c = sx[2 * (n - j)], d = sx[2 * (n - j) + 1];
    const f2 e[kFT / 2] = {mk2(a.x, a.y), mk2(a.z, a.w), mk2(b.x, b.y), mk2(b.z, b.w)};
    const f2 o[kFT / 2] = {mk2(c.x, c.y), mk2(c.z, c.w), mk2(d.x, d.y), mk2(d.z, d.w)};
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const f2 wc = mk2(w[i].x, w[i].x), ws = mk2(-w[i].y, -w[i].y);
#pragma unroll
      for (int h = 0; h < kFT / 2; ++h) {
        re[i][h] = fma2(e[h], wc, re[i][h]);
        im[i][h] = fma2(o[h], ws, im[i][h]);
      }
    }
    q += mstep;
    if (q >= n) q -= n;
    qu += ustep;
    if (qu >= n) qu -= n;
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int m = m0 + lane + 32 * i;
    if (m >= M) continue;
    float fc = 0.f, fs = 0.f;
    const bool on = m < nm;
    if (on) {
      double ps, pc;
      sincos(-(double)m * phi0, &ps, &pc);
      fc = (float)pc;
      fs = (float)ps;
    }
#pragma unroll
    for (int h = 0; h < kFT / 2; ++h) {
      float r0, r1, i0, i1;
      un2(re[i][h], r0, r1);
      un2(im[i][h], i0, i1);
      const int f = 2 * h;
      if (c0 + f < C) tmp[((long long)(c0 + f) * K + t) * M + m] = on ? make_float2(r0 * fc - i0 * fs, r0 * fs + i0 * fc) : make_float2(0.f, 0.f);
      if (c0 + f + 1 < C)
        tmp[((long long)(c0 + f + 1) * K + t) * M + m] = on ? make_float2(r1 * fc - i1 * fs, r1 * fs + i1 * fc) : make_float2(0.f, 0.f);
    }
  }
}

__global__ void __launch_bounds__(32) hpx_ring_dft_fwd_kernel(const float* __restrict__ x, int C, int nside, int K, int M,
                                                             float2* __restrict__ tmp) {
  extern __shared__ float4 hpx_smem[];
  const int cap = 4 * nside;
  float4* sx = hpx_smem;                                     // [cap][kFT / 4]
  float2* tw = reinterpret_cast<float2*>(hpx_smem + cap * (kFT / 4));  // [cap]
  const int t = blockIdx.y, c0 = blockIdx.x * kFT, lane = threadIdx.x;
  const Ring r = ring_of(t, nside);
  const int n = r.nphi, half = n / 2;  // n is a multiple of 4
  const long long npix = 12LL * nside * nside;
  for (int i = lane; i < n; i += 32) {
    float sn, cs;
    sincospif(2.f * (float)i / (float)n, &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
  float* sxf = reinterpret_cast<float*>(sx);
  for (int f = 0; f < kFT; ++f) {
    const bool live = c0 + f < C;
    const float* xr = x + (long long)(c0 + f) * npix + r.start;
    for (int j = lane; j < n; j += 32) sxf[j * kFT + f] = live ? xr[j] : 0.f;
  }
  __syncwarp();
  for (int i = lane; i < kFT * (half - 1); i += 32) {
    const int f = i % kFT, j = 1 + i / kFT;
    const float a = sxf[j * kFT + f], b = sxf[(n - j) * kFT + f];
    sxf[j * kFT + f] = a + b;
    sxf[(n - j) * kFT + f] = a - b;
  }
  __syncwarp();
  const int nm = min(half + 1, M);
  int m0 = 0;
  // warp-uniform dispatch: as many orders per lane as the ring has left
  for (; m0 < nm; ) {
    const int left = nm - m0;
    if (left > 64) {
      hpx_fwd_orders<4>(sx, tw, n, nm, m0, lane, M, C, c0, K, t, r.phi0, tmp);
      m0 += 128;
    } else if (left > 32) {
      hpx_fwd_orders<2>(sx, tw, n, nm, m0, lane, M, C, c0, K, t, r.phi0, tmp);
      m0 += 64;
    } else {
      hpx_fwd_orders<1>(sx, tw, n, nm, m0, lane, M, C, c0, K, t, r.phi0, tmp);
      m0 += 32;
    }
  }
  // orders the ring does not resolve: zero
  for (int m = m0 + lane; m < M; m += 32)
    for (int f = 0; f < kFT; ++f)
      if (c0 + f < C) tmp[((long long)(c0 + f) * K + t) * M + m] = make_float2(0.f, 0.f);
}

// tmp [C][K][2M] fp32 (2M = interleaved re / im of the orders) -> X1 planes [(2m + reim)][C][Kp]; 64 (rings) x 32 (rows) tiles, two
// rings per lane on the store side (bf16x2: 128 bytes per warp and plane row; Kp is even, the pad column K of an odd ring count
// receives zero).  grid (C, ceil(2M/32), ceil(K/64))
__global__ void __launch_bounds__(256) hpx_tmp_to_x1_kernel(const float* __restrict__ tmp, int C, int K, int M2, int Kp,
                                                           bf16* __restrict__ x1, long long plane) {
  __shared__ float tile[64][33];
  const int t0 = blockIdx.z * 64, r0 = blockIdx.y * 32, f = blockIdx.x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 64; i += 8) {
    const int t = t0 + i, mr = r0 + tx;
    tile[i][tx] = (t < K && mr < M2) ? tmp[((long long)f * K + t) * M2 + mr] : 0.f;
  }
  __syncthreads();
  const int t = t0 + 2 * tx;
  if (t >= Kp) return;
  for (int i = ty; i < 32; i += 8) {
    const int mr = r0 + i;
    if (mr >= M2) break;
    const float v0 = tile[2 * tx][i], v1 = tile[2 * tx + 1][i];
    const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v0 - __low2float(hi), v1 - __high2float(hi));
    bf16* d = x1 + ((long long)mr * C + f) * Kp + t;
    *reinterpret_cast<__nv_bfloat162*>(d) = hi;
    *reinterpret_cast<__nv_bfloat162*>(d + plane) = lo;
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
  __shared__ float tile[32][65];
  const int t0 = blockIdx.z * 64, r0 = blockIdx.y * 32, f = blockIdx.x;  // 64 rings x 32 rows per block; Kg is a multiple of 64
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int mr = r0 + i;
    float v0 = 0.f, v1 = 0.f;
    if (mr < M2) {
      const int row = (((mr >> 1) & 1) ? Ke : 0) + 2 * (mr >> 2) + (mr & 1);  // order m = mr / 2, reim = mr & 1
      const bf16* s = g + ((long long)row * C + f) * Kg + t0 + 2 * tx;        // two rings per lane (bf16x2)
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(s), l = *reinterpret_cast<const __nv_bfloat162*>(s + plane);
      v0 = __low2float(h) + __low2float(l);
      v1 = __high2float(h) + __high2float(l);
    }
    tile[i][2 * tx] = v0;
    tile[i][2 * tx + 1] = v1;
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int t = t0 + i, mr = r0 + tx;
    if (t < K && mr < M2) tmp[((long long)f * K + t) * M2 + mr] = tile[tx][i];
  }
}

// One warp per (ring, 8 fields): a lane owns up to kMT = 4 pixel pairs (j, n - j), j = j0 + lane + 32 i, and their 8 fields' cosine /
// sine sums (64 FMAs, issued as 32 FFMA2, per order m against 64 bytes of coefficients and one per-lane twiddle read, like the
// forward kernel); the Nyquist pixel j = n/2 (cosine (-1)^m, no sine) is a warp reduction.
// grid (ceil(C / kFT), rings), 32 threads, hpx_inv_smem(nside) bytes.
template <int MT>
__device__ __forceinline__ void hpx_inv_pixels(const float4* sg, const float2* tw, int n, int nm, int j0, int lane, int C, int c0, float* __restrict__ yr0,
                                               long long npix) {
  const int nyq = n / 2;
  f2 A[MT][kFT / 2], S[MT][kFT / 2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int h = 0; h < kFT / 2; ++h) A[i][h] = S[i][h] = mk2(0.f, 0.f);
  const int jstep = (j0 + lane) % n, ustep = 32 % n;
  int q = 0, qu = 0;  // (m * j) mod n for this lane's first pixel; (m * 32) mod n
  for (int m = 0; m < nm; ++m) {
    float2 w[MT];
    w[0] = tw[q];
    if (MT > 1) {
      const float2 u = tw[qu];
#pragma unroll
      for (int i = 1; i < MT; ++i) w[i] = cmul(w[i - 1], u);
    }
    const float4 a = sg[4 * m], b = sg[4 * m + 1], c = sg[4 * m + 2], d = sg[4 * m + 3];  // row m: 8 real parts, then 8 imaginary parts
    const f2 vr[kFT / 2] = {mk2(a.x, a.y), mk2(a.z, a.w), mk2(b.x, b.y), mk2(b.z, b.w)};
    const f2 vi[kFT / 2] = {mk2(c.x, c.y), mk2(c.z, c.w), mk2(d.x, d.y), mk2(d.z, d.w)};
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const f2 wc = mk2(w[i].x, w[i].x), ws = mk2(w[i].y, w[i].y);
#pragma unroll
      for (int h = 0; h < kFT / 2; ++h) {
        A[i][h] = fma2(vr[h], wc, A[i][h]);
        S[i][h] = fma2(vi[h], ws, S[i][h]);
      }
    }
    q += jstep;
    if (q >= n) q -= n;
    qu += ustep;
    if (qu >= n) qu -= n;
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int j = j0 + lane + 32 * i;
    if (j >= nyq) continue;
#pragma unroll
    for (int h = 0; h < kFT / 2; ++h) {
      float a0, a1, s0, s1;
      un2(A[i][h], a0, a1);
      un2(S[i][h], s0, s1);
      const int f = 2 * h;
      if (c0 + f < C) {
        float* yr = yr0 + (long long)f * npix;
        yr[j] = a0 - s0;
        if (j > 0) yr[n - j] = a0 + s0;
      }
      if (c0 + f + 1 < C) {
        float* yr = yr0 + (long long)(f + 1) * npix;
        yr[j] = a1 - s1;
        if (j > 0) yr[n - j] = a1 + s1;
      }
    }
  }
}

__global__ void __launch_bounds__(32) hpx_ring_dft_inv_kernel(const float2* __restrict__ tmp, int C, int nside, int K, int M,
                                                             float* __restrict__ y) {
  extern __shared__ float4 hpx_smem[];
  float4* sg = hpx_smem;                                                        // [2 nside + 1][4] = 8 real parts, 8 imaginary parts
  float2* tw = reinterpret_cast<float2*>(hpx_smem + (2 * nside + 1) * (kFT / 2));  // [4 nside]
  const int t = blockIdx.y, c0 = blockIdx.x * kFT, lane = threadIdx.x;
  const Ring r = ring_of(t, nside);
  const long long npix = 12LL * nside * nside;
  const int n = r.nphi, nyq = n / 2;
  const int nm = min(nyq + 1, M);
  for (int i = lane; i < n; i += 32) {
    float sn, cs;
    sincospif(2.f * (float)i / (float)n, &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
  float* sgf = reinterpret_cast<float*>(sg);
  for (int m = lane; m < nm; m += 32) {
    double ps, pc;
    sincos((double)m * r.phi0, &ps, &pc);
    const float fc = (float)pc, fs = (float)ps;
    const bool edge = (m == 0 || m == nyq);
    // Hermitian weights of the c2r transform; the imaginary parts of m = 0 and of the Nyquist mode are ignored
    const float wgt = edge ? 1.f : 2.f;
    for (int f = 0; f < kFT; ++f) {
      float2 v = make_float2(0.f, 0.f);
      if (c0 + f < C) {
        const float2 gc = tmp[((long long)(c0 + f) * K + t) * M + m];
        v.x = (gc.x * fc - gc.y * fs) * wgt;
        v.y = edge ? 0.f : (gc.x * fs + gc.y * fc) * wgt;
      }
      sgf[m * 2 * kFT + f] = v.x;
      sgf[m * 2 * kFT + kFT + f] = v.y;
    }
  }
  __syncwarp();
  float* yr0 = y + (long long)c0 * npix + r.start;
  for (int j0 = 0; j0 < nyq;) {  // warp-uniform dispatch: as many pixel pairs per lane as the ring has left
    const int left = nyq - j0;
    if (left > 64) {
      hpx_inv_pixels<4>(sg, tw, n, nm, j0, lane, C, c0, yr0, npix);
      j0 += 128;
    } else if (left > 32) {
      hpx_inv_pixels<2>(sg, tw, n, nm, j0, lane, C, c0, yr0, npix);
      j0 += 64;
    } else {
      hpx_inv_pixels<1>(sg, tw, n, nm, j0, lane, C, c0, yr0, npix);
      j0 += 32;
    }
  }
  {  // Nyquist pixel j = n/2: y = sum_m vr[m] (-1)^m
    float acc[kFT];
#pragma unroll
    for (int f = 0; f < kFT; ++f) acc[f] = 0.f;
    for (int m = lane; m < nm; m += 32) {
      const float sgn = (m & 1) ? -1.f : 1.f;
#pragma unroll
      for (int f = 0; f < kFT; ++f) acc[f] = fmaf(sgf[m * 2 * kFT + f], sgn, acc[f]);
    }
#pragma unroll
    for (int f = 0; f < kFT; ++f) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], d);
      if (lane == 0 && c0 + f < C) yr0[(long long)f * npix + nyq] = acc[f];
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
    dim3 grid((C + kFT - 1) / kFT, p.K);  // field groups on x (no 65 535 limit), rings on y
    hpx_ring_dft_fwd_kernel<<<grid, 32, hpx_fwd_smem(nside), s>>>(x_dev, C, nside, p.K, p.M, p.ws_hpx.as<float2>());
    after_launch("hpx_ring_dft_fwd");
  }
  {
    ProfileScope prof("hpx.tmp_to_x1", s);
    dim3 grid(C, (2 * p.M + 31) / 32, (p.K + 63) / 64);
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
    dim3 grid(C, (2 * p.M + 31) / 32, (p.K + 63) / 64);
    hpx_g_to_tmp_kernel<<<grid, 256, 0, s>>>(p.ws_g2.as<bf16>(), gp, C, p.K, 2 * p.M, p.Kg, p.Ke, p.ws_hpx.as<float>());
    after_launch("hpx_g_to_tmp");
  }
  {
    ProfileScope prof("hpx.ring_dft_inv", s);
    dim3 grid((C + kFT - 1) / kFT, p.K);  // field groups on x (no 65 535 limit), rings on y
    hpx_ring_dft_inv_kernel<<<grid, 32, hpx_inv_smem(nside), s>>>(p.ws_hpx.as<float2>(), C, nside, p.K, p.M, x_dev);
    after_launch("hpx_ring_dft_inv");
  }
  ACE_API_END
}
