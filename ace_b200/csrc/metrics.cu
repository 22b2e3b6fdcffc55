// Device reductions of the inference aggregators: area-weighted horizontal moments,
// zonal means and the spherical power spectrum.  HBM-bound streaming kernels; fp64 accumulation.
//
// Reference arithmetic (/root/reference):
//   fme/core/metrics.py:35-90    weighted_sum / weighted_mean  (zero-weight points contribute 0, even if NaN)
//   fme/core/metrics.py:118-197  weighted_std, weighted_mean_bias, root_mean_squared_error
//   fme/core/gridded_ops.py:284-360 LatLonOperations (area weights over the last two dims), zonal mean = mean over lon
//   fme/core/metrics.py:388-408  spherical_power_spectrum = sum_m |c_lm|^2
#include "common.cuh"

namespace ace {
namespace {

constexpr int kT = 256;

// out[f][0..4] += { sum w x, sum w x^2, sum w (x - t), sum w (x - t)^2, sum w }   (t optional)
__global__ void __launch_bounds__(kT) weighted_moments_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                             const float* __restrict__ w, long long hw,
                                                             double* __restrict__ out) {
  const long long f = blockIdx.y;
  const float* xf = x + f * hw;
  const float* tf = t ? t + f * hw : nullptr;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < hw; i += (long long)gridDim.x * kT) {
    const float wi = __ldg(w + i);
    if (wi != 0.f) {  // "expected NaNs" under zero weight are dropped (metrics.py:56-58)
      const double wd = wi, xv = xf[i];
      a0 += wd * xv;
      a1 += wd * xv * xv;
      if (tf) {
        const double d = xv - (double)tf[i];
        a2 += wd * d;
        a3 += wd * d * d;
      }
      a4 += wd;
    }
  }
  __shared__ double red[5][kT / 32];
  double v[5] = {a0, a1, a2, a3, a4};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0;
    for (int i = 0; i < kT / 32; ++i) s += red[threadIdx.x][i];
    atomicAdd(out + f * 5 + threadIdx.x, s);
  }
}

// out[f][h] = mean over w of x[f][h][w]; one warp per latitude row
__global__ void __launch_bounds__(kT) zonal_mean_kernel(const float* __restrict__ x, long long rows, int W,
                                                       float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* r = x + row * W;
  double s = 0;
  for (int j = threadIdx.x & 31; j < W; j += 32) s += (double)r[j];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) out[row] = (float)(s / (double)W);
}

// out[f][l] = sum_m |c[f][l][m]|^2 ; one warp per (f, l)
__global__ void __launch_bounds__(kT) power_spectrum_kernel(const float2* __restrict__ c, long long rows, int M,
                                                           float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float2* r = c + row * M;
  double s = 0;
  for (int m = threadIdx.x & 31; m < M; m += 32) {
    const float2 v = r[m];
    s += (double)v.x * v.x + (double)v.y * v.y;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) out[row] = (float)s;
}

// acc[i] += sum over (a, b) of x[a * stride_a + b * stride_b + i]: the window's contribution to the time-mean maps
// (fme/ace/aggregator/inference/time_mean.py:103-124: tensor[:, time_slice].sum(dim=time).sum(dim=sample) added to a running
// fp32 map).  The window sum is formed in fp64 and rounded once before it joins the fp32 running sum.
__global__ void __launch_bounds__(kT) time_sum_kernel(const float* __restrict__ x, int na, long long stride_a, int nb,
                                                     long long stride_b, long long n, float* __restrict__ acc) {
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n; i += (long long)gridDim.x * kT) {
    double s = 0.0;
    for (int a = 0; a < na; ++a) {
      const float* xa = x + a * stride_a + i;
      double sa = 0.0;
      for (int b = 0; b < nb; ++b) sa += (double)__ldg(xa + b * stride_b);
      s += sa;
    }
    acc[i] += (float)s;
  }
}

}  // namespace
}  // namespace ace

using namespace ace;

extern "C" int ace_time_sum(const float* x_dev, int n_outer, long long stride_outer, int n_inner, long long stride_inner,
                            long long n_elems, float* acc_dev, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(x_dev && acc_dev && n_outer > 0 && n_inner > 0 && n_elems > 0, "ace_time_sum: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  ProfileScope prof("time_sum", s);
  const int grid = (int)std::min<long long>((n_elems + kT - 1) / kT, 148 * 8);
  time_sum_kernel<<<grid, kT, 0, s>>>(x_dev, n_outer, stride_outer, n_inner, stride_inner, n_elems, acc_dev);
  after_launch("time_sum");
  ACE_API_END
}

extern "C" int ace_weighted_moments(const float* x_dev, const float* t_dev, const float* weights_dev, long long nfields,
                                    long long hw, double* out_dev, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(x_dev && weights_dev && out_dev, "ace_weighted_moments: null argument");
  ACE_REQUIRE(nfields > 0 && nfields <= 65535 && hw > 0, "ace_weighted_moments: bad sizes (%lld fields of %lld points)", nfields, hw);
  cudaStream_t s = (cudaStream_t)stream;
  ACE_CHECK_CUDA(cudaMemsetAsync(out_dev, 0, (size_t)nfields * 5 * sizeof(double), s));
  long long per = (hw + kT - 1) / kT;
  int gx = (int)std::min<long long>(per, std::max<long long>(1, (148LL * 8 + nfields - 1) / nfields));
  ProfileScope prof("weighted_moments", s);
  weighted_moments_kernel<<<dim3(gx, (unsigned)nfields), kT, 0, s>>>(x_dev, t_dev, weights_dev, hw, out_dev);
  after_launch("weighted_moments");
  ACE_API_END
}

extern "C" int ace_zonal_mean(const float* x_dev, long long nfields, int h, int w, float* out_dev, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(x_dev && out_dev && nfields > 0 && h > 0 && w > 0, "ace_zonal_mean: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const long long rows = nfields * h;
  ProfileScope prof("zonal_mean", s);
  zonal_mean_kernel<<<(unsigned)((rows + kT / 32 - 1) / (kT / 32)), kT, 0, s>>>(x_dev, rows, w, out_dev);
  after_launch("zonal_mean");
  ACE_API_END
}

extern "C" int ace_power_spectrum(const float* coeffs_dev, long long nfields, int lmax, int mmax, float* out_dev, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(coeffs_dev && out_dev && nfields > 0 && lmax > 0 && mmax > 0, "ace_power_spectrum: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const long long rows = nfields * lmax;
  ProfileScope prof("power_spectrum", s);
  power_spectrum_kernel<<<(unsigned)((rows + kT / 32 - 1) / (kT / 32)), kT, 0, s>>>(reinterpret_cast<const float2*>(coeffs_dev), rows, mmax,
                                                                                 out_dev);
  after_launch("power_spectrum");
  ACE_API_END
}
