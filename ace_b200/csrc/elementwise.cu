// HBM-bound streaming kernels around the GEMMs: InstanceNorm apply + bf16 split, weight
// re-layout, spectral layout conversions, the stepper's pack/normalise/denormalise.
#include "kernels.cuh"

namespace ace {

namespace {

constexpr int kThreads = 256;

inline int grid_for(long long work_items, int per_block, int max_blocks = 148 * 16) {
  long long b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------
// norm + split: one (b, c) row per blockIdx.y, grid-stride over HW with 4-wide vectors
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) norm_split_kernel(const float* __restrict__ src, int C, long long HW,
                                                             const double* __restrict__ stats,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             bf16* __restrict__ dst, long long plane, long long dst_b,
                                                             long long dst_c) {
  const int bc = blockIdx.y;
  const int b = bc / C, c = bc % C;
  float alpha = 1.f, shift = 0.f;
  if (stats != nullptr) {
    double s = stats[(long long)bc * 2], q = stats[(long long)bc * 2 + 1];
    double mean = s / (double)HW;
    double var = q / (double)HW - mean * mean;
    if (var < 0.0) var = 0.0;
    double invstd = 1.0 / sqrt(var + (double)eps);
    double a = invstd * (double)gamma[c];
    alpha = (float)a;
    shift = (float)((double)beta[c] - mean * a);
  }
  const float* s = src + (long long)bc * HW;
  bf16* dh = dst + (long long)b * dst_b + (long long)c * dst_c;
  bf16* dl = dh + plane;
  const bool vec = ((HW & 3) == 0) && ((((uintptr_t)s) & 15) == 0) && ((((uintptr_t)dh) & 7) == 0) &&
                   ((((uintptr_t)dl) & 7) == 0);
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const long long n4 = HW >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 v = __ldg(reinterpret_cast<const float4*>(s) + i);
      float f[4] = {fmaf(v.x, alpha, shift), fmaf(v.y, alpha, shift), fmaf(v.z, alpha, shift), fmaf(v.w, alpha, shift)};
      __align__(8) bf16 h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_bf16(f[j], h[j], l[j]);
      *reinterpret_cast<uint2*>(dh + 4 * i) = *reinterpret_cast<uint2*>(h);
      *reinterpret_cast<uint2*>(dl + 4 * i) = *reinterpret_cast<uint2*>(l);
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += stride) {
      bf16 h, l;
      split_bf16(fmaf(s[i], alpha, shift), h, l);
      dh[i] = h;
      dl[i] = l;
    }
  }
}

__global__ void __launch_bounds__(kThreads) split_pad_kernel(const float* __restrict__ src, long long rows, int cols,
                                                            int cols_pad, bf16* __restrict__ dst, long long plane) {
  const long long total = rows * cols_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cols_pad;
    int c = (int)(i % cols_pad);
    float v = (c < cols) ? src[r * cols + c] : 0.f;
    bf16 h, l;
    split_bf16(v, h, l);
    dst[i] = h;
    dst[i + plane] = l;
  }
}

// dhconv weight [Cin][Cout][L][2] (reference layout) -> planes [L][2 (re, im)][Cout][Cinp] for the complex GEMM mode
__global__ void __launch_bounds__(kThreads) prep_dhconv_cplx_kernel(const float* __restrict__ w, int Cin, int Cout, int L,
                                                                   long long s_l, long long s_o, long long s_i, int Cinp,
                                                                   bf16* __restrict__ dst, long long plane) {
  const long long total = (long long)L * 2 * Cout * Cinp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int i = (int)(idx % Cinp);
    long long t = idx / Cinp;
    int o = (int)(t % Cout);
    t /= Cout;
    int part = (int)(t & 1);
    int l = (int)(t >> 1);
    float v = (i < Cin) ? w[(l * s_l + o * s_o + i * s_i) * 2 + part] : 0.f;
    bf16 h, lo;
    split_bf16(v, h, lo);
    dst[idx] = h;
    dst[idx + plane] = lo;
  }
}

// grouped dhconv weights [G][L][cg (o')][cg (i')][2] fp32 -> planes [L][2][C = G cg][cgp]: row o = g cg + o' keeps its own group's inputs
__global__ void __launch_bounds__(kThreads) prep_dhconv_grouped_kernel(const float* __restrict__ w, int G, int cg, int L, int cgp,
                                                                      bf16* __restrict__ dst, long long plane) {
  const int C = G * cg;
  const long long total = (long long)L * 2 * C * cgp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int i = (int)(idx % cgp);
    long long t = idx / cgp;
    int o = (int)(t % C);
    t /= C;
    int part = (int)(t & 1);
    int l = (int)(t >> 1);
    const int g = o / cg, o2 = o - g * cg;
    float v = (i < cg) ? w[(((((long long)g * L + l) * cg + o2) * cg) + i) * 2 + part] : 0.f;
    bf16 h, lo;
    split_bf16(v, h, lo);
    dst[idx] = h;
    dst[idx + plane] = lo;
  }
}

// one thread per (b, l, m, o); x is broadcast across the o threads of a warp
__global__ void __launch_bounds__(kThreads) diagonal_contract_kernel(const bf16* __restrict__ c1, long long c1_plane,
                                                                    const float* __restrict__ w, int B, int C, int L,
                                                                    int M, int Lp, bf16* __restrict__ c2,
                                                                    long long c2_plane) {
  const long long total = (long long)B * L * M * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int o = (int)(idx % C);
    long long t = idx / C;
    int m = (int)(t % M);
    t /= M;
    int l = (int)(t % L);
    int b = (int)(t / L);
    const bf16* x = c1 + (((long long)b * L + l) * M + m) * 2 * C;
    float yr = 0.f, yi = 0.f;
    for (int i = 0; i < C; ++i) {
      float xr = __bfloat162float(x[i]) + __bfloat162float(x[i + c1_plane]);
      float xi = __bfloat162float(x[C + i]) + __bfloat162float(x[C + i + c1_plane]);
      const float* p = w + ((((long long)i * C + o) * L + l) * M + m) * 2;
      float wr = p[0], wi = p[1];
      yr = fmaf(xr, wr, fmaf(-xi, wi, yr));
      yi = fmaf(xr, wi, fmaf(xi, wr, yi));
    }
    bf16* y = c2 + (((long long)b * M + m) * Lp + l) * 2 * C;
    bf16 h, lo;
    split_bf16(yr, h, lo);
    y[o] = h;
    y[o + c2_plane] = lo;
    split_bf16(yi, h, lo);
    y[C + o] = h;
    y[C + o + c2_plane] = lo;
  }
}

// Layout converters of the standalone transform API: 32 x 32 tiles transposed through shared memory so that both the plane
// side (channel contiguous) and the complex side (order m contiguous) are accessed in 128-256 byte runs.
// c1 planes [L*M][2C] -> out complex64 [C][L*M]
__global__ void __launch_bounds__(256) spec_planes_to_complex_kernel(const bf16* __restrict__ c1, long long plane, int C,
                                                                    long long LM, float* __restrict__ out) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long lm0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = ty; i < 32; i += 8) {
    const long long lm = lm0 + i;
    const int c = c0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (lm < LM && c < C) {
      const bf16* p = c1 + lm * 2 * C;
      v.x = __bfloat162float(p[c]) + __bfloat162float(p[c + plane]);
      v.y = __bfloat162float(p[C + c]) + __bfloat162float(p[C + c + plane]);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const long long lm = lm0 + tx;
    if (c < C && lm < LM) reinterpret_cast<float2*>(out)[(long long)c * LM + lm] = tile[tx][i];
  }
}

// in complex64 [C][L][M] -> c2 planes [M][Lp][2C]; grid (M tiles, C tiles, L)
__global__ void __launch_bounds__(256) spec_complex_to_planes_kernel(const float* __restrict__ in, int C, int L, int M, int Lp,
                                                                    bf16* __restrict__ c2, long long plane) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32, l = blockIdx.z;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, m = m0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (c < C && m < M) v = reinterpret_cast<const float2*>(in)[((long long)c * L + l) * M + m];
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int m = m0 + i, c = c0 + tx;
    if (m < M && c < C) {
      const float2 v = tile[tx][i];
      bf16* y = c2 + ((long long)m * Lp + l) * 2 * C;
      bf16 h, lo;
      split_bf16(v.x, h, lo);
      y[c] = h;
      y[c + plane] = lo;
      split_bf16(v.y, h, lo);
      y[C + c] = h;
      y[C + c + plane] = lo;
    }
  }
}

// The same converters for an even channel count (and even plane offsets): 64 channels x 32 spectral positions per block, a lane
// owns two adjacent channels on the plane side (bf16x2: 128 bytes per warp and row instead of 64)
__global__ void __launch_bounds__(256) spec_planes_to_complex2_kernel(const bf16* __restrict__ c1, long long plane, int C,
                                                                     long long LM, float* __restrict__ out) {
  __shared__ float2 tile[32][65];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long lm0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 64;
  const int c = c0 + 2 * tx;
  for (int i = ty; i < 32; i += 8) {
    const long long lm = lm0 + i;
    float2 v0 = make_float2(0.f, 0.f), v1 = v0;
    if (lm < LM && c < C) {
      const bf16* p = c1 + lm * 2 * C + c;
      const __nv_bfloat162 rh = *reinterpret_cast<const __nv_bfloat162*>(p), rl = *reinterpret_cast<const __nv_bfloat162*>(p + plane);
      const __nv_bfloat162 ih = *reinterpret_cast<const __nv_bfloat162*>(p + C), il = *reinterpret_cast<const __nv_bfloat162*>(p + C + plane);
      v0 = make_float2(__low2float(rh) + __low2float(rl), __low2float(ih) + __low2float(il));
      v1 = make_float2(__high2float(rh) + __high2float(rl), __high2float(ih) + __high2float(il));
    }
    tile[i][2 * tx] = v0;
    tile[i][2 * tx + 1] = v1;
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int cc = c0 + i;
    const long long lm = lm0 + tx;
    if (cc < C && lm < LM) reinterpret_cast<float2*>(out)[(long long)cc * LM + lm] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(256) spec_complex_to_planes2_kernel(const float* __restrict__ in, int C, int L, int M, int Lp,
                                                                     bf16* __restrict__ c2, long long plane) {
  __shared__ float2 tile[64][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 64, l = blockIdx.z;
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i, m = m0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (c < C && m < M) v = reinterpret_cast<const float2*>(in)[((long long)c * L + l) * M + m];
    tile[i][tx] = v;
  }
  __syncthreads();
  const int c = c0 + 2 * tx;
  if (c >= C) return;
  for (int i = ty; i < 32; i += 8) {
    const int m = m0 + i;
    if (m >= M) break;
    const float2 v0 = tile[2 * tx][i], v1 = tile[2 * tx + 1][i];
    bf16* y = c2 + ((long long)m * Lp + l) * 2 * C + c;
    const __nv_bfloat162 rh = __floats2bfloat162_rn(v0.x, v1.x), ih = __floats2bfloat162_rn(v0.y, v1.y);
    const __nv_bfloat162 rl = __floats2bfloat162_rn(v0.x - __low2float(rh), v1.x - __high2float(rh));
    const __nv_bfloat162 il = __floats2bfloat162_rn(v0.y - __low2float(ih), v1.y - __high2float(ih));
    *reinterpret_cast<__nv_bfloat162*>(y) = rh;
    *reinterpret_cast<__nv_bfloat162*>(y + plane) = rl;
    *reinterpret_cast<__nv_bfloat162*>(y + C) = ih;
    *reinterpret_cast<__nv_bfloat162*>(y + C + plane) = il;
  }
}

// Deferred InstanceNorm.  From the (sum, sum of squares) statistics of h[B][C][HW], per sample b:
//   a[c] = gamma[c] / sqrt(var_c + eps),   s[c] = beta[c] - mean_c * a[c]        (a*h + s is the normalised tensor)
// and fold them into the 1x1 convolution that consumes it:  W (a*h + s) + bias = (W diag(a)) h + (bias + W s):
//   wout planes [B][O][Ip] = split(W[o][i] * a[i]),   bout[B][O] = bias[o] + sum_i W[o][i] * s[i]
// Block (0, b) also publishes a, s and shift0 = 2*pi*s (the m = 0 DFT coefficient of the constant field s).
__global__ void __launch_bounds__(128) prep_norm_conv_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, long long HW, int C,
                                                            const float* __restrict__ w, const float* __restrict__ bias, int O,
                                                            int Ip, bf16* __restrict__ wout, long long wplane,
                                                            float* __restrict__ bout, float* __restrict__ a_out,
                                                            float* __restrict__ s_out, float* __restrict__ shift0_out) {
  // griddepcontrol: the statistics come from the previous kernel in the stream
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // one block per output row o (blockIdx.x < O) plus one extra block (blockIdx.x == O) that publishes a, s, 2*pi*s;
  // every thread derives the a_i, s_i it needs itself (4 channels per thread at C = 384), no shared staging
  __shared__ float red[4];
  const int b = blockIdx.y, o = blockIdx.x;
  const bool publish = (o == O) || (w == nullptr);
  const float* wr = publish ? nullptr : w + (long long)o * C;
  bf16* wo = publish ? nullptr : wout + ((long long)b * O + o) * Ip;
  float dot = 0.f;
  const double inv_hw = 1.0 / (double)HW;
  for (int i = threadIdx.x; i < Ip; i += blockDim.x) {
    float a = 0.f, sh = 0.f;
    if (i < C) {
      // the cancellation E[x^2] - E[x]^2 is done in fp64 (two multiplies and one FMA); the reciprocal square root runs on the
      // fp32 units (IEEE sqrt + division: relative error 1e-7, native_batch_norm's own accuracy) -- an fp64 rsqrt in every
      // thread of all O + 1 blocks was most of this kernel's 12 us
      const double sum = stats[((long long)b * C + i) * 2], sq = stats[((long long)b * C + i) * 2 + 1];
      const double mean = sum * inv_hw;
      double var = fma(-mean, mean, sq * inv_hw);
      if (var < 0.0) var = 0.0;
      a = __ldg(gamma + i) / sqrtf((float)var + eps);
      sh = (float)((double)__ldg(beta + i) - mean * (double)a);
    }
    if (publish) {
      if (i < C) {
        a_out[(long long)b * C + i] = a;
        s_out[(long long)b * C + i] = sh;
        shift0_out[(long long)b * C + i] = 6.283185307179586f * sh;
      }
    } else {
      float v = 0.f;
      if (i < C) {
        const float wi = __ldg(wr + i);
        v = wi * a;
        dot = fmaf(wi, sh, dot);
      }
      bf16 h, l;
      split_bf16(v, h, l);
      wo[i] = h;
      wo[i + wplane] = l;
    }
  }
  if (publish) return;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) bout[(long long)b * O + o] = (bias ? bias[o] : 0.f) + red[0] + red[1] + red[2] + red[3];
}

__global__ void vec_add_kernel(const float* a, const float* b, float* out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = a[i] + b[i];
}

__global__ void __launch_bounds__(kThreads) pack_normalize_kernel(const float* __restrict__ prog,
                                                                 const float* __restrict__ forcing, int n_prog,
                                                                 int n_forcing, const int* __restrict__ kind,
                                                                 const int* __restrict__ index,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ std, int n_in, long long HW,
                                                                 float* __restrict__ x) {
  const int bc = blockIdx.y;
  const int b = bc / n_in, c = bc % n_in;
  const float* s = (kind[c] == 0) ? prog + ((long long)b * n_prog + index[c]) * HW
                                  : forcing + ((long long)b * n_forcing + index[c]) * HW;
  float* d = x + (long long)bc * HW;
  const float mu = mean[c], sd = std[c];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x)
    d[i] = (s[i] - mu) / sd;
}

__global__ void __launch_bounds__(kThreads) unpack_denormalize_kernel(const float* __restrict__ y,
                                                                     const float* __restrict__ x_norm,
                                                                     const int* __restrict__ out_prog_index,
                                                                     const int* __restrict__ prog_in_chan,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ std, int residual,
                                                                     int n_out, int n_in, int n_prog, long long HW,
                                                                     const int* __restrict__ clamp, int ocean_out,
                                                                     int ocean_interp, const float* __restrict__ ocean,
                                                                     float* __restrict__ out,
                                                                     float* __restrict__ next_prog) {
  const int bc = blockIdx.y;
  const int b = bc / n_out, c = bc % n_out;
  const int p = out_prog_index[c];
  const float* s = y + (long long)bc * HW;
  const float* r = nullptr;
  if (residual && p >= 0 && prog_in_chan[p] >= 0) r = x_norm + ((long long)b * n_in + prog_in_chan[p]) * HW;
  float* d = out + (long long)bc * HW;
  float* np = (p >= 0 && next_prog != nullptr) ? next_prog + ((long long)b * n_prog + p) * HW : nullptr;
  const float mu = mean[c], sd = std[c];
  const bool pos = clamp[c] != 0;
  const float* om = (c == ocean_out) ? ocean + (long long)b * 2 * HW : nullptr;  // {mask, target}
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    float v = s[i];
    if (r) v += r[i];
    v = v * sd + mu;
    if (pos) v = (v < 0.f) ? 0.f : v;  // torch.clamp(min=0): NaN stays NaN
    if (om) {
      const float m = om[i], tgt = om[HW + i];
      if (ocean_interp) v = m * tgt + (1.f - m) * v;
      else if ((int)rintf(m) == 1) v = tgt;
    }
    d[i] = v;
    if (np) np[i] = v;
  }
}

__global__ void __launch_bounds__(kThreads) ocean_prescribe_kernel(float* __restrict__ out, float* __restrict__ next_prog,
                                                                  const int* __restrict__ out_prog_index, int n_out, int n_prog,
                                                                  long long HW, int ocean_out, int ocean_interp,
                                                                  const float* __restrict__ ocean) {
  const int b = blockIdx.y;
  float* d = out + ((long long)b * n_out + ocean_out) * HW;
  const int p = out_prog_index[ocean_out];
  float* np = (p >= 0 && next_prog != nullptr) ? next_prog + ((long long)b * n_prog + p) * HW : nullptr;
  const float* om = ocean + (long long)b * 2 * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    float v = d[i];
    const float m = om[i], tgt = om[HW + i];
    if (ocean_interp) v = m * tgt + (1.f - m) * v;
    else if ((int)rintf(m) == 1) v = tgt;
    d[i] = v;
    if (np) np[i] = v;
  }
}

// slab ocean (fme/core/ocean.py:64-88,223-243): target = T_in + (F_net + Q) / (rho * depth * c_p) * dt with F_net the net surface
// energy flux of the generated (corrected) fields without frozen precipitation (metrics.py:299-334), then the prescriber.
// ocean = [B][3][HW] {mask, q_flux, mixed layer depth}
__global__ void __launch_bounds__(kThreads) ocean_slab_kernel(float* __restrict__ out, float* __restrict__ next_prog,
                                                             const float* __restrict__ prev_prog, const int* __restrict__ out_prog_index,
                                                             int n_out, int n_prog, long long HW, int ocean_out, int ocean_interp,
                                                             const float* __restrict__ ocean, SlabOceanIdx ix) {
  const int b = blockIdx.y;
  float* ob = out + (long long)b * n_out * HW;
  float* d = ob + (long long)ocean_out * HW;
  const int p = out_prog_index[ocean_out];
  float* np = (p >= 0 && next_prog != nullptr) ? next_prog + ((long long)b * n_prog + p) * HW : nullptr;
  const float* om = ocean + (long long)b * 3 * HW;
  const float* tin = prev_prog + ((long long)b * n_prog + ix.prog_sst) * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    auto o = [&](int c) { return ob[(long long)c * HW + i]; };
    const float rad = o(ix.dsw) - o(ix.usw) + o(ix.dlw) - o(ix.ulw);
    const float turb = -o(ix.lhf) - o(ix.shf);
    const float fnet = rad + turb - 0.f;
    const float tend = (fnet + om[HW + i]) / (1000.f * om[2 * HW + i] * 4000.f);  // DENSITY_OF_WATER, SPECIFIC_HEAT_OF_WATER
    const float tgt = tin[i] + tend * ix.dt;
    float v = d[i];
    const float m = om[i];
    if (ocean_interp) v = m * tgt + (1.f - m) * v;
    else if ((int)rintf(m) == 1) v = tgt;
    d[i] = v;
    if (np) np[i] = v;
  }
}

}  // namespace

void launch_norm_split(const float* src, int B, int C, long long HW, const double* stats, const float* gamma,
                       const float* beta, float eps, bf16* dst, long long plane, long long dst_b, long long dst_c,
                       cudaStream_t stream) {
  ProfileScope prof("norm_split", stream);
  ACE_REQUIRE((long long)B * C <= 65535, "norm_split: B*C = %lld exceeds grid.y", (long long)B * C);
  int gx = grid_for((HW + 3) / 4, kThreads, 32);
  dim3 grid(gx, B * C);
  norm_split_kernel<<<grid, kThreads, 0, stream>>>(src, C, HW, stats, gamma, beta, eps, dst, plane, dst_b, dst_c);
  after_launch("norm_split");
}

void launch_split_pad(const float* src, long long rows, int cols, int cols_pad, bf16* dst, long long plane,
                      cudaStream_t stream) {
  ProfileScope prof("split_pad", stream);
  split_pad_kernel<<<grid_for(rows * cols_pad, kThreads), kThreads, 0, stream>>>(src, rows, cols, cols_pad, dst, plane);
  after_launch("split_pad");
}

void launch_prep_dhconv_cplx(const float* w, int Cin, int Cout, int L, int Cinp, bf16* dst, long long plane, cudaStream_t stream) {
  ProfileScope prof("prep_dhconv", stream);
  prep_dhconv_cplx_kernel<<<grid_for((long long)L * 2 * Cout * Cinp, kThreads), kThreads, 0, stream>>>(w, Cin, Cout, L, 1, L, (long long)Cout * L,
                                                                                                       Cinp, dst, plane);
  after_launch("prep_dhconv_cplx");
}

void launch_prep_dhconv_cplx_strided(const float* w, int Cin, int Cout, int L, long long s_l, long long s_o, long long s_i, int Cinp, bf16* dst,
                                     long long plane, cudaStream_t stream) {
  ProfileScope prof("prep_dhconv", stream);
  prep_dhconv_cplx_kernel<<<grid_for((long long)L * 2 * Cout * Cinp, kThreads), kThreads, 0, stream>>>(w, Cin, Cout, L, s_l, s_o, s_i, Cinp, dst,
                                                                                                       plane);
  after_launch("prep_dhconv_cplx");
}

void launch_prep_dhconv_grouped(const float* w, int G, int cg, int L, int cgp, bf16* dst, long long plane, cudaStream_t stream) {
  ProfileScope prof("prep_dhconv", stream);
  prep_dhconv_grouped_kernel<<<grid_for((long long)L * 2 * G * cg * cgp, kThreads), kThreads, 0, stream>>>(w, G, cg, L, cgp, dst, plane);
  after_launch("prep_dhconv_grouped");
}

void launch_diagonal_contract(const bf16* c1, long long c1_plane, const float* w, int B, int C, int L, int M, int Lp,
                              bf16* c2, long long c2_plane, cudaStream_t stream) {
  ProfileScope prof("diagonal_contract", stream);
  diagonal_contract_kernel<<<grid_for((long long)B * L * M * C, kThreads), kThreads, 0, stream>>>(c1, c1_plane, w, B, C, L, M, Lp, c2, c2_plane);
  after_launch("diagonal_contract");
}

void launch_spec_planes_to_complex(const bf16* c1, long long plane, int C, int L, int M, float* out, cudaStream_t stream) {
  ProfileScope prof("spec_planes_to_complex", stream);
  const long long LM = (long long)L * M;
  if (C % 2 == 0 && plane % 2 == 0 && ((uintptr_t)c1 & 3) == 0)
    spec_planes_to_complex2_kernel<<<dim3((unsigned)((LM + 31) / 32), (unsigned)((C + 63) / 64)), 256, 0, stream>>>(c1, plane, C, LM, out);
  else
    spec_planes_to_complex_kernel<<<dim3((unsigned)((LM + 31) / 32), (unsigned)((C + 31) / 32)), 256, 0, stream>>>(c1, plane, C, LM, out);
  after_launch("spec_planes_to_complex");
}

void launch_spec_complex_to_planes(const float* in, int C, int L, int M, int Lp, bf16* c2, long long plane, cudaStream_t stream) {
  ProfileScope prof("spec_complex_to_planes", stream);
  ACE_REQUIRE(L <= 65535 && (C + 31) / 32 <= 65535, "spec_complex_to_planes: extent too large");
  if (C % 2 == 0 && plane % 2 == 0 && ((uintptr_t)c2 & 3) == 0)
    spec_complex_to_planes2_kernel<<<dim3((unsigned)((M + 31) / 32), (unsigned)((C + 63) / 64), (unsigned)L), 256, 0, stream>>>(in, C, L, M, Lp, c2, plane);
  else
    spec_complex_to_planes_kernel<<<dim3((unsigned)((M + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)L), 256, 0, stream>>>(in, C, L, M, Lp, c2, plane);
  after_launch("spec_complex_to_planes");
}

void launch_prep_norm_conv(const double* stats, const float* gamma, const float* beta, float eps, long long HW, int B, int C,
                           const float* w, const float* bias, int O, int Ip, bf16* wout, long long wplane, float* bout,
                           float* a_out, float* s_out, float* shift0_out, cudaStream_t stream) {
  ProfileScope prof("prep_norm_conv", stream);
  dim3 grid(w ? O + 1 : 1, B);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = options().pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ACE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, prep_norm_conv_kernel, stats, gamma, beta, eps, HW, C, w, bias, O, Ip, wout, wplane, bout,
                                    a_out, s_out, shift0_out));
  after_launch("prep_norm_conv");
}

void launch_vec_add(const float* a, const float* b, float* out, long long n, cudaStream_t stream) {
  ProfileScope prof("vec_add", stream);
  vec_add_kernel<<<grid_for(n, kThreads), kThreads, 0, stream>>>(a, b, out, n);
  after_launch("vec_add");
}

void launch_pack_normalize(const float* prog, const float* forcing, int n_prog, int n_forcing, const int* kind,
                           const int* index, const float* mean, const float* std, int B, int n_in, long long HW,
                           float* x, cudaStream_t stream) {
  ProfileScope prof("pack_normalize", stream);
  dim3 grid(grid_for(HW, kThreads, 64), B * n_in);
  pack_normalize_kernel<<<grid, kThreads, 0, stream>>>(prog, forcing, n_prog, n_forcing, kind, index, mean, std, n_in, HW, x);
  after_launch("pack_normalize");
}

void launch_unpack_denormalize(const float* y, const float* x_norm, const int* out_prog_index, const int* prog_in_chan,
                               const float* mean, const float* std, int residual, int B, int n_out, int n_in,
                               int n_prog, long long HW, const int* clamp, int ocean_out, int ocean_interp, const float* ocean,
                               float* out, float* next_prog, cudaStream_t stream) {
  ProfileScope prof("unpack_denormalize", stream);
  dim3 grid(grid_for(HW, kThreads, 64), B * n_out);
  unpack_denormalize_kernel<<<grid, kThreads, 0, stream>>>(y, x_norm, out_prog_index, prog_in_chan, mean, std, residual, n_out, n_in, n_prog, HW, clamp, ocean_out, ocean_interp, ocean, out, next_prog);
  after_launch("unpack_denormalize");
}

void launch_ocean_prescribe(float* out, float* next_prog, const int* out_prog_index, int B, int n_out, int n_prog, long long HW,
                            int ocean_out, int ocean_interp, const float* ocean, cudaStream_t stream) {
  ProfileScope prof("ocean_prescribe", stream);
  dim3 grid(grid_for(HW, kThreads, 64), B);
  ocean_prescribe_kernel<<<grid, kThreads, 0, stream>>>(out, next_prog, out_prog_index, n_out, n_prog, HW, ocean_out, ocean_interp, ocean);
  after_launch("ocean_prescribe");
}

void launch_ocean_slab(float* out, float* next_prog, const float* prev_prog, const int* out_prog_index, int B, int n_out, int n_prog, long long HW,
                       int ocean_out, int ocean_interp, const float* ocean, const SlabOceanIdx& ix, cudaStream_t stream) {
  ProfileScope prof("ocean_slab", stream);
  dim3 grid(grid_for(HW, kThreads, 64), B);
  ocean_slab_kernel<<<grid, kThreads, 0, stream>>>(out, next_prog, prev_prog, out_prog_index, n_out, n_prog, HW, ocean_out, ocean_interp, ocean, ix);
  after_launch("ocean_slab");
}

}  // namespace ace
