// ConditionalLayerNorm of the noise-conditioned SFNO.
//
// Reference: /root/reference/fme/core/models/conditional_sfno/layers.py:95-141 (ChannelLayerNorm: per pixel, biased variance
// over the channel axis) and :285-320 (ConditionalLayerNorm.forward):
//     y = LN(x) * scale + bias,
//     scale = [W_scale(scalar) | 1] + W_scale_2d(noise) + W_scale_labels(labels) + W_scale_pos(pos)
//     bias  = [W_bias(scalar)  | 0] + W_bias_2d(noise)  + W_bias_labels(labels)  + W_bias_pos(pos)
// where the *_2d / *_pos maps are bias-free 1x1 convolutions of per-pixel context fields and the others are Linear layers of
// per-sample vectors.  Here: the per-(sample, channel) part is a tiny pre-kernel (cln_vector_terms), the per-pixel context
// channels (noise, then positional) are concatenated once per forward into ctx [B][E][HW], and ONE streaming kernel per norm
// does statistics, normalisation, the two E-term dot products per output element and the split-bf16 store.
//
// Kernel shape (cond_layer_norm_kernel2, HW even): a block of 8 warps owns 64 consecutive pixels; lane l of EVERY warp holds
// pixels 2l, 2l+1 (bf16x2 loads / stores: 128 B per warp and channel row) and warp w takes the channels c = w (mod 8), so a
// sample offers HW/64 * 8 warps of parallelism (8 100 at 1 degree).  Pass 1: per-warp partial sums / sums of squares in fp64,
// combined through shared memory.  Pass 2 re-reads the tile (L2) with the two pixels' E context values in fp32x2 register
// pairs; the (scale, bias) weights of 32 channels at a time sit in shared memory and every warp-uniform read of one context
// channel's pair feeds two FFMA2 (both pixels' scale and bias accumulators).  Algorithmic traffic per norm:
// 2 reads + 1 write of the [C][HW] planes (4 B per element each) + E * HW * 4 B of context; 2 * E FMA per element.
// cond_layer_norm_kernel (one thread per pixel) is the generic path for odd HW.
#include "kernels.cuh"

namespace ace {
namespace {

constexpr int kPix = 128;  // pixels (= threads) per block
constexpr int kCh = 16;    // channels whose conditioning weights are staged per shared-memory chunk

template <int E>
__global__ void __launch_bounds__(kPix) cond_layer_norm_kernel(const bf16* __restrict__ x, long long x_plane, long long x_b, int C,
                                                              long long HW, const float* __restrict__ lnw,
                                                              const float* __restrict__ lnb, const float* __restrict__ sb0,
                                                              const float* __restrict__ w2, const float* __restrict__ ctx, float eps,
                                                              bf16* __restrict__ out, long long o_plane, long long o_b) {
  __shared__ float wsm[kCh * (E > 0 ? E : 1) * 2];
  const int b = blockIdx.y;
  const long long p = (long long)blockIdx.x * kPix + threadIdx.x;
  const bool live = p < HW;
  const bf16* xb = x + (long long)b * x_b + p;
  double s = 0.0, q = 0.0;
  if (live) {
    for (int c = 0; c < C; ++c) {
      const float v = __bfloat162float(xb[(long long)c * HW]) + __bfloat162float(xb[(long long)c * HW + x_plane]);
      s += (double)v;
      q += (double)v * (double)v;
    }
  }
  const double mean_d = s / C;
  const float mean = (float)mean_d;
  const float rstd = rsqrtf((float)fmax(q / C - mean_d * mean_d, 0.0) + eps);
  float cx[E > 0 ? E : 1];
  if (E > 0) {
#pragma unroll
    for (int e = 0; e < E; ++e) cx[e] = live ? __ldg(ctx + ((long long)b * E + e) * HW + p) : 0.f;
  }
  bf16* ob = out + (long long)b * o_b + p;
  for (int c0 = 0; c0 < C; c0 += kCh) {
    const int nch = min(kCh, C - c0);
    if (E > 0) {
      __syncthreads();
      for (int i = threadIdx.x; i < nch * E * 2; i += kPix) wsm[i] = __ldg(w2 + (long long)c0 * E * 2 + i);
      __syncthreads();
    }
    if (!live) continue;
    for (int cc = 0; cc < nch; ++cc) {
      const int c = c0 + cc;
      float sc = sb0 ? __ldg(sb0 + ((long long)b * C + c) * 2) : 1.f;
      float bi = sb0 ? __ldg(sb0 + ((long long)b * C + c) * 2 + 1) : 0.f;
      if (E > 0) {
        const float2* wp = reinterpret_cast<const float2*>(wsm + cc * E * 2);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float2 w = wp[e];
          sc = fmaf(w.x, cx[e], sc);
          bi = fmaf(w.y, cx[e], bi);
        }
      }
      const float v = __bfloat162float(xb[(long long)c * HW]) + __bfloat162float(xb[(long long)c * HW + x_plane]);
      float y = (v - mean) * rstd;
      if (lnw) y = fmaf(y, __ldg(lnw + c), __ldg(lnb + c));
      y = fmaf(y, sc, bi);
      bf16 hi, lo;
      split_bf16(y, hi, lo);
      ob[(long long)c * HW] = hi;
      ob[(long long)c * HW + o_plane] = lo;
    }
  }
}

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

constexpr int kWarps = 8;     // warps per block = channel interleave
constexpr int kPix2 = 64;     // pixels per block (2 per lane)
constexpr int kCh2 = 32;      // channels per shared-memory weight chunk

__device__ __forceinline__ void load_pair(const bf16* p, long long plane, float& v0, float& v1) {
  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(p);
  const __nv_bfloat162 l = *reinterpret_cast<const __nv_bfloat162*>(p + plane);
  v0 = __low2float(h) + __low2float(l);
  v1 = __high2float(h) + __high2float(l);
}

template <int E>
__global__ void __launch_bounds__(kWarps * 32) cond_layer_norm_kernel2(const bf16* __restrict__ x, long long x_plane, long long x_b, int C,
                                                                      long long HW, const float* __restrict__ lnw,
                                                                      const float* __restrict__ lnb, const float* __restrict__ sb0,
                                                                      const float* __restrict__ w2, const float* __restrict__ ctx, float eps,
                                                                      bf16* __restrict__ out, long long o_plane, long long o_b) {
  __shared__ __align__(16) float wsm[2 * kCh2 * (E > 0 ? E : 1) * 2];  // two buffers of [channel][e] -> (ws, wb)
  __shared__ double red[kWarps][32][4];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * kPix2 + 2 * lane;  // HW is even: a pair is live or dead as a whole
  const bool live = p < HW;
  const bf16* xb = x + (long long)b * x_b + p;
  // ---- pass 1: statistics over the channel axis
  double s0 = 0, q0 = 0, s1 = 0, q1 = 0;
  if (live) {
#pragma unroll 8
    for (int c = warp; c < C; c += kWarps) {
      float v0, v1;
      load_pair(xb + (long long)c * HW, x_plane, v0, v1);
      s0 += (double)v0;
      q0 += (double)v0 * (double)v0;
      s1 += (double)v1;
      q1 += (double)v1 * (double)v1;
    }
  }
  red[warp][lane][0] = s0;
  red[warp][lane][1] = q0;
  red[warp][lane][2] = s1;
  red[warp][lane][3] = q1;
  __syncthreads();
  s0 = q0 = s1 = q1 = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    s0 += red[w][lane][0];
    q0 += red[w][lane][1];
    s1 += red[w][lane][2];
    q1 += red[w][lane][3];
  }
  const double m0 = s0 / C, m1 = s1 / C;
  const float mean0 = (float)m0, mean1 = (float)m1;
  const float rstd0 = rsqrtf((float)fmax(q0 / C - m0 * m0, 0.0) + eps), rstd1 = rsqrtf((float)fmax(q1 / C - m1 * m1, 0.0) + eps);
  // ---- the two pixels' context values as (v, v) pairs: ptxas turns them into the scalar-broadcast operand of FFMA2, so the
  // (ws, wb) weight pair of a context channel -- 8 bytes of a warp-uniform shared-memory read -- feeds both pixels'
  // (scale, bias) accumulators with two FFMA2.
  f2 cx[E > 0 ? 2 * E : 1];
  if (E > 0) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float2 v = make_float2(0.f, 0.f);
      if (live) v = __ldg(reinterpret_cast<const float2*>(ctx + ((long long)b * E + e) * HW + p));
      cx[2 * e] = mk2(v.x, v.x);
      cx[2 * e + 1] = mk2(v.y, v.y);
    }
  }
  bf16* ob = out + (long long)b * o_b + p;
  // ---- pass 2: normalise, conditional affine map, split store.  The weights of 32 channels at a time are double-buffered
  // in shared memory with cp.async so that the next chunk's global loads overlap this chunk's arithmetic.
  constexpr int kChunkVec = kCh2 * (E > 0 ? E : 1) / 2;  // 16-byte vectors per chunk ((ws, wb) pairs of two context channels)
  auto stage = [&](int c0, int buf) {
    if (E > 0) {
      const int nvec = min(kCh2, C - c0) * E / 2;
      const float4* src = reinterpret_cast<const float4*>(w2 + (long long)c0 * E * 2);
      float4* dst = reinterpret_cast<float4*>(wsm) + buf * kChunkVec;
      for (int i = threadIdx.x; i < nvec; i += kWarps * 32) {
        const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(src + i) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0, 0);
  // software pipeline of the tile re-read: the (hi, lo) words of this warp's next two channels are always in flight, so the
  // L2 latency hides behind two channels' worth of FFMA2 instead of stalling every iteration
  auto ldraw = [&](int c, unsigned& h, unsigned& l) {
    if (live && c < C) {
      h = __ldg(reinterpret_cast<const unsigned*>(xb + (long long)c * HW));
      l = __ldg(reinterpret_cast<const unsigned*>(xb + (long long)c * HW + x_plane));
    }
  };
  auto unpack = [](unsigned h, unsigned l, float& v0, float& v1) {
    v0 = __uint_as_float(h << 16) + __uint_as_float(l << 16);  // bf16 -> fp32 is a 16-bit shift
    v1 = __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u);
  };
  unsigned hA = 0, lA = 0, hB = 0, lB = 0, hC = 0, lC = 0;
  ldraw(warp, hA, lA);
  ldraw(warp + kWarps, hB, lB);
  // running pointers (the per-iteration strides are loop constants: no 64-bit multiplies on the FMA pipe inside the loop)
  const long long cstep = (long long)kWarps * HW;
  const unsigned* xpre = reinterpret_cast<const unsigned*>(xb + (long long)(warp + 2 * kWarps) * HW);  // bf16x2 words
  bf16* op = ob + (long long)warp * HW;
  const float* sbp = sb0 ? sb0 + ((long long)b * C + warp) * 2 : nullptr;
  int buf = 0;
  for (int c0 = 0; c0 < C; c0 += kCh2, buf ^= 1) {
    const int nch = min(kCh2, C - c0);
    if (c0 + kCh2 < C) {
      stage(c0 + kCh2, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();  // chunk `buf` has landed for every thread
    if (live) {
      const float2* wp = reinterpret_cast<const float2*>(wsm) + (buf * kCh2 + warp) * E;
      for (int cc = warp; cc < nch; cc += kWarps, wp += kWarps * E, xpre += cstep / 2, op += cstep) {
        const int c = c0 + cc;
        if (c + 2 * kWarps < C) {
          hC = __ldg(xpre);
          lC = __ldg(xpre + x_plane / 2);
        }
        float sc_i = 1.f, bi_i = 0.f;
        if (sbp) {
          sc_i = __ldg(sbp);
          bi_i = __ldg(sbp + 1);
          sbp += 2 * kWarps;
        }
        f2 a0 = mk2(sc_i, bi_i), a1 = a0;  // (scale, bias) of pixel 0 / pixel 1
        if (E > 0) {
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const float2 w = wp[e];
            const f2 wv = mk2(w.x, w.y);
            a0 = fma2(wv, cx[2 * e], a0);
            a1 = fma2(wv, cx[2 * e + 1], a1);
          }
        }
        float v0, v1, sc0, sc1, bi0, bi1;
        unpack(hA, lA, v0, v1);
        hA = hB, lA = lB, hB = hC, lB = lC;
        un2(a0, sc0, bi0);
        un2(a1, sc1, bi1);
        float y0 = (v0 - mean0) * rstd0, y1 = (v1 - mean1) * rstd1;
        if (lnw) {
          const float g = __ldg(lnw + c), h = __ldg(lnb + c);
          y0 = fmaf(y0, g, h);
          y1 = fmaf(y1, g, h);
        }
        y0 = fmaf(y0, sc0, bi0);
        y1 = fmaf(y1, sc1, bi1);
        // split-bf16 store: hi = round-to-nearest bf16 of y, lo = bf16 of the remainder; two pixels per 32-bit word
        const __nv_bfloat162 hi = __floats2bfloat162_rn(y0, y1);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(y0 - __low2float(hi), y1 - __high2float(hi));
        *reinterpret_cast<__nv_bfloat162*>(op) = hi;
        *reinterpret_cast<__nv_bfloat162*>(op + o_plane) = lo;
      }
    }
    __syncthreads();  // everyone is done with chunk `buf` before it is refilled two iterations later
  }
}

// sb0[b][c] = {scale0, bias0}: the Linear layers on the per-sample context vectors (layers.py:289-311)
__global__ void cln_vector_terms_kernel(const float* __restrict__ scalar, int Es, const float* __restrict__ labels, int El,
                                        const float* __restrict__ Ws, const float* __restrict__ bs, const float* __restrict__ Wb,
                                        const float* __restrict__ bb, const float* __restrict__ Wsl, const float* __restrict__ bsl,
                                        const float* __restrict__ Wbl, const float* __restrict__ bbl, int C, float* __restrict__ sb0) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sc = 1.f, bi = 0.f;
  if (Es > 0) {
    sc = bs[c];
    bi = bb[c];
    for (int e = 0; e < Es; ++e) {
      const float v = scalar[(long long)b * Es + e];
      sc = fmaf(Ws[(long long)c * Es + e], v, sc);
      bi = fmaf(Wb[(long long)c * Es + e], v, bi);
    }
  }
  if (El > 0) {
    float s2 = bsl[c], b2 = bbl[c];
    for (int e = 0; e < El; ++e) {
      const float v = labels[(long long)b * El + e];
      s2 = fmaf(Wsl[(long long)c * El + e], v, s2);
      b2 = fmaf(Wbl[(long long)c * El + e], v, b2);
    }
    sc += s2;
    bi += b2;
  }
  sb0[((long long)b * C + c) * 2] = sc;
  sb0[((long long)b * C + c) * 2 + 1] = bi;
}

__global__ void concat_ctx_kernel(const float* __restrict__ noise, int En, const float* __restrict__ pos, int Epos, long long HW, int Ep,
                                  float* __restrict__ ctx, long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long p = idx % HW;
    const long long t = idx / HW;
    const int e = (int)(t % Ep);
    const long long b = t / Ep;
    float v = 0.f;
    if (e < En) v = noise[(b * En + e) * HW + p];
    else if (e < En + Epos) v = pos[(b * Epos + (e - En)) * HW + p];
    ctx[idx] = v;
  }
}

// w2[c][e] = {scale weight, bias weight} of context channel e (noise channels first, then positional, zero padding)
__global__ void build_cln_w2_kernel(const float* __restrict__ ws_n, const float* __restrict__ wb_n, int En, const float* __restrict__ ws_p,
                                    const float* __restrict__ wb_p, int Epos, int C, int Ep, float* __restrict__ w2) {
  const long long total = (long long)C * Ep;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(idx % Ep);
    const long long c = idx / Ep;
    float s = 0.f, b = 0.f;
    if (e < En) {
      s = ws_n[c * En + e];
      b = wb_n[c * En + e];
    } else if (e < En + Epos) {
      s = ws_p[c * Epos + (e - En)];
      b = wb_p[c * Epos + (e - En)];
    }
    w2[idx * 2] = s;
    w2[idx * 2 + 1] = b;
  }
}

// ---- helpers of the tensor-core ConditionalLayerNorm (GemmOp::cln, gemm.cuh) ----

// {mean, rstd} over the channel axis per pixel (pass 1 of cond_layer_norm_kernel2 on its own): musr[b][p]
__global__ void __launch_bounds__(kWarps * 32) cln_stats_kernel(const bf16* __restrict__ x, long long x_plane, long long x_b, int C, long long HW,
                                                               float eps, float2* __restrict__ musr) {
  __shared__ double red[kWarps][32][4];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * kPix2 + 2 * lane;
  const bool live = p < HW;
  const bf16* xb = x + (long long)b * x_b + p;
  double s0 = 0, q0 = 0, s1 = 0, q1 = 0;
  if (live) {
#pragma unroll 8
    for (int c = warp; c < C; c += kWarps) {
      float v0, v1;
      load_pair(xb + (long long)c * HW, x_plane, v0, v1);
      s0 += (double)v0;
      q0 += (double)v0 * (double)v0;
      s1 += (double)v1;
      q1 += (double)v1 * (double)v1;
    }
  }
  red[warp][lane][0] = s0;
  red[warp][lane][1] = q0;
  red[warp][lane][2] = s1;
  red[warp][lane][3] = q1;
  __syncthreads();
  if (warp != 0 || !live) return;
  s0 = q0 = s1 = q1 = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    s0 += red[w][lane][0];
    q0 += red[w][lane][1];
    s1 += red[w][lane][2];
    q1 += red[w][lane][3];
  }
  const double m0 = s0 / C, m1 = s1 / C;
  const float r0 = rsqrtf((float)fmax(q0 / C - m0 * m0, 0.0) + eps), r1 = rsqrtf((float)fmax(q1 / C - m1 * m1, 0.0) + eps);
  *reinterpret_cast<float4*>(musr + (long long)b * HW + p) = make_float4((float)m0, r0, (float)m1, r1);
}

// The same statistics with four pixels per lane (8-byte loads: a warp covers 128 pixels of a channel row per request instead of 64,
// twice the bytes in flight per thread) -- needs HW % 4 == 0 and 8-byte aligned channel rows, which the tensor-core path has anyway
constexpr int kPix4 = 128;
__global__ void __launch_bounds__(kWarps * 32) cln_stats4_kernel(const bf16* __restrict__ x, long long x_plane, long long x_b, int C, long long HW,
                                                                float eps, float2* __restrict__ musr) {
  __shared__ double red[kWarps][32][8];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * kPix4 + 4 * lane;
  const bool live = p < HW;  // HW % 4 == 0: a quad is live or dead as a whole
  const bf16* xb = x + (long long)b * x_b + p;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (live) {
#pragma unroll 4
    for (int c = warp; c < C; c += kWarps) {
      const uint2 h = __ldg(reinterpret_cast<const uint2*>(xb + (long long)c * HW));
      const uint2 l = __ldg(reinterpret_cast<const uint2*>(xb + (long long)c * HW + x_plane));
      const float v[4] = {__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16), __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u),
                          __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16), __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double d = (double)v[i];
        s[i] += d;
        q[i] = fma(d, d, q[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[warp][lane][2 * i] = s[i];
    red[warp][lane][2 * i + 1] = q[i];
  }
  __syncthreads();
  // 32 lanes x 4 pixels = 128 results: thread t < 128 finishes pixel t (lane t / 4, slot t % 4)
  const int t = threadIdx.x;
  if (t >= kPix4) return;
  const int ln = t >> 2, i = t & 3;
  const long long pp = (long long)blockIdx.x * kPix4 + t;
  if (pp >= HW) return;
  double ss = 0, qq = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    ss += red[w][ln][2 * i];
    qq += red[w][ln][2 * i + 1];
  }
  const double m = ss / C;
  musr[(long long)b * HW + pp] = make_float2((float)m, rsqrtf((float)fmax(qq / C - m * m, 0.0) + eps));
}

// ctx [B][Ep][HW] fp32 -> K-major B operand planes [B][HW][2 Ep]: the Ep context channels twice along k (the first half meets
// the scale weights, the second half the bias weights); 32-pixel x 64-channel tiles are transposed through shared memory
// (any Ep: the channel axis is walked 64 at a time)
__global__ void __launch_bounds__(256) cln_ctx_planes_kernel(const float* __restrict__ ctx, int EP, long long HW, bf16* __restrict__ out,
                                                            long long plane) {
  __shared__ float tile[64][33];
  const int b = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * 32;
  for (int e0 = 0; e0 < EP; e0 += 64) {
    const int ne = min(64, EP - e0);
    __syncthreads();
    for (int i = threadIdx.x; i < ne * 32; i += 256) {
      const int e = i >> 5, j = i & 31;
      tile[e][j] = (p0 + j < HW) ? ctx[((long long)b * EP + e0 + e) * HW + p0 + j] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * ne; i += 256) {
      const int j = i / ne, e = i % ne;
      if (p0 + j >= HW) continue;
      bf16 h, l;
      split_bf16(tile[e][j], h, l);
      bf16* o = out + ((long long)b * HW + p0 + j) * (2 * EP) + e0 + e;
      o[0] = h;
      o[EP] = h;
      o[plane] = l;
      o[plane + EP] = l;
    }
  }
}

// Context wider than 64 channels where the tensor-core path (GemmOp::cln) does not apply (fewer than 128 channels: the
// conditionally normalised big skip of the network input; pixel counts that are not a multiple of 4): any E, the pixel's context
// values are re-read per channel through L1 (E * 64 pixels * 4 B per block) instead of living in registers.  A block owns 64
// pixels; its four thread groups of two warps take the channels c = group (mod 4).
constexpr int kPixW = 64, kGroupsW = 4;
__global__ void __launch_bounds__(kPixW * kGroupsW) cond_layer_norm_wide_kernel(const bf16* __restrict__ x, long long x_plane, long long x_b, int C,
                                                                              long long HW, const float* __restrict__ lnw,
                                                                              const float* __restrict__ lnb, const float* __restrict__ sb0,
                                                                              const float* __restrict__ w2, const float* __restrict__ ctx, int E,
                                                                              float eps, bf16* __restrict__ out, long long o_plane, long long o_b) {
  __shared__ double red[kGroupsW][kPixW][2];
  const int b = blockIdx.y, pix = threadIdx.x % kPixW, grp = threadIdx.x / kPixW;
  const long long p = (long long)blockIdx.x * kPixW + pix;
  const bool live = p < HW;
  const bf16* xb = x + (long long)b * x_b + p;
  double s = 0.0, q = 0.0;
  if (live) {
    for (int c = grp; c < C; c += kGroupsW) {
      const float v = __bfloat162float(xb[(long long)c * HW]) + __bfloat162float(xb[(long long)c * HW + x_plane]);
      s += (double)v;
      q += (double)v * (double)v;
    }
  }
  red[grp][pix][0] = s;
  red[grp][pix][1] = q;
  __syncthreads();
  if (!live) return;
  s = q = 0.0;
#pragma unroll
  for (int g = 0; g < kGroupsW; ++g) {
    s += red[g][pix][0];
    q += red[g][pix][1];
  }
  const double mean_d = s / C;
  const float mean = (float)mean_d;
  const float rstd = rsqrtf((float)fmax(q / C - mean_d * mean_d, 0.0) + eps);
  const float* cp = ctx + (long long)b * E * HW + p;
  bf16* ob = out + (long long)b * o_b + p;
  for (int c = grp; c < C; c += kGroupsW) {
    float sc = sb0 ? __ldg(sb0 + ((long long)b * C + c) * 2) : 1.f;
    float bi = sb0 ? __ldg(sb0 + ((long long)b * C + c) * 2 + 1) : 0.f;
    const float2* wp = reinterpret_cast<const float2*>(w2) + (long long)c * E;
    float sc1 = 0.f, bi1 = 0.f;  // two independent chains per accumulator
    int e = 0;
    for (; e + 1 < E; e += 2) {
      const float2 wa = __ldg(wp + e), wb = __ldg(wp + e + 1);
      const float va = __ldg(cp + (long long)e * HW), vb = __ldg(cp + (long long)(e + 1) * HW);
      sc = fmaf(wa.x, va, sc);
      bi = fmaf(wa.y, va, bi);
      sc1 = fmaf(wb.x, vb, sc1);
      bi1 = fmaf(wb.y, vb, bi1);
    }
    if (e < E) {
      const float2 wa = __ldg(wp + e);
      const float va = __ldg(cp + (long long)e * HW);
      sc = fmaf(wa.x, va, sc);
      bi = fmaf(wa.y, va, bi);
    }
    sc += sc1;
    bi += bi1;
    const float v = __bfloat162float(xb[(long long)c * HW]) + __bfloat162float(xb[(long long)c * HW + x_plane]);
    float y = (v - mean) * rstd;
    if (lnw) y = fmaf(y, __ldg(lnw + c), __ldg(lnb + c));
    y = fmaf(y, sc, bi);
    bf16 hi, lo;
    split_bf16(y, hi, lo);
    ob[(long long)c * HW] = hi;
    ob[(long long)c * HW + o_plane] = lo;
  }
}

// w2 [C][Ep][2] -> K-major A operand planes [C][2 Ep] = [scale weights | bias weights]
__global__ void cln_w_planes_kernel(const float* __restrict__ w2, int C, int Ep, bf16* __restrict__ out, long long plane) {
  const long long total = (long long)C * Ep;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / Ep;
    const int e = (int)(idx % Ep);
    bf16 h, l;
    split_bf16(w2[idx * 2], h, l);
    out[c * 2 * Ep + e] = h;
    out[c * 2 * Ep + e + plane] = l;
    split_bf16(w2[idx * 2 + 1], h, l);
    out[c * 2 * Ep + Ep + e] = h;
    out[c * 2 * Ep + Ep + e + plane] = l;
  }
}

}  // namespace

void launch_cln_stats(const bf16* x, long long x_plane, long long x_b, int B, int C, long long HW, float eps, float* musr, cudaStream_t stream) {
  ProfileScope prof("cln_stats", stream);
  if (HW % 4 == 0 && x_plane % 4 == 0 && x_b % 4 == 0 && ((uintptr_t)x & 7) == 0)
    cln_stats4_kernel<<<dim3((unsigned)((HW + kPix4 - 1) / kPix4), (unsigned)B), kWarps * 32, 0, stream>>>(x, x_plane, x_b, C, HW, eps,
                                                                                                      reinterpret_cast<float2*>(musr));
  else
    cln_stats_kernel<<<dim3((unsigned)((HW + kPix2 - 1) / kPix2), (unsigned)B), kWarps * 32, 0, stream>>>(x, x_plane, x_b, C, HW, eps,
                                                                                                      reinterpret_cast<float2*>(musr));
  after_launch("cln_stats");
}

void launch_cln_ctx_planes(const float* ctx, int B, int Ep, long long HW, bf16* out, long long plane, cudaStream_t stream) {
  ProfileScope prof("cln_ctx_planes", stream);
  const dim3 grid((unsigned)((HW + 31) / 32), (unsigned)B);
  if (Ep <= 0 || Ep % 32 != 0) throw Error(ACE_ERR_INVALID, "cln_ctx_planes: padded context width must be a positive multiple of 32");
  cln_ctx_planes_kernel<<<grid, 256, 0, stream>>>(ctx, Ep, HW, out, plane);
  after_launch("cln_ctx_planes");
}

void launch_cln_w_planes(const float* w2, int C, int Ep, bf16* out, long long plane, cudaStream_t stream) {
  const long long total = (long long)C * Ep;
  cln_w_planes_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, stream>>>(w2, C, Ep, out, plane);
  after_launch("cln_w_planes");
}

int cln_padded_context(int E) {
  if (E <= 0) return 0;
  if (E <= 8) return 8;
  if (E <= 16) return 16;
  if (E <= 32) return 32;
  if (E <= 64) return 64;
  return (E + 31) / 32 * 32;  // wider: the tensor-core path (GemmOp::cln, K = 2 Ep) or cond_layer_norm_wide_kernel
}

void launch_cond_layer_norm(const bf16* x, long long x_plane, long long x_b, int B, int C, long long HW, const float* lnw, const float* lnb,
                            const float* sb0, const float* w2, const float* ctx, int Ep, float eps, bf16* out, long long o_plane,
                            long long o_b, cudaStream_t stream) {
  ProfileScope prof("cond_layer_norm", stream);
  // the paired kernel needs 4-byte aligned bf16x2 / 8-byte aligned context pairs at every channel row of every sample
  const bool paired = (HW % 2 == 0) && (x_plane % 2 == 0) && (x_b % 2 == 0) && (o_plane % 2 == 0) && (o_b % 2 == 0) &&
                      ((uintptr_t)x % 4 == 0) && ((uintptr_t)out % 4 == 0) && ((uintptr_t)ctx % 8 == 0) && !options().force_simt;
  const dim3 grid(paired ? (unsigned)((HW + kPix2 - 1) / kPix2) : (unsigned)((HW + kPix - 1) / kPix), (unsigned)B);
#define ACE_CLN(E)                                                                                                                      \
  do {                                                                                                                                  \
    if (paired)                                                                                                                         \
      cond_layer_norm_kernel2<E><<<grid, kWarps * 32, 0, stream>>>(x, x_plane, x_b, C, HW, lnw, lnb, sb0, w2, ctx, eps, out, o_plane, o_b); \
    else                                                                                                                                \
      cond_layer_norm_kernel<E><<<grid, kPix, 0, stream>>>(x, x_plane, x_b, C, HW, lnw, lnb, sb0, w2, ctx, eps, out, o_plane, o_b);      \
  } while (0)
  switch (Ep) {
    case 0: ACE_CLN(0); break;
    case 8: ACE_CLN(8); break;
    case 16: ACE_CLN(16); break;
    case 32: ACE_CLN(32); break;
    case 64: ACE_CLN(64); break;
    default:
      if (Ep < 0) throw Error(ACE_ERR_INVALID, "cond_layer_norm: negative context width");
      cond_layer_norm_wide_kernel<<<dim3((unsigned)((HW + kPixW - 1) / kPixW), (unsigned)B), kPixW * kGroupsW, 0, stream>>>(
          x, x_plane, x_b, C, HW, lnw, lnb, sb0, w2, ctx, Ep, eps, out, o_plane, o_b);
  }
#undef ACE_CLN
  after_launch("cond_layer_norm");
}

void launch_cln_vector_terms(const float* scalar, int Es, const float* labels, int El, const float* Ws, const float* bs, const float* Wb,
                             const float* bb, const float* Wsl, const float* bsl, const float* Wbl, const float* bbl, int B, int C, float* sb0,
                             cudaStream_t stream) {
  ProfileScope prof("cln_vector_terms", stream);
  cln_vector_terms_kernel<<<dim3((unsigned)((C + 127) / 128), (unsigned)B), 128, 0, stream>>>(scalar, Es, labels, El, Ws, bs, Wb, bb, Wsl, bsl,
                                                                                          Wbl, bbl, C, sb0);
  after_launch("cln_vector_terms");
}

void launch_concat_ctx(const float* noise, int En, const float* pos, int Epos, int B, long long HW, int Ep, float* ctx, cudaStream_t stream) {
  ProfileScope prof("concat_ctx", stream);
  const long long total = (long long)B * Ep * HW;
  concat_ctx_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, stream>>>(noise, En, pos, Epos, HW, Ep, ctx, total);
  after_launch("concat_ctx");
}

void launch_build_cln_w2(const float* ws_n, const float* wb_n, int En, const float* ws_p, const float* wb_p, int Epos, int C, int Ep, float* w2,
                         cudaStream_t stream) {
  const long long total = (long long)C * Ep;
  build_cln_w2_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, stream>>>(ws_n, wb_n, En, ws_p, wb_p, Epos, C, Ep, w2);
  after_launch("build_cln_w2");
}

}  // namespace ace
