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
// Kernel shape: one thread per pixel (coalesced along the pixel axis for every channel row), 128 pixels per block, grid
// (ceil(HW / 128), B).  Pass 1 reads the C channel values of the pixel and accumulates sum / sum of squares in fp64; pass 2
// re-reads them (L2), with the pixel's E context values in registers and the (scale, bias) weight pairs of 16 channels at a
// time staged in shared memory (read as warp-uniform broadcasts).  Algorithmic traffic per norm: 2 reads + 1 write of the
// [C][HW] planes (4 B per element each) + E * HW * 4 B of context.
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

}  // namespace

int cln_padded_context(int E) {
  if (E <= 0) return 0;
  if (E <= 8) return 8;
  if (E <= 16) return 16;
  if (E <= 32) return 32;
  if (E <= 64) return 64;
  return -1;
}

void launch_cond_layer_norm(const bf16* x, long long x_plane, long long x_b, int B, int C, long long HW, const float* lnw, const float* lnb,
                            const float* sb0, const float* w2, const float* ctx, int Ep, float eps, bf16* out, long long o_plane,
                            long long o_b, cudaStream_t stream) {
  ProfileScope prof("cond_layer_norm", stream);
  const dim3 grid((unsigned)((HW + kPix - 1) / kPix), (unsigned)B);
#define ACE_CLN(E) cond_layer_norm_kernel<E><<<grid, kPix, 0, stream>>>(x, x_plane, x_b, C, HW, lnw, lnb, sb0, w2, ctx, eps, out, o_plane, o_b)
  switch (Ep) {
    case 0: ACE_CLN(0); break;
    case 8: ACE_CLN(8); break;
    case 16: ACE_CLN(16); break;
    case 32: ACE_CLN(32); break;
    case 64: ACE_CLN(64); break;
    default: throw Error(ACE_ERR_INVALID, "cond_layer_norm: padded context width must be 0, 8, 16, 32 or 64");
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
