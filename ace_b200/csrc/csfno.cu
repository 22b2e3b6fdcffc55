// Noise-conditioned SFNO forward (ace_csfno_*): parameter re-layout + kernel orchestration.
//
// Mirrors /root/reference/fme/core/models/conditional_sfno/sfnonet.py:773-824 (net forward), :376-436 (block forward),
// s2convolutions.py:359-433 (SpectralConvS2.forward, filter_type="linear", one group) and layers.py:285-320
// (ConditionalLayerNorm) as a fixed sequence of GemmOps (gemm.cuh) and streaming kernels.  Per block
// (C = embed_dim, activations are split-bf16 planes [B][C][HW]):
//
//   h --CLN0(ctx)--> xn --G1 DFT--> X1 --G2 Legendre--> c1 --G3 dhconv--> c2 --G4 Legendre--> g --G5 iDFT--> T (fp32)
//   r = xn, or iSHT(SHT(xn)) when the block's forward / inverse grids differ (first / last block on an equiangular data grid)
//       (or in every block with filter_residual)
//   t = GELU(T + b_filter + W_skip r + b_skip) --CLN1(ctx)--> tn --fc1 + GELU--> hmid --fc2 + b + r--> h (next block)
//
// filter_residual also passes the big-skip copy of the input through trans_down -> itrans_up, filter_output the network output
// (sfnonet.py:581-591,775-778,822): four GemmOps each on the data-grid plan with their own (zero-initialised) spectral buffers.
// Grouped / LoRA / mean-preserving / bottlenecked spectral operators arrive as the equivalent dense [L][C][C] operator
// (ace_b200/csfno.py folds them when a parameter changes), so the filter is always the one complex GEMM per degree below.
//
// Unlike the InstanceNorm of the deterministic SFNO (sfno.cu), the conditional layer norm is per PIXEL over channels with a
// per-pixel, per-channel affine map (functions of the noise field), so it cannot be folded into the neighbouring GEMMs'
// per-row epilogues; it is a statistics pass + one GEMM with a normalising epilogue, or one streaming kernel, per norm (cln.cu).
#include <map>
#include <string>
#include <vector>

#include "netops.cuh"

using namespace ace;

namespace {

// one ConditionalLayerNorm over `C` channels
struct ClnW {
  int C = 0;
  DevBuf lnw, lnb;                          // elementwise affine of the channel LayerNorm (affine_norms)
  DevBuf Ws, bs, Wb, bb, Wsl, bsl, Wbl, bbl;  // Linear layers on the scalar embedding / labels
  DevBuf ws_n, wb_n, ws_p, wb_p;            // 1x1-conv weights on the noise / positional context, fp32 as given
  DevBuf w2;                                // [C][Ep][2] built by finalize
  DevBuf wP;                                // planes [C][2 Ep] = [scale | bias] weights: A operand of the tensor-core path
  long long wP_plane = 0;
};

// buffers of one outer-grid SHT round trip at a fixed channel count (the never-written l < m region of c1 must stay zero, so
// round trips of different widths do not share them)
struct RtBuf {
  DevBuf x1, c1, g;
  long long p_x1 = 0, p_c1 = 0, p_g = 0;
  int B = 0;
  void ensure(const ace_sht_plan& p, int C, int batch) {
    if (batch <= B) return;
    p_x1 = batch * p.x1_elems(C);
    p_c1 = batch * p.c1_elems(C);
    p_g = batch * p.g_elems(C);
    x1.ensure(2 * (size_t)p_x1 * sizeof(bf16));
    c1.ensure(2 * (size_t)p_c1 * sizeof(bf16));
    g.ensure(2 * (size_t)p_g * sizeof(bf16));
    B = batch;
  }
};

struct CBlockW {
  ClnW n0, n1;
  DevBuf spec;  // planes [L][2 (re, im)][C][Cp]; grouped (groups > 1): [L][2][C][cgp], row o holds its group's C / groups inputs
  long long spec_plane = 0;
  int groups = 1;
  DevBuf fbias, skip_total;
  ConvW skip, fc1, fc2;
};

}  // namespace

struct ace_csfno {
  ace_csfno_config cfg;
  ace_sht_plan* outer;
  ace_sht_plan* inner;
  long long HW;
  int Ctot, E2, Ep;  // concat channels; context channels (noise + pos) and their padded count
  ConvW enc0, enc1, dec0, dec1;
  DevBuf pos;
  ClnW nbs;  // norm_big_skip
  std::vector<CBlockW> blocks;
  std::map<std::string, bool> params;
  bool finalized = false;

  int wsB = 0;
  DevBuf xin, hcat, e1, hP, xn, rr, x1, c1, c2, g, g2, T, tP, tn, hmid, d1, ctx, sb0;
  DevBuf ctxP, musr;  // tensor-core ConditionalLayerNorm: context as K-major planes [B][HW][2 Ep]; per-pixel {mean, rstd}
  long long p_ctxP = 0;
  long long p_xin, p_hcat, p_act, p_x1, p_c1, p_c2, p_g, p_g2, p_hmid;
  RtBuf rt_in, rt_out;  // filter_residual on the big skip (in_chans wide) / filter_output (out_chans wide)
  DevBuf xrt, yP;       // planes [B][in_chans][HW]: filtered input ahead of norm_big_skip; [B][out_chans][HW]: unfiltered output
  long long p_yP = 0;
  DevBuf gm_min, gm_max, gm_stats;  // clip_latent_global_means: envelope [C] each; per-(sample, channel) {sum, sum of squares} (double)
};

namespace {

void declare_cln(std::map<std::string, bool>& p, const std::string& pre, const ace_csfno_config& c) {
  if (c.embed_dim_scalar > 0) p[pre + "W_scale.weight"] = p[pre + "W_scale.bias"] = p[pre + "W_bias.weight"] = p[pre + "W_bias.bias"] = false;
  if (c.embed_dim_labels > 0)
    p[pre + "W_scale_labels.weight"] = p[pre + "W_scale_labels.bias"] = p[pre + "W_bias_labels.weight"] = p[pre + "W_bias_labels.bias"] = false;
  if (c.embed_dim_noise > 0) p[pre + "W_scale_2d.weight"] = p[pre + "W_bias_2d.weight"] = false;
  if (c.embed_dim_pos > 0) p[pre + "W_scale_pos.weight"] = p[pre + "W_bias_pos.weight"] = false;
  if (c.affine_norms) p[pre + "norm.weight"] = p[pre + "norm.bias"] = false;
}

void declare_params(ace_csfno& n) {
  auto& p = n.params;
  const ace_csfno_config& c = n.cfg;
  if (c.pos_embed) p["pos_embed"] = false;
  p["encoder.0.weight"] = p["encoder.0.bias"] = p["encoder.2.weight"] = false;
  for (int i = 0; i < c.num_layers; ++i) {
    std::string b = "blocks." + std::to_string(i) + ".";
    declare_cln(p, b + "norm0.", c);
    declare_cln(p, b + "norm1.", c);
    p[b + "filter.filter.weight"] = p[b + "filter.filter.bias"] = false;
    p[b + "inner_skip.weight"] = p[b + "inner_skip.bias"] = false;
    p[b + "mlp.fwd.0.weight"] = p[b + "mlp.fwd.0.bias"] = p[b + "mlp.fwd.2.weight"] = p[b + "mlp.fwd.2.bias"] = false;
  }
  p["decoder.0.weight"] = p["decoder.0.bias"] = p["decoder.2.weight"] = false;
  if (c.big_skip && c.normalize_big_skip) declare_cln(p, "norm_big_skip.", c);
  if (c.clip_latent_global_means) p["_gm_min"] = p["_gm_max"] = false;
}

// clip_latent_global_means, eval branch (sfnonet.py:803-812): x[b][c] += clamp(mean, lo[c], hi[c]) - mean with mean the plain
// spatial mean of channel c of sample b (from the producing GEMM's row statistics); nothing happens while any hi[] is non-finite.
// grid (chunks of the plane, B * C); a block whose channel needs no shift returns before touching the plane.
__global__ void __launch_bounds__(256) clip_latent_means_kernel(const double* __restrict__ stats, const float* __restrict__ lo,
                                                               const float* __restrict__ hi, int C, long long HW, bf16* __restrict__ x,
                                                               long long plane, long long x_b) {
  int ok = 1;
  for (int i = threadIdx.x; i < C; i += blockDim.x) ok &= isfinite(hi[i]) ? 1 : 0;
  if (!__syncthreads_and(ok)) return;
  const int b = blockIdx.y / C, c = blockIdx.y % C;
  const float mean = (float)(stats[((long long)b * C + c) * 2] / (double)HW);
  const float shift = fminf(fmaxf(mean, lo[c]), hi[c]) - mean;
  if (shift == 0.f) return;
  bf16* p = x + (long long)b * x_b + (long long)c * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    const float v = __bfloat162float(p[i]) + __bfloat162float(p[i + plane]) + shift;
    bf16 h, l;
    split_bf16(v, h, l);
    p[i] = h;
    p[i + plane] = l;
  }
}

// returns true if `rest` (the key below the norm's prefix) named one of this norm's parameters
bool set_cln_param(ClnW& w, const ace_csfno_config& c, const std::string& rest, const float* src, long long numel, const char* name, cudaStream_t s) {
  auto take = [&](DevBuf& d, long long expect) {
    ACE_REQUIRE(numel == expect, "%s: expected %lld elements, got %lld", name, expect, numel);
    copy_f32(d, src, numel, s);
    return true;
  };
  const long long C = w.C;
  if (rest == "W_scale.weight") return take(w.Ws, C * c.embed_dim_scalar);
  if (rest == "W_scale.bias") return take(w.bs, C);
  if (rest == "W_bias.weight") return take(w.Wb, C * c.embed_dim_scalar);
  if (rest == "W_bias.bias") return take(w.bb, C);
  if (rest == "W_scale_labels.weight") return take(w.Wsl, C * c.embed_dim_labels);
  if (rest == "W_scale_labels.bias") return take(w.bsl, C);
  if (rest == "W_bias_labels.weight") return take(w.Wbl, C * c.embed_dim_labels);
  if (rest == "W_bias_labels.bias") return take(w.bbl, C);
  if (rest == "W_scale_2d.weight") return take(w.ws_n, C * c.embed_dim_noise);
  if (rest == "W_bias_2d.weight") return take(w.wb_n, C * c.embed_dim_noise);
  if (rest == "W_scale_pos.weight") return take(w.ws_p, C * c.embed_dim_pos);
  if (rest == "W_bias_pos.weight") return take(w.wb_p, C * c.embed_dim_pos);
  if (rest == "norm.weight") return take(w.lnw, C);
  if (rest == "norm.bias") return take(w.lnb, C);
  return false;
}

void finalize_cln(ace_csfno& n, ClnW& w, cudaStream_t s) {
  if (n.Ep == 0) return;
  w.w2.ensure((size_t)w.C * n.Ep * 2 * sizeof(float));
  launch_build_cln_w2(w.ws_n.as<float>(), w.wb_n.as<float>(), n.cfg.embed_dim_noise, w.ws_p.as<float>(), w.wb_p.as<float>(), n.cfg.embed_dim_pos,
                      w.C, n.Ep, w.w2.as<float>(), s);
  if (n.Ep % 32 == 0) {
    w.wP_plane = (long long)w.C * 2 * n.Ep;
    w.wP.ensure(2 * (size_t)w.wP_plane * sizeof(bf16));
    launch_cln_w_planes(w.w2.as<float>(), w.C, n.Ep, w.wP.as<bf16>(), w.wP_plane, s);
  }
}

void ensure_ws(ace_csfno& n, int B) {
  if (B <= n.wsB) return;
  const ace_csfno_config& c = n.cfg;
  const int C = c.embed_dim;
  const long long HW = n.HW;
  const ace_sht_plan& p = *n.inner;
  n.p_xin = (long long)B * c.in_chans * HW;
  n.p_hcat = (long long)B * n.Ctot * HW;
  n.p_act = (long long)B * C * HW;
  n.p_x1 = B * p.x1_elems(C);
  n.p_c1 = B * p.c1_elems(C);
  n.p_c2 = B * p.c2_elems(C);
  n.p_g = B * p.g_elems(C);
  n.p_g2 = B * p.g2_elems(C);
  n.p_hmid = (long long)B * c.mlp_hidden * HW;
  const size_t e = sizeof(bf16);
  n.xin.ensure(2 * (size_t)n.p_xin * e);
  n.hcat.ensure(2 * (size_t)n.p_hcat * e);
  for (DevBuf* d : {&n.e1, &n.hP, &n.xn, &n.rr, &n.tP, &n.tn, &n.d1}) d->ensure(2 * (size_t)n.p_act * e);
  n.x1.ensure(2 * (size_t)n.p_x1 * e);
  n.c1.ensure(2 * (size_t)n.p_c1 * e);
  n.c2.ensure(2 * (size_t)n.p_c2 * e);
  n.g.ensure(2 * (size_t)n.p_g * e);
  n.g2.ensure(2 * (size_t)n.p_g2 * e);
  n.T.ensure((size_t)n.p_act * sizeof(float));
  n.hmid.ensure(2 * (size_t)n.p_hmid * e);
  if (n.Ep > 0) n.ctx.ensure((size_t)B * n.Ep * HW * sizeof(float));
  if (n.Ep > 0 && n.Ep % 32 == 0) {
    n.p_ctxP = (long long)B * HW * 2 * n.Ep;
    n.ctxP.ensure(2 * (size_t)n.p_ctxP * sizeof(bf16));
    n.musr.ensure((size_t)B * HW * 2 * sizeof(float));
  }
  if (c.big_skip && c.filter_residual) {
    n.rt_in.ensure(*n.outer, c.in_chans, B);
    if (c.normalize_big_skip) n.xrt.ensure(2 * (size_t)n.p_xin * e);
  }
  if (c.filter_output) {
    n.rt_out.ensure(*n.outer, c.out_chans, B);
    n.p_yP = (long long)B * c.out_chans * HW;
    n.yP.ensure(2 * (size_t)n.p_yP * e);
  }
  n.sb0.ensure((size_t)B * std::max(C, c.in_chans) * 2 * sizeof(float));
  if (c.clip_latent_global_means) n.gm_stats.ensure((size_t)B * C * 2 * sizeof(double));
  n.wsB = B;
}

// The tensor-core ConditionalLayerNorm (GemmOp::cln) is the default wherever the context is padded to a multiple of 32 (option
// cln_gemm: 1.6 % of the ERA5-baseline forward at Ep = 32) and the only fast path beyond 64 context channels,
// where the 2 Ep FMAs per element of the streaming kernel would cost more than the rest of the block
bool cln_on_tensor_cores(const ace_csfno& n) {
  if (options().force_simt || n.Ep <= 0 || n.Ep % 32 != 0) return false;
  return n.Ep > 64 || options().cln_gemm;
}

void run_cln(ace_csfno& n, const ClnW& w, const bf16* x, long long x_plane, long long x_b, const float* scalar, const float* labels, int B, bf16* out,
             long long o_plane, long long o_b, cudaStream_t s) {
  const ace_csfno_config& c = n.cfg;
  const float* sb0 = nullptr;
  if (c.embed_dim_scalar > 0 || c.embed_dim_labels > 0) {
    launch_cln_vector_terms(scalar, c.embed_dim_scalar, labels, c.embed_dim_labels, w.Ws.as<float>(), w.bs.as<float>(), w.Wb.as<float>(),
                            w.bb.as<float>(), w.Wsl.as<float>(), w.bsl.as<float>(), w.Wbl.as<float>(), w.bbl.as<float>(), B, w.C,
                            n.sb0.as<float>(), s);
    sb0 = n.sb0.as<float>();
  }
  // tensor-core path (GemmOp::cln): a statistics pass, then ONE GEMM whose two accumulators are the scale / bias modulation
  // (K = 2 Ep) and whose epilogue normalises x and applies them (the streaming kernel below is bound by its 2 Ep FMAs per
  // element on the FMA pipe: 200 us per norm at C = 512, 1 degree)
  if (cln_on_tensor_cores(n) && w.C >= 128 && w.wP.p && n.HW % 4 == 0) {
    GemmOp op = make_gemm_op("cond_layer_norm");
    op.cln = 1;
    op.M = w.C;
    op.N = (int)n.HW;
    op.K = 2 * n.Ep;
    op.k_split = n.Ep;
    op.Z2 = B;
    op.A = {w.wP.as<bf16>(), w.wP_plane, 2LL * n.Ep, 1, 0, 0};
    op.B = {n.ctxP.as<bf16>(), n.p_ctxP, 2LL * n.Ep, 1, 0, n.HW * 2LL * n.Ep};
    op.epi.flags = EPI_RES_PLANES | EPI_OUT_PLANES;
    op.epi.res = x;
    op.epi.res_plane = x_plane;
    op.epi.res_z2 = x_b;
    op.epi.res_m0 = n.HW;
    op.epi.res_n = 1;
    op.epi.out = out;
    op.epi.out_plane = o_plane;
    op.epi.o_z2 = o_b;
    op.epi.o_m0 = n.HW;
    op.epi.o_n = 1;
    op.epi.cln_musr = reinterpret_cast<const float2*>(n.musr.as<float>());
    op.epi.cln_musr_z2 = n.HW;
    op.epi.cln_lnw = c.affine_norms ? w.lnw.as<float>() : nullptr;
    op.epi.cln_lnb = c.affine_norms ? w.lnb.as<float>() : nullptr;
    op.epi.cln_sb0 = sb0;
    op.epi.cln_sb0_z2 = w.C;
    launch_cln_stats(x, x_plane, x_b, B, w.C, n.HW, c.norm_eps, n.musr.as<float>(), s);
    {
      ProfileScope prof("cond_layer_norm", s);
      if (run_gemm_cln(op, s)) return;
    }
  }
  launch_cond_layer_norm(x, x_plane, x_b, B, w.C, n.HW, c.affine_norms ? w.lnw.as<float>() : nullptr, c.affine_norms ? w.lnb.as<float>() : nullptr,
                         sb0, n.Ep > 0 ? w.w2.as<float>() : nullptr, n.Ep > 0 ? n.ctx.as<float>() : nullptr, n.Ep, c.norm_eps, out, o_plane,
                         o_b, s);
}

// itrans_up(trans_down(x)) for `Cc` channels of split planes -> split planes `out`, or fp32 `outf` when given
void round_trip_outer(ace_csfno& n, RtBuf& r, const bf16* x, long long x_plane, long long x_b, int Cc, int B, bf16* out, long long o_plane,
                      long long o_b, float* outf, long long f_b, cudaStream_t s) {
  const ace_sht_plan& p = *n.outer;
  run_gemm(sht_op_dft_fwd(p, x, x_plane, x_b, Cc, B, r.x1.as<bf16>(), r.p_x1), s);
  run_gemm(sht_op_legendre_fwd(p, r.x1.as<bf16>(), r.p_x1, Cc, B, r.c1.as<bf16>(), r.p_c1), s);
  run_gemm(sht_op_legendre_inv_from_c1(p, r.c1.as<bf16>(), r.p_c1, Cc, B, r.g.as<bf16>(), r.p_g), s);
  if (outf) run_gemm(sht_op_dft_inv(p, r.g.as<bf16>(), r.p_g, Cc, B, outf, f_b), s);
  else run_gemm(sht_op_dft_inv_planes(p, r.g.as<bf16>(), r.p_g, Cc, B, out, o_plane, o_b), s);
}

void forward(ace_csfno& n, const float* x, const float* scalar, const float* labels, const float* noise, const float* posctx, float* y, int B,
             cudaStream_t s) {
  const ace_csfno_config& c = n.cfg;
  const int C = c.embed_dim, Cin = c.in_chans, NL = c.num_layers;
  const long long HW = n.HW;
  ensure_ws(n, B);
  bf16* hcat = n.hcat.as<bf16>();
  const long long P_hcat = n.p_hcat, P_act = n.p_act, P_x1 = n.p_x1, P_c1 = n.p_c1, P_c2 = n.p_c2, P_g = n.p_g, P_hmid = n.p_hmid,
                  P_xin = n.p_xin;
  const long long act_b = (long long)C * HW, cat_b = (long long)n.Ctot * HW, in_b = (long long)Cin * HW;

  if (n.Ep > 0) launch_concat_ctx(noise, c.embed_dim_noise, posctx, c.embed_dim_pos, B, HW, n.Ep, n.ctx.as<float>(), s);
  if (cln_on_tensor_cores(n))
    launch_cln_ctx_planes(n.ctx.as<float>(), B, n.Ep, HW, n.ctxP.as<bf16>(), n.p_ctxP, s);

  // network input -> split planes; the big skip is the (optionally conditionally normalised) input, stored in the tail channels
  // of the concat buffer (sfnonet.py:775-778)
  const bf16* enc_in;
  long long enc_plane, enc_b;
  bf16* const skip_dst = hcat + (long long)C * HW;
  const bool rt_skip = c.big_skip && c.filter_residual;  // residual = itrans_up(trans_down(x)); the encoder still reads x itself
  if (c.big_skip && !c.normalize_big_skip && !rt_skip) {
    launch_norm_split(x, B, Cin, HW, nullptr, nullptr, nullptr, 0.f, skip_dst, P_hcat, cat_b, HW, s);
    enc_in = skip_dst;
    enc_plane = P_hcat;
    enc_b = cat_b;
  } else {
    launch_norm_split(x, B, Cin, HW, nullptr, nullptr, nullptr, 0.f, n.xin.as<bf16>(), P_xin, in_b, HW, s);
    enc_in = n.xin.as<bf16>();
    enc_plane = P_xin;
    enc_b = in_b;
    if (rt_skip && !c.normalize_big_skip) {
      round_trip_outer(n, n.rt_in, enc_in, P_xin, in_b, Cin, B, skip_dst, P_hcat, cat_b, nullptr, 0, s);
    } else if (c.big_skip) {
      const bf16* src = enc_in;
      if (rt_skip) {
        round_trip_outer(n, n.rt_in, enc_in, P_xin, in_b, Cin, B, n.xrt.as<bf16>(), P_xin, in_b, nullptr, 0, s);
        src = n.xrt.as<bf16>();
      }
      run_cln(n, n.nbs, src, P_xin, in_b, scalar, labels, B, skip_dst, P_hcat, cat_b, s);
    }
  }

  // encoder: Conv(Cin->C)+bias, GELU, Conv(C->C) ; + pos_embed          (sfnonet.py:613-640, :785-786)
  {
    GemmOp op = conv_op("encoder.0", enc_in, enc_plane, enc_b, HW, B, n.enc0, Cin);
    op.epi.flags |= EPI_GELU;
    out_planes(op, n.e1.as<bf16>(), P_act, act_b, HW);
    run_gemm(op, s);
  }
  bf16* hP = n.hP.as<bf16>();
  {
    GemmOp op = conv_op("encoder.2", n.e1.as<bf16>(), P_act, act_b, HW, B, n.enc1, C);
    if (c.pos_embed) add_f32(op, n.pos.as<float>(), 0, HW);
    out_planes(op, hP, P_act, act_b, HW);
    if (c.clip_latent_global_means) {
      ACE_CHECK_CUDA(cudaMemsetAsync(n.gm_stats.p, 0, (size_t)B * C * 2 * sizeof(double), s));
      row_stats(op, n.gm_stats.as<double>(), C);
    }
    run_gemm(op, s);
    if (c.clip_latent_global_means) {
      ACE_REQUIRE((long long)B * C <= 65535, "clip_latent_global_means: batch * embed_dim = %lld exceeds the grid limit 65535", (long long)B * C);
      ProfileScope prof("clip_latent_means", s);
      dim3 grid((unsigned)std::min<long long>((HW + 255) / 256, 64), (unsigned)(B * C));
      clip_latent_means_kernel<<<grid, 256, 0, s>>>(n.gm_stats.as<double>(), n.gm_min.as<float>(), n.gm_max.as<float>(), C, HW, hP, P_act, act_b);
      after_launch("clip_latent_means");
    }
  }

  for (int i = 0; i < NL; ++i) {
    CBlockW& w = n.blocks[i];
    const ace_sht_plan& pf = (i == 0) ? *n.outer : *n.inner;
    const ace_sht_plan& pi = (i == NL - 1) ? *n.outer : *n.inner;
    const bool round_trip = c.filter_residual || pf.table_id != pi.table_id;  // s2convolutions.py:195-199
    bf16* xn = n.xn.as<bf16>();
    run_cln(n, w.n0, hP, P_act, act_b, scalar, labels, B, xn, P_act, act_b, s);
    run_gemm(sht_op_dft_fwd(pf, xn, P_act, act_b, C, B, n.x1.as<bf16>(), P_x1), s);
    run_gemm(sht_op_legendre_fwd(pf, n.x1.as<bf16>(), P_x1, C, B, n.c1.as<bf16>(), P_c1), s);
    const bf16* resid = xn;
    if (round_trip) {
      run_gemm(sht_op_legendre_inv_from_c1(pi, n.c1.as<bf16>(), P_c1, C, B, n.g.as<bf16>(), P_g), s);
      run_gemm(sht_op_dft_inv_planes(pi, n.g.as<bf16>(), P_g, C, B, n.rr.as<bf16>(), P_act, act_b), s);
      resid = n.rr.as<bf16>();
    }
    // complex GEMM per degree l over the orders m <= l (s2convolutions.py:118-136)
    if (w.groups > 1)
      run_gemm(dhconv_grouped_op(n.c1.as<bf16>(), P_c1, w.spec.as<bf16>(), w.spec_plane, pf, C, w.groups, B, n.c2.as<bf16>(), P_c2), s);
    else
      run_gemm(dhconv_op(n.c1.as<bf16>(), P_c1, w.spec.as<bf16>(), w.spec_plane, pf, C, C, B, n.c2.as<bf16>(), P_c2), s);
    if (options().inv2) {
      run_gemm(sht_op_legendre_inv2(pi, n.c2.as<bf16>(), P_c2, C, B, n.g2.as<bf16>(), n.p_g2), s);
      run_gemm(sht_op_dft_inv2(pi, n.g2.as<bf16>(), n.p_g2, C, B, n.T.as<float>(), act_b), s);
    } else {
      run_gemm(sht_op_legendre_inv(pi, n.c2.as<bf16>(), P_c2, C, B, n.g.as<bf16>(), P_g), s);
      run_gemm(sht_op_dft_inv(pi, n.g.as<bf16>(), P_g, C, B, n.T.as<float>(), act_b), s);
    }
    // x = GELU(filter(xn) + b_filter + inner_skip(residual))   (sfnonet.py:387-399)
    {
      GemmOp op = conv_op("inner_skip", resid, P_act, act_b, HW, B, w.skip, C);
      op.epi.row_bias = w.skip_total.as<float>();
      op.epi.flags |= EPI_GELU;
      add_f32(op, n.T.as<float>(), act_b, HW);
      out_planes(op, n.tP.as<bf16>(), P_act, act_b, HW);
      run_gemm(op, s);
    }
    run_cln(n, w.n1, n.tP.as<bf16>(), P_act, act_b, scalar, labels, B, n.tn.as<bf16>(), P_act, act_b, s);
    // MLP + identity outer skip of the block's residual   (sfnonet.py:413-434)
    {
      GemmOp op = conv_op("mlp.fc1", n.tn.as<bf16>(), P_act, act_b, HW, B, w.fc1, C);
      op.epi.flags |= EPI_GELU;
      out_planes(op, n.hmid.as<bf16>(), P_hmid, (long long)c.mlp_hidden * HW, HW);
      run_gemm(op, s);
    }
    {
      GemmOp op = conv_op("mlp.fc2", n.hmid.as<bf16>(), P_hmid, (long long)c.mlp_hidden * HW, HW, B, w.fc2, c.mlp_hidden);
      op.epi.flags |= EPI_RES_PLANES;
      op.epi.res = resid;
      op.epi.res_plane = P_act;
      op.epi.res_z2 = act_b;
      op.epi.res_m0 = HW;
      op.epi.res_n = 1;
      if (i == NL - 1) out_planes(op, hcat, P_hcat, cat_b, HW);
      else out_planes(op, hP, P_act, act_b, HW);
      run_gemm(op, s);
    }
  }

  // decoder on cat(x, big skip)   (sfnonet.py:813-822)
  {
    GemmOp op = conv_op("decoder.0", hcat, P_hcat, cat_b, HW, B, n.dec0, c.big_skip ? n.Ctot : C);
    op.epi.flags |= EPI_GELU;
    out_planes(op, n.d1.as<bf16>(), P_act, act_b, HW);
    run_gemm(op, s);
  }
  {
    const long long out_b = (long long)c.out_chans * HW;
    GemmOp op = conv_op("decoder.2", n.d1.as<bf16>(), P_act, act_b, HW, B, n.dec1, C);
    if (c.filter_output) out_planes(op, n.yP.as<bf16>(), n.p_yP, out_b, HW);
    else out_f32(op, y, out_b, HW);
    run_gemm(op, s);
    if (c.filter_output) round_trip_outer(n, n.rt_out, n.yP.as<bf16>(), n.p_yP, out_b, c.out_chans, B, nullptr, 0, 0, y, out_b, s);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------ C ABI

extern "C" int ace_csfno_create(const ace_csfno_config* cfg, ace_sht_plan* plan_outer, ace_sht_plan* plan_inner, ace_csfno** out) {
  ACE_API_BEGIN
  ACE_REQUIRE(cfg && plan_outer && plan_inner && out, "ace_csfno_create: null argument");
  const ace_csfno_config& c = *cfg;
  ACE_REQUIRE(c.img_h > 0 && c.img_w > 0 && c.in_chans > 0 && c.out_chans > 0 && c.embed_dim > 0 && c.num_layers > 0 && c.mlp_hidden > 0,
              "ace_csfno_create: non-positive size");
  ACE_REQUIRE(c.embed_dim_scalar >= 0 && c.embed_dim_labels >= 0 && c.embed_dim_noise >= 0 && c.embed_dim_pos >= 0,
              "ace_csfno_create: negative context width");
  const int E2 = c.embed_dim_noise + c.embed_dim_pos;
  const int Ep = cln_padded_context(E2);
  for (ace_sht_plan* p : {plan_outer, plan_inner})
    ACE_REQUIRE(p->K == c.img_h && p->W == c.img_w && p->L == c.lmax && p->M == c.mmax,
                "ace_csfno_create: plan (%d,%d,%d,%d) does not match config (%d,%d,%d,%d)", p->K, p->W, p->L, p->M, c.img_h,
                c.img_w, c.lmax, c.mmax);
  ace_csfno* n = new ace_csfno();
  try {
    n->cfg = c;
    n->outer = plan_outer;
    n->inner = plan_inner;
    n->HW = (long long)c.img_h * c.img_w;
    n->Ctot = c.embed_dim + c.in_chans;
    n->E2 = E2;
    n->Ep = Ep;
    const int C = c.embed_dim;
    n->enc0.init(C, c.in_chans, true);
    n->enc1.init(C, C, false);
    n->dec0.init(C, c.big_skip ? n->Ctot : C, true);
    n->dec1.init(c.out_chans, C, false);
    n->nbs.C = c.in_chans;
    n->blocks.resize(c.num_layers);
    for (CBlockW& b : n->blocks) {
      b.n0.C = b.n1.C = C;
      b.skip.init(C, C, true);
      b.fc1.init(c.mlp_hidden, C, true);
      b.fc2.init(C, c.mlp_hidden, true);
      b.fbias.ensure((size_t)C * sizeof(float));
      b.skip_total.ensure((size_t)C * sizeof(float));
    }
    declare_params(*n);
  } catch (...) {
    delete n;
    throw;
  }
  *out = n;
  ACE_API_END
}

extern "C" void ace_csfno_destroy(ace_csfno* net) { delete net; }

extern "C" int ace_csfno_query(ace_csfno* net, int* in_chans, int* out_chans, long long* hw) {
  ACE_API_BEGIN
  ACE_REQUIRE(net && in_chans && out_chans && hw, "ace_csfno_query: null argument");
  *in_chans = net->cfg.in_chans;
  *out_chans = net->cfg.out_chans;
  *hw = net->HW;
  ACE_API_END
}

extern "C" int ace_csfno_set_param(ace_csfno* net, const char* name, const float* data_dev, long long numel, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(net && name && data_dev, "ace_csfno_set_param: null argument");
  ace_csfno& n = *net;
  cudaStream_t s = (cudaStream_t)stream;
  const ace_csfno_config& c = n.cfg;
  const int C = c.embed_dim;
  std::string nm(name);
  auto it = n.params.find(nm);
  ACE_REQUIRE(it != n.params.end(), "ace_csfno_set_param: unexpected parameter '%s' for this configuration", name);
  if (nm == "pos_embed") {
    ACE_REQUIRE(numel == (long long)C * n.HW, "pos_embed: expected %lld elements, got %lld", (long long)C * n.HW, numel);
    copy_f32(n.pos, data_dev, numel, s);
  } else if (nm == "_gm_min" || nm == "_gm_max") {
    ACE_REQUIRE(numel == C, "%s: expected %d elements, got %lld", name, C, numel);
    copy_f32(nm == "_gm_min" ? n.gm_min : n.gm_max, data_dev, numel, s);
  } else if (nm == "encoder.0.weight") set_conv_w(n.enc0, data_dev, numel, name, s);
  else if (nm == "encoder.0.bias") set_conv_b(n.enc0, data_dev, numel, name, s);
  else if (nm == "encoder.2.weight") set_conv_w(n.enc1, data_dev, numel, name, s);
  else if (nm == "decoder.0.weight") set_conv_w(n.dec0, data_dev, numel, name, s);
  else if (nm == "decoder.0.bias") set_conv_b(n.dec0, data_dev, numel, name, s);
  else if (nm == "decoder.2.weight") set_conv_w(n.dec1, data_dev, numel, name, s);
  else if (nm.rfind("norm_big_skip.", 0) == 0) {
    ACE_REQUIRE(set_cln_param(n.nbs, c, nm.substr(14), data_dev, numel, name, s), "ace_csfno_set_param: unhandled parameter '%s'", name);
  } else {
    int bi = -1, consumed = 0;
    ACE_REQUIRE(sscanf(name, "blocks.%d.%n", &bi, &consumed) == 1 && bi >= 0 && bi < c.num_layers, "bad block index in '%s'", name);
    CBlockW& b = n.blocks[bi];
    std::string rest(name + consumed);
    auto vecC = [&](DevBuf& d) {
      ACE_REQUIRE(numel == C, "%s: expected %d elements, got %lld", name, C, numel);
      copy_f32(d, data_dev, numel, s);
    };
    if (rest.rfind("norm0.", 0) == 0) {
      ACE_REQUIRE(set_cln_param(b.n0, c, rest.substr(6), data_dev, numel, name, s), "ace_csfno_set_param: unhandled parameter '%s'", name);
    } else if (rest.rfind("norm1.", 0) == 0) {
      ACE_REQUIRE(set_cln_param(b.n1, c, rest.substr(6), data_dev, numel, name, s), "ace_csfno_set_param: unhandled parameter '%s'", name);
    } else if (rest == "filter.filter.bias") vecC(b.fbias);
    else if (rest == "filter.filter.weight") {
      // [G][L][O / G][I / G][2]  (s2convolutions.py:229-236); G = 1: the dense operator, G > 1: the diagonal blocks of a grouped one
      const long long dense = (long long)C * C * c.lmax * 2;
      const int G = (numel > 0 && dense % numel == 0) ? (int)(dense / numel) : 0;
      ACE_REQUIRE(G >= 1 && C % G == 0, "%s: expected %lld elements (or 1/G of it for a grouped filter), got %lld", name, dense, numel);
      if (G == 1) {
        const int Cp = (int)round_up(C, 8);
        b.spec_plane = (long long)c.lmax * 2 * C * Cp;
        b.spec.ensure(2 * (size_t)b.spec_plane * sizeof(bf16));
        launch_prep_dhconv_cplx_strided(data_dev, C, C, c.lmax, (long long)C * C, C, 1, Cp, b.spec.as<bf16>(), b.spec_plane, s);
      } else {
        const int cg = C / G, cgp = (int)round_up(cg, 8);
        ACE_REQUIRE(cg == 64 || cg == 128, "%s: grouped filters are multiplied natively for 64 or 128 channels per group (got %d); fold "
                    "other group sizes into the dense operator", name, cg);
        b.spec_plane = (long long)c.lmax * 2 * C * cgp;
        b.spec.release();  // the grouped planes are 1/G of the dense ones: do not keep a dense-sized buffer around
        b.spec.ensure(2 * (size_t)b.spec_plane * sizeof(bf16));
        launch_prep_dhconv_grouped(data_dev, G, cg, c.lmax, cgp, b.spec.as<bf16>(), b.spec_plane, s);
      }
      b.groups = G;
    } else if (rest == "inner_skip.weight") set_conv_w(b.skip, data_dev, numel, name, s);
    else if (rest == "inner_skip.bias") set_conv_b(b.skip, data_dev, numel, name, s);
    else if (rest == "mlp.fwd.0.weight") set_conv_w(b.fc1, data_dev, numel, name, s);
    else if (rest == "mlp.fwd.0.bias") set_conv_b(b.fc1, data_dev, numel, name, s);
    else if (rest == "mlp.fwd.2.weight") set_conv_w(b.fc2, data_dev, numel, name, s);
    else if (rest == "mlp.fwd.2.bias") set_conv_b(b.fc2, data_dev, numel, name, s);
    else ACE_REQUIRE(false, "ace_csfno_set_param: unhandled parameter '%s'", name);
  }
  it->second = true;
  n.finalized = false;
  ACE_API_END
}

extern "C" int ace_csfno_finalize(ace_csfno* net, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(net, "ace_csfno_finalize: null argument");
  for (auto& kv : net->params)
    if (!kv.second) throw Error(ACE_ERR_STATE, "ace_csfno_finalize: parameter '" + kv.first + "' has not been set");
  cudaStream_t s = (cudaStream_t)stream;
  ace_csfno& n = *net;
  if (n.cfg.big_skip && n.cfg.normalize_big_skip) finalize_cln(n, n.nbs, s);
  for (CBlockW& b : n.blocks) {
    finalize_cln(n, b.n0, s);
    finalize_cln(n, b.n1, s);
    launch_vec_add(b.skip.bias.as<float>(), b.fbias.as<float>(), b.skip_total.as<float>(), n.cfg.embed_dim, s);
  }
  n.finalized = true;
  ACE_API_END
}

extern "C" int ace_csfno_forward(ace_csfno* net, const float* x_dev, const float* scalar_dev, const float* labels_dev, const float* noise_dev,
                                 const float* pos_dev, float* y_dev, int batch, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(net && x_dev && y_dev, "ace_csfno_forward: null argument");
  ACE_REQUIRE(batch > 0, "ace_csfno_forward: batch must be positive");
  if (!net->finalized) throw Error(ACE_ERR_STATE, "ace_csfno_forward: call ace_csfno_finalize first");
  const ace_csfno_config& c = net->cfg;
  ACE_REQUIRE(scalar_dev || c.embed_dim_scalar == 0, "ace_csfno_forward: embedding_scalar must be provided");
  ACE_REQUIRE(labels_dev || c.embed_dim_labels == 0, "ace_csfno_forward: labels must be provided");
  ACE_REQUIRE(noise_dev || c.embed_dim_noise == 0, "ace_csfno_forward: noise must be provided");
  ACE_REQUIRE(pos_dev || c.embed_dim_pos == 0, "ace_csfno_forward: embedding_pos must be provided");
  forward(*net, x_dev, scalar_dev, labels_dev, noise_dev, pos_dev, y_dev, batch, (cudaStream_t)stream);
  ACE_API_END
}

// ------------------------------------------------------------------------------------ label conditioning of the wrapper
// fme/ace/registry/stochastic_sfno.py:152-165: labels -> Linear(n_labels, label_embed_dim); positional context per sample =
// pos_embed + einsum("bl,lpxy->bpxy", labels, label_pos_embed).  Tiny streaming kernels (a few kB .. MB per step).
namespace {
__global__ void small_linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int n_in, int n_out,
                                    long long total, float* __restrict__ y) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / n_out;
    const int o = (int)(idx % n_out);
    float acc = 0.f;
    for (int i = 0; i < n_in; ++i) acc = fmaf(x[b * n_in + i], w[(long long)o * n_in + i], acc);
    y[idx] = acc + (bias ? bias[o] : 0.f);
  }
}
__global__ void label_pos_embed_kernel(const float* __restrict__ base, const float* __restrict__ labels, const float* __restrict__ lpe, int n_labels,
                                       long long phw, long long total, float* __restrict__ out) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / phw, r = idx % phw;
    float acc = 0.f;
    for (int l = 0; l < n_labels; ++l) acc = fmaf(labels[b * n_labels + l], lpe[(long long)l * phw + r], acc);
    out[idx] = base[r] + acc;
  }
}
}  // namespace

extern "C" int ace_label_embed(const float* labels_dev, const float* weight_dev, const float* bias_dev, int batch, int n_labels, int embed_dim,
                               float* out_dev, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(labels_dev && weight_dev && out_dev && batch > 0 && n_labels > 0 && embed_dim > 0, "ace_label_embed: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = (long long)batch * embed_dim;
  ProfileScope prof("label_embed", s);
  small_linear_kernel<<<(unsigned)std::min<long long>((total + 127) / 128, 148 * 8), 128, 0, s>>>(labels_dev, weight_dev, bias_dev, n_labels, embed_dim,
                                                                                             total, out_dev);
  after_launch("label_embed");
  ACE_API_END
}

extern "C" int ace_label_pos_embed(const float* pos_dev, const float* labels_dev, const float* label_pos_dev, int batch, int n_labels, long long phw,
                                   float* out_dev, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(pos_dev && labels_dev && label_pos_dev && out_dev && batch > 0 && n_labels > 0 && phw > 0, "ace_label_pos_embed: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = (long long)batch * phw;
  ProfileScope prof("label_pos_embed", s);
  label_pos_embed_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, s>>>(pos_dev, labels_dev, label_pos_dev, n_labels,
                                                                                                 phw, total, out_dev);
  after_launch("label_pos_embed");
  ACE_API_END
}

// ------------------------------------------------------------------------------------ isotropic noise
// fme/ace/registry/stochastic_sfno.py:21-47: two N(0,1) draws -> a_lm with Im(a_l0) = 0, Re / Im of m > 0 divided by sqrt(2),
// everything scaled by sqrt(4 pi) / lmax (unit pointwise variance), then the inverse SHT.
namespace {
__global__ void isotropic_coeffs_kernel(const float* __restrict__ re, const float* __restrict__ im, int M, float scale, long long total,
                                        float2* __restrict__ out) {
  const float sqrt2 = 1.4142135623730951f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % M);
    float r = re[idx], i = im[idx];
    if (m == 0) {
      i = 0.f;
    } else {
      r /= sqrt2;
      i /= sqrt2;
    }
    out[idx] = make_float2(r * scale, i * scale);
  }
}
}  // namespace

extern "C" int ace_isotropic_noise(ace_sht_plan* plan, const float* real_dev, const float* imag_dev, float* coeffs_scratch_dev, float* noise_dev,
                                   long long nfields, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(plan && real_dev && imag_dev && coeffs_scratch_dev && noise_dev && nfields > 0, "ace_isotropic_noise: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = nfields * plan->L * plan->M;
  const float scale = (float)(sqrt(4.0 * 3.14159265358979323846) / (double)plan->L);
  {
    ProfileScope prof("isotropic_coeffs", s);
    isotropic_coeffs_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, s>>>(real_dev, imag_dev, plan->M, scale, total,
                                                                                                    reinterpret_cast<float2*>(coeffs_scratch_dev));
    after_launch("isotropic_coeffs");
  }
  return ace_sht_inverse(plan, coeffs_scratch_dev, noise_dev, nfields, stream);
  ACE_API_END
}
