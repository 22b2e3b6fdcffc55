// SphericalFourierNeuralOperatorNet forward (ace_sfno_*): parameter re-layout + kernel orchestration.
//
// Mirrors /root/reference/fme/ace/models/modulus/sfnonet.py:713-749 (net), :217-252 (block),
// s2convolutions.py:162-197 (spectral conv) as a fixed sequence of GemmOps and streaming kernels.
// Per block (C = embed_dim, HW = nlat*nlon):
//
//   hP --G1 DFT (x a0, + 2 pi s0 at m = 0)--> X1 --G2 Legendre--> c1 --G3 dhconv--> c2 --G4 Legendre-->
//   g --G5 iDFT--> T (fp32) ; tP <- GELU(T + (Wskip a0) hP + bskip' + bfilter) [stats1]
//   --fc1 (W1 a1, b1') + GELU--> hmid --fc2 + bias + (a0 hP + s0)--> hP (next block) [stats0 of the next block]
//
// InstanceNorm is "deferred": the epilogue of the GEMM that produces a tensor stores it UN-normalised as
// split-bf16 planes and accumulates per-(sample, channel) sums (fp64 atomics); a tiny per-block prep kernel
// turns the sums into y = a*h + s and folds that affine map into every consumer -- the 1x1 convolution
// weights/bias (per sample), the row scale of the DFT GEMM (the constant s only reaches the m = 0
// coefficient) and the residual term of fc2.  No separate normalisation pass over the activations exists.
#include <map>
#include <string>
#include <vector>

#include "netops.cuh"

using namespace ace;

namespace {

struct BlockW {
  DevBuf g0, b0, g1, b1;  // InstanceNorm affine
  DevBuf spec;            // dhconv: planes [L][2 (re, im)][C][Cp]; diagonal: fp32 [C][C][L][M][2]
  long long spec_plane = 0;
  DevBuf fbias;           // spectral conv bias [C]
  DevBuf skip_total;      // inner_skip.bias + fbias
  ConvW skip, fc1, fc2;
};

}  // namespace

struct ace_sfno {
  ace_sfno_config cfg;
  ace_sht_plan* outer;
  ace_sht_plan* inner;
  long long HW;
  int Ctot;  // channels of the concat buffer: embed_dim + in_chans
  ConvW enc0, enc1, dec0, dec1;
  DevBuf pos;  // fp32 [C][HW]
  std::vector<BlockW> blocks;
  std::map<std::string, bool> params;  // name -> set?
  bool finalized = false;

  // workspace (allocated on first use for a batch size; grows monotonically)
  int wsB = 0;
  DevBuf hcat, e1, hP, xn, x1, c1, c2, g, g2, T, tP, hmid, d1, stats;
  DevBuf skipW, skipB, fc1W, fc1B;   // per-sample convolution parameters with the InstanceNorm folded in
  DevBuf na0, ns0, nsh0, na1, ns1, nsh1;  // per-(sample, channel) a, s, 2*pi*s of norm0 / norm1
  long long p_hcat, p_act, p_x1, p_c1, p_c2, p_g, p_g2, p_hmid, p_skipW, p_fc1W;  // plane offsets (elements)
  // tP and hmid are only ever 1x1-convolution operands (MN-major B: one TMA box row = 64 pixels = 128 bytes of one channel), so
  // their channel pitch is padded to whole 128-byte lines: with the natural pitch H*W*2 B (= 1012.5 lines at 180x360) every odd
  // channel's box rows straddle two L2 lines.  xn / hcat keep the natural pitch.
  long long HWp = 0, p_tp = 0;
  // hP (the block input: DFT operand, inner_skip operand, fc2 residual) carries Kp - nlat pad rows per channel: its channel pitch
  // Kp * nlon is a whole number of 128-byte lines at 180x360 and the forward DFT's rows (channel, latitude) then map onto the X1
  // layout [..][C][Kp] without gaps, i.e. every tile column is stored as one contiguous run (measured: dft_fwd 83 -> 73 us when
  // nlat itself is 184).  The pad rows stay zero from allocation: the convolutions only write the H*W logical columns.
  long long HWq = 0, p_hp = 0;
};

namespace {

void declare_params(ace_sfno& n) {
  auto& p = n.params;
  const ace_sfno_config& c = n.cfg;
  if (c.pos_embed) p["pos_embed"] = false;
  p["encoder.0.weight"] = p["encoder.0.bias"] = p["encoder.2.weight"] = false;
  for (int i = 0; i < c.num_layers; ++i) {
    std::string b = "blocks." + std::to_string(i) + ".";
    if (c.normalization == 1) p[b + "norm0.weight"] = p[b + "norm0.bias"] = p[b + "norm1.weight"] = p[b + "norm1.bias"] = false;
    p[b + "filter.filter.weight"] = p[b + "filter.filter.bias"] = false;
    p[b + "inner_skip.weight"] = p[b + "inner_skip.bias"] = false;
    p[b + "mlp.fwd.0.weight"] = p[b + "mlp.fwd.0.bias"] = p[b + "mlp.fwd.2.weight"] = p[b + "mlp.fwd.2.bias"] = false;
  }
  p["decoder.0.weight"] = p["decoder.0.bias"] = p["decoder.2.weight"] = false;
}

void ensure_ws(ace_sfno& n, int B) {
  if (B <= n.wsB) return;
  const ace_sfno_config& c = n.cfg;
  const int C = c.embed_dim;
  const long long HW = n.HW;
  const ace_sht_plan& p = *n.inner;  // outer has identical L, M, K, W -> identical layouts
  n.p_hcat = (long long)B * n.Ctot * HW;
  n.p_act = (long long)B * C * HW;
  n.p_x1 = B * p.x1_elems(C);
  n.p_c1 = B * p.c1_elems(C);
  n.p_c2 = B * p.c2_elems(C);
  n.p_g = B * p.g_elems(C);
  n.p_g2 = B * p.g2_elems(C);
  n.HWp = round_up(HW, 64);
  n.HWq = (long long)p.Kp * p.W;
  n.p_hp = (long long)B * C * n.HWq;
  n.p_hmid = (long long)B * c.mlp_hidden * n.HWp;
  n.p_tp = (long long)B * C * n.HWp;
  const size_t e = sizeof(bf16);
  n.hcat.ensure(2 * (size_t)n.p_hcat * e);
  n.e1.ensure(2 * (size_t)n.p_act * e);
  n.hP.ensure(2 * (size_t)n.p_hp * e);
  n.xn.ensure(2 * (size_t)n.p_act * e);
  n.x1.ensure(2 * (size_t)n.p_x1 * e);
  n.c1.ensure(2 * (size_t)n.p_c1 * e);
  n.c2.ensure(2 * (size_t)n.p_c2 * e);
  n.g.ensure(2 * (size_t)n.p_g * e);
  n.g2.ensure(2 * (size_t)n.p_g2 * e);
  n.T.ensure((size_t)n.p_act * sizeof(float));
  n.tP.ensure(2 * (size_t)n.p_tp * e);
  {
    const int Cp = (int)round_up(C, 8);
    n.p_skipW = (long long)B * C * Cp;
    n.p_fc1W = (long long)B * c.mlp_hidden * Cp;
    n.skipW.ensure(2 * (size_t)n.p_skipW * e);
    n.fc1W.ensure(2 * (size_t)n.p_fc1W * e);
    n.skipB.ensure((size_t)B * C * sizeof(float));
    n.fc1B.ensure((size_t)B * c.mlp_hidden * sizeof(float));
    for (DevBuf* d : {&n.na0, &n.ns0, &n.nsh0, &n.na1, &n.ns1, &n.nsh1}) d->ensure((size_t)B * C * sizeof(float));
  }
  n.hmid.ensure(2 * (size_t)n.p_hmid * e);
  n.d1.ensure(2 * (size_t)n.p_act * e);
  n.stats.ensure((size_t)2 * c.num_layers * B * C * 2 * sizeof(double));
  n.wsB = B;
}

void forward(ace_sfno& n, const float* x, float* y, int B, cudaStream_t s) {
  const ace_sfno_config& c = n.cfg;
  const int C = c.embed_dim, Cin = c.in_chans, NL = c.num_layers;
  const long long HW = n.HW;
  const bool inorm = c.normalization == 1;
  ensure_ws(n, B);
  // plane offsets are those of the allocated capacity; per-sample strides do not depend on it
  bf16* hcat = n.hcat.as<bf16>();
  const long long P_hcat = n.p_hcat, P_act = n.p_act, P_x1 = n.p_x1, P_c1 = n.p_c1, P_c2 = n.p_c2, P_g = n.p_g,
                  P_hmid = n.p_hmid;
  const long long act_b = (long long)C * HW, cat_b = (long long)n.Ctot * HW, HWp = n.HWp;
  const long long HWq = n.HWq, P_hp = n.p_hp, hp_b = (long long)C * HWq;
  double* stats = n.stats.as<double>();
  const long long stats_per = (long long)B * C * 2;
  if (inorm) ACE_CHECK_CUDA(cudaMemsetAsync(stats, 0, (size_t)2 * NL * stats_per * sizeof(double), s));

  // network input -> split planes, stored in the tail channels of the concat buffer (zero-copy big skip)
  launch_norm_split(x, B, Cin, HW, nullptr, nullptr, nullptr, 0.f, hcat + (long long)C * HW, P_hcat, cat_b, HW, s);

  // encoder: Conv(Cin->C)+bias, GELU, Conv(C->C) ; + pos_embed          (sfnonet.py:566-577, :733)
  {
    GemmOp op = conv_op("encoder.0", hcat + (long long)C * HW, P_hcat, cat_b, HW, B, n.enc0, Cin);
    op.epi.flags |= EPI_GELU;
    out_planes(op, n.e1.as<bf16>(), P_act, act_b, HW);
    run_gemm(op, s);
  }
  bf16* hP = n.hP.as<bf16>();
  {
    GemmOp op = conv_op("encoder.2", n.e1.as<bf16>(), P_act, act_b, HW, B, n.enc1, C);
    if (c.pos_embed) add_f32(op, n.pos.as<float>(), 0, HW);
    out_planes(op, hP, P_hp, hp_b, HWq);
    if (inorm) row_stats(op, stats, C);
    run_gemm(op, s);
  }

  for (int i = 0; i < NL; ++i) {
    BlockW& w = n.blocks[i];
    const ace_sht_plan& pf = (i == 0) ? *n.outer : *n.inner;
    const ace_sht_plan& pi = (i == NL - 1) ? *n.outer : *n.inner;
    double* st0 = stats + (long long)(2 * i) * stats_per;
    double* st1 = stats + (long long)(2 * i + 1) * stats_per;
    // forward and inverse grids differ -> scale_residual (s2convolutions.py:81-85,170-173): the residual used by
    // inner_skip and the outer skip is inverse_transform(forward_transform(x_norm)) instead of x_norm itself
    const bool scale_res = pf.table_id != pi.table_id;
    // the residual operand of inner_skip / fc2: hP (un-normalised, affine deferred) or the round trip (already normalised)
    const bool res_deferred = inorm && !scale_res;
    bf16* xn = n.xn.as<bf16>();

    // norm0 (sfnonet.py:218-221), deferred: a0, s0 and the inner_skip weights with diag(a0) folded in
    if (inorm)
      launch_prep_norm_conv(st0, w.g0.as<float>(), w.b0.as<float>(), c.norm_eps, HW, B, C, res_deferred ? w.skip.wf.as<float>() : nullptr,
                            w.skip_total.as<float>(), C, w.skip.Ip, n.skipW.as<bf16>(), n.p_skipW, n.skipB.as<float>(),
                            n.na0.as<float>(), n.ns0.as<float>(), n.nsh0.as<float>(), s);
    // spectral convolution (s2convolutions.py:162-197)
    {
      GemmOp op = sht_op_dft_fwd(pf, hP, P_hp, hp_b, C, B, n.x1.as<bf16>(), P_x1, pf.Kp);
      if (inorm) {
        op.epi.flags |= EPI_ROW_AFFINE;  // DFT(a h + s) = a DFT(h) + 2 pi s [m = 0]
        op.epi.ra_scale = n.na0.as<float>();
        op.epi.ra_shift0 = n.nsh0.as<float>();
        op.epi.ra_z2 = C;
      }
      run_gemm(op, s);
    }
    run_gemm(sht_op_legendre_fwd(pf, n.x1.as<bf16>(), P_x1, C, B, n.c1.as<bf16>(), P_c1), s);
    if (scale_res) {
      run_gemm(sht_op_legendre_inv_from_c1(pi, n.c1.as<bf16>(), P_c1, C, B, n.g.as<bf16>(), P_g), s);
      run_gemm(sht_op_dft_inv_planes(pi, n.g.as<bf16>(), P_g, C, B, xn, P_act, act_b), s);
    }
    if (c.operator_type == 1) {
      // complex GEMM per degree l: D[(ro,o)][m] = sum_i W[l][o][i] (complex) * c1[l][m][i] (complex)  (contractions.py:184-195)
      run_gemm(dhconv_op(n.c1.as<bf16>(), P_c1, w.spec.as<bf16>(), w.spec_plane, pf, C, C, B, n.c2.as<bf16>(), P_c2), s);
    } else {
      launch_diagonal_contract(n.c1.as<bf16>(), P_c1, w.spec.as<float>(), B, C, pf.L, pf.M, pf.Lp, n.c2.as<bf16>(), P_c2, s);
    }
    if (options().inv2) {
      run_gemm(sht_op_legendre_inv2(pi, n.c2.as<bf16>(), P_c2, C, B, n.g2.as<bf16>(), n.p_g2), s);
      run_gemm(sht_op_dft_inv2(pi, n.g2.as<bf16>(), n.p_g2, C, B, n.T.as<float>(), act_b), s);
    } else {
      run_gemm(sht_op_legendre_inv(pi, n.c2.as<bf16>(), P_c2, C, B, n.g.as<bf16>(), P_g), s);
      run_gemm(sht_op_dft_inv(pi, n.g.as<bf16>(), P_g, C, B, n.T.as<float>(), act_b), s);
    }

    // residual operand of this block: x_norm = a0 hP + s0 (deferred), the round trip (scale_res), or hP itself (no norm)
    const bf16* resid = scale_res ? xn : hP;
    const long long r_plane = scale_res ? P_act : P_hp, r_b = scale_res ? act_b : hp_b, r_pitch = scale_res ? HW : HWq;
    // x = GELU(filter(x_norm) + bias_f + inner_skip(residual)) (sfnonet.py:223-232) -> tP (un-normalised) + stats1
    {
      GemmOp op = conv_op("inner_skip", resid, r_plane, r_b, HW, B, w.skip, C);
      op.B.s_k = r_pitch;
      op.epi.row_bias = w.skip_total.as<float>();
      if (res_deferred) use_folded(op, w.skip, n.skipW, n.p_skipW, n.skipB);
      op.epi.flags |= EPI_GELU;
      add_f32(op, n.T.as<float>(), act_b, HW);
      out_planes(op, n.tP.as<bf16>(), n.p_tp, (long long)C * HWp, HWp);
      if (inorm) row_stats(op, st1, C);
      run_gemm(op, s);
    }
    // norm1 (sfnonet.py:234-238), deferred into fc1
    if (inorm)
      launch_prep_norm_conv(st1, w.g1.as<float>(), w.b1.as<float>(), c.norm_eps, HW, B, C, w.fc1.wf.as<float>(), w.fc1.bias.as<float>(),
                            c.mlp_hidden, w.fc1.Ip, n.fc1W.as<bf16>(), n.p_fc1W, n.fc1B.as<float>(), n.na1.as<float>(),
                            n.ns1.as<float>(), n.nsh1.as<float>(), s);
    // MLP (layers.py:117-124) + outer skip (identity) with the block's residual (sfnonet.py:249-250)
    {
      GemmOp op = conv_op("mlp.fc1", n.tP.as<bf16>(), n.p_tp, (long long)C * HWp, HW, B, w.fc1, C);
      op.B.s_k = HWp;  // channel pitch of the padded buffer (N stays H*W)
      if (inorm) use_folded(op, w.fc1, n.fc1W, n.p_fc1W, n.fc1B);
      op.epi.flags |= EPI_GELU;
      out_planes(op, n.hmid.as<bf16>(), P_hmid, (long long)c.mlp_hidden * HWp, HWp);
      run_gemm(op, s);
    }
    {
      GemmOp op = conv_op("mlp.fc2", n.hmid.as<bf16>(), P_hmid, (long long)c.mlp_hidden * HWp, HW, B, w.fc2, c.mlp_hidden);
      op.B.s_k = HWp;
      op.epi.flags |= EPI_RES_PLANES;
      op.epi.res = resid;
      op.epi.res_plane = r_plane;
      op.epi.res_z2 = r_b;
      op.epi.res_m0 = r_pitch;
      op.epi.res_n = 1;
      if (res_deferred) {
        op.epi.flags |= EPI_RES_AFFINE;
        op.epi.res_a = n.na0.as<float>();
        op.epi.res_s = n.ns0.as<float>();
        op.epi.rsa_z2 = C;
      }
      if (i == NL - 1) {
        out_planes(op, hcat, P_hcat, cat_b, HW);  // head channels of the concat buffer
      } else {
        out_planes(op, hP, P_hp, hp_b, HWq);  // in place when resid == hP: every chunk is read before it is written
        if (inorm) row_stats(op, stats + (long long)(2 * i + 2) * stats_per, C);
      }
      run_gemm(op, s);
    }
  }

  // decoder on cat(x, input) (sfnonet.py:741-747)
  {
    GemmOp op = conv_op("decoder.0", hcat, P_hcat, cat_b, HW, B, n.dec0, c.big_skip ? n.Ctot : C);
    op.epi.flags |= EPI_GELU;
    out_planes(op, n.d1.as<bf16>(), P_act, act_b, HW);
    run_gemm(op, s);
  }
  {
    GemmOp op = conv_op("decoder.2", n.d1.as<bf16>(), P_act, act_b, HW, B, n.dec1, C);
    out_f32(op, y, (long long)c.out_chans * HW, HW);
    run_gemm(op, s);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------ C ABI

extern "C" int ace_sfno_create(const ace_sfno_config* cfg, ace_sht_plan* plan_outer, ace_sht_plan* plan_inner, ace_sfno** out) {
  ACE_API_BEGIN
  ACE_REQUIRE(cfg && plan_outer && plan_inner && out, "ace_sfno_create: null argument");
  const ace_sfno_config& c = *cfg;
  ACE_REQUIRE(c.img_h > 0 && c.img_w > 0 && c.in_chans > 0 && c.out_chans > 0 && c.embed_dim > 0 && c.num_layers > 0,
              "ace_sfno_create: non-positive size");
  ACE_REQUIRE(c.operator_type == 0 || c.operator_type == 1, "ace_sfno_create: operator_type must be 0 (diagonal) or 1 (dhconv)");
  ACE_REQUIRE(c.normalization == 0 || c.normalization == 1, "ace_sfno_create: normalization must be 0 (none) or 1 (instance_norm)");
  ACE_REQUIRE(c.mlp_hidden > 0, "ace_sfno_create: use_mlp=False is not supported");
  for (ace_sht_plan* p : {plan_outer, plan_inner})
    ACE_REQUIRE(p->K == c.img_h && p->W == c.img_w && p->L == c.lmax && p->M == c.mmax,
                "ace_sfno_create: plan (%d,%d,%d,%d) does not match config (%d,%d,%d,%d)", p->K, p->W, p->L, p->M, c.img_h,
                c.img_w, c.lmax, c.mmax);
  ace_sfno* n = new ace_sfno();
  try {
    n->cfg = c;
    n->outer = plan_outer;
    n->inner = plan_inner;
    n->HW = (long long)c.img_h * c.img_w;
    n->Ctot = c.embed_dim + c.in_chans;
    const int C = c.embed_dim;
    n->enc0.init(C, c.in_chans, true);
    n->enc1.init(C, C, false);
    n->dec0.init(C, c.big_skip ? n->Ctot : C, true);
    n->dec1.init(c.out_chans, C, false);
    n->blocks.resize(c.num_layers);
    for (BlockW& b : n->blocks) {
      b.skip.keep_f32 = b.fc1.keep_f32 = (c.normalization == 1);
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

extern "C" void ace_sfno_destroy(ace_sfno* net) { delete net; }

extern "C" int ace_sfno_query(ace_sfno* net, int* in_chans, int* out_chans, long long* hw) {
  ACE_API_BEGIN
  ACE_REQUIRE(net && in_chans && out_chans && hw, "ace_sfno_query: null argument");
  *in_chans = net->cfg.in_chans;
  *out_chans = net->cfg.out_chans;
  *hw = net->HW;
  ACE_API_END
}

extern "C" int ace_sfno_set_param(ace_sfno* net, const char* name, const float* data_dev, long long numel, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(net && name && data_dev, "ace_sfno_set_param: null argument");
  ace_sfno& n = *net;
  cudaStream_t s = (cudaStream_t)stream;
  const ace_sfno_config& c = n.cfg;
  const int C = c.embed_dim;
  std::string nm(name);
  auto it = n.params.find(nm);
  ACE_REQUIRE(it != n.params.end(), "ace_sfno_set_param: unexpected parameter '%s' for this configuration", name);
  if (nm == "pos_embed") {
    ACE_REQUIRE(numel == (long long)C * n.HW, "pos_embed: expected %lld elements, got %lld", (long long)C * n.HW, numel);
    copy_f32(n.pos, data_dev, numel, s);
  } else if (nm == "encoder.0.weight") set_conv_w(n.enc0, data_dev, numel, name, s);
  else if (nm == "encoder.0.bias") set_conv_b(n.enc0, data_dev, numel, name, s);
  else if (nm == "encoder.2.weight") set_conv_w(n.enc1, data_dev, numel, name, s);
  else if (nm == "decoder.0.weight") set_conv_w(n.dec0, data_dev, numel, name, s);
  else if (nm == "decoder.0.bias") set_conv_b(n.dec0, data_dev, numel, name, s);
  else if (nm == "decoder.2.weight") set_conv_w(n.dec1, data_dev, numel, name, s);
  else {
    int bi = -1, consumed = 0;
    ACE_REQUIRE(sscanf(name, "blocks.%d.%n", &bi, &consumed) == 1 && bi >= 0 && bi < c.num_layers, "bad block index in '%s'", name);
    BlockW& b = n.blocks[bi];
    std::string rest(name + consumed);
    auto vecC = [&](DevBuf& d) {
      ACE_REQUIRE(numel == C, "%s: expected %d elements, got %lld", name, C, numel);
      copy_f32(d, data_dev, numel, s);
    };
    if (rest == "norm0.weight") vecC(b.g0);
    else if (rest == "norm0.bias") vecC(b.b0);
    else if (rest == "norm1.weight") vecC(b.g1);
    else if (rest == "norm1.bias") vecC(b.b1);
    else if (rest == "filter.filter.bias") vecC(b.fbias);
    else if (rest == "filter.filter.weight") {
      if (c.operator_type == 1) {
        ACE_REQUIRE(numel == (long long)C * C * c.lmax * 2, "%s: expected %lld elements, got %lld", name, (long long)C * C * c.lmax * 2, numel);
        const int Cp = (int)round_up(C, 8);
        b.spec_plane = (long long)c.lmax * 2 * C * Cp;
        b.spec.ensure(2 * (size_t)b.spec_plane * sizeof(bf16));
        launch_prep_dhconv_cplx(data_dev, C, C, c.lmax, Cp, b.spec.as<bf16>(), b.spec_plane, s);
      } else {
        ACE_REQUIRE(numel == (long long)C * C * c.lmax * c.mmax * 2, "%s: expected %lld elements, got %lld", name, (long long)C * C * c.lmax * c.mmax * 2, numel);
        copy_f32(b.spec, data_dev, numel, s);
      }
    } else if (rest == "inner_skip.weight") set_conv_w(b.skip, data_dev, numel, name, s);
    else if (rest == "inner_skip.bias") set_conv_b(b.skip, data_dev, numel, name, s);
    else if (rest == "mlp.fwd.0.weight") set_conv_w(b.fc1, data_dev, numel, name, s);
    else if (rest == "mlp.fwd.0.bias") set_conv_b(b.fc1, data_dev, numel, name, s);
    else if (rest == "mlp.fwd.2.weight") set_conv_w(b.fc2, data_dev, numel, name, s);
    else if (rest == "mlp.fwd.2.bias") set_conv_b(b.fc2, data_dev, numel, name, s);
    else ACE_REQUIRE(false, "ace_sfno_set_param: unhandled parameter '%s'", name);
    if (rest == "inner_skip.bias" || rest == "filter.filter.bias") {
      std::string pre = "blocks." + std::to_string(bi) + ".";
      bool other = n.params[pre + (rest == "inner_skip.bias" ? "filter.filter.bias" : "inner_skip.bias")];
      if (other) launch_vec_add(b.skip.bias.as<float>(), b.fbias.as<float>(), b.skip_total.as<float>(), C, s);
    }
  }
  it->second = true;
  n.finalized = false;
  ACE_API_END
}

extern "C" int ace_sfno_finalize(ace_sfno* net) {
  ACE_API_BEGIN
  ACE_REQUIRE(net, "ace_sfno_finalize: null argument");
  for (auto& kv : net->params)
    if (!kv.second) throw Error(ACE_ERR_STATE, "ace_sfno_finalize: parameter '" + kv.first + "' has not been set");
  net->finalized = true;
  ACE_API_END
}

extern "C" int ace_sfno_forward(ace_sfno* net, const float* x_dev, float* y_dev, int batch, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(net && x_dev && y_dev, "ace_sfno_forward: null argument");
  ACE_REQUIRE(batch > 0, "ace_sfno_forward: batch must be positive");
  if (!net->finalized) throw Error(ACE_ERR_STATE, "ace_sfno_forward: call ace_sfno_finalize first");
  forward(*net, x_dev, y_dev, batch, (cudaStream_t)stream);
  ACE_API_END
}
