// Building blocks shared by the two network forward passes (sfno.cu, csfno.cu): 1x1-convolution weights as split planes and
// the GemmOp builders / epilogue setters for activations laid out [batch][channel][space].
#pragma once
#include "kernels.cuh"
#include "sht.cuh"

namespace ace {

struct ConvW {
  DevBuf w;  // planes [O][Ip]
  long long plane = 0;
  int O = 0, I = 0, Ip = 0;
  DevBuf bias;  // fp32 [O]
  DevBuf wf;    // fp32 [O][I] copy (only for convolutions that a deferred InstanceNorm is folded into)
  bool keep_f32 = false;
  bool has_bias = false;
  void init(int o, int i, bool b) {
    O = o;
    I = i;
    Ip = (int)round_up(i, 8);
    plane = (long long)O * Ip;
    w.ensure(2 * (size_t)plane * sizeof(bf16));
    has_bias = b;
    if (b) bias.ensure((size_t)O * sizeof(float));
  }
};

inline void copy_f32(DevBuf& dst, const float* src, long long n, cudaStream_t s) {
  dst.ensure((size_t)n * sizeof(float));
  ACE_CHECK_CUDA(cudaMemcpyAsync(dst.p, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s));
}

inline void set_conv_w(ConvW& w, const float* src, long long numel, const char* name, cudaStream_t s) {
  ACE_REQUIRE(numel == (long long)w.O * w.I, "%s: expected %lld elements, got %lld", name, (long long)w.O * w.I, numel);
  launch_split_pad(src, w.O, w.I, w.Ip, w.w.as<bf16>(), w.plane, s);
  if (w.keep_f32) copy_f32(w.wf, src, numel, s);
}
inline void set_conv_b(ConvW& w, const float* src, long long numel, const char* name, cudaStream_t s) {
  ACE_REQUIRE(w.has_bias && numel == w.O, "%s: expected %d elements, got %lld", name, w.O, numel);
  copy_f32(w.bias, src, numel, s);
}

// 1x1 convolution as a GEMM with channels on the accumulator rows: D[o][hw] = sum_i W[o][i] * x[i][hw].
// A = weights (K-major), B = activations [channel][space] (MN-major: space contiguous), so the output
// row of a thread is a channel and its columns are contiguous in memory (NC epilogue, gemm_umma.cu).
inline GemmOp conv_op(const char* name, const bf16* x, long long x_plane, long long x_batch_stride, long long HW, int B,
               const ConvW& w, int k_channels) {
  GemmOp op = make_gemm_op(name);
  op.M = w.O;
  op.N = (int)HW;
  op.K = k_channels;
  op.Z2 = B;
  op.A = {w.w.as<bf16>(), w.plane, (long long)w.Ip, 1, 0, 0};
  op.B = {x, x_plane, 1, HW, 0, x_batch_stride};
  if (w.has_bias) {
    op.epi.flags |= EPI_ROW_BIAS;
    op.epi.row_bias = w.bias.as<float>();
  }
  return op;
}
// same convolution with per-sample weights / bias (InstanceNorm folded in by prep_norm_conv)
inline void use_folded(GemmOp& op, const ConvW& w, const DevBuf& wfold, long long wfold_plane, const DevBuf& bfold) {
  op.A = {wfold.as<bf16>(), wfold_plane, (long long)w.Ip, 1, 0, (long long)w.O * w.Ip};
  op.epi.flags |= EPI_ROW_BIAS;
  op.epi.row_bias = bfold.as<float>();
  op.epi.rb_z2 = w.O;
}

inline void out_planes(GemmOp& op, bf16* out, long long plane, long long batch_stride, long long HW) {
  op.epi.flags |= EPI_OUT_PLANES;
  op.epi.out = out;
  op.epi.out_plane = plane;
  op.epi.o_z2 = batch_stride;
  op.epi.o_m0 = HW;
  op.epi.o_n = 1;
}
inline void out_f32(GemmOp& op, float* out, long long batch_stride, long long HW) {
  op.epi.flags |= EPI_OUT_F32;
  op.epi.outf = out;
  op.epi.f_z2 = batch_stride;
  op.epi.f_m0 = HW;
  op.epi.f_n = 1;
}
inline void add_f32(GemmOp& op, const float* add, long long batch_stride, long long HW) {
  op.epi.flags |= EPI_ADD_F32;
  op.epi.add = add;
  op.epi.add_z2 = batch_stride;
  op.epi.add_m0 = HW;
  op.epi.add_n = 1;
}
// dhconv (contractions.py:184-195): per degree l a complex GEMM  c2[m][l][o] = sum_i c1[l][m][i] * W[l][o][i]  over the orders
// m <= l.  Weights: planes [L][2 (re, im)][Cout][Cinp]; c1: [B][L][M][2 Cin]; c2: [B][M][Lp][2 Cout].
//   option dhconv_t = 0 (default): weights on the accumulator rows, the l + 1 orders on the columns (cplx == 1, ROWC epilogue):
//     the fewest multiplications -- under the power cap that wins (83 vs 89 us at ACE2 size, profiles/r02_probe2.jsonl);
//   option dhconv_t = 1: orders on the accumulator rows, output channels on the columns (GemmOp::cplx == 2) -- every MMA is
//     128 x 128, the output is stored along its contiguous (re/im, channel) axis (NC epilogue).
inline GemmOp dhconv_op(const bf16* c1, long long c1_plane, const bf16* w, long long w_plane, const ace_sht_plan& p, int Cin, int Cout,
                        int B, bf16* c2, long long c2_plane) {
  GemmOp op = make_gemm_op("dhconv");
  const int Cp = (int)round_up(Cin, 8);
  op.K = Cin;
  op.Z1 = p.L;
  op.Z2 = B;
  op.epi.flags = EPI_OUT_PLANES;
  op.epi.out = c2;
  op.epi.out_plane = c2_plane;
  op.epi.o_z2 = (long long)p.M * p.Lp * 2 * Cout;
  op.epi.o_z1 = 2LL * Cout;
  if (options().dhconv_t && Cin % 8 == 0) {
    op.cplx = 2;
    op.M = p.M;
    op.N = Cout;
    op.m_hi_z1 = 1;  // order m <= degree l
    op.A = {c1, c1_plane, 2LL * Cin, 1, (long long)p.M * 2 * Cin, (long long)p.L * p.M * 2 * Cin};
    op.B = {w, w_plane, (long long)Cp, 1, 2LL * Cout * Cp, 0};
    op.b_part = (long long)Cout * Cp;
    op.epi.o_m0 = (long long)p.Lp * 2 * Cout;
    op.epi.o_n = 1;
  } else {
    op.cplx = 1;
    op.M = Cout;
    op.N = p.M;
    op.a_part = (long long)Cout * Cp;
    op.A = {w, w_plane, (long long)Cp, 1, 2LL * Cout * Cp, 0};
    op.B = {c1, c1_plane, 2LL * Cin, 1, (long long)p.M * 2 * Cin, (long long)p.L * p.M * 2 * Cin};
    op.n_hi_z1 = 1;  // order m <= degree l
    op.epi.o_n = (long long)p.Lp * 2 * Cout;
    op.epi.o_m0 = 1;
  }
  return op;
}

// dhconv with filter_num_groups = G > 1 (fme/core/models/conditional_sfno/s2convolutions.py:119-135): the per-degree operator is
// block diagonal.  Weights: planes [L][2][C][cgp] -- row o = g * cg + o' holds the cg inputs of ITS group only (1/G of the
// dense operator's bytes and multiplications); the GEMM is the grouped complex mode of gemm.cuh (orders on the rows).
inline GemmOp dhconv_grouped_op(const bf16* c1, long long c1_plane, const bf16* w, long long w_plane, const ace_sht_plan& p, int C, int G,
                                int B, bf16* c2, long long c2_plane) {
  GemmOp op = make_gemm_op("dhconv");
  const int cg = C / G, cgp = (int)round_up(cg, 8);
  op.cplx = 2;
  op.group_n = cg;
  op.a_part_k = C;
  op.M = p.M;
  op.N = C;
  op.K = cg;
  op.Z1 = p.L;
  op.Z2 = B;
  op.m_hi_z1 = 1;
  op.A = {c1, c1_plane, 2LL * C, 1, (long long)p.M * 2 * C, (long long)p.L * p.M * 2 * C};
  op.B = {w, w_plane, (long long)cgp, 1, 2LL * C * cgp, 0};
  op.b_part = (long long)C * cgp;
  op.epi.flags = EPI_OUT_PLANES;
  op.epi.out = c2;
  op.epi.out_plane = c2_plane;
  op.epi.o_z2 = (long long)p.M * p.Lp * 2 * C;
  op.epi.o_z1 = 2LL * C;
  op.epi.o_m0 = (long long)p.Lp * 2 * C;
  op.epi.o_n = 1;
  return op;
}

inline void row_stats(GemmOp& op, double* stats, int C) {
  op.epi.flags |= EPI_ROW_STATS;
  op.epi.stats = stats;
  op.epi.stats_z2 = C;
}


}  // namespace ace
