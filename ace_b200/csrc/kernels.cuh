// Launchers of the non-GEMM kernels (elementwise.cu).  All are HBM-bound streaming kernels:
// 16-byte vector accesses where alignment allows, grid sized in multiples of the SM count.
#pragma once
#include "common.cuh"

namespace ace {

// fp32 [B][C][HW] (contiguous) -> split planes at dst + b*dst_b + c*dst_c + hw  (lo plane at +plane).
// With stats != nullptr applies InstanceNorm first:
//   mean = S/HW, var = Q/HW - mean^2 (biased), y = (x-mean) * rsqrt(var+eps) * gamma[c] + beta[c]
// (nn.InstanceNorm2d(eps, affine=True, track_running_stats=False), sfnonet.py:593-601),
// stats = double [B][C][2] = {sum, sum of squares} accumulated by a GEMM epilogue.
void launch_norm_split(const float* src, int B, int C, long long HW, const double* stats, const float* gamma,
                       const float* beta, float eps, bf16* dst, long long plane, long long dst_b, long long dst_c,
                       cudaStream_t stream);

// fp32 [rows][cols] -> zero-padded split planes [rows][cols_pad]
void launch_split_pad(const float* src, long long rows, int cols, int cols_pad, bf16* dst, long long plane,
                      cudaStream_t stream);

// dhconv weight [Cin][Cout][L][2] -> planes [L][2 (re, im)][Cout][Cinp] for the complex GEMM mode (gemm.cuh)
void launch_prep_dhconv_cplx(const float* w, int Cin, int Cout, int L, int Cinp, bf16* dst, long long plane, cudaStream_t stream);

// diagonal operator (contractions.py:170-180) on the spectral layouts:
//   c1 [B][L][M][2C] planes -> c2 [B][M][Lp][2C] planes, w fp32 [C][C][L][M][2]
void launch_diagonal_contract(const bf16* c1, long long c1_plane, const float* w, int B, int C, int L, int M, int Lp,
                              bf16* c2, long long c2_plane, cudaStream_t stream);

// standalone SHT API layout conversions
// c1 planes [L][M][2C] -> complex64 [C][L][M]
void launch_spec_planes_to_complex(const bf16* c1, long long plane, int C, int L, int M, float* out, cudaStream_t stream);
// complex64 [C][L][M] -> c2 planes [M][Lp][2C]
void launch_spec_complex_to_planes(const float* in, int C, int L, int M, int Lp, bf16* c2, long long plane, cudaStream_t stream);

// Deferred InstanceNorm: statistics -> (a, s, 2*pi*s) per (sample, channel) and, when w != nullptr, the conv
// weights / bias with the normalisation folded in (per sample): wout planes [B][O][Ip], bout [B][O].
void launch_prep_norm_conv(const double* stats, const float* gamma, const float* beta, float eps, long long HW, int B, int C,
                           const float* w, const float* bias, int O, int Ip, bf16* wout, long long wplane, float* bout,
                           float* a_out, float* s_out, float* shift0_out, cudaStream_t stream);

// out[i] = a[i] + b[i]
void launch_vec_add(const float* a, const float* b, float* out, long long n, cudaStream_t stream);

// stepper: gather + normalise network input  (fme/core/normalizer.py:213-243, packer.py:45-52)
//   x[b][c] = (src_c[b][idx_c] - mean[c]) / std[c],  src = prog or forcing according to kind[c]
void launch_pack_normalize(const float* prog, const float* forcing, int n_prog, int n_forcing, const int* kind,
                           const int* index, const float* mean, const float* std, int B, int n_in, long long HW,
                           float* x, cudaStream_t stream);
// stepper: (residual add) + denormalise + ForcePositive clamp + ocean prescriber + scatter to next state
//   clamp[c] != 0: v = max(v, 0) (corrector/utils.py:26-43); channel ocean_out: v = target where round(mask) == 1, or the
//   linear blend when ocean_interp (prescriber.py:94-108); ocean = [B][2][HW] {mask, target}
void launch_unpack_denormalize(const float* y, const float* x_norm, const int* out_prog_index, const int* prog_in_chan,
                               const float* mean, const float* std, int residual, int B, int n_out, int n_in,
                               int n_prog, long long HW, const int* clamp, int ocean_out, int ocean_interp, const float* ocean,
                               float* out, float* next_prog, cudaStream_t stream);

// ocean prescriber as a separate pass over the surface-temperature channel (used when conservation correctors run between
// ForcePositive and the ocean); ocean = [B][2][HW] {mask, target}
void launch_ocean_prescribe(float* out, float* next_prog, const int* out_prog_index, int B, int n_out, int n_prog, long long HW,
                            int ocean_out, int ocean_interp, const float* ocean, cudaStream_t stream);

// slab ocean instead of the prescribed target: ocean = [B][3][HW] {mask, q_flux, mixed layer depth}; target = T_in + (F_net + Q) /
// (rho depth c_p) dt (fme/core/ocean.py:64-88,223-243); channel indices into out (fluxes) / the prognostic input state (T_in)
struct SlabOceanIdx {
  int prog_sst, dlw, ulw, dsw, usw, lhf, shf;
  float dt;
};
void launch_ocean_slab(float* out, float* next_prog, const float* prev_prog, const int* out_prog_index, int B, int n_out, int n_prog, long long HW,
                       int ocean_out, int ocean_interp, const float* ocean, const SlabOceanIdx& ix, cudaStream_t stream);

}  // namespace ace

namespace ace {

// ---- noise-conditioned SFNO (cln.cu) ----
// ConditionalLayerNorm (conditional_sfno/layers.py:95-141,285-320) on split planes [B][C][HW]:
//   y = ((x - mean) rsqrt(var + eps) lnw[c] + lnb[c]) * (sb0[b][c].s + sum_e w2[c][e].s ctx[b][e][hw]) + (sb0[b][c].b + sum_e w2[c][e].b ctx[b][e][hw])
// mean / var over the C channels of the pixel (biased); lnw/lnb, sb0 may be null (1, 0); Ep in {0, 8, 16, 32, 64}.
void launch_cond_layer_norm(const bf16* x, long long x_plane, long long x_b, int B, int C, long long HW, const float* lnw, const float* lnb,
                            const float* sb0, const float* w2, const float* ctx, int Ep, float eps, bf16* out, long long o_plane,
                            long long o_b, cudaStream_t stream);
// helpers of the tensor-core ConditionalLayerNorm (GemmOp::cln): per-pixel {mean, rstd} [B][HW][2]; the context as K-major
// planes [B][HW][2 Ep] (channels twice along k); the weights as K-major planes [C][2 Ep] = [scale | bias]
void launch_cln_stats(const bf16* x, long long x_plane, long long x_b, int B, int C, long long HW, float eps, float* musr, cudaStream_t stream);
void launch_cln_ctx_planes(const float* ctx, int B, int Ep, long long HW, bf16* out, long long plane, cudaStream_t stream);
void launch_cln_w_planes(const float* w2, int C, int Ep, bf16* out, long long plane, cudaStream_t stream);
// smallest supported padded context width >= E, or -1
int cln_padded_context(int E);
// sb0 [B][C][2]: the Linear maps of the per-sample context vectors (scalar embedding, labels) incl. their biases
void launch_cln_vector_terms(const float* scalar, int Es, const float* labels, int El, const float* Ws, const float* bs, const float* Wb,
                             const float* bb, const float* Wsl, const float* bsl, const float* Wbl, const float* bbl, int B, int C, float* sb0,
                             cudaStream_t stream);
// ctx [B][Ep][HW] = concat(noise [B][En][HW], pos [B][Epos][HW], zero padding)
void launch_concat_ctx(const float* noise, int En, const float* pos, int Epos, int B, long long HW, int Ep, float* ctx, cudaStream_t stream);
// w2 [C][Ep][2] from the 1x1-convolution weights W_scale_2d / W_bias_2d [C][En] and W_scale_pos / W_bias_pos [C][Epos]
void launch_build_cln_w2(const float* ws_n, const float* wb_n, int En, const float* ws_p, const float* wb_p, int Epos, int C, int Ep, float* w2,
                         cudaStream_t stream);
// dhconv weight with element (l, o, i, part) at w[(l*s_l + o*s_o + i*s_i)*2 + part] -> planes [L][2 (re, im)][Cout][Cinp]
void launch_prep_dhconv_grouped(const float* w, int G, int cg, int L, int cgp, bf16* dst, long long plane, cudaStream_t stream);
void launch_prep_dhconv_cplx_strided(const float* w, int Cin, int Cout, int L, long long s_l, long long s_o, long long s_i, int Cinp, bf16* dst,
                                     long long plane, cudaStream_t stream);

}  // namespace ace
