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

}  // namespace ace
