// Spherical harmonic transform plan: device tables + the four GemmOps a transform pair is made of.
//
//   forward  (fme/sht_fix.py:125-151):   x --G1: longitude DFT--> X1 --G2: Legendre/quadrature--> c1
//   inverse  (fme/sht_fix.py:206-226):   c2 --G4: Legendre--> g --G5: inverse longitude DFT--> y
//
// Device layouts (split-bf16 planes unless noted; element counts per plane):
//   x   [B][C][K][W]            grid-point fields (K = nlat, W = nlon)
//   X1  [B][2M][C][Kp]          row n = 2m+reim of the DFT output, latitude contiguous (Kp = K padded to 8)
//   c1  [B][L][M][2C]           forward coefficients, (reim, channel) contiguous
//   c2  [B][M][Lp][2C]          inverse input (after the spectral operator), (reim, channel) contiguous
//   g   [B][2M][C][K]           Legendre-synthesised Fourier coefficients, latitude contiguous
//   y   fp32 [B][C][K][W]
#pragma once
#include "gemm.cuh"

struct ace_sht_plan {
  int K, W, L, M;      // nlat, nlon, lmax, mmax
  int Kp, Lp, Wp, K2p; // padded strides (Lp: multiple of 8 elements = 16 bytes; Kp, Wp, K2p: of 64 elements = one 128-byte line)
  int Kt, Lt;          // row pitches of the forward / inverse Legendre tables (multiples of 64 elements)
  // inverse pair, second generation (sht_op_legendre_inv2 / sht_op_dft_inv2):
  //   g2  [B][K2a][C][Kg]   Kg = nlat padded to 64 rows per channel (128-byte rows, no ragged chunk in the Legendre epilogue);
  //                         k index of the inverse DFT = even orders first (2 (m / 2) + reim), odd orders from Ke on
  //   idft2 planes [W][K2q] inverse DFT rows in that k order (rows j and j + W/2 differ only in the sign of the odd part)
  int Kg, Ke, K2, K2a, K2q;
  unsigned long long table_id = 0;  // FNV-1a of the two host tables: equal ids <=> same grid/normalisation
  ace::DevBuf wt;       // planes [M][L][Kt]    P_l^m(cos th_k) w_k
  ace::DevBuf pinv;     // planes [M][Kg][Lt]   P_l^m(cos th_k), l contiguous (rows k >= K are zero)
  ace::DevBuf idft2;
  long long idft2_plane;
  ace::DevBuf fdft;     // planes [2M][Wp]      forward DFT rows (2pi/W)(cos, -sin)
  ace::DevBuf idft;     // planes [W][K2p]      inverse DFT rows, Hermitian weights folded in
  long long wt_plane, pinv_plane, fdft_plane, idft_plane;
  // lazily grown workspace of the standalone transform API
  ace::DevBuf ws_x, ws_x1, ws_c1, ws_c2, ws_g;
  ace::DevBuf ws_g2;    // HEALPix inverse: the padded, parity-split g2 layout (the lat-lon standalone API uses ws_g)
  ace::DevBuf ws_hpx;   // HEALPix ring stage (healpix.cu): [fields][rings][M] complex fp32 between the ring DFT and the plane layouts
  long long ws_fields = 0;         // capacity
  long long ws_fields_layout = 0;  // field count the c1/c2 zero regions are currently valid for

  long long x1_elems(int C) const { return 2LL * M * C * Kp; }
  long long c1_elems(int C) const { return (long long)L * M * 2 * C; }
  long long c2_elems(int C) const { return (long long)M * Lp * 2 * C; }
  long long g_elems(int C) const { return 2LL * M * C * K; }
  long long g2_elems(int C) const { return (long long)K2a * C * Kg; }
};

namespace ace {

// All builders take per-plane element offsets (`*_plane`) and per-sample strides implied by the
// layouts above with batch B folded as the outermost dimension.
// rows_per_channel (0 = nlat): the grid-point buffer may carry pad rows after each channel's nlat latitude rows (channel pitch =
// rows_per_channel * nlon, rows_per_channel <= Kp); pad rows must hold finite values, their transforms land in the pad slots of
// X1 that the Legendre stage never reads.  With rows_per_channel == Kp the X1 stores of a tile are one contiguous run per column.
GemmOp sht_op_dft_fwd(const ace_sht_plan& p, const bf16* x, long long x_plane, long long x_batch_stride, int C, int B,
                      bf16* x1, long long x1_plane, int rows_per_channel = 0);
GemmOp sht_op_legendre_fwd(const ace_sht_plan& p, const bf16* x1, long long x1_plane, int C, int B, bf16* c1,
                           long long c1_plane);
GemmOp sht_op_legendre_inv(const ace_sht_plan& p, const bf16* c2, long long c2_plane, int C, int B, bf16* g,
                           long long g_plane);
GemmOp sht_op_dft_inv(const ace_sht_plan& p, const bf16* g, long long g_plane, int C, int B, float* y,
                      long long y_batch_stride);
// Second-generation inverse pair (the networks' main path): padded, parity-split g2 and the butterfly inverse DFT
//   y[j] = E[j] + O[j],  y[j + W/2] = E[j] - O[j]   (E / O = contributions of the even / odd orders m: half the multiplications)
GemmOp sht_op_legendre_inv2(const ace_sht_plan& p, const bf16* c2, long long c2_plane, int C, int B, bf16* g2, long long g2_plane);
GemmOp sht_op_dft_inv2(const ace_sht_plan& p, const bf16* g2, long long g2_plane, int C, int B, float* y, long long y_batch_stride);
// Round-trip residual of SpectralConvS2 when the forward and inverse grids differ
// (s2convolutions.py:81-85,170-173): inverse Legendre stage reading the c1 layout in place ...
GemmOp sht_op_legendre_inv_from_c1(const ace_sht_plan& p, const bf16* c1, long long c1_plane, int C, int B, bf16* g,
                                   long long g_plane);
// ... and the inverse DFT storing split planes [B][C][K][W] instead of fp32
GemmOp sht_op_dft_inv_planes(const ace_sht_plan& p, const bf16* g, long long g_plane, int C, int B, bf16* y,
                             long long y_plane, long long y_batch_stride);

}  // namespace ace
