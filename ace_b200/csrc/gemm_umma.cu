// tcgen05 / TMA implementation of the fast subsets of the GemmOp contract (gemm.cuh) for sm_100a.
//
// One persistent CTA per SM, 512 threads, warp-specialised:
//   warp 0    TMA producer   -- cp.async.bulk.tensor loads of the hi and lo bf16 planes of the A and B
//                               tiles into a multi-stage shared-memory ring (hardware swizzle)
//   warp 1    MMA issuer     -- one thread issues tcgen05.mma (M=128, N<=256, K=16, bf16 x bf16 -> fp32
//                               in TMEM); per K-step THREE products: Ahi*Bhi + Ahi*Blo + Alo*Bhi, which
//                               reproduces an fp32 product to ~2^-17 relative (error budget)
//   warp 2    TMEM allocator -- 2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 4-15 epilogue      -- three warps per TMEM lane quarter, interleaved 32-column chunks:
//                               tcgen05.ld -> registers -> fused epilogue -> global memory
//
// Operand layouts: A and B may each be K-major (rows of K contiguous; 64-byte swizzle, BK = 32) or
// MN-major (the M / N index contiguous: activations seen as [channel][space]; 128-byte swizzle atoms of
// 64 elements).  Out-of-range parts of any tile are zero-filled by TMA, so M, N, K need no padding.
//
// Epilogue shapes (compile-time; everything else goes to the SIMT kernel):
//   ROWC  output rows contiguous in memory (o_m0 == 1): thread = accumulator row, each column is a
//         lane-coalesced 2-byte store per plane.  Used by the transposing stages of the SHT
//         (DFT -> m-major, Legendre -> l-major, dhconv -> m-major).  Split-plane output only.
//   NC    output columns contiguous (o_n == 1 / f_n == 1, rows affine): the 32x32 chunk a warp holds
//         (thread = row) is transposed through a 4 KB per-warp shared-memory tile so that global loads of
//         the fp32 addend / split-plane residual and the stores are 64-128 B coalesced per row.
//         Compile-time flags: ADD_F32, RES_PLANES, GELU, OUT_PLANES | OUT_F32; run-time: ROW_BIAS,
//         ROW_STATS (per-row = per-channel InstanceNorm sums, thread-local then two fp64 atomics per tile).
//
// Triangular ops (gemm.cuh): the N extent of the MMA shrinks to the non-zero column range of the tile
// (UMMA N is a runtime field of the instruction descriptor) and K-chunks below k_lo are skipped.
#include <cuda.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "gemm.cuh"
#include "ptx.cuh"

namespace ace {

namespace {

struct UmmaParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  GemmOp op;
  int tiles_m, tiles_n;
  int nterms;
  int a_z1_on, a_z2_on, b_z1_on, b_z2_on;  // 0 when the operand does not vary along that batch axis
  int dbg;                                 // development switches (option "dbg"): 1 skip epilogue, 2 skip TMA, 4 skip MMA
  int scalar_store;                        // epilogue stores element-wise (row groups / row starts not vector-aligned, e.g. odd nlat)
  int m_fastest;                           // tile order: consecutive tiles share the B (1) or the A (0) tile
  // Triangular ops: explicit list of the NON-EMPTY tiles {tm, tn, z1, z2}, heaviest first (host-built, cached per shape).
  // Walking the full tm x tn x z1 x z2 box round-robin leaves CTAs idle whenever the empty slots correlate with the CTA
  // index (dhconv: for l < 128 every odd slot is empty and the grid size is even -> half the SMs idled for 70 % of the
  // kernel) and gives every CTA a different amount of work when the tile cost depends on z1 (Legendre stages).
  const int4* tiles;
  long long n_listed;
  // development trace (option "trace"): [CTA][tile iteration < kTraceIters][kTraceSlots] clock64 samples of the three roles
  long long* trace;
  // SP variants with split-plane output: 4-D map (position, channel, sample, plane) of the output for the epilogue's TMA stores
  // (sp_tma = 0: strides not 16-byte aligned -> element-wise stores from registers)
  CUtensorMap tmO;
  int sp_tma;
  // ... and of the epilogue operand (fp32 addend, 3-D + a unit axis, or split-plane residual) for its TMA loads
  CUtensorMap tmX;
  int sp_tmx;
  int mma_batch;  // option "mma_batch": the MMA warp issues two ring slots per barrier round when both have landed
};

constexpr int kTraceIters = 41, kTraceSlots = 16;
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_put(const UmmaParams& p, uint32_t it, int slot, long long v) {
  if (p.trace && it < (uint32_t)kTraceIters) p.trace[((long long)blockIdx.x * kTraceIters + it) * kTraceSlots + slot] = v;
}

constexpr uint32_t EF_MASK = EPI_ADD_F32 | EPI_RES_PLANES | EPI_GELU | EPI_OUT_PLANES | EPI_OUT_F32;
constexpr int kEpiWarps = 12;
constexpr int kThreadsUmma = 32 * (4 + kEpiWarps);
constexpr int kStgBytesPerWarp = 2048;
constexpr int kMaxSmem = 232448;  // 227 KB opt-in limit per CTA

template <int BN_, bool A_MN_, bool B_MN_, uint32_t EF_, bool NC_, int BK_ = 32, bool CPLX_ = false, bool PAIR_ = false, bool BFLY_ = false,
          bool SP_ = false, int SP_STAGES_ = 8, bool CLN_ = false>
struct Cfg {
  static constexpr bool CLN = CLN_;  // ConditionalLayerNorm mode (GemmOp::cln): two accumulators split along k like the butterfly mode
  // SP ("space on the rows"): a 1x1 convolution executed with the operand roles exchanged -- A = activations (MN-major, 256
  // spatial positions per CTA pair), B = weights (K-major, the output channels on the accumulator columns).  A thread of the
  // epilogue then owns ONE spatial position, and for every channel the 32 lanes of a warp touch 32 consecutive positions of
  // that channel's row: the fp32 addend, the split-plane residual and the output are read / written straight from registers
  // with coalesced accesses, no shared-memory transposition (the convolutions are shared-memory-bandwidth bound otherwise:
  // MMA operand reads + TMA fills + epilogue staging ~ 120 B/clk of the SM's 128 B/clk).
  static constexpr bool SP = SP_;
  static constexpr int BM = 128, BN = BN_, BK = BK_;
  static constexpr bool A_MN = A_MN_, B_MN = B_MN_, NC = NC_, CPLX = CPLX_;
  // complex mode with the NC epilogue = GemmOp::cplx == 2: [Ar | Ai] along the k axis of A, {Br, Bi} as two matrices, the
  // imaginary accumulator lands N columns further (dhconv with the orders m on the rows and the output channels on the columns)
  static constexpr bool CT = CPLX_ && NC_;
  // PAIR: two CTAs of a cluster compute one 256 x BN tile with tcgen05.mma.cta_group::2 -- each CTA stages its own 128
  // rows of A and HALF of the B tile (the tensor cores read the other half from the peer), which cuts the shared-memory
  // and L2->SM operand traffic per SM by a third for a 256-wide tile
  static constexpr bool PAIR = PAIR_;
  static constexpr int BNL = PAIR ? BN / 2 : BN;  // B columns staged by this CTA
  static constexpr int TILE_M = PAIR ? 256 : 128;
  // butterfly mode (GemmOp::bfly): K-chunks below / from k_split accumulate into two accumulators E, O; the epilogue stores
  // E + O at column n and E - O at column n + N
  static constexpr bool BFLY = BFLY_;
  static constexpr int NACC = (CPLX || BFLY || CLN_) ? 2 : 1;  // accumulators per tile (complex mode: real and imaginary part)
  static constexpr uint32_t EF = EF_;
  static constexpr int UMMA_K = 16;
  static constexpr int A_PLANE = BM * BK * 2;  // bytes
  static constexpr int B_PLANE = BNL * BK * 2;
  static constexpr int MN_ATOM = 2 * BK * 128;  // MN-major tiles: one 64-wide atom = [plane][BK k-rows][128 B]; the lo plane sits BK * 128 B after the hi plane
  static constexpr int A_LO = A_MN ? BK * 128 : A_PLANE, B_LO = B_MN ? BK * 128 : B_PLANE;  // byte offset of the lo plane inside the tile
  static constexpr int STAGE = 2 * (CPLX ? 2 : 1) * (A_PLANE + B_PLANE);  // complex mode: {Ar, Ai} x {hi, lo}, then {Br, Bi} x {hi, lo}
  // SP: per warp group (the four lane-quarter warps that share a 32-channel chunk) a [plane][32 channels][128 positions] bf16
  // tile the output is stored from by TMA, then the per-channel statistics of the tile
  // (SP_X: a second tile per group that TMA loads the addend [32][128] fp32 / the residual [plane][32][128] bf16 into;
  //  SP_VEC: per warp the bias / residual scale / residual shift of the chunk's 32 channels)
  static constexpr int SP_TILE = 2 * 32 * 128 * 2, SP_NG = kEpiWarps / 4;
  //  a split-plane residual has the layout of the output tile and is loaded into the Y tile itself (SP_SHARED): a thread reads and
  //  later overwrites exactly its own elements, and the two extra ring slots are worth more than the chunk of load latency)
  static constexpr bool SP_OPND = (EF_ & (EPI_ADD_F32 | EPI_RES_PLANES)) != 0;
  static constexpr bool SP_SHARED = SP_OPND;  // (the fp32 addend tile has the same size; barrier A separates its reads from the output writes)
  static constexpr int SP_X = SP_SHARED ? 0 : SP_NG * SP_TILE;
  static constexpr int SP_STATS = SP_NG * SP_TILE + ((SP_OPND && !SP_SHARED) ? SP_NG * SP_TILE : 0), SP_VEC = SP_STATS + 2048;
  static constexpr int CLN_COLS = kEpiWarps * kStgBytesPerWarp;  // CLN: per warp the {mean, rstd} of the chunk's 32 columns (256 B)
  static constexpr int STG_BYTES = SP ? SP_VEC + kEpiWarps * 3 * 32 * 4 : kEpiWarps * kStgBytesPerWarp + (CLN_ ? kEpiWarps * 256 : 0);
  // SP: the epilogue reads its addend / residual with ordinary global loads, which need L1 lines to land in; with the whole
  // 227 KB configured as shared memory hardly any are left (measured: those loads cost 24 - 37 us per launch).  Four ring
  // slots keep the kernel inside the 164 KB carve-out, i.e. 64 KB of L1.
  static constexpr int MAX_STAGES = SP_ ? SP_STAGES_ : 8;
  static constexpr int BAR_BYTES = (2 * MAX_STAGES + 4) * 8 + 16 + 32;  // ring, accumulator, TMEM slot, SP operand-tile barriers
  static constexpr int STAGES_RAW = (kMaxSmem - 1024 - STG_BYTES - BAR_BYTES) / STAGE;
  static constexpr int STAGES = STAGES_RAW > MAX_STAGES ? MAX_STAGES : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * NACC * BN <= 256) ? 256 : 512;
  static_assert(2 * NACC * BN <= 512, "TMEM budget");
  static_assert(!CPLX || (!A_MN && !B_MN), "complex mode: K-major operands");
  static_assert(!CT || EF_ == EPI_OUT_PLANES, "complex mode with the NC epilogue: plain split-plane output");
  static_assert(!PAIR || (!CPLX && (B_MN_ ? BNL % 64 == 0 : BNL % 16 == 0)), "pair mode");
  static_assert(!SP || (PAIR && A_MN_ && !B_MN_ && !NC_ && !CPLX_ && !BFLY_), "SP mode: CTA pairs, MN-major activations x K-major weights");
  static constexpr int SMEM_BYTES = STAGES * STAGE + STG_BYTES + BAR_BYTES + 1024;  // + alignment slack
  static_assert(BN % (B_MN_ ? 64 : 32) == 0 && BN <= 256, "BN");
  static_assert(!BFLY || (NC && !CPLX && !B_MN), "butterfly mode: NC epilogue, K-major B");
  static_assert(!CLN_ || (NC_ && !CPLX_ && !PAIR_ && !BFLY_ && !SP_ && !A_MN_ && !B_MN_ && EF_ == (EPI_RES_PLANES | EPI_OUT_PLANES)), "CLN mode");
  static constexpr uint32_t K_LAYOUT = (BK == 64) ? 2u : 4u;  // K-major tiles: 128B swizzle (BK = 64) or 64B (BK = 32)
  static_assert(BK == 32 || BK == 64, "BK");
  static_assert(STAGES >= 2, "pipeline too shallow");
  static_assert(A_PLANE % 1024 == 0 && B_PLANE % 1024 == 0, "tiles must keep 1024B alignment");
  static_assert(SMEM_BYTES <= kMaxSmem, "shared memory budget");
};

struct Tile {
  int m0, n_begin, n_count, n_end, z1, z2, k_begin, num_kc, tn;
};

template <int BN, int BK, int TILE_M = 128>
__device__ __forceinline__ bool decode_tile(const UmmaParams& p, long long t, Tile& ti) {
  const GemmOp& op = p.op;
  int tn, tm;
  long long r;
  if (p.tiles) {
    const int4 e = __ldg(p.tiles + t);
    tm = e.x;
    tn = e.y;
    r = (long long)e.z * op.Z2 + e.w;
  } else if (p.m_fastest) {
    tm = (int)(t % p.tiles_m);
    r = t / p.tiles_m;
    tn = (int)(r % p.tiles_n);
    r /= p.tiles_n;
  } else {
    tn = (int)(t % p.tiles_n);
    r = t / p.tiles_n;
    tm = (int)(r % p.tiles_m);
    r /= p.tiles_m;
  }
  ti.z2 = (int)(r % op.Z2);
  ti.z1 = (int)(r / op.Z2);
  const int n_lo = op.n_lo_z1 ? ti.z1 : 0;
  const int n_hi = op.n_hi_z1 ? min(op.N, ti.z1 + 1) : op.N;
  const int k_lo = op.k_lo_z1 ? ti.z1 : 0;
  ti.tn = tn;
  ti.m0 = tm * TILE_M;  // pair mode: the caller adds 128 * cluster rank
  ti.n_begin = max(tn * BN, (n_lo / 16) * 16);
  ti.n_end = min(tn * BN + BN, n_hi);
  ti.n_count = ti.n_end - ti.n_begin;
  ti.k_begin = (k_lo / BK) * BK;
  ti.num_kc = (op.K - ti.k_begin + BK - 1) / BK;
  if (op.m_hi_z1 && ti.m0 > ti.z1) return false;  // all rows of the tile are orders m > degree z1
  return ti.n_count > 0 && ti.n_end > n_lo && ti.num_kc > 0;
}

// ---- packed fp32x2 arithmetic (sm_100: one FFMA2 / FADD2 / FMUL2 per register pair) ----
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
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.u) : "l"(a.u), "l"(b.u));
  return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.u) : "l"(a.u), "l"(b.u));
  return d;
}
__device__ __forceinline__ float ex2f(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
}
// gelu_fast (common.cuh) on a pair: same polynomial, packed FMAs
__device__ __forceinline__ void gelu2(float& x0, float& x1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  f2 t = mul2(mk2(a0, a1), mk2(0.70710678118654752440f, 0.70710678118654752440f));
  float t0, t1;
  un2(t, t0, t1);
  t = mk2(fminf(t0, 4.6f), fminf(t1, 4.6f));
  f2 q = mk2(4.565055586e-07f, 4.565055586e-07f);
  q = fma2(q, t, mk2(-1.097697806e-05f, -1.097697806e-05f));
  q = fma2(q, t, mk2(1.118192688e-04f, 1.118192688e-04f));
  q = fma2(q, t, mk2(-6.078477993e-04f, -6.078477993e-04f));
  q = fma2(q, t, mk2(1.635471654e-03f, 1.635471654e-03f));
  q = fma2(q, t, mk2(8.210374325e-04f, 8.210374325e-04f));
  q = fma2(q, t, mk2(-2.841062484e-02f, -2.841062484e-02f));
  q = fma2(q, t, mk2(1.485603089e-01f, 1.485603089e-01f));
  q = fma2(q, t, mk2(9.184083273e-01f, 9.184083273e-01f));
  q = fma2(q, t, mk2(1.627907927e+00f, 1.627907927e+00f));
  float u0, u1;
  un2(mul2(q, t), u0, u1);
  const f2 e = mk2(ex2f(-u0), ex2f(-u1));
  const f2 r = fma2(mul2(mk2(a0, a1), mk2(-0.5f, -0.5f)), e, mk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  un2(r, x0, x1);
}

// two fp32 -> packed bf16x2 (element 0 in the low half), round to nearest even: one F2FP instruction
__device__ __forceinline__ uint32_t pack_bf16x2(float x0, float x1) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  return *reinterpret_cast<uint32_t*>(&h);
}
// split (x0, x1) into packed hi and lo words: v ~= hi + lo, |v - (hi + lo)| <= 2^-17 |v|
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(x0, x1);
  float r0, r1;
  un2(add2(mk2(x0, x1), mk2(-__uint_as_float(hi << 16), -__uint_as_float(hi & 0xffff0000u))), r0, r1);
  lo = pack_bf16x2(r0, r1);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- epilogue: ROWC (rows contiguous in memory), split-plane output only ----
// thread = accumulator row; each plane of a 32x32 chunk is transposed through the warp's 2 KB staging tile
// ([column][32 rows x 2 B]) so that the global stores are 8 bytes per lane, 64 contiguous bytes per column.
template <class C>
__device__ __forceinline__ void epilogue_rowc(const UmmaParams& p, const Tile& ti, uint32_t tacc, int q, int sub, int lane,
                                              uint8_t* stg_raw) {
  const GemmOp& op = p.op;
  const EpiParams& e = op.epi;
  unsigned short* s2 = reinterpret_cast<unsigned short*>(stg_raw);  // [32 columns][32 rows]
  const uint2* s8 = reinterpret_cast<const uint2*>(stg_raw);        // [32 columns][8 x (4 rows)]
  const int row0 = ti.m0 + 32 * q;
  // cooperative role: column cc + 4*it, rows row0 + 4*pc .. +3 (never straddles an mdiv boundary: mdiv % 4 == 0)
  const int cc = lane >> 3, pc = lane & 7;
  const int grow = row0 + pc * 4;
  const bool g_ok = grow < op.M;  // M % 4 == 0 (host-checked)
  const int m1 = grow / e.mdiv, mr = grow - m1 * e.mdiv;
  bf16* gbase = e.out + epi_z1_off(e, ti.z1) + (long long)ti.z2 * e.o_z2 + (long long)m1 * e.o_m1 + mr +
                (long long)(ti.n_begin + cc) * e.o_n;
  const int nch = (ti.n_count + 31) >> 5;  // 32-column chunks per accumulator
  const long long on4 = 4 * e.o_n;
  // element-wise fallback (rows of a 4-row group not contiguous / not 8-byte aligned): thread = its own row
  unsigned short* own = nullptr;
  if (p.scalar_store) {
    const int orow2 = min(row0 + lane, op.M - 1);
    const int om1 = orow2 / e.mdiv, omr = orow2 - om1 * e.mdiv;
    own = reinterpret_cast<unsigned short*>(e.out) + epi_z1_off(e, ti.z1) + (long long)ti.z2 * e.o_z2 + (long long)om1 * e.o_m1 + omr +
          (long long)ti.n_begin * e.o_n;
  }
  // deferred InstanceNorm of the A operand: per-row scale, constant folded into column 0 (thread = its own row)
  float a_scale = 1.f, a_shift0 = 0.f;
  if (e.flags & EPI_ROW_AFFINE) {
    const int orow = min(row0 + lane, op.M - 1);
    const long long i = (long long)ti.z2 * e.ra_z2 + orow / e.mdiv;
    a_scale = __ldg(e.ra_scale + i);
    a_shift0 = __ldg(e.ra_shift0 + i);
  }
  for (int cc2 = sub; cc2 < C::NACC * nch; cc2 += kEpiWarps / 4) {
    // complex mode: the second accumulator (imaginary rows) sits BN columns further and lands op.M rows further
    const int part = (C::NACC == 2 && cc2 >= nch) ? 1 : 0;
    const int c = cc2 - part * nch;
    float v[32];
    ptx::tmem_ld_32x32(tacc + part * C::BN + c * 32, v);
    ptx::tmem_ld_wait();
    if (e.flags & EPI_ROW_AFFINE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= a_scale;
      if (ti.n_begin + c * 32 == 0) v[0] += a_shift0;
    }
    uint32_t hw[16], lw[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
    const int nleft = ti.n_count - c * 32;  // columns of this chunk inside [n_begin, n_end)
    if (p.scalar_store) {
      if (row0 + lane < op.M) {
        unsigned short* h = own + (long long)(c * 32) * e.o_n + (long long)part * op.M;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (2 * j < nleft) {
            h[0] = (unsigned short)hw[j];
            h[e.out_plane] = (unsigned short)lw[j];
          }
          if (2 * j + 1 < nleft) {
            h[e.o_n] = (unsigned short)(hw[j] >> 16);
            h[e.o_n + e.out_plane] = (unsigned short)(lw[j] >> 16);
          }
          h += 2 * e.o_n;
        }
      }
      continue;
    }
    bf16* g = gbase + (long long)(c * 32) * e.o_n + (long long)part * op.M;
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t w = pl ? lw[j] : hw[j];
        s2[(2 * j) * 32 + lane] = (unsigned short)w;
        s2[(2 * j + 1) * 32 + lane] = (unsigned short)(w >> 16);
      }
      __syncwarp();
      // all eight shared-memory reads first, then the stores: issued pairwise (as ptxas orders the lo-plane pass when the
      // loads sit inside the predicated store loop) every STG waits for its own LDS
      uint2 w[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) w[it] = s8[(it * 4 + cc) * 8 + pc];
#pragma unroll
      for (int it = 0; it < 8; ++it)
        if (g_ok && it * 4 + cc < nleft) *reinterpret_cast<uint2*>(g + it * on4 + (pl ? e.out_plane : 0)) = w[it];
      __syncwarp();
    }
  }
}

// ---- epilogue: NC (thread = row, columns contiguous in memory, transposed through shared memory) ----
// A warp owns a 2 KB staging tile and moves its 32x32 chunk through it in passes:
//   fp32 pass   16 columns: [32 rows][4 x 16 B], slot = r*4 + (p ^ ((r >> 1) & 3))
//   plane pass  32 columns of ONE plane: [32 rows][8 x 8 B], slot = r*8 + (p ^ ((r >> 1) & 7))
// Row pitch is 64 B in both views; own-row accesses (lane = row) and cooperative accesses (lane -> (row, piece),
// whole 64 B row segments per quarter / half warp) are bank-conflict free with these swizzles.
__device__ __forceinline__ int swz16(int r, int p) { return r * 4 + (p ^ ((r >> 1) & 3)); }
__device__ __forceinline__ int swz8(int r, int p) { return r * 8 + (p ^ ((r >> 1) & 7)); }

// L2 prefetch of the fp32 addend / residual planes a warp will need for its chunks of this tile; issued before the
// warp blocks on the accumulator barrier, i.e. several microseconds ahead of use.
template <class C>
__device__ __forceinline__ void epilogue_nc_prefetch(const UmmaParams& p, const Tile& ti, int q, int sub, int lane) {
  constexpr uint32_t EF = C::EF;
  if (!(EF & (EPI_ADD_F32 | EPI_RES_PLANES))) return;
  const EpiParams& e = p.op.epi;
  const int row = ti.m0 + 32 * q + lane;
  if (row >= p.op.M) return;
  const int pm1 = row / e.mdiv, pm0 = row - pm1 * e.mdiv;
  for (int c = sub; c * 32 < ti.n_count; c += kEpiWarps / 4) {
    const long long n0 = ti.n_begin + c * 32;
    if (EF & EPI_ADD_F32) prefetch_l2(e.add + (long long)ti.z2 * e.add_z2 + (long long)pm1 * e.add_m1 + (long long)pm0 * e.add_m0 + n0);
    if (EF & EPI_RES_PLANES) {
      const bf16* r = e.res + (long long)ti.z2 * e.res_z2 + (long long)pm1 * e.res_m1 + (long long)pm0 * e.res_m0 + n0;
      prefetch_l2(r);
      prefetch_l2(r + e.res_plane);
    }
  }
}

template <class C, bool FULL>
__device__ __forceinline__ void epilogue_nc_chunk(const EpiParams& e, float (&v)[32], int nvalid, int rows_valid, int lane,
                                                  uint4* s16, uint2* s8, const float* g_add, const bf16* g_res,
                                                  float* g_f32, bf16* g_pl, float rbias, float res_a, float res_s, bool do_stats,
                                                  f2& ssum, f2& ssq, bool scalar_store, int dbg = 0) {
  constexpr uint32_t EF = C::EF;
  const int fr = lane >> 2, fp = lane & 3;  // fp32 pass: rows it*8 + fr (it < 4), 16-byte piece fp
  const int pr = lane >> 3, pp = lane & 7;  // plane pass: rows it*4 + pr (it < 8), 8-byte piece pp
  if (dbg & 8) do_stats = false;  // development switches (option "dbg"): 8 no statistics, 16 no addend / residual, 64 no stores

  // All global loads of the chunk are issued before the first one is consumed (one L2 round trip per chunk instead of one per
  // 16-column half / per plane: the epilogue warps are latency-bound, three of them share a scheduler).
  if ((EF & EPI_ADD_F32) && !(dbg & 16)) {
    uint4 t[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        t[h][it] = make_uint4(0u, 0u, 0u, 0u);
        if (FULL || (16 * h + fp * 4 < nvalid && it * 8 + fr < rows_valid))
          t[h][it] = __ldg(reinterpret_cast<const uint4*>(g_add + (long long)(it * 8) * e.add_m0 + 16 * h));
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int it = 0; it < 4; ++it) s16[swz16(it * 8 + fr, fp)] = t[h][it];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 w = s16[swz16(lane, k)];
        v[16 * h + 4 * k + 0] += __uint_as_float(w.x);
        v[16 * h + 4 * k + 1] += __uint_as_float(w.y);
        v[16 * h + 4 * k + 2] += __uint_as_float(w.z);
        v[16 * h + 4 * k + 3] += __uint_as_float(w.w);
      }
      __syncwarp();
    }
  }
  if ((EF & EPI_RES_PLANES) && !(dbg & 16)) {
    // (one plane at a time: holding both planes' loads in registers spills in the RS | P variant and measured slower, fc2 114 -> 118 us)
    const f2 ra2 = mk2(res_a, res_a);
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      uint2 t[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        t[it] = make_uint2(0u, 0u);
        if (FULL || (pp * 4 < nvalid && it * 4 + pr < rows_valid))
          t[it] = __ldg(reinterpret_cast<const uint2*>(g_res + (long long)(it * 4) * e.res_m0 + (pl ? e.res_plane : 0)));
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) s8[swz8(it * 4 + pr, pp)] = t[it];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint2 a = s8[swz8(lane, k)];
        // bf16 -> fp32 is a 16-bit shift; element 2i sits in the low half of word i
        f2 s0 = fma2(ra2, mk2(__uint_as_float(a.x << 16), __uint_as_float(a.x & 0xffff0000u)), mk2(v[4 * k], v[4 * k + 1]));
        f2 s1 = fma2(ra2, mk2(__uint_as_float(a.y << 16), __uint_as_float(a.y & 0xffff0000u)), mk2(v[4 * k + 2], v[4 * k + 3]));
        un2(s0, v[4 * k], v[4 * k + 1]);
        un2(s1, v[4 * k + 2], v[4 * k + 3]);
      }
      __syncwarp();
    }
  }

  const f2 rb = mk2(rbias + res_s, rbias + res_s);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    f2 x = add2(mk2(v[j], v[j + 1]), rb);
    un2(x, v[j], v[j + 1]);
    if (EF & EPI_GELU) gelu2(v[j], v[j + 1]);
  }
  if (do_stats) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const bool ok = FULL || j < nvalid;  // nvalid is even
      const f2 x = mk2(ok ? v[j] : 0.f, ok ? v[j + 1] : 0.f);
      ssum = add2(ssum, x);
      ssq = fma2(x, x, ssq);
    }
  }

  if (dbg & 64) return;
  if (EF & EPI_OUT_F32) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        s16[swz16(lane, k)] = make_uint4(__float_as_uint(v[16 * h + 4 * k]), __float_as_uint(v[16 * h + 4 * k + 1]),
                                         __float_as_uint(v[16 * h + 4 * k + 2]), __float_as_uint(v[16 * h + 4 * k + 3]));
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const uint4 w = s16[swz16(it * 8 + fr, fp)];
        if (FULL || (16 * h + fp * 4 < nvalid && it * 8 + fr < rows_valid))
          *reinterpret_cast<uint4*>(g_f32 + (long long)(it * 8) * e.f_m0 + 16 * h) = w;
      }
      __syncwarp();
    }
  }
  if (EF & EPI_OUT_PLANES) {
    // pass 0 stores the hi plane and leaves the residuals (v - hi) in v; pass 1 stores their bf16 rounding (lo)
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        uint32_t w[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float& x0 = v[4 * k + 2 * i];
          float& x1 = v[4 * k + 2 * i + 1];
          w[i] = pack_bf16x2(x0, x1);
          if (pl == 0) un2(add2(mk2(x0, x1), mk2(-__uint_as_float(w[i] << 16), -__uint_as_float(w[i] & 0xffff0000u))), x0, x1);
        }
        s8[swz8(lane, k)] = make_uint2(w[0], w[1]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint2 w = s8[swz8(it * 4 + pr, pp)];
        if (!FULL && scalar_store) {
          // row starts not 8-byte aligned / N not a multiple of 4 (odd nlat): same coalescing, 2-byte stores
          if (it * 4 + pr < rows_valid) {
            unsigned short* d = reinterpret_cast<unsigned short*>(g_pl + (long long)(it * 4) * e.o_m0 + (pl ? e.out_plane : 0));
            if (pp * 4 + 0 < nvalid) d[0] = (unsigned short)w.x;
            if (pp * 4 + 1 < nvalid) d[1] = (unsigned short)(w.x >> 16);
            if (pp * 4 + 2 < nvalid) d[2] = (unsigned short)w.y;
            if (pp * 4 + 3 < nvalid) d[3] = (unsigned short)(w.y >> 16);
          }
        } else if (FULL || (pp * 4 < nvalid && it * 4 + pr < rows_valid))
          *reinterpret_cast<uint2*>(g_pl + (long long)(it * 4) * e.o_m0 + (pl ? e.out_plane : 0)) = w;
      }
      __syncwarp();
    }
  }
}

template <class C>
__device__ __forceinline__ void epilogue_nc(const UmmaParams& p, const Tile& ti, uint32_t tacc, int q, int sub, int lane,
                                            uint8_t* stg_raw) {
  constexpr uint32_t EF = C::EF;
  const GemmOp& op = p.op;
  const EpiParams& e = op.epi;
  uint4* s16 = reinterpret_cast<uint4*>(stg_raw);
  uint2* s8 = reinterpret_cast<uint2*>(stg_raw);
  const int row0 = ti.m0 + 32 * q;  // first row of this warp
  const int row = row0 + lane;
  const int m_hi = op.m_hi_z1 ? min(op.M, ti.z1 + 1) : op.M;
  const bool row_ok = row < m_hi;
  int rows_valid = m_hi - row0;  // >= 32 for full tiles
  // rows flattened as (m1, m0) with m0 < mdiv (mdiv % 32 == 0, host-checked: a warp's 32 rows share m1) and rows
  // m0 >= mrows being padding that is never stored
  const int rm1 = row0 / e.mdiv, rm0 = row0 - rm1 * e.mdiv;
  if (e.mrows) rows_valid = min(rows_valid, e.mrows - rm0);
  const long long z1off = epi_z1_off(e, ti.z1);
  const bool do_stats = (e.flags & EPI_ROW_STATS) != 0;
  float rbias = 0.f;
  if ((e.flags & EPI_ROW_BIAS) && row_ok) rbias = __ldg(e.row_bias + (long long)ti.z2 * e.rb_z2 + row);
  float res_a = 1.f, res_s = 0.f;  // deferred InstanceNorm of the residual: r -> res_a * r + res_s (per row = channel)
  if ((EF & EPI_RES_PLANES) && (e.flags & EPI_RES_AFFINE) && row_ok) {
    res_a = __ldg(e.res_a + (long long)ti.z2 * e.rsa_z2 + row);
    res_s = __ldg(e.res_s + (long long)ti.z2 * e.rsa_z2 + row);
  }
  f2 ssum = mk2(0.f, 0.f), ssq = mk2(0.f, 0.f);
  const int fr = lane >> 2, fp = lane & 3, pr = lane >> 3, pp = lane & 7;
  // per-lane global pointers at (first cooperative row, n_begin, piece)
  const float* g_add = nullptr;
  const bf16* g_res = nullptr;
  float* g_f32 = nullptr;
  bf16* g_pl = nullptr;
  if (EF & EPI_ADD_F32)
    g_add = e.add + (long long)ti.z2 * e.add_z2 + (long long)rm1 * e.add_m1 + (long long)(rm0 + fr) * e.add_m0 + ti.n_begin + fp * 4;
  if (EF & EPI_RES_PLANES)
    g_res = e.res + (long long)ti.z2 * e.res_z2 + (long long)rm1 * e.res_m1 + (long long)(rm0 + pr) * e.res_m0 + ti.n_begin + pp * 4;
  if (EF & EPI_OUT_F32)
    g_f32 = e.outf + (long long)ti.z1 * e.f_z1 + (long long)ti.z2 * e.f_z2 + (long long)rm1 * e.f_m1 + (long long)(rm0 + fr) * e.f_m0 +
            ti.n_begin + fp * 4;
  if (EF & EPI_OUT_PLANES)
    g_pl = e.out + z1off + (long long)ti.z2 * e.o_z2 + (long long)rm1 * e.o_m1 + (long long)(rm0 + pr) * e.o_m0 + ti.n_begin + pp * 4;

  const int nch = (ti.n_count + 31) >> 5;
  for (int cc2 = sub; cc2 < C::NACC * nch; cc2 += kEpiWarps / 4) {
    // complex mode: the second accumulator (imaginary part) sits BN TMEM columns further and lands op.N columns further;
    // butterfly mode: pass 0 stores E + O at column n, pass 1 stores E - O at column n + N
    const int part = (C::NACC == 2 && cc2 >= nch) ? 1 : 0;
    const int c = cc2 - part * nch;
    if (rows_valid <= 0) break;  // (triangular M range / padding rows) nothing of this warp's rows exists
    const int nvalid = min(32, ti.n_count - c * 32);  // multiple of 4 (host-checked) unless scalar_store
    float v[32];
    if constexpr (C::BFLY) {
      float u[32];
      ptx::tmem_ld_32x32(tacc + c * 32, v);
      ptx::tmem_ld_32x32(tacc + C::BN + c * 32, u);
      ptx::tmem_ld_wait();
      const f2 sgn = part ? mk2(-1.f, -1.f) : mk2(1.f, 1.f);
#pragma unroll
      for (int j = 0; j < 32; j += 2) un2(fma2(sgn, mk2(u[j], u[j + 1]), mk2(v[j], v[j + 1])), v[j], v[j + 1]);
    } else {
      ptx::tmem_ld_32x32(tacc + part * C::BN + c * 32, v);
      ptx::tmem_ld_wait();
    }
    const int co = c * 32 + part * op.N;
    if (nvalid == 32 && rows_valid >= 32 && !p.scalar_store)
      epilogue_nc_chunk<C, true>(e, v, nvalid, rows_valid, lane, s16, s8, g_add + co, g_res + co, g_f32 + co, g_pl + co, rbias,
                                 res_a, res_s, do_stats, ssum, ssq, false, p.dbg);
    else
      epilogue_nc_chunk<C, false>(e, v, nvalid, rows_valid, lane, s16, s8, g_add + co, g_res + co, g_f32 + co, g_pl + co, rbias,
                                  res_a, res_s, do_stats && row_ok, ssum, ssq, p.scalar_store != 0, p.dbg);
  }
  if (do_stats && row_ok) {
    float s0, s1, q0, q1;
    un2(ssum, s0, s1);
    un2(ssq, q0, q1);
    double* st = e.stats + ((long long)ti.z2 * e.stats_z2 + row) * 2;
    atomicAdd(st, (double)(s0 + s1));
    atomicAdd(st + 1, (double)(q0 + q1));
  }
}

// ---- epilogue: CLN (ConditionalLayerNorm mode; thread = row = channel, columns = pixels, contiguous in memory) ----
//   acc0 = S (scale modulation), acc1 = T (bias modulation);  y = x a + o  with  a = rstd_n lnw_m (scale0_m + S),
//   o = lnb_m (scale0_m + S) - mean_n a + bias0_m + T;  x (the norm's input) is read plane by plane through the staging tile like
//   the residual of the NC epilogue, y leaves as split planes the same way.
template <class C>
__device__ __forceinline__ void epilogue_cln(const UmmaParams& p, const Tile& ti, uint32_t tacc, int q, int sub, int lane, uint8_t* stg_raw,
                                             float2* s_cols) {
  const GemmOp& op = p.op;
  const EpiParams& e = op.epi;
  uint2* s8 = reinterpret_cast<uint2*>(stg_raw);
  const int row0 = ti.m0 + 32 * q, row = row0 + lane;
  const bool row_ok = row < op.M;
  const int rows_valid = op.M - row0;
  if (rows_valid <= 0) return;
  const int rowc = row_ok ? row : op.M - 1;
  const float lnw = e.cln_lnw ? __ldg(e.cln_lnw + rowc) : 1.f, lnb = e.cln_lnb ? __ldg(e.cln_lnb + rowc) : 0.f;
  float sc0 = 1.f, bi0 = 0.f;
  if (e.cln_sb0) {
    const float2 sb = __ldg(reinterpret_cast<const float2*>(e.cln_sb0 + ((long long)ti.z2 * e.cln_sb0_z2 + rowc) * 2));
    sc0 = sb.x;
    bi0 = sb.y;
  }
  const int pr = lane >> 3, pp = lane & 7;  // plane pass: rows it*4 + pr (it < 8), 8-byte piece pp
  const bf16* g_res = e.res + (long long)ti.z2 * e.res_z2 + (long long)(row0 + pr) * e.res_m0 + ti.n_begin + pp * 4;
  bf16* g_pl = e.out + (long long)ti.z2 * e.o_z2 + (long long)(row0 + pr) * e.o_m0 + ti.n_begin + pp * 4;
  const float2* g_cols = e.cln_musr + (long long)ti.z2 * e.cln_musr_z2 + ti.n_begin;
  const int nch = (ti.n_count + 31) >> 5;
  for (int c = sub; c < nch; c += kEpiWarps / 4) {
    const int nvalid = min(32, ti.n_count - c * 32);  // multiple of 4 (host-checked)
    const bool full = nvalid == 32 && rows_valid >= 32;
    float a[32], o[32];
    ptx::tmem_ld_32x32(tacc + c * 32, a);
    ptx::tmem_ld_32x32(tacc + C::BN + c * 32, o);
    // the chunk's column statistics: one coalesced load, then warp-uniform shared-memory reads
    s_cols[lane] = (lane < nvalid) ? __ldg(g_cols + c * 32 + lane) : make_float2(0.f, 0.f);
    __syncwarp();
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float4 mr = *reinterpret_cast<const float4*>(s_cols + j);  // {mean_j, rstd_j, mean_j+1, rstd_j+1}
      const float t0 = sc0 + a[j], t1 = sc0 + a[j + 1];
      a[j] = mr.y * lnw * t0;
      a[j + 1] = mr.w * lnw * t1;
      o[j] = fmaf(lnb, t0, bi0 + o[j]) - mr.x * a[j];
      o[j + 1] = fmaf(lnb, t1, bi0 + o[j + 1]) - mr.z * a[j + 1];
    }
    // x planes: o += a * (hi + lo) through the staging tile; the loads of both planes are in flight together
    const bf16* gr = g_res + c * 32;
    uint2 t[2][8];
#pragma unroll
    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        t[pl][it] = make_uint2(0u, 0u);
        if (full || (pp * 4 < nvalid && it * 4 + pr < rows_valid))
          t[pl][it] = __ldg(reinterpret_cast<const uint2*>(gr + (long long)(it * 4) * e.res_m0 + (pl ? e.res_plane : 0)));
      }
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
      for (int it = 0; it < 8; ++it) s8[swz8(it * 4 + pr, pp)] = t[pl][it];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint2 w = s8[swz8(lane, k)];
        o[4 * k] = fmaf(a[4 * k], __uint_as_float(w.x << 16), o[4 * k]);
        o[4 * k + 1] = fmaf(a[4 * k + 1], __uint_as_float(w.x & 0xffff0000u), o[4 * k + 1]);
        o[4 * k + 2] = fmaf(a[4 * k + 2], __uint_as_float(w.y << 16), o[4 * k + 2]);
        o[4 * k + 3] = fmaf(a[4 * k + 3], __uint_as_float(w.y & 0xffff0000u), o[4 * k + 3]);
      }
      __syncwarp();
    }
    // split-plane store of o (pass 0: hi plane, residuals stay in o; pass 1: lo plane)
    bf16* gp = g_pl + c * 32;
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        uint32_t w[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float& x0 = o[4 * k + 2 * i];
          float& x1 = o[4 * k + 2 * i + 1];
          w[i] = pack_bf16x2(x0, x1);
          if (pl == 0) {
            x0 -= __uint_as_float(w[i] << 16);
            x1 -= __uint_as_float(w[i] & 0xffff0000u);
          }
        }
        s8[swz8(lane, k)] = make_uint2(w[0], w[1]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint2 w = s8[swz8(it * 4 + pr, pp)];
        if (full || (pp * 4 < nvalid && it * 4 + pr < rows_valid))
          *reinterpret_cast<uint2*>(gp + (long long)(it * 4) * e.o_m0 + (pl ? e.out_plane : 0)) = w;
      }
      __syncwarp();
    }
  }
}

// ---- epilogue: SP (thread = spatial position, accumulator columns = output channels; see Cfg::SP) ----
// EpiParams keep the caller's meaning (row m = channel, column n = spatial position); p.op is the exchanged op (op.M = positions,
// op.N = channels).

// sum over the 32 lanes of each of 32 per-lane values in 31 shuffles: on return lane l holds the total of element l
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? v[i] : v[i + w];
      const float keep = up ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

// operands of one 32-channel chunk for this thread's position, loaded into registers (issued a whole chunk ahead of their use:
// before the accumulator barrier for the first chunk of a tile, before the previous chunk's arithmetic for the others)
template <class C>
struct SpOperands {
  static constexpr uint32_t EF = C::EF;
  static constexpr int NA = (EF & EPI_ADD_F32) ? 32 : 1, NR = (EF & EPI_RES_PLANES) ? 32 : 1;
  float add[NA];
  uint32_t rh[NR], rl[NR];  // the 32-bit words that hold this position's hi / lo residual element
  __device__ __forceinline__ void load(const EpiParams& e, const float* g_add, const uint32_t* g_res, int nvalid) {
    if (EF & EPI_ADD_F32) {
#pragma unroll
      for (int j = 0; j < NA; ++j) add[j] = (j < nvalid) ? __ldg(g_add + (long long)j * e.add_m0) : 0.f;
    }
    if (EF & EPI_RES_PLANES) {
      // 32-bit loads (lane pairs share a word; 16-bit non-coherent loads run at a fraction of the rate) through the coherent
      // path: fc2 reads its residual from the buffer it overwrites in place
      const long long m0w = e.res_m0 >> 1, planew = e.res_plane >> 1;  // in words (both even: sp_eligible)
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const uint32_t* r = g_res + (long long)j * m0w;
        rh[j] = (j < nvalid) ? *r : 0u;
        rl[j] = (j < nvalid) ? r[planew] : 0u;
      }
    }
  }
  __device__ __forceinline__ void apply(float (&v)[32], bool odd, const float* res_a, const float* res_s) const {
    if (EF & EPI_ADD_F32) {
#pragma unroll
      for (int j = 0; j < NA; ++j) v[j] += add[j];
    }
    if (EF & EPI_RES_PLANES) {
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const float r = __uint_as_float(odd ? (rh[j] & 0xffff0000u) : (rh[j] << 16)) + __uint_as_float(odd ? (rl[j] & 0xffff0000u) : (rl[j] << 16));
        if (res_a) v[j] += fmaf(__ldg(res_a + j), r, __ldg(res_s + j));
        else v[j] += r;
      }
    }
  }
};

// everything after the operands: bias, GELU, stores (staging tile or element-wise), statistics
template <class C, bool FULL>
__device__ __forceinline__ void epilogue_sp_chunk(const EpiParams& e, float (&v)[32], int nvalid, bool sp_ok, int lane, float* g_f32,
                                                  unsigned short* g_pl, const float* bias, bool do_stats, float* s_stats,
                                                  unsigned short* s_tile, int dbg) {
  constexpr uint32_t EF = C::EF;
  if (dbg & 8) do_stats = false;
  {
    // `bias` = this warp's shared-memory copy of the chunk's 32 biases (zeros without a bias)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(bias + 4 * k);
      v[4 * k] += b.x;
      v[4 * k + 1] += b.y;
      v[4 * k + 2] += b.z;
      v[4 * k + 3] += b.w;
    }
  }
  if (EF & EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu2(v[j], v[j + 1]);
  }
  if (EF & EPI_OUT_F32) {
    if (sp_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (FULL || j < nvalid) g_f32[(long long)j * e.f_m0] = v[j];
    }
  }
  if (dbg & 64) {
  } else if ((EF & EPI_OUT_PLANES) && s_tile) {
    // staging tile [plane][channel][128 positions] (s_tile points at this thread's position): a warp writes 64 contiguous
    // bytes per channel and plane; the tile leaves by TMA (epilogue_sp)
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      uint32_t hw, lw;
      split2(v[j], v[j + 1], hw, lw);
      s_tile[j * 128] = (unsigned short)hw;
      s_tile[(j + 1) * 128] = (unsigned short)(hw >> 16);
      s_tile[32 * 128 + j * 128] = (unsigned short)lw;
      s_tile[32 * 128 + (j + 1) * 128] = (unsigned short)(lw >> 16);
    }
  } else if (EF & EPI_OUT_PLANES) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      uint32_t hw, lw;
      split2(v[j], v[j + 1], hw, lw);
      if (sp_ok) {
        if (FULL || j < nvalid) {
          unsigned short* d = g_pl + (long long)j * e.o_m0;
          d[0] = (unsigned short)hw;
          d[e.out_plane] = (unsigned short)lw;
        }
        if (FULL || j + 1 < nvalid) {
          unsigned short* d = g_pl + (long long)(j + 1) * e.o_m0;
          d[0] = (unsigned short)(hw >> 16);
          d[e.out_plane] = (unsigned short)(lw >> 16);
        }
      }
    }
  }
  if (do_stats) {
    // per-channel sums over this warp's 32 positions (rows beyond the image contribute nothing), then one shared-memory atomic
    // per channel and warp; the tile's totals go to the fp64 accumulators once per tile (epilogue_sp_flush)
    float w[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = sp_ok ? v[j] : 0.f;
      w[j] = v[j] * v[j];
    }
    const float sq = warp_colsum32(w, lane);
    const float sm = warp_colsum32(v, lane);
    if (FULL || lane < nvalid) {
      atomicAdd(s_stats + 2 * lane, sm);
      atomicAdd(s_stats + 2 * lane + 1, sq);
    }
  }
}

// waits for the tile's accumulator itself (tfull barrier) so that the first chunk's operand tile is in flight meanwhile.
// Per group of four warps (lane quarters q = 0..3 of the same chunk index `sub`) and chunk of 32 channels:
//   X tile <- TMA load of the addend / residual of the chunk (issued a chunk ahead by the group's first lane, mbarrier xbar)
//   v <- accumulator + X (shared-memory reads, thread = position: conflict free) ; barrier A: X consumed, Y free again
//   bias, GELU, statistics ; Y tile <- split planes ; barrier B ; TMA store of Y
template <class C>
__device__ __forceinline__ void epilogue_sp(const UmmaParams& p, const Tile& ti, uint32_t tacc, int q, int sub, int lane, uint8_t* stg_tiles,
                                            float* s_stats, float* s_vec, uint32_t xbar, uint32_t& xphase, uint32_t tfull, uint32_t tfull_parity) {
  constexpr uint32_t EF = C::EF;
  const GemmOp& op = p.op;
  const EpiParams& e = op.epi;
  const int sp = ti.m0 + 32 * q + lane;
  const bool sp_ok = sp < op.M;
  const long long spc = sp_ok ? sp : op.M - 1;  // loads of the lanes beyond the image stay in bounds (their results are never stored)
  const bool do_stats = (e.flags & EPI_ROW_STATS) != 0;
  const long long cb = ti.n_begin;  // first channel of the tile
  const float* g_add = nullptr;
  const uint32_t* g_res = nullptr;  // the 32-bit word that holds position spc (rows start on even elements: sp_eligible)
  float* g_f32 = nullptr;
  unsigned short* g_pl = nullptr;
  const bool odd = (spc & 1) != 0;
  if (EF & EPI_ADD_F32) g_add = e.add + (long long)ti.z2 * e.add_z2 + cb * e.add_m0 + spc;
  if (EF & EPI_RES_PLANES) g_res = reinterpret_cast<const uint32_t*>(e.res + (long long)ti.z2 * e.res_z2 + cb * e.res_m0) + (spc >> 1);
  if (EF & EPI_OUT_F32) g_f32 = e.outf + (long long)ti.z2 * e.f_z2 + cb * e.f_m0 + spc;
  if (EF & EPI_OUT_PLANES) g_pl = reinterpret_cast<unsigned short*>(e.out) + (long long)ti.z2 * e.o_z2 + cb * e.o_m0 + spc;
  const float* bias = (e.flags & EPI_ROW_BIAS) ? e.row_bias + (long long)ti.z2 * e.rb_z2 + cb : nullptr;
  const bool raff = (EF & EPI_RES_PLANES) && (e.flags & EPI_RES_AFFINE);
  const float* res_a = raff ? e.res_a + (long long)ti.z2 * e.rsa_z2 + cb : nullptr;
  const float* res_s = raff ? e.res_s + (long long)ti.z2 * e.rsa_z2 + cb : nullptr;
  const int nch = (ti.n_count + 31) >> 5;
  const long long res_m0w = e.res_m0 >> 1;
  const bool staged = (EF & EPI_OUT_PLANES) && p.sp_tma;
  const bool xtile = C::SP_OPND && p.sp_tmx && staged && !(p.dbg & 16);  // operands by TMA (else: loads from registers' side, SpOperands)
  unsigned short* tile = reinterpret_cast<unsigned short*>(stg_tiles + sub * C::SP_TILE);
  uint8_t* xt = C::SP_SHARED ? reinterpret_cast<uint8_t*>(tile) : stg_tiles + C::SP_X + sub * C::SP_TILE;
  const bool issuer = q == 0 && lane == 0;
  const bool skip = (p.dbg & 1) != 0;
  auto load_x = [&](int c) {  // (issuer only) the X tile is free: every thread of the group has passed barrier A of the previous chunk
    ptx::mbar_arrive_expect_tx(xbar, (uint32_t)C::SP_TILE);
    ptx::tma_load_4d(ptx::smem_u32(xt), &p.tmX, xbar, ti.m0, ti.n_begin + c * 32, ti.z2, 0);
  };
  // (SP_SHARED: the issuer waited for the previous store's shared-memory reads right after issuing it)
  if (xtile && issuer && sub < nch && !skip) load_x(sub);
  ptx::mbar_wait(tfull, tfull_parity);
  ptx::tc_fence_after();
  if (skip) return;
  float* vec = s_vec + ((sub * 4 + q) * 3) * 32;  // this warp's [bias | residual scale | residual shift] of the chunk
  for (int c = sub; c < nch; c += kEpiWarps / 4) {
    const int nvalid = min(32, ti.n_count - c * 32);
    float v[32];
    ptx::tmem_ld_32x32(tacc + c * 32, v);
    const long long co = (long long)c * 32;
    // the chunk's per-channel vectors: one coalesced load per vector and warp instead of broadcast loads per channel (with
    // the whole shared-memory carve-out in use there is no L1 to serve those)
    if (!(p.dbg & 32)) {
      const bool ok = lane < nvalid;
      vec[lane] = (bias && ok) ? __ldg(bias + co + lane) : 0.f;
      if (EF & EPI_RES_PLANES) {
        vec[32 + lane] = (res_a && ok) ? __ldg(res_a + co + lane) : 1.f;
        vec[64 + lane] = (res_s && ok) ? __ldg(res_s + co + lane) : 0.f;
      }
    } else {
      vec[lane] = 0.f;
      vec[32 + lane] = 1.f;
      vec[64 + lane] = 0.f;
    }
    __syncwarp();
    ptx::tmem_ld_wait();
    if (xtile) {
      ptx::mbar_wait(xbar, xphase);
      xphase ^= 1u;
      if (EF & EPI_ADD_F32) {
        const float* xa = reinterpret_cast<const float*>(xt) + 32 * q + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += xa[j * 128];
      }
      if (EF & EPI_RES_PLANES) {
        const unsigned short* xr = reinterpret_cast<const unsigned short*>(xt) + 32 * q + lane;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 ra = *reinterpret_cast<const float4*>(vec + 32 + j), rs = *reinterpret_cast<const float4*>(vec + 64 + j);
          const float a4[4] = {ra.x, ra.y, ra.z, ra.w}, s4[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float r = __uint_as_float((uint32_t)xr[(j + k) * 128] << 16) + __uint_as_float((uint32_t)xr[32 * 128 + (j + k) * 128] << 16);
            v[j + k] += fmaf(a4[k], r, s4[k]);
          }
        }
      }
    } else if ((EF & (EPI_ADD_F32 | EPI_RES_PLANES)) && !(p.dbg & 16)) {
      SpOperands<C> ops;
      ops.load(e, g_add + co * e.add_m0, g_res + co * res_m0w, nvalid);
      ops.apply(v, odd, res_a ? res_a + co : nullptr, res_s ? res_s + co : nullptr);
    }
    if (staged) {
      if (issuer && !(C::SP_SHARED && xtile)) ptx::bulk_wait_read0();  // the previous store has finished reading the Y tile
      ptx::named_bar_sync(2 + sub, 128);   // barrier A
      const int cn = c + kEpiWarps / 4;
      if (!C::SP_SHARED && xtile && issuer && cn < nch) load_x(cn);
    }
    unsigned short* s_tile = staged ? tile + 32 * q + lane : nullptr;
    if (nvalid == 32)
      epilogue_sp_chunk<C, true>(e, v, nvalid, sp_ok, lane, g_f32 + co * e.f_m0, g_pl + co * e.o_m0, vec, do_stats, s_stats + 2 * co, s_tile, p.dbg);
    else
      epilogue_sp_chunk<C, false>(e, v, nvalid, sp_ok, lane, g_f32 + co * e.f_m0, g_pl + co * e.o_m0, vec, do_stats, s_stats + 2 * co, s_tile, p.dbg);
    if (staged) {
      ptx::fence_proxy_async_smem();  // the tile was written through the generic proxy, TMA reads it through the async proxy
      ptx::named_bar_sync(2 + sub, 128);  // barrier B
      if (issuer) {
        // rows beyond the image and channels beyond the tensor are clipped by TMA
        ptx::tma_store_4d(&p.tmO, ptx::smem_u32(tile), ti.m0, ti.n_begin + c * 32, ti.z2, 0);
        ptx::bulk_commit();
        if (C::SP_SHARED && xtile) {
          // the next residual tile lands in the tile this store reads from
          ptx::bulk_wait_read0();
          const int cn = c + kEpiWarps / 4;
          if (cn < nch) load_x(cn);
        }
      }
    }
  }
}

// the tile's per-channel sums (all twelve epilogue warps have added theirs) -> fp64 accumulators; the staging array is zero again
// afterwards.  Named barrier 1 = the 384 epilogue threads.
template <class C>
__device__ __forceinline__ void epilogue_sp_flush(const UmmaParams& p, const Tile& ti, float* s_stats) {
  const EpiParams& e = p.op.epi;
  asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
  for (int i = (int)threadIdx.x - 128; i < 2 * C::BN; i += 32 * kEpiWarps) {
    const int col = i >> 1;
    if (col < ti.n_count) {
      const float val = s_stats[i];
      s_stats[i] = 0.f;
      atomicAdd(e.stats + ((long long)ti.z2 * e.stats_z2 + ti.n_begin + col) * 2 + (i & 1), (double)val);
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
}

template <class C>
__global__ void __launch_bounds__(kThreadsUmma, 1) gemm_umma_kernel(const __grid_constant__ UmmaParams p) {
  constexpr int BN = C::BN, BK = C::BK, STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;  // swizzle atoms need 1024B alignment
  uint8_t* smem = smem_raw + (sbase - raw);
  uint8_t* stg_all = smem + STAGES * C::STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_all + C::STG_BYTES);
  const uint32_t bar0 = sbase + STAGES * C::STAGE + C::STG_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * STAGES + 4);
  auto x_bar = [&](int g) { return bar0 + 8u * (2 * STAGES + 6 + g); };  // SP: operand tile of epilogue group g has landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmOp& op = p.op;
  constexpr bool PAIR = C::PAIR;
  // trace header row (iteration kTraceIters - 1): {globaltimer, clock64} at kernel entry / after the set-up / at the end
  if (p.trace && threadIdx.x == 0) {
    trace_put(p, kTraceIters - 1, 0, globaltimer_ns());
    trace_put(p, kTraceIters - 1, 1, clock64());
  }
  const uint32_t crank = PAIR ? ptx::cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs), 1 = peer
  constexpr int NCTA = PAIR ? 2 : 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmA);
    ptx::prefetch_tensormap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), NCTA);  // pair: one arrive per CTA's producer, all bytes credited to the leader
      ptx::mbar_init(empty_bar(s), 1);
    }
    if constexpr (C::SP) {
      for (int g = 0; g < kEpiWarps / 4; ++g) ptx::mbar_init(x_bar(g), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), NCTA * kEpiWarps);  // one arrive per epilogue warp (of both CTAs)
    }
    ptx::fence_barrier_init();
  }
  if constexpr (C::SP) {
    for (int i = threadIdx.x; i < 2 * BN; i += kThreadsUmma) reinterpret_cast<float*>(stg_all + C::SP_STATS)[i] = 0.f;
    if (warp == 0 && lane == 0) {
      ptx::prefetch_tensormap(&p.tmO);
      ptx::prefetch_tensormap(&p.tmX);
    }
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      ptx::tmem_alloc_2sm(ptx::smem_u32((const void*)tmem_slot), C::TMEM_COLS);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(ptx::smem_u32((const void*)tmem_slot), C::TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync();  // barriers of both CTAs exist before any remote arrive / multicast commit
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel in the stream; its
  // results (and the buffers it was still reading) may only be touched after this point
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p.trace && threadIdx.x == 0) {
    trace_put(p, kTraceIters - 1, 2, globaltimer_ns());
    trace_put(p, kTraceIters - 1, 3, clock64());
  }

  const long long total = p.tiles ? p.n_listed : (long long)p.tiles_m * p.tiles_n * op.Z1 * op.Z2;
  const long long t_first = PAIR ? (blockIdx.x >> 1) : blockIdx.x, t_step = PAIR ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    uint32_t it = 0;
    Tile ti;
    for (long long t = t_first; t < total; t += t_step) {
      if (!decode_tile<BN, BK, C::TILE_M>(p, t, ti)) continue;
      const int az1 = ti.z1 * p.a_z1_on, az2 = ti.z2 * p.a_z2_on, bz1 = ti.z1 * p.b_z1_on, bz2 = ti.z2 * p.b_z2_on;
      long long tr_wait = 0;
      if (p.trace && lane == 0) trace_put(p, it, 4, clock64());
      for (int kc = 0; kc < ti.num_kc; ++kc) {
        if (p.trace) {
          const long long c0 = clock64();
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          tr_wait += clock64() - c0;
        } else
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        if (ptx::elect_one()) {
          const uint32_t fb = full_bar(stage);
          if constexpr (PAIR) {
            // both CTAs fill their own stage; every byte is credited to the leader's barrier, which also collects one
            // arrive per CTA
            if (crank == 0) ptx::mbar_arrive_expect_tx(fb, (uint32_t)(2 * C::STAGE));
            else ptx::mbar_arrive_leader(fb);
            const uint32_t sA = sbase + stage * C::STAGE;
            const uint32_t sB = sA + 2 * C::A_PLANE;
            const int k0 = ti.k_begin + kc * BK;
            // the MMA takes the first N/2 accumulator columns from the leader's B tile and the rest from the peer's
            const int n_eff = (ti.n_count + 15) & ~15;
            const int mrow = ti.m0 + 128 * (int)crank, ncol = ti.n_begin + (n_eff >> 1) * (int)crank;
            // one box carries both planes; MN-major tiles are stored [64-wide atom][plane][k][64]
            if (!C::A_MN) {
              ptx::tma_load_5d_2sm(sA, &p.tmA, fb, k0, mrow, az1, az2, 0);
            } else {
#pragma unroll
              for (int a = 0; a < 2; ++a) ptx::tma_load_5d_2sm(sA + a * C::MN_ATOM, &p.tmA, fb, mrow + 64 * a, k0, az1, az2, 0);
            }
            if (!C::B_MN) {
              ptx::tma_load_5d_2sm(sB, &p.tmB, fb, k0, ncol, bz1, bz2, 0);
            } else {
#pragma unroll
              for (int a = 0; a < C::BNL / 64; ++a) ptx::tma_load_5d_2sm(sB + a * C::MN_ATOM, &p.tmB, fb, ncol + 64 * a, k0, bz1, bz2, 0);
            }
          } else if (p.dbg & 2) {
            ptx::mbar_arrive(fb);
          } else {
            ptx::mbar_arrive_expect_tx(fb, (uint32_t)C::STAGE);
            const uint32_t sA = sbase + stage * C::STAGE;
            const uint32_t sB = sA + 2 * C::A_PLANE;
            const int k0 = ti.k_begin + kc * BK;
            if constexpr (C::CPLX) {
              // {Ar, Ai} x {hi, lo} then {Br, Bi} x {hi, lo}; the part index rides on the z2 axis of the A map and
              // is a K offset of op.K on the B map
              const uint32_t sBc = sA + 4 * C::A_PLANE;
#pragma unroll
              for (int part = 0; part < 2; ++part) {
                if constexpr (C::CT) {
                  // grouped form: n tile = group (BN == group_n), which contracts its own k range of each part of A
                  const int ka = k0 + part * (op.group_n ? op.a_part_k : op.K) + (op.group_n ? ti.tn * op.group_n : 0);
                  ptx::tma_load_5d(sA + 2 * part * C::A_PLANE, &p.tmA, fb, ka, ti.m0, az1, az2, 0);
                  ptx::tma_load_5d(sBc + 2 * part * C::B_PLANE, &p.tmB, fb, k0, ti.n_begin, bz1, part, 0);
                } else {
                  ptx::tma_load_5d(sA + 2 * part * C::A_PLANE, &p.tmA, fb, k0, ti.m0, az1, part, 0);
                  ptx::tma_load_5d(sBc + 2 * part * C::B_PLANE, &p.tmB, fb, k0 + part * op.K, ti.n_begin, bz1, bz2, 0);
                }
              }
            } else {
              if (!C::A_MN) {
                ptx::tma_load_5d(sA, &p.tmA, fb, k0, ti.m0, az1, az2, 0);
              } else {
                // 64-wide MN atoms, each [plane][BK rows][128 B]
#pragma unroll
                for (int a = 0; a < 2; ++a) ptx::tma_load_5d(sA + a * C::MN_ATOM, &p.tmA, fb, ti.m0 + 64 * a, k0, az1, az2, 0);
              }
              if (!C::B_MN) {
                ptx::tma_load_5d(sB, &p.tmB, fb, k0, ti.n_begin, bz1, bz2, 0);
              } else {
#pragma unroll
                for (int a = 0; a < BN / 64; ++a) ptx::tma_load_5d(sB + a * C::MN_ATOM, &p.tmB, fb, ti.n_begin + 64 * a, k0, bz1, bz2, 0);
              }
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (p.trace && lane == 0) {
        trace_put(p, it, 5, tr_wait);
        trace_put(p, it, 6, clock64());
      }
      ++it;
    }
  } else if (warp == 1 && crank == 0) {
    // ===================== MMA issuer (warp-uniform control flow keeps the descriptors in uniform registers) ==========
    // K-major (64B swizzle for BK = 32, 128B for BK = 64): rows of 2*BK bytes, 8-row groups SBO apart; a K-step is
    // 32 bytes inside the swizzled row.  MN-major (128B swizzle): [k][64 mn] atoms, 8-k groups SBO = 1024 B apart, the next
    // 64-mn atom LBO = BK*128 B away; a K-step is 16 k-rows = 2048 B.  Descriptor address fields are in 16-byte units.
    const uint64_t descA0 = C::A_MN ? ptx::smem_desc(0, C::MN_ATOM, 1024, 2u) : ptx::smem_desc(0, 0, 8 * BK * 2, C::K_LAYOUT);
    const uint64_t descB0 = C::B_MN ? ptx::smem_desc(0, C::MN_ATOM, 1024, 2u) : ptx::smem_desc(0, 0, 8 * BK * 2, C::K_LAYOUT);
    constexpr uint32_t kstepA = (C::A_MN ? 2048 : 32) >> 4, kstepB = (C::B_MN ? 2048 : 32) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t it = 0;
    Tile ti;
    const int dbg = p.dbg, nterms = p.nterms, op_K = op.K, op_ksplit = op.k_split;
    const bool tracing = p.trace != nullptr, batch2 = p.mma_batch != 0;
    for (long long t = t_first; t < total; t += t_step) {
      if (!decode_tile<BN, BK, C::TILE_M>(p, t, ti)) continue;
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      if (p.trace && lane == 0) trace_put(p, it, 0, clock64());
      ptx::mbar_wait(tempty_bar(as), aph ^ 1u);
      ptx::tc_fence_after();
      long long tr_wait = 0;
      if (p.trace && lane == 0) {
        trace_put(p, it, 1, clock64());
        trace_put(p, it, 13, ((long long)ti.num_kc << 32) | (uint32_t)ti.n_count);
      }
      const int n_eff = (ti.n_count + 15) & ~15;
      const uint32_t idesc = ptx::instr_desc_bf16(C::TILE_M, n_eff, C::A_MN ? 1 : 0, C::B_MN ? 1 : 0);
      const uint32_t tmem_d = tmem_base + as * (C::NACC * BN);
      // all MMAs of the K-chunk kc (operands in ring slot st) + the commit that frees the slot; called by the elected lane
      auto issue_stage = [&](int st, int kc) {
        const uint32_t sA = sbase + st * C::STAGE;
        const uint64_t a_hi = descA0 + (uint64_t)(sA >> 4), a_lo = a_hi + (uint64_t)(C::A_LO >> 4);
        const uint64_t b_hi = a_hi - descA0 + descB0 + (uint64_t)((2 * C::A_PLANE) >> 4), b_lo = b_hi + (uint64_t)(C::B_LO >> 4);
        if constexpr (C::CPLX) {
          // planes in the stage: Ar_hi Ar_lo Ai_hi Ai_lo | Br_hi Br_lo Bi_hi Bi_lo
          constexpr uint64_t AP = C::A_PLANE >> 4, BP = C::B_PLANE >> 4;
          const uint64_t b0 = a_hi - descA0 + descB0 + 4 * AP;
          const uint32_t tmem_r = tmem_base + as * (2 * BN), tmem_i = tmem_r + BN;
          const uint32_t ineg = idesc | (1u << 13);  // negate A: -Ai * Bi
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t ar_h = a_hi + kk * kstepA, ar_l = ar_h + AP, ai_h = ar_h + 2 * AP, ai_l = ar_h + 3 * AP;
            const uint64_t br_h = b0 + kk * kstepB, br_l = br_h + BP, bi_h = br_h + 2 * BP, bi_l = br_h + 3 * BP;
            const uint32_t first = (kc | kk) != 0 ? 1u : 0u;
            ptx::umma_bf16(tmem_r, ar_h, br_h, idesc, first);
            ptx::umma_bf16(tmem_r, ar_h, br_l, idesc, 1u);
            ptx::umma_bf16(tmem_r, ar_l, br_h, idesc, 1u);
            ptx::umma_bf16(tmem_r, ai_h, bi_h, ineg, 1u);
            ptx::umma_bf16(tmem_r, ai_h, bi_l, ineg, 1u);
            ptx::umma_bf16(tmem_r, ai_l, bi_h, ineg, 1u);
            ptx::umma_bf16(tmem_i, ai_h, br_h, idesc, first);
            ptx::umma_bf16(tmem_i, ai_h, br_l, idesc, 1u);
            ptx::umma_bf16(tmem_i, ai_l, br_h, idesc, 1u);
            ptx::umma_bf16(tmem_i, ar_h, bi_h, idesc, 1u);
            ptx::umma_bf16(tmem_i, ar_h, bi_l, idesc, 1u);
            ptx::umma_bf16(tmem_i, ar_l, bi_h, idesc, 1u);
          }
        } else if constexpr (PAIR) {
          // K-steps that lie entirely beyond K (zero-filled by TMA) are not issued (forward DFT: K = 360 -> 23 of 24)
          const int k0 = ti.k_begin + kc * BK;
          const int kk_n = min(BK / 16, (op_K - k0 + 15) >> 4);
          // butterfly mode: the chunks from k_split on go to the second accumulator (k_split is a multiple of BK)
          const bool second = C::BFLY && k0 >= op_ksplit;
          const uint32_t tmem_t = tmem_d + (second ? BN : 0);
          const bool fresh = kc == 0 || (C::BFLY && k0 == op_ksplit);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            if (kk < kk_n) {
              ptx::umma_bf16_2sm(tmem_t, a_hi + kk * kstepA, b_hi + kk * kstepB, idesc, (fresh && kk == 0) ? 0u : 1u);
              ptx::umma_bf16_2sm(tmem_t, a_hi + kk * kstepA, b_lo + kk * kstepB, idesc, 1u);
              ptx::umma_bf16_2sm(tmem_t, a_lo + kk * kstepA, b_hi + kk * kstepB, idesc, 1u);
            }
          }
        } else if (!(dbg & 4)) {
          // butterfly mode: the chunks from k_split on go to the second accumulator (k_split is a multiple of BK)
          const int k0 = ti.k_begin + kc * BK;
          const bool second = (C::BFLY || C::CLN) && k0 >= op_ksplit;
          const uint32_t tmem_t = tmem_d + (second ? BN : 0);
          const bool fresh = kc == 0 || ((C::BFLY || C::CLN) && k0 == op_ksplit);
          const int kk_n = min(BK / 16, (op_K - k0 + 15) >> 4);  // K-steps entirely beyond K are not issued
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            if (kk < kk_n) {
              ptx::umma_bf16(tmem_t, a_hi + kk * kstepA, b_hi + kk * kstepB, idesc, (fresh && kk == 0) ? 0u : 1u);
              if (nterms == 3) {
                ptx::umma_bf16(tmem_t, a_hi + kk * kstepA, b_lo + kk * kstepB, idesc, 1u);
                ptx::umma_bf16(tmem_t, a_lo + kk * kstepA, b_hi + kk * kstepB, idesc, 1u);
              }
            }
          }
        }
        if constexpr (PAIR) ptx::umma_commit_2sm(empty_bar(st));  // frees the slot in both CTAs
        else ptx::umma_commit(empty_bar(st));                     // frees the smem slot when these MMAs retire
      };
      for (int kc = 0; kc < ti.num_kc;) {
        if (tracing) {
          const long long c0 = clock64();
          ptx::mbar_wait(full_bar(stage), phase);
          tr_wait += clock64() - c0;
        } else {
          ptx::mbar_wait(full_bar(stage), phase);
        }
        // The barrier round trip (wait, fence, election, commit, re-convergence) is ~250 clocks of serial latency per ring slot,
        // which the tensor pipe's short queue does not hide when a slot holds only 6 MMAs of <= 96 clocks: when the NEXT slot
        // has landed too (non-blocking test), its MMAs are issued in the same round.
        const int stage2 = (stage + 1 == STAGES) ? 0 : stage + 1;
        const uint32_t phase2 = (stage + 1 == STAGES) ? (phase ^ 1u) : phase;
        bool two = false;
        if (batch2 && kc + 1 < ti.num_kc) two = __all_sync(0xffffffffu, ptx::mbar_test_wait(full_bar(stage2), phase2));
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_stage(stage, kc);
          if (two) issue_stage(stage2, kc + 1);
        }
        __syncwarp();
        const int adv = two ? 2 : 1;
        kc += adv;
        stage += adv;
        if (stage >= STAGES) { stage -= STAGES; phase ^= 1u; }
      }
      if (ptx::elect_one()) {  // accumulator ready for the epilogue (of both CTAs)
        if constexpr (PAIR) ptx::umma_commit_2sm(tfull_bar(as));
        else ptx::umma_commit(tfull_bar(as));
      }
      __syncwarp();
      if (p.trace && lane == 0) {
        trace_put(p, it, 2, tr_wait);
        trace_put(p, it, 3, clock64());
      }
      ++it;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;           // TMEM lane quarter this warp may access (warp id % 4)
    const int sub = (warp - 4) >> 2;  // which of the kEpiWarps/4 warps of the quarter
    uint8_t* stg = C::SP ? stg_all : stg_all + (warp - 4) * kStgBytesPerWarp;
    uint32_t it = 0;
    uint32_t xphase = 0;  // SP: parity of the group's operand-tile barrier
    Tile ti;
    for (long long t = t_first; t < total; t += t_step) {
      if (!decode_tile<BN, BK, C::TILE_M>(p, t, ti)) continue;
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      ti.m0 += 128 * (int)crank;  // pair mode: this CTA owns the second 128 rows of the 256-row tile
      if constexpr (C::NC) epilogue_nc_prefetch<C>(p, ti, q, sub, lane);
      const bool tr = p.trace && lane == 0 && (warp == 4 || warp == 15);
      const int trs = warp == 4 ? 7 : 10;
      if (tr) trace_put(p, it, trs, clock64());
      const uint32_t tacc = tmem_base + as * (C::NACC * BN) + ((uint32_t)(32 * q) << 16);
      if constexpr (C::SP) {
        if (tr) trace_put(p, it, trs + 1, clock64());  // (the wait is inside: work = wait + chunks for this variant)
        epilogue_sp<C>(p, ti, tacc, q, sub, lane, stg_all, reinterpret_cast<float*>(stg_all + C::SP_STATS),
                       reinterpret_cast<float*>(stg_all + C::SP_VEC), x_bar(sub), xphase, tfull_bar(as), aph);
      } else {
        ptx::mbar_wait(tfull_bar(as), aph);
        ptx::tc_fence_after();
        if (tr) trace_put(p, it, trs + 1, clock64());
        if (!(p.dbg & 1)) {
          if constexpr (C::CLN) epilogue_cln<C>(p, ti, tacc, q, sub, lane, stg, reinterpret_cast<float2*>(stg_all + C::CLN_COLS + (warp - 4) * 256));
          else if constexpr (C::NC) epilogue_nc<C>(p, ti, tacc, q, sub, lane, stg);
          else epilogue_rowc<C>(p, ti, tacc, q, sub, lane, stg);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (tr) trace_put(p, it, trs + 2, clock64());
      if (lane == 0) {
        if constexpr (PAIR) ptx::mbar_arrive_leader(tempty_bar(as));
        else ptx::mbar_arrive(tempty_bar(as));
      }
      if constexpr (C::SP) {
        if ((op.epi.flags & EPI_ROW_STATS) && !(p.dbg & 1)) epilogue_sp_flush<C>(p, ti, reinterpret_cast<float*>(stg_all + C::SP_STATS));
      }
      ++it;
    }
    if constexpr (C::SP) ptx::bulk_wait0();  // this thread's TMA stores have completed (only the issuing lanes have any)
  }

  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync();  // the peer may still read this CTA's B half / signal its barriers
  else __syncthreads();
  if (p.trace && threadIdx.x == 0) {
    trace_put(p, kTraceIters - 1, 4, globaltimer_ns());
    trace_put(p, kTraceIters - 1, 5, clock64());
  }
  if (warp == 2) {
    ptx::tc_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  if (!fn) throw Error(ACE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    ACE_CHECK_CUDA(cudaGetDevice(&dev));
    ACE_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

// 5-D map: (inner, outer, z1, z2, plane).  Batch axes the operand does not vary along get extent 1.
void make_tmap(CUtensorMap* tm, const Operand& o, bool mn_major, long long rows, long long kext, int Z1, int Z2,
               int box_inner, int box_outer, CUtensorMapSwizzle swz, int* z1_on, int* z2_on, const char* name) {
  cuuint64_t dims[5], strides[4];
  // one box = the tile of BOTH planes (hi, lo): half the TMA instructions per stage (the producer warp is issue-bound otherwise)
  cuuint32_t box[5] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1, 1, 2}, estr[5] = {1, 1, 1, 1, 1};
  if (!mn_major) {
    dims[0] = (cuuint64_t)kext;
    dims[1] = (cuuint64_t)rows;
    strides[0] = (cuuint64_t)o.s_row * 2;
  } else {
    dims[0] = (cuuint64_t)rows;
    dims[1] = (cuuint64_t)kext;
    strides[0] = (cuuint64_t)o.s_k * 2;
  }
  *z1_on = (Z1 > 1 && o.s_z1 != 0) ? 1 : 0;
  *z2_on = (Z2 > 1 && o.s_z2 != 0) ? 1 : 0;
  dims[2] = *z1_on ? (cuuint64_t)Z1 : 1;
  strides[1] = *z1_on ? (cuuint64_t)o.s_z1 * 2 : strides[0] * dims[1];
  dims[3] = *z2_on ? (cuuint64_t)Z2 : 1;
  strides[2] = *z2_on ? (cuuint64_t)o.s_z2 * 2 : strides[1] * dims[2];
  dims[4] = 2;
  strides[3] = (cuuint64_t)o.plane * 2;
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)o.ptr, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(ACE_ERR_CUDA,
                strprintf("gemm %s: cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu,%llu) box=(%u,%u)",
                          name, (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                          (unsigned long long)dims[3], (unsigned long long)dims[4], (unsigned long long)strides[0],
                          (unsigned long long)strides[1], (unsigned long long)strides[2], (unsigned long long)strides[3],
                          box[0], box[1]));
}

inline bool aligned8(long long v) { return (v & 7) == 0; }
inline bool aligned4(long long v) { return (v & 3) == 0; }

thread_local int t_scalar_store = 0;  // set by dispatch() for the launch it is about to make

// ---- tile lists of the triangular ops (see UmmaParams::tiles) ----
struct TileKey {
  int M, N, K, Z1, Z2, flags, tile_m, bn, bk, m_fastest, workers;
  bool operator<(const TileKey& o) const { return memcmp(this, &o, sizeof(TileKey)) < 0; }
};
struct TileList {
  DevBuf buf;
  long long n = 0;
};

// Dense ops with several N tiles per M tile (forward DFT: N = 2 mmax = 362 -> a 256-wide and a 112-wide tile): the implicit walk
// gives worker w the tiles w, w + G, ... and with an even G every worker only ever sees ONE of the two N tiles, i.e. half the
// workers do 2.3x the MMA work of the others.  Listed order instead: worker w takes every N tile of M tile w, then of M tile
// w + G, ... (the second read of the A tile hits L2 right after the first); slots beyond the last M tile are empty entries.
const TileList& grouped_tile_list(const GemmOp& op, int tile_m, int bn, int workers) {
  static std::mutex mu;
  static std::map<TileKey, TileList> cache;
  TileKey key;
  memset(&key, 0, sizeof(key));
  key.M = op.M; key.N = op.N; key.K = op.K; key.Z1 = op.Z1; key.Z2 = op.Z2;
  key.flags = 16; key.tile_m = tile_m; key.bn = bn; key.workers = workers;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const int tiles_m = (op.M + tile_m - 1) / tile_m, tiles_n = (op.N + bn - 1) / bn;
  const long long groups = (long long)tiles_m * op.Z1 * op.Z2;
  const long long rounds = (groups + workers - 1) / workers;
  std::vector<int4> host;
  host.reserve((size_t)(rounds * tiles_n * workers));
  for (long long r = 0; r < rounds; ++r)
    for (int tn = 0; tn < tiles_n; ++tn)
      for (int w = 0; w < workers; ++w) {
        const long long g = r * workers + w;
        if (g >= groups) {
          host.push_back(make_int4(0, tiles_n, 0, 0));  // empty slot: decode_tile() rejects it (no columns)
          continue;
        }
        const int tm = (int)(g % tiles_m);
        const long long z = g / tiles_m;
        host.push_back(make_int4(tm, tn, (int)(z / op.Z2), (int)(z % op.Z2)));
      }
  while (!host.empty() && host.back().y == tiles_n) host.pop_back();
  TileList& tl = cache[key];
  tl.n = (long long)host.size();
  tl.buf.ensure(std::max<size_t>(16, host.size() * sizeof(int4)));
  if (!host.empty()) ACE_CHECK_CUDA(cudaMemcpy(tl.buf.p, host.data(), host.size() * sizeof(int4), cudaMemcpyHostToDevice));
  return tl;
}

const TileList& tile_list(const GemmOp& op, int tile_m, int bn, int bk, int m_fastest, int workers) {
  static std::mutex mu;
  static std::map<TileKey, TileList> cache;
  TileKey key;
  memset(&key, 0, sizeof(key));
  key.M = op.M; key.N = op.N; key.K = op.K; key.Z1 = op.Z1; key.Z2 = op.Z2;
  key.flags = (op.n_lo_z1 ? 1 : 0) | (op.n_hi_z1 ? 2 : 0) | (op.k_lo_z1 ? 4 : 0) | (op.m_hi_z1 ? 8 : 0);
  key.tile_m = tile_m; key.bn = bn; key.bk = bk; key.m_fastest = m_fastest;
  key.workers = options().tile_serpentine ? workers : 0;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const int tiles_m = (op.M + tile_m - 1) / tile_m, tiles_n = (op.N + bn - 1) / bn;
  struct Ent { int4 t; long long cost; long long order; };
  std::vector<Ent> ents;
  long long order = 0;
  // same enumeration order as the implicit walk (locality: neighbouring entries share an operand tile), then a stable
  // sort by descending cost: CTA b takes entries b, b + G, b + 2G, ... = one tile of every cost stratum
  for (int z1 = 0; z1 < op.Z1; ++z1)
    for (int z2 = 0; z2 < op.Z2; ++z2)
      for (int o = 0; o < tiles_m * tiles_n; ++o) {
        const int tm = m_fastest ? o % tiles_m : o / tiles_n, tn = m_fastest ? o / tiles_m : o % tiles_n;
        const int n_lo = op.n_lo_z1 ? z1 : 0, n_hi = op.n_hi_z1 ? std::min(op.N, z1 + 1) : op.N, k_lo = op.k_lo_z1 ? z1 : 0;
        const int n_begin = std::max(tn * bn, (n_lo / 16) * 16), n_end = std::min(tn * bn + bn, n_hi);
        const int k_begin = (k_lo / bk) * bk, num_kc = (op.K - k_begin + bk - 1) / bk;
        if (op.m_hi_z1 && tm * tile_m > z1) continue;
        if (!(n_end - n_begin > 0 && n_end > n_lo && num_kc > 0)) continue;
        const long long n_eff = std::max(64, (n_end - n_begin + 15) & ~15);  // MMAs narrower than ~64 columns cost the same
        ents.push_back({make_int4(tm, tn, z1, z2), n_eff * num_kc, order++});
      }
  std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.cost > b.cost; });
  // worker w takes entries w, w + G, w + 2G, ...: with every stratum of G entries in descending order worker 0 would get the
  // heaviest tile of each stratum and worker G - 1 the lightest; odd strata are reversed (serpentine) so the sums even out
  if (workers > 0 && options().tile_serpentine) {
    for (size_t b = (size_t)workers; b < ents.size(); b += 2 * (size_t)workers)
      std::reverse(ents.begin() + b, ents.begin() + std::min(ents.size(), b + (size_t)workers));
  }
  TileList& tl = cache[key];
  tl.n = (long long)ents.size();
  std::vector<int4> host(ents.size());
  for (size_t i = 0; i < ents.size(); ++i) host[i] = ents[i].t;
  tl.buf.ensure(std::max<size_t>(16, host.size() * sizeof(int4)));
  if (!host.empty()) ACE_CHECK_CUDA(cudaMemcpy(tl.buf.p, host.data(), host.size() * sizeof(int4), cudaMemcpyHostToDevice));
  return tl;
}

// SP variants run the caller's op with the operand roles exchanged (the epilogue parameters keep the caller's meaning)
inline GemmOp exchanged(const GemmOp& o) {
  GemmOp x = o;
  x.M = o.N;
  x.N = o.M;
  x.A = o.B;
  x.B = o.A;
  return x;
}

template <class C>
void launch(const GemmOp& op_in, cudaStream_t stream) {
  const GemmOp op = C::SP ? exchanged(op_in) : op_in;
  static bool attr_set = false;
  if (!attr_set) {
    ACE_CHECK_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  UmmaParams p;
  p.op = op;
  p.tiles_m = (op.M + C::TILE_M - 1) / C::TILE_M;
  p.tiles_n = (op.N + C::BN - 1) / C::BN;
  p.nterms = options().split_terms;
  p.dbg = options().dbg;
  p.mma_batch = options().mma_batch;
  p.scalar_store = t_scalar_store;
  // the operand that is re-read by neighbouring tiles should be the small one: keep the big streaming operand's
  // tile shared by consecutive CTAs (they run concurrently, so the second reader hits L2)
  p.m_fastest = (C::B_MN || C::CT) ? 1 : 0;
  const CUtensorMapSwizzle kswz = (C::BK == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (C::CT) {
    // [Ar | Ai] along k: one map over 2K columns; {Br, Bi}: the part index rides on the z2 axis of the B map
    make_tmap(&p.tmA, op.A, false, op.M, op.group_n ? 2LL * op.a_part_k : 2LL * op.K, op.Z1, op.Z2, C::BK, 128, kswz, &p.a_z1_on, &p.a_z2_on,
              op.name);
    Operand b = op.B;
    b.s_z2 = op.b_part;
    make_tmap(&p.tmB, b, false, op.N, op.K, op.Z1, 2, C::BK, C::BN, kswz, &p.b_z1_on, &p.b_z2_on, op.name);
  } else if (C::CPLX) {
    Operand a = op.A;
    a.s_z2 = op.a_part;  // the {real, imaginary} matrix index rides on the z2 axis of the map
    make_tmap(&p.tmA, a, false, op.M, op.K, op.Z1, 2, C::BK, 128, kswz, &p.a_z1_on, &p.a_z2_on, op.name);
    Operand b = op.B;
    make_tmap(&p.tmB, b, false, op.N, 2LL * op.K, op.Z1, op.Z2, C::BK, C::BN, kswz, &p.b_z1_on, &p.b_z2_on, op.name);
  } else if (!C::A_MN)
    make_tmap(&p.tmA, op.A, false, op.M, op.K, op.Z1, op.Z2, C::BK, 128, kswz, &p.a_z1_on, &p.a_z2_on, op.name);
  else
    make_tmap(&p.tmA, op.A, true, op.M, op.K, op.Z1, op.Z2, 64, C::BK, CU_TENSOR_MAP_SWIZZLE_128B, &p.a_z1_on, &p.a_z2_on, op.name);
  if (C::CPLX) {
  } else if (!C::B_MN)
    make_tmap(&p.tmB, op.B, false, op.N, op.K, op.Z1, op.Z2, C::BK, C::BNL, kswz, &p.b_z1_on, &p.b_z2_on, op.name);
  else
    make_tmap(&p.tmB, op.B, true, op.N, op.K, op.Z1, op.Z2, 64, C::BK, CU_TENSOR_MAP_SWIZZLE_128B, &p.b_z1_on, &p.b_z2_on, op.name);
  long long total = (long long)p.tiles_m * p.tiles_n * op.Z1 * op.Z2;
  p.tiles = nullptr;
  p.n_listed = 0;
  if ((op.n_lo_z1 || op.n_hi_z1 || op.k_lo_z1 || op.m_hi_z1) && options().tile_list) {
    const TileList& tl = tile_list(op, C::TILE_M, C::BN, C::BK, p.m_fastest, C::PAIR ? sm_count() / 2 : sm_count());
    if (tl.n == 0) return;  // nothing to compute
    p.tiles = tl.buf.as<int4>();
    p.n_listed = total = tl.n;
  } else if (!p.m_fastest && p.tiles_n > 1 && options().tile_list && options().group_order &&
             4 * (op.N - (p.tiles_n - 1) * C::BN) < 3 * C::BN) {  // a last N tile much narrower than the others (measured: -5 us forward DFT; +6 us inverse DFT, whose two tiles are 96 and 84 wide)
    const int workers = C::PAIR ? sm_count() / 2 : sm_count();
    if ((long long)p.tiles_m * op.Z1 * op.Z2 >= workers) {
      const TileList& tl = grouped_tile_list(op, C::TILE_M, C::BN, workers);
      p.tiles = tl.buf.as<int4>();
      p.n_listed = total = tl.n;
    }
  }
  int grid = C::PAIR ? 2 * (int)std::min<long long>(total, sm_count() / 2) : (int)std::min<long long>(total, sm_count());
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreadsUmma);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (options().pdl == 2 || (options().pdl && !C::PAIR)) ? 1 : 0;  // (2: cluster launches too)
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (C::PAIR) {
    attr[cfg.numAttrs].id = cudaLaunchAttributeClusterDimension;
    attr[cfg.numAttrs].val.clusterDim.x = 2;
    attr[cfg.numAttrs].val.clusterDim.y = 1;
    attr[cfg.numAttrs].val.clusterDim.z = 1;
    ++cfg.numAttrs;
  }
  if (options().l2_persist && (op.epi.flags & EPI_OUT_PLANES) && op.epi.out_plane > 0) {
    // Experiment for the "no intermediate tensor in HBM" question: mark the split-plane OUTPUT of this
    // GEMM (an intermediate the next kernel consumes) as persisting in the L2 set-aside, everything else as streaming.
    static size_t max_win = 0, set_aside = 0;
    if (!max_win) {
      int dev = 0, v = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, dev);
      max_win = (size_t)std::max(v, 1);
      cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev);
      set_aside = (size_t)std::max(v, 0);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside);
    }
    const size_t bytes = std::min<size_t>(max_win, 2 * (size_t)op.epi.out_plane * sizeof(bf16));
    attr[cfg.numAttrs].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[cfg.numAttrs].val.accessPolicyWindow.base_ptr = (void*)op.epi.out;
    attr[cfg.numAttrs].val.accessPolicyWindow.num_bytes = bytes;
    attr[cfg.numAttrs].val.accessPolicyWindow.hitRatio = bytes ? std::min(1.0f, (float)set_aside / (float)bytes) : 0.f;
    attr[cfg.numAttrs].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[cfg.numAttrs].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++cfg.numAttrs;
  }
  p.sp_tma = 0;
  if constexpr (C::SP) {
    const EpiParams& e = op.epi;
    if ((C::EF & EPI_OUT_PLANES) && options().sp_tma && aligned8(e.o_m0) && aligned8(e.o_z2) && aligned8(e.out_plane) && !(((uintptr_t)e.out) & 15)) {
      // (position, channel, sample, plane); one store = [2 planes][32 channels][128 positions] from the staging tile
      cuuint64_t dims[4] = {(cuuint64_t)op.M, (cuuint64_t)op.N, (cuuint64_t)std::max(1, op.Z2), 2};
      cuuint64_t strides[3] = {(cuuint64_t)e.o_m0 * 2, (cuuint64_t)(op.Z2 > 1 ? e.o_z2 : e.o_m0 * (long long)op.N) * 2, (cuuint64_t)e.out_plane * 2};
      cuuint32_t box[4] = {128, 32, 1, 2}, estr[4] = {1, 1, 1, 1};
      CUresult r = encode_fn()(&p.tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)e.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      p.sp_tma = (r == CUDA_SUCCESS) ? 1 : 0;
    }
    if (!p.sp_tma) memcpy(&p.tmO, &p.tmA, sizeof(CUtensorMap));  // a valid descriptor for the prefetch
    p.sp_tmx = 0;
    if (p.sp_tma && C::SP_OPND && options().sp_tmx) {
      CUresult r = CUDA_ERROR_INVALID_VALUE;
      cuuint32_t estr[4] = {1, 1, 1, 1};
      if ((C::EF & EPI_RES_PLANES) && aligned8(e.res_m0) && aligned8(e.res_z2) && aligned8(e.res_plane) && !(((uintptr_t)e.res) & 15)) {
        cuuint64_t dims[4] = {(cuuint64_t)op.M, (cuuint64_t)op.N, (cuuint64_t)std::max(1, op.Z2), 2};
        cuuint64_t strides[3] = {(cuuint64_t)e.res_m0 * 2, (cuuint64_t)(op.Z2 > 1 ? e.res_z2 : e.res_m0 * (long long)op.N) * 2, (cuuint64_t)e.res_plane * 2};
        cuuint32_t box[4] = {128, 32, 1, 2};
        r = encode_fn()(&p.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)e.res, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else if ((C::EF & EPI_ADD_F32) && aligned4(e.add_m0) && aligned4(e.add_z2) && !(((uintptr_t)e.add) & 15)) {
        const cuuint64_t per_sample = (cuuint64_t)(op.Z2 > 1 ? e.add_z2 : e.add_m0 * (long long)op.N) * 4;
        cuuint64_t dims[4] = {(cuuint64_t)op.M, (cuuint64_t)op.N, (cuuint64_t)std::max(1, op.Z2), 1};
        cuuint64_t strides[3] = {(cuuint64_t)e.add_m0 * 4, per_sample, per_sample * (cuuint64_t)std::max(1, op.Z2)};
        cuuint32_t box[4] = {128, 32, 1, 1};
        r = encode_fn()(&p.tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)e.add, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      p.sp_tmx = (r == CUDA_SUCCESS) ? 1 : 0;
    }
    if (!p.sp_tmx) memcpy(&p.tmX, &p.tmA, sizeof(CUtensorMap));
  }
  p.trace = nullptr;
  static DevBuf trace_buf;
  const size_t trace_bytes = (size_t)grid * kTraceIters * kTraceSlots * sizeof(long long);
  const char* trace_path = options().trace ? getenv("ACE_B200_TRACE_FILE") : nullptr;
  if (trace_path) {
    trace_buf.ensure((size_t)sm_count() * kTraceIters * kTraceSlots * sizeof(long long));
    ACE_CHECK_CUDA(cudaMemsetAsync(trace_buf.p, 0, trace_bytes, stream));
    p.trace = trace_buf.as<long long>();
  }
  ACE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_umma_kernel<C>, p));
  if (trace_path) {
    // development only (never under graph capture): record = {char name[64], int grid, iters, slots, stages, tile_m, bn, bk, pair} + samples
    std::vector<long long> host(trace_bytes / sizeof(long long));
    ACE_CHECK_CUDA(cudaStreamSynchronize(stream));
    ACE_CHECK_CUDA(cudaMemcpy(host.data(), trace_buf.p, trace_bytes, cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_path, "ab")) {
      char name[64] = {0};
      strncpy(name, op.name ? op.name : "?", 63);
      const int hdr[8] = {grid, kTraceIters, kTraceSlots, C::STAGES, C::TILE_M, C::BN, C::BK, C::PAIR ? 1 : 0};
      fwrite(name, 1, 64, f);
      fwrite(hdr, sizeof(int), 8, f);
      fwrite(host.data(), 1, trace_bytes, f);
      fclose(f);
    }
  }
  after_launch(op.name);
  g_umma_count.fetch_add(1, std::memory_order_relaxed);
}

// ---- which compiled variant (if any) serves this op ----
struct Variant {
  bool a_mn, b_mn, nc;
  bool scalar;  // element-wise stores (alignment of the vector paths not met)
  uint32_t ef;
};

bool classify(const GemmOp& op, Variant& v, const char** why) {
  auto fail = [&](const char* w) {
    if (why) *why = w;
    return false;
  };
  const bool a_k = op.A.s_k == 1, a_mn = op.A.s_row == 1 && !a_k;
  const bool b_k = op.B.s_k == 1, b_mn = op.B.s_row == 1 && !b_k;
  if (!a_k && !a_mn) return fail("A is neither K-major nor MN-major");
  if (!b_k && !b_mn) return fail("B is neither K-major nor MN-major");
  if (a_k && !aligned8(op.A.s_row)) return fail("A row stride not 16B aligned");
  if (a_mn && !aligned8(op.A.s_k)) return fail("A k stride not 16B aligned");
  if (b_k && !aligned8(op.B.s_row)) return fail("B row stride not 16B aligned");
  if (b_mn && !aligned8(op.B.s_k)) return fail("B k stride not 16B aligned");
  if (!aligned8(op.A.plane) || !aligned8(op.B.plane)) return fail("plane offset not 16B aligned");
  if (!aligned8(op.A.s_z1) || !aligned8(op.A.s_z2) || !aligned8(op.B.s_z1) || !aligned8(op.B.s_z2)) return fail("batch stride not 16B aligned");
  if ((((uintptr_t)op.A.ptr) & 15) || (((uintptr_t)op.B.ptr) & 15)) return fail("base pointer not 16B aligned");
  if (op.M < 1 || op.N < 1 || op.K < 1) return fail("empty");
  if (op.A.plane <= 0 || op.B.plane <= 0) return fail("plane offset must be positive");
  const EpiParams& e = op.epi;
  const uint32_t f = e.flags;
  if (op.cplx && (!a_k || !b_k || !aligned8(op.a_part) || !aligned8(op.b_part) || op.k_lo_z1 || (op.K & 7)))
    return fail("complex mode needs K-major operands, 16B-aligned parts and K % 8 == 0");
  if (op.cplx == 2) {
    // roles exchanged: NC epilogue, plain split-plane output with the imaginary part N columns further
    if ((f & ~(uint32_t)EPI_OUT_PLANES) != 0 || !(f & EPI_OUT_PLANES)) return fail("complex mode 2 takes the plain split-plane output only");
    if (op.n_lo_z1 || op.n_hi_z1) return fail("complex mode 2 takes no triangular N range");
    if (e.mdiv < op.M) return fail("complex mode 2 needs affine rows");
    if (e.o_n != 1 || !aligned4(op.N) || !aligned4(e.o_m0) || !aligned4(e.o_z1) || !aligned4(e.o_z2) || !aligned4(e.out_plane) || (((uintptr_t)e.out) & 7))
      return fail("complex mode 2: plane output not 8B-vectorisable");
    if (op.B.s_z2 != 0) return fail("complex mode 2: B must not vary along z2 (the axis carries the {re, im} index)");
    if (op.group_n && ((op.group_n != 64 && op.group_n != 128) || op.N % op.group_n || op.K > op.group_n || op.a_part_k < op.N))
      return fail("grouped complex mode: groups of 64 or 128 columns");
    v.a_mn = v.b_mn = false;
    v.nc = true;
    v.scalar = false;
    v.ef = EPI_OUT_PLANES;
    return true;
  }
  if (op.m_hi_z1) return fail("a triangular M range is only compiled for complex mode 2");
  if (op.bfly && (e.o_m0 == 1 && e.o_n != 1 && (f & EPI_OUT_PLANES))) return fail("butterfly mode needs the NC epilogue");
  v.a_mn = a_mn;
  v.b_mn = b_mn;
  v.ef = f & EF_MASK;
  const bool planes = (f & EPI_OUT_PLANES) != 0, f32 = (f & EPI_OUT_F32) != 0;
  if (planes == f32) return fail("exactly one of OUT_PLANES / OUT_F32 is supported");
  // ROWC: split-plane output with contiguous rows, nothing else
  if ((f & EPI_ROW_AFFINE) && !(v.ef == EPI_OUT_PLANES && e.o_m0 == 1 && e.o_n != 1)) return fail("ROW_AFFINE is only compiled for the ROWC epilogue");
  if ((f & EPI_RES_AFFINE) && !(f & EPI_RES_PLANES)) return fail("RES_AFFINE without RES_PLANES");
  if (v.ef == EPI_OUT_PLANES && !(f & (EPI_ROW_BIAS | EPI_ROW_STATS)) && e.o_m0 == 1 && e.o_n != 1) {
    v.nc = false;
    // transposed 8-byte stores need groups of 4 rows that are contiguous and 8-byte aligned; otherwise (odd nlat) the
    // same kernel stores element-wise
    v.scalar = !aligned4(op.M) || (e.mdiv < op.M && !aligned4(e.mdiv)) || !aligned4(e.o_m1) || !aligned4(e.o_n) || !aligned4(e.o_z1) ||
               !aligned4(e.o_z2) || !aligned4(e.out_plane) || (((uintptr_t)e.out) & 7);
    if (v.scalar && op.cplx) return fail("complex mode needs the vectorised ROWC stores");
    return true;
  }
  // NC: columns contiguous, rows affine, everything 4-element aligned
  v.nc = true;
  v.scalar = false;
  const bool remap = e.mdiv < op.M;  // rows flattened as (m1, m0): a warp's 32 rows must share m1
  if (remap && (e.mdiv % 32 != 0 || (f & (EPI_ROW_BIAS | EPI_ROW_STATS | EPI_RES_AFFINE)))) return fail("NC epilogue: row remapping needs mdiv % 32 == 0 and no per-row vectors");
  if (remap && ((planes && !aligned4(e.o_m1)) || (f32 && !aligned4(e.f_m1)) || ((f & EPI_ADD_F32) && !aligned4(e.add_m1)) ||
                ((f & EPI_RES_PLANES) && !aligned4(e.res_m1))))
    return fail("NC epilogue: m1 strides not vectorisable");
  if (e.mrows && !remap && e.mrows < op.M) return fail("mrows without row remapping");
  if (op.n_lo_z1 || op.n_hi_z1) return fail("NC epilogue does not take triangular N ranges");
  if (op.bfly) {
    if (op.cplx || !b_k || v.ef != EPI_OUT_F32 || (f & ~(uint32_t)EPI_OUT_F32) || op.k_lo_z1 || op.k_split <= 0 || op.k_split >= op.K || (op.k_split % 32) != 0)
      return fail("butterfly mode: K-major B, plain fp32 output, 0 < k_split < K, k_split % 32 == 0");
  }
  if (planes && v.ef == EPI_OUT_PLANES && e.o_n == 1 &&
      (!aligned4(op.N) || !aligned4(e.o_m0) || !aligned4(e.o_z1) || !aligned4(e.o_z2) || !aligned4(e.out_plane) || (((uintptr_t)e.out) & 7))) {
    v.scalar = true;  // plain plane output with unaligned row starts (inverse Legendre at odd nlat): element-wise stores
    return true;
  }
  if (!aligned4(op.N)) return fail("NC epilogue needs N % 4 == 0");
  if (planes) {
    if (e.o_n != 1 || !aligned4(e.o_m0) || !aligned4(e.o_z1) || !aligned4(e.o_z2) || !aligned4(e.out_plane) || (((uintptr_t)e.out) & 7))
      return fail("plane output not 8B-vectorisable");
  } else {
    if (e.f_n != 1 || !aligned4(e.f_m0) || !aligned4(e.f_z1) || !aligned4(e.f_z2) || (((uintptr_t)e.outf) & 15))
      return fail("fp32 output not 16B-vectorisable");
  }
  if (f & EPI_ADD_F32) {
    if (e.add_n != 1 || !aligned4(e.add_m0) || !aligned4(e.add_z2) || (((uintptr_t)e.add) & 15))
      return fail("fp32 addend not 16B-vectorisable");
  }
  if (f & EPI_RES_PLANES) {
    if (e.res_n != 1 || !aligned4(e.res_m0) || !aligned4(e.res_z2) || !aligned4(e.res_plane) || (((uintptr_t)e.res) & 7))
      return fail("plane residual not 8B-vectorisable");
  }
  return true;
}

constexpr uint32_t P = EPI_OUT_PLANES, F = EPI_OUT_F32, G = EPI_GELU, AD = EPI_ADD_F32, RS = EPI_RES_PLANES;

bool launch_variant_pair(const GemmOp& op, const Variant& v, cudaStream_t s) {
  if (!v.nc) {
    if (v.a_mn || v.b_mn) return false;
    if (options().umma_bk == 64 || (options().umma_bk == 0 && op.bk_hint == 64)) launch<Cfg<256, false, false, P, false, 64, false, true>>(op, s);
    else launch<Cfg<256, false, false, P, false, 32, false, true>>(op, s);
    return true;
  }
  if (v.b_mn) return false;
  if (v.a_mn) {
    if (v.ef == P) { launch<Cfg<256, true, false, P, true, 32, false, true>>(op, s); return true; }
    if (v.ef == F) { launch<Cfg<256, true, false, F, true, 32, false, true>>(op, s); return true; }
    return false;
  }
  if (v.ef == P) { launch<Cfg<256, false, false, P, true, 32, false, true>>(op, s); return true; }
  if (v.ef == F) { launch<Cfg<256, false, false, F, true, 32, false, true>>(op, s); return true; }
  return false;
}
template <int BN>
bool launch_variant(const GemmOp& op, const Variant& v, cudaStream_t s) {
  if (!v.nc) {
    if (v.a_mn || v.b_mn) return false;
    if (options().umma_bk == 64 || (options().umma_bk == 0 && op.bk_hint == 64)) launch<Cfg<BN, false, false, P, false, 64>>(op, s);
    else launch<Cfg<BN, false, false, P, false, 32>>(op, s);
    return true;
  }
  if (v.a_mn && !v.b_mn) {  // inverse SHT stages
    if (v.ef == P) { launch<Cfg<BN, true, false, P, true>>(op, s); return true; }
    if (v.ef == F) { launch<Cfg<BN, true, false, F, true>>(op, s); return true; }
    return false;
  }
  if (!v.a_mn && !v.b_mn) {  // generic K-major x K-major with contiguous columns (tests, forward SHT variants)
    if (v.ef == P) { launch<Cfg<BN, false, false, P, true>>(op, s); return true; }
    if (v.ef == F) { launch<Cfg<BN, false, false, F, true>>(op, s); return true; }
    return false;
  }
  return false;
}

// 1x1 convolutions: weights (K-major) x activations (MN-major); BN = 256 (default) or 192 (option "conv_bn")
bool launch_conv_pair(const GemmOp& op, const Variant& v, cudaStream_t s) {
  switch (v.ef) {
    case G | P: launch<Cfg<256, false, true, G | P, true, 32, false, true>>(op, s); return true;
    case AD | P: launch<Cfg<256, false, true, AD | P, true, 32, false, true>>(op, s); return true;
    case AD | G | P: launch<Cfg<256, false, true, AD | G | P, true, 32, false, true>>(op, s); return true;
    case RS | P: launch<Cfg<256, false, true, RS | P, true, 32, false, true>>(op, s); return true;
    case F: launch<Cfg<256, false, true, F, true, 32, false, true>>(op, s); return true;
    case P: launch<Cfg<256, false, true, P, true, 32, false, true>>(op, s); return true;
    default: return false;
  }
}
template <int BN>
bool launch_conv_bn(const GemmOp& op, const Variant& v, cudaStream_t s) {
  switch (v.ef) {
    case G | P: launch<Cfg<BN, false, true, G | P, true>>(op, s); return true;
    case AD | F: launch<Cfg<BN, false, true, AD | F, true>>(op, s); return true;
    case AD | P: launch<Cfg<BN, false, true, AD | P, true>>(op, s); return true;
    case AD | G | P: launch<Cfg<BN, false, true, AD | G | P, true>>(op, s); return true;
    case AD | G | F: launch<Cfg<BN, false, true, AD | G | F, true>>(op, s); return true;
    case RS | F: launch<Cfg<BN, false, true, RS | F, true>>(op, s); return true;
    case RS | P: launch<Cfg<BN, false, true, RS | P, true>>(op, s); return true;
    case F: launch<Cfg<BN, false, true, F, true>>(op, s); return true;
    case P: launch<Cfg<BN, false, true, P, true>>(op, s); return true;
    case G | F: launch<Cfg<BN, false, true, G | F, true>>(op, s); return true;
    default: return false;
  }
}
// SP variants (Cfg::SP): channels on the accumulator columns, N tile 256 when the channel count is a multiple of it, else 192
template <int BN>
bool launch_conv_sp_bn(const GemmOp& op, const Variant& v, cudaStream_t s) {
  switch (v.ef) {
    case G | P: launch<Cfg<BN, true, false, G | P, false, 32, false, true, false, true>>(op, s); return true;
    case AD | P: launch<Cfg<BN, true, false, AD | P, false, 32, false, true, false, true>>(op, s); return true;
    case AD | G | P: launch<Cfg<BN, true, false, AD | G | P, false, 32, false, true, false, true>>(op, s); return true;
    case RS | P: launch<Cfg<BN, true, false, RS | P, false, 32, false, true, false, true>>(op, s); return true;
    case P: launch<Cfg<BN, true, false, P, false, 32, false, true, false, true>>(op, s); return true;
    case F: launch<Cfg<BN, true, false, F, false, 32, false, true, false, true>>(op, s); return true;
    default: return false;
  }
}
bool sp_eligible(const GemmOp& op, const Variant& v) {
  const EpiParams& e = op.epi;
  if (!v.nc || v.a_mn || !v.b_mn || op.M < 128 || (op.M & 31) || e.mdiv < op.M || op.Z1 != 1) return false;  // whole 32-channel chunks
  // the per-channel vectors are read 16 bytes at a time from a 32-channel boundary
  auto vec_ok = [](const float* p, long long z2) { return p && (((uintptr_t)p) & 15) == 0 && (z2 & 3) == 0; };
  if ((e.flags & EPI_ROW_BIAS) && !vec_ok(e.row_bias, e.rb_z2)) return false;
  if ((e.flags & EPI_RES_AFFINE) && !(vec_ok(e.res_a, e.rsa_z2) && vec_ok(e.res_s, e.rsa_z2))) return false;
  if (e.flags & EPI_ROW_AFFINE) return false;
  if ((op.epi.flags & EPI_RES_PLANES) && ((e.res_m0 | e.res_plane | e.res_z2) & 1 || (((uintptr_t)e.res) & 3))) return false;  // residual read as 32-bit words
  // Where it pays (measured at ACE2 size, profiles/r02_sp_probe.jsonl): the exchanged main loop runs at the MMA-only time
  // (no shared-memory contention), but its epilogue spends more instructions per element (2-byte stores, shuffle-reduced
  // statistics), so it wins when the main loop is long (K >= 512: fc2) or the epilogue light (no addend / statistics), and
  // never against the 256-row CTA-pair variants that channel counts of a multiple of 256 already use.
  if (options().sp != 2) {
    if (op.M % 256 == 0 || op.K < 256) return false;
    // (an addend tile loaded a chunk at a time leaves its latency exposed on the short K = 384 tiles: inner_skip 91 vs 76.5 us)
    if (e.flags & EPI_ADD_F32) return false;
    if (e.flags & (EPI_ROW_STATS | EPI_RES_PLANES)) {
      // only with the TMA-stored output / TMA-loaded residual tile (16-byte aligned strides); element-wise accesses lose
      if (!(options().sp_tma && options().sp_tmx)) return false;
      if (!(aligned8(e.o_m0) && aligned8(e.o_z2) && aligned8(e.out_plane) && !(((uintptr_t)e.out) & 15))) return false;
      if ((e.flags & EPI_RES_PLANES) && !(aligned8(e.res_m0) && aligned8(e.res_z2) && aligned8(e.res_plane) && !(((uintptr_t)e.res) & 15))) return false;
    }
  }
  return true;
}
bool launch_conv_sp(const GemmOp& op, const Variant& v, cudaStream_t s) {
  return (op.M % 256 == 0) ? launch_conv_sp_bn<256>(op, v, s) : launch_conv_sp_bn<192>(op, v, s);
}

bool launch_conv(const GemmOp& op, const Variant& v, cudaStream_t s) {
  if (!v.nc || v.a_mn || !v.b_mn) return false;
  if (options().sp && sp_eligible(op, v) && launch_conv_sp(op, v, s)) return true;
  // CTA pairs pay off when no half tile is wasted (measured: fc1 M=768 111 -> 95 us; M=384 ops gain nothing)
  const bool want_pair = options().pair == 1 ? op.M > 128 : (options().pair < 0 && op.M % 256 == 0);
  if (want_pair && launch_conv_pair(op, v, s)) return true;
  int bn = options().conv_bn;
  if (bn != 192 && bn != 256) {
    // whole waves of tiles over the SMs; the narrower tile re-reads the weights more often (measured ~6 %)
    const long long tm = (op.M + 127) / 128, z = (long long)op.Z1 * op.Z2, sms = sm_count();
    const long long w256 = (tm * ((op.N + 255) / 256) * z + sms - 1) / sms * 256;
    const long long w192 = (tm * ((op.N + 191) / 192) * z + sms - 1) / sms * 192;
    bn = (w192 * 106 < w256 * 100) ? 192 : 256;
  }
  return bn == 192 ? launch_conv_bn<192>(op, v, s) : launch_conv_bn<256>(op, v, s);
}

bool dispatch(const GemmOp& op, bool dry, cudaStream_t s, const char** why) {
  Variant v;
  v.scalar = false;
  if (!classify(op, v, why)) return false;
  t_scalar_store = v.scalar ? 1 : 0;
  if (v.b_mn) {
    if (v.a_mn) { if (why) *why = "MN-major x MN-major is not compiled"; return false; }
    static const uint32_t ok[] = {G | P, AD | F, AD | P, AD | G | P, AD | G | F, RS | F, RS | P, F, P, G | F};
    bool found = false;
    for (uint32_t x : ok) found |= (x == v.ef);
    if (!v.nc || !found) { if (why) *why = "epilogue combination not compiled for MN-major B"; return false; }
    if (!dry) launch_conv(op, v, s);
    return true;
  }
  if (op.cplx == 2) {
    if (dry) return true;
    if (op.group_n == 64) launch<Cfg<64, false, false, P, true, 32, true>>(op, s);
    else launch<Cfg<128, false, false, P, true, 32, true>>(op, s);
    return true;
  }
  if (op.cplx) {
    if (v.nc || (op.epi.flags & EPI_ROW_AFFINE)) { if (why) *why = "complex mode is compiled for the plain ROWC epilogue only"; return false; }
    if (!dry) launch<Cfg<128, false, false, P, false, 32, true>>(op, s);
    return true;
  }
  if (op.bfly) {
    if (!(v.nc && v.a_mn && v.ef == F)) { if (why) *why = "butterfly mode is compiled for MN-major A with fp32 output"; return false; }
    if (dry) return true;
    // CTA pairs: the small-N MMAs of this stage are bound by their shared-memory operand reads (4 KB of A per 96-column MMA);
    // a pair halves the B bytes each CTA stages and reads
    if (options().pair != 0 && options().bfly_pair && op.M % 256 == 0) launch<Cfg<96, true, false, F, true, 32, false, true, true>>(op, s);
    else launch<Cfg<96, true, false, F, true, 32, false, false, true>>(op, s);
    return true;
  }
  const bool ok = (!v.nc && !v.a_mn) || (v.nc && (v.ef == P || v.ef == F));
  if (!ok) { if (why) *why = "epilogue combination not compiled for K-major B"; return false; }
  if (dry) return true;
  // N tile: least padded columns, ties to the larger tile
  // forward SHT stages (K-major x K-major, ROWC) gain ~4 us each from CTA pairs when no half tile is wasted
  const bool want_pair = options().pair == 1 ? op.M > 128 : (options().pair < 0 && !v.nc && !v.a_mn && !v.b_mn && op.M % 256 == 0);
  if (want_pair && launch_variant_pair(op, v, s)) return true;
  int bn = options().umma_bn;
  if (bn == 128 && !v.nc) return launch_variant<128>(op, v, s);
  if (bn != 192 && bn != 256) {
    long long w192 = (op.N + 191) / 192 * 192 - op.N, w256 = (op.N + 255) / 256 * 256 - op.N;
    bn = (w192 < w256) ? 192 : 256;
  }
  return bn == 192 ? launch_variant<192>(op, v, s) : launch_variant<256>(op, v, s);
}

}  // namespace

bool umma_eligible(const GemmOp& op, const char** why) { return dispatch(op, true, nullptr, why); }

// ConditionalLayerNorm mode: returns false when the op is outside the compiled variant (the caller then uses cln.cu's kernel)
bool run_gemm_cln(const GemmOp& op, cudaStream_t stream) {
  const EpiParams& e = op.epi;
  const bool ok = op.cln && op.A.s_k == 1 && op.B.s_k == 1 && aligned8(op.A.s_row) && aligned8(op.B.s_row) && aligned8(op.A.plane) &&
                  aligned8(op.B.plane) && aligned8(op.A.s_z2) && aligned8(op.B.s_z2) && !(((uintptr_t)op.A.ptr) & 15) &&
                  !(((uintptr_t)op.B.ptr) & 15) && op.K % 64 == 0 && op.k_split * 2 == op.K && op.Z1 == 1 && aligned4(op.N) &&
                  e.flags == (EPI_RES_PLANES | EPI_OUT_PLANES) && e.res_n == 1 && e.o_n == 1 && aligned4(e.res_m0) && aligned4(e.res_z2) &&
                  aligned4(e.res_plane) && !(((uintptr_t)e.res) & 7) && aligned4(e.o_m0) && aligned4(e.o_z2) && aligned4(e.out_plane) &&
                  !(((uintptr_t)e.out) & 7) && e.cln_musr && !(((uintptr_t)e.cln_musr) & 15) && (e.cln_musr_z2 & 1) == 0 && op.M >= 1 &&
                  e.mdiv >= op.M;
  if (!ok) return false;
  t_scalar_store = 0;
  launch<Cfg<128, false, false, RS | P, true, 32, false, false, false, false, 8, true>>(op, stream);
  return true;
}

void run_gemm_umma(const GemmOp& op, cudaStream_t stream) {
  const char* why = nullptr;
  bool ok = dispatch(op, false, stream, &why);
  ACE_REQUIRE(ok, "gemm %s: not eligible for the tcgen05 kernel: %s", op.name, why ? why : "?");
}

}  // namespace ace
