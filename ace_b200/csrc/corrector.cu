// Conservation correctors of the post-step state: everything the ACE2 baseline configs
// enable in `corrector:` besides ForcePositive (configs/baselines/era5/ace-train-config-1-step-pretrain.yaml:119-133).
//
// Reference (/root/reference/fme/core/corrector/atmosphere.py), in the order :349-398 builds the sequence:
//   :404-427  _seed_global_dry_air_mass     target = area-weighted mean of (ps - g * total_water_path) of the initial condition
//   :430-463  _adjust_gen_dry_air_to_target add a global constant to the dry-air pressure, solve for the surface pressure
//   :467-490  _force_zero_global_mean_moisture_advection   subtract the global mean of the advective tendency
//   :518-608  _force_conserve_moisture      scale precipitation (or evaporation) so the GLOBAL moisture budget closes, then
//                                           (advection_and_*) recompute the advective tendency as the COLUMN budget residual
//   :493-515  _clip_frozen_precipitation    frozen precipitation <= corrected total precipitation
//   :611-695  _force_conserve_total_energy  one global temperature offset so that the global mean column energy
//                                           (cv T + Lv q + g z, hydrostatic z) changes by the predicted net flux * dt
// with fme/core/atmosphere_data.py:180-197,217-289,291-365,376-418 (dry air, total water path, energy fluxes, layer thickness,
// interface heights, total_energy_ace2_path) and fme/core/coordinates.py:241-284 (interface pressure a_k + b_k ps, vertical
// integral (1/g) sum x dp).  Column integrals and area-weighted means accumulate in fp64 (the reference uses fp64 for the
// dry-air pin only and fp32 elsewhere).
#include <vector>

#include "common.cuh"

namespace {
constexpr double kGravity = 9.80665;               // fme/core/constants.py
constexpr double kLatentHeatVaporization = 2.5e6;
constexpr double kLatentHeatFreezing = 334000.0;
constexpr double kRvgas = 461.5, kRdgas = 287.05;
constexpr double kCv = 1004.6 - 287.05;            // SPECIFIC_HEAT_OF_DRY_AIR_CONST_VOLUME
constexpr int kMaxNz = 32;
// per-sample fp64 accumulators
enum Slot { S_DRY = 0, S_TEND, S_EVAP, S_PRECIP, S_ADV, S_EGEN, S_EIN, S_FLUX, S_FACTOR, kSlots = 12 };
}  // namespace

struct ace_corrector {
  int n_out, n_prog, nz, n_forcing;
  long long hw;
  int out_ps, out_precip, out_lhf, out_adv, prog_ps, out_frozen, forcing_hgt;
  int out_flux[7];  // dlw_sfc, ulw_sfc, dsw_sfc, usw_sfc, shf, usw_toa, ulw_toa
  int out_wat[kMaxNz], prog_wat[kMaxNz], out_temp[kMaxNz], prog_temp[kMaxNz];
  int conserve_dry_air, moisture_mode, zero_adv, clip_frozen, energy_mode;
  double dt, wsum, unaccounted_heating;
  ace::DevBuf w, ak, bk, akd, bkd, out_prog, target, sums, tend;
  int cap_b = 0;
  bool seeded = false;
};

namespace ace {
namespace {

struct Idx {
  int out_ps, out_precip, out_lhf, out_adv, prog_ps, out_frozen, forcing_hgt, nz;
  int out_flux[7];
  int out_wat[kMaxNz], prog_wat[kMaxNz], out_temp[kMaxNz], prog_temp[kMaxNz];
};

__device__ __forceinline__ double block_sum_atomic(double v, double* dst) {
  __shared__ double red[8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(dst, s);
  }
  __syncthreads();
  return v;
}

// sum_k (akd_k + bkd_k * ps) * wat_k = g * total_water_path (coordinates.py:282-284)
template <class F>
__device__ __forceinline__ double g_twp(F wat, double ps, const double* akd, const double* bkd, int nz) {
  double s = 0;
  for (int k = 0; k < nz; ++k) s += (akd[k] + bkd[k] * ps) * (double)wat(k);
  return s;
}

// Column total energy path (atmosphere_data.py:341-365) and the temperature-correction factor (atmosphere.py:668-695), one
// bottom-up pass: the reversed cumulative sums of the reference are running sums from the surface.
template <class FT, class FQ>
__device__ __forceinline__ double column_energy(FT temp, FQ wat, double ps, double hgt, const double* ak, const double* bk, int nz,
                                                double* factor) {
  const double hsfc = hgt < 0.0 ? 0.0 : hgt;  // negative surface heights are filled with 0 (atmosphere_data.py:408-411)
  double cum = 0, cumq = 0, e = 0, f = 0;
  double p_lo = ak[nz] + bk[nz] * ps;
  double log_lo = log(fmax(p_lo, 1.0));      // TOA pressure clamped to 1 Pa before the log (atmosphere_data.py:386-392)
  for (int k = nz - 1; k >= 0; --k) {
    const double p_hi = ak[k] + bk[k] * ps, log_hi = log(fmax(p_hi, 1.0));
    const double dp = p_lo - p_hi, dlogp = log_lo - log_hi;
    const double t = (double)temp(k), q = (double)wat(k);
    const double thick = dlogp * kRdgas * (t * (1.0 + (kRvgas / kRdgas - 1.0) * q)) / kGravity;
    const double h_bot = cum + hsfc;
    cum += thick;
    const double h_mid = 0.5 * (h_bot + cum + hsfc);
    e += (t * kCv + q * kLatentHeatVaporization + h_mid * kGravity) * dp;
    const double qd = thick * kGravity / t;
    cumq += qd;
    f += (kCv - 0.5 * qd + cumq) * dp;
    p_lo = p_hi;
    log_lo = log_hi;
  }
  *factor = f / kGravity;
  return e / kGravity;
}

// sums[b][S_DRY] += sum_pixels w * (ps - g*twp): `prog` layout (seed from the initial condition) or `out` layout (generated)
__global__ void __launch_bounds__(256) dry_air_reduce_kernel(const float* __restrict__ data, int nchan, Idx ix, bool from_prog,
                                                            long long hw, const float* __restrict__ w,
                                                            const double* __restrict__ akd, const double* __restrict__ bkd,
                                                            double* __restrict__ sums) {
  const int b = blockIdx.y;
  const float* base = data + (long long)b * nchan * hw;
  double acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const double ps = base[(long long)(from_prog ? ix.prog_ps : ix.out_ps) * hw + i];
    auto wat = [&](int k) { return base[(long long)(from_prog ? ix.prog_wat[k] : ix.out_wat[k]) * hw + i]; };
    acc += (double)w[i] * (ps - g_twp(wat, ps, akd, bkd, ix.nz));
  }
  block_sum_atomic(acc, sums + b * kSlots + S_DRY);
}

// dry-air pin (atmosphere.py:430-463) + the global sums of the moisture budget (:556-561) and of the advective tendency
// (:483-485); tend scratch [B][hw]
__global__ void __launch_bounds__(256) dry_air_apply_kernel(float* __restrict__ out, const float* __restrict__ prev, float* next_prog,
                                                           int n_out, int n_prog, Idx ix, const int* __restrict__ out_prog,
                                                           long long hw, const float* __restrict__ w, const double* __restrict__ akd,
                                                           const double* __restrict__ bkd, const double* __restrict__ target,
                                                           double wsum, double* __restrict__ sums, int conserve_dry_air,
                                                           int moisture, int zero_adv, double dt, float* __restrict__ tend) {
  const int b = blockIdx.y;
  float* ob = out + (long long)b * n_out * hw;
  const float* pb = prev + (long long)b * n_prog * hw;
  const double err = conserve_dry_air ? sums[b * kSlots + S_DRY] / wsum - target[b] : 0.0;
  double a_t = 0, a_e = 0, a_p = 0, a_a = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    auto wat = [&](int k) { return ob[(long long)ix.out_wat[k] * hw + i]; };
    float ps = ob[(long long)ix.out_ps * hw + i];
    const double wi = w[i];
    if (conserve_dry_air) {
      const double dry = (double)ps - g_twp(wat, (double)ps, akd, bkd, ix.nz) - err;
      double sa = 0, sb = 0;
      for (int k = 0; k < ix.nz; ++k) {
        sa += akd[k] * (double)wat(k);
        sb += bkd[k] * (double)wat(k);
      }
      ps = (float)((dry + sa) / (1.0 - sb));
      ob[(long long)ix.out_ps * hw + i] = ps;
      const int p = out_prog[ix.out_ps];
      if (p >= 0 && next_prog) next_prog[((long long)b * n_prog + p) * hw + i] = ps;
    }
    if (moisture) {
      auto wat_in = [&](int k) { return pb[(long long)ix.prog_wat[k] * hw + i]; };
      const double twp_gen = g_twp(wat, (double)ps, akd, bkd, ix.nz) / kGravity;
      const double twp_in = g_twp(wat_in, (double)pb[(long long)ix.prog_ps * hw + i], akd, bkd, ix.nz) / kGravity;
      const float td = (float)((twp_gen - twp_in) / dt);
      tend[(long long)b * hw + i] = td;
      a_t += wi * td;
      a_e += wi * ((double)ob[(long long)ix.out_lhf * hw + i] / kLatentHeatVaporization);
      a_p += wi * (double)ob[(long long)ix.out_precip * hw + i];
    }
    if (zero_adv) a_a += wi * (double)ob[(long long)ix.out_adv * hw + i];
  }
  if (moisture) {
    block_sum_atomic(a_t, sums + b * kSlots + S_TEND);
    block_sum_atomic(a_e, sums + b * kSlots + S_EVAP);
    block_sum_atomic(a_p, sums + b * kSlots + S_PRECIP);
  }
  if (zero_adv) block_sum_atomic(a_a, sums + b * kSlots + S_ADV);
}

// mode 0 none, 1 precipitation, 2 advection_and_precipitation, 3 evaporation, 4 advection_and_evaporation (atmosphere.py:562-608);
// zero_adv: subtract the global mean of the advective tendency first (it is overwritten when the mode recomputes advection);
// clip_frozen: frozen precipitation <= corrected total precipitation.
__global__ void __launch_bounds__(256) moisture_apply_kernel(float* __restrict__ out, float* next_prog, int n_out, int n_prog, Idx ix,
                                                            const int* __restrict__ out_prog, long long hw,
                                                            const double* __restrict__ sums, double wsum, int mode, int zero_adv,
                                                            int clip_frozen, const float* __restrict__ tend) {
  const int b = blockIdx.y;
  float* ob = out + (long long)b * n_out * hw;
  // global means and their ratio in fp64 (m_t + m_p can cancel to a fraction of either term); rounded once
  const double m_t = sums[b * kSlots + S_TEND] / wsum, m_e = sums[b * kSlots + S_EVAP] / wsum, m_p = sums[b * kSlots + S_PRECIP] / wsum;
  const bool fix_precip = (mode == 1 || mode == 2);
  const float ratio = mode ? (float)(fix_precip ? (m_e - m_t) / m_p : (m_t + m_p) / m_e) : 1.f;
  const float adv_mean = (float)(sums[b * kSlots + S_ADV] / wsum);
  const float lv = (float)kLatentHeatVaporization;
  auto put = [&](int chan, long long i, float v) {
    ob[(long long)chan * hw + i] = v;
    const int p = out_prog[chan];
    if (p >= 0 && next_prog) next_prog[((long long)b * n_prog + p) * hw + i] = v;
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    if (zero_adv && mode != 2 && mode != 4) put(ix.out_adv, i, ob[(long long)ix.out_adv * hw + i] - adv_mean);
    if (!mode) continue;
    float precip = ob[(long long)ix.out_precip * hw + i];
    float evap = ob[(long long)ix.out_lhf * hw + i] / lv;
    if (fix_precip) {
      precip *= ratio;
      put(ix.out_precip, i, precip);
    } else {
      const float lhf = evap * ratio * lv;  // the reference stores the flux and re-derives the rate from it
      put(ix.out_lhf, i, lhf);
      evap = lhf / lv;
    }
    if (mode == 2 || mode == 4) put(ix.out_adv, i, tend[(long long)b * hw + i] - (evap - precip));
    if (clip_frozen) put(ix.out_frozen, i, fminf(ob[(long long)ix.out_frozen * hw + i], precip));
  }
}

// global sums of the energy budget (atmosphere.py:634-646,653-655): column energy of the generated and of the input state,
// predicted net flux into the atmosphere, temperature-correction factor.  next [B][2][hw] = (TOA downward shortwave, surface
// height) at the output time; the input state's surface height is channel forcing_hgt of the step's forcing input.
__global__ void __launch_bounds__(256) energy_reduce_kernel(const float* __restrict__ out, const float* __restrict__ prev,
                                                           const float* __restrict__ prev_forcing, const float* __restrict__ next,
                                                           int n_out, int n_prog, int n_forcing, Idx ix, long long hw,
                                                           const float* __restrict__ w, const double* __restrict__ ak,
                                                           const double* __restrict__ bk, double* __restrict__ sums) {
  const int b = blockIdx.y;
  const float* ob = out + (long long)b * n_out * hw;
  const float* pb = prev + (long long)b * n_prog * hw;
  const float* fb = prev_forcing + (long long)b * n_forcing * hw;
  const float* nb = next + (long long)b * 2 * hw;
  double a_g = 0, a_i = 0, a_f = 0, a_c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const double wi = w[i];
    double factor, unused;
    const double e_gen = column_energy([&](int k) { return ob[(long long)ix.out_temp[k] * hw + i]; },
                                       [&](int k) { return ob[(long long)ix.out_wat[k] * hw + i]; },
                                       (double)ob[(long long)ix.out_ps * hw + i], (double)nb[hw + i], ak, bk, ix.nz, &factor);
    const double e_in = column_energy([&](int k) { return pb[(long long)ix.prog_temp[k] * hw + i]; },
                                      [&](int k) { return pb[(long long)ix.prog_wat[k] * hw + i]; },
                                      (double)pb[(long long)ix.prog_ps * hw + i], (double)fb[(long long)ix.forcing_hgt * hw + i], ak, bk,
                                      ix.nz, &unused);
    auto o = [&](int c) { return (double)ob[(long long)c * hw + i]; };
    const double frozen = ix.out_frozen >= 0 ? o(ix.out_frozen) * kLatentHeatFreezing : 0.0;
    // metrics.py:299-352: surface = (dsw - usw + dlw - ulw) + (-lhf - shf) - frozen * Lf, toa = dsw_toa - usw_toa - ulw_toa
    const double sfc = (o(ix.out_flux[2]) - o(ix.out_flux[3]) + o(ix.out_flux[0]) - o(ix.out_flux[1])) + (-o(ix.out_lhf) - o(ix.out_flux[4])) - frozen;
    const double toa = (double)nb[i] - o(ix.out_flux[5]) - o(ix.out_flux[6]);
    a_g += wi * e_gen;
    a_i += wi * e_in;
    a_f += wi * (toa - sfc);
    a_c += wi * factor;
  }
  block_sum_atomic(a_g, sums + b * kSlots + S_EGEN);
  block_sum_atomic(a_i, sums + b * kSlots + S_EIN);
  block_sum_atomic(a_f, sums + b * kSlots + S_FLUX);
  block_sum_atomic(a_c, sums + b * kSlots + S_FACTOR);
}

// the same temperature offset on every level (atmosphere.py:648-665)
__global__ void __launch_bounds__(256) energy_apply_kernel(float* __restrict__ out, float* next_prog, int n_out, int n_prog, Idx ix,
                                                          const int* __restrict__ out_prog, long long hw, const double* __restrict__ sums,
                                                          double wsum, double dt, double unaccounted_heating) {
  const int b = blockIdx.y;
  float* ob = out + (long long)b * n_out * hw;
  const double* s = sums + b * kSlots;
  const double desired = s[S_EIN] / wsum + (s[S_FLUX] / wsum + unaccounted_heating) * dt;
  const float d_t = (float)((desired - s[S_EGEN] / wsum) / (s[S_FACTOR] / wsum));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    for (int k = 0; k < ix.nz; ++k) {
      const int c = ix.out_temp[k];
      const float v = ob[(long long)c * hw + i] + d_t;
      ob[(long long)c * hw + i] = v;
      const int p = out_prog[c];
      if (p >= 0 && next_prog) next_prog[((long long)b * n_prog + p) * hw + i] = v;
    }
  }
}

Idx make_idx(const ace_corrector& c) {
  Idx ix;
  ix.out_ps = c.out_ps;
  ix.out_precip = c.out_precip;
  ix.out_lhf = c.out_lhf;
  ix.out_adv = c.out_adv;
  ix.prog_ps = c.prog_ps;
  ix.out_frozen = c.out_frozen;
  ix.forcing_hgt = c.forcing_hgt;
  ix.nz = c.nz;
  for (int k = 0; k < 7; ++k) ix.out_flux[k] = c.out_flux[k];
  for (int k = 0; k < kMaxNz; ++k) {
    ix.out_wat[k] = c.out_wat[k];
    ix.prog_wat[k] = c.prog_wat[k];
    ix.out_temp[k] = c.out_temp[k];
    ix.prog_temp[k] = c.prog_temp[k];
  }
  return ix;
}

void ensure_batch(ace_corrector& c, int B) {
  if (B <= c.cap_b) return;
  c.target.ensure((size_t)B * sizeof(double));
  c.sums.ensure((size_t)B * kSlots * sizeof(double));
  c.tend.ensure((size_t)B * c.hw * sizeof(float));
  c.cap_b = B;
  c.seeded = false;
}

dim3 grid_for(const ace_corrector& c, int B) { return dim3((unsigned)std::min<long long>((c.hw + 255) / 256, 148 * 4 / std::max(1, B) + 1), B); }

}  // namespace
}  // namespace ace

using namespace ace;

extern "C" int ace_corrector_create(const ace_corrector_config* cfg, ace_corrector** out) {
  ACE_API_BEGIN
  ACE_REQUIRE(cfg && out, "ace_corrector_create: null argument");
  ACE_REQUIRE(cfg->nz >= 1 && cfg->nz <= kMaxNz, "ace_corrector_create: nz must be in [1, %d]", kMaxNz);
  ACE_REQUIRE(cfg->hw > 0 && cfg->n_out > 0 && cfg->n_prog > 0, "ace_corrector_create: bad sizes");
  ACE_REQUIRE(cfg->area_weights_host && cfg->ak_host && cfg->bk_host && cfg->out_wat_host && cfg->prog_wat_host && cfg->out_prog_index_host,
              "ace_corrector_create: null table");
  ACE_REQUIRE(cfg->moisture_mode >= 0 && cfg->moisture_mode <= 4, "ace_corrector_create: moisture_mode must be 0..4");
  ACE_REQUIRE(cfg->energy_mode == 0 || cfg->energy_mode == 1, "ace_corrector_create: energy_mode must be 0 or 1 (constant_temperature)");
  auto chan_ok = [&](int c, int n) { return c >= 0 && c < n; };
  ACE_REQUIRE(chan_ok(cfg->out_ps, cfg->n_out) && chan_ok(cfg->prog_ps, cfg->n_prog), "ace_corrector_create: surface pressure index out of range");
  if (cfg->moisture_mode) {
    ACE_REQUIRE(chan_ok(cfg->out_precip, cfg->n_out) && chan_ok(cfg->out_lhf, cfg->n_out), "ace_corrector_create: precipitation / latent heat flux index out of range");
    ACE_REQUIRE(cfg->timestep_seconds > 0, "ace_corrector_create: timestep must be positive");
  }
  if (cfg->moisture_mode == 2 || cfg->moisture_mode == 4 || cfg->zero_global_mean_moisture_advection)
    ACE_REQUIRE(chan_ok(cfg->out_adv, cfg->n_out), "ace_corrector_create: advective tendency index out of range");
  if (cfg->out_frozen >= 0) ACE_REQUIRE(chan_ok(cfg->out_frozen, cfg->n_out), "ace_corrector_create: frozen precipitation index out of range");
  const bool clip = cfg->clip_frozen_precipitation && cfg->moisture_mode && cfg->out_frozen >= 0;  // no-op when the field is not predicted
  if (cfg->energy_mode) {
    ACE_REQUIRE(cfg->out_temp_host && cfg->prog_temp_host, "ace_corrector_create: energy correction needs the air temperature channels");
    ACE_REQUIRE(chan_ok(cfg->forcing_hgt, cfg->n_forcing), "ace_corrector_create: surface height channel out of range");
    ACE_REQUIRE(chan_ok(cfg->out_lhf, cfg->n_out), "ace_corrector_create: latent heat flux index out of range");
    const int fl[7] = {cfg->out_dlw_sfc, cfg->out_ulw_sfc, cfg->out_dsw_sfc, cfg->out_usw_sfc, cfg->out_shf, cfg->out_usw_toa, cfg->out_ulw_toa};
    for (int k = 0; k < 7; ++k) ACE_REQUIRE(chan_ok(fl[k], cfg->n_out), "ace_corrector_create: energy flux channel %d out of range", k);
    ACE_REQUIRE(cfg->timestep_seconds > 0, "ace_corrector_create: timestep must be positive");
  }
  ace_corrector* c = new ace_corrector();
  try {
    c->n_out = cfg->n_out;
    c->n_prog = cfg->n_prog;
    c->n_forcing = cfg->n_forcing;
    c->nz = cfg->nz;
    c->hw = cfg->hw;
    c->out_ps = cfg->out_ps;
    c->out_precip = cfg->out_precip;
    c->out_lhf = cfg->out_lhf;
    c->out_adv = cfg->out_adv;
    c->prog_ps = cfg->prog_ps;
    c->out_frozen = cfg->out_frozen;
    c->forcing_hgt = cfg->forcing_hgt;
    const int fl[7] = {cfg->out_dlw_sfc, cfg->out_ulw_sfc, cfg->out_dsw_sfc, cfg->out_usw_sfc, cfg->out_shf, cfg->out_usw_toa, cfg->out_ulw_toa};
    for (int k = 0; k < 7; ++k) c->out_flux[k] = fl[k];
    c->conserve_dry_air = cfg->conserve_dry_air ? 1 : 0;
    c->moisture_mode = cfg->moisture_mode;
    c->zero_adv = cfg->zero_global_mean_moisture_advection ? 1 : 0;
    c->clip_frozen = clip ? 1 : 0;
    c->energy_mode = cfg->energy_mode;
    c->unaccounted_heating = cfg->unaccounted_heating;
    c->dt = cfg->timestep_seconds;
    for (int k = 0; k < kMaxNz; ++k) c->out_wat[k] = c->prog_wat[k] = c->out_temp[k] = c->prog_temp[k] = 0;
    for (int k = 0; k < cfg->nz; ++k) {
      ACE_REQUIRE(chan_ok(cfg->out_wat_host[k], cfg->n_out) && chan_ok(cfg->prog_wat_host[k], cfg->n_prog), "ace_corrector_create: water index out of range");
      c->out_wat[k] = cfg->out_wat_host[k];
      c->prog_wat[k] = cfg->prog_wat_host[k];
      if (cfg->energy_mode) {
        ACE_REQUIRE(chan_ok(cfg->out_temp_host[k], cfg->n_out) && chan_ok(cfg->prog_temp_host[k], cfg->n_prog), "ace_corrector_create: air temperature index out of range");
        c->out_temp[k] = cfg->out_temp_host[k];
        c->prog_temp[k] = cfg->prog_temp_host[k];
      }
    }
    std::vector<double> akd(cfg->nz), bkd(cfg->nz);
    for (int k = 0; k < cfg->nz; ++k) {
      akd[k] = cfg->ak_host[k + 1] - cfg->ak_host[k];
      bkd[k] = cfg->bk_host[k + 1] - cfg->bk_host[k];
    }
    double ws = 0;
    for (long long i = 0; i < cfg->hw; ++i) ws += (double)cfg->area_weights_host[i];
    c->wsum = ws;
    auto up = [&](DevBuf& d, const void* src, size_t bytes) {
      d.ensure(bytes);
      ACE_CHECK_CUDA(cudaMemcpy(d.p, src, bytes, cudaMemcpyHostToDevice));
    };
    up(c->w, cfg->area_weights_host, (size_t)cfg->hw * sizeof(float));
    up(c->ak, cfg->ak_host, (size_t)(cfg->nz + 1) * sizeof(double));
    up(c->bk, cfg->bk_host, (size_t)(cfg->nz + 1) * sizeof(double));
    up(c->akd, akd.data(), akd.size() * sizeof(double));
    up(c->bkd, bkd.data(), bkd.size() * sizeof(double));
    up(c->out_prog, cfg->out_prog_index_host, (size_t)cfg->n_out * sizeof(int));
  } catch (...) {
    delete c;
    throw;
  }
  *out = c;
  ACE_API_END
}

extern "C" void ace_corrector_destroy(ace_corrector* c) { delete c; }

extern "C" int ace_corrector_reset(ace_corrector* c) {
  ACE_API_BEGIN
  ACE_REQUIRE(c, "ace_corrector_reset: null argument");
  c->seeded = false;
  ACE_API_END
}

extern "C" int ace_corrector_seed(ace_corrector* c, const float* prog_dev, int batch, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(c && prog_dev && batch > 0, "ace_corrector_seed: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  ensure_batch(*c, batch);
  if (c->conserve_dry_air) {
    ACE_CHECK_CUDA(cudaMemsetAsync(c->sums.p, 0, (size_t)batch * kSlots * sizeof(double), s));
    dry_air_reduce_kernel<<<grid_for(*c, batch), 256, 0, s>>>(prog_dev, c->n_prog, make_idx(*c), true, c->hw, c->w.as<float>(),
                                                             c->akd.as<double>(), c->bkd.as<double>(), c->sums.as<double>());
    after_launch("dry_air_seed");
    // target[b] = sums[b][S_DRY] / wsum, kept in fp64 on the device
    std::vector<double> h((size_t)batch * kSlots);
    ACE_CHECK_CUDA(cudaMemcpyAsync(h.data(), c->sums.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    ACE_CHECK_CUDA(cudaStreamSynchronize(s));  // once per rollout
    std::vector<double> t(batch);
    for (int b = 0; b < batch; ++b) t[b] = h[(size_t)b * kSlots + S_DRY] / c->wsum;
    ACE_CHECK_CUDA(cudaMemcpyAsync(c->target.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    ACE_CHECK_CUDA(cudaStreamSynchronize(s));
  }
  c->seeded = true;
  ACE_API_END
}

extern "C" int ace_corrector_is_seeded(ace_corrector* c) { return (c && c->seeded) ? 1 : 0; }

// The reference threads the dry-air target through CorrectorState.global_dry_air_mass (fme/core/corrector/state.py:15-29) from
// one prediction window to the next; these two calls read / install it (fp64, one value per sample, host memory).
extern "C" int ace_corrector_get_state(ace_corrector* c, double* target_host, int batch, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(c && target_host && batch > 0, "ace_corrector_get_state: bad argument");
  if (!c->seeded || batch > c->cap_b) throw Error(ACE_ERR_STATE, "ace_corrector_get_state: the corrector holds no state for this batch size");
  cudaStream_t s = (cudaStream_t)stream;
  if (c->conserve_dry_air) {
    ACE_CHECK_CUDA(cudaMemcpyAsync(target_host, c->target.p, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, s));
    ACE_CHECK_CUDA(cudaStreamSynchronize(s));
  } else {
    for (int b = 0; b < batch; ++b) target_host[b] = 0.0;
  }
  ACE_API_END
}

extern "C" int ace_corrector_set_state(ace_corrector* c, const double* target_host, int batch, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(c && target_host && batch > 0, "ace_corrector_set_state: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  ensure_batch(*c, batch);
  if (c->conserve_dry_air) {
    ACE_CHECK_CUDA(cudaMemcpyAsync(c->target.p, target_host, (size_t)batch * sizeof(double), cudaMemcpyHostToDevice, s));
    ACE_CHECK_CUDA(cudaStreamSynchronize(s));  // target_host may be a temporary of the caller
  }
  c->seeded = true;
  ACE_API_END
}

extern "C" int ace_corrector_needs_next(ace_corrector* c) { return (c && c->energy_mode) ? 1 : 0; }

extern "C" int ace_corrector_apply(ace_corrector* c, const float* prev_prog_dev, const float* prev_forcing_dev, const float* next_dev,
                                   float* out_dev, float* next_prog_dev, int batch, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(c && prev_prog_dev && out_dev && batch > 0, "ace_corrector_apply: bad argument");
  if (!c->seeded || batch > c->cap_b)
    throw Error(ACE_ERR_STATE, "ace_corrector_apply: call ace_corrector_seed with the initial condition first (once per rollout)");
  if (c->energy_mode && (!prev_forcing_dev || !next_dev))
    throw Error(ACE_ERR_INVALID, "ace_corrector_apply: the energy budget correction needs the step's forcing input (surface height) and "
                             "next_dev = [batch][2][hw] (TOA downward shortwave, surface height) at the output time");
  if (!c->conserve_dry_air && !c->moisture_mode && !c->zero_adv && !c->energy_mode) return ACE_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const Idx ix = make_idx(*c);
  const dim3 grid = grid_for(*c, batch);
  ACE_CHECK_CUDA(cudaMemsetAsync(c->sums.p, 0, (size_t)batch * kSlots * sizeof(double), s));
  if (c->conserve_dry_air) {
    ProfileScope prof("corrector.dry_air_reduce", s);
    dry_air_reduce_kernel<<<grid, 256, 0, s>>>(out_dev, c->n_out, ix, false, c->hw, c->w.as<float>(), c->akd.as<double>(),
                                               c->bkd.as<double>(), c->sums.as<double>());
    after_launch("dry_air_reduce");
  }
  if (c->conserve_dry_air || c->moisture_mode || c->zero_adv) {
    ProfileScope prof("corrector.dry_air_apply", s);
    dry_air_apply_kernel<<<grid, 256, 0, s>>>(out_dev, prev_prog_dev, next_prog_dev, c->n_out, c->n_prog, ix, c->out_prog.as<int>(), c->hw,
                                              c->w.as<float>(), c->akd.as<double>(), c->bkd.as<double>(), c->target.as<double>(), c->wsum,
                                              c->sums.as<double>(), c->conserve_dry_air, c->moisture_mode, c->zero_adv, c->dt,
                                              c->tend.as<float>());
    after_launch("dry_air_apply");
  }
  if (c->moisture_mode || c->zero_adv) {
    ProfileScope prof("corrector.moisture_apply", s);
    moisture_apply_kernel<<<grid, 256, 0, s>>>(out_dev, next_prog_dev, c->n_out, c->n_prog, ix, c->out_prog.as<int>(), c->hw,
                                               c->sums.as<double>(), c->wsum, c->moisture_mode, c->zero_adv, c->clip_frozen,
                                               c->tend.as<float>());
    after_launch("moisture_apply");
  }
  if (c->energy_mode) {
    {
      ProfileScope prof("corrector.energy_reduce", s);
      energy_reduce_kernel<<<grid, 256, 0, s>>>(out_dev, prev_prog_dev, prev_forcing_dev, next_dev, c->n_out, c->n_prog, c->n_forcing, ix,
                                                c->hw, c->w.as<float>(), c->ak.as<double>(), c->bk.as<double>(), c->sums.as<double>());
      after_launch("energy_reduce");
    }
    ProfileScope prof("corrector.energy_apply", s);
    energy_apply_kernel<<<grid, 256, 0, s>>>(out_dev, next_prog_dev, c->n_out, c->n_prog, ix, c->out_prog.as<int>(), c->hw,
                                             c->sums.as<double>(), c->wsum, c->dt, c->unaccounted_heating);
    after_launch("energy_apply");
  }
  ACE_API_END
}
