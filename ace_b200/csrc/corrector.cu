// Conservation correctors of the post-step state: the two that the ACE2 baseline
// config enables (configs/baselines/amip-c96-shield/train-ace2.yaml:128-140) besides ForcePositive.
//
// Reference (/root/reference/fme/core/corrector/atmosphere.py):
//   :404-427  _seed_global_dry_air_mass     target = area-weighted mean of (ps - g * total_water_path) of the initial condition
//   :430-463  _adjust_gen_dry_air_to_target add a global constant to the dry-air pressure, solve for the surface pressure
//   :518-608  _force_conserve_moisture      scale precipitation (or evaporation) so the GLOBAL moisture budget closes, then
//                                           (advection_and_*) recompute the advective tendency as the COLUMN budget residual
// with fme/core/atmosphere_data.py:180-197,253-289 (dry air, total water path, evaporation = LHF / Lv) and
// fme/core/coordinates.py:241-284 (interface pressure a_k + b_k ps, vertical integral (1/g) sum x dp).
// Area-weighted means accumulate in fp64 (the reference uses fp64 for the dry-air pin and fp32 for the budget means).
#include <vector>

#include "common.cuh"

namespace {
constexpr double kGravity = 9.80665;               // fme/core/constants.py:3
constexpr double kLatentHeatVaporization = 2.5e6;  // fme/core/constants.py:1
constexpr int kMaxNz = 32;
}  // namespace

struct ace_corrector {
  int n_out, n_prog, nz;
  long long hw;
  int out_ps, out_precip, out_lhf, out_adv, prog_ps;
  int out_wat[kMaxNz], prog_wat[kMaxNz];
  int conserve_dry_air, moisture_mode;
  double dt, wsum;
  ace::DevBuf w, akd, bkd, out_prog, target, sums, tend;
  int cap_b = 0;
  bool seeded = false;
};

namespace ace {
namespace {

struct Idx {
  int out_ps, out_precip, out_lhf, out_adv, prog_ps, nz;
  int out_wat[kMaxNz], prog_wat[kMaxNz];
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

// sums[b][0] += sum_pixels w * (ps - g*twp): `prog` layout (seed from the initial condition) or `out` layout (generated)
__global__ void __launch_bounds__(256) dry_air_reduce_kernel(const float* __restrict__ data, int nchan, Idx ix, bool from_prog,
                                                            long long hw, const float* __restrict__ w,
                                                            const double* __restrict__ akd, const double* __restrict__ bkd,
                                                            double* __restrict__ sums, int slot) {
  const int b = blockIdx.y;
  const float* base = data + (long long)b * nchan * hw;
  double acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const double ps = base[(long long)(from_prog ? ix.prog_ps : ix.out_ps) * hw + i];
    auto wat = [&](int k) { return base[(long long)(from_prog ? ix.prog_wat[k] : ix.out_wat[k]) * hw + i]; };
    acc += (double)w[i] * (ps - g_twp(wat, ps, akd, bkd, ix.nz));
  }
  block_sum_atomic(acc, sums + b * 4 + slot);
}

// dry-air pin (atmosphere.py:430-463) + the three global means of the moisture budget (:556-561); tend scratch [B][hw]
__global__ void __launch_bounds__(256) dry_air_apply_kernel(float* __restrict__ out, const float* __restrict__ prev, float* next_prog,
                                                           int n_out, int n_prog, Idx ix, const int* __restrict__ out_prog,
                                                           long long hw, const float* __restrict__ w, const double* __restrict__ akd,
                                                           const double* __restrict__ bkd, const double* __restrict__ target,
                                                           double wsum, double* __restrict__ sums, int conserve_dry_air,
                                                           int moisture, double dt, float* __restrict__ tend) {
  const int b = blockIdx.y;
  float* ob = out + (long long)b * n_out * hw;
  const float* pb = prev + (long long)b * n_prog * hw;
  const double err = conserve_dry_air ? sums[b * 4 + 0] / wsum - target[b] : 0.0;
  double a_t = 0, a_e = 0, a_p = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    auto wat = [&](int k) { return ob[(long long)ix.out_wat[k] * hw + i]; };
    float ps = ob[(long long)ix.out_ps * hw + i];
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
      const double wi = w[i];
      a_t += wi * td;
      a_e += wi * ((double)ob[(long long)ix.out_lhf * hw + i] / kLatentHeatVaporization);
      a_p += wi * (double)ob[(long long)ix.out_precip * hw + i];
    }
  }
  if (moisture) {
    block_sum_atomic(a_t, sums + b * 4 + 1);
    block_sum_atomic(a_e, sums + b * 4 + 2);
    block_sum_atomic(a_p, sums + b * 4 + 3);
  }
}

// mode 1 precipitation, 2 advection_and_precipitation, 3 evaporation, 4 advection_and_evaporation (atmosphere.py:562-608)
__global__ void __launch_bounds__(256) moisture_apply_kernel(float* __restrict__ out, float* next_prog, int n_out, int n_prog, Idx ix,
                                                            const int* __restrict__ out_prog, long long hw,
                                                            const double* __restrict__ sums, double wsum, int mode,
                                                            const float* __restrict__ tend) {
  const int b = blockIdx.y;
  float* ob = out + (long long)b * n_out * hw;
  // global means and their ratio in fp64 (m_t + m_p can cancel to a fraction of either term); rounded once
  const double m_t = sums[b * 4 + 1] / wsum, m_e = sums[b * 4 + 2] / wsum, m_p = sums[b * 4 + 3] / wsum;
  const bool fix_precip = (mode == 1 || mode == 2);
  const float ratio = (float)(fix_precip ? (m_e - m_t) / m_p : (m_t + m_p) / m_e);
  const float lv = (float)kLatentHeatVaporization;
  auto put = [&](int chan, long long i, float v) {
    ob[(long long)chan * hw + i] = v;
    const int p = out_prog[chan];
    if (p >= 0 && next_prog) next_prog[((long long)b * n_prog + p) * hw + i] = v;
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
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
  }
}

Idx make_idx(const ace_corrector& c) {
  Idx ix;
  ix.out_ps = c.out_ps;
  ix.out_precip = c.out_precip;
  ix.out_lhf = c.out_lhf;
  ix.out_adv = c.out_adv;
  ix.prog_ps = c.prog_ps;
  ix.nz = c.nz;
  for (int k = 0; k < kMaxNz; ++k) {
    ix.out_wat[k] = c.out_wat[k];
    ix.prog_wat[k] = c.prog_wat[k];
  }
  return ix;
}

void ensure_batch(ace_corrector& c, int B) {
  if (B <= c.cap_b) return;
  c.target.ensure((size_t)B * sizeof(double));
  c.sums.ensure((size_t)B * 4 * sizeof(double));
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
  auto chan_ok = [&](int c, int n) { return c >= 0 && c < n; };
  ACE_REQUIRE(chan_ok(cfg->out_ps, cfg->n_out) && chan_ok(cfg->prog_ps, cfg->n_prog), "ace_corrector_create: surface pressure index out of range");
  if (cfg->moisture_mode) {
    ACE_REQUIRE(chan_ok(cfg->out_precip, cfg->n_out) && chan_ok(cfg->out_lhf, cfg->n_out), "ace_corrector_create: precipitation / latent heat flux index out of range");
    if (cfg->moisture_mode == 2 || cfg->moisture_mode == 4)
      ACE_REQUIRE(chan_ok(cfg->out_adv, cfg->n_out), "ace_corrector_create: advective tendency index out of range");
    ACE_REQUIRE(cfg->timestep_seconds > 0, "ace_corrector_create: timestep must be positive");
  }
  ace_corrector* c = new ace_corrector();
  try {
    c->n_out = cfg->n_out;
    c->n_prog = cfg->n_prog;
    c->nz = cfg->nz;
    c->hw = cfg->hw;
    c->out_ps = cfg->out_ps;
    c->out_precip = cfg->out_precip;
    c->out_lhf = cfg->out_lhf;
    c->out_adv = cfg->out_adv;
    c->prog_ps = cfg->prog_ps;
    c->conserve_dry_air = cfg->conserve_dry_air ? 1 : 0;
    c->moisture_mode = cfg->moisture_mode;
    c->dt = cfg->timestep_seconds;
    for (int k = 0; k < kMaxNz; ++k) c->out_wat[k] = c->prog_wat[k] = 0;
    for (int k = 0; k < cfg->nz; ++k) {
      ACE_REQUIRE(chan_ok(cfg->out_wat_host[k], cfg->n_out) && chan_ok(cfg->prog_wat_host[k], cfg->n_prog), "ace_corrector_create: water index out of range");
      c->out_wat[k] = cfg->out_wat_host[k];
      c->prog_wat[k] = cfg->prog_wat_host[k];
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
    ACE_CHECK_CUDA(cudaMemsetAsync(c->sums.p, 0, (size_t)batch * 4 * sizeof(double), s));
    dry_air_reduce_kernel<<<grid_for(*c, batch), 256, 0, s>>>(prog_dev, c->n_prog, make_idx(*c), true, c->hw, c->w.as<float>(),
                                                             c->akd.as<double>(), c->bkd.as<double>(), c->sums.as<double>(), 0);
    after_launch("dry_air_seed");
    // target[b] = sums[b][0] / wsum, kept in fp64 on the device
    std::vector<double> h((size_t)batch * 4);
    ACE_CHECK_CUDA(cudaMemcpyAsync(h.data(), c->sums.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    ACE_CHECK_CUDA(cudaStreamSynchronize(s));  // once per rollout
    std::vector<double> t(batch);
    for (int b = 0; b < batch; ++b) t[b] = h[(size_t)b * 4] / c->wsum;
    ACE_CHECK_CUDA(cudaMemcpyAsync(c->target.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    ACE_CHECK_CUDA(cudaStreamSynchronize(s));
  }
  c->seeded = true;
  ACE_API_END
}

extern "C" int ace_corrector_is_seeded(ace_corrector* c) { return (c && c->seeded) ? 1 : 0; }

extern "C" int ace_corrector_apply(ace_corrector* c, const float* prev_prog_dev, float* out_dev, float* next_prog_dev, int batch,
                                   void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(c && prev_prog_dev && out_dev && batch > 0, "ace_corrector_apply: bad argument");
  if (!c->seeded || batch > c->cap_b)
    throw Error(ACE_ERR_STATE, "ace_corrector_apply: call ace_corrector_seed with the initial condition first (once per rollout)");
  if (!c->conserve_dry_air && !c->moisture_mode) return ACE_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const Idx ix = make_idx(*c);
  const dim3 grid = grid_for(*c, batch);
  ACE_CHECK_CUDA(cudaMemsetAsync(c->sums.p, 0, (size_t)batch * 4 * sizeof(double), s));
  if (c->conserve_dry_air) {
    ProfileScope prof("corrector.dry_air_reduce", s);
    dry_air_reduce_kernel<<<grid, 256, 0, s>>>(out_dev, c->n_out, ix, false, c->hw, c->w.as<float>(), c->akd.as<double>(),
                                               c->bkd.as<double>(), c->sums.as<double>(), 0);
    after_launch("dry_air_reduce");
  }
  {
    ProfileScope prof("corrector.dry_air_apply", s);
    dry_air_apply_kernel<<<grid, 256, 0, s>>>(out_dev, prev_prog_dev, next_prog_dev, c->n_out, c->n_prog, ix, c->out_prog.as<int>(), c->hw,
                                              c->w.as<float>(), c->akd.as<double>(), c->bkd.as<double>(), c->target.as<double>(), c->wsum,
                                              c->sums.as<double>(), c->conserve_dry_air, c->moisture_mode, c->dt, c->tend.as<float>());
    after_launch("dry_air_apply");
  }
  if (c->moisture_mode) {
    ProfileScope prof("corrector.moisture_apply", s);
    moisture_apply_kernel<<<grid, 256, 0, s>>>(out_dev, next_prog_dev, c->n_out, c->n_prog, ix, c->out_prog.as<int>(), c->hw,
                                               c->sums.as<double>(), c->wsum, c->moisture_mode, c->tend.as<float>());
    after_launch("moisture_apply");
  }
  ACE_API_END
}
