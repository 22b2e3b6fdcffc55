// Fused single-module step (ace_stepper_*): the device work of
// /root/reference/fme/core/step/single_module.py:595-665 with corrector/ocean disabled:
//   normalise (normalizer.py:213-243) -> pack (packer.py:45-52) -> network -> unpack ->
//   [+ normalised input for residual prediction] -> denormalise -> next prognostic state
// in two streaming kernels around ace_sfno_forward instead of ~100 elementwise launches.
#include <vector>

#include "kernels.cuh"

extern "C" int ace_sfno_forward(ace_sfno* net, const float* x_dev, float* y_dev, int batch, void* stream);
extern "C" int ace_csfno_forward(ace_csfno* net, const float* x_dev, const float* scalar_dev, const float* labels_dev, const float* noise_dev,
                                 const float* pos_dev, float* y_dev, int batch, void* stream);
extern "C" int ace_csfno_query(ace_csfno* net, int* in_chans, int* out_chans, long long* hw);

struct ace_stepper {
  ace_sfno* net = nullptr;
  ace_csfno* cnet = nullptr;               // noise-conditioned network instead of `net`
  const float* noise = nullptr;            // its context fields: persistent device addresses the caller refills before each step
  const float* pos = nullptr;
  int n_in, n_out, n_prog, n_forcing, residual;
  long long HW;
  ace::DevBuf in_kind, in_index, out_prog, prog_in_chan, in_mean, in_std, out_mean, out_std, out_clamp;
  int ocean_out = -1, ocean_interp = 0;
  bool slab = false;
  ace::SlabOceanIdx slab_ix;
  ace_corrector* corrector = nullptr;
  ace::DevBuf x, y;
  int wsB = 0;
};

using namespace ace;

// provided by corrector.cu
extern "C" int ace_corrector_apply(ace_corrector* c, const float* prev_prog_dev, const float* prev_forcing_dev, const float* next_dev,
                                   float* out_dev, float* next_prog_dev, int batch, void* stream);
extern "C" int ace_corrector_needs_next(ace_corrector* c);
extern "C" int ace_corrector_seed(ace_corrector* c, const float* prog_dev, int batch, void* stream);
extern "C" int ace_corrector_is_seeded(ace_corrector* c);

// provided by sfno.cu
extern "C" int ace_sfno_query(ace_sfno* net, int* in_chans, int* out_chans, long long* hw);

template <class T>
static void upload(DevBuf& d, const T* host, size_t n) {
  d.ensure(n * sizeof(T));
  ACE_CHECK_CUDA(cudaMemcpy(d.p, host, n * sizeof(T), cudaMemcpyHostToDevice));
}

static ace_stepper* make_stepper(int cin, int cout, long long hw, const ace_step_config* cfg);

extern "C" int ace_stepper_create(ace_sfno* net, const ace_step_config* cfg, ace_stepper** out) {
  ACE_API_BEGIN
  ACE_REQUIRE(net && cfg && out, "ace_stepper_create: null argument");
  int cin = 0, cout = 0;
  long long hw = 0;
  ACE_REQUIRE(ace_sfno_query(net, &cin, &cout, &hw) == ACE_OK, "ace_stepper_create: bad net");
  ace_stepper* st = make_stepper(cin, cout, hw, cfg);
  st->net = net;
  *out = st;
  ACE_API_END
}

extern "C" int ace_stepper_create_conditional(ace_csfno* net, const ace_step_config* cfg, ace_stepper** out) {
  ACE_API_BEGIN
  ACE_REQUIRE(net && cfg && out, "ace_stepper_create_conditional: null argument");
  int cin = 0, cout = 0;
  long long hw = 0;
  ACE_REQUIRE(ace_csfno_query(net, &cin, &cout, &hw) == ACE_OK, "ace_stepper_create_conditional: bad net");
  ace_stepper* st = make_stepper(cin, cout, hw, cfg);
  st->cnet = net;
  *out = st;
  ACE_API_END
}

extern "C" int ace_stepper_set_context(ace_stepper* st, const float* noise_dev, const float* pos_dev) {
  ACE_API_BEGIN
  ACE_REQUIRE(st && st->cnet, "ace_stepper_set_context: the stepper does not drive a noise-conditioned network");
  st->noise = noise_dev;
  st->pos = pos_dev;
  ACE_API_END
}

static ace_stepper* make_stepper(int cin, int cout, long long hw, const ace_step_config* cfg) {
  ACE_REQUIRE(cfg->n_in == cin && cfg->n_out == cout, "ace_stepper_create: net has %d->%d channels, step config %d->%d", cin,
              cout, cfg->n_in, cfg->n_out);
  ACE_REQUIRE(cfg->n_prog > 0 && cfg->n_forcing >= 0, "ace_stepper_create: bad state sizes");
  std::vector<int> prog_in_chan(cfg->n_prog, -1);
  for (int c = 0; c < cfg->n_in; ++c) {
    int kind = cfg->in_kind_host[c], idx = cfg->in_index_host[c];
    ACE_REQUIRE(kind == 0 || kind == 1, "in_kind[%d] must be 0 (prognostic) or 1 (forcing)", c);
    ACE_REQUIRE(idx >= 0 && idx < (kind == 0 ? cfg->n_prog : cfg->n_forcing), "in_index[%d]=%d out of range", c, idx);
    ACE_REQUIRE(cfg->in_std_host[c] != 0.f, "in_std[%d] is zero", c);
    if (kind == 0) prog_in_chan[idx] = c;
  }
  for (int c = 0; c < cfg->n_out; ++c)
    ACE_REQUIRE(cfg->out_prog_index_host[c] >= -1 && cfg->out_prog_index_host[c] < cfg->n_prog, "out_prog_index[%d] out of range", c);
  ace_stepper* st = new ace_stepper();
  try {
    st->n_in = cfg->n_in;
    st->n_out = cfg->n_out;
    st->n_prog = cfg->n_prog;
    st->n_forcing = cfg->n_forcing;
    st->residual = cfg->residual_prediction;
    st->HW = hw;
    upload(st->in_kind, cfg->in_kind_host, cfg->n_in);
    upload(st->in_index, cfg->in_index_host, cfg->n_in);
    upload(st->out_prog, cfg->out_prog_index_host, cfg->n_out);
    upload(st->prog_in_chan, prog_in_chan.data(), prog_in_chan.size());
    upload(st->in_mean, cfg->in_mean_host, cfg->n_in);
    upload(st->in_std, cfg->in_std_host, cfg->n_in);
    upload(st->out_mean, cfg->out_mean_host, cfg->n_out);
    upload(st->out_std, cfg->out_std_host, cfg->n_out);
    {
      std::vector<int> clamp(cfg->n_out, 0);
      if (cfg->out_force_positive_host)
        for (int c = 0; c < cfg->n_out; ++c) clamp[c] = cfg->out_force_positive_host[c] ? 1 : 0;
      upload(st->out_clamp, clamp.data(), clamp.size());
    }
    ACE_REQUIRE(cfg->ocean_out_index >= -1 && cfg->ocean_out_index < cfg->n_out, "ocean_out_index out of range");
    st->ocean_out = cfg->ocean_out_index;
    st->ocean_interp = cfg->ocean_interpolate ? 1 : 0;
  } catch (...) {
    delete st;
    throw;
  }
  return st;
}

extern "C" int ace_stepper_set_slab_ocean(ace_stepper* st, const ace_slab_ocean_config* cfg) {
  ACE_API_BEGIN
  ACE_REQUIRE(st, "ace_stepper_set_slab_ocean: null stepper");
  if (!cfg) {
    st->slab = false;
    return ACE_OK;
  }
  ACE_REQUIRE(st->ocean_out >= 0, "ace_stepper_set_slab_ocean: the step has no ocean surface temperature channel (ocean_out_index)");
  auto ok = [&](int c, int n) { return c >= 0 && c < n; };
  ACE_REQUIRE(ok(cfg->prog_sst, st->n_prog), "ace_stepper_set_slab_ocean: surface temperature must be a prognostic input");
  const int fl[6] = {cfg->out_dlw_sfc, cfg->out_ulw_sfc, cfg->out_dsw_sfc, cfg->out_usw_sfc, cfg->out_lhf, cfg->out_shf};
  for (int k = 0; k < 6; ++k) ACE_REQUIRE(ok(fl[k], st->n_out), "ace_stepper_set_slab_ocean: flux channel %d out of range", k);
  ACE_REQUIRE(cfg->timestep_seconds > 0, "ace_stepper_set_slab_ocean: timestep must be positive");
  st->slab_ix = {cfg->prog_sst, cfg->out_dlw_sfc, cfg->out_ulw_sfc, cfg->out_dsw_sfc, cfg->out_usw_sfc, cfg->out_lhf, cfg->out_shf,
                 (float)cfg->timestep_seconds};
  st->slab = true;
  ACE_API_END
}

extern "C" void ace_stepper_destroy(ace_stepper* st) { delete st; }

extern "C" int ace_stepper_set_corrector(ace_stepper* st, ace_corrector* c) {
  ACE_API_BEGIN
  ACE_REQUIRE(st, "ace_stepper_set_corrector: null stepper");
  st->corrector = c;
  ACE_API_END
}

extern "C" int ace_stepper_step(ace_stepper* st, const float* prog_dev, const float* forcing_dev, const float* ocean_dev,
                                const float* corrector_next_dev, float* out_dev, float* next_prog_dev, int batch, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(st && prog_dev && out_dev && batch > 0, "ace_stepper_step: bad argument");
  ACE_REQUIRE(forcing_dev || st->n_forcing == 0, "ace_stepper_step: forcing is null");
  ACE_REQUIRE(ocean_dev || st->ocean_out < 0, "ace_stepper_step: an ocean model is configured but ocean_dev is null");
  ACE_REQUIRE(corrector_next_dev || !ace_corrector_needs_next(st->corrector),
              "ace_stepper_step: the corrector's energy budget correction needs corrector_next_dev");
  cudaStream_t s = (cudaStream_t)stream;
  if (batch > st->wsB) {
    st->x.ensure((size_t)batch * st->n_in * st->HW * sizeof(float));
    st->y.ensure((size_t)batch * st->n_out * st->HW * sizeof(float));
    st->wsB = batch;
  }
  launch_pack_normalize(prog_dev, forcing_dev, st->n_prog, st->n_forcing, st->in_kind.as<int>(), st->in_index.as<int>(),
                        st->in_mean.as<float>(), st->in_std.as<float>(), batch, st->n_in, st->HW, st->x.as<float>(), s);
  int rc = st->cnet ? ace_csfno_forward(st->cnet, st->x.as<float>(), nullptr, nullptr, st->noise, st->pos, st->y.as<float>(), batch, stream)
                    : ace_sfno_forward(st->net, st->x.as<float>(), st->y.as<float>(), batch, stream);
  if (rc != ACE_OK) return rc;
  // reference order (single_module.py:670-709): ForcePositive -> conservation correctors -> ocean prescriber.  Without a
  // corrector the ocean overwrite rides in the denormalisation kernel; with one it is a separate pass over its one channel.
  const bool split = (st->corrector != nullptr || st->slab) && st->ocean_out >= 0;
  launch_unpack_denormalize(st->y.as<float>(), st->x.as<float>(), st->out_prog.as<int>(), st->prog_in_chan.as<int>(),
                            st->out_mean.as<float>(), st->out_std.as<float>(), st->residual, batch, st->n_out, st->n_in,
                            st->n_prog, st->HW, st->out_clamp.as<int>(), split ? -1 : st->ocean_out, st->ocean_interp, ocean_dev, out_dev,
                            next_prog_dev, s);
  if (st->corrector) {
    if (!ace_corrector_is_seeded(st->corrector)) {
      rc = ace_corrector_seed(st->corrector, prog_dev, batch, stream);  // first step of a rollout: the input IS the initial condition
      if (rc != ACE_OK) return rc;
    }
    rc = ace_corrector_apply(st->corrector, prog_dev, forcing_dev, corrector_next_dev, out_dev, next_prog_dev, batch, stream);
    if (rc != ACE_OK) return rc;
  }
  if (split && st->slab)
    launch_ocean_slab(out_dev, next_prog_dev, prog_dev, st->out_prog.as<int>(), batch, st->n_out, st->n_prog, st->HW, st->ocean_out,
                      st->ocean_interp, ocean_dev, st->slab_ix, s);
  else if (split)
    launch_ocean_prescribe(out_dev, next_prog_dev, st->out_prog.as<int>(), batch, st->n_out, st->n_prog, st->HW, st->ocean_out,
                           st->ocean_interp, ocean_dev, s);
  ACE_API_END
}
