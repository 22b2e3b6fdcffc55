/*
 * ace_b200 -- C ABI of the B200-native SFNO inference hot path.
 *
 * Drop-in boundary for the one path of ai2cm/ace (fme 2026.5.1) that this
 * library replaces.  The reference is pure Python, so there is no existing FFI
 * to mirror; each entry point below names the reference interface whose device
 * work it takes over (paths relative to the reference tree):
 *
 *   ace_sht_*      fme/sht_fix.py:60-151 (RealSHT), :153-226 (InverseRealSHT),
 *                  fme/fft.py:60-96 (rfft/irfft conventions)
 *   ace_sfno_*     fme/ace/models/modulus/sfnonet.py:341-749
 *                  (SphericalFourierNeuralOperatorNet.__init__/forward) incl.
 *                  sfnonet.py:123-252 (block), s2convolutions.py:47-197
 *                  (SpectralConvS2), contractions.py:170-195, layers.py:97-137,
 *                  as built by fme/ace/registry/sfno.py:44-61 and called from
 *                  fme/core/step/single_module.py:425-428
 *   ace_step_*     fme/core/step/single_module.py:595-665 (normalise, pack,
 *                  network, unpack, residual add, denormalise) and the feedback
 *                  loop of fme/ace/stepper/single_module.py:1124-1167
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a device pointer owned
 *     by the caller (PyTorch's allocator in practice) and only borrowed for the
 *     duration of the call; *_host pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it,
 *     nothing synchronises the host, so calls can be captured in a CUDA graph
 *     once workspaces exist (first call with a given batch allocates them).
 *   - return value 0 = success; anything else is an error code and
 *     ace_last_error() returns a thread-local message.  Nothing calls exit().
 *   - plans / nets are immutable after finalize; one in-flight call per object.
 *   - fp32 in, fp32 out.  Inside, GEMM-shaped work runs on tcgen05 tensor cores
 *     as 3-term split-bf16 products (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32
 *     accumulate in TMEM); see DESIGN.md for the error budget.
 */
#ifndef ACE_B200_H_
#define ACE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ACE_OK 0
#define ACE_ERR_INVALID 1   /* bad argument / unsupported configuration */
#define ACE_ERR_CUDA 2      /* a CUDA runtime / driver call failed */
#define ACE_ERR_STATE 3     /* call sequence error (e.g. forward before finalize) */

typedef struct ace_sht_plan ace_sht_plan;
typedef struct ace_sfno ace_sfno;
typedef struct ace_stepper ace_stepper;
typedef struct ace_corrector ace_corrector;
typedef struct ace_csfno ace_csfno;

int ace_version(void);
const char* ace_last_error(void);

/* Runtime options (development / A-B testing):
 *   "force_simt"  1 = route every GEMM through the generic SIMT CUDA kernel
 *                 instead of the tcgen05 kernel (both are CUDA; there is no CPU path)
 *   "split_terms" 3 (default) or 1 = plain bf16 products (fast, ~1e-2 accurate)
 *   "profile"     1 = time every launch with CUDA events (see ace_profile_report)
 *   "nvtx"        1 = one NVTX range per operator launch (default 0; env ACE_B200_NVTX)
 *   "umma_bn"     0 (default: per-op choice) or 128 / 192 / 256: N tile of the tcgen05 kernel (SHT stages)
 *   "umma_bk"     0 (default: per-op hint) or 32 / 64: K extent per pipeline stage of the forward SHT stages
 *   "conv_bn"     0 (default: per-op choice by wave quantisation) or 192 / 256: N tile of the 1x1-conv GEMMs
 *   "pair"        -1 (default: per-op choice) / 0 / 1: CTA-pair (tcgen05 cta_group::2) variants of the GEMM kernel
 *   "pdl"         1 = launch the tcgen05 / prep kernels with programmatic dependent launch (default 0)
 *   "dbg"         development switches of the tcgen05 kernel (non-zero values produce wrong results)
 * Read-only counters through ace_get_option: "count_umma" / "count_simt" = GEMMs launched on the
 * tcgen05 / SIMT kernel since load.                                                            */
int ace_set_option(const char* key, int value);
int ace_get_option(const char* key);
/* Number of kernels this library has launched since load (all streams). */
long long ace_launch_count(void);
/* With option "profile" = 1 every launch is bracketed by CUDA events on its stream.  This call
 * synchronises, writes one "name count total_ms" line per kernel name into buf, clears the
 * records and returns the number of bytes written. */
int ace_profile_report(char* buf, int buflen);
/* Scope hook: `cb(name, 1, user)` is called on the host before an operator's kernels are enqueued and `cb(name, 0, user)` after
 * (same names as ace_profile_report); NULL removes it.  A host records its own CUDA events / ranges there: ace_b200.timing
 * drives the reference's hierarchical Timer protocol with it (fme/core/benchmark/timer.py:48-51; the conditional SFNO block
 * threads `timer.child("filter")`, ... through its forward, fme/core/models/conditional_sfno/sfnonet.py:388-437).
 * With option "nvtx" = 1 (env ACE_B200_NVTX) every scope is also an NVTX range. */
typedef void (*ace_scope_callback)(const char* name, int begin, void* user);
int ace_set_scope_callback(ace_scope_callback cb, void* user);
/* Opens and closes the scope `name` without launching anything (tests of the hook / NVTX plumbing; no GPU needed). */
int ace_debug_scope(const char* name);

/* ---- spherical harmonic transforms ---------------------------------------------------
 * legendre_fwd_host / legendre_inv_host: float64 [mmax][lmax][nlat], the tables of
 * fme/sht_fix.py:113-117 (P_l^m(cos theta_k) * w_k) and :195-198 (P_l^m(cos theta_k)).
 * They are cast to fp32 exactly as the reference does before any further processing. */
int ace_sht_plan_create(int nlat, int nlon, int lmax, int mmax,
                        const double* legendre_fwd_host, const double* legendre_inv_host,
                        ace_sht_plan** out);
void ace_sht_plan_destroy(ace_sht_plan* plan);
/* x_dev: float32 [nfields][nlat][nlon]  ->  coeffs_dev: complex64 [nfields][lmax][mmax] */
int ace_sht_forward(ace_sht_plan* plan, const float* x_dev, float* coeffs_dev, long long nfields, void* stream);
/* coeffs_dev: complex64 [nfields][lmax][mmax] -> x_dev: float32 [nfields][nlat][nlon] */
int ace_sht_inverse(ace_sht_plan* plan, const float* coeffs_dev, float* x_dev, long long nfields, void* stream);

/* ---- the network ------------------------------------------------------------------ */
typedef struct ace_sfno_config {
  int img_h, img_w;          /* nlat, nlon */
  int in_chans, out_chans;
  int embed_dim, num_layers;
  int lmax, mmax;            /* modes_lat, modes_lon (sfnonet.py:471-472) */
  int mlp_hidden;            /* int(embed_dim * mlp_ratio); 0 = no MLP */
  int operator_type;         /* 0 = diagonal, 1 = dhconv */
  int normalization;         /* 0 = none, 1 = instance_norm */
  int pos_embed;             /* 0/1 */
  int big_skip;              /* 0/1 */
  float norm_eps;            /* 1e-6 in the reference */
} ace_sfno_config;

/* plan_outer: data-grid plan used by the first block's forward and the last block's
 * inverse transform (trans_down / itrans_up); plan_inner: legendre-gauss plan (trans / itrans).
 * They may be the same object.  Plans must outlive the net. */
int ace_sfno_create(const ace_sfno_config* cfg, ace_sht_plan* plan_outer, ace_sht_plan* plan_inner, ace_sfno** out);
void ace_sfno_destroy(ace_sfno* net);
/* name = the reference's state_dict key (e.g. "blocks.3.filter.filter.weight"); data_dev is
 * float32, contiguous, in the reference's shape.  Copied/re-laid-out on `stream`. */
int ace_sfno_set_param(ace_sfno* net, const char* name, const float* data_dev, long long numel, void* stream);
/* Verifies every parameter has been set. */
int ace_sfno_finalize(ace_sfno* net);
/* x_dev float32 [batch][in_chans][H][W] -> y_dev float32 [batch][out_chans][H][W] */
int ace_sfno_forward(ace_sfno* net, const float* x_dev, float* y_dev, int batch, void* stream);

/* Shape query (used by the stepper and the Python wrapper). */
int ace_sfno_query(ace_sfno* net, int* in_chans, int* out_chans, long long* hw);

/* ---- the noise-conditioned network -------------------------------
 * fme/core/models/conditional_sfno/sfnonet.py:496-824 (SphericalFourierNeuralOperatorNet with a Context) as built by
 * get_lat_lon_sfnonet (:443-493) for fme/ace/registry/stochastic_sfno.py:181-397 ("NoiseConditionedSFNO"): filter_type
 * "linear" (dhconv weights [1][L][O][I][2]), one filter group, scale_factor 1, encoder_layers 1, GELU, MLP, identity outer
 * skip, ConditionalLayerNorm (layers.py:143-320, channel LayerNorm per pixel, eps 1e-5) conditioned on any of: a scalar
 * embedding [batch][embed_dim_scalar], labels [batch][embed_dim_labels], a noise field [batch][embed_dim_noise][H][W], a
 * positional embedding [batch][embed_dim_pos][H][W] (any width: up to 64 channels one streaming kernel per norm or -- the
 * default from 32 padded channels on -- a statistics pass + one tcgen05 GEMM; beyond 64 only the latter). */
typedef struct ace_csfno_config {
  int img_h, img_w;
  int in_chans, out_chans;
  int embed_dim, num_layers;
  int lmax, mmax;
  int mlp_hidden;            /* int(embed_dim * mlp_ratio) */
  int pos_embed;             /* 0/1 learned additive position embedding after the encoder */
  int big_skip;              /* 0/1 */
  int normalize_big_skip;    /* 0/1: ConditionalLayerNorm on the big-skip copy of the input (sfnonet.py:738-747) */
  int affine_norms;          /* 0/1: elementwise affine of the channel LayerNorms (norm.weight / norm.bias) */
  int embed_dim_scalar, embed_dim_labels, embed_dim_noise, embed_dim_pos; /* ContextConfig (layers.py:33-43) */
  float norm_eps;            /* 1e-5 in the reference */
  int filter_residual;       /* 0/1: every block's residual and the big skip are SHT round trips (s2convolutions.py:195-199,
                                sfnonet.py:581-586,775-778) */
  int filter_output;         /* 0/1: the network output is passed through trans_down -> itrans_up (sfnonet.py:586-591,822) */
  int clip_latent_global_means; /* 0/1: eval branch of sfnonet.py:792-812 -- the per-channel spatial mean of the post-encoder latent
                                is clamped into the envelope buffers "_gm_min" / "_gm_max" (set like parameters, [embed_dim] each);
                                a no-op while any "_gm_max" entry is non-finite (envelope never trained) */
} ace_csfno_config;
/* plan_outer / plan_inner as for ace_sfno_create (trans_down + itrans_up on the data grid, trans + itrans on Legendre-Gauss). */
int ace_csfno_create(const ace_csfno_config* cfg, ace_sht_plan* plan_outer, ace_sht_plan* plan_inner, ace_csfno** out);
void ace_csfno_destroy(ace_csfno* net);
/* name = the reference's state_dict key (e.g. "blocks.0.norm0.W_scale_2d.weight", "norm_big_skip.norm.bias"). */
int ace_csfno_set_param(ace_csfno* net, const char* name, const float* data_dev, long long numel, void* stream);
/* Verifies every parameter has been set and builds the derived tables on `stream`. */
int ace_csfno_finalize(ace_csfno* net, void* stream);
/* x_dev float32 [batch][in_chans][H][W] + Context (pointers may be NULL when the corresponding width is 0)
 * -> y_dev float32 [batch][out_chans][H][W].  Enqueue-only, graph-capturable after the first call for a batch size. */
int ace_csfno_forward(ace_csfno* net, const float* x_dev, const float* scalar_dev, const float* labels_dev, const float* noise_dev,
                      const float* pos_dev, float* y_dev, int batch, void* stream);
int ace_csfno_query(ace_csfno* net, int* in_chans, int* out_chans, long long* hw);
/* Isotropic Gaussian noise fields of unit pointwise variance (fme/ace/registry/stochastic_sfno.py:21-47) from the caller's two
 * N(0,1) draws real_dev / imag_dev float32 [nfields][lmax][mmax]: Im(a_l0) = 0, Re / Im of m > 0 divided by sqrt(2), all scaled
 * by sqrt(4 pi) / lmax, then the inverse SHT of `plan` -> noise_dev float32 [nfields][nlat][nlon].
 * coeffs_scratch_dev: complex64 [nfields][lmax][mmax]. */
int ace_isotropic_noise(ace_sht_plan* plan, const float* real_dev, const float* imag_dev, float* coeffs_scratch_dev, float* noise_dev,
                        long long nfields, void* stream);
/* Label conditioning of NoiseConditionedModel (fme/ace/registry/stochastic_sfno.py:96-103,152-165).
 * ace_label_embed: out[batch][embed_dim] = labels[batch][n_labels] * weight[embed_dim][n_labels]^T + bias[embed_dim] (bias may be
 * NULL) -- the learned label embedding, torch.nn.Linear(n_labels, label_embed_dim).
 * ace_label_pos_embed: out[batch][phw] = pos[phw] + sum_l labels[batch][l] * label_pos[l][phw], phw = embed_dim_pos * H * W -- the
 * per-sample positional context "pos_embed + einsum('bl,lpxy->bpxy', labels, label_pos_embed)". */
int ace_label_embed(const float* labels_dev, const float* weight_dev, const float* bias_dev, int batch, int n_labels, int embed_dim,
                    float* out_dev, void* stream);
int ace_label_pos_embed(const float* pos_dev, const float* labels_dev, const float* label_pos_dev, int batch, int n_labels, long long phw,
                        float* out_dev, void* stream);

/* ---- fused step: normalise -> pack -> net -> (residual) -> denormalise -> feed back ----
 * State layout: prognostic/forcing/diagnostic fields as float32 [batch][n][H][W] tensors.
 *   in_index_*:  for network input channel c: source kind (0 = prognostic state, 1 = forcing)
 *                and index into that tensor
 *   out_prog_index: for output channel c, index of the prognostic it updates, or -1 (diagnostic) */
typedef struct ace_step_config {
  int n_in, n_out, n_prog, n_forcing;
  const int* in_kind_host;      /* [n_in] */
  const int* in_index_host;     /* [n_in] */
  const int* out_prog_index_host; /* [n_out] */
  const float* in_mean_host;    /* [n_in]  */
  const float* in_std_host;     /* [n_in]  */
  const float* out_mean_host;   /* [n_out] */
  const float* out_std_host;    /* [n_out] */
  int residual_prediction;      /* 0/1: add normalised input of the same prognostic to the output */
  /* post-step adjustments on the DENORMALISED outputs, in the reference's order (single_module.py:670-709):
   * corrector ForcePositive (fme/core/corrector/utils.py:26-43: clamp(min=0) of the named fields), then the ocean
   * prescriber (fme/core/ocean.py:165-215 + fme/core/prescriber.py:71-109): overwrite the surface temperature with the
   * target where round(ocean_fraction) == 1 (fme/core/spatial_masking.py:11-30), or blend linearly when `interpolate`. */
  const int* out_force_positive_host; /* [n_out] 0/1, or NULL */
  int ocean_out_index;          /* output channel of the surface temperature, or -1: no ocean model */
  int ocean_interpolate;        /* 0: replace on rounded mask == 1; 1: mask * target + (1 - mask) * generated */
} ace_step_config;
int ace_stepper_create(ace_sfno* net, const ace_step_config* cfg, ace_stepper** out);
void ace_stepper_destroy(ace_stepper* st);
/* One 6-hour step.  prog_dev [batch][n_prog][H][W] is read; out_dev [batch][n_out][H][W]
 * receives the denormalised (and adjusted) outputs; next_prog_dev [batch][n_prog][H][W] receives the state
 * for the next step (outputs that are prognostic; may alias nothing).  ocean_dev: float32 [batch][2][H][W] =
 * {ocean fraction, target surface temperature} at the OUTPUT time (the reference's next_step_input_data); required
 * iff ocean_out_index >= 0.  corrector_next_dev: float32 [batch][2][H][W] = {DSWRFtoa, HGTsfc} at the OUTPUT time; required
 * iff the attached corrector has its energy budget correction on (ace_corrector_needs_next). */
/* The same fused step around a noise-conditioned network (scalar / label contexts are not supported here).  Its context fields are
 * read from persistent device buffers the caller refills before every step: noise_dev float32 [batch][embed_dim_noise][H][W] (a
 * fresh draw per step, fme/ace/registry/stochastic_sfno.py:128-146) and pos_dev [batch][embed_dim_pos][H][W] (or NULL). */
int ace_stepper_create_conditional(ace_csfno* net, const ace_step_config* cfg, ace_stepper** out);
int ace_stepper_set_context(ace_stepper* st, const float* noise_dev, const float* pos_dev);
int ace_stepper_step(ace_stepper* st, const float* prog_dev, const float* forcing_dev, const float* ocean_dev,
                     const float* corrector_next_dev, float* out_dev, float* next_prog_dev, int batch, void* stream);

/* Slab ocean instead of a prescribed target (fme/core/ocean.py:64-88,223-243; configs/baselines/shield-som): the surface
 * temperature written where the mask selects ocean is T_in + (F_net + Q) / (rho depth c_p) dt, F_net = net surface energy flux of
 * the generated fields (after the correctors) without frozen precipitation (fme/core/metrics.py:299-334).  ocean_dev of
 * ace_stepper_step is then float32 [batch][3][H][W] = {ocean fraction, q_flux, mixed layer depth} at the OUTPUT time.  NULL: back
 * to the prescribed target. */
typedef struct ace_slab_ocean_config {
  int prog_sst;                      /* surface temperature in the prognostic INPUT state */
  int out_dlw_sfc, out_ulw_sfc, out_dsw_sfc, out_usw_sfc, out_lhf, out_shf; /* DLWRFsfc ULWRFsfc DSWRFsfc USWRFsfc LHTFLsfc SHTFLsfc */
  double timestep_seconds;
} ace_slab_ocean_config;
int ace_stepper_set_slab_ocean(ace_stepper* st, const ace_slab_ocean_config* cfg);

/* ---- conservation correctors of the post-step state -------------
 * fme/core/corrector/atmosphere.py:404-463 (conserve_dry_air: pin the area-weighted global mean of ps - g * total water path
 * to its value at the initial condition by a globally constant dry-air pressure offset, solving for ps) and :518-608
 * (moisture_budget_correction: scale precipitation / evaporation so the global budget closes, then optionally recompute the
 * advective tendency as the column residual).  Channel indices refer to the denormalised output tensor [batch][n_out][hw] and
 * to the prognostic INPUT state [batch][n_prog][hw] of the step; out_prog_index maps an output channel to the prognostic it
 * feeds (-1: diagnostic), so corrected fields also reach the next state. */
typedef struct ace_corrector_config {
  int n_out, n_prog, nz;             /* nz = number of vertical layers; ak / bk have nz + 1 interface values */
  long long hw;
  const float* area_weights_host;    /* [hw] */
  const double* ak_host;             /* [nz + 1] hybrid sigma-pressure coefficients (fme/core/coordinates.py:241-255) */
  const double* bk_host;
  const int* out_prog_index_host;    /* [n_out] */
  int out_ps;                        /* surface pressure (PRESsfc) */
  const int* out_wat_host;           /* [nz] specific_total_water_k */
  int out_precip, out_lhf, out_adv;  /* PRATEsfc, LHTFLsfc, tendency_of_total_water_path_due_to_advection (-1 if unused) */
  int prog_ps;
  const int* prog_wat_host;          /* [nz] */
  int conserve_dry_air;              /* 0/1 */
  int moisture_mode;                 /* 0 none, 1 precipitation, 2 advection_and_precipitation, 3 evaporation, 4 advection_and_evaporation */
  double timestep_seconds;
  /* the remaining corrections of the reference's sequence (atmosphere.py:349-398); zero / -1 = off */
  int zero_global_mean_moisture_advection; /* 0/1: subtract the global mean of out_adv (:467-490), before the moisture budget */
  int out_frozen;                    /* total_frozen_precipitation_rate, or -1 (enters the surface energy flux) */
  int clip_frozen_precipitation;     /* 0/1: frozen = min(frozen, corrected precipitation) (:493-515); needs moisture_mode != 0 */
  int energy_mode;                   /* 0 none, 1 constant_temperature (:611-695) */
  double unaccounted_heating;        /* W/m2 (constant_unaccounted_heating) */
  const int* out_temp_host;          /* [nz] air_temperature_k in out (energy_mode only) */
  const int* prog_temp_host;         /* [nz] in the prognostic state */
  int n_forcing, forcing_hgt;        /* the step's forcing input [batch][n_forcing][hw] and its surface height channel (HGTsfc) */
  int out_dlw_sfc, out_ulw_sfc, out_dsw_sfc, out_usw_sfc, out_shf, out_usw_toa, out_ulw_toa; /* DLWRFsfc ULWRFsfc DSWRFsfc USWRFsfc SHTFLsfc USWRFtoa ULWRFtoa */
} ace_corrector_config;
int ace_corrector_create(const ace_corrector_config* cfg, ace_corrector** out);
void ace_corrector_destroy(ace_corrector* c);
/* Capture the dry-air reference from the initial condition (the reference does this the first time the corrector runs,
 * atmosphere.py:404-427); synchronises the stream once.  ace_corrector_reset forgets it (new rollout). */
int ace_corrector_seed(ace_corrector* c, const float* prog_dev, int batch, void* stream);
int ace_corrector_reset(ace_corrector* c);
int ace_corrector_is_seeded(ace_corrector* c);
/* CorrectorState.global_dry_air_mass (fme/core/corrector/state.py:15-29): the reference carries the dry-air target from one
 * prediction window to the next inside stepper_state (fme/ace/stepper/single_module.py:1160-1165).  get: copy the fp64 target of
 * each sample to host memory (ACE_ERR_STATE when unseeded); set: install it and mark the corrector seeded, so that the next
 * step does NOT re-seed from its input.  Both synchronise `stream` once. */
int ace_corrector_get_state(ace_corrector* c, double* target_host, int batch, void* stream);
int ace_corrector_set_state(ace_corrector* c, const double* target_host, int batch, void* stream);
/* 1 when ace_corrector_apply needs prev_forcing_dev and next_dev (energy budget correction). */
int ace_corrector_needs_next(ace_corrector* c);
/* In place on out_dev (and next_prog_dev for corrected prognostic fields); prev_prog_dev / prev_forcing_dev = the step's input
 * state and forcing; next_dev [batch][2][hw] = (TOA downward shortwave DSWRFtoa, surface height HGTsfc) valid at the OUTPUT time
 * (the reference reads them from next_step_input_data, atmosphere.py:626-633).  The last two may be NULL unless energy_mode. */
int ace_corrector_apply(ace_corrector* c, const float* prev_prog_dev, const float* prev_forcing_dev, const float* next_dev,
                        float* out_dev, float* next_prog_dev, int batch, void* stream);
/* Attach (or detach with NULL) a corrector to the fused step: it then runs after ForcePositive and before the ocean
 * prescriber, the reference's order (fme/core/step/single_module.py:670-709). */
int ace_stepper_set_corrector(ace_stepper* st, ace_corrector* c);

/* ---- HEALPix spherical harmonic transform ----------------------
 * fme/core/cuhpx/sht.py:32-98 (SHT) and :101-153 (iSHT) with fme/core/cuhpx/tools.py:34-83 (per-ring rfft / irfft +
 * phase shift).  `plan` is an ace_sht_plan created with nlat = 4*nside - 1 (the iso-latitude rings), nlon = 4*nside and
 * the HEALPix Legendre tables [mmax][lmax][4*nside - 1] (ring quadrature weights folded into the forward table, no
 * Condon-Shortley sign: tools.py:288-336).  Pixels in RING order.
 * x_dev: float32 [nfields][12*nside^2]  <->  coeffs_dev: complex64 [nfields][lmax][mmax] */
int ace_hpx_forward(ace_sht_plan* plan, int nside, const float* x_dev, float* coeffs_dev, long long nfields, void* stream);
int ace_hpx_inverse(ace_sht_plan* plan, int nside, const float* coeffs_dev, float* x_dev, long long nfields, void* stream);

/* ---- device reductions of the inference aggregators ----------
 * fme/core/metrics.py:35-197 (weighted_sum / weighted_mean / weighted_std / weighted_mean_bias /
 * root_mean_squared_error) as used by fme/core/gridded_ops.py:284-360 (LatLonOperations): one pass,
 * fp64 accumulation.  x_dev, t_dev: float32 [nfields][hw] (t_dev may be NULL), weights_dev: float32 [hw];
 * out_dev: float64 [nfields][5] = {sum w x, sum w x^2, sum w (x-t), sum w (x-t)^2, sum w}; points with
 * zero weight contribute nothing even if NaN (metrics.py:56-58). */
int ace_weighted_moments(const float* x_dev, const float* t_dev, const float* weights_dev, long long nfields,
                         long long hw, double* out_dev, void* stream);
/* Time-mean maps (fme/ace/aggregator/inference/time_mean.py:103-124 `_add_or_initialize_time_mean`): one pass over a
 * window adds its sum over two leading axes (sample, time) to a running fp32 map:
 *   acc_dev[i] += sum_{a < n_outer, b < n_inner} x_dev[a * stride_outer + b * stride_inner + i],   i < n_elems
 * (strides in elements; for a window tensor [sample][time][field][H][W]: n_elems = fields * H * W). */
int ace_time_sum(const float* x_dev, int n_outer, long long stride_outer, int n_inner, long long stride_inner,
                 long long n_elems, float* acc_dev, void* stream);
/* zonal mean (mean over longitude; gridded_ops.py:313-315): x_dev float32 [nfields][h][w] -> out_dev [nfields][h] */
int ace_zonal_mean(const float* x_dev, long long nfields, int h, int w, float* out_dev, void* stream);
/* fme/core/metrics.py:388-408 spherical_power_spectrum: complex64 [nfields][lmax][mmax] -> float32 [nfields][lmax] */
int ace_power_spectrum(const float* coeffs_dev, long long nfields, int lmax, int mmax, float* out_dev, void* stream);

/* ---- development hook: run one split-bf16 GEMM through both kernels (tests only) ------
 * D[z][m][n] = sum_k A[z][m][k] * B[z][n][k], fp32 in/out.  layout bit 0: A is stored MN-major
 * ([z][k][m]); bit 1: B is stored MN-major ([z][k][n]).  impl: 0 = SIMT kernel, 1 = tcgen05
 * kernel (error if the shape is not eligible, e.g. n % 4 != 0 or unaligned strides). */
int ace_dev_gemm(const float* a_dev, const float* b_dev, float* d_dev, int m, int n, int k, int nbatch,
                 int layout, int impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACE_B200_H_ */
