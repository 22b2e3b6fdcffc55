"""Fused single-module step + autoregressive rollout around the B200 SFNO.

Device work mirrored (corrector / ocean / masking disabled, as in the throughput benchmark):

* ``/root/reference/fme/core/step/single_module.py:595-665`` ``step_with_adjustments``:
  ``normalizer.normalize`` -> ``in_packer.pack`` -> module -> ``out_packer.unpack`` ->
  [``residual_prediction``: add normalised prognostic inputs] -> ``normalizer.denormalize``
  (``fme/core/normalizer.py:213-243``: ``(x - mean) / std`` and ``x * std + mean`` per name,
  ``fme/core/packer.py:45-52``: channel order = order of the name lists);
* ``/root/reference/fme/ace/stepper/single_module.py:1124-1167`` ``predict_generator``: the state
  fed to step t+1 is the prognostic subset of step t's output; forcing comes from the data.

The reference spends ~330 small launches per step on this; here it is two streaming kernels
around ``ace_sfno_forward`` (``ace_stepper_step`` in the C ABI), and ``rollout`` replays the
whole step from a CUDA graph.
"""
import ctypes
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .sfno import SphericalFourierNeuralOperatorNet


class PrognosticDict(dict):
    """name -> [B, 1, H, W] of the prognostic variables plus the terminal ``stepper_state`` of the window that produced it
    (the reference's ``PrognosticState`` carries it the same way, fme/ace/stepper/single_module.py:507-521)."""

    stepper_state = None


class FusedStepper:
    def __init__(
        self,
        module: SphericalFourierNeuralOperatorNet,
        in_names: Sequence[str],
        out_names: Sequence[str],
        means: Mapping[str, float],
        stds: Mapping[str, float],
        residual_prediction: bool = False,
        force_positive_names: Sequence[str] = (),
        ocean: Optional[Mapping[str, object]] = None,
        corrector: Optional[Mapping[str, object]] = None,
        next_step_forcing_names: Sequence[str] = (),
        prescribed_prognostic_names: Sequence[str] = (),
    ):
        """``next_step_forcing_names`` / ``prescribed_prognostic_names``: ``SingleModuleStepConfig``'s fields of the same names
        (``fme/core/step/single_module.py:64-66,92-93,103-117``): input-only variables read from the OUTPUT time by ``predict*``,
        and outputs overwritten from the next-step data after the ocean step (``:709-714``).
        ``force_positive_names``: outputs clamped to >= 0 after denormalisation (the corrector's ForcePositive,
        ``fme/core/corrector/utils.py:26-43``).  ``ocean``: ``{"surface_temperature_name": ..., "interpolate": False}`` enables
        the prescribed-SST ocean (``fme/core/ocean.py:165-215``): every step then takes ``ocean`` data ``[B, 2, H, W]`` =
        (ocean fraction, target surface temperature) valid at the OUTPUT time.  With ``"slab": {"mixed_layer_depth_name": ...,
        "q_flux_name": ..., "timestep_seconds": 21600.0}`` the target is the slab-ocean prediction instead (``fme/core/ocean.py:64-88``:
        input surface temperature + (net surface energy flux of the corrected outputs + q-flux) / (rho depth c_p) dt) and the ocean
        data is ``[B, 3, H, W]`` = (ocean fraction, q-flux, mixed layer depth).
        ``corrector``: the conservation correctors of ``fme/core/corrector/atmosphere.py`` (``conserve_dry_air``,
        ``moisture_budget_correction``), run between ForcePositive and the ocean like the reference:
        ``dict(conserve_dry_air=True, moisture_budget_correction="advection_and_precipitation", ak=..., bk=...,
        area_weights=[H, W], timestep_seconds=21600.0)``; field names are resolved with the reference's prefixes
        (``fme/core/atmosphere_data.py:17-41``: ``PRESsfc``, ``specific_total_water_k``, ``PRATEsfc``, ``LHTFLsfc``,
        ``tendency_of_total_water_path_due_to_advection``).  The dry-air reference is captured from the first state the
        stepper sees after ``reset_corrector_state()`` (the initial condition of a rollout)."""
        self.module = module
        # a NoiseConditionedModel wrapper (ace_b200/csfno.py): the step drives its conditional network and draws fresh noise per step
        self._wrapper = module if hasattr(module, "conditional_model") else None
        if self._wrapper is not None:
            module = self.module = self._wrapper.conditional_model
            cc = module.context_config
            if cc.embed_dim_scalar or cc.embed_dim_labels:
                raise NotImplementedError("FusedStepper: scalar / label conditioning is not supported in the fused step")
        self._ctx = None
        self.in_names: List[str] = list(in_names)
        self.out_names: List[str] = list(out_names)
        if len(self.in_names) != module.in_chans or len(self.out_names) != module.out_chans:
            raise ValueError(
                f"module has {module.in_chans}->{module.out_chans} channels, names give {len(self.in_names)}->{len(self.out_names)}"
            )
        # fme/core/step/single_module.py: prognostic = outputs that are also inputs; forcing = inputs only
        self.prognostic_names = [n for n in self.out_names if n in self.in_names]
        self.forcing_names = [n for n in self.in_names if n not in self.out_names]
        self.diagnostic_names = [n for n in self.out_names if n not in self.in_names]
        self.residual_prediction = bool(residual_prediction)
        self.force_positive_names = list(force_positive_names)
        for n in self.force_positive_names:
            if n not in self.out_names:
                raise ValueError(f"force_positive name '{n}' is not an output")
        self.ocean = dict(ocean) if ocean is not None else None
        if self.ocean is not None and self.ocean["surface_temperature_name"] not in self.out_names:
            raise ValueError("ocean surface_temperature_name must be an output")
        self._slab = dict(self.ocean["slab"]) if self.ocean is not None and self.ocean.get("slab") else None
        self.n_ocean = 0 if self.ocean is None else (3 if self._slab else 2)
        if self._slab is not None and self.ocean["surface_temperature_name"] not in self.in_names:
            raise ValueError("slab ocean: the surface temperature must be a prognostic variable (input and output)")
        self._means = {k: float(v) for k, v in means.items()}
        self._stds = {k: float(v) for k, v in stds.items()}
        for n in set(self.in_names) | set(self.out_names):
            if n not in self._means or n not in self._stds:
                raise KeyError(f"normalization statistics missing for '{n}'")
        self.corrector = dict(corrector) if corrector is not None else None
        self.next_step_forcing_names = list(next_step_forcing_names)
        for n in self.next_step_forcing_names:
            if n not in self.in_names:
                raise ValueError(f"next_step_forcing_name '{n}' not in in_names: {self.in_names}")
            if n in self.out_names:
                raise ValueError(f"next_step_forcing_name is an output variable: '{n}'")
        self.prescribed_prognostic_names = list(prescribed_prognostic_names)
        for n in self.prescribed_prognostic_names:
            if n not in self.out_names:
                raise ValueError(f"prescribed_prognostic_name '{n}' must be in out_names: {self.out_names}")
        self._corrector_handle = None
        self._handle = None
        self._handle_net = None
        self._graph = None
        self._static = None

    # ------------------------------------------------------------------ native object
    def _ensure(self, device):
        net = self.module.native_handle()
        if self._handle is not None and self._handle_net is not None and net is not None and self._handle_net.value == net.value:
            return
        self._destroy()
        if net is None:
            # materialise the device net (uploads parameters) with a throw-away forward
            with torch.no_grad():
                (self._wrapper or self.module)(torch.zeros(1, self.module.in_chans, *self.module.img_shape, device=device))
            net = self.module.native_handle()
        kind = np.array([0 if n in self.prognostic_names else 1 for n in self.in_names], dtype=np.int32)
        index = np.array(
            [self.prognostic_names.index(n) if n in self.prognostic_names else self.forcing_names.index(n) for n in self.in_names],
            dtype=np.int32,
        )
        out_prog = np.array([self.prognostic_names.index(n) if n in self.prognostic_names else -1 for n in self.out_names], dtype=np.int32)
        clamp = np.array([1 if n in self.force_positive_names else 0 for n in self.out_names], dtype=np.int32)
        f32 = lambda names, d: np.array([d[n] for n in names], dtype=np.float32)  # noqa: E731
        in_mean, in_std = f32(self.in_names, self._means), f32(self.in_names, self._stds)
        out_mean, out_std = f32(self.out_names, self._means), f32(self.out_names, self._stds)
        ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))  # noqa: E731
        fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))  # noqa: E731
        cfg = _lib.StepConfig(
            n_in=len(self.in_names), n_out=len(self.out_names), n_prog=len(self.prognostic_names),
            n_forcing=len(self.forcing_names), in_kind_host=ip(kind), in_index_host=ip(index),
            out_prog_index_host=ip(out_prog), in_mean_host=fp(in_mean), in_std_host=fp(in_std),
            out_mean_host=fp(out_mean), out_std_host=fp(out_std), residual_prediction=int(self.residual_prediction),
            out_force_positive_host=ip(clamp),
            ocean_out_index=self.out_names.index(self.ocean["surface_temperature_name"]) if self.ocean is not None else -1,
            ocean_interpolate=int(bool(self.ocean.get("interpolate", False))) if self.ocean is not None else 0,
        )
        handle = ctypes.c_void_p()
        create = _lib.load().ace_stepper_create_conditional if self._wrapper is not None else _lib.load().ace_stepper_create
        _lib.check(create(net, ctypes.byref(cfg), ctypes.byref(handle)))
        self._handle, self._handle_net = handle, net
        self._graph = None
        self._ctx = None
        if self._slab is not None:
            from .corrector import _find  # the reference's field-name prefixes (fme/core/atmosphere_data.py:17-41)

            keys = dict(out_dlw_sfc="sfc_down_lw_radiative_flux", out_ulw_sfc="sfc_up_lw_radiative_flux", out_dsw_sfc="sfc_down_sw_radiative_flux",
                        out_usw_sfc="sfc_up_sw_radiative_flux", out_lhf="latent_heat_flux", out_shf="sensible_heat_flux")
            idx = {k: _find(self.out_names, v) for k, v in keys.items()}
            missing = [keys[k] for k, i in idx.items() if i < 0]
            if missing:
                raise ValueError(f"slab ocean: no output field for {missing}")
            scfg = _lib.SlabOceanConfig(prog_sst=self.prognostic_names.index(self.ocean["surface_temperature_name"]),
                                        timestep_seconds=float(self._slab.get("timestep_seconds", 21600.0)), **idx)
            _lib.check(_lib.load().ace_stepper_set_slab_ocean(self._handle, ctypes.byref(scfg)))
        if self.corrector is not None:
            self._build_corrector()
            _lib.check(_lib.load().ace_stepper_set_corrector(self._handle, self._corrector_handle))

    def _build_corrector(self):
        from .corrector import AtmosphereCorrector

        self._corrector_obj = AtmosphereCorrector(self.out_names, self.prognostic_names, self.module.img_shape,
                                                  forcing_names=self.forcing_names, **self.corrector)
        self._corrector_handle = self._corrector_obj._handle

    @property
    def corrector_needs_next(self) -> bool:
        """True when every step also needs (DSWRFtoa, HGTsfc) at the output time (total-energy budget correction)."""
        c = self.corrector or {}
        return c.get("total_energy_budget_correction") is not None

    def corrector_next_from_forcing(self, forcing_next: torch.Tensor) -> torch.Tensor:
        """Select the two next-step fields the energy correction reads from a forcing tensor [..., n_forcing, H, W]."""
        from .corrector import next_step_names

        a, b = next_step_names(self.forcing_names)
        return forcing_next[..., [self.forcing_names.index(a), self.forcing_names.index(b)], :, :].contiguous()

    def _refresh_context(self, B: int, device, noise: Optional[torch.Tensor]):
        """Noise-conditioned network: (re)fill the persistent context buffers the native step reads -- a fresh noise draw per step
        (``fme/ace/registry/stochastic_sfno.py:128-146``), or ``noise`` when the caller injects it (parity tests)."""
        w = self._wrapper
        H, Wd = self.module.img_shape
        pkey = None if w.pos_embed is None else (w.pos_embed.data_ptr(), w.pos_embed._version)
        if self._ctx is not None and self._ctx["B"] == B and self._ctx["device"] == device:
            self._refresh_pos()
        if self._ctx is None or self._ctx["B"] != B or self._ctx["device"] != device:
            nz = torch.empty(B, w.embed_dim, H, Wd, device=device) if w.embed_dim > 0 else None
            pos = w.pos_embed.detach().float().expand(B, -1, -1, -1).contiguous() if w.pos_embed is not None else None
            self._ctx = dict(B=B, device=device, noise=nz, pos=pos, pkey=pkey)
            _lib.check(_lib.load().ace_stepper_set_context(
                self._handle, ctypes.c_void_p(nz.data_ptr()) if nz is not None else None, ctypes.c_void_p(pos.data_ptr()) if pos is not None else None))
        if self._ctx["noise"] is not None:
            self._ctx["noise"].copy_(noise if noise is not None else w.draw_noise(B, device))

    def _refresh_pos(self):
        """pos_embed edited / reloaded since it was cached: refresh the copy IN PLACE (a captured graph reads this buffer)."""
        w, c = self._wrapper, self._ctx
        if w is None or c is None or w.pos_embed is None:
            return
        pkey = (w.pos_embed.data_ptr(), w.pos_embed._version)
        if c["pkey"] != pkey:
            c["pos"].copy_(w.pos_embed.detach().float().expand(c["B"], -1, -1, -1))
            c["pkey"] = pkey

    def reset_corrector_state(self):
        """Forget the dry-air reference: the next step's input state is treated as the initial condition of a new rollout."""
        if self._corrector_handle is not None:
            _lib.check(_lib.load().ace_corrector_reset(self._corrector_handle))

    def get_stepper_state(self, batch: int):
        """The per-sample state the reference threads from one ``predict`` window to the next (``StepperState.corrector_state``,
        fme/core/stepper_state.py; ``CorrectorState.global_dry_air_mass``, fme/core/corrector/state.py:15-29) as a plain dict
        ``{"corrector_state": {"global_dry_air_mass": fp64 [batch, 1, 1]}}``, or None when there is no corrector / nothing seeded."""
        if self._corrector_handle is None or not _lib.load().ace_corrector_is_seeded(self._corrector_handle):
            return None
        return {"corrector_state": {"global_dry_air_mass": self._corrector_obj.get_state(batch)}}

    def _begin_window(self, state: torch.Tensor, stepper_state):
        """Start of a rollout / prediction window (fme/core/corrector/atmosphere.py:404-427): with ``stepper_state`` carrying a
        dry-air target the corrector keeps it (the mass stays pinned to the very first initial condition, as the reference's
        ``predict -> prognostic_state -> next initial condition`` chain does); otherwise the target is seeded from ``state``."""
        if self.corrector is None:
            return
        with torch.cuda.device(state.device):
            self._ensure(state.device)
            gm = None
            if stepper_state is not None:
                cs = stepper_state.get("corrector_state") if isinstance(stepper_state, Mapping) else getattr(stepper_state, "corrector_state", None)
                if cs is not None:
                    gm = cs.get("global_dry_air_mass") if isinstance(cs, Mapping) else getattr(cs, "global_dry_air_mass", None)
            if gm is not None:
                if gm.numel() != state.shape[0]:
                    raise ValueError(f"stepper_state carries {gm.numel()} samples, the state has {state.shape[0]}")
                self._corrector_obj.set_state(gm)
            else:
                self.reset_corrector_state()
                self._seed_corrector(state)

    def _seed_corrector(self, prog: torch.Tensor):
        if self._corrector_handle is not None and not _lib.load().ace_corrector_is_seeded(self._corrector_handle):
            _lib.check(_lib.load().ace_corrector_seed(self._corrector_handle, ctypes.c_void_p(prog.data_ptr()), prog.shape[0],
                                                      _lib.current_stream_ptr()))

    def _destroy(self):
        self._corrector_handle = None  # the stepper goes first (it borrows the corrector), AtmosphereCorrector frees itself
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().ace_stepper_destroy(self._handle)
            except Exception:  # noqa: BLE001
                pass
        self._handle = None
        self._handle_net = None

    def __del__(self):
        self._destroy()

    # ------------------------------------------------------------------ one step, packed tensors
    def step_packed(self, prog: torch.Tensor, forcing: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                    next_prog: Optional[torch.Tensor] = None, ocean: Optional[torch.Tensor] = None,
                    corrector_next: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                    prescribed: Optional[torch.Tensor] = None):
        """prog [B, n_prog, H, W], forcing [B, n_forcing, H, W] (, ocean [B, 2, H, W]) (, corrector_next [B, 2, H, W] =
        (DSWRFtoa, HGTsfc) at the output time) (, prescribed [B, n_prescribed, H, W] = the ``prescribed_prognostic_names`` at the
        output time) -> (out [B, n_out, H, W], next_prog)."""
        B = prog.shape[0]
        H, W = self.module.img_shape
        prog = prog.float().contiguous()
        if forcing is not None:
            forcing = forcing.float().contiguous()
        if (self.ocean is not None) != (ocean is not None):
            raise ValueError("ocean data must be given exactly when an ocean model is configured")
        if ocean is not None:
            ocean = ocean.float().contiguous()
            if tuple(ocean.shape) != (B, self.n_ocean, H, W):
                raise ValueError(f"ocean data must be [B, {self.n_ocean}, H, W] = (ocean fraction, "
                                 f"{'q-flux, mixed layer depth' if self._slab else 'target surface temperature'}), got {tuple(ocean.shape)}")
        if self.corrector_needs_next != (corrector_next is not None):
            raise ValueError("corrector_next = (DSWRFtoa, HGTsfc) at the output time must be given exactly when the energy budget "
                             "correction is configured")
        if corrector_next is not None:
            corrector_next = corrector_next.float().contiguous()
            if tuple(corrector_next.shape) != (B, 2, H, W):
                raise ValueError(f"corrector_next must be [B, 2, H, W], got {tuple(corrector_next.shape)}")
        if bool(self.prescribed_prognostic_names) != (prescribed is not None):
            raise ValueError("prescribed data must be given exactly when prescribed_prognostic_names is configured")
        if prescribed is not None and tuple(prescribed.shape) != (B, len(self.prescribed_prognostic_names), H, W):
            raise ValueError(f"prescribed must be [B, {len(self.prescribed_prognostic_names)}, H, W], got {tuple(prescribed.shape)}")
        if prescribed is not None:
            prescribed = prescribed.to(device=prog.device, dtype=torch.float32)
        if out is None:
            out = torch.empty(B, len(self.out_names), H, W, device=prog.device, dtype=torch.float32)
        if next_prog is None:
            next_prog = torch.empty(B, len(self.prognostic_names), H, W, device=prog.device, dtype=torch.float32)
        # the library receives raw pointers: a view, a half / double buffer or another device would be silent corruption
        for nm, t, c in (("prog", prog, len(self.prognostic_names)), ("forcing", forcing, len(self.forcing_names)),
                         ("out", out, len(self.out_names)), ("next_prog", next_prog, len(self.prognostic_names)),
                         ("ocean", ocean, self.n_ocean), ("corrector_next", corrector_next, 2)):
            if t is None:
                continue
            if t.device != prog.device or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != (B, c, H, W):
                raise ValueError(f"FusedStepper.step_packed: `{nm}` must be a contiguous fp32 tensor [{B}, {c}, {H}, {W}] on {prog.device}, "
                                 f"got {t.dtype} {tuple(t.shape)} on {t.device}{'' if t.is_contiguous() else ' (non-contiguous)'}")
        self._native_step(prog, forcing, ocean, corrector_next, noise, out, next_prog)
        # prescribed overwrite, last of all (fme/core/step/single_module.py:709-714): device-to-device copies on the same stream
        for j, n in enumerate(self.prescribed_prognostic_names):
            out[:, self.out_names.index(n)].copy_(prescribed[:, j])
            if n in self.prognostic_names:
                next_prog[:, self.prognostic_names.index(n)].copy_(prescribed[:, j])
        return out, next_prog

    def _native_step(self, prog, forcing, ocean, corrector_next, noise, out, next_prog):
        """``ace_stepper_step`` on validated, contiguous fp32 device tensors (the one place the fused step enters the library)."""
        if not prog.is_cuda:
            raise _lib.AceError("FusedStepper: tensors must be on a CUDA device (there is no CPU path)")
        B = prog.shape[0]
        with torch.cuda.device(prog.device):
            self._ensure(prog.device)
            stream = _lib.current_stream_ptr()
            self.module._sync_params(stream)
            if self._wrapper is not None:
                self._refresh_context(B, prog.device, noise)
            _lib.check(_lib.load().ace_stepper_step(
                self._handle, ctypes.c_void_p(prog.data_ptr()),
                ctypes.c_void_p(forcing.data_ptr()) if forcing is not None else None,
                ctypes.c_void_p(ocean.data_ptr()) if ocean is not None else None,
                ctypes.c_void_p(corrector_next.data_ptr()) if corrector_next is not None else None,
                ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(next_prog.data_ptr()), B, stream))

    def _sync_before_replay(self, device):
        """A graph replay never passes through ``_native_step``: push edited / reloaded parameters (``load_state_dict``,
        in-place updates; version-counter tracked by the module) to the device buffers the captured kernels read, OUTSIDE
        capture.  The uploads reuse the same device buffers, so the captured graph stays valid; a rebuilt native net
        (``_ensure`` drops the graph) is re-captured by the caller."""
        with torch.cuda.device(device):
            self._ensure(device)
            self.module._sync_params(_lib.current_stream_ptr())
            self._refresh_pos()

    # ------------------------------------------------------------------ one step, name dicts (reference API shape)
    def step(self, input: Mapping[str, torch.Tensor], next_step_input_data: Optional[Mapping[str, torch.Tensor]] = None,
             stepper_state=None) -> Dict[str, torch.Tensor]:
        """``input``: every name in ``in_names`` -> [B, H, W] (denormalised).  Returns every ``out_names`` entry.
        Like the reference's ``step`` (fme/core/step/single_module.py:670-690) the corrector's dry-air target comes from
        ``stepper_state`` when given and is seeded from THIS call's input otherwise; the updated state is
        ``get_stepper_state(B)`` afterwards."""
        prog = torch.stack([input[n] for n in self.prognostic_names], dim=1)
        self._begin_window(prog.float().contiguous(), stepper_state)
        forcing = torch.stack([input[n] for n in self.forcing_names], dim=1) if self.forcing_names else None
        ocean = None
        if self.ocean is not None:
            # the reference reads both from next_step_input_data (fme/core/step/single_module.py:708-709)
            nxt = next_step_input_data or {}
            if self._slab is not None:
                ocean = torch.stack([nxt[self.ocean["ocean_fraction_name"]], nxt[self._slab["q_flux_name"]],
                                     nxt[self._slab["mixed_layer_depth_name"]]], dim=1)
            else:
                ocean = torch.stack([nxt[self.ocean["ocean_fraction_name"]], nxt[self.ocean["surface_temperature_name"]]], dim=1)
        cnext = None
        if self.corrector_needs_next:
            from .corrector import next_step_names

            nxt = next_step_input_data or {}
            cnext = torch.stack([nxt[n] for n in next_step_names(self.forcing_names)], dim=1)
        presc = None
        if self.prescribed_prognostic_names:
            nxt = next_step_input_data or {}
            for n in self.prescribed_prognostic_names:
                if n not in nxt:
                    raise ValueError(f"prescribed_prognostic_name '{n}' not in next_step_input_data")
            presc = torch.stack([nxt[n] for n in self.prescribed_prognostic_names], dim=1).float()
        out, _ = self.step_packed(prog, forcing, ocean=ocean, corrector_next=cnext, prescribed=presc)
        return {n: out[:, i] for i, n in enumerate(self.out_names)}

    # ------------------------------------------------------------------ rollout
    def rollout(self, prog0: torch.Tensor, forcing_seq: Optional[torch.Tensor], n_steps: int, use_cuda_graph: bool = True,
                keep_outputs: bool = True, ocean_seq: Optional[torch.Tensor] = None, prescribed_seq: Optional[torch.Tensor] = None,
                corrector_next_seq: Optional[torch.Tensor] = None, stepper_state=None):
        """Autoregressive loop (predict_generator).  forcing_seq [n_steps, B, n_forcing, H, W] resident on device
        ([n_steps + 1, ...] with the energy budget correction, which reads DSWRFtoa / HGTsfc of the output time, like the
        reference's forcing windows of n + 1 times; ``corrector_next_seq`` [n_steps, B, 2, H, W] overrides that selection).
        ocean_seq [n_steps, B, n_ocean, H, W] / prescribed_seq [n_steps, B, n_prescribed, H, W]: data of the OUTPUT time of each step.

        ``stepper_state``: what ``get_stepper_state`` returned after the previous window (the dry-air target stays pinned to the
        first initial condition across windows, fme/ace/stepper/single_module.py:1160-1165); None = ``prog0`` is a new initial condition.

        Returns (outputs [n_steps, B, n_out, H, W] or None, final prognostic state).
        """
        B = prog0.shape[0]
        H, W = self.module.img_shape
        outs = torch.empty(n_steps, B, len(self.out_names), H, W, device=prog0.device) if keep_outputs else None
        state = None
        for t, out, state in self._iter_steps(prog0, forcing_seq, n_steps, use_cuda_graph, ocean_seq, prescribed_seq, corrector_next_seq,
                                              out_bufs=outs, stepper_state=stepper_state):
            if outs is not None and out.data_ptr() != outs[t].data_ptr():
                outs[t].copy_(out, non_blocking=True)
        if state is None:  # n_steps == 0
            return outs, prog0.float().contiguous().clone()
        return outs, state.clone()

    def _iter_steps(self, prog0, forcing_seq, n_steps, use_cuda_graph, ocean_seq, prescribed_seq, corrector_next_seq, out_bufs=None,
                    stepper_state=None):
        """Generator behind ``rollout`` / ``predict_generator``: yields ``(t, out, state)`` after every step, where ``out``
        [B, n_out, H, W] and ``state`` [B, n_prog, H, W] are buffers that the next iteration overwrites (copy what must survive)."""
        B = prog0.shape[0]
        H, W = self.module.img_shape
        dev = prog0.device
        n_out, n_prog, n_presc = len(self.out_names), len(self.prognostic_names), len(self.prescribed_prognostic_names)
        state = prog0.float().contiguous().clone()
        # a rollout starts from an initial condition (capture the dry-air reference from it) or continues a carried state
        self._begin_window(state, stepper_state)
        needs_next = self.corrector_needs_next
        if needs_next and corrector_next_seq is None and (forcing_seq is None or forcing_seq.shape[0] < n_steps + 1):
            raise ValueError("the energy budget correction needs forcing at n_steps + 1 times")
        if n_presc and (prescribed_seq is None or prescribed_seq.shape[0] < n_steps):
            raise ValueError("prescribed_prognostic_names is configured: prescribed_seq [n_steps, B, n_prescribed, H, W] is required")

        def cnext_at(t):
            if not needs_next:
                return None
            return corrector_next_seq[t] if corrector_next_seq is not None else self.corrector_next_from_forcing(forcing_seq[t + 1])

        if not use_cuda_graph:
            out_buf = torch.empty(B, n_out, H, W, device=dev)
            nxt = torch.empty_like(state)
            for t in range(n_steps):
                f = forcing_seq[t] if forcing_seq is not None else None
                o = out_buf if out_bufs is None else out_bufs[t]
                self.step_packed(state, f, o, nxt, ocean=ocean_seq[t] if ocean_seq is not None else None, corrector_next=cnext_at(t),
                                 prescribed=prescribed_seq[t] if n_presc else None)
                state, nxt = nxt, state
                yield t, o, state
            return
        self._sync_before_replay(dev)
        st = self._static
        if self._graph is None or st is None or st["B"] != B or st["prog"].device != dev:
            st = dict(
                B=B, prog=torch.empty_like(state),
                forcing=torch.empty(B, len(self.forcing_names), H, W, device=dev) if self.forcing_names else None,
                out=torch.empty(B, n_out, H, W, device=dev), nxt=torch.empty(B, n_prog, H, W, device=dev),
                ocean=torch.ones(B, self.n_ocean, H, W, device=dev) if self.ocean is not None else None,
                cnext=torch.zeros(B, 2, H, W, device=dev) if needs_next else None,
                presc=torch.zeros(B, n_presc, H, W, device=dev) if n_presc else None,
            )
            st["prog"].copy_(state)
            if st["forcing"] is not None:
                st["forcing"].zero_()
            kw = dict(ocean=st["ocean"], corrector_next=st["cnext"], prescribed=st["presc"])
            self.step_packed(st["prog"], st["forcing"], st["out"], st["nxt"], **kw)  # warm-up: allocations, func attributes
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.step_packed(st["prog"], st["forcing"], st["out"], st["nxt"], **kw)
                st["prog"].copy_(st["nxt"])  # feed back inside the graph
            self._graph, self._static = g, st
        st["prog"].copy_(state)
        for t in range(n_steps):
            if st["forcing"] is not None:
                st["forcing"].copy_(forcing_seq[t], non_blocking=True)
            if st["ocean"] is not None:
                st["ocean"].copy_(ocean_seq[t], non_blocking=True)
            if st["cnext"] is not None:
                st["cnext"].copy_(cnext_at(t), non_blocking=True)
            if st["presc"] is not None:
                st["presc"].copy_(prescribed_seq[t], non_blocking=True)
            self._graph.replay()
            yield t, st["out"], st["prog"]

    # ------------------------------------------------------------------ the reference Stepper's prediction API, TensorMapping level
    TIME_DIM = 1  # fme/ace/stepper/single_module.py: tensors are [n_batch, n_time, n_lat, n_lon]
    n_ic_timesteps = 1  # Stepper.n_ic_timesteps for a single_module step

    @property
    def next_step_input_names(self) -> List[str]:
        """fme/core/step/single_module.py:191-199: what ``next_step_input_data`` carries (input-only names, the ocean's forcing
        names, the prescribed prognostic names)."""
        names = list(self.forcing_names)
        if self.ocean is not None:
            extra = [self.ocean["ocean_fraction_name"]]
            extra += [self._slab["q_flux_name"], self._slab["mixed_layer_depth_name"]] if self._slab else [self.ocean["surface_temperature_name"]]
            names += [n for n in extra if n not in names]
        names += [n for n in self.prescribed_prognostic_names if n not in names]
        return names

    def _window_tensors(self, forcing_dict: Mapping[str, torch.Tensor], n_forward_steps: int, device):
        """The per-step device sequences of one forcing window: inputs of step t come from time t (time t + 1 for
        ``next_step_forcing_names``), ocean / prescribed / corrector data from time t + 1 (single_module.py:1136-1147)."""
        T = n_forward_steps

        def at(name, shift):
            v = forcing_dict[name]
            if v.shape[self.TIME_DIM] < T + 1:
                raise ValueError(f"forcing '{name}' has {v.shape[self.TIME_DIM]} times, {T + 1} are needed for {T} forward steps")
            return v[:, shift:shift + T].to(device=device, dtype=torch.float32)

        def seq(names, shift_of):  # -> [T, B, len(names), H, W]
            return torch.stack([at(n, shift_of(n)) for n in names], dim=2).transpose(0, 1).contiguous()

        fseq = seq(self.forcing_names, lambda n: 1 if n in self.next_step_forcing_names else 0) if self.forcing_names else None
        oseq = None
        if self.ocean is not None:
            frac = self.ocean["ocean_fraction_name"]
            names = [frac, self._slab["q_flux_name"], self._slab["mixed_layer_depth_name"]] if self._slab else \
                [frac, self.ocean["surface_temperature_name"]]
            oseq = seq(names, lambda n: 1)
        pseq = seq(self.prescribed_prognostic_names, lambda n: 1) if self.prescribed_prognostic_names else None
        cseq = None
        if self.corrector_needs_next:
            from .corrector import next_step_names

            cseq = seq(list(next_step_names(self.forcing_names)), lambda n: 1)
        return fseq, oseq, pseq, cseq

    def predict_generator(self, ic_dict: Mapping[str, torch.Tensor], forcing_dict: Mapping[str, torch.Tensor], n_forward_steps: int,
                          use_cuda_graph: bool = True, stepper_state=None):
        """``Stepper.predict_generator`` (fme/ace/stepper/single_module.py:1124-1167) on name -> tensor mappings:
        ``ic_dict[name]`` [B, 1, H, W] for every prognostic name, ``forcing_dict[name]`` [B, n_forward_steps + 1, H, W] for the
        names in ``next_step_input_names``.  Yields, per step, the dict of all ``out_names`` -> [B, H, W] (fresh tensors); the
        prognostic subset is fed back as the next state on the device (one CUDA-graph replay per step)."""
        missing = [n for n in self.prognostic_names if n not in ic_dict]
        if missing:
            raise KeyError(f"initial condition lacks prognostic variables {missing}")
        prog0 = torch.stack([ic_dict[n].squeeze(self.TIME_DIM) for n in self.prognostic_names], dim=1)
        fseq, oseq, pseq, cseq = self._window_tensors(forcing_dict, n_forward_steps, prog0.device)
        for _, out, _ in self._iter_steps(prog0, fseq, n_forward_steps, use_cuda_graph, oseq, pseq, cseq, stepper_state=stepper_state):
            snap = out.clone()
            yield {n: snap[:, i] for i, n in enumerate(self.out_names)}

    def predict(self, initial_condition: Mapping[str, torch.Tensor], forcing: Mapping[str, torch.Tensor], use_cuda_graph: bool = True,
                stepper_state=None):
        """``Stepper.predict`` (fme/ace/stepper/single_module.py:1169-1262) on name -> tensor mappings, without derived variables:
        returns ``(data, new_initial_condition)`` with ``data[name]`` [B, n_forward_steps, H, W] for every output name and
        ``new_initial_condition[name]`` [B, 1, H, W] for every prognostic name (the last predicted time, usable as the next
        window's initial condition).  ``n_forward_steps`` = forcing times - 1, as in the reference.  The reference attaches the
        terminal ``stepper_state`` to the returned prognostic state (:507-521); here ``new_initial_condition`` is a dict subclass
        whose ``.stepper_state`` attribute holds it -- pass it as ``stepper_state=`` of the next window so the dry-air target stays
        pinned to the FIRST initial condition."""
        names = self.next_step_input_names
        if not names:
            raise ValueError("predict: the number of forward steps is taken from the forcing data, which is empty")
        for n in self.prognostic_names:
            if initial_condition[n].shape[self.TIME_DIM] != 1:
                raise ValueError(f"Initial condition must have 1 timesteps, got {initial_condition[n].shape[self.TIME_DIM]}.")
        n_forward_steps = forcing[names[0]].shape[self.TIME_DIM] - 1
        if stepper_state is None:
            stepper_state = getattr(initial_condition, "stepper_state", None)
        prog0 = torch.stack([initial_condition[n].squeeze(self.TIME_DIM) for n in self.prognostic_names], dim=1)
        fseq, oseq, pseq, cseq = self._window_tensors(forcing, n_forward_steps, prog0.device)
        outs, final = self.rollout(prog0, fseq, n_forward_steps, use_cuda_graph, True, oseq, pseq, cseq, stepper_state=stepper_state)
        data = {n: outs[:, :, i].transpose(0, 1) for i, n in enumerate(self.out_names)}
        new_ic = PrognosticDict({n: final[:, i].unsqueeze(self.TIME_DIM) for i, n in enumerate(self.prognostic_names)})
        new_ic.stepper_state = self.get_stepper_state(prog0.shape[0])
        return data, new_ic

    def predict_paired(self, initial_condition: Mapping[str, torch.Tensor], forcing: Mapping[str, torch.Tensor], use_cuda_graph: bool = True,
                       stepper_state=None):
        """``Stepper.predict_paired`` (fme/ace/stepper/single_module.py:1261-1311) on name -> tensor mappings: the prediction paired
        with the reference ("target / forcing") values of every variable of ``forcing`` at the predicted times, i.e. the time axis
        without the initial-condition time (``get_forward_data``, :1313-1327).  Returns ``(prediction, reference), new_ic``."""
        if stepper_state is None:
            stepper_state = getattr(initial_condition, "stepper_state", None)
        prediction, new_ic = self.predict(initial_condition, forcing, use_cuda_graph, stepper_state=stepper_state)
        reference = {n: v[:, self.n_ic_timesteps:] for n, v in forcing.items()}
        return (prediction, reference), new_ic

    def _ensure_graph(self, prog0: torch.Tensor):
        """Capture (once per batch size / device) the CUDA graph of one fused step on static buffers."""
        B = prog0.shape[0]
        st = self._static
        if self._graph is None or st is None or st["B"] != B or st["prog"].device != prog0.device:
            H, W = self.module.img_shape
            nt = 2 if self.corrector_needs_next else 1
            n_presc = len(self.prescribed_prognostic_names)
            self.rollout(prog0, torch.zeros(nt, B, len(self.forcing_names), H, W, device=prog0.device) if self.forcing_names else None, 1,
                         use_cuda_graph=True, keep_outputs=False,
                         ocean_seq=torch.ones(1, B, self.n_ocean, H, W, device=prog0.device) if self.ocean is not None else None,
                         prescribed_seq=torch.zeros(1, B, n_presc, H, W, device=prog0.device) if n_presc else None)
        return self._static, self._graph

    def rollout_host(self, prog0: torch.Tensor, forcing_host: Optional[torch.Tensor], n_steps: int,
                     out_host: Optional[torch.Tensor] = None, ocean_host: Optional[torch.Tensor] = None,
                     prescribed_host: Optional[torch.Tensor] = None, corrector_next_host: Optional[torch.Tensor] = None,
                     stepper_state=None):
        """Autoregressive loop with HOST-resident forcing and outputs (the inference driver's situation: forcing windows come
        from the data loader, outputs go to the writers; ``fme/core/generics/inference.py:117-166``).

        forcing_host [>= n_steps (cycled), B, n_forcing, H, W] pinned; out_host [n_steps, B, n_out, H, W] pinned (or None);
        ocean_host / prescribed_host [>= n_steps (cycled), B, n, H, W] pinned: data of each step's OUTPUT time.
        The packed host forcing must already carry the ``next_step_forcing_names`` shift (the value of such a channel at index t is
        the one of time t + 1), as ``_window_tensors`` produces it.
        With the energy budget correction the (DSWRFtoa, HGTsfc) pair of the OUTPUT time is taken from ``corrector_next_host[t]``
        ([>= n_steps, B, 2, H, W] pinned) when given, else from ``forcing_host[t + 1]`` -- which then needs ``n_steps + 1`` forcing
        times (no wrap-around: silently re-using time 0 as the next-time insolation would be wrong; ``rollout`` raises likewise).
        Every step copies its forcing host->device and its outputs device->host; both copies run on side streams and overlap the
        neighbouring steps' compute (double-buffered staging), the step itself is one CUDA-graph replay.
        Returns the final prognostic state (device).  The caller synchronises before reading ``out_host``.
        """
        dev = prog0.device
        if bool(self.prescribed_prognostic_names) != (prescribed_host is not None):
            raise ValueError("prescribed_host must be given exactly when prescribed_prognostic_names is configured")
        if prescribed_host is not None and not self.forcing_names:
            raise NotImplementedError("rollout_host: prescribed data is staged alongside the forcing; a network without forcing inputs is not handled")
        if self.corrector_needs_next and corrector_next_host is None and (forcing_host is None or forcing_host.shape[0] < n_steps + 1):
            raise ValueError("the energy budget correction reads (DSWRFtoa, HGTsfc) at the output time: give forcing_host with "
                             "n_steps + 1 times or corrector_next_host [n_steps, B, 2, H, W]")
        self._sync_before_replay(dev)
        st, graph = self._ensure_graph(prog0)
        self._begin_window(prog0.float().contiguous(), stepper_state)
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_h"):
            self._h = None
        h = self._h
        if h is None or h["B"] != st["B"] or h["dev"] != dev:
            h = dict(B=st["B"], dev=dev, s_in=torch.cuda.Stream(dev), s_out=torch.cuda.Stream(dev),
                     fst=[torch.empty_like(st["forcing"]) for _ in range(2)] if st["forcing"] is not None else None,
                     cst=[torch.empty_like(st["ocean"]) for _ in range(2)] if st["ocean"] is not None else None,
                     nst=[torch.empty_like(st["cnext"]) for _ in range(2)] if st["cnext"] is not None else None,
                     pst=[torch.empty_like(st["presc"]) for _ in range(2)] if st["presc"] is not None else None,
                     ost=[torch.empty_like(st["out"]) for _ in range(2)])
            self._h = h
        ev_in = [torch.cuda.Event() for _ in range(2)]        # forcing staged on device
        ev_used = [torch.cuda.Event() for _ in range(2)]      # forcing stage consumed by the step
        ev_ready = [torch.cuda.Event() for _ in range(2)]     # output staged for the host copy
        ev_done = [torch.cuda.Event() for _ in range(2)]      # host copy of the output finished
        st["prog"].copy_(prog0)
        nf = forcing_host.shape[0] if forcing_host is not None else 0

        def stage_forcing(t):
            b = t & 1
            with torch.cuda.stream(h["s_in"]):
                if t >= 2:
                    h["s_in"].wait_event(ev_used[b])
                h["fst"][b].copy_(forcing_host[t % nf], non_blocking=True)
                if h["cst"] is not None:
                    h["cst"][b].copy_(ocean_host[t % ocean_host.shape[0]], non_blocking=True)
                if h["nst"] is not None:
                    if corrector_next_host is not None:
                        h["nst"][b].copy_(corrector_next_host[t % corrector_next_host.shape[0]], non_blocking=True)
                    else:
                        h["nst"][b].copy_(self.corrector_next_from_forcing(forcing_host[t + 1]), non_blocking=True)
                if h["pst"] is not None:
                    h["pst"][b].copy_(prescribed_host[t % prescribed_host.shape[0]], non_blocking=True)
                ev_in[b].record(h["s_in"])

        if h["fst"] is not None and n_steps > 0:
            h["s_in"].wait_stream(cur)
            stage_forcing(0)
        for t in range(n_steps):
            b = t & 1
            if h["fst"] is not None:
                if t + 1 < n_steps:
                    stage_forcing(t + 1)
                cur.wait_event(ev_in[b])
                st["forcing"].copy_(h["fst"][b], non_blocking=True)
                if h["cst"] is not None:
                    st["ocean"].copy_(h["cst"][b], non_blocking=True)
                if h["nst"] is not None:
                    st["cnext"].copy_(h["nst"][b], non_blocking=True)
                if h["pst"] is not None:
                    st["presc"].copy_(h["pst"][b], non_blocking=True)
                ev_used[b].record(cur)
            graph.replay()
            if out_host is not None:
                if t >= 2:
                    cur.wait_event(ev_done[b])
                h["ost"][b].copy_(st["out"], non_blocking=True)
                ev_ready[b].record(cur)
                with torch.cuda.stream(h["s_out"]):
                    h["s_out"].wait_event(ev_ready[b])
                    out_host[t].copy_(h["ost"][b], non_blocking=True)
                    ev_done[b].record(h["s_out"])
        cur.wait_stream(h["s_out"])
        cur.wait_stream(h["s_in"])
        return st["prog"].clone()
