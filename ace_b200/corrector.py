"""Conservation correctors of the post-step state, on the device.

Mirrors ``AtmosphereCorrector.__call__`` of the reference (``fme/core/corrector/atmosphere.py:349-398``) for its two global
budget options: ``conserve_dry_air`` (``:404-463``) and ``moisture_budget_correction`` (``:518-608``).  ForcePositive is the
fused step's clamp (``ace_b200/stepper.py``).  Also the remaining members of the sequence the reference builds at
``:349-398``: ``zero_global_mean_moisture_advection`` (``:467-490``), ``clip_frozen_precipitation`` (``:493-515``) and
``total_energy_budget_correction`` with method ``constant_temperature`` (``:611-695``).

Works on the packed tensors of the fused step: ``out [B, n_out, H, W]`` (denormalised) and the prognostic state
``[B, n_prog, H, W]``.  Field names are resolved like ``fme/core/atmosphere_data.py:17-41``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

MOISTURE_MODES = {None: 0, "precipitation": 1, "advection_and_precipitation": 2, "evaporation": 3, "advection_and_evaporation": 4}

_FIELD_NAMES = {  # fme/core/atmosphere_data.py:17-41
    "surface_pressure": ["PRESsfc", "PS"],
    "precipitation_rate": ["PRATEsfc", "surface_precipitation_rate"],
    "latent_heat_flux": ["LHTFLsfc", "LHFLX"],
    "tendency_of_total_water_path_due_to_advection": ["tendency_of_total_water_path_due_to_advection"],
    "sensible_heat_flux": ["SHTFLsfc", "SHFLX"],
    "sfc_down_sw_radiative_flux": ["DSWRFsfc", "FSDS"],
    "sfc_up_sw_radiative_flux": ["USWRFsfc", "surface_upward_shortwave_flux"],
    "sfc_down_lw_radiative_flux": ["DLWRFsfc", "FLDS"],
    "sfc_up_lw_radiative_flux": ["ULWRFsfc", "surface_upward_longwave_flux"],
    "toa_up_lw_radiative_flux": ["ULWRFtoa", "FLUT"],
    "toa_up_sw_radiative_flux": ["USWRFtoa", "top_of_atmos_upward_shortwave_flux"],
    "toa_down_sw_radiative_flux": ["DSWRFtoa", "SOLIN"],
    "frozen_precipitation_rate": ["total_frozen_precipitation_rate"],
    "surface_height": ["HGTsfc"],
}
_WATER_PREFIX = "specific_total_water_"
_TEMP_PREFIXES = ["air_temperature_", "T_"]


def _find(names: Sequence[str], key: str) -> int:
    for p in _FIELD_NAMES[key]:
        if p in names:
            return list(names).index(p)
    return -1


def _levels(names: Sequence[str], prefixes=(_WATER_PREFIX,)):
    for prefix in prefixes:
        lv = sorted((int(n[len(prefix):]), i) for i, n in enumerate(names) if n.startswith(prefix) and n[len(prefix):].isdigit())
        if lv:
            return [i for _, i in lv]
    return []


def next_step_names(forcing_names: Sequence[str]):
    """Names of the two next-step fields the energy budget correction reads (atmosphere.py:626-633): (DSWRFtoa, HGTsfc)."""
    i, j = _find(forcing_names, "toa_down_sw_radiative_flux"), _find(forcing_names, "surface_height")
    if i < 0 or j < 0:
        raise ValueError("total_energy_budget_correction needs DSWRFtoa (or SOLIN) and HGTsfc among the forcing inputs")
    return forcing_names[i], forcing_names[j]


class AtmosphereCorrector:
    def __init__(self, out_names: Sequence[str], prognostic_names: Sequence[str], img_shape, ak, bk, area_weights,
                 timestep_seconds: float = 21600.0, conserve_dry_air: bool = False,
                 moisture_budget_correction: Optional[str] = None, zero_global_mean_moisture_advection: bool = False,
                 clip_frozen_precipitation: bool = False, total_energy_budget_correction=None,
                 forcing_names: Sequence[str] = (), force_positive_names: Sequence[str] = ()):
        """Options are the fields of the reference's ``AtmosphereCorrectorConfig`` (``atmosphere.py:223-337``);
        ``total_energy_budget_correction``: ``None`` or ``dict(method="constant_temperature", constant_unaccounted_heating=0.0)``;
        ``forcing_names``: channel names of the step's forcing input (the energy correction reads the surface height there)."""
        del force_positive_names  # the fused step's clamp (FusedStepper(force_positive_names=...))
        if moisture_budget_correction not in MOISTURE_MODES:
            raise ValueError(f"moisture_budget_correction must be one of {list(MOISTURE_MODES)}")
        energy = total_energy_budget_correction
        if energy is not None:
            energy = dict(energy) if not hasattr(energy, "method") else dict(
                method=energy.method, constant_unaccounted_heating=getattr(energy, "constant_unaccounted_heating", 0.0))
            if energy.get("method", "constant_temperature") != "constant_temperature":
                raise NotImplementedError(f"total_energy_budget_correction method {energy.get('method')!r} (the reference implements "
                                          "constant_temperature only, atmosphere.py:625-628)")
        H, W = img_shape
        self.out_names, self.prognostic_names, self.forcing_names = list(out_names), list(prognostic_names), list(forcing_names)
        ak, bk = np.asarray(ak, dtype=np.float64), np.asarray(bk, dtype=np.float64)
        nz = len(ak) - 1
        out_wat, prog_wat = _levels(self.out_names), _levels(self.prognostic_names)
        if len(out_wat) != nz or len(prog_wat) != nz:
            raise ValueError(f"corrector: {nz} vertical layers need {_WATER_PREFIX}0..{nz - 1} among the prognostic outputs")
        w = np.ascontiguousarray(np.asarray(area_weights, dtype=np.float32).reshape(-1))
        if w.size != H * W:
            raise ValueError("corrector: area_weights must be [H, W]")
        # [n_out]: the prognostic channel an output channel feeds back into, or -1 (diagnostic)
        out_prog = np.array([self.prognostic_names.index(n) if n in self.prognostic_names else -1 for n in self.out_names], dtype=np.int32)
        out_wat_a, prog_wat_a = np.array(out_wat, dtype=np.int32), np.array(prog_wat, dtype=np.int32)

        def ip(a):
            return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))

        out_temp = np.zeros(nz, dtype=np.int32)
        prog_temp = np.zeros(nz, dtype=np.int32)
        flux = dict(out_dlw_sfc=-1, out_ulw_sfc=-1, out_dsw_sfc=-1, out_usw_sfc=-1, out_shf=-1, out_usw_toa=-1, out_ulw_toa=-1)
        forcing_hgt = -1
        if energy is not None:
            ot, pt = _levels(self.out_names, _TEMP_PREFIXES), _levels(self.prognostic_names, _TEMP_PREFIXES)
            if len(ot) != nz or len(pt) != nz:
                raise ValueError(f"total_energy_budget_correction: {nz} layers need air_temperature_0..{nz - 1} among the prognostic outputs")
            out_temp, prog_temp = np.array(ot, dtype=np.int32), np.array(pt, dtype=np.int32)
            keys = dict(out_dlw_sfc="sfc_down_lw_radiative_flux", out_ulw_sfc="sfc_up_lw_radiative_flux", out_dsw_sfc="sfc_down_sw_radiative_flux",
                        out_usw_sfc="sfc_up_sw_radiative_flux", out_shf="sensible_heat_flux", out_usw_toa="toa_up_sw_radiative_flux",
                        out_ulw_toa="toa_up_lw_radiative_flux")
            for k, std in keys.items():
                flux[k] = _find(self.out_names, std)
                if flux[k] < 0:
                    raise ValueError(f"total_energy_budget_correction: no output field for {std} ({_FIELD_NAMES[std]})")
            if any(n in self.out_names for n in ("ICEsfc", "GRAUPELsfc", "SNOWsfc")) and _find(self.out_names, "frozen_precipitation_rate") < 0:
                raise NotImplementedError("frozen precipitation as ICEsfc + GRAUPELsfc + SNOWsfc (atmosphere_data.py:205-211) is not implemented; "
                                          "predict total_frozen_precipitation_rate")
            self.next_step_names = next_step_names(self.forcing_names)
            forcing_hgt = _find(self.forcing_names, "surface_height")
        self.needs_next = energy is not None
        cfg = _lib.CorrectorConfig(
            n_out=len(self.out_names), n_prog=len(self.prognostic_names), nz=nz, hw=H * W,
            area_weights_host=w.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
            ak_host=ak.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), bk_host=bk.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            out_prog_index_host=ip(out_prog),
            out_ps=_find(self.out_names, "surface_pressure"), out_wat_host=ip(out_wat_a),
            out_precip=_find(self.out_names, "precipitation_rate"), out_lhf=_find(self.out_names, "latent_heat_flux"),
            out_adv=_find(self.out_names, "tendency_of_total_water_path_due_to_advection"),
            prog_ps=_find(self.prognostic_names, "surface_pressure"), prog_wat_host=ip(prog_wat_a),
            conserve_dry_air=int(bool(conserve_dry_air)), moisture_mode=MOISTURE_MODES[moisture_budget_correction],
            timestep_seconds=float(timestep_seconds),
            zero_global_mean_moisture_advection=int(bool(zero_global_mean_moisture_advection)),
            out_frozen=_find(self.out_names, "frozen_precipitation_rate"), clip_frozen_precipitation=int(bool(clip_frozen_precipitation)),
            energy_mode=int(energy is not None),
            unaccounted_heating=float(energy.get("constant_unaccounted_heating", 0.0)) if energy is not None else 0.0,
            out_temp_host=ip(out_temp), prog_temp_host=ip(prog_temp), n_forcing=len(self.forcing_names), forcing_hgt=forcing_hgt, **flux,
        )
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().ace_corrector_create(ctypes.byref(cfg), ctypes.byref(handle)))
        self._handle = handle
        self.shape = (H, W)

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _lib.load().ace_corrector_destroy(h)
            except Exception:  # noqa: BLE001
                pass

    # ------------------------------------------------------------------
    def reset(self):
        """Forget the dry-air reference: the next seed() (or the next fused step's input) is a new initial condition."""
        _lib.check(_lib.load().ace_corrector_reset(self._handle))

    @property
    def seeded(self) -> bool:
        return bool(_lib.load().ace_corrector_is_seeded(self._handle))

    def get_state(self, batch: int) -> torch.Tensor:
        """``CorrectorState.global_dry_air_mass`` (fme/core/corrector/state.py:15-29): the fp64 dry-air target of each sample,
        shape ``(batch, 1, 1)`` on the host.  Raises ``AceError`` when no reference has been seeded."""
        import numpy as np

        buf = np.empty(batch, dtype=np.float64)
        _lib.check(_lib.load().ace_corrector_get_state(self._handle, buf.ctypes.data_as(ctypes.c_void_p), batch, _lib.current_stream_ptr()))
        return torch.from_numpy(buf).view(batch, 1, 1)

    def set_state(self, global_dry_air_mass: torch.Tensor):
        """Install a dry-air target carried over from an earlier window; the next step does not re-seed from its input."""
        import numpy as np

        buf = np.ascontiguousarray(global_dry_air_mass.detach().double().cpu().reshape(-1).numpy())
        _lib.check(_lib.load().ace_corrector_set_state(self._handle, buf.ctypes.data_as(ctypes.c_void_p), buf.shape[0], _lib.current_stream_ptr()))

    def seed(self, prog: torch.Tensor):
        """Capture the global dry-air mass of the initial condition ``prog [B, n_prog, H, W]`` (atmosphere.py:404-427)."""
        self._check(prog, len(self.prognostic_names))
        with torch.cuda.device(prog.device):
            _lib.check(_lib.load().ace_corrector_seed(self._handle, ctypes.c_void_p(prog.data_ptr()), prog.shape[0], _lib.current_stream_ptr()))

    def apply(self, prev_prog: torch.Tensor, out: torch.Tensor, next_prog: Optional[torch.Tensor] = None,
              prev_forcing: Optional[torch.Tensor] = None, next_step: Optional[torch.Tensor] = None):
        """Correct ``out`` (and the matching channels of ``next_prog``) in place; ``prev_prog`` / ``prev_forcing`` are the step's
        input state and forcing, ``next_step [B, 2, H, W]`` = (DSWRFtoa, HGTsfc) at the output time (energy correction only)."""
        self._check(prev_prog, len(self.prognostic_names))
        self._check(out, len(self.out_names))
        if next_prog is not None:
            self._check(next_prog, len(self.prognostic_names))
        if self.needs_next:
            if prev_forcing is None or next_step is None:
                raise ValueError("the energy budget correction needs prev_forcing and next_step = (DSWRFtoa, HGTsfc) at the output time")
            self._check(prev_forcing, len(self.forcing_names))
            self._check(next_step, 2)

        def ptr(t):
            return ctypes.c_void_p(t.data_ptr()) if t is not None else None

        with torch.cuda.device(out.device):
            _lib.check(_lib.load().ace_corrector_apply(
                self._handle, ptr(prev_prog), ptr(prev_forcing) if self.needs_next else None, ptr(next_step) if self.needs_next else None,
                ptr(out), ptr(next_prog), out.shape[0], _lib.current_stream_ptr()))
        return out

    def _check(self, t: torch.Tensor, c: int):
        if not t.is_cuda:
            raise _lib.AceError("AtmosphereCorrector: tensors must be on a CUDA device (there is no CPU path)")
        if t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape[1:]) != (c, *self.shape):
            raise ValueError(f"AtmosphereCorrector: expected contiguous fp32 [B, {c}, {self.shape[0]}, {self.shape[1]}], got {t.dtype} {tuple(t.shape)}")
