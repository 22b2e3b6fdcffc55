"""Conservation correctors of the post-step state, on the device.

Mirrors ``AtmosphereCorrector.__call__`` of the reference (``fme/core/corrector/atmosphere.py:349-398``) for its two global
budget options: ``conserve_dry_air`` (``:404-463``) and ``moisture_budget_correction`` (``:518-608``).  ForcePositive is the
fused step's clamp (``ace_b200/stepper.py``); the remaining options of the reference config (total-energy budget, zero
global-mean moisture advection) are not on this path and raise ``NotImplementedError``.

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
}
_WATER_PREFIX = "specific_total_water_"


def _find(names: Sequence[str], key: str) -> int:
    for p in _FIELD_NAMES[key]:
        if p in names:
            return list(names).index(p)
    return -1


def _levels(names: Sequence[str]):
    lv = sorted((int(n[len(_WATER_PREFIX):]), i) for i, n in enumerate(names)
                if n.startswith(_WATER_PREFIX) and n[len(_WATER_PREFIX):].isdigit())
    return [i for _, i in lv]


class AtmosphereCorrector:
    def __init__(self, out_names: Sequence[str], prognostic_names: Sequence[str], img_shape, ak, bk, area_weights,
                 timestep_seconds: float = 21600.0, conserve_dry_air: bool = False,
                 moisture_budget_correction: Optional[str] = None, **unsupported):
        for k, v in unsupported.items():
            if k in ("zero_global_mean_moisture_advection", "total_energy_budget_correction") and v:
                raise NotImplementedError(f"AtmosphereCorrector: option {k!r} is not implemented on the B200 path")
            if k not in ("zero_global_mean_moisture_advection", "total_energy_budget_correction", "force_positive_names"):
                raise TypeError(f"AtmosphereCorrector: unknown option {k!r}")
        if moisture_budget_correction not in MOISTURE_MODES:
            raise ValueError(f"moisture_budget_correction must be one of {list(MOISTURE_MODES)}")
        H, W = img_shape
        self.out_names, self.prognostic_names = list(out_names), list(prognostic_names)
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

    def seed(self, prog: torch.Tensor):
        """Capture the global dry-air mass of the initial condition ``prog [B, n_prog, H, W]`` (atmosphere.py:404-427)."""
        self._check(prog, len(self.prognostic_names))
        with torch.cuda.device(prog.device):
            _lib.check(_lib.load().ace_corrector_seed(self._handle, ctypes.c_void_p(prog.data_ptr()), prog.shape[0], _lib.current_stream_ptr()))

    def apply(self, prev_prog: torch.Tensor, out: torch.Tensor, next_prog: Optional[torch.Tensor] = None):
        """Correct ``out`` (and the matching channels of ``next_prog``) in place; ``prev_prog`` is the step's input state."""
        self._check(prev_prog, len(self.prognostic_names))
        self._check(out, len(self.out_names))
        if next_prog is not None:
            self._check(next_prog, len(self.prognostic_names))
        with torch.cuda.device(out.device):
            _lib.check(_lib.load().ace_corrector_apply(
                self._handle, ctypes.c_void_p(prev_prog.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                ctypes.c_void_p(next_prog.data_ptr()) if next_prog is not None else None, out.shape[0], _lib.current_stream_ptr()))
        return out

    def _check(self, t: torch.Tensor, c: int):
        if not t.is_cuda:
            raise _lib.AceError("AtmosphereCorrector: tensors must be on a CUDA device (there is no CPU path)")
        if t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape[1:]) != (c, *self.shape):
            raise ValueError(f"AtmosphereCorrector: expected contiguous fp32 [B, {c}, {self.shape[0]}, {self.shape[1]}], got {t.dtype} {tuple(t.shape)}")
