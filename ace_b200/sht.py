"""RealSHT / InverseRealSHT with the reference's object contract, executed by libace_b200.

Contract mirrored (``/root/reference/fme/sht_fix.py:60-226``, the classes fme monkey-patches
over ``torch_harmonics.RealSHT/InverseRealSHT`` at ``:228-229``):
constructor ``(nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True)``,
attributes ``nlat nlon lmax mmax grid norm csphase``, ``forward(x[..., nlat, nlon]) ->
complex64[..., lmax, mmax]`` (inverse symmetrical), usable under ``.float()`` / ``.to()``.
The Legendre tables are NOT parameters or buffers (same as the reference, fme/sht_fix.py:117).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .legendre import grid_nodes, sht_tables


class ShtPlan:
    """Owns one ace_sht_plan (device tables for a forward+inverse pair). Shared between modules."""

    _cache = {}

    def __init__(self, nlat, nlon, lmax, mmax, grid, norm, csphase):
        fwd, inv, lmax, mmax = sht_tables(nlat, nlon, lmax, mmax, grid, norm, csphase)
        self.nlat, self.nlon, self.lmax, self.mmax = nlat, nlon, lmax, mmax
        lib = _lib.load()
        handle = ctypes.c_void_p()
        fwd = np.ascontiguousarray(fwd, dtype=np.float64)
        inv = np.ascontiguousarray(inv, dtype=np.float64)
        _lib.check(lib.ace_sht_plan_create(nlat, nlon, lmax, mmax, fwd.ctypes.data_as(ctypes.c_void_p),
                                           inv.ctypes.data_as(ctypes.c_void_p), ctypes.byref(handle)))
        self.handle = handle
        self.device = torch.cuda.current_device()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().ace_sht_plan_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    @classmethod
    def get(cls, nlat, nlon, lmax, mmax, grid, norm="ortho", csphase=True):
        if not torch.cuda.is_available():
            raise _lib.AceError("ace_b200 needs a CUDA device (sm_100a); there is no CPU path")
        key = (torch.cuda.current_device(), nlat, nlon, lmax, mmax, grid, norm, csphase)
        plan = cls._cache.get(key)
        if plan is None:
            plan = cls(nlat, nlon, lmax, mmax, grid, norm, csphase)
            cls._cache[key] = plan
        return plan


def _check_input(x, what):
    if not x.is_cuda:
        raise _lib.AceError(f"{what}: input must be a CUDA tensor (ace_b200 has no CPU path)")


class _ShtBase(nn.Module):
    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
        super().__init__()
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        _, _, lmax_default = grid_nodes(grid, nlat)  # also validates `grid` exactly like the reference
        self.lmax = lmax or lmax_default
        self.mmax = mmax or nlon // 2 + 1
        self._plan = None

    def plan(self):
        if self._plan is None or self._plan.device != torch.cuda.current_device():
            self._plan = ShtPlan.get(self.nlat, self.nlon, self.lmax, self.mmax, self.grid, self.norm, self.csphase)
        return self._plan

    def extra_repr(self):
        return f"nlat={self.nlat}, nlon={self.nlon},\n lmax={self.lmax}, mmax={self.mmax},\n grid={self.grid}, csphase={self.csphase}"


class RealSHT(_ShtBase):
    """Forward real SHT; replaces fme.sht_fix.RealSHT (fme/sht_fix.py:60-151)."""

    def forward(self, x: torch.Tensor):
        assert x.shape[-2] == self.nlat
        assert x.shape[-1] == self.nlon
        _check_input(x, "RealSHT")
        x = x.float().contiguous()
        lead = x.shape[:-2]
        nf = int(np.prod(lead)) if len(lead) else 1
        out = torch.empty(*lead, self.lmax, self.mmax, dtype=torch.complex64, device=x.device)
        if nf == 0:
            return out
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ace_sht_forward(self.plan().handle, ctypes.c_void_p(x.data_ptr()),
                                                   ctypes.c_void_p(out.data_ptr()), nf, _lib.current_stream_ptr()))
        return out


class InverseRealSHT(_ShtBase):
    """Inverse real SHT; replaces fme.sht_fix.InverseRealSHT (fme/sht_fix.py:153-226)."""

    def forward(self, x: torch.Tensor):
        assert x.shape[-2] == self.lmax
        assert x.shape[-1] == self.mmax
        _check_input(x, "InverseRealSHT")
        x = x.to(torch.complex64).contiguous()
        lead = x.shape[:-2]
        nf = int(np.prod(lead)) if len(lead) else 1
        out = torch.empty(*lead, self.nlat, self.nlon, dtype=torch.float32, device=x.device)
        if nf == 0:
            return out
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ace_sht_inverse(self.plan().handle, ctypes.c_void_p(x.data_ptr()),
                                                   ctypes.c_void_p(out.data_ptr()), nf, _lib.current_stream_ptr()))
        return out


def patch_torch_harmonics():
    """Install these classes the way fme does (fme/sht_fix.py:228-229) if torch_harmonics is importable."""
    import torch_harmonics  # noqa: PLC0415

    from .registry import disabled_by_env  # noqa: PLC0415

    if disabled_by_env():  # ACE_B200_DISABLE=1: A/B kill switch, the reference transforms stay in place
        return
    torch_harmonics.RealSHT = RealSHT
    torch_harmonics.InverseRealSHT = InverseRealSHT
