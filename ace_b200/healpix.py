"""HEALPix spherical harmonic transform, executed by libace_b200.

Mirrors the object contract of the reference's vendored cuHPX Python transform
(``/root/reference/fme/core/cuhpx/sht.py:32-98`` ``SHT`` and ``:101-153`` ``iSHT``, used by
``fme/core/gridded_ops.py:414-460,519-529`` with ``lmax = mmax = 2*nside - 1``):

* ``HealpixSHT(nside, lmax, mmax, quad_weights="ring", ring_weights=None)``:
  ``forward(x[..., 12*nside**2]) -> complex64[..., lmax, mmax]``, pixels in RING order;
* ``HealpixISHT(nside, lmax, mmax)``: the inverse.

The per-ring FFT loop of the reference (4*nside - 1 ``torch.fft`` calls + phase shifts,
``fme/core/cuhpx/tools.py:34-83``) is a ring-DFT kernel plus a tile transposition per direction; the Legendre contraction reuses the tcgen05 GEMMs
of the lat-lon transform.  Unlike the reference's loop -- which takes the ring count from ``ftm.shape[0]`` and is
therefore only correct for unbatched 1-D input -- any leading batch dims are transformed field by field.

``quad_weights="ring"`` needs healpy's per-ring quadrature weights, which the reference ships as data files
(``fme/core/cuhpx/data/weight_ring_n*.npy``): pass ``ring_weights = fme.core.cuhpx.tools.apply_ring_weight(nside)``
(the full per-ring array ``4 pi / npix * (1 + w)``, length ``4*nside - 1``).  ``quad_weights="none"`` uses the uniform
``4 pi / npix`` (``tools.py:228-231``).  No CPU path.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .legendre import legendre_table


def ring_cos_theta(nside: int) -> np.ndarray:
    """cos(colatitude) of the 4*nside - 1 rings as ``healpix_weights`` returns it (fme/core/cuhpx/tools.py:222-240)."""
    t = np.arange(4 * nside - 1)
    z = np.zeros(t.shape, dtype=np.float64)
    cap_n = t < (nside - 1)
    belt = (t >= (nside - 1)) & (t <= (3 * nside - 1))
    cap_s = t > (3 * nside - 1)
    z[cap_n] = 1 - ((t[cap_n] + 1) ** 2) / (3 * nside**2)
    z[belt] = 4 / 3 - 2 * (t[belt] + 1) / (3 * nside)
    z[cap_s] = ((4 * nside - 1 - t[cap_s]) ** 2) / (3 * nside**2) - 1
    return np.flip(z)


def healpix_tables(nside, lmax, mmax, weights):
    """(forward incl. ring weights, inverse) float64 [mmax, lmax, 4*nside-1]; no Condon-Shortley sign (tools.py:332-334)."""
    x = np.cos(np.flip(np.arccos(ring_cos_theta(nside))))
    fwd = legendre_table(mmax, lmax, x, norm="ortho", inverse=False, csphase=False) * np.asarray(weights, dtype=np.float64)[None, None, :]
    inv = legendre_table(mmax, lmax, x, norm="ortho", inverse=True, csphase=False)
    return np.ascontiguousarray(fwd), np.ascontiguousarray(inv)


class _HpxPlan:
    _cache = {}

    def __init__(self, nside, lmax, mmax, weights):
        fwd, inv = healpix_tables(nside, lmax, mmax, weights)
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().ace_sht_plan_create(4 * nside - 1, 4 * nside, lmax, mmax, fwd.ctypes.data_as(ctypes.c_void_p),
                                                   inv.ctypes.data_as(ctypes.c_void_p), ctypes.byref(handle)))
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().ace_sht_plan_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001
            pass

    @classmethod
    def get(cls, nside, lmax, mmax, weights):
        if not torch.cuda.is_available():
            raise _lib.AceError("ace_b200 needs a CUDA device (sm_100a); there is no CPU path")
        key = (torch.cuda.current_device(), nside, lmax, mmax, np.asarray(weights, dtype=np.float64).tobytes())
        plan = cls._cache.get(key)
        if plan is None:
            plan = cls._cache[key] = cls(nside, lmax, mmax, weights)
        return plan


class _HpxBase(nn.Module):
    def __init__(self, nside, lmax=None, mmax=None, grid="healpix", quad_weights="ring", ring_weights=None, norm="ortho",
                 csphase=True):
        super().__init__()
        if grid != "healpix":
            raise ValueError("Unknown quadrature mode")
        if norm != "ortho":
            raise NotImplementedError("ace_b200 HEALPix SHT implements norm='ortho' only")
        self.nside, self.grid, self.norm, self.csphase = nside, grid, norm, csphase
        self.nlat, self.nlon = 4 * nside - 1, 4 * nside
        self.lmax = lmax or self.nlat
        self.mmax = mmax or (self.nlon // 2 + 1)
        self.quad_weights = quad_weights
        npix = 12 * nside**2
        if quad_weights == "ring":
            if ring_weights is None:
                raise ValueError("quad_weights='ring' needs ring_weights (fme.core.cuhpx.tools.apply_ring_weight(nside)); "
                                 "use quad_weights='none' for the uniform 4 pi / npix weights")
            w = np.asarray(ring_weights, dtype=np.float64)
            if w.shape != (self.nlat,):
                raise ValueError(f"ring_weights must have shape ({self.nlat},), got {w.shape}")
        else:
            w = 4.0 * np.pi / npix * np.ones(self.nlat)
        self._w = w
        self._plan = None

    def plan(self):
        if self._plan is None:
            self._plan = _HpxPlan.get(self.nside, self.lmax, self.mmax, self._w)
        return self._plan


class HealpixSHT(_HpxBase):
    """fme/core/cuhpx/sht.py:32-98."""

    def forward(self, x: torch.Tensor):
        if torch.is_complex(x):
            raise ValueError("Input tensor must be real.")
        if not x.is_cuda:
            raise _lib.AceError("HealpixSHT: input must be a CUDA tensor (ace_b200 has no CPU path)")
        npix = 12 * self.nside**2
        assert x.shape[-1] == npix
        x = x.float().contiguous()
        lead = x.shape[:-1]
        nf = int(np.prod(lead)) if len(lead) else 1
        out = torch.empty(*lead, self.lmax, self.mmax, dtype=torch.complex64, device=x.device)
        if nf:
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().ace_hpx_forward(self.plan().handle, self.nside, ctypes.c_void_p(x.data_ptr()),
                                                       ctypes.c_void_p(out.data_ptr()), nf, _lib.current_stream_ptr()))
        return out


class HealpixISHT(_HpxBase):
    """fme/core/cuhpx/sht.py:101-153."""

    def __init__(self, nside, lmax=None, mmax=None, grid="healpix", norm="ortho", csphase=True):
        super().__init__(nside, lmax, mmax, grid, quad_weights="none", norm=norm, csphase=csphase)

    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise _lib.AceError("HealpixISHT: input must be a CUDA tensor (ace_b200 has no CPU path)")
        assert x.shape[-2] == self.lmax and x.shape[-1] == self.mmax
        x = x.to(torch.complex64).contiguous()
        lead = x.shape[:-2]
        nf = int(np.prod(lead)) if len(lead) else 1
        out = torch.empty(*lead, 12 * self.nside**2, dtype=torch.float32, device=x.device)
        if nf:
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().ace_hpx_inverse(self.plan().handle, self.nside, ctypes.c_void_p(x.data_ptr()),
                                                       ctypes.c_void_p(out.data_ptr()), nf, _lib.current_stream_ptr()))
        return out
