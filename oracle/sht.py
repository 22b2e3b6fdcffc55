"""RealSHT / InverseRealSHT -- oracle restatement on torch-CPU (test infrastructure).

Follows /root/reference/fme/sht_fix.py:60-226 and /root/reference/fme/fft.py:60-96
operation by operation (rfft norm="forward" x 2*pi, truncate/zero-pad to mmax,
separate real/imag Legendre contractions against fp32 tables, irfft after zeroing
Im(m=0) and Im(Nyquist)), so that on the same fp32 input it reproduces the
reference's stored goldens to rounding.  Tables are built in float64 by
``oracle.legendre`` / ``oracle.quadrature`` and cast to fp32 exactly as
fme/sht_fix.py:113-117 and :195-198 do.
"""
import numpy as np
import torch

from .legendre import precompute_legpoly
from .quadrature import nodes_and_weights


def _colatitudes(cost):
    # fme/sht_fix.py:107 / :191 -- arccos, then flip
    return np.flip(np.arccos(np.asarray(cost, dtype=np.float64)), axis=0).copy()


def forward_table(nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
    """fp64 [mmax, lmax, nlat] table = P_l^m(cos theta_k) * w_k (fme/sht_fix.py:91-117)."""
    cost, w, lmax_default = nodes_and_weights(grid, nlat)
    lmax = lmax or lmax_default
    mmax = mmax or nlon // 2 + 1
    tq = _colatitudes(cost)
    pct = precompute_legpoly(mmax, lmax, tq, norm=norm, csphase=csphase)
    return pct * np.asarray(w, dtype=np.float64)[None, None, :], lmax, mmax


def inverse_table(nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
    """fp64 [mmax, lmax, nlat] table = P_l^m(cos theta_k) (fme/sht_fix.py:175-198)."""
    cost, _, lmax_default = nodes_and_weights(grid, nlat)
    lmax = lmax or lmax_default
    mmax = mmax or nlon // 2 + 1
    t = _colatitudes(cost)
    pct = precompute_legpoly(mmax, lmax, t, norm=norm, inverse=True, csphase=csphase)
    return pct, lmax, mmax


def rfft(x, nmodes=None, dim=-1, **kwargs):
    """fme/fft.py:60-76."""
    x = torch.fft.rfft(x, dim=dim, **kwargs)
    if nmodes is not None and nmodes > x.shape[dim]:
        pad = [0, 0] * x.ndim
        d = dim if dim >= 0 else x.ndim + dim
        pad[(x.ndim - 1 - d) * 2 + 1] = nmodes - x.shape[dim]
        x = torch.nn.functional.pad(x, tuple(pad), value=0.0)
    elif nmodes is not None and nmodes < x.shape[dim]:
        x = x.narrow(dim, 0, nmodes)
    return x


def irfft(x, n=None, dim=-1, **kwargs):
    """fme/fft.py:78-96 (operates on a private copy; the reference mutates its input)."""
    if n is None:
        n = 2 * (x.size(dim) - 1)
    x = x.clone()
    x[..., 0].imag = 0.0
    if (n % 2 == 0) and (n // 2 < x.size(dim)):
        x[..., n // 2].imag = 0.0
    return torch.fft.irfft(x, n=n, dim=dim, **kwargs)


class RealSHT(torch.nn.Module):
    """Oracle of fme.sht_fix.RealSHT (fme/sht_fix.py:60-151)."""

    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
        super().__init__()
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        table, self.lmax, self.mmax = forward_table(nlat, nlon, lmax, mmax, grid, norm, csphase)
        self.weights = torch.from_numpy(table).float()
        self._dev_tables = {}

    def _table(self, device):
        # plain attribute like fme/sht_fix.py:117 (which places it on get_device()); cached per device so that the GPU-eager
        # baseline of bench.py does not re-upload it per call
        if device.type == "cpu":
            return self.weights
        if device not in self._dev_tables:
            self._dev_tables[device] = self.weights.to(device)
        return self._dev_tables[device]

    def forward(self, x):
        assert x.shape[-2] == self.nlat and x.shape[-1] == self.nlon
        x = x.float()
        x = 2.0 * torch.pi * rfft(x, nmodes=self.mmax, dim=-1, norm="forward")
        x = torch.view_as_real(x.transpose(-2, -1).contiguous())
        out_shape = list(x.size())
        out_shape[-3] = self.lmax
        out_shape[-2] = self.mmax
        xout = torch.zeros(out_shape, dtype=x.dtype, device=x.device)
        w = self._table(x.device).to(x.dtype)
        xout[..., 0] = torch.einsum("...mk,mlk->...lm", x[..., : self.mmax, :, 0], w)
        xout[..., 1] = torch.einsum("...mk,mlk->...lm", x[..., : self.mmax, :, 1], w)
        return torch.view_as_complex(xout)


class InverseRealSHT(torch.nn.Module):
    """Oracle of fme.sht_fix.InverseRealSHT (fme/sht_fix.py:153-226)."""

    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
        super().__init__()
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        table, self.lmax, self.mmax = inverse_table(nlat, nlon, lmax, mmax, grid, norm, csphase)
        self.pct = torch.from_numpy(table).float()
        self._dev_tables = {}

    def _table(self, device):
        if device.type == "cpu":
            return self.pct
        if device not in self._dev_tables:
            self._dev_tables[device] = self.pct.to(device)
        return self._dev_tables[device]

    def forward(self, x):
        assert x.shape[-2] == self.lmax and x.shape[-1] == self.mmax
        x = torch.view_as_real(x.transpose(-1, -2).contiguous()).float()
        pct = self._table(x.device).to(x.dtype)
        rl = torch.einsum("...ml,mlk->...km", x[..., 0], pct)
        im = torch.einsum("...ml,mlk->...km", x[..., 1], pct)
        x = torch.view_as_complex(torch.stack((rl, im), -1))
        return irfft(x, n=self.nlon, dim=-1, norm="forward")
