"""Oracle restatement of the HEALPix spherical harmonic transform (torch CPU).  TEST INFRASTRUCTURE.

Follows the vendored cuHPX Python path of the reference:
  /root/reference/fme/core/cuhpx/sht.py:32-98    SHT   (per-ring rfft + phase shift, Legendre-quadrature einsum)
  /root/reference/fme/core/cuhpx/sht.py:101-153  iSHT  (Legendre einsum, phase shift, per-ring irfft)
  /root/reference/fme/core/cuhpx/tools.py:34-83  healpix_rfft_torch / healpix_irfft_torch
  /root/reference/fme/core/cuhpx/tools.py:178-240 apply_ring_weight / healpix_weights (ring colatitudes, weights)
  /root/reference/fme/core/cuhpx/tools.py:257-286 p2phi_ring / nphi_ring
  /root/reference/fme/core/cuhpx/tools.py:288-336 legpoly -- NOTE its Condon-Shortley line multiplies by +1 (:332-334),
                                                   i.e. the tables carry NO (-1)^m phase, unlike the lat-lon path
The reference's ring loops read the ring count from ``ftm.shape[0]`` (tools.py:40,59), which is the ring axis only for
unbatched 1-D input (the shape its own test uses); this restatement is the per-field transform for any leading dims.
Ring weights ("quad_weights='ring'") come from data files of the reference (healpy's ring weights); the oracle takes the
per-ring weight array as an argument (tests read it from the golden fixtures).
Pinned against the reference's own classes in tests/test_oracle_healpix.py.
"""
import numpy as np
import torch

from .legendre import legpoly


def nphi_ring(t, nside):
    if t < nside - 1:
        return 4 * (t + 1)
    if t <= 3 * nside - 1:
        return 4 * nside
    return 4 * (4 * nside - t - 1)


def phi0_ring(t, nside):
    """Longitude of pixel 0 of ring t (tools.py:257-272 with p = 0)."""
    shift = 0.5
    if nside <= t + 1 <= 3 * nside:
        shift *= (t - nside + 2) % 2
        return np.pi / (2 * nside) * shift
    if t + 1 > 3 * nside:
        return np.pi / (2 * (4 * nside - t - 1)) * shift
    return np.pi / (2 * (t + 1)) * shift


def ring_cos_theta(nside):
    """cos(colatitude) per ring exactly as healpix_weights returns it (tools.py:222-240): z, flipped."""
    t = np.arange(4 * nside - 1)
    z = np.zeros_like(t, dtype=float)
    m1 = t < (nside - 1)
    m2 = (t >= (nside - 1)) & (t <= (3 * nside - 1))
    m3 = (t > (3 * nside - 1)) & (t <= (4 * nside - 2))
    z[m1] = 1 - ((t[m1] + 1) ** 2) / (3 * nside**2)
    z[m2] = 4 / 3 - 2 * (t[m2] + 1) / (3 * nside)
    z[m3] = ((4 * nside - 1 - t[m3]) ** 2) / (3 * nside**2) - 1
    return np.flip(z)


def uniform_weights(nside):
    return 4.0 * np.pi / (12 * nside**2) * np.ones(4 * nside - 1)


def tables(nside, lmax, mmax, weights):
    """(forward [M,L,T] incl. ring weights, inverse [M,L,T]) float64, as SHT.__init__ / iSHT.__init__ build them."""
    cost = ring_cos_theta(nside)
    tq = np.flip(np.arccos(cost))
    x = np.cos(tq)
    # the reference's legpoly applies "*= 1" for csphase (tools.py:332-334): no Condon-Shortley sign
    fwd = legpoly(mmax, lmax, x, norm="ortho", inverse=False, csphase=False) * np.asarray(weights)[None, None, :]
    inv = legpoly(mmax, lmax, x, norm="ortho", inverse=True, csphase=False)
    return fwd, inv


def ring_rfft(f, L, nside):
    lead = f.shape[:-1]
    T = 4 * nside - 1
    ftm = torch.zeros(lead + (T, L), dtype=torch.complex64)
    index = 0
    for t in range(T):
        nphi = nphi_ring(t, nside)
        fm = torch.fft.rfft(f[..., index:index + nphi], norm="backward")
        n = min(nphi // 2 + 1, L)
        ftm[..., t, :n] = fm[..., :n]
        index += nphi
        ftm[..., t, :] *= torch.exp(-1j * torch.arange(L) * phi0_ring(t, nside))
    return ftm


def ring_irfft(ftm, L, nside):
    lead = ftm.shape[:-2]
    T = 4 * nside - 1
    f = torch.zeros(lead + (12 * nside**2,), dtype=torch.float32)
    index = 0
    ftm = ftm.clone()
    for t in range(T):
        ftm[..., t, :] *= torch.exp(1j * torch.arange(L) * phi0_ring(t, nside))
        nphi = nphi_ring(t, nside)
        f[..., index:index + nphi] = torch.fft.irfft(ftm[..., t, :], n=nphi, norm="forward")
        index += nphi
    return f


class SHT:
    def __init__(self, nside, lmax, mmax, weights):
        self.nside, self.lmax, self.mmax = nside, lmax, mmax
        fwd, _ = tables(nside, lmax, mmax, weights)
        self.weights = torch.from_numpy(fwd).float()

    def __call__(self, x):
        x = torch.view_as_real(ring_rfft(x, self.mmax, self.nside))
        re = torch.einsum("...km,mlk->...lm", x[..., : self.mmax, 0], self.weights)
        im = torch.einsum("...km,mlk->...lm", x[..., : self.mmax, 1], self.weights)
        return torch.view_as_complex(torch.stack((re, im), -1).contiguous())


class iSHT:
    def __init__(self, nside, lmax, mmax):
        self.nside, self.lmax, self.mmax = nside, lmax, mmax
        _, inv = tables(nside, lmax, mmax, uniform_weights(nside))
        self.pct = torch.from_numpy(inv).float()

    def __call__(self, x):
        x = torch.view_as_real(x)
        rl = torch.einsum("...lm,mlk->...km", x[..., 0], self.pct)
        im = torch.einsum("...lm,mlk->...km", x[..., 1], self.pct)
        return ring_irfft(torch.view_as_complex(torch.stack((rl, im), -1).contiguous()), self.mmax, self.nside)
