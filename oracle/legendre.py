"""Orthonormal associated Legendre tables -- oracle restatement (test infrastructure).

Restates ``torch_harmonics.legendre._precompute_legpoly`` / ``legpoly``
(torch-harmonics 0.8.0, pinned /root/reference/pyproject.toml:41; call sites
/root/reference/fme/sht_fix.py:51, :113, :195).  The same recursion is vendored
in the reference at /root/reference/fme/core/cuhpx/tools.py:288-336 (that copy's
Condon-Shortley line is a no-op; upstream multiplies odd m by -1, which the
sht-regression.pt golden confirms).

Deliberately written as the plain (m, l) double loop in float64 so that it is an
independent check of the vectorised product implementation in
``ace_b200/legendre.py``.
"""
import numpy as np


def legpoly(mmax, lmax, x, norm="ortho", inverse=False, csphase=True):
    """P[m, l, k] = normalised associated Legendre function of degree l, order m at x[k].

    float64, shape (mmax, lmax, len(x)); zero for l < m.
    """
    x = np.asarray(x, dtype=np.float64)
    nmax = max(mmax, lmax)
    p = np.zeros((nmax, nmax, len(x)), dtype=np.float64)

    scale = 1.0 if norm == "ortho" else np.sqrt(4 * np.pi)
    if inverse:
        scale = 1.0 / scale
    p[0, 0, :] = scale / np.sqrt(4 * np.pi)

    # sectoral (m = l) and first off-diagonal (m = l - 1) terms
    for l in range(1, nmax):
        p[l - 1, l, :] = np.sqrt(2 * l + 1) * x * p[l - 1, l - 1, :]
        p[l, l, :] = np.sqrt((2 * l + 1) * (1 + x) * (1 - x) / 2 / l) * p[l - 1, l - 1, :]

    # three-term recursion in l for every m <= l - 2
    for l in range(2, nmax):
        for m in range(0, l - 1):
            a = np.sqrt((2 * l - 1) / (l - m) * (2 * l + 1) / (l + m))
            b = np.sqrt((l + m - 1) / (l - m) * (2 * l + 1) / (2 * l - 3) * (l - m - 1) / (l + m))
            p[m, l, :] = x * a * p[m, l - 1, :] - b * p[m, l - 2, :]

    if norm == "schmidt":
        for l in range(0, nmax):
            if inverse:
                p[:, l, :] = p[:, l, :] * np.sqrt(2 * l + 1)
            else:
                p[:, l, :] = p[:, l, :] / np.sqrt(2 * l + 1)

    p = p[:mmax, :lmax]

    if csphase:
        for m in range(1, mmax, 2):
            p[m] *= -1

    return p


def precompute_legpoly(mmax, lmax, t, norm="ortho", inverse=False, csphase=True):
    """``_precompute_legpoly``: tables at colatitudes ``t`` (radians)."""
    return legpoly(mmax, lmax, np.cos(np.asarray(t, dtype=np.float64)), norm=norm, inverse=inverse, csphase=csphase)
