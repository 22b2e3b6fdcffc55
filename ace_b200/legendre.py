"""Host-side (init-time) constants of the spherical harmonic transform: quadrature and Legendre tables.

Product code (float64 numpy).  Produces the same tables the reference builds at
``/root/reference/fme/sht_fix.py:91-117`` (forward: P_l^m(cos theta_k) * w_k) and ``:175-198``
(inverse: P_l^m(cos theta_k)) with torch-harmonics 0.8.0's ``legendre_gauss_weights`` /
``lobatto_weights`` / ``clenshaw_curtiss_weights`` / ``_precompute_legpoly``; the device library
casts them to fp32 exactly as the reference does.

Unlike the upstream O(L^2) Python double loop (~4 s per table at L = 180) the three-term
recursion here is vectorised over the order m, so a 180x360 plan builds in ~0.1 s and a
721x1440 one in seconds.  Per element the arithmetic and its order are those of the
upstream recursion, so the float64 results are bit-identical to it; ``tests/`` checks that
against the independent loop restatement in ``oracle/legendre.py``.
"""
import numpy as np


def legendre_gauss_nodes(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return x, w


def lobatto_nodes(n, tol=1e-16, maxiter=100):
    """Gauss-Lobatto nodes/weights: Newton iteration on t P_{n-1}(t) - P_{n-2}(t) from Chebyshev points."""
    t = -np.cos(np.pi * np.arange(n) / (n - 1))
    p = np.zeros((n, n))
    for _ in range(maxiter):
        t_old = t
        p[:, 0] = 1.0
        p[:, 1] = t
        for k in range(2, n):
            p[:, k] = ((2 * k - 1) * t * p[:, k - 1] - (k - 1) * p[:, k - 2]) / k
        t = t_old - (t * p[:, n - 1] - p[:, n - 2]) / (n * p[:, n - 1])
        if np.max(np.abs(t - t_old)) < tol:
            break
    w = 2.0 / ((n * (n - 1)) * p[:, n - 1] ** 2)
    return t, w


def clenshaw_curtis_nodes(n):
    """Clenshaw-Curtis nodes cos(pi..0) and weights by the closed-form cosine sum."""
    assert n > 1
    n1 = n - 1
    theta = np.linspace(np.pi, 0, n)
    x = np.cos(theta)
    if n == 2:
        return x, np.array([1.0, 1.0])
    k = np.arange(n)
    w = np.ones(n)
    for j in range(1, n1 // 2 + 1):
        b = 1.0 if 2 * j == n1 else 2.0
        w -= b / (4.0 * j * j - 1.0) * np.cos(2.0 * j * k * np.pi / n1)
    c = np.where((k == 0) | (k == n1), 1.0, 2.0)
    return x, c * w / n1


def grid_nodes(grid, nlat):
    """(cos(theta) ascending, weights, default lmax) for the grids RealSHT accepts (fme/sht_fix.py:91-104)."""
    if grid == "legendre-gauss":
        x, w = legendre_gauss_nodes(nlat)
        return x, w, nlat
    if grid == "lobatto":
        x, w = lobatto_nodes(nlat)
        return x, w, nlat - 1
    if grid == "equiangular":
        x, w = clenshaw_curtis_nodes(nlat)
        return x, w, nlat
    if grid == "healpix":
        raise NotImplementedError("'healpix' grid not supported by RealSHT")
    raise ValueError("Unknown quadrature mode")


def legendre_table(mmax, lmax, x, norm="ortho", inverse=False, csphase=True):
    """float64 [mmax, lmax, len(x)]: normalised associated Legendre functions, zero for l < m."""
    x = np.asarray(x, dtype=np.float64)
    nmax = max(mmax, lmax)
    p = np.zeros((nmax, nmax, x.shape[0]), dtype=np.float64)
    scale = 1.0 if norm == "ortho" else np.sqrt(4 * np.pi)
    if inverse:
        scale = 1.0 / scale
    p[0, 0, :] = scale / np.sqrt(4 * np.pi)
    # diagonal (m = l) and first off-diagonal (m = l - 1): a sequential chain along l
    for l in range(1, nmax):
        p[l - 1, l, :] = np.sqrt(2 * l + 1) * x * p[l - 1, l - 1, :]
        p[l, l, :] = np.sqrt((2 * l + 1) * (1 + x) * (1 - x) / 2 / l) * p[l - 1, l - 1, :]
    # three-term recursion in l, all orders m <= l - 2 at once
    for l in range(2, nmax):
        m = np.arange(0, l - 1, dtype=np.float64)
        a = np.sqrt((2 * l - 1) / (l - m) * (2 * l + 1) / (l + m))
        b = np.sqrt((l + m - 1) / (l - m) * (2 * l + 1) / (2 * l - 3) * (l - m - 1) / (l + m))
        p[: l - 1, l, :] = x[None, :] * a[:, None] * p[: l - 1, l - 1, :] - b[:, None] * p[: l - 1, l - 2, :]
    if norm == "schmidt":
        fac = np.sqrt(2 * np.arange(nmax) + 1.0)
        p = p * fac[None, :, None] if inverse else p / fac[None, :, None]
    p = p[:mmax, :lmax]
    if csphase:
        p[1::2] *= -1.0
    return np.ascontiguousarray(p)


def sht_tables(nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
    """(forward [M,L,K] with quadrature weights, inverse [M,L,K], lmax, mmax) as float64."""
    cost, w, lmax_default = grid_nodes(grid, nlat)
    lmax = lmax or lmax_default
    mmax = mmax or nlon // 2 + 1
    theta = np.flip(np.arccos(cost))  # fme/sht_fix.py:107
    ct = np.cos(theta)
    fwd = legendre_table(mmax, lmax, ct, norm=norm, inverse=False, csphase=csphase) * np.asarray(w)[None, None, :]
    inv = legendre_table(mmax, lmax, ct, norm=norm, inverse=True, csphase=csphase)
    return np.ascontiguousarray(fwd), inv, lmax, mmax
