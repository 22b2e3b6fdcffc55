"""Quadrature nodes/weights on [-1, 1] -- oracle restatement (test infrastructure).

Restates the three init-time routines the reference imports from
torch-harmonics 0.8.0 (``torch_harmonics.quadrature``; pinned in
/root/reference/pyproject.toml:41, call sites /root/reference/fme/sht_fix.py:50,
:92, :95, :98 and :176-182).  The package itself is not vendored in the
reference tree, so these follow the published algorithms:

  * ``legendre_gauss_weights``   -- Gauss-Legendre via numpy's ``leggauss``
  * ``lobatto_weights``          -- Gauss-Lobatto, Newton iteration on the
                                    Legendre three-term recurrence
  * ``clenshaw_curtiss_weights`` -- Clenshaw-Curtis, Waldvogel's FFT formula

All return ``(nodes, weights)`` as float64 numpy arrays, nodes ascending.
Validated through the SHT goldens (``oracle/make_golden.py``): lobatto ->
sht-regression.pt, equiangular + legendre-gauss -> test_sfnonet_output_is_unchanged.pt.
"""
import numpy as np


def _affine(nodes, weights, a, b):
    half = (b - a) * 0.5
    return half * nodes + (b + a) * 0.5, weights * half


def legendre_gauss_weights(n, a=-1.0, b=1.0):
    nodes, weights = np.polynomial.legendre.leggauss(n)
    return _affine(nodes, weights, a, b)


def lobatto_weights(n, a=-1.0, b=1.0, tol=1e-16, maxiter=100):
    # Chebyshev-Gauss-Lobatto points as the first guess, then Newton on
    # q(t) = t P_{n-1}(t) - P_{n-2}(t), whose roots are the Lobatto nodes.
    t = -np.cos(np.pi * np.arange(n) / (n - 1))
    leg = np.zeros((n, n))  # leg[:, k] = P_k(t)
    for _ in range(maxiter):
        t_prev = t
        leg[:, 0] = 1.0
        leg[:, 1] = t
        for k in range(2, n):
            leg[:, k] = ((2 * k - 1) * t * leg[:, k - 1] - (k - 1) * leg[:, k - 2]) / k
        t = t_prev - (t * leg[:, n - 1] - leg[:, n - 2]) / (n * leg[:, n - 1])
        if np.max(np.abs(t - t_prev)) < tol:
            break
    w = 2.0 / ((n * (n - 1)) * leg[:, n - 1] ** 2)
    return _affine(t, w, a, b)


def clenshaw_curtiss_weights(n, a=-1.0, b=1.0):
    assert n > 1
    nodes = np.cos(np.linspace(np.pi, 0, n))
    if n == 2:
        weights = np.array([1.0, 1.0])
    else:
        # Waldvogel (2006), "Fast construction of the Fejer and Clenshaw-Curtis
        # quadrature rules": weights are the inverse FFT of v + g.
        n1 = n - 1
        odd = np.arange(1, n1, 2)
        n_odd = len(odd)
        rest = n1 - n_odd
        v = np.concatenate([2.0 / odd / (odd - 2), 1.0 / odd[-1:], np.zeros(rest)])
        v = 0 - v[:-1] - v[-1:0:-1]
        g0 = -np.ones(n1)
        g0[n_odd] = g0[n_odd] + n1
        g0[rest] = g0[rest] + n1
        g = g0 / (n1**2 - 1 + (n1 % 2))
        weights = np.fft.ifft(v + g).real
        weights = np.concatenate((weights, weights[:1]))
    return _affine(nodes, weights, a, b)


def nodes_and_weights(grid, nlat):
    """Grid dispatch of RealSHT/InverseRealSHT (fme/sht_fix.py:91-104, :175-188).

    Returns (cost, w, default_lmax).
    """
    if grid == "legendre-gauss":
        cost, w = legendre_gauss_weights(nlat, -1, 1)
        return cost, w, nlat
    if grid == "lobatto":
        cost, w = lobatto_weights(nlat, -1, 1)
        return cost, w, nlat - 1
    if grid == "equiangular":
        cost, w = clenshaw_curtiss_weights(nlat, -1, 1)
        return cost, w, nlat
    if grid == "healpix":
        raise NotImplementedError("'healpix' grid not supported by RealSHT")
    raise ValueError("Unknown quadrature mode")
