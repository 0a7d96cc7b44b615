"""Composite Simpson quadrature as the reference uses it (TEST INFRASTRUCTURE ONLY).

poc/main.py:179-186 `integra3d` nests three calls of `scipy.integrate.simps(f, x)` (the legacy
name; SciPy removed it in 1.14 - the pinned dependency of the reference's era is SciPy <= 1.10,
whose `simps` has `even='avg'`).  Published algorithm (SciPy 1.10 `_quadrature.py: simps`):
  * odd number of samples N: composite Simpson 1/3 on the N-1 intervals;
  * even N, 'avg': the mean of (Simpson on the first N-1 samples + trapezoid on the last
    interval) and (trapezoid on the first interval + Simpson on the last N-1 samples).
Both are linear in f, so each 1-D rule is a weight vector and the nested 3-D integral equals
the weighted sum with the product weights - that is what the device kernel accumulates.
Pinned here: exactness on cubics (odd N), agreement of weights with the literal algorithm,
and agreement with today's `scipy.integrate.simpson` on odd N (tests/test_oracle_train.py).
"""
import numpy as np


def _simpson_odd(y, h):
    """composite Simpson 1/3 for an odd number of equally spaced samples"""
    return h / 3.0 * (y[0] + y[-1] + 4.0 * y[1:-1:2].sum() + 2.0 * y[2:-1:2].sum())


def simps_avg(y, x):
    """scipy<=1.10 simps(y, x, even='avg') for equally spaced x (what linspace gives the reference)"""
    y = np.asarray(y, np.float64)
    n = y.size
    h = (x[-1] - x[0]) / (n - 1)
    if n % 2 == 1:
        return _simpson_odd(y, h)
    first = _simpson_odd(y[:-1], h) + 0.5 * h * (y[-2] + y[-1])
    last = 0.5 * h * (y[0] + y[1]) + _simpson_odd(y[1:], h)
    return 0.5 * (first + last)


def weights_avg(n, h):
    """weight vector w with simps_avg(y, x) == w @ y"""
    def odd(m):
        w = np.zeros(m)
        w[0] = w[-1] = 1.0
        w[1:-1:2] = 4.0
        w[2:-1:2] = 2.0
        return w * h / 3.0
    if n % 2 == 1:
        return odd(n)
    a = np.zeros(n)
    a[:-1] += odd(n - 1)
    a[-2:] += 0.5 * h
    b = np.zeros(n)
    b[1:] += odd(n - 1)
    b[:2] += 0.5 * h
    return 0.5 * (a + b)


def integra3d(x, y, z, f):
    """poc/main.py:179-186: f[i,j,k] on meshgrid('ij') of x, y, z.  The reference iterates `for fy in f`, i.e. over the
    FIRST axis with the z samples, and integrates the last axis over x - harmless for its cubic grids with equal axes;
    restated literally."""
    return simps_avg([simps_avg([simps_avg(fx, x) for fx in fy], y) for fy in f], z)


def weights_simpson(n, h):
    """scipy >= 1.11 `simpson` (default rule) for equal spacing: for even N the last interval uses the parabola
    through the last three samples (Cartwright).  Offered because SciPy >= 1.14 has no `simps` any more."""
    if n % 2 == 1:
        return weights_avg(n, h)
    w = np.zeros(n)
    w[:-1] += weights_avg(n - 1, h)
    w[-3:] += h * np.array([-1.0 / 12.0, 2.0 / 3.0, 5.0 / 12.0])
    return w
