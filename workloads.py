"""The five BASELINE.json workloads, written once against a module parameter.

Every function takes ``xp`` -- any module exposing the DelayRepay drop-in surface
(``numpy`` itself, the reference package run with ``DELAY_CPU=1``, the oracle
``oracle/refcpu.py`` or this package) -- so the *same source* is evaluated by the
oracle and by the CUDA engine (SURVEY.md section 8d lists the inputs and seeds).

``erf`` is ``scipy.special.erf``: a genuine NumPy ufunc, so every implementation
captures it through ``__array_ufunc__`` (reference: delayarray.py:46-61).
"""
import numpy as _np
from scipy.special import erf as _erf

INV_SQRT2 = 0.7071067811865476      # Python float: stays "weak" under NEP 50


# ----------------------------------------------------------------------------- inputs
def make_inputs(name, n, seed=None):
    """Seeded host inputs for workload ``name`` at problem size ``n``."""
    if name == "axpy":
        rng = _np.random.default_rng(1 if seed is None else seed)
        return dict(a=1.5, x=rng.standard_normal(n), y=rng.standard_normal(n))
    if name == "black_scholes":
        rng = _np.random.default_rng(2 if seed is None else seed)
        return dict(S=rng.uniform(5.0, 30.0, n).astype(_np.float32),
                    K=rng.uniform(1.0, 100.0, n).astype(_np.float32),
                    T=rng.uniform(0.25, 10.0, n).astype(_np.float32))
    if name == "l2":
        rng = _np.random.default_rng(3 if seed is None else seed)
        return dict(a=rng.standard_normal(n), b=rng.standard_normal(n))
    if name == "heat":
        rng = _np.random.default_rng(4 if seed is None else seed)
        return dict(u=rng.random((n, n), dtype=_np.float32))
    if name == "nbody":
        rng = _np.random.default_rng(5 if seed is None else seed)
        return dict(pos=rng.standard_normal((n, 3)).astype(_np.float32),
                    m=rng.uniform(0.5, 1.5, n).astype(_np.float32))
    raise KeyError(name)


# ----------------------------------------------------------------------------- C1
def axpy(xp, a, x, y):
    return a * x + y


# ----------------------------------------------------------------------------- C2
def black_scholes(xp, S, K, T, r=0.02, v=0.30):
    """European call/put; returns the two lazy results (call, put)."""
    sqrt_t = xp.sqrt(T)
    d1 = (xp.log(S / K) + (r + 0.5 * v * v) * T) / (v * sqrt_t)
    d2 = d1 - v * sqrt_t

    def cnd(x):
        return 0.5 * (1.0 + _erf(x * INV_SQRT2))

    disc = K * xp.exp(-r * T)
    call = S * cnd(d1) - disc * cnd(d2)
    put = disc * cnd(-d2) - S * cnd(-d1)
    return call, put


# ----------------------------------------------------------------------------- C3
def l2_distance(xp, a, b):
    return xp.sqrt(xp.sum((a - b) ** 2))


def dot(xp, a, b):
    return xp.dot(a, b)


def norm(xp, a):
    return xp.sqrt(xp.dot(a, a))


# ----------------------------------------------------------------------------- C4
def heat_step(xp, u, c=0.1):
    """One Jacobi step of the 5-point heat stencil, written as slice arithmetic."""
    u[1:-1, 1:-1] = u[1:-1, 1:-1] + c * (
        u[2:, 1:-1] + u[:-2, 1:-1] + u[1:-1, 2:] + u[1:-1, :-2] - 4.0 * u[1:-1, 1:-1])
    return u


def heat(xp, u, steps, c=0.1):
    for _ in range(steps):
        heat_step(xp, u, c)
    return u


# ----------------------------------------------------------------------------- C5
def nbody_acc(xp, pos, m, eps=1e-3):
    """All-pairs acceleration: acc_i = sum_j W_ij (pos_j - pos_i), W = m_j r2^-1.5."""
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    dx = x[None, :] - x[:, None]
    dy = y[None, :] - y[:, None]
    dz = z[None, :] - z[:, None]
    r2 = dx ** 2 + dy ** 2 + dz ** 2 + eps
    w = m[None, :] * r2 ** -1.5
    return w @ pos - pos * w.sum(1)[:, None]
