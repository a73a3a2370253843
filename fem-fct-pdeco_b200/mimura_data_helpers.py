"""Drop-in for the reference's `mimura_data_helpers.py` (config 3, chemotaxis_mimura_FCT_PGD.py:10,98,182-184,222): the
initial condition and the UFL form builders of the two-species chemotaxis model, on the GPU assembly kernels through the
UFL-like front end (forms.py).  Same names, argument order and return types as mimura_data_helpers.py:19-109, except
that `mat_chtx_p` returns a sparse matrix where the reference's `+ np.zeros(Ad.shape)` densifies it (an n x n dense
array: 2.2 GB at the script's 129^2 mesh) -- same values."""
import numpy as np

from .forms import assemble, assemble_sparse, dot, dx, exp, grad


def m_initial_condition(a1, a2, deltax):
    """mimura_data_helpers.py:19-63: "simplified feathers model" start, np.random.seed(5)"""
    X = np.arange(a1, a2 + deltax, deltax)
    n = X.shape[0]
    np.random.seed(5)
    return 1.5 + 0.1 * (0.5 - np.random.rand(n, n))


def rhs_chtx_m(m_fun, v):
    """mimura_data_helpers.py:65-71: IMEX reaction term m^2 (1 - m) on the right-hand side"""
    return np.asarray(assemble(m_fun ** 2 * (1 - m_fun) * v * dx))


def rhs_chtx_f(f_fun, m_fun, dt, v):
    """mimura_data_helpers.py:73-80"""
    return np.asarray(assemble(f_fun * v * dx + dt * m_fun * v * dx))


def mat_chtx_m(f_fun, m_fun, Dm, chi, u, v):
    """mimura_data_helpers.py:83-101: -Dm K + chi A_a, A_a = exp(-beta m) grad f . grad v u, beta = 0.5"""
    Ad = assemble_sparse(dot(grad(u), grad(v)) * dx)
    beta = 0.5
    Aa = assemble_sparse(exp(-beta * m_fun) * dot(grad(f_fun), grad(v)) * u * dx)
    return - Dm * Ad + chi * Aa


def mat_chtx_p(f_fun, m_fun, Dm, chi, u, v):
    """mimura_data_helpers.py:103-109: -Dm K - chi A_a - chi A_df + A_r with A_a = grad f . grad v u; A_df = div(grad f) u v
    vanishes identically for a P1 field f (dolfin assembles zeros) and A_r is np.zeros in the reference"""
    Ad = assemble_sparse(dot(grad(u), grad(v)) * dx)
    Aa = assemble_sparse(dot(grad(f_fun), grad(v)) * u * dx)
    return - Dm * Ad - chi * Aa
