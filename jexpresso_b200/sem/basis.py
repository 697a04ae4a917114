"""LGL quadrature and Lagrange basis -- host-side setup inputs of the RHS path.

These arrays (xi, omega, psi, dpsi) are what Jexpresso's ``sem_setup`` hands to
``rhs!`` through ``params.ω`` / ``params.basis.dψ``.  They are *inputs* of the
hot path, generated here so that tests/bench can build synthetic problems
without Julia.  Restated (not copied) from

  * src/kernel/bases/basis_structs.jl:387-485   LegendreGaussLobattoNodesAndWeights!
  * src/kernel/bases/basis_structs.jl:487-564   LegendreAndDerivativeAndQ!
  * src/kernel/bases/basis_structs.jl:596-662   LagrangeInterpolatingPolynomials_classic

Scalar python floats are IEEE binary64, so the scalar recurrences below follow
the reference operation order literally.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = ["legendre_and_derivative_and_q", "lgl_nodes_weights", "lagrange_basis", "build_basis"]


def legendre_and_derivative_and_q(nop: int, x: float):
    """Legendre polynomial L_p, L'_p, q = L_{p+1}-L_{p-1} and q'.

    Follows basis_structs.jl:487-564 (Kopriva alg. 24) including the reference's
    derivative recurrence as written there.
    """
    if nop == 0:
        return 1.0, 0.0, x, 1.0
    if nop == 1:
        return x, 1.0, 0.5 * (3 * x**2 - 2) - 1, 3 * x
    phim2, phim1 = 1.0, x
    dphim2, dphim1 = 0.0, 1.0
    phi = dphi = q = dq = 0.0
    for k in range(2, nop + 1):
        phi = x * phim1 * (2.0 * k - 1.0) / k - phim2 * (k - 1.0) / k
        dphi = dphim2 + (2.0 * k - 1.0) * phim1
        phip1 = x * phi * (2.0 * k + 1.0) / (k + 1) - phim1 * k / (k + 1)
        dphip1 = dphim1 + (2.0 * k - 1.0) * phi
        q = phip1 - phim1
        dq = dphip1 - dphim1
        phim2, phim1 = phim1, phi
        dphim2, dphim1 = dphim1, dphi
    return phi, dphi, q, dq


def lgl_nodes_weights(nop: int):
    """LGL nodes xi[0..nop] and weights omega[0..nop] (basis_structs.jl:387-485)."""
    NITER = 100
    TOL = 4 * np.finfo(np.float64).eps
    xi = np.zeros(nop + 1)
    om = np.ones(nop + 1)
    if nop == 1:
        xi[0], xi[1] = -1.0, 1.0
        om[0] = om[1] = 1.0
    else:
        xi[0], xi[nop] = -1.0, 1.0
        om[0] = om[nop] = 2.0 / (nop * (nop + 1))
        for jj in range(2, (nop + 1) // 2 + 1):  # 1-based jj as in the reference
            j = jj - 1
            xj = -math.cos((j + 0.25) * math.pi / nop - 3.0 / (8.0 * nop * math.pi * (j + 0.25)))
            for _ in range(NITER + 1):
                _, _, q, dq = legendre_and_derivative_and_q(nop, xj)
                delta = -q / dq
                xj = xj + delta
                if abs(delta) <= TOL * abs(xj):
                    break
            L, _, _, _ = legendre_and_derivative_and_q(nop, xj)
            xi[jj - 1] = xj
            xi[nop + 1 - j - 1] = -xj
            L2 = L * L
            om[jj - 1] = 2.0 / (nop * (nop + 1.0) * L2)
            om[nop + 1 - j - 1] = om[jj - 1]
    if nop % 2 == 0:
        L, _, _, _ = legendre_and_derivative_and_q(nop, 0.0)
        xi[nop // 2] = 0.0
        om[nop // 2] = 2.0 / (nop * (nop + 1.0) * (L * L))
    return xi, om


def lagrange_basis(xi: np.ndarray, xiq: np.ndarray):
    """psi[i,l] = L_i(xiq_l), dpsi[i,l] = L_i'(xiq_l)  (basis_structs.jl:596-662)."""
    N = len(xi) - 1
    Q = len(xiq) - 1
    psi = np.zeros((N + 1, Q + 1))
    dpsi = np.zeros((N + 1, Q + 1))
    for l in range(Q + 1):
        xl = float(xiq[l])
        for i in range(N + 1):
            x_i = float(xi[i])
            Lil = 1.0
            dL = 0.0
            for j in range(N + 1):
                xj = float(xi[j])
                if j != i:
                    Lil = Lil * (xl - xj) / (x_i - xj)
                ddL = 1.0
                if j != i:
                    for k in range(N + 1):
                        xk = float(xi[k])
                        if k != i and k != j:
                            ddL = ddL * (xl - xk) / (x_i - xk)
                    dL = dL + ddL / (x_i - xj)
            psi[i, l] = Lil
            dpsi[i, l] = dL
    return psi, dpsi


def build_basis(nop: int):
    """Return dict(xi, omega, psi, dpsi) for inexact integration (Q == N)."""
    xi, om = lgl_nodes_weights(nop)
    psi, dpsi = lagrange_basis(xi, xi)
    return {"nop": nop, "ngl": nop + 1, "xi": xi, "omega": om, "psi": psi, "dpsi": dpsi}
