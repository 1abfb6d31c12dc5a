"""TEST INFRASTRUCTURE (oracle).  krylov_exp.apply restated on NumPy vectors.

Follows /root/reference/krylov_exp.py:50-109: shift-and-invert Arnoldi with modified
Gram-Schmidt in the M inner product, per-iteration inv(H)/expm for the residual
estimate, y = beta * sum_j u_j V_j with u from the last iteration at s = 1.
Callbacks: shifted_solve(v) = (M + gamma K')^-1 M v   (solver.solve(), :70-71)
           op(v)            = operator_solver.solve()  (:92-93, residual constant only)
           mdot(a, b)       = assemble(inner(a,b)*dx) = a^T M b  (:55-56,74)
"""
import numpy as np
import scipy.linalg as sla


def krylov_apply(shifted_solve, op, mdot, x, t, gamma, kdim, its=-999, residuals=True):
    norm = lambda v: mdot(v, v) ** 0.5
    beta = norm(x)                                     # :58
    V = [x / beta]                                     # :63-64
    H = np.zeros((kdim + 1, kdim))                     # :45
    if its < 0:
        its = kdim                                     # :66-67
    assert its <= kdim                                 # :68
    res_hist = []
    u = None
    for j in range(its):                               # :69
        w = shifted_solve(V[j])                        # :70-71
        for i in range(j + 1):                         # :73-77
            H[i, j] = mdot(w, V[i])
            w = w - H[i, j] * V[i]
        H[j + 1, j] = norm(w)                          # :78
        V.append(w / H[j + 1, j])                      # :79-80
        Hj = H[:j + 1, :j + 1]                         # :87
        Hinv = sla.inv(Hj)                             # :88
        H1 = (Hinv - np.eye(j + 1)) / gamma            # :89
        e1 = np.zeros(j + 1)
        e1[0] = 1.0
        C = norm(op(w)) if residuals else 0.0          # :92-94 (w already orthogonalised, unnormalised)
        res = np.zeros(3)
        for q, s in enumerate((1. / 3, 2. / 3, 1.)):   # :95-99
            u = sla.expm(-H1 * s * t) @ e1
            res[q] = C / gamma * (Hinv @ u)[j]
        res_hist.append(np.max(np.abs(res)))           # :100
    y = np.zeros_like(x)                               # :102
    for j in range(its):                               # :104-107
        y += u[j] * V[j]
    return beta * y, H, u, np.array(res_hist)          # :109
