"""TEST INFRASTRUCTURE (oracle).  NumPy restatement of the nonlinear slow right-hand side.

Follows /root/reference/averaged_sw_explicit.py:177-213 term by term (vector_invariant = True, :192-201):

    L = dT [ inner(perp(grad(inner(v, perp(u0)))), u0) dx                      (T1)
           - inner(both(perp(n) inner(v, perp(u0))), both(Upwind u0)) dS       (T2)
           + div(v) K dx                                                        (T3)
           + inner(grad(phi), u0 (eta0 - b)) dx                                 (T4)
           - jump(phi) (uup('+')(eta0('+') - b('+')) - uup('-')(eta0('-') - b('-'))) dS ]   (T5)

with perp(u) = cross(outward_normals, u) (:49-50), K = 0.5 inner(u0, u0) (:186), Upwind = 0.5 (sign(dot(u0, n)) + 1)
(:184), uup = 0.5 (dot(u0, n) + abs(dot(u0, n))) (:187), both(u) = 2 avg(u) (:185), and SlowSolver.solve() = M^-1 L (:203-205).

The form is evaluated from the geometry / basis tables of krylov_exp_b200.slow_rhs.SlowTables (input generation in the
role of Firedrake's mesh and form compiler, shared with the GPU path like the assembled K and M are), cell by cell in
vectorised NumPy -- straightforward quadrature loops, no attempt at speed.  PARITY UNPINNED against Firedrake itself
(not installable here); pinned by construction properties in tests/test_slow_rhs.py: mass conservation, the K-gradient
identity, the steady Williamson-2 state, and convergence under refinement.
"""
import numpy as np


def _cross2(a, b):
    return a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]


def slow_form(T, U, b=None, dT=1.0):
    """L(U) (n,) for one state U = [u ; eta] (n,)."""
    nc, nu = T.n_cells, T.n_u
    b = np.zeros(T.n_eta) if b is None else np.asarray(b, dtype=np.float64)
    Uc = U[T.cell_u_dofs] * T.cell_u_sign                    # (C,12) signed local BDM coefficients
    Ec = U[nu:].reshape(nc, 3)                               # (C,3) DG1 nodal depth
    Dc = Ec - b.reshape(nc, 3)                               # eta0 - b
    eu = np.zeros((nc, 12))
    ep = np.zeros((nc, 3))

    # ---------------- cell integrals
    for q in range(T.nq):
        Phi, gPhi, dPhi, Psi = T.Phi[q], T.gPhi[q], T.dPhi[q], T.Psi[q]
        t = T.cell[:, q, :]
        gi00, gi01, gi11, g00, g01, g11, wm, mu, dmu0, dmu1, ke = (t[:, i] for i in range(11))
        uh = Uc @ Phi                                         # (C,2)
        duh = np.einsum('cj,jab->cab', Uc, gPhi)              # (C,2,2): d uh_a / d xi_b
        uperp = np.stack([-uh[:, 1], uh[:, 0]], axis=1)
        h0 = wm * (gi00 * uperp[:, 0] + gi01 * uperp[:, 1])   # w m G^-1 uh_perp
        h1 = wm * (gi01 * uperp[:, 0] + gi11 * uperp[:, 1])
        kin = ke * (g00 * uh[:, 0] ** 2 + 2.0 * g01 * uh[:, 0] * uh[:, 1] + g11 * uh[:, 1] ** 2)   # w K / detJ
        Dq = Dc @ Psi                                         # (C,)
        for i in range(12):
            cx = Phi[i, 0] * uh[:, 1] - Phi[i, 1] * uh[:, 0]                         # Phi_i x uh
            t1 = 0.0
            for bb, (hb, dmub) in enumerate(((h0, dmu0), (h1, dmu1))):
                dcx = (gPhi[i, 0, bb] * uh[:, 1] - gPhi[i, 1, bb] * uh[:, 0]) + (Phi[i, 0] * duh[:, 1, bb] - Phi[i, 1] * duh[:, 0, bb])
                t1 = t1 + hb * (dmub * cx + mu * dcx)                                # -(grad s)_b h_b, s = -mu (Phi_i x uh)
            eu[:, i] += t1 + dPhi[i] * kin                                           # T1 + T3
        for i in range(3):
            ep[:, i] += T.qw[q] * (T.gPsi[i, 0] * uh[:, 0] + T.gPsi[i, 1] * uh[:, 1]) * Dq       # T4

    # ---------------- interior facets: own trace and the neighbour's at the same physical points
    nf = T.nfq
    for e in range(3):
        nb, ne = T.nbr_cell[:, e], T.nbr_edge[:, e]
        for f in range(nf):
            fo = T.fac[:, e, f, :]
            fn = T.fac[nb, ne, nf - 1 - f, :]                                        # the neighbour runs along the edge the other way
            PhiO, PsiO = T.PhiF[e, f], T.PsiF[e, f]
            uhO = Uc @ PhiO                                                          # (C,2)
            uhN = np.einsum('cj,cja->ca', Uc[nb], T.PhiF[ne, nf - 1 - f])            # (C,2) neighbour's reference trace
            u3O = fo[:, 0:3] * uhO[:, 0:1] + fo[:, 3:6] * uhO[:, 1:2]                # u0 as a 3-vector, own side
            u3N = fn[:, 0:3] * uhN[:, 0:1] + fn[:, 3:6] * uhN[:, 1:2]
            unO = (u3O * fo[:, 6:9]).sum(1)                                          # dot(u0, n), each side with its own n
            unN = (u3N * fn[:, 6:9]).sum(1)
            upO, upN = 0.5 * (np.sign(unO) + 1.0), 0.5 * (np.sign(unN) + 1.0)       # Upwind (:184)
            Yv = upO[:, None] * u3O + upN[:, None] * u3N                             # both(Upwind u0)
            PY = (fo[:, 9:12] * Yv).sum(1)                                           # perp(n) . both(Upwind u0)
            ds, muf = fo[:, 12], fo[:, 13]
            DO = Dc @ PsiO
            DN = np.einsum('cj,cj->c', Dc[nb], T.PsiF[ne, nf - 1 - f])
            flux = 0.5 * (unO + np.abs(unO)) * DO - 0.5 * (unN + np.abs(unN)) * DN   # uup+ D+ - uup- D-  seen from this cell
            for i in range(12):
                cx = PhiO[i, 0] * uhO[:, 1] - PhiO[i, 1] * uhO[:, 0]
                eu[:, i] += ds * muf * cx * PY                                       # -T2: s_i = -mu (Phi_i x uh)
            for i in range(3):
                ep[:, i] -= ds * PsiO[i] * flux                                      # -T5

    L = np.zeros_like(U)
    np.add.at(L, T.cell_u_dofs.ravel(), (dT * eu * T.cell_u_sign).ravel())
    L[nu:] = dT * ep.ravel()
    return L


def slow_solve(T, Msolve, U, b=None, dT=1.0):
    """SlowSolver.solve(): USlow_out = M^-1 L(USlow_in)   (:203-205).  Msolve(v) = M^-1 v."""
    return Msolve(slow_form(T, U, b, dT))
