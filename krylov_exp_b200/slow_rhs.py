"""The nonlinear slow right-hand side N(U) of the averaged model on the GPU.

Reference: /root/reference/averaged_sw_explicit.py:177-213 -- `SlowSolver` solves  a(w, test) = L(U; test)  with the
mass form `a` and the vector-invariant, upwinded nonlinear 1-form

    L = dT [  <perp(grad(v . perp(u))), u> dx  -  <both(perp(n) (v . perp(u))), both(Upwind u)> dS  +  div(v) K dx
            + <grad(phi), u> (eta - b) dx      -  jump(phi) (uup+ (eta+ - b+) - uup- (eta- - b-)) dS ]

    perp(u) = k x u (k = DG interpolant of the cell normal, :49-50), K = |u|^2 / 2, Upwind = (sign(u.n) + 1)/2,
    uup = (u.n + |u.n|)/2, both(X) = X+ + X-                                                   (:183-201)

i.e. N(U) = M^-1 L(U).  Here: `SlowTables` is host-side input generation in the role of Firedrake's form compiler and
mesh (per-cell geometry factors of the BDM2 x DG1 pair on the cubic icosahedral sphere at the quadrature points, basis
tabulations, facet neighbours), exactly like krylov_exp_b200.assembly is for K and M; `SlowRHS` evaluates L(U) for all
ensemble members with the CUDA kernel `slow_form_kernel` (csrc/kexp_slow.cu, one warp = one cell x 32 members) and
applies M^-1 with the library's batched mass solve, all on the library's stream with device-resident (B, n) arrays --
the `slow` callback of DeviceAveragedStepper.  The NumPy restatement used as the checker lives with the test infrastructure (slow_rhs_ref.py).

Everything is evaluated in reference coordinates.  With u = J uh / detJ (contravariant Piola), G = J^T J, m = k.(J_0 x J_1),
mu = m / detJ^2, a x b = a_0 b_1 - a_1 b_0 and uh_perp = (-uh_1, uh_0):
    v . perp(u)            = -mu (vh x uh)
    <perp(grad s), u> dx   = -(grad_h s)^T G^-1 (m uh_perp) d(xi)
    div(v) K dx            = divh(vh) (uh^T G uh) / (2 detJ^2) d(xi)
    <grad phi, u> D dx     = (grad_h phi . uh) D d(xi)
and on a facet with reference normal nh: n = J G^-1 nh / |.|, dS = |J th| ds, u.n from the 3-vectors of either side.
"""
import ctypes as C

import numpy as np

from .assembly import BDM2, Lagrange, cell_geometry, duffy_quadrature, _LOCAL_EDGES, _REF_V


class SlowTables:
    """Geometry / basis tables of the nonlinear form for one mesh (host, once)."""

    NCELL = 11       # doubles per (cell, quadrature point):  Ginv00 Ginv01 Ginv11 G00 G01 G11 w*m mu dmu0 dmu1 w/(2 detJ^2)
    NFAC = 14        # doubles per (cell, edge, facet point):  Jd0[3] Jd1[3] n[3] P[3] ds mu

    def __init__(self, sw, nquad=6, nfq=6):
        V, Cc, X, geo = cell_geometry(sw.ref_level, sw.degree, sw.radius, sw.p3_variant)
        nc = Cc.shape[0]
        self.n_cells, self.n_u, self.n_eta = nc, sw.n_u, sw.n_eta
        self.cell_u_dofs = np.ascontiguousarray(sw.cell_u_dofs, dtype=np.int32)
        self.cell_u_sign = np.ascontiguousarray(sw.cell_u_sign, dtype=np.float64)
        bdm, dg = BDM2(), Lagrange(_REF_V.copy(), 1)
        cross = np.cross

        # DG(degree) interpolant of the unit cell normal (averaged_sw_explicit.py:49): nodal values at the geometry nodes
        gxn, gyn = geo.tab_grad(geo.nodes)
        J0n, J1n = np.einsum('na,cad->cnd', gxn, X), np.einsum('na,cad->cnd', gyn, X)
        nun = cross(J0n, J1n)
        knode = nun / np.linalg.norm(nun, axis=2, keepdims=True)
        assert np.all(np.einsum('cnd,cnd->cn', knode, X) > 0), "cells must be oriented outward"

        def geom(pts):
            """J columns, their reference derivatives, interpolated normal and its derivatives at reference points."""
            N = geo.tab(pts)
            gx, gy = geo.tab_grad(pts)
            hxx, hxy, hyy = geo.tab_hess(pts)
            e = lambda T, A: np.einsum('qa,cad->cqd', T, A)
            return dict(J0=e(gx, X), J1=e(gy, X), dJ0=(e(hxx, X), e(hxy, X)), dJ1=(e(hxy, X), e(hyy, X)),
                        k=e(N, knode), dk=(e(gx, knode), e(gy, knode)))

        def metric(g):
            J0, J1 = g["J0"], g["J1"]
            G00, G01, G11 = (J0 * J0).sum(-1), (J0 * J1).sum(-1), (J1 * J1).sum(-1)
            det2 = G00 * G11 - G01 ** 2
            cr = cross(J0, J1)
            return G00, G01, G11, det2, np.sqrt(det2), cr, (g["k"] * cr).sum(-1)

        # ---- cell quadrature
        qp, qw = duffy_quadrature(nquad)
        self.nq = qp.shape[0]
        g = geom(qp)
        G00, G01, G11, det2, detJ, cr, m = metric(g)
        nu = cr / detJ[..., None]
        dmu = []
        for b in range(2):
            dcr = cross(g["dJ0"][b], g["J1"]) + cross(g["J0"], g["dJ1"][b])
            ddet = (nu * dcr).sum(-1)
            dm = (g["dk"][b] * cr).sum(-1) + (g["k"] * dcr).sum(-1)
            dmu.append(dm / det2 - 2.0 * m * ddet / (det2 * detJ))
        cell = np.empty((nc, self.nq, self.NCELL))
        cell[..., 0], cell[..., 1], cell[..., 2] = G11 / det2, -G01 / det2, G00 / det2
        cell[..., 3], cell[..., 4], cell[..., 5] = G00, G01, G11
        cell[..., 6] = qw[None, :] * m
        cell[..., 7] = m / det2
        cell[..., 8], cell[..., 9] = dmu
        cell[..., 10] = 0.5 * qw[None, :] / det2
        self.cell = np.ascontiguousarray(cell)
        self.qw = np.ascontiguousarray(qw)
        Phi, dPhi = bdm.tab(qp)
        self.Phi = np.ascontiguousarray(Phi)                         # (nq,12,2)
        self.dPhi = np.ascontiguousarray(dPhi)                       # (nq,12)    reference divergence
        self.gPhi = np.ascontiguousarray(bdm.tab_grad(qp))           # (nq,12,2,2) [q,j,a,b] = d Phi_j,a / d xi_b
        self.Psi = np.ascontiguousarray(dg.tab(qp))                  # (nq,3)
        gpx, gpy = dg.tab_grad(qp[:1])
        self.gPsi = np.ascontiguousarray(np.stack([gpx[0], gpy[0]], axis=1))   # (3,2) constant P1 gradients

        # ---- facets: Gauss points on each local edge, counter-clockwise; the neighbour meets them in reverse order
        from scipy.special import roots_legendre
        xs, ws = roots_legendre(nfq)
        sf, wf = 0.5 * (xs + 1.0), 0.5 * ws
        self.nfq = nfq
        fac = np.empty((nc, 3, nfq, self.NFAC))
        PhiF = np.empty((3, nfq, 12, 2))
        PsiF = np.empty((3, nfq, 3))
        for e, (p, q) in enumerate(_LOCAL_EDGES):
            th = _REF_V[q] - _REF_V[p]
            nh = np.array([th[1], -th[0]])
            pts = _REF_V[p][None, :] + sf[:, None] * th[None, :]
            PhiF[e], _ = bdm.tab(pts)
            PsiF[e] = dg.tab(pts)
            g = geom(pts)
            G00, G01, G11, det2, detJ, cr, m = metric(g)
            v0 = (G11 * nh[0] - G01 * nh[1]) / det2                  # G^-1 nh
            v1 = (-G01 * nh[0] + G00 * nh[1]) / det2
            nph = g["J0"] * v0[..., None] + g["J1"] * v1[..., None]
            nph /= np.linalg.norm(nph, axis=-1, keepdims=True)
            tph = g["J0"] * th[0] + g["J1"] * th[1]
            fac[:, e, :, 0:3] = g["J0"] / detJ[..., None]
            fac[:, e, :, 3:6] = g["J1"] / detJ[..., None]
            fac[:, e, :, 6:9] = nph
            fac[:, e, :, 9:12] = cross(g["k"], nph)                   # perp(n)
            fac[:, e, :, 12] = np.linalg.norm(tph, axis=-1) * wf[None, :]
            fac[:, e, :, 13] = m / det2
        self.fac = np.ascontiguousarray(fac)
        self.PhiF, self.PsiF = np.ascontiguousarray(PhiF), np.ascontiguousarray(PsiF)

        # ---- facet neighbours: cell and local edge on the other side of each local edge
        owner = {}
        nbr_cell = np.empty((nc, 3), dtype=np.int32)
        nbr_edge = np.empty((nc, 3), dtype=np.int32)
        for ci in range(nc):
            for e, (p, q) in enumerate(_LOCAL_EDGES):
                owner[(int(Cc[ci, p]), int(Cc[ci, q]))] = (ci, e)
        for (a, b), (ci, e) in owner.items():
            nbr_cell[ci, e], nbr_edge[ci, e] = owner[(b, a)]          # closed, consistently oriented surface
        self.nbr_cell, self.nbr_edge = nbr_cell, nbr_edge


class SlowRHS:
    """N(.) = M^-1 L(.) for a batch of ensemble members on the device: the `slow` callback of DeviceAveragedStepper.

    op_mass: a krylov_exp_b200.Operator built with K = identity (so that its batched operator application is M^-1);
    `from_sw` builds it.  dt: the dT of the form (averaged_sw_explicit.py:190).  b: DG1 nodal values of the topography."""

    def __init__(self, tables, op_mass, dt, b=None):
        import torch
        self.torch, self.T, self.op, self.dt = torch, tables, op_mass, float(dt)
        self.lib = op_mass.lib
        dev = torch.device("cuda", torch.cuda.current_device())
        up = lambda a, dt_=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt_, device=dev)
        T = tables
        self.n = T.n_u + T.n_eta
        if op_mass.n != self.n:
            raise ValueError("mass operator and tables belong to different spaces")
        self._dev = dict(cell=up(T.cell), fac=up(T.fac), Phi=up(T.Phi), dPhi=up(T.dPhi), gPhi=up(T.gPhi), Psi=up(T.Psi), gPsi=up(T.gPsi),
                         PhiF=up(T.PhiF), PsiF=up(T.PsiF), qw=up(T.qw), sign=up(T.cell_u_sign),
                         dofs=up(T.cell_u_dofs, torch.int32), nbr_cell=up(T.nbr_cell, torch.int32), nbr_edge=up(T.nbr_edge, torch.int32),
                         b=up(np.zeros(T.n_eta) if b is None else b))
        self.stream = torch.cuda.ExternalStream(op_mass.stream_ptr())
        self._batches = {}
        op_mass.set_async(True)                            # mass solves queue without host joins; checked in synchronize()

    @classmethod
    def from_sw(cls, sw, dt, b=None, device=0, rtol=1e-12, max_it=60, **kw):
        """Tables and the M^-1 operator (K = I) from the host assembler's output."""
        import scipy.sparse as sp
        from .operator import Operator
        op_mass = Operator(sp.identity(sw.n, format="csr"), sw.Mu, sw.Meta_inv_blocks, device=device, rtol=rtol, max_it=max_it)
        return cls(SlowTables(sw, **kw), op_mass, dt, b=b)

    def form(self, U, out=None):
        """L(U) for the rows of the device tensor U (B, n) -> (B, n); enqueued on the current torch stream."""
        torch = self.torch
        B = U.shape[0]
        if U.dtype != torch.float64 or not U.is_cuda or U.shape[1] != self.n or U.stride(1) != 1:
            raise ValueError(f"U must be a CUDA float64 tensor of shape (B, {self.n}) with unit inner stride")
        if out is None:
            out = torch.empty((B, self.n), dtype=torch.float64, device=U.device)
        d, T = self._dev, self.T
        p = lambda t: C.c_void_p(t.data_ptr())
        rc = self.lib.kexp_slow_form_dev(C.c_void_p(torch.cuda.current_stream().cuda_stream), T.n_cells, T.n_u, T.n_eta, T.nq, T.nfq,
                                         p(d["cell"]), p(d["fac"]), p(d["Phi"]), p(d["dPhi"]), p(d["gPhi"]), p(d["Psi"]), p(d["gPsi"]),
                                         p(d["PhiF"]), p(d["PsiF"]), p(d["qw"]), p(d["sign"]), p(d["dofs"]), p(d["nbr_cell"]), p(d["nbr_edge"]),
                                         p(d["b"]), p(U), U.stride(0) if B > 1 else self.n, p(out), out.stride(0) if B > 1 else self.n, B,
                                         C.c_double(self.dt))
        if rc != 0:
            raise RuntimeError(f"kexp_slow_form_dev failed ({rc})")
        return out

    def __call__(self, F):
        """SlowSolver.solve() for every member: rows of F (B, n) -> N(F) (B, n), on the library's stream."""
        torch = self.torch
        B = F.shape[0]
        cur = torch.cuda.current_stream()
        # F is produced on the caller's stream and the result is consumed there: device-side waits on both sides order the two
        # streams (no host join).  The result is allocated on the caller's stream; no tensor is tied to the library's stream,
        # which may be destroyed (Operator.close) before the tensors are.
        L = torch.empty((B, self.n), dtype=torch.float64, device=F.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.form(F, out=L)
            if B not in self._batches:
                self._batches[B] = (self.op.batch(B), self.op.batch(B))
            bi, bo = self._batches[B]
            bi.upload_dev(L)
            self.op._check(self.lib.kexp_operator_apply_dev(self.op.ctx, bi.handle, bo.handle), "kexp_operator_apply_dev")
            bo.export_dev(L)
        cur.wait_stream(self.stream)
        return L

    def synchronize(self):
        """Joins the mass-solve stream and raises if one of the queued mass solves missed its tolerance."""
        self.op.synchronize()

    def close(self):
        for bi, bo in self._batches.values():
            bi.close(); bo.close()
        self._batches = {}
        self.op.close()
