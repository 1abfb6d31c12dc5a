"""The nonlinear slow right-hand side N(U) = M^-1 L(U) (averaged_sw_explicit.py:177-213).

CPU (not gpu): the NumPy restatement oracle/slow_rhs_ref.py is pinned by construction properties of the form -- Firedrake
itself cannot run here, so parity with the reference's assembled vector is UNPINNED and these properties are the anchor:
  * mass conservation of the depth equation (phi = 1 test function: cell terms vanish, upwind fluxes cancel);
  * the momentum terms converge (second order) to the analytic -(zeta k x u + grad K) of the solid-body flow;
  * the depth terms converge (first order: gradient of a P1 interpolant) to the analytic -u.grad(eta).
GPU: slow_form_kernel against the restatement (<= 1e-12), SlowRHS = M^-1 L against a sparse-LU mass solve, and one
device-resident averaged time step with the CUDA slow right-hand side against the same step with the host restatement.
"""
import numpy as np
import pytest
from scipy.sparse.linalg import splu

from conftest import get_sw

U_MAX = 20.0


def _geometry(sw):
    from krylov_exp_b200.assembly import cell_geometry, duffy_quadrature
    V, C, X, geo = cell_geometry(sw.ref_level, sw.degree, sw.radius, sw.p3_variant)
    qp, qw = duffy_quadrature(6)
    gx, gy = geo.tab_grad(qp)
    J0, J1 = np.einsum('qa,cad->cqd', gx, X), np.einsum('qa,cad->cqd', gy, X)
    xq = np.einsum('qa,cad->cqd', geo.tab(qp), X)
    return qp, qw, J0, J1, xq


def _project_u(sw, field_of_x):
    """rhs_i = int v_i . field dx for a tangent 3-vector field (contravariant Piola: v = J vh / detJ)."""
    from krylov_exp_b200.assembly import BDM2
    qp, qw, J0, J1, xq = _geometry(sw)
    F = field_of_x(xq)
    Phi, _ = BDM2().tab(qp)
    loc = np.einsum('q,cq,qi->ci', qw, (J0 * F).sum(-1), Phi[:, :, 0]) + np.einsum('q,cq,qi->ci', qw, (J1 * F).sum(-1), Phi[:, :, 1])
    out = np.zeros(sw.n_u)
    np.add.at(out, sw.cell_u_dofs.ravel(), (loc * sw.cell_u_sign).ravel())
    return out


def solid_body_state(sw, eta_fn=None):
    """L2 projection of the Williamson-2 zonal flow into BDM2 and a nodal DG1 depth (averaged_sw_explicit.py:60-63)."""
    from krylov_exp_b200.assembly import OMEGA, GRAV
    R = sw.radius
    rhs = _project_u(sw, lambda x: np.stack([-U_MAX * x[..., 1] / R, U_MAX * x[..., 0] / R, 0 * x[..., 0]], axis=-1))
    u = splu(sw.Mu.tocsc()).solve(rhs)
    if eta_fn is None:
        eta_fn = lambda x, y, z: -((R * OMEGA * U_MAX + U_MAX ** 2 / 2.0) * (z * z / (R * R))) / GRAV
    return np.concatenate([u, sw.interpolate_eta(eta_fn)])


def _momentum_error(ref):
    from krylov_exp_b200.slow_rhs import SlowTables
    from oracle.slow_rhs_ref import slow_form
    sw = get_sw(ref)
    U = solid_body_state(sw)
    R = sw.radius

    def tendency(x):                                       # -(zeta k x u + grad K) = -(u_max^2 / 2R) sin(2 lat) e_lat
        lat, lon = np.arcsin(x[..., 2] / R), np.arctan2(x[..., 1], x[..., 0])
        e_lat = np.stack([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)], axis=-1)
        return (-(U_MAX ** 2) / (2 * R) * np.sin(2 * lat))[..., None] * e_lat
    exact = _project_u(sw, tendency)
    L = slow_form(SlowTables(sw), U, None, 1.0)
    lu = splu(sw.Mu.tocsc())
    dn = lambda v: np.sqrt(v @ lu.solve(v))
    return dn(L[:sw.n_u] - exact) / dn(exact), L, sw


def test_momentum_terms_converge_to_the_analytic_tendency():
    e2, L2, sw2 = _momentum_error(2)
    e3, L3, sw3 = _momentum_error(3)
    assert e2 < 2.5e-2 and e3 < 6e-3 and e2 / e3 > 3.5        # second order
    # the steady zonal flow does not change the depth: u . grad(eta) = 0, div u = 0 (up to the discretisation)
    for L, sw in ((L2, sw2), (L3, sw3)):
        assert abs(L[sw.n_u:].sum()) < 1e-9 * np.abs(L[sw.n_u:]).sum()


def _depth_error(ref):
    from krylov_exp_b200.slow_rhs import SlowTables
    from krylov_exp_b200.assembly import Lagrange, _REF_V
    from oracle.slow_rhs_ref import slow_form
    sw = get_sw(ref)
    R, A = sw.radius, 500.0
    U = solid_body_state(sw, lambda x, y, z: A * x / R)        # eta = A cos(lat) cos(lon)
    qp, qw, J0, J1, xq = _geometry(sw)
    detJ = np.linalg.norm(np.cross(J0, J1), axis=-1)
    tend = U_MAX * A * xq[..., 1] / R ** 2                     # -u . grad(eta), u = u_max (-y, x, 0)/R
    rhs = np.einsum('q,cq,cq,qi->ci', qw, detJ, tend, Lagrange(_REF_V.copy(), 1).tab(qp)).ravel()
    L = slow_form(SlowTables(sw), U, None, 1.0)[sw.n_u:]
    Mi = sw.Meta_inv_blocks
    dn = lambda v: np.sqrt(np.einsum('ci,cij,cj->', v.reshape(-1, 3), Mi, v.reshape(-1, 3)))
    return dn(L - rhs) / dn(rhs), L


def test_depth_terms_converge_and_conserve_mass():
    e2, L2 = _depth_error(2)
    e3, L3 = _depth_error(3)
    assert e2 < 7e-2 and e3 < 3.5e-2 and e2 / e3 > 1.8          # first order (gradient of the P1 interpolant)
    for L in (L2, L3):
        assert abs(L.sum()) < 1e-9 * np.abs(L).sum()            # sum_i L(phi_i) = L(1) = 0: upwind fluxes cancel pairwise


def test_topography_and_time_step_scale_linearly_where_they_should():
    """dT scales the whole form; b enters only through eta - b."""
    from krylov_exp_b200.slow_rhs import SlowTables
    from oracle.slow_rhs_ref import slow_form
    sw = get_sw(2)
    T = SlowTables(sw)
    rng = np.random.default_rng(3)
    U = solid_body_state(sw) + np.concatenate([1e-2 * rng.standard_normal(sw.n_u), 5.0 * rng.standard_normal(sw.n_eta)])
    b = 30.0 * rng.standard_normal(sw.n_eta)
    L1 = slow_form(T, U, b, 1.0)
    assert np.allclose(slow_form(T, U, b, 450.0), 450.0 * L1, rtol=1e-12, atol=450.0 * 1e-13 * np.abs(L1).max())
    Ub = U.copy()
    Ub[sw.n_u:] -= b
    assert np.allclose(slow_form(T, Ub, None, 1.0), L1, rtol=1e-12, atol=1e-12 * np.abs(L1).max())


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("ref,B", [(2, 1), (2, 5), (3, 8), (3, 33)])
def test_gpu_slow_form_matches_the_restatement(ref, B):
    import torch
    from krylov_exp_b200.slow_rhs import SlowRHS
    from oracle.slow_rhs_ref import slow_form, slow_solve
    sw = get_sw(ref)
    rng = np.random.default_rng(11)
    b = 100.0 * rng.standard_normal(sw.n_eta)
    dt = 900.0
    slow = SlowRHS.from_sw(sw, dt, b=b, rtol=1e-13, max_it=80)
    base = solid_body_state(sw)
    X = np.stack([base * (1.0 + 0.05 * k) + np.concatenate([0.5 * rng.standard_normal(sw.n_u), 20.0 * rng.standard_normal(sw.n_eta)])
                  for k in range(B)])
    Xd = torch.tensor(X, dtype=torch.float64, device="cuda")
    L = slow.form(Xd).cpu().numpy()
    ref_L = np.stack([slow_form(slow.T, X[k], b, dt) for k in range(B)])
    scale = np.abs(ref_L).max(axis=1, keepdims=True)
    assert np.abs(L - ref_L).max() <= 1e-12 * scale.max(), np.abs((L - ref_L) / scale).max()
    N = slow(Xd)
    torch.cuda.synchronize()
    lu = splu(sw.M().tocsc())
    M = sw.M()
    Nh = N.cpu().numpy()
    for k in range(0, B, max(1, B // 4)):
        r = slow_solve(slow.T, lu.solve, X[k], b, dt)
        d = Nh[k] - r
        assert np.sqrt((d @ (M @ d)) / (r @ (M @ r))) < 1e-10, k
    slow.close()


@pytest.mark.gpu
def test_device_time_step_with_the_cuda_slow_rhs():
    """One rk2 and one rk4 step of the averaged model with EVERYTHING on the device -- exponentials, nonlinear right-hand
    side (CUDA form + batched mass solve), stage algebra -- against the same schemes on host arrays with the NumPy
    restatement of N(.).  No state vector crosses PCIe during the device step."""
    import torch
    import krylov_exp_b200 as kx
    from krylov_exp_b200.slow_rhs import SlowRHS
    from krylov_exp_b200.timestepping import AveragedStepper, DeviceAveragedStepper
    from oracle.slow_rhs_ref import slow_solve
    sw = get_sw(2)
    dt, rho, Mbar = 900.0, 1.0, 5
    eig = 0.0017
    L = eig * dt * rho
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
    cheby = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    cheby2 = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    t, w = kx.sample_times_weights(Mbar, dt, rho)
    slow_dev = SlowRHS.from_sw(sw, dt, rtol=1e-13, max_it=80)
    lu = splu(sw.M().tocsc())
    slow_host = lambda F: np.stack([slow_solve(slow_dev.T, lu.solve, F[k], None, dt) for k in range(F.shape[0])])
    U0 = solid_body_state(sw)
    M = sw.M()
    host = AveragedStepper(cheby, cheby2, slow_host, t, w, dt, fused=True)
    dev = DeviceAveragedStepper(cheby, cheby2, slow_dev, t, w, dt)
    for scheme in ("rk2", "rk4"):
        Uh = getattr(host, scheme)(U0.copy())
        Ud = torch.tensor(U0, dtype=torch.float64, device="cuda")
        getattr(dev, scheme)(Ud.clone())                       # warm-up (allocations, iteration tuning)
        dev.synchronize()
        op.stats_reset()
        getattr(dev, scheme)(Ud)
        dev.synchronize()
        st = op.stats()
        assert st["h2d_vector_bytes"] == 0 and st["d2h_vector_bytes"] == 0
        d = Ud.cpu().numpy() - Uh
        assert np.sqrt((d @ (M @ d)) / (Uh @ (M @ Uh))) < 1e-10, scheme
        # the nonlinear term matters at this size: the step differs from the purely linear propagation
        lin = np.empty_like(U0)
        cheby2.apply(U0, lin, dt)
        assert np.sqrt(((Uh - lin) @ (M @ (Uh - lin))) / (lin @ (M @ lin))) > 1e-6
    dev.close()
    slow_dev.close()
    op.close()
