"""Batched averaged time-steppers (krylov_exp_b200/timestepping.py) against golden vectors produced by the
reference's own timestepping_methods.py run as 5 threaded ensemble members (tests/golden/make_golden_timestepping.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT

SCHEMES = ("rk2", "rk4", "heuns", "ssprk3", "leapfrog")


def _load():
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    ts = np.load(os.path.join(ROOT, "tests", "golden", "timestepping_golden.npz"))
    nu, ne, eig = g["op_meta"]
    nu, ne = int(nu), int(ne)
    n = nu + ne
    K = sp.csr_matrix((g["op_K_data"], g["op_K_indices"], g["op_K_indptr"]), shape=(n, n))
    M = sp.csr_matrix((g["op_M_data"], g["op_M_indices"], g["op_M_indptr"]), shape=(n, n))
    return g, ts, K, M, nu, ne


def _slow(scale):
    return lambda F: 1e-2 * F * np.roll(F, 1, axis=-1) / scale


class OracleCheby:
    """The batched interface of krylov_exp_b200.cheby_exp on top of the CPU oracle (host-logic test)."""

    def __init__(self, A, L):
        from oracle import coeffs
        c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
        self.a, self.L, self.A = coeffs.real_recurrence_coeffs(c, n), L, A

    def apply(self, x, y, t):
        from oracle import cheby_ref
        y[:] = cheby_ref.apply_real(self.a, self.L, lambda v: self.A @ v, np.asarray(x), t)

    def apply_batch(self, X, t, shared=False):
        from oracle import cheby_ref
        X = np.repeat(np.asarray(X)[None, :], len(t), axis=0) if shared else np.asarray(X)
        return cheby_ref.apply_batch(self.a, self.L, lambda V: self.A @ V, X.T, np.asarray(t)).T

    def apply_average(self, X, t, w, shared=False, fused=False):
        return np.asarray(w) @ self.apply_batch(X, t, shared)


def _run(stepper, ts, scheme):
    U = ts["U0"].copy()
    if scheme == "leapfrog":
        U_old = 0.97 * ts["U0"]
        stepper.leapfrog(U, U_old, float(ts["meta"][4]))
        return U, U_old
    getattr(stepper, scheme)(U)
    return U, None


@pytest.mark.parametrize("scheme", SCHEMES)
def test_batched_formulation_matches_reference_ensemble(scheme):
    import krylov_exp_b200 as kx
    from krylov_exp_b200.timestepping import AveragedStepper
    g, ts, K, M, nu, ne = _load()
    Mbar, dt, rho, Lc, _ = ts["meta"]
    A = np.linalg.solve(M.toarray(), K.toarray())
    t, w = kx.sample_times_weights(int(Mbar), dt, rho)
    ch = OracleCheby(A, Lc)
    st = AveragedStepper(ch, ch, _slow(np.abs(ts["U0"]).max()), t, w, dt)
    U, U_old = _run(st, ts, scheme)
    ref = ts[f"{scheme}_U"]
    assert np.linalg.norm(U - ref) <= 1e-11 * np.linalg.norm(ref)
    if U_old is not None:
        assert np.linalg.norm(U_old - ts["leapfrog_Uold"]) <= 1e-11 * np.linalg.norm(ts["leapfrog_Uold"])


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("scheme", SCHEMES)
def test_gpu_time_step_matches_reference_ensemble(scheme, fused):
    """fused=True: backward half of every averaged evaluation as one Clenshaw recurrence; False: Mbar applications."""
    import krylov_exp_b200 as kx
    from krylov_exp_b200.timestepping import AveragedStepper
    g, ts, K, M, nu, ne = _load()
    Mbar, dt, rho, Lc, _ = ts["meta"]
    Mu = M[:nu][:, :nu].tocsr()
    Meta = M[nu:][:, nu:].tobsr(blocksize=(3, 3)).data
    op = kx.Operator(K, Mu, np.linalg.inv(Meta), rtol=1e-13, max_it=100)
    cheby = kx.cheby_exp(op, None, None, 10000, 1e-8, Lc)
    cheby2 = kx.cheby_exp(op, None, None, 10000, 1e-8, Lc)
    t, w = kx.sample_times_weights(int(Mbar), dt, rho)
    st = AveragedStepper(cheby, cheby2, _slow(np.abs(ts["U0"]).max()), t, w, dt, fused=fused)
    U, U_old = _run(st, ts, scheme)
    ref = ts[f"{scheme}_U"]
    d = U - ref
    assert np.sqrt((d @ (M @ d)) / (ref @ (M @ ref))) < 1e-10
    if U_old is not None:
        d = U_old - ts["leapfrog_Uold"]
        assert np.sqrt((d @ (M @ d)) / (ts["leapfrog_Uold"] @ (M @ ts["leapfrog_Uold"]))) < 1e-10
    op.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", SCHEMES)
def test_device_resident_time_step(scheme):
    """DeviceAveragedStepper: state, stages, ensemble batches and cheby2 results stay in HBM, `slow` is a device
    callback (torch ops on the library's stream).  Matches the reference ensemble's golden step and copies no state
    vector between host and device during the step."""
    import torch
    import krylov_exp_b200 as kx
    from krylov_exp_b200.timestepping import DeviceAveragedStepper
    g, ts, K, M, nu, ne = _load()
    Mbar, dt, rho, Lc, asselin = ts["meta"]
    Mu = M[:nu][:, :nu].tocsr()
    Meta = M[nu:][:, nu:].tobsr(blocksize=(3, 3)).data
    op = kx.Operator(K, Mu, np.linalg.inv(Meta), rtol=1e-13, max_it=100)
    cheby = kx.cheby_exp(op, None, None, 10000, 1e-8, Lc)
    cheby2 = kx.cheby_exp(op, None, None, 10000, 1e-8, Lc)
    t, w = kx.sample_times_weights(int(Mbar), dt, rho)
    scale = float(np.abs(ts["U0"]).max())
    slow_dev = lambda F: 1e-2 * F * torch.roll(F, 1, dims=-1) / scale        # the golden's synthetic N(.) on the device
    st = DeviceAveragedStepper(cheby, cheby2, slow_dev, t, w, dt)
    U = torch.tensor(ts["U0"], dtype=torch.float64, device="cuda")
    U_old = torch.tensor(0.97 * ts["U0"], dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    if scheme == "leapfrog":                                 # warm-up step on copies (allocations, iteration tuning)
        st.leapfrog(U.clone(), U_old.clone(), float(asselin))
    else:
        getattr(st, scheme)(U.clone())
    st.synchronize()
    op.stats_reset()
    if scheme == "leapfrog":
        st.leapfrog(U, U_old, float(asselin))
    else:
        getattr(st, scheme)(U)
    st.synchronize()
    stats = op.stats()
    assert stats["h2d_vector_bytes"] == 0 and stats["d2h_vector_bytes"] == 0
    Uh = U.cpu().numpy()
    ref = ts[f"{scheme}_U"]
    d = Uh - ref
    assert np.sqrt((d @ (M @ d)) / (ref @ (M @ ref))) < 1e-10
    if scheme == "leapfrog":
        d = U_old.cpu().numpy() - ts["leapfrog_Uold"]
        assert np.sqrt((d @ (M @ d)) / (ts["leapfrog_Uold"] @ (M @ ts["leapfrog_Uold"]))) < 1e-10
    st.close()
    op.close()
