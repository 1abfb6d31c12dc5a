"""Export K and M from a Firedrake LinearVariationalSolver as CSR (host, once).

Only usable where Firedrake is installed; nothing in this repository's tests can run it
(no Firedrake in the build image).  It follows the reference's own precedent for pulling
the operators out as PETSc AIJ matrices, /root/reference/L_eigs.py:44-45:
    petsc_a = assemble(a, mat_type='aij').M.handle
For the solver built at averaged_sw_explicit.py:163-164, problem.J is the mass form `a`
and problem.F = a(operator_out) - F(operator_in), so K = d(-F_residual)/d(operator_in).
"""
import numpy as np
import scipy.sparse as sp

from .operator import Operator

_CACHE = {}


def _petsc_to_csr(mat):
    indptr, indices, data = mat.getValuesCSR()
    return sp.csr_matrix((data, indices, indptr), shape=mat.getSize())


def operator_from_firedrake(operator_solver, operator_in, device=0, **kw):
    key = id(operator_solver)
    if key in _CACHE:
        return _CACHE[key]
    import firedrake as fd                                  # noqa: F401 (fails loudly without Firedrake)
    problem = operator_solver._problem
    W = operator_in.function_space()
    V1 = W.sub(0)
    n_u = V1.dof_count
    # residual F(u_out; operator_in) = a(u_out, v) - rhs(operator_in; v); K = -dF/d(operator_in)
    Kform = -fd.derivative(problem.F, operator_in)
    K = _petsc_to_csr(fd.assemble(Kform, mat_type="aij").M.handle)
    M = _petsc_to_csr(fd.assemble(problem.J, mat_type="aij").M.handle)
    Mu = M[:n_u][:, :n_u].tocsr()
    Meta = M[n_u:][:, n_u:].tobsr(blocksize=(3, 3))
    nb = Meta.shape[0] // 3
    if Meta.indices.shape[0] != nb or np.any(Meta.indices != np.arange(nb)):
        raise ValueError("DG1 mass matrix is not 3x3 block diagonal in this dof order")
    op = Operator(K, Mu, np.linalg.inv(Meta.data), device=device, **kw)
    _CACHE[key] = op
    return op


def shifted_matrices_from_firedrake(solver, solver_in, gamma):
    """(K, Mu, Meta_blocks) from the shift-and-invert solver of test_krylov.py:34-62: problem.J is the bilinear form
    a = M + gamma K' (K' = -K in this package's sign: a = <u,v> - gamma <f perp(u), v> + gamma c h div v + ... ), and the
    residual problem.F = a(solver_out) - <v, u_in> - <phi, h_in> gives M = -dF/d(solver_in).  krylov_exp.apply evaluates
    exp(-t M^-1 K') with K' = (A_shifted - M)/gamma, which is what ShiftedOperator(K=K') solves."""
    import firedrake as fd                                  # noqa: F401 (fails loudly without Firedrake)
    problem = solver._problem
    W = solver_in.function_space()
    n_u = W.sub(0).dof_count
    A = _petsc_to_csr(fd.assemble(problem.J, mat_type="aij").M.handle)
    M = _petsc_to_csr(fd.assemble(-fd.derivative(problem.F, solver_in), mat_type="aij").M.handle)
    K = ((A - M) / float(gamma)).tocsr()
    Mu = M[:n_u][:, :n_u].tocsr()
    Meta = M[n_u:][:, n_u:].tobsr(blocksize=(3, 3))
    nb = Meta.shape[0] // 3
    if Meta.indices.shape[0] != nb or np.any(Meta.indices != np.arange(nb)):
        raise ValueError("DG1 mass matrix is not 3x3 block diagonal in this dof order")
    return K, Mu, Meta.data
