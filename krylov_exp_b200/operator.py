"""Host side of the operator M^-1 K: CSR export -> (optional) static condensation ->
upload through the C ABI.

Replaces the construction of `OperatorSolver` at
/root/reference/averaged_sw_explicit.py:142-164: K and M are assembled once on the host
(Firedrake, or any assembler that yields CSR), everything after that runs on the GPU.

Static condensation (exact algebra, done once with SciPy on the host): BDM2 has 3 dofs
per cell that couple only inside the cell, so M_u splits as [[M_EE, M_EI],[M_IE, M_II]]
with M_II block diagonal.  With R = -M_II^-1 M_IE the mass solve becomes

    S w_E = z_E + R^T z_I,      S = M_EE - M_EI M_II^-1 M_IE   (edge dofs only, 0.6 n_u)
    w_I   = M_II^-1 z_I + R w_E

and because z = K s is itself a sparse product, R^T, M_II^-1 and M_eta^-1 are folded into
K on the host:  Kt = [K_E + R^T K_I ; M_II^-1 K_I ; M_eta^-1 K_eta].  The device then runs
T = Kt s (one SpMM), CG on S, and w_I = T_I + R w_E inside the recurrence kernel.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components

from . import _lib

KCLASS = {"spmm_K": 0, "cg_init": 1, "spmm_cg": 2, "cg_update_xr": 3, "cg_update_p": 4, "cheb_update": 5, "misc": 6}


def _csr(A, dtype=np.float64):
    A = sp.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    return A


def _csr_args(A):
    """(rowptr int32, col int32, val float64) contiguous arrays kept alive by the caller."""
    if A is None:
        return None, None, None
    if A.nnz >= 2 ** 31:
        raise ValueError("matrix too large for 32-bit CSR indices")
    return (np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64))


def detect_interior_dofs(Mu):
    """Dofs that can be eliminated exactly: rows of M_u with the smallest stencil whose mutual
    coupling graph splits into components of exactly 3 dofs (BDM2 cell-interior dofs).
    Returns (index array grouped by component, shape (nblocks,3)) or None."""
    Mu = _csr(Mu)
    rownnz = np.diff(Mu.indptr)
    cand = np.flatnonzero(rownnz == rownnz.min())
    if len(cand) == 0 or len(cand) == Mu.shape[0] or len(cand) % 3:
        return None
    sub = Mu[cand][:, cand]
    ncomp, lab = connected_components(sub, directed=False)
    if ncomp * 3 != len(cand) or np.any(np.bincount(lab) != 3):
        return None
    order = np.argsort(lab, kind="stable")
    return cand[order].reshape(-1, 3)


def _block_diag_inv(A, groups):
    """Inverse of the block-diagonal matrix A[groups][:, groups] as dense (nb,3,3)."""
    idx = groups.ravel()
    sub = _csr(A[idx][:, idx]).tobsr(blocksize=(3, 3))
    nb = groups.shape[0]
    if sub.indices.shape[0] != nb or np.any(sub.indices != np.arange(nb)):
        raise ValueError("eliminated block is not block diagonal")
    return np.linalg.inv(sub.data)


def _blockdiag(blocks):
    nb = blocks.shape[0]
    indptr = np.arange(nb + 1)
    return sp.bsr_matrix((blocks, np.arange(nb), indptr), shape=(3 * nb, 3 * nb)).tocsr()


class CondensedPlan:
    """Matrices of the condensed operator plan in internal order [E ; I ; eta]."""

    def __init__(self, K, Mu, Meta_inv_blocks, interior=None):
        K, Mu = _csr(K), _csr(Mu)
        n_u = Mu.shape[0]
        n = K.shape[0]
        n_eta = n - n_u
        if interior is None:
            interior = detect_interior_dofs(Mu)
        if interior is None:
            raise ValueError("no eliminable dofs found")
        I = interior.ravel()
        mask = np.ones(n_u, dtype=bool)
        mask[I] = False
        E = np.flatnonzero(mask)
        self.n, self.n_s, self.n_b, self.n_d = n, len(E), len(I), n_eta
        order = np.concatenate([E, I, np.arange(n_u, n)])          # internal -> user
        self.perm = np.empty(n, dtype=np.int32)                    # user -> internal
        self.perm[order] = np.arange(n, dtype=np.int32)
        MII_inv = _blockdiag(_block_diag_inv(Mu, interior))
        MEE, MEI, MIE = Mu[E][:, E], Mu[E][:, I], Mu[I][:, E]
        R = _csr(-(MII_inv @ MIE))                                 # n_I x n_E
        self.As = _csr(MEE + MEI @ R)
        Kp = K[order][:, order].tocsr()
        nE, nI = len(E), len(I)
        KE, KI, Keta = Kp[:nE], Kp[nE:nE + nI], Kp[nE + nI:]
        Meta_inv = _blockdiag(np.ascontiguousarray(Meta_inv_blocks, dtype=np.float64))
        self.Kt = _csr(sp.vstack([KE + R.T @ KI, MII_inv @ KI, Meta_inv @ Keta]))
        self.Rt = R
        self.order = order


def block_jacobi_spectrum(As, safety=1.015):
    """Enclosure [lmin, lmax] of the spectrum of P^-1 S, P the 3x3 block-Jacobi of the SPD matrix S
    (host, once; SciPy Lanczos on the symmetrised L^-1 S L^-T, D = L L^T)."""
    from scipy.sparse.linalg import eigsh
    As = _csr(As)
    n = As.shape[0]
    nb = n // 3
    B = As.tobsr(blocksize=(3, 3))
    rows = np.repeat(np.arange(nb), np.diff(B.indptr))
    D = B.data[B.indices == rows]
    Li = np.linalg.inv(np.linalg.cholesky(D))
    LiM = sp.bsr_matrix((Li, np.arange(nb), np.arange(nb + 1)), shape=(n, n)).tocsr()
    Sym = (LiM @ As @ LiM.T).tocsc()
    # plain Lanczos (no factorisation): Ritz values lie inside the spectrum, hence the safety factor
    lmax = float(eigsh(Sym, k=1, which="LA", return_eigenvectors=False, tol=1e-4, ncv=40)[0])
    lmin = float(eigsh(Sym, k=1, which="SA", return_eigenvectors=False, tol=1e-4, ncv=40)[0])
    return lmin / safety, lmax * safety


class Operator:
    """Device-resident M^-1 K.  `condense=True` uploads the statically condensed plan.
    solver: "cg" (batched block-Jacobi PCG) or "chebyshev" (Chebyshev semi-iteration on the same
    preconditioned system, spectrum bounds estimated here on the host)."""

    def __init__(self, K, Mu, Meta_inv_blocks, device=0, condense=True, rtol=1e-12, max_it=60, precond=3,
                 solver="chebyshev", spectrum=None):
        self.lib = _lib.load()
        self._ctx = C.c_void_p()
        rc = self.lib.kexp_create(C.byref(self._ctx), int(device))
        if rc != 0:
            self._ctx = None
            raise _lib.KexpError(f"kexp_create failed ({rc}): a CUDA sm_100a device is required; there is no CPU path")
        K, Mu = _csr(K), _csr(Mu)
        self.n_u = Mu.shape[0]
        self.n = K.shape[0]
        self.n_eta = self.n - self.n_u
        Meta_inv_blocks = np.ascontiguousarray(Meta_inv_blocks, dtype=np.float64)
        self.condensed = False
        self._keep = []
        self._check(self.lib.kexp_set_mass_solver(self._ctx, rtol, max_it, precond), "kexp_set_mass_solver")
        plan = None
        if condense:
            interior = detect_interior_dofs(Mu)
            if interior is not None:
                plan = CondensedPlan(K, Mu, Meta_inv_blocks, interior)
        if plan is not None:
            kt, as_, rt = _csr_args(plan.Kt), _csr_args(plan.As), _csr_args(plan.Rt)
            perm = np.ascontiguousarray(plan.perm)
            self._check(self.lib.kexp_set_operator_plan(
                self._ctx, plan.n, plan.n_s, plan.n_b,
                _lib.ptr(kt[0]), _lib.ptr(kt[1]), _lib.ptr(kt[2]),
                _lib.ptr(as_[0]), _lib.ptr(as_[1]), _lib.ptr(as_[2]),
                _lib.ptr(rt[0]), _lib.ptr(rt[1]), _lib.ptr(rt[2]),
                None, None, _lib.ptr(perm)), "kexp_set_operator_plan")
            self.condensed = True
            self.n_solve = plan.n_s
            self.nnz = {"Kt": plan.Kt.nnz, "As": plan.As.nnz, "Rt": plan.Rt.nnz}
            As_host = plan.As
        else:
            k, m = _csr_args(K), _csr_args(Mu)
            self._check(self.lib.kexp_set_operator(
                self._ctx, self.n_u, self.n_eta, _lib.ptr(k[0]), _lib.ptr(k[1]), _lib.ptr(k[2]),
                _lib.ptr(m[0]), _lib.ptr(m[1]), _lib.ptr(m[2]), _lib.ptr(Meta_inv_blocks), self.n_eta // 3),
                "kexp_set_operator")
            self.n_solve = self.n_u
            self.nnz = {"Kt": K.nnz, "As": Mu.nnz, "Rt": 0}
            As_host = Mu
        self.solver = solver
        self.spectrum = None
        if solver == "chebyshev":
            self.spectrum = tuple(spectrum) if spectrum is not None else block_jacobi_spectrum(As_host)
            self._check(self.lib.kexp_set_mass_solver_chebyshev(self._ctx, rtol, self.spectrum[0], self.spectrum[1], max_it),
                        "kexp_set_mass_solver_chebyshev")
        elif solver != "cg":
            raise ValueError("solver must be 'cg' or 'chebyshev'")
        self._next_cheby_id = 0

    # -- plumbing
    def _check(self, rc, what):
        _lib.check(self.lib, self._ctx, rc, what)

    @property
    def ctx(self):
        if self._ctx is None:
            raise _lib.KexpError("operator has been closed")
        return self._ctx

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.kexp_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mass_solver(self, rtol=1e-12, max_it=60, precond=3):
        self._check(self.lib.kexp_set_mass_solver(self.ctx, rtol, max_it, precond), "kexp_set_mass_solver")

    def new_cheby_id(self):
        i = self._next_cheby_id
        self._next_cheby_id += 1
        return i

    def set_async(self, on=True):
        """Applies on device batches return once enqueued; the convergence record of their mass solves is checked by a
        later call (at the latest synchronize / download / stats), which raises if a solve missed rtol."""
        self._check(self.lib.kexp_set_async(self.ctx, int(bool(on))), "kexp_set_async")

    def stream_ptr(self):
        """The library's compute stream (cudaStream_t as an int), e.g. for torch.cuda.ExternalStream."""
        st = C.c_void_p()
        self._check(self.lib.kexp_get_stream(self.ctx, C.byref(st)), "kexp_get_stream")
        return int(st.value or 0)

    def wait_transfers(self):
        """Completes the asynchronous device-to-host copies (cheby_exp.apply_batch(wait=False), DeviceBatch.download(wait=False))."""
        self._check(self.lib.kexp_wait_transfers(self.ctx), "kexp_wait_transfers")

    def synchronize(self):
        self._check(self.lib.kexp_synchronize(self.ctx), "kexp_synchronize")

    def stats(self):
        s = _lib.Stats()
        self._check(self.lib.kexp_stats(self.ctx, C.byref(s)), "kexp_stats")
        return {f: getattr(s, f) for f, _ in s._fields_}

    def stats_reset(self):
        self._check(self.lib.kexp_stats_reset(self.ctx), "kexp_stats_reset")

    def set_kernel_timing(self, on):
        self._check(self.lib.kexp_set_kernel_timing(self.ctx, int(on)), "kexp_set_kernel_timing")

    def class_stats(self, name):
        ms, ln, b, mb = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
        self._check(self.lib.kexp_get_class_stats(self.ctx, KCLASS[name], C.byref(ms), C.byref(ln), C.byref(b)), "class_stats")
        self._check(self.lib.kexp_get_class_matrix_bytes(self.ctx, KCLASS[name], C.byref(mb)), "class_matrix_bytes")
        return {"ms": ms.value, "launches": ln.value, "bytes_per_col": b.value, "matrix_bytes": mb.value}

    def timer_start(self):
        self._check(self.lib.kexp_timer_start(self.ctx), "kexp_timer_start")

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.lib.kexp_timer_stop(self.ctx, C.byref(ms)), "kexp_timer_stop")
        return ms.value

    def last_device_ms(self):
        ms = C.c_double()
        self._check(self.lib.kexp_last_device_ms(self.ctx, C.byref(ms)), "kexp_last_device_ms")
        return ms.value

    # -- batches
    def batch(self, B):
        return DeviceBatch(self, B)

    def apply(self, X):
        """One batched operator_solver.solve(): columns of X (B x n, one vector per row)."""
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        bi, bo = self.batch(X.shape[0]), self.batch(X.shape[0])
        bi.upload(X)
        self._check(self.lib.kexp_operator_apply_dev(self.ctx, bi.handle, bo.handle), "kexp_operator_apply_dev")
        Y = bo.download()
        bi.close(); bo.close()
        return Y

    def largest_imag_eig(self, tol=1e-10, ncv=None, v0=None):
        """Largest |imag| eigenvalue of M^-1 K -- the `eigs` table of averaged_sw_explicit.py:77-82, which the
        reference hard-codes per refinement level and computes offline with SLEPc (L_eigs.py:36-58,
        eps_largest_imaginary).  M^-1 K is skew-adjoint in the energy inner product, so (M^-1 K)^2 has the real
        spectrum -omega^2; ARPACK (host) drives the device operator for the products.  Needed for levels the
        table lacks (ref >= 6, averaged_sw_williamson6.py:112-117)."""
        from scipy.sparse.linalg import LinearOperator, eigs
        bi, bo = self.batch(1), self.batch(1)
        try:
            def A2(v):
                for _ in range(2):
                    bi.upload(np.ascontiguousarray(v, dtype=np.float64).reshape(1, -1))
                    self._check(self.lib.kexp_operator_apply_dev(self.ctx, bi.handle, bo.handle), "kexp_operator_apply_dev")
                    v = bo.download()[0]
                return v
            if v0 is None:
                v0 = np.random.default_rng(0).standard_normal(self.n)
            vals = eigs(LinearOperator((self.n, self.n), matvec=A2, dtype=np.float64), k=1, which="LM", tol=tol,
                        ncv=ncv, v0=v0, return_eigenvectors=False)
        finally:
            bi.close(); bo.close()
        return float(np.sqrt(np.max(-vals.real)))

    def spmm_K(self, X):
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        bi, bo = self.batch(X.shape[0]), self.batch(X.shape[0])
        bi.upload(X)
        self._check(self.lib.kexp_spmm_K_dev(self.ctx, bi.handle, bo.handle), "kexp_spmm_K_dev")
        Y = bo.download()
        bi.close(); bo.close()
        return Y


class DeviceBatch:
    """B state vectors resident in HBM in the batch-minor panel layout."""

    def __init__(self, op, B):
        self.op, self.B = op, int(B)
        self.handle = C.c_void_p()
        op._check(op.lib.kexp_batch_create(op.ctx, self.B, C.byref(self.handle)), "kexp_batch_create")

    @property
    def num_panels(self):
        return int(self.op.lib.kexp_batch_num_panels(self.handle))

    def set_order(self, order=None):
        """Stored column j holds the caller's column order[j]; uploads, downloads, sample times and weights keep the
        caller's order.  `np.argsort(np.abs(t), kind="stable")` groups sample times of similar series length."""
        if order is None:
            self.op._check(self.op.lib.kexp_batch_set_order(self.handle, None), "kexp_batch_set_order")
            return
        order = np.ascontiguousarray(order, dtype=np.int32)
        assert order.shape == (self.B,)
        self.op._check(self.op.lib.kexp_batch_set_order(self.handle, _lib.ptr(order)), "kexp_batch_set_order")

    def upload(self, X, shared=False):
        X = np.asarray(X, dtype=np.float64)
        if shared:
            X = np.ascontiguousarray(X.reshape(-1))
            assert X.shape[0] == self.op.n
            ld = self.op.n
        else:
            X = np.atleast_2d(X)
            assert X.shape == (self.B, self.op.n), (X.shape, self.B, self.op.n)
            if X.strides[1] != 8 or X.strides[0] % 8 or (self.B > 1 and X.strides[0] < 8 * self.op.n):
                X = np.ascontiguousarray(X)
            ld = X.strides[0] // 8 if self.B > 1 else self.op.n     # a single row: NumPy's stride[0] is arbitrary (0 for x[None, :])
        self.op._check(self.op.lib.kexp_batch_upload(self.handle, X.ctypes.data_as(C.c_void_p), ld, int(shared)), "kexp_batch_upload")

    # -- device-resident vectors in the caller's layout (e.g. torch tensors): no host copy, no synchronisation
    @staticmethod
    def _dev_array(T, rows, n):
        """(data pointer, leading dimension) of a CUDA float64 tensor / __cuda_array_interface__ object of shape (rows, n) or (n,)."""
        if hasattr(T, "data_ptr"):                                    # torch.Tensor
            import torch
            if T.dtype != torch.float64 or not T.is_cuda:
                raise ValueError("device vectors must be CUDA float64 tensors")
            shape, strides, ptr = tuple(T.shape), tuple(T.stride()), T.data_ptr()
        else:
            ci = T.__cuda_array_interface__
            if ci["typestr"] not in ("<f8", "=f8"):
                raise ValueError("device vectors must be float64")
            shape, ptr = tuple(ci["shape"]), ci["data"][0]
            strides = tuple(s // 8 for s in ci["strides"]) if ci.get("strides") else None
            if strides is None:
                strides = (shape[1], 1) if len(shape) == 2 else (1,)
        if len(shape) == 1:
            if shape[0] != n or strides[0] != 1 or rows != 1:
                raise ValueError(f"expected a contiguous vector of n = {n} values")
            return ptr, n
        if shape != (rows, n) or strides[1] != 1 or (rows > 1 and strides[0] < n):
            raise ValueError(f"expected shape ({rows}, {n}) with unit inner stride, got {shape} / {strides}")
        return ptr, (strides[0] if rows > 1 else n)

    def upload_dev(self, T, shared=False):
        """Device array (B, n) -- or one shared vector (n,) -- in user dof order -> panels; enqueued on the library stream."""
        ptr, ld = self._dev_array(T, 1 if shared else self.B, self.op.n)
        self.op._check(self.op.lib.kexp_batch_upload_dev(self.handle, C.c_void_p(ptr), ld, int(shared)), "kexp_batch_upload_dev")

    def export_dev(self, T):
        """Panels -> device array (B, n) in user dof order; enqueued on the library stream."""
        ptr, ld = self._dev_array(T, self.B, self.op.n)
        self.op._check(self.op.lib.kexp_batch_export_dev(self.handle, C.c_void_p(ptr), ld), "kexp_batch_export_dev")

    def download(self, out=None, wait=True):
        if out is None:
            out = np.empty((self.B, self.op.n))
        assert out.dtype == np.float64 and out.shape == (self.B, self.op.n) and out.strides[1] == 8
        fn = self.op.lib.kexp_batch_download if wait else self.op.lib.kexp_batch_download_async
        ld = out.strides[0] // 8 if self.B > 1 else self.op.n
        assert ld >= self.op.n
        self.op._check(fn(self.handle, out.ctypes.data_as(C.c_void_p), ld), "kexp_batch_download")
        return out

    def average_into(self, w, v_dev_ptr):
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.shape == (self.B,)
        self.op._check(self.op.lib.kexp_batch_average_dev(self.handle, _lib.ptr(w), C.c_void_p(v_dev_ptr)), "kexp_batch_average_dev")

    def close(self):
        if self.handle:
            self.op.lib.kexp_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.op._ctx is not None:
                self.close()
        except Exception:
            pass
