"""TEST INFRASTRUCTURE (oracle).  ctypes wrapper of oracle/cpu_cheby.c (C/OpenMP port of
the hot path): CPU checker for large cases and the CPU baseline timed by bench.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libcpu_cheby.so")


def build(force=False):
    src = os.path.join(_HERE, "cpu_cheby.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True)
    return _LIB


def _block_jacobi(Mu):
    nb = Mu.shape[0] // 3
    D = sp.csr_matrix(Mu).tobsr(blocksize=(3, 3))
    blocks = np.zeros((nb, 3, 3))
    rows = np.repeat(np.arange(nb), np.diff(D.indptr))
    diag = D.indices == rows
    blocks[rows[diag]] = D.data[diag]
    return np.ascontiguousarray(np.linalg.inv(blocks))


class CpuCheby:
    def __init__(self, K, Mu, Meta_inv_blocks, rtol=1e-12, max_it=500):
        build()
        self.lib = C.CDLL(_LIB)
        self.lib.cpu_cheby_apply_batch.restype = C.c_int64
        self.K = sp.csr_matrix(K); self.K.sort_indices()
        self.Mu = sp.csr_matrix(Mu); self.Mu.sort_indices()
        self.n_u = self.Mu.shape[0]
        self.n = self.K.shape[0]
        self.Meta_inv = np.ascontiguousarray(Meta_inv_blocks, dtype=np.float64)
        self.Pinv = _block_jacobi(self.Mu)
        self.rtol, self.max_it = rtol, max_it
        self._a = [np.ascontiguousarray(a) for a in (
            self.K.indptr.astype(np.int32), self.K.indices.astype(np.int32), self.K.data.astype(np.float64),
            self.Mu.indptr.astype(np.int32), self.Mu.indices.astype(np.int32), self.Mu.data.astype(np.float64))]

    def num_threads(self):
        return int(self.lib.cpu_num_threads())

    def apply_batch(self, coef, ncheb, L, X, t, shared=False, literal=False, nthreads=0):
        """coef: real table a_n (literal=False) or complex ChebCoeffs (literal=True)."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        B = t.shape[0]
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.empty((B, self.n))
        if literal:
            c = np.ascontiguousarray(np.asarray(coef[:ncheb + 1], dtype=np.complex128)).view(np.float64)
        else:
            c = np.ascontiguousarray(coef[:ncheb + 1], dtype=np.float64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        its = self.lib.cpu_cheby_apply_batch(
            C.c_int64(self.n_u), C.c_int64(self.n - self.n_u), *[p(a) for a in self._a],
            p(self.Meta_inv), p(self.Pinv), p(c), C.c_int(ncheb), C.c_double(L), C.c_int(int(literal)),
            p(X), C.c_int(int(shared)), p(t), C.c_int(B), p(Y), C.c_double(self.rtol), C.c_int(self.max_it),
            C.c_int(nthreads))
        return Y, int(its)
