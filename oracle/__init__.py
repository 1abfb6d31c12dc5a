"""TEST INFRASTRUCTURE: CPU oracle for the exp(sL)v hot path of colinjcotter/krylov_exp.

Restates the reference's algorithm (file:line cited per function) in NumPy/SciPy and C.
Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import or execute anything in this package; the product
package krylov_exp_b200 never does.

Parity status: the reference's arithmetic lives in Firedrake/PETSc
(un-vendored, unpinned, not installable here).  The oracle is pinned by
  * golden vectors produced by running the reference's OWN cheby_exp.py / krylov_exp.py
    classes (imported unmodified from /root/reference through a NumPy stand-in for the
    firedrake namespace; tests/golden/make_golden.py),
  * /root/reference/eigs.dat for the assembled operator (to ~1e-4: the reference's
    quadrature on curved cells is unknown, see oracle/assemble_sw.py).
At the Firedrake/PETSc assembly boundary itself parity remains unpinned.
"""
