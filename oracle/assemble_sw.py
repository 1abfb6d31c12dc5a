"""TEST INFRASTRUCTURE (oracle).  The discrete operator the oracle works on.

The standalone BDM2 x DG1 icosahedral assembler lives in krylov_exp_b200/assembly.py
(host-side input generation, standing in for Firedrake's assembly which the north star
keeps on the host).  The oracle re-exports it so that the CUDA path and the CPU oracle
are compared on identical K, M_u, M_eta; everything downstream of the matrices (operator
application, mass solves, recurrences, norms) is restated independently in this package.
"""
from krylov_exp_b200.assembly import *          # noqa: F401,F403
from krylov_exp_b200.assembly import assemble_sw, test_pulse, largest_imag_eig, SWOperator  # noqa: F401
