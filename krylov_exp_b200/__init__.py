"""krylov_exp_b200: B200-native exp(sL)v hot path of colinjcotter/krylov_exp.

    from krylov_exp_b200 import cheby_exp, krylov_exp, Operator

`cheby_exp` / `krylov_exp` keep the reference's constructor and apply(x, y, t) signatures
(/root/reference/cheby_exp.py:9-10,110; krylov_exp.py:6-7,50).  All arithmetic runs in
libkexp.so (hand-written sm_100a CUDA); there is no CPU fallback.
"""
from ._lib import KexpError, load as load_library, LIB_PATH
from .operator import Operator, DeviceBatch, CondensedPlan, detect_interior_dofs
from .cheby import cheby_exp, chebyshev_coefficients, real_table
from .krylov import krylov_exp, ShiftedOperator
from .function import MixedVector
from .samples import sample_times_weights, natural_Mbar, EIGS
from .ensemble import SampleEnsemble, block_partition, balanced_partition, partition_cost

__all__ = ["cheby_exp", "krylov_exp", "Operator", "DeviceBatch", "CondensedPlan", "MixedVector",
           "ShiftedOperator", "chebyshev_coefficients", "real_table", "sample_times_weights",
           "natural_Mbar", "EIGS", "SampleEnsemble", "block_partition", "balanced_partition", "partition_cost", "KexpError", "load_library", "LIB_PATH", "detect_interior_dofs"]
