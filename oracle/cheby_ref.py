"""TEST INFRASTRUCTURE (oracle).  cheby_exp.apply restated on NumPy vectors.

apply_literal follows /root/reference/cheby_exp.py:110-165 statement by statement
(two real vectors for Re/Im, 2*ncheb-1 operator applications, half of them on zero
vectors).  apply_real is the collapsed one-vector recurrence the GPU path runs;
apply_batch runs it for many sample times t_k at once.  `op(v)` is one
operator_solver.solve(): v -> M^-1 K v.
"""
import numpy as np


def apply_literal(c, ncheb, L, op, x, t):
    Tm1_r = x.copy()                                   # :114
    Tm1_i = np.zeros_like(x)                           # :115
    y = np.zeros_like(x)                               # :119
    y += np.real(c[0]) * Tm1_r                         # :120-123
    T_r = np.zeros_like(x)                             # :129
    T_i = op(x) * (-t / L)                             # :127-131
    y += -np.imag(c[1]) * T_i                          # :133-136
    for i in range(2, ncheb + 1):                      # :138
        Tm2_r, Tm2_i = Tm1_r, Tm1_i                    # :139-140
        Tm1_r, Tm1_i = T_r, T_i                        # :141-142
        T_i = op(Tm1_r) * (-2 * t / L)                 # :145-148
        T_r = op(Tm1_i) * (2 * t / L)                  # :149-152
        T_i = T_i - Tm2_i                              # :154
        T_r = T_r - Tm2_r                              # :155
        y += np.real(c[i]) * T_r                       # :157-160
        y += -np.imag(c[i]) * T_i                      # :162-165
    return y


def apply_real(a, L, op, x, t):
    """y = sum a_n S_n, S_n = 2 (t/L) op(S_{n-1}) + S_{n-2}."""
    s = t / L
    Sm2 = x.copy()
    y = a[0] * Sm2
    if len(a) == 1:
        return y
    Sm1 = s * op(x)
    y += a[1] * Sm1
    for n in range(2, len(a)):
        S = 2.0 * s * op(Sm1) + Sm2
        y += a[n] * S
        Sm2, Sm1 = Sm1, S
    return y


def apply_batch(a, L, op_mat, X, t):
    """Columns of X (N x B) advanced with their own t[k]; op_mat acts column-wise."""
    s = (np.asarray(t, dtype=np.float64) / L)[None, :]
    Sm2 = X.copy()
    Y = a[0] * Sm2
    if len(a) == 1:
        return Y
    Sm1 = s * op_mat(X)
    Y += a[1] * Sm1
    for n in range(2, len(a)):
        S = 2.0 * s * op_mat(Sm1) + Sm2
        Y += a[n] * S
        Sm2, Sm1 = Sm1, S
    return Y
