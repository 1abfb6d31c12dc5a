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


# ---------------------------------------------------------------------------------------------
# NumPy restatements of the three exact rewritings the GPU path uses; they are
# checked here, on the CPU, against apply_batch -- i.e. against the reference's own recurrence.
# ---------------------------------------------------------------------------------------------
def reexpand(a, s):
    """b with sum_n a_n S_n(s z) = sum_m b_m S_m(z): multiplication by z in the S basis is
    z S_0 = S_1, z S_m = (S_{m+1} - S_{m-1})/2, so E_0 = e_0, E_1 = s shift(E_0), E_n = 2 s shift(E_{n-1}) + E_{n-2}."""
    nb = len(a)

    def shift(e):
        o = np.zeros(nb + 2)
        o[1] += e[0]
        o[2:nb + 2] += 0.5 * e[1:nb + 1]
        o[0:nb] -= 0.5 * e[1:nb + 1]
        return o
    e0 = np.zeros(nb + 2); e0[0] = 1.0
    b = a[0] * e0[:nb]
    if nb == 1:
        return b
    e1 = s * shift(e0)
    b = b + a[1] * e1[:nb]
    for n in range(2, nb):
        e2 = 2.0 * s * shift(e1) + e0
        b = b + a[n] * e2[:nb]
        e0, e1 = e1, e2
    return b


def truncated(b, eps):
    keep = len(b)
    tail, total = 0.0, np.abs(b).sum()
    while keep > 1 and tail + abs(b[keep - 1]) <= eps * total:
        tail += abs(b[keep - 1]); keep -= 1
    return keep


def apply_batch_economised(a, L, op_mat, X, t, t_valid, eps=1e-13, margin=0.01):
    """Per-column economised series: the same polynomial as apply_batch, every column through its own shorter
    expansion in the common operator C = (t_valid / ((1+margin) L)) * op."""
    t = np.asarray(t, dtype=np.float64)
    cs = t_valid / ((1.0 + margin) * L)
    bs = [reexpand(a, abs(tk) * (1.0 + margin) / t_valid) for tk in t]
    M = max(truncated(b, eps) for b in bs) - 1
    coef = np.array([b[:M + 1] for b in bs]).T                 # (M+1) x B
    sign = np.where(t < 0, -1.0, 1.0)[None, :]
    Sm2 = X.copy()
    Y = coef[0][None, :] * Sm2
    if M == 0:
        return Y, M
    Sm1 = cs * sign * op_mat(X)
    Y += coef[1][None, :] * Sm1
    for n in range(2, M + 1):
        S = 2.0 * cs * sign * op_mat(Sm1) + Sm2
        Y += coef[n][None, :] * S
        Sm2, Sm1 = Sm1, S
    return Y, M


def average_clenshaw(a, L, op, X, t, w, eps=1e-13):
    """sum_k w_k p_k(op) X[:, k] by ONE Clenshaw recurrence on a single vector: z_m = sum_k w_k b_m^(k) X_k,
    u_m = z_m + 2 c op(u_{m+1}) + u_{m+2}, v = z_0 + c op(u_1) + u_2  (c = max|t| / L)."""
    t = np.asarray(t, dtype=np.float64)
    tstar = max(np.abs(t).max(), 1e-300)
    c = tstar / L
    D = np.array([reexpand(a, tk / tstar) for tk in t])        # B x nb
    M = max(truncated(D[k], eps) for k in range(len(t))) - 1
    Z = X @ (np.asarray(w)[:, None] * D[:, :M + 1])            # N x (M+1)
    u1 = Z[:, M].copy()
    u2 = np.zeros_like(u1)
    for m in range(M - 1, -1, -1):
        cs = 1.0 if m == 0 else 2.0
        u1, u2 = Z[:, m] + cs * c * op(u1) + u2, u1
    return u1, M
