"""TEST INFRASTRUCTURE (oracle).  Parity norms of /root/reference/calculate_norm.py:93-96.

errornorm(a,b)/norm(b) in L2 for eta and u, and in H(div) for u, as matrix-weighted
vector norms: ||v||_L2^2 = v^T M v, ||u||_Hdiv^2 = u^T (M_u + D) u, D = <div u, div v>.
"""
import numpy as np


def rel_l2_eta(op, a, b):
    Me = op.Meta()
    d = a - b
    return np.sqrt((d @ (Me @ d)) / (b @ (Me @ b)))


def rel_l2_u(op, a, b):
    d = a - b
    return np.sqrt((d @ (op.Mu @ d)) / (b @ (op.Mu @ b)))


def rel_hdiv_u(op, a, b):
    A = op.Mu + op.Ddiv
    d = a - b
    return np.sqrt((d @ (A @ d)) / (b @ (A @ b)))


def rel_mixed(op, a, b):
    """Relative M-weighted L2 norm of a full mixed vector [u; eta]."""
    M = op.M()
    d = a - b
    return np.sqrt((d @ (M @ d)) / (b @ (M @ b)))


def calculate_norm(op, a, b):
    """(etanorm, unorm_Hdiv, unorm_L2) of calculate_norm.py:94-96 for mixed vectors."""
    nu = op.n_u
    return (rel_l2_eta(op, a[nu:], b[nu:]), rel_hdiv_u(op, a[:nu], b[:nu]),
            rel_l2_u(op, a[:nu], b[:nu]))
