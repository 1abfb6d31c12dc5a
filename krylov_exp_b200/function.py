"""Vector access for the drop-in classes: Firedrake mixed Functions, this package's
MixedVector stand-in, or plain NumPy arrays, all seen as one contiguous [u ; eta] array
(the order of Firedrake's mixed Function: f.subfunctions[0].dat.data, [1].dat.data)."""
import numpy as np


class MixedVector:
    """Minimal stand-in for a Firedrake Function on W = V1 x V2 (used where Firedrake is absent)."""

    def __init__(self, n_u, n_eta, data=None):
        self.n_u, self.n_eta = int(n_u), int(n_eta)
        self.data = np.zeros(self.n_u + self.n_eta) if data is None else np.array(data, dtype=np.float64)
        assert self.data.shape == (self.n_u + self.n_eta,)

    @property
    def subfunctions(self):
        return (self.data[:self.n_u], self.data[self.n_u:])

    def assign(self, other):
        self.data[:] = as_array(other) if not np.isscalar(other) else other
        return self

    def copy(self):
        return MixedVector(self.n_u, self.n_eta, self.data.copy())


def _firedrake_parts(f):
    subs = getattr(f, "subfunctions", None)
    if subs is None and hasattr(f, "split"):
        subs = f.split()
    return [s.dat.data for s in subs]


def as_array(f):
    if isinstance(f, np.ndarray):
        return f
    if isinstance(f, MixedVector):
        return f.data
    if hasattr(f, "dat"):                      # Firedrake Function
        parts = _firedrake_parts(f)
        return np.concatenate([np.asarray(p).reshape(-1) for p in parts])
    raise TypeError(f"cannot read a vector from {type(f)}")


def write_back(f, values):
    if isinstance(f, np.ndarray):
        f[:] = values
    elif isinstance(f, MixedVector):
        f.data[:] = values
    elif hasattr(f, "dat"):
        o = 0
        for p in _firedrake_parts(f):
            m = p.size
            p[...] = values[o:o + m].reshape(p.shape)
            o += m
    else:
        raise TypeError(f"cannot write a vector to {type(f)}")
