"""Field checkpoints of the time loop and the parity norms computed from them (second half).

Reference: /root/reference/averaged_sw_explicit.py:319-326 writes `<filename>_<hours>h` with the fields "Velocity" and
"Elevation" and the attributes time / tdump / tcheck through Firedrake's DumbCheckpoint (HDF5);
/root/reference/calculate_norm.py:66-96 reads two such files and prints  errornorm(eta0, eta1)/norm(eta1)  in L2 and
errornorm(u0, u1)/norm(u1)  in H(div) and L2.

Here the state is the flat [u ; eta] vector of this package, so a checkpoint holds the two arrays under the reference's
field names plus the attributes.  Container: HDF5 with the reference's dataset / attribute names when h5py is importable
(`/fields/Velocity`, `/fields/Elevation`, root attributes), otherwise NumPy .npz with the same keys -- h5py is not part
of this image, so the tests exercise the .npz path.  This is host-side post-processing, not part of the measured path.
"""
import os

import numpy as np


def _h5py():
    try:
        import h5py
        return h5py
    except Exception:
        return None


def store(filename, hours, n_u, U, time, tdump=0.0, tcheck=0.0, eta_name="Elevation"):
    """chk = DumbCheckpoint(filename + "_" + str(hours) + "h", FILE_CREATE); chk.store(un); chk.store(etan); attributes (:319-326)."""
    U = np.asarray(U, dtype=np.float64)
    base = f"{filename}_{int(hours)}h"
    h5 = _h5py()
    if h5 is not None:
        with h5.File(base + ".h5", "w") as f:
            g = f.create_group("fields")
            g.create_dataset("Velocity", data=U[:n_u])
            g.create_dataset(eta_name, data=U[n_u:])
            f.attrs["time"], f.attrs["tdump"], f.attrs["tcheck"] = float(time), float(tdump), float(tcheck)
        return base + ".h5"
    np.savez(base + ".npz", **{"Velocity": U[:n_u], eta_name: U[n_u:], "time": float(time), "tdump": float(tdump), "tcheck": float(tcheck)})
    return base + ".npz"


def load(filename, hours, eta_name="Elevation"):
    """chkfile.load(u, name="Velocity"); chkfile.load(eta, name="Elevation")  (calculate_norm.py:71-76) -> (U, attributes)."""
    base = f"{filename}_{int(hours)}h"
    if os.path.exists(base + ".h5"):
        h5 = _h5py()
        if h5 is None:
            raise RuntimeError(f"{base}.h5 needs h5py")
        with h5.File(base + ".h5", "r") as f:
            u, eta = f["fields/Velocity"][...], f["fields/" + eta_name][...]
            attrs = {k: float(f.attrs[k]) for k in ("time", "tdump", "tcheck") if k in f.attrs}
        return np.concatenate([u, eta]), attrs
    z = np.load(base + ".npz")
    return np.concatenate([z["Velocity"], z[eta_name]]), {k: float(z[k]) for k in ("time", "tdump", "tcheck") if k in z.files}


def calculate_norm(sw, file0, file1, hours):
    """The table of calculate_norm.py:64-105: for every hour (etanorm, unorm_Hdiv, unorm_L2) of file0 against file1.
    `sw`: the assembled operator (mass and div-div matrices give the L2 / H(div) norms)."""
    Me, Mu, A = sw.Meta(), sw.Mu, sw.Mu + sw.Ddiv
    nu = sw.n_u
    rel = lambda M, a, b: float(np.sqrt(((a - b) @ (M @ (a - b))) / (b @ (M @ b))))
    out = []
    for h in hours:
        U0, _ = load(file0, h)
        U1, _ = load(file1, h)
        out.append((rel(Me, U0[nu:], U1[nu:]), rel(A, U0[:nu], U1[:nu]), rel(Mu, U0[:nu], U1[:nu])))   # :93-96
    return out
