"""TEST INFRASTRUCTURE (oracle).  Chebyshev coefficients of exp(ix) on [-L, L].

Literal restatement of /root/reference/cheby_exp.py:31-96 including its quirks:
  :31-34  ncheb+2 Chebyshev points from theta=pi down to 0
  :43,46  optional (1+(x/fL)^2)^-4 filters
  :48-49  even extension, FFT, division by ncheb (not ncheb+1)
  :51-53  first ncheb+2 coefficients, first and last halved
  :57-63  tail compression while the removed L1 mass stays < tol; must remove >= 1 term
  :69-83  renormalise so the truncated series is exactly 1 at A=0 (complex divisor)
"""
import numpy as np
from scipy import fftpack


def cheby_coeffs(ncheb, tol, L, filter=False, filter_val=0.75, filter_freq=False):
    """Returns (ChebCoeffs complex128[ncheb0+2], ncheb_compressed)."""
    dpi = np.pi / (ncheb + 1)
    t1 = np.arange(np.pi, -dpi / 2, -dpi)
    x = L * np.cos(t1)
    fvals = np.exp(1j * x)
    fL = 0.003465 * 60 * 60                       # cheby_exp.py:37-38
    if filter_freq:
        fvals /= (1 + (x / fL) ** 2) ** 4
    elif filter:
        fvals /= (1 + (x / (filter_val * L)) ** 2) ** 4
    unit = np.concatenate((np.flipud(fvals), fvals[1:-1]))
    c = (fftpack.fft(unit) / ncheb)[:ncheb + 2].copy()
    c[0] /= 2
    c[-1] /= 2
    nrm = 0.0
    compressed = False
    while nrm + abs(c[ncheb + 1]) < tol:
        nrm += abs(c[ncheb + 1])
        ncheb -= 1
        compressed = True
    assert compressed
    # value of the truncated series at A = 0:  T_0=1, T_1=0, T_n = -T_{n-2}
    tm1, tn = 1.0, 0.0
    f0 = c[0] * tm1 + c[1] * tn
    for i in range(2, ncheb + 1):
        tm2, tm1 = tm1, tn
        tn = -tm2
        f0 = f0 + c[i] * tn
    c = c / f0
    return c, ncheb


def real_recurrence_coeffs(c, ncheb):
    """a_n = Re(c_n (-i)^n): with S_0=x, S_1=Bx, S_n=2B S_{n-1}+S_{n-2}, B=(t/L)A,
    the reference's y (cheby_exp.py:110-165) equals sum_n a_n S_n, because T_n is
    purely real for even n and purely imaginary for odd n."""
    n = np.arange(ncheb + 1)
    return np.real(c[:ncheb + 1] * (-1j) ** n)
