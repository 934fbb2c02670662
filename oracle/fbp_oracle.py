"""TEST INFRASTRUCTURE ONLY -- python face of oracle/fbp_oracle.c.

Builds the geometry tables with the numpy expressions of the reference
`FBP.__init__` / `getrphi` (Recon/FBP_kernel.py:27-84) and calls the C loops.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfbp_oracle.so")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "fbp_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "-s"])
    return _SO


class Tables:
    """FBP_kernel.py:32-60 restated (same expressions, same dtypes)."""

    def __init__(self):
        self.D = 59.5
        m = 912
        self.da = 0.0010125
        self.theta = np.arange(0, 359.82 + 0.18, 0.18) / 180 * np.pi                      # :38  f64 [2000]
        self.nda = np.arange((-m / 2 + 0.5 + 3.75) * self.da, (m / 2 - 0.5 + 3.75 + 1) * self.da,
                             self.da).astype("float32")                                 # :39-40 f32 [912]
        n = self.nda.size
        h = np.zeros((2 * n - 1, 1))
        ng = np.arange(-n + 1, n, 2) * self.da
        h[0:2 * n - 1:2] = (-0.5 / np.pi ** 2. / (np.sin(ng) ** 2))[:, None]             # :54
        h[n - 1] = 1 / 8 / self.da ** 2                                                  # :55
        self.h = np.ascontiguousarray((h * self.da).astype("float32")[:, 0])             # :56 f32 [1823]
        g, L = 512, 21
        isect = np.arange(0, g * g)
        i, j = np.unravel_index(isect, (g, g))
        i = i + 1
        j = j + 1
        y = (g + 1 - i - g / 2 - 0.5) * 2 * L / g                                        # :77
        x = (j - g / 2 - 0.5) * 2 * L / g                                                # :78
        self.r = np.sqrt(x ** 2 + y ** 2)
        with np.errstate(divide="ignore", invalid="ignore"):
            phi = np.arctan(y / x)
        phi[x < 0] = phi[x < 0] + np.pi
        phi[phi < 0] = phi[phi < 0] + 2 * np.pi
        self.phi = phi
        self.wcos = np.ascontiguousarray((self.D * np.cos(self.nda)).astype(np.float32))  # :104 (f32 * weak python float)
        self.dtheta = float(self.theta[1] - self.theta[0])                               # :105 (f64 scalar)


_tables = None
_lib = None


def _load():
    global _tables, _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _tables = Tables()
    return _lib, _tables


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def convert(pj, flip=True):
    """Reference `FBP.convert` on [B,2000,912] or [2000,912] f32 -> [B,512,512] f32."""
    lib, t = _load()
    pj = np.ascontiguousarray(pj, dtype=np.float32)
    if pj.ndim == 2:
        pj = pj[None]
    assert pj.shape[1:] == (2000, 912), pj.shape
    out = np.empty((pj.shape[0], 512, 512), dtype=np.float32)
    rc = lib.fbp_oracle_convert(_p(pj, ctypes.c_float), _p(out, ctypes.c_float), ctypes.c_int(pj.shape[0]),
                                ctypes.c_int(1 if flip else 0), _p(t.wcos, ctypes.c_float), ctypes.c_double(t.dtheta),
                                _p(t.h, ctypes.c_float), _p(t.theta, ctypes.c_double), _p(t.nda, ctypes.c_float),
                                _p(t.r, ctypes.c_double), _p(t.phi, ctypes.c_double), ctypes.c_double(t.D),
                                ctypes.c_double(t.da))
    if rc != 0:
        raise MemoryError("fbp_oracle_convert")
    return out


def weight_and_ramp(pj, flip=True):
    """Stages 1-2 only (weighted + ramp-filtered sinogram [2000,912] f32) for per-kernel parity."""
    lib, t = _load()
    pj = np.ascontiguousarray(pj, dtype=np.float32)
    w = np.empty_like(pj)
    q = np.empty_like(pj)
    lib.fbp_oracle_weight(_p(pj, ctypes.c_float), _p(w, ctypes.c_float), _p(t.wcos, ctypes.c_float),
                          ctypes.c_double(t.dtheta), ctypes.c_int(1 if flip else 0))
    lib.fbp_oracle_ramp(_p(w, ctypes.c_float), _p(q, ctypes.c_float), _p(t.h, ctypes.c_float))
    return w, q
