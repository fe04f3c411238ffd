"""ctypes wrapper of oracle/c/oracle_budget.c (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

The C/OpenMP twin evaluates the seven budget diagnostics per cell exactly as the NumPy oracle does (no sharing of
intermediates, one sweep per diagnostic), fast enough to time as the CPU baseline and to check the CUDA path at
sizes the NumPy oracle cannot reach.  tests/test_oracle_c_twin.py pins it against the NumPy oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "liboracle_budget.so")
DIAGS = ("KE", "ADV", "PR", "WB", "EPS", "CHI", "CDIFF")


class ob_state(C.Structure):
    _fields_ = [("Nx", C.c_int), ("Ny", C.c_int), ("Nz", C.c_int), ("H", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("dz", C.c_double), ("nu", C.c_double), ("kappa", C.c_double), ("ghat", C.c_double * 3),
                ("u", C.c_void_p), ("v", C.c_void_p), ("w", C.c_void_p), ("b", C.c_void_p), ("p", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} missing: run __graft_entry__.build_oracle()")
        _lib = C.CDLL(LIB)
        _lib.ob_integral.restype = C.c_double
        _lib.ob_integral.argtypes = [C.POINTER(ob_state), C.c_int]
        _lib.ob_field.restype = C.c_int
        _lib.ob_field.argtypes = [C.POINTER(ob_state), C.c_int, C.c_void_p]
        _lib.ob_num_threads.restype = C.c_int
        _lib.ob_set_num_threads.argtypes = [C.c_int]
    return _lib


class BudgetState:
    """Parent arrays in the library's layout: numpy arrays of shape (Nz(+1)+2H, Ny+2H, Nx+2H), C-contiguous."""

    def __init__(self, N, H, d, nu, kappa, u, v, w, b, p, ghat=(0.0, 0.0, 1.0)):
        self.arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in (u, v, w, b, p)]
        Nx, Ny, Nz = N
        for a, nz in zip(self.arrays, (Nz, Nz, Nz + 1, Nz, Nz)):
            assert a.shape == (nz + 2 * H, Ny + 2 * H, Nx + 2 * H), a.shape
        s = ob_state()
        s.Nx, s.Ny, s.Nz, s.H = Nx, Ny, Nz, H
        s.dx, s.dy, s.dz = d
        s.nu, s.kappa = nu, kappa
        for i in range(3):
            s.ghat[i] = ghat[i]
        s.u, s.v, s.w, s.b, s.p = [a.ctypes.data for a in self.arrays]
        self.s, self.N = s, N

    def integral(self, diag):
        return lib().ob_integral(C.byref(self.s), DIAGS.index(diag))

    def field(self, diag):
        Nx, Ny, Nz = self.N
        out = np.empty((Nz, Ny, Nx))
        rc = lib().ob_field(C.byref(self.s), DIAGS.index(diag), out.ctypes.data)
        assert rc == 0
        return out.transpose(2, 1, 0)    # [i, j, k]


def from_oracle_fields(og, ou, ov, ow, ob, op, nu, kappa):
    """Build a BudgetState from NumPy-oracle fields (OField.data is [i, j, k])."""
    t = lambda f: np.ascontiguousarray(f.data.transpose(2, 1, 0))
    d = (float(og.dc[0][1]), float(og.dc[1][1]), float(og.dc[2][1]))
    return BudgetState(og.N, og.H[0], d, nu, kappa, t(ou), t(ov), t(ow), t(ob), t(op))
