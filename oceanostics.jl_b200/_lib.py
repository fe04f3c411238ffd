"""ctypes binding of liboceanostics_b200.so -- the same C ABI (include/oceanostics_b200.h) a Julia
`ccall` layer binds.  There is no CPU fallback: importing the package without the built
library raises, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCB_LIB_PATH") or os.path.join(HERE, "lib", "liboceanostics_b200.so")   # OCB_LIB_PATH: another build of the same ABI (A/B runs)

OCB_F32, OCB_F64 = 0, 1
OCB_PERIODIC, OCB_BOUNDED = 0, 1
OCB_CENTER, OCB_FACE, OCB_NOTHING = 0, 1, 2
OCB_FIELD_ARRAY, OCB_FIELD_ZERO, OCB_FIELD_CONSTANT = 0, 1, 2
OCB_MAX_CLOSURE_FIELDS = 2
OCB_N_AUX = 7
OCB_MAX_REQ = 16
OCB_MAX_FILTER_WIDTH = 64
OCB_MAX_RANKS = 16
OCB_RING_SYNC_WORDS = 1024
OCB_ABI_VERSION = 2

OCB_REQ_PERTURBATION = 1
OCB_REQ_TENSOR_DIM1, OCB_REQ_TENSOR_DIM2, OCB_REQ_TENSOR_DIM3 = 1 << 4, 1 << 5, 1 << 6
OCB_REQ_COLLOCATED = 1 << 7
OCB_REQ_CARRIED_HEIGHT = 1 << 8
OCB_REQ_HYDROSTATIC_SPLIT = 1 << 9
OCB_REQ_CENTERED4 = 1 << 10
OCB_REQ_TERM_SHIFT = 12
OCB_TERM_TENDENCY, OCB_TERM_ADVECTION, OCB_TERM_BUOYANCY, OCB_TERM_CORIOLIS, OCB_TERM_PRESSURE, OCB_TERM_DIFFUSION, OCB_TERM_FORCING = range(7)

_DIAG_NAMES = [
    "KE", "ADV", "PR", "WB", "EPS", "EPS_ISO", "TKE", "SPX", "SPY", "SPZ", "SP", "CHI", "CDIFF", "SMOD", "OMOD", "Q",
    "S11", "S22", "S33", "S12", "S13", "S23", "O12", "O13", "O23", "T11", "T22", "T33", "T11C", "T22C", "T33C",
    "T12", "T13", "T23", "RI", "RO", "EPV", "TWPV", "DEPV", "VF11", "VF22", "VF33", "VF12", "VF13", "VF23",
    "FEPS", "SFEPS", "SFKE", "PI", "BPE", "APE", "UPSILON", "APE_EPS", "KE_STRESS", "KE_FORCING", "C_TENDENCY", "KE_TENDENCY", "NU_SMAG",
    "U_TERM", "V_TERM", "W_TERM", "C_TERM", "QX", "QY", "QZ"]
DIAG = {name: i for i, name in enumerate(_DIAG_NAMES)}
OCB_DIAG_COUNT = len(_DIAG_NAMES)

# output location of each diagnostic ('c' / 'f' per axis), mirroring ocb_diag_location()
DIAG_LOC = {name: "ccc" for name in _DIAG_NAMES}
for _n in ("S12", "O12", "T12", "VF12"):
    DIAG_LOC[_n] = "ffc"
for _n in ("S13", "O13", "T13", "VF13"):
    DIAG_LOC[_n] = "fcf"
for _n in ("S23", "O23", "T23", "VF23"):
    DIAG_LOC[_n] = "cff"
DIAG_LOC.update(T11="fcc", T22="cfc", T33="ccf", RI="ccf", RO="fff", EPV="fff", TWPV="fff", DEPV="fff",
                U_TERM="fcc", V_TERM="cfc", W_TERM="ccf", QX="fcc", QY="cfc", QZ="ccf")

OCB_FILTER_BOX, OCB_FILTER_GAUSSIAN, OCB_FILTER_GAUSSIAN_STRETCHED = 0, 1, 2
OCB_POLICY_PERIODIC, OCB_POLICY_SHRINK, OCB_POLICY_EDGE, OCB_POLICY_CONSTANT = 0, 1, 2, 3
OCB_BPE_THREE_DIMENSIONAL_SORT, OCB_BPE_HEAVISIDE_INTEGRAL = 0, 1
OCB_OP_COPY, OCB_OP_ADD, OCB_OP_SUB, OCB_OP_MUL = 0, 1, 2, 3


class ocb_grid(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("N", C.c_int32 * 3), ("H", C.c_int32 * 3), ("topo", C.c_int32 * 3),
                ("L", C.c_double * 3), ("d", C.c_double * 3), ("z_bottom", C.c_double),
                ("dzc", C.c_void_p), ("dzf", C.c_void_p), ("zc", C.c_void_p), ("zf", C.c_void_p)]


class ocb_field(C.Structure):
    _fields_ = [("data", C.c_void_p), ("kind", C.c_int32), ("loc", C.c_int32 * 3), ("halo", C.c_int32 * 3),
                ("size", C.c_int32 * 3), ("stride", C.c_int64 * 3), ("value", C.c_double)]


class ocb_closure(C.Structure):
    _fields_ = [("nu_const", C.c_double), ("n_nu_fields", C.c_int32), ("nu_fields", ocb_field * OCB_MAX_CLOSURE_FIELDS),
                ("kappa_const", C.c_double), ("n_kappa_fields", C.c_int32),
                ("kappa_fields", ocb_field * OCB_MAX_CLOSURE_FIELDS), ("kappa_scale", C.c_double * OCB_MAX_CLOSURE_FIELDS)]


class ocb_request(C.Structure):
    _fields_ = [("diag", C.c_int32), ("flags", C.c_int32), ("out", C.c_void_p), ("out_halo", C.c_int32 * 3),
                ("out_stride", C.c_int64 * 3), ("window_lo", C.c_int32 * 3), ("window_hi", C.c_int32 * 3),
                ("out_offset", C.c_int64 * 3), ("integral_slot", C.c_int32)]


class ocb_sweep_inputs(C.Structure):
    _fields_ = [("u", ocb_field), ("v", ocb_field), ("w", ocb_field), ("b", ocb_field), ("p", ocb_field),
                ("U", ocb_field), ("V", ocb_field), ("W", ocb_field), ("B", ocb_field),
                ("aux", ocb_field * OCB_N_AUX), ("closure", ocb_closure),
                ("f", C.c_double * 3), ("dir", C.c_double * 3), ("f_dir", C.c_double), ("ghat", C.c_double * 3),
                ("ro_bg", C.c_double * 6), ("les", C.c_double * 4)]


class ocb_filter_pass(C.Structure):
    _fields_ = [("dim", C.c_int32), ("kind", C.c_int32), ("width", C.c_int32), ("policy", C.c_int32),
                ("extent", C.c_int32), ("left", C.c_double), ("right", C.c_double),
                ("weights", C.c_double * (2 * OCB_MAX_FILTER_WIDTH + 1)), ("sigma", C.c_double), ("period", C.c_double),
                ("nodes", C.c_void_p), ("spacings", C.c_void_p)]


class ocb_ring(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("sync", C.c_void_p * OCB_MAX_RANKS), ("timeout_seconds", C.c_double)]


class OcbError(RuntimeError):
    """A non-zero status from the library (OCB_ERR_CUDA / ALLOC / UNSUPPORTED)."""


_lib = None


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OcbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    L.ocb_abi_version.restype = C.c_int
    L.ocb_last_error.restype = C.c_char_p
    L.ocb_device_count.restype = C.c_int
    L.ocb_plan_create.argtypes = [P(ocb_grid), P(ocb_sweep_inputs), P(ocb_request), C.c_int32, P(C.c_void_p)]
    L.ocb_plan_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ocb_plan_destroy.argtypes = [C.c_void_p]
    L.ocb_plan_info.argtypes = [C.c_void_p, P(C.c_int32), P(C.c_int32), P(C.c_double)]
    L.ocb_fill_halo.argtypes = [P(ocb_grid), P(ocb_field), C.c_int32, C.c_double, C.c_double, C.c_void_p]
    L.ocb_filter_execute.argtypes = [C.c_int32, P(ocb_field), C.c_int32 * 3, P(ocb_filter_pass), C.c_int32,
                                     P(ocb_field), C.c_int32 * 3, C.c_int32 * 3, C.c_void_p]
    L.ocb_bpe_reference_height.argtypes = [P(ocb_grid), P(ocb_field), C.c_int32, P(ocb_field), P(ocb_field),
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    L.ocb_sort_pairs.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.ocb_gather_sorted.argtypes = [C.c_int32, P(ocb_field), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.ocb_bpe_profile_lookup.argtypes = [P(ocb_grid), P(ocb_field), C.c_void_p, C.c_void_p, C.c_int64, P(ocb_field), P(ocb_field),
                                         C.c_void_p]
    L.ocb_pointwise.argtypes = [P(ocb_grid), C.c_int32, C.c_double, P(ocb_field), P(ocb_field), P(ocb_field), C.c_void_p]
    L.ocb_halo_pack_y.argtypes = [C.c_int32, P(ocb_field), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.ocb_halo_unpack_y.argtypes = [C.c_int32, P(ocb_field), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.ocb_halo_pack_y_many.argtypes = [C.c_int32, P(ocb_field), C.c_int32, C.c_int32, P(C.c_void_p), C.c_int32, C.c_void_p]
    L.ocb_ipc_export.argtypes = [C.c_void_p, C.c_void_p, P(C.c_int64)]
    L.ocb_ipc_import.argtypes = [C.c_void_p, P(C.c_void_p)]
    L.ocb_ipc_close.argtypes = [C.c_void_p]
    L.ocb_halo_push_y_many.argtypes = [C.c_int32, P(ocb_field), C.c_int32, C.c_int32, P(C.c_void_p), P(C.c_int32), C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_void_p]
    L.ocb_halo_wait.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
    L.ocb_plan_ring_capable.argtypes = [C.c_void_p]
    L.ocb_plan_execute_ring.argtypes = [C.c_void_p, C.c_void_p, P(ocb_ring), C.c_int64, C.c_void_p]
    for name in ("ocb_plan_create", "ocb_plan_execute", "ocb_plan_destroy", "ocb_plan_info", "ocb_fill_halo",
                 "ocb_filter_execute", "ocb_bpe_reference_height", "ocb_sort_pairs", "ocb_pointwise",
                 "ocb_halo_pack_y", "ocb_halo_unpack_y", "ocb_halo_pack_y_many", "ocb_gather_sorted", "ocb_bpe_profile_lookup",
                 "ocb_ipc_export", "ocb_ipc_import", "ocb_ipc_close", "ocb_halo_push_y_many", "ocb_halo_wait",
                 "ocb_plan_ring_capable", "ocb_plan_execute_ring"):
        getattr(L, name).restype = C.c_int
    if L.ocb_abi_version() != OCB_ABI_VERSION:
        raise OcbError("liboceanostics_b200.so ABI version mismatch")
    _lib = L
    return L


EXPORTED_SYMBOLS = ["ocb_abi_version", "ocb_last_error", "ocb_device_count", "ocb_plan_create", "ocb_plan_execute",
                    "ocb_plan_destroy", "ocb_plan_info", "ocb_fill_halo", "ocb_filter_execute",
                    "ocb_bpe_reference_height", "ocb_sort_pairs", "ocb_pointwise", "ocb_halo_pack_y", "ocb_halo_unpack_y",
                    "ocb_halo_pack_y_many", "ocb_gather_sorted", "ocb_bpe_profile_lookup", "ocb_ipc_export", "ocb_ipc_import",
                    "ocb_ipc_close", "ocb_halo_push_y_many", "ocb_halo_wait", "ocb_plan_ring_capable", "ocb_plan_execute_ring"]


def check(status: int):
    """Turn a status code into the exception the reference would raise: OCB_ERR_INVALID -> ValueError
    (the reference's construction-time ArgumentError, e.g. src/Oceanostics.jl:114-120), anything else ->
    OcbError."""
    if status == 0:
        return
    msg = lib().ocb_last_error().decode("utf-8", "replace")
    if status == 1:
        raise ValueError(msg)
    raise OcbError(f"liboceanostics_b200 status {status}: {msg}")
