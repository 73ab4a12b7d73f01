"""ctypes mirror of ``include/melvin_b200.h`` (the C ABI of libmelvin_b200.so).

Only declarations live here: structures, constants and argument types.  Every
function takes raw device pointers (integers); the array layer above passes
``tensor.data_ptr()`` values.  A non-zero status is turned into ``MlvError``
carrying ``mlv_last_error()``.
"""
import ctypes as C

ABI_VERSION = 2

# operator codes (include/melvin_b200.h)
OP_IDENT, OP_PSI, OP_UX, OP_UZ, OP_DDX, OP_DDZ, OP_D2DX2, OP_D2DZ2, OP_LAP, OP_INVLAP = range(10)
OP_FDM_D2DZ2, OP_FDM_DDZ, OP_FDM_NABLA2, OP_FDX_SYM = range(10, 14)
MAXLIN = 6
SYM_ONE, SYM_FDX, SYM_FDZ = range(3)
SCHEME_SI_LAP, SCHEME_SI_ARR, SCHEME_EXPLICIT = range(3)
EW_ADD, EW_SUB, EW_MUL, EW_DIV, EW_COPY, EW_POW = range(6)
KIND_REAL, KIND_CPLX, KIND_SCALAR = range(3)
RED_SUM, RED_MAX, RED_MIN, RED_SUMSQ, RED_SUMPROD = range(5)

EXPORTS = [
    "mlv_create", "mlv_destroy", "mlv_set_stream", "mlv_get_info", "mlv_long_lines", "mlv_set_sharding", "mlv_set_forward_blocks", "mlv_p2p_alloc", "mlv_p2p_copy", "mlv_p2p_open",
    "mlv_p2p_close", "mlv_set_peer_buffers", "mlv_set_peer_flags", "mlv_last_error",
    "mlv_abi_version", "mlv_launch_count", "mlv_to_physical", "mlv_to_spectral", "mlv_x_inverse",
    "mlv_z_inverse", "mlv_z_forward", "mlv_x_forward", "mlv_advect_z", "mlv_advect_z_rows", "mlv_set_reduction_partials", "mlv_set_reductions", "mlv_reduce_partials", "mlv_advect_phys",
    "mlv_spec_lincomb", "mlv_lap_array", "mlv_stencil", "mlv_solve_fdm", "mlv_solve_fdm_o4", "mlv_fdm_velocity",
    "mlv_fdm_advect", "mlv_integrate",
    "mlv_elementwise", "mlv_reduce", "mlv_trig_axis",
]
EXT_PERIODIC, EXT_EVEN, EXT_ODD = range(3)


class MlvError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nz", C.c_int32), ("fdm_z", C.c_int32),
                ("fd_order", C.c_int32), ("lx", C.c_double), ("lz", C.c_double),
                ("kx0", C.c_double), ("kz0", C.c_double), ("d2x", C.c_double),
                ("d2z", C.c_double)]


class Info(C.Structure):
    _fields_ = [("nn", C.c_int32), ("nm", C.c_int32), ("spec_rows", C.c_int32),
                ("spec_cols", C.c_int32), ("ipitch", C.c_int32), ("nm_local", C.c_int32),
                ("ibytes", C.c_int64), ("red_doubles", C.c_int64)]


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("row_stride", C.c_int64), ("col_stride", C.c_int64)]


class LinTerms(C.Structure):
    _fields_ = [("n", C.c_int32), ("op", C.c_int32 * MAXLIN), ("src", C.c_void_p * MAXLIN),
                ("cre", C.c_double * MAXLIN), ("cim", C.c_double * MAXLIN)]


class Integ(C.Structure):
    _fields_ = [("ab_order", C.c_int32), ("scheme", C.c_int32), ("dt", C.c_double),
                ("alpha", C.c_double), ("lcoef", C.c_double), ("larr", C.c_void_p),
                ("q_in", C.c_void_p), ("q_out", C.c_void_p), ("f0", C.c_void_p),
                ("f0_set", C.c_int32), ("reserved_", C.c_int32),
                ("fm1", C.c_void_p), ("fm2", C.c_void_p), ("fm3", C.c_void_p)]


class XFwd(C.Structure):
    _fields_ = [("nf", C.c_int32), ("mode", C.c_int32), ("src", C.c_void_p * 4),
                ("sym", C.c_int32 * 4), ("coef", C.c_double * 4), ("dst", C.c_void_p),
                ("lin", LinTerms), ("integ", Integ)]


class Ew(C.Structure):
    _fields_ = [("op", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32),
                ("out_kind", C.c_int32), ("a_kind", C.c_int32), ("b_kind", C.c_int32),
                ("out", View), ("a", View), ("b", View),
                ("a_re", C.c_double), ("a_im", C.c_double),
                ("b_re", C.c_double), ("b_im", C.c_double)]


class Trig(C.Structure):
    _fields_ = [("inverse", C.c_int32), ("ext", C.c_int32), ("period", C.c_int32),
                ("n_samp", C.c_int32), ("n_modes", C.c_int32), ("two_sided", C.c_int32),
                ("hermitian", C.c_int32), ("samp_complex", C.c_int32), ("batch_fastest", C.c_int32),
                ("nbatch", C.c_int32),
                ("samp_stride", C.c_int64), ("samp_batch_stride", C.c_int64),
                ("mode_stride", C.c_int64), ("mode_batch_stride", C.c_int64),
                ("in_", C.c_void_p), ("out", C.c_void_p),
                ("scale_re", C.c_double), ("scale_im", C.c_double), ("w0", C.c_double)]


def make_lin_terms(terms):
    """terms: iterable of (coef: complex, op: int, src_ptr: int)"""
    lt = LinTerms()
    terms = list(terms)
    if len(terms) > MAXLIN:
        raise MlvError(f"at most {MAXLIN} linear terms per call")
    lt.n = len(terms)
    for i, (coef, op, ptr) in enumerate(terms):
        coef = complex(coef)
        lt.op[i] = op
        lt.src[i] = ptr
        lt.cre[i] = coef.real
        lt.cim[i] = coef.imag
    return lt


def declare(lib):
    """Attach argtypes / restypes to a loaded library and verify its ABI."""
    vp, i32, f64 = C.c_void_p, C.c_int, C.c_double
    sig = {
        "mlv_create": [C.POINTER(Params), C.POINTER(vp)],
        "mlv_destroy": [vp],
        "mlv_set_stream": [vp, vp],
        "mlv_get_info": [vp, C.POINTER(Info)],
        "mlv_long_lines": [vp],
        "mlv_set_sharding": [vp, i32, i32, i32, i32],
        "mlv_set_forward_blocks": [vp, i32],
        "mlv_p2p_alloc": [vp, C.c_int64, C.POINTER(vp), vp],
        "mlv_p2p_open": [vp, vp, C.POINTER(vp)],
        "mlv_p2p_close": [vp, vp, i32],
        "mlv_p2p_copy": [vp, vp, vp, C.c_int64, vp],
        "mlv_set_peer_buffers": [vp, i32, C.POINTER(vp)],
        "mlv_set_peer_flags": [vp, C.POINTER(vp)],
        "mlv_abi_version": [],
        "mlv_to_physical": [vp, vp, vp, vp],
        "mlv_to_spectral": [vp, vp, vp, vp],
        "mlv_x_inverse": [vp, i32, C.POINTER(vp), C.POINTER(C.c_int32), C.POINTER(vp)],
        "mlv_z_inverse": [vp, vp, vp],
        "mlv_z_forward": [vp, vp, vp],
        "mlv_x_forward": [vp, C.POINTER(XFwd)],
        "mlv_advect_z": [vp, vp, vp, vp, vp, vp, vp],
        "mlv_advect_z_rows": [vp, vp, vp, vp, vp, vp, i32, i32, vp],
        "mlv_set_reduction_partials": [vp, vp],
        "mlv_set_reductions": [vp, i32],
        "mlv_reduce_partials": [vp, vp, vp],
        "mlv_advect_phys": [vp, vp, vp, vp, vp],
        "mlv_spec_lincomb": [vp, C.POINTER(LinTerms), vp],
        "mlv_lap_array": [vp, f64, vp],
        "mlv_stencil": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f64],
        "mlv_solve_fdm": [vp, vp, vp],
        "mlv_solve_fdm_o4": [vp, vp, vp],
        "mlv_fdm_velocity": [vp, vp, vp, vp, vp],
        "mlv_fdm_advect": [vp, vp, vp, vp, vp, vp, vp],
        "mlv_integrate": [vp, C.POINTER(LinTerms), C.POINTER(Integ)],
        "mlv_elementwise": [vp, C.POINTER(Ew)],
        "mlv_reduce": [vp, i32, i32, i32, C.POINTER(View), C.POINTER(View), vp],
        "mlv_trig_axis": [vp, C.POINTER(Trig)],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.mlv_last_error.argtypes = []
    lib.mlv_last_error.restype = C.c_char_p
    lib.mlv_launch_count.argtypes = []
    lib.mlv_launch_count.restype = C.c_longlong
    if lib.mlv_abi_version() != ABI_VERSION:
        raise MlvError("libmelvin_b200 ABI version mismatch")
    return lib


def check(lib, status):
    if status != 0:
        msg = lib.mlv_last_error()
        raise MlvError(f"libmelvin_b200 error {status}: {msg.decode() if msg else '?'}")
