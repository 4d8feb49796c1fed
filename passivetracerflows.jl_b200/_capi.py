"""ctypes binding of the C ABI in include/ptf_b200.h (the same binding a Julia ``ccall`` wrapper makes;
see INTEGRATION.md).  No compute happens in Python: every call below lands in libptf_b200.so."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libptf_b200.so")

# status codes
OK, EINVAL, ECUDA, ECUFFT, ENCCL, ENOMEM, EUNSUPPORTED, ENODEVICE = range(8)

STEPPER_IDS = {"ForwardEuler": 0, "RK4": 1, "ETDRK4": 2, "LSRK54": 3, "AB3": 4}
STEPPER_FILTERED = 16
FLOW_STEADY, FLOW_CALLBACK, FLOW_SEPARABLE, FLOW_LAYERED, FLOW_EXPR = 0, 1, 2, 3, 4
ENGINE_AUTO, ENGINE_CUFFT, ENGINE_FUSED = 0, 1, 2
ENGINE_NAMES = {"auto": ENGINE_AUTO, "cufft": ENGINE_CUFFT, "fused": ENGINE_FUSED}
DECOMP_NONE, DECOMP_BATCH, DECOMP_SLAB = 0, 1, 2


class PtfDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("ndim", C.c_int32),
        ("n", C.c_int64 * 3),
        ("L", C.c_double * 3),
        ("nbatch", C.c_int32),
        ("stepper", C.c_int32),
        ("kappa", C.c_double * 3),
        ("kappa_h", C.c_double),
        ("n_kappa_h", C.c_int32),
        ("dealias", C.c_int32),
        ("aliased_fraction", C.c_double),
        ("dt", C.c_double),
        ("nyquist_sign", C.c_int32),
        ("flow_kind", C.c_int32),
        ("velocity_per_batch", C.c_int32),
        ("engine", C.c_int32),
        ("device", C.c_int32),
        ("decomposition", C.c_int32),
        ("nranks", C.c_int32),
        ("rank", C.c_int32),
        ("nccl_id", C.c_uint8 * 128),
        ("filter_order", C.c_double),
        ("filter_inner_k", C.c_double),
        ("filter_outer_k", C.c_double),
        ("filter_tol", C.c_double),
        ("use_graph", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class PtfMqgDesc(C.Structure):
    """ptf_mqg_desc (include/ptf_b200.h): MultiLayerQG.Problem keyword arguments."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("nlayers", C.c_int32),
        ("nx", C.c_int64),
        ("ny", C.c_int64),
        ("Lx", C.c_double),
        ("Ly", C.c_double),
        ("f0", C.c_double),
        ("beta", C.c_double),
        ("H", C.POINTER(C.c_double)),
        ("b", C.POINTER(C.c_double)),
        ("U", C.POINTER(C.c_double)),
        ("U_is_profile", C.c_int32),
        ("n_nu", C.c_int32),
        ("eta", C.POINTER(C.c_double)),
        ("topographic_pv_gradient", C.c_double * 2),
        ("mu", C.c_double),
        ("nu", C.c_double),
        ("dt", C.c_double),
        ("stepper", C.c_int32),
        ("device", C.c_int32),
        ("aliased_fraction", C.c_double),
        ("use_graph", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


VELOCITY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double),
                          C.POINTER(C.c_double))
COEFF_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_double))

_DP = C.POINTER(C.c_double)
_H = C.c_void_p

# name -> (restype, argtypes): every symbol include/ptf_b200.h declares
SIGNATURES = {
    "ptf_version": (C.c_int32, [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ptf_device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
    "ptf_error_string": (C.c_char_p, [C.c_int32]),
    "ptf_last_error": (C.c_char_p, [_H]),
    "ptf_nccl_unique_id": (C.c_int32, [C.POINTER(C.c_uint8)]),
    "ptf_desc_init": (C.c_int32, [C.POINTER(PtfDesc)]),
    "ptf_create": (C.c_int32, [C.POINTER(PtfDesc), C.POINTER(_H)]),
    "ptf_destroy": (C.c_int32, [_H]),
    "ptf_local_shape": (C.c_int32, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int64)]),
    "ptf_set_velocity": (C.c_int32, [_H, C.c_int32, _DP, C.c_int64]),
    "ptf_set_velocity_callback": (C.c_int32, [_H, VELOCITY_FN, C.c_void_p]),
    "ptf_set_velocity_separable": (C.c_int32, [_H, C.c_int32, C.c_int32, _DP, _DP, _DP, _DP]),
    "ptf_set_coeff_callback": (C.c_int32, [_H, COEFF_FN, C.c_void_p]),
    "ptf_set_velocity_expr": (C.c_int32, [_H, C.c_int32, C.c_char_p]),
    "ptf_set_layered_velocity": (C.c_int32, [_H, _DP, _DP, _DP]),
    "ptf_set_c": (C.c_int32, [_H, _DP, C.c_int32]),
    "ptf_get_c": (C.c_int32, [_H, _DP]),
    "ptf_set_sol": (C.c_int32, [_H, _DP]),
    "ptf_get_sol": (C.c_int32, [_H, _DP]),
    "ptf_get_clock": (C.c_int32, [_H, _DP, C.POINTER(C.c_int64), _DP]),
    "ptf_set_clock": (C.c_int32, [_H, C.c_double, C.c_int64]),
    "ptf_set_dt": (C.c_int32, [_H, C.c_double]),
    "ptf_step": (C.c_int32, [_H, C.c_int64]),
    "ptf_step_until": (C.c_int32, [_H, C.c_double]),
    "ptf_step_timed": (C.c_int32, [_H, C.c_int64, C.POINTER(C.c_float)]),
    "ptf_sync": (C.c_int32, [_H]),
    "ptf_engine": (C.c_int32, [_H, C.POINTER(C.c_int32)]),
    "ptf_launch_count": (C.c_int32, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ptf_kernel_timed": (C.c_int32, [_H, C.c_char_p, C.c_int32, C.POINTER(C.c_float)]),
    "ptf_device_bytes": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "ptf_diag": (C.c_int32, [_H, _DP, _DP, _DP]),
    "ptf_selftest_fft": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _DP, _DP]),
    # MultiLayerQG flow solver
    "ptf_mqg_desc_init": (C.c_int32, [C.POINTER(PtfMqgDesc)]),
    "ptf_mqg_create": (C.c_int32, [C.POINTER(PtfMqgDesc), C.POINTER(_H)]),
    "ptf_mqg_destroy": (C.c_int32, [_H]),
    "ptf_mqg_last_error": (C.c_char_p, [_H]),
    "ptf_mqg_set_q": (C.c_int32, [_H, _DP]),
    "ptf_mqg_set_psi": (C.c_int32, [_H, _DP]),
    "ptf_mqg_set_sol": (C.c_int32, [_H, _DP]),
    "ptf_mqg_get_sol": (C.c_int32, [_H, _DP]),
    "ptf_mqg_updatevars": (C.c_int32, [_H]),
    "ptf_mqg_get_var": (C.c_int32, [_H, C.c_int32, _DP]),
    "ptf_mqg_get_background": (C.c_int32, [_H, _DP, _DP]),
    "ptf_mqg_step": (C.c_int32, [_H, C.c_int64]),
    "ptf_mqg_step_until": (C.c_int32, [_H, C.c_double]),
    "ptf_mqg_step_timed": (C.c_int32, [_H, C.c_int64, C.c_int32, C.POINTER(C.c_float)]),
    "ptf_mqg_get_clock": (C.c_int32, [_H, _DP, C.POINTER(C.c_int64), _DP]),
    "ptf_mqg_set_dt": (C.c_int32, [_H, C.c_double]),
    "ptf_mqg_launch_count": (C.c_int32, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ptf_mqg_couple": (C.c_int32, [_H, _H]),
    "ptf_mqg_forget_tracer": (C.c_int32, [_H, _H]),
    "ptf_mqg_step_coupled": (C.c_int32, [_H, _H, C.c_int64, C.POINTER(C.c_float)]),
}

_lib = None


class PtfError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"[status {status}] {message}")
        self.status = status


def _preload_nccl():
    """libptf_b200.so needs ``libnccl.so.2``.  PyTorch bundles a NEWER NCCL under the same SONAME than the system one; the
    dynamic loader keeps whichever copy is loaded first, so loading this library before ``import torch`` would pin the
    older system copy and break torch (undefined symbol ncclDevCommCreate).  Loading the bundled copy first (when there
    is one) makes the order irrelevant; the few stable entry points used here (send/recv/group/all-reduce/broadcast)
    exist in both."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    if spec is None or not spec.submodule_search_locations:
        return
    for base in spec.submodule_search_locations:
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL if hasattr(C, "RTLD_GLOBAL") else 0)
            except OSError:
                pass
            return


def load():
    """Load libptf_b200.so.  Fails loudly when the CUDA extension has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PTF_LIB_PATH") or LIB_PATH     # PTF_LIB_PATH: developer knob (A/B builds, build.py --variant)
    if not os.path.exists(path):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python passivetracerflows.jl_b200/build.py` "
                          "(__graft_entry__.build()).  This package has no CPU / PyTorch fallback.")
    _preload_nccl()
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL if hasattr(C, "RTLD_GLOBAL") else 0)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, handle=None):
    if status == OK:
        return
    lib = load()
    msg = lib.ptf_last_error(handle)
    msg = msg.decode(errors="replace") if msg else ""
    generic = lib.ptf_error_string(status).decode()
    if status == EINVAL:
        raise ValueError(f"{generic}: {msg}")   # mirrors the reference's ArgumentError (TAD.jl:234)
    raise PtfError(status, f"{generic}: {msg}")


def check_mqg(status, handle=None):
    """Same as :func:`check` for MultiLayerQG handles (their message lives in ptf_mqg_last_error)."""
    if status == OK:
        return
    lib = load()
    msg = lib.ptf_mqg_last_error(handle)
    msg = msg.decode(errors="replace") if msg else ""
    generic = lib.ptf_error_string(status).decode()
    if status == EINVAL:
        raise ValueError(f"{generic}: {msg}")
    raise PtfError(status, f"{generic}: {msg}")


def as_dp(arr):
    return arr.ctypes.data_as(_DP)
