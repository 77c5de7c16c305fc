"""ctypes binding of the C-ABI in include/dde_b200.h.

The shared library is the product: if it is missing or cannot be loaded this
module raises -- there is no Python or CPU fallback for the solvers.
"""
import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("DDE_B200_LIB", _HERE / "libdde_b200.so"))

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint64_p = C.POINTER(C.c_uint64)


class DopriControl(C.Structure):
    """struct dde_dopri_control (include/dde_b200.h)."""

    _fields_ = [
        ("atol", C.c_double),
        ("rtol", C.c_double),
        ("step_size_min", C.c_double),
        ("step_size_max", C.c_double),
        ("step_size_initial", C.c_double),
        ("step_max_n", C.c_uint64),
        ("stiff_check", C.c_uint64),
        ("n_history", C.c_uint64),
        ("method", C.c_int32),
        ("step_size_min_allow", C.c_int32),
        ("grow_history", C.c_int32),
        ("return_initial", C.c_int32),
    ]


# every symbol include/dde_b200.h declares: (name, restype, argtypes)
_SZ = C.c_size_t
_DOPRI_BATCH_ARGS = [
    C.c_char_p, C.POINTER(DopriControl), _SZ, _SZ, _SZ, c_double_p, c_double_p, _SZ, _SZ,
    c_double_p, _SZ, c_double_p, _SZ, c_double_p, c_double_p, c_int32_p, c_int32_p, c_double_p,
    c_double_p,
]
_DIFEQ_BATCH_ARGS = [
    C.c_char_p, _SZ, _SZ, _SZ, c_double_p, c_double_p, _SZ, _SZ, c_uint64_p, _SZ, _SZ, C.c_int32,
    c_double_p, c_double_p, c_int32_p, c_double_p,
]
SYMBOLS = [
    ("dde_dopri_control_default", None, [C.POINTER(DopriControl), C.c_int32]),
    ("dde_last_error", C.c_char_p, []),
    ("dde_version", C.c_char_p, []),
    ("dde_model_count", _SZ, []),
    ("dde_model_name", C.c_char_p, [_SZ]),
    ("dde_model_info", C.c_int, [C.c_char_p, c_int32_p, C.POINTER(_SZ), C.POINTER(_SZ),
                                 C.POINTER(_SZ)]),
    ("dde_dopri_strerror", C.c_int, [C.c_int32, C.c_double, C.c_uint64, C.c_char_p, _SZ]),
    ("dde_dopri_batch", C.c_int, _DOPRI_BATCH_ARGS),
    ("dde_dopri_batch_alloc", C.c_void_p, [C.c_char_p, C.POINTER(DopriControl), _SZ, _SZ, _SZ,
                                           _SZ]),
    ("dde_dopri_batch_free", None, [C.c_void_p]),
    ("dde_dopri_batch_upload", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _SZ]),
    ("dde_dopri_batch_set_times", C.c_int, [C.c_void_p, c_double_p, _SZ, c_double_p, _SZ]),
    ("dde_dopri_batch_set_events", C.c_int, [C.c_void_p, c_int32_p, _SZ]),
    ("dde_dopri_batch_set_order", C.c_int, [C.c_void_p, C.c_void_p]),
    ("dde_dopri_batch_order_by_last_run", C.c_int, [C.c_void_p]),
    ("dde_dopri_batch_run", C.c_int, [C.c_void_p]),
    ("dde_dopri_batch_continue", C.c_int, [C.c_void_p, C.c_void_p, c_double_p, _SZ, c_double_p,
                                           _SZ]),
    ("dde_dopri_batch_copy", C.c_void_p, [C.c_void_p]),
    ("dde_dopri_batch_sync", C.c_int, [C.c_void_p]),
    ("dde_dopri_batch_download", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    ("dde_dopri_batch_download_range", C.c_int, [C.c_void_p, _SZ, _SZ, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("dde_dopri_batch_run_download", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p]),
    ("dde_dopri_batch_n_history", C.c_uint64, [C.c_void_p]),
    ("dde_dopri_batch_last_ms", C.c_double, [C.c_void_p]),
    ("dde_dopri_batch_last_launches", C.c_int, [C.c_void_p]),
    ("dde_dopri_batch_history", C.c_long, [C.c_void_p, _SZ, c_double_p, _SZ]),
    ("dde_dopri_batch_device_ptr", C.c_void_p, [C.c_void_p, C.c_int]),
    ("dde_dopri_interpolate_batch", C.c_int, [c_double_p, _SZ, _SZ, _SZ, _SZ, c_double_p, _SZ,
                                              c_double_p]),
    ("dde_dopri_batch_interpolate", C.c_int, [C.c_void_p, c_double_p, _SZ, c_double_p]),
    ("dde_difeq_batch", C.c_int, _DIFEQ_BATCH_ARGS),
    ("dde_difeq_batch_alloc", C.c_void_p, [C.c_char_p, _SZ, _SZ, _SZ, _SZ, _SZ]),
    ("dde_difeq_batch_free", None, [C.c_void_p]),
    ("dde_difeq_batch_upload", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _SZ]),
    ("dde_difeq_batch_set_steps", C.c_int, [C.c_void_p, c_uint64_p, _SZ, C.c_int32]),
    ("dde_difeq_batch_run", C.c_int, [C.c_void_p]),
    ("dde_difeq_batch_continue", C.c_int, [C.c_void_p, C.c_void_p, c_uint64_p, _SZ, C.c_int32]),
    ("dde_difeq_batch_copy", C.c_void_p, [C.c_void_p]),
    ("dde_difeq_batch_restartable", C.c_int, [C.c_void_p, C.c_int32]),
    ("dde_difeq_batch_sync", C.c_int, [C.c_void_p]),
    ("dde_difeq_batch_download", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    ("dde_difeq_batch_history", C.c_long, [C.c_void_p, _SZ, c_double_p, _SZ]),
    ("dde_difeq_batch_last_ms", C.c_double, [C.c_void_p]),
    ("dde_device_count", C.c_int, []),
    ("dde_dopri_batch_sharded", C.c_int, _DOPRI_BATCH_ARGS + [c_int32_p, _SZ]),
    ("dde_dopri_batch_sharded_events", C.c_int,
     _DOPRI_BATCH_ARGS[:13] + [c_int32_p] + _DOPRI_BATCH_ARGS[13:] + [c_int32_p, _SZ]),
    ("dde_difeq_batch_sharded", C.c_int, _DIFEQ_BATCH_ARGS + [c_int32_p, _SZ]),
    ("dde_dopri_shards_alloc", C.c_void_p, [C.c_char_p, C.POINTER(DopriControl), _SZ, _SZ, _SZ, _SZ,
                                            c_int32_p, _SZ]),
    ("dde_dopri_shards_free", None, [C.c_void_p]),
    ("dde_dopri_shards_count", _SZ, [C.c_void_p]),
    ("dde_dopri_shards_upload", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _SZ]),
    ("dde_dopri_shards_set_times", C.c_int, [C.c_void_p, c_double_p, _SZ, c_double_p, _SZ]),
    ("dde_dopri_shards_set_events", C.c_int, [C.c_void_p, c_int32_p, _SZ]),
    ("dde_dopri_shards_order_by_last_run", C.c_int, [C.c_void_p, C.c_int32]),
    ("dde_dopri_shards_run_download", C.c_int, [C.c_void_p] + [C.c_void_p] * 6),
    ("dde_dopri_shards_last_ms", C.c_double, [C.c_void_p]),
    ("dde_shard_range", None, [_SZ, _SZ, _SZ, C.POINTER(_SZ), C.POINTER(_SZ)]),
    ("dde_host_alloc", C.c_void_p, [_SZ]),
    ("dde_host_free", None, [C.c_void_p]),
    ("dde_pow_host", C.c_double, [C.c_double, C.c_double]),
    ("dde_pow2_host", None, [C.c_double, C.c_double, C.c_double, c_double_p]),
    ("dde_exp_host", C.c_double, [C.c_double]),
    ("dde_fp64_peak", C.c_double, [C.c_int]),
]

_lib = None


def load():
    """Load libdde_b200.so (once) and declare the signatures.  Raises if the
    library has not been built: run ./build.sh or __graft_entry__.build()."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            "%s not found: the CUDA extension is the product and has no fallback; "
            "build it with ./build.sh" % LIB_PATH)
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    for path in filter(None, os.environ.get("DDE_B200_MODEL_LIBS", "").split(os.pathsep)):
        load_model_library(path)
    return lib


_model_libs = {}


def load_model_library(path):
    """dyn.load() for a dde_b200 model library (R/common.R:89-98): a shared object built
    from a model translation unit (include/dde_b200/dde_model.cuh).  Its
    static initialisers register its models by name with the product library
    (dde_register_model); they are then available to every entry point.  Libraries named
    in DDE_B200_MODEL_LIBS (os.pathsep-separated) are loaded with the product library."""
    load()
    path = str(Path(path).resolve())
    if path not in _model_libs:
        _model_libs[path] = C.CDLL(path, mode=C.RTLD_GLOBAL)
    return _model_libs[path]


def last_error():
    return load().dde_last_error().decode()
