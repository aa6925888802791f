"""ctypes binding of libgtb200.so (the C ABI declared in include/gtb200.h).

The library is the product; this module only loads it.  There is no fallback of any kind: if the shared object is
missing, importing this module raises, and without a CUDA device every compute entry point returns GTB_ERR_CUDA
which `check` turns into RuntimeError (the reference throws std::runtime_error from GT_CUDA_CHECK,
common/cuda_util.hpp:20-35).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgtb200.so")

GTB_BC_VALUE, GTB_BC_COPY = 0, 1
GTB_OK, GTB_ERR_ARG, GTB_ERR_LAYOUT, GTB_ERR_CUDA, GTB_ERR_ALLOC, GTB_ERR_STATE = range(6)
HALO_BLOB_BYTES = 512


class Field(C.Structure):
    """gtb_field: pointer to element (0,0,0) of the compute domain + element strides."""

    _fields_ = [("ptr", C.c_void_p), ("stride_i", C.c_int64), ("stride_j", C.c_int64), ("stride_k", C.c_int64)]


class HaloDesc(C.Structure):
    """gtb_halo_desc == common/halo_descriptor.hpp:44-227 (minus, plus, begin, end inclusive, total)."""

    _fields_ = [("minus", C.c_int), ("plus", C.c_int), ("begin", C.c_int), ("end", C.c_int), ("total", C.c_int)]


class HaloField(C.Structure):
    """gtb_halo_field: a field with its own three halo descriptors (field_on_the_fly)."""

    _fields_ = [("ptr", C.c_void_p), ("desc", HaloDesc * 3)]


class GtbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libgtb200 status %d: %s" % (status, message))
        self.status = status


# every symbol include/gtb200.h declares: name -> (restype, argtypes)
_FP = C.POINTER(Field)
_SIG = {
    "gtb_version": (C.c_int, []),
    "gtb_last_error": (C.c_char_p, []),
    "gtb_device_count": (C.c_int, []),
    "gtb_init": (C.c_int, [C.c_int]),
    "gtb_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gtb_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "gtb_get_option": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "gtb_release_scratch": (C.c_int, []),
    "gtb_launch_count": (C.c_int64, []),
    "gtb_last_kernel": (C.c_char_p, []),
    "gtb_copy_box_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p]),
    "gtb_debug_trace": (C.c_int, [C.c_void_p, C.c_int64]),
    "gtb_copy": (C.c_int, [_FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_hori_diff_f64": (C.c_int, [_FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_hori_diff_f32": (C.c_int, [_FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_simple_hori_diff_f64": (C.c_int, [_FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_simple_hori_diff_f32": (C.c_int, [_FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_vert_adv_f64": (C.c_int, [_FP, _FP, _FP, _FP, _FP, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_vert_adv_f32": (C.c_int, [_FP, _FP, _FP, _FP, _FP, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_tridiagonal_f64": (C.c_int, [_FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_prepare_tracers_f64": (C.c_int, [_FP, _FP, C.c_int, _FP, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_halo_create": (C.c_int, [C.POINTER(HaloDesc), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_void_p)]),
    "gtb_halo_destroy": (C.c_int, [C.c_void_p]),
    "gtb_halo_send_bytes": (C.c_int64, [C.c_void_p, C.c_int, C.c_int]),
    "gtb_halo_recv_bytes": (C.c_int64, [C.c_void_p, C.c_int, C.c_int]),
    "gtb_halo_send_buffer": (C.c_void_p, [C.c_void_p, C.c_int]),
    "gtb_halo_recv_buffer": (C.c_void_p, [C.c_void_p, C.c_int]),
    "gtb_halo_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gtb_halo_connect": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "gtb_halo_pack": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "gtb_halo_pack_send": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "gtb_halo_send": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gtb_halo_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gtb_halo_unpack": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "gtb_halo_wait_unpack": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "gtb_halo_exchange": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "gtb_halo_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "gtb_halo_next_epoch": (C.c_int, [C.c_void_p]),
    "gtb_halo_poll_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "gtb_halo_set_trace": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gtb_stamp": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gtb_halo_generic_pack_send": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gtb_halo_generic_wait_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gtb_boundary_apply": (C.c_int, [C.POINTER(HaloDesc), C.POINTER(C.c_int), C.c_int, C.c_double, C.POINTER(C.c_void_p),
                                     C.c_int, C.c_int, C.c_void_p]),
    "gtb_halo_set_boundary": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "gtb_halo_unpacked_flag": (C.c_void_p, [C.c_void_p]),
    "gtb_halo_epoch": (C.c_uint64, [C.c_void_p]),
    "gtb_stencil_gate": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "gtb_halo_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "gtb_gate_timeouts": (C.c_int, [C.POINTER(C.c_int64)]),
    "gtb_seq_add_stencil_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "gtb_seq_add_halo_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "gtb_seq_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "gtb_seq_destroy": (C.c_int, [C.c_void_p]),
    "gtb_seq_size": (C.c_int, [C.c_void_p]),
    "gtb_seq_add_hori_diff": (C.c_int, [C.c_void_p, C.c_int, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_seq_add_vert_adv": (C.c_int, [C.c_void_p, C.c_int, _FP, _FP, _FP, _FP, _FP, C.c_double, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p]),
    "gtb_seq_add_prepare_tracers": (C.c_int, [C.c_void_p, _FP, _FP, C.c_int, _FP, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gtb_seq_add_halo_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "gtb_seq_add_record": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gtb_seq_add_wait": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gtb_seq_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "gtb_tensor_map_3d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                          C.POINTER(C.c_int)]),
    "gtb_device_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "gtb_device_free": (C.c_int, [C.c_void_p]),
    "gtb_host_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "gtb_host_free": (C.c_int, [C.c_void_p]),
    "gtb_staged_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "gtb_staged_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "gtb_stream_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "gtb_stream_destroy": (C.c_int, [C.c_void_p]),
    "gtb_stream_after_default": (C.c_int, [C.c_void_p]),
    "gtb_stream_synchronize": (C.c_int, [C.c_void_p]),
    "gtb_seq_add_stamp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gtb_seq_add_mark": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gtb_seq_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
}

_lib = None


def lib():
    """The loaded library.  Raises ImportError (never falls back) when libgtb200.so has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m gridtools_b200.build` (needs nvcc). "
                "gridtools_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIG.items():
            fn = getattr(handle, name)  # AttributeError if the C ABI and this table ever diverge
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIG)


def last_error():
    return lib().gtb_last_error().decode("utf-8", "replace")


def check(status):
    if status != GTB_OK:
        raise GtbError(status, last_error())


def set_option(key, value):
    check(lib().gtb_set_option(key.encode(), int(value)))


def get_option(key):
    v = C.c_int()
    check(lib().gtb_get_option(key.encode(), C.byref(v)))
    return v.value


def gate_timeouts():
    n = C.c_int64()
    check(lib().gtb_gate_timeouts(C.byref(n)))
    return n.value


def last_kernel():
    """Name of the kernel this thread launched last (which variant the automatic choice took)."""
    return lib().gtb_last_kernel().decode()


def launch_count():
    return int(lib().gtb_launch_count())


def device_info():
    sm, l2, hbm = C.c_int(), C.c_int64(), C.c_int64()
    check(lib().gtb_device_info(C.byref(sm), C.byref(l2), C.byref(hbm)))
    return {"sm_count": sm.value, "l2_bytes": l2.value, "hbm_bytes": hbm.value}
