"""ctypes binding of the CPU kernel-logic emulator (tests/hostemu/hostemu.cc) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(os.path.join(_HERE, "libj40b_hostemu.so"))
        L.hostemu_decode.restype = C.c_uint32
        L.hostemu_decode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.hostemu_free.argtypes = [C.c_void_p]
        L.hostemu_dq_matrix.restype = C.c_int
        L.hostemu_dq_matrix.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.hostemu_natural_order.restype = C.c_int
        L.hostemu_natural_order.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.hostemu_srgb_thresholds.argtypes = [C.c_int, C.c_void_p]
        L.hostemu_srgb_lookup.restype = C.c_int
        L.hostemu_srgb_lookup.argtypes = [C.c_void_p, C.c_float]
        L.hostemu_srgb_lut_lookup.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hostemu_inverse_transform.argtypes = [C.c_int, C.c_void_p]
        L.hostemu_forward_llf.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hostemu_dump.restype = C.c_size_t
        L.hostemu_dump.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        _LIB = L
    return _LIB


def err_str(code):
    return "".join(chr((code >> s) & 0xff) for s in (24, 16, 8, 0)) if code else ""


def decode(data: bytes):
    L = lib()
    px = C.POINTER(C.c_uint8)()
    w, h, stride = C.c_int32(), C.c_int32(), C.c_int32()
    err = L.hostemu_decode(data, len(data), C.byref(px), C.byref(w), C.byref(h), C.byref(stride))
    out = None
    if px:
        raw = np.ctypeslib.as_array(px, shape=(h.value, stride.value)).copy()
        out = raw[:, : w.value * 4].reshape(h.value, w.value, 4).copy()
        L.hostemu_free(px)
    return out, err_str(err), stride.value


def last_tree_lanes() -> int:
    return int(lib().hostemu_last_tree_lanes())


def dump(data: bytes, lf_group: int, what: int, out):
    """intermediate array `what` (include/j40b.h, j40b_batch_debug_dump) of an LF group into numpy array `out`"""
    return int(lib().hostemu_dump(data, len(data), lf_group, what, out.ctypes.data, out.nbytes))
