"""ctypes binding of oracle/_ref/libj40ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the unmodified reference j40.h compiled by oracle/Makefile (see oracle/ref_harness.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_ref", "libj40ref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libj40ref.so missing: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(path)
        L.ref_decode_rgba.restype = C.c_int
        L.ref_decode_rgba.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(C.c_uint8)),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_uint32), C.c_char_p]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_time_decode.restype = C.c_int
        L.ref_time_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        L.ref_staged_open.restype = C.c_void_p
        L.ref_staged_open.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32)]
        L.ref_staged_info.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.ref_staged_lf_group_info.restype = C.c_int
        L.ref_staged_lf_group_info.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]
        L.ref_staged_lf_group_array.restype = C.c_int64
        L.ref_staged_lf_group_array.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64]
        L.ref_staged_advance.restype = C.c_uint32
        L.ref_staged_advance.argtypes = [C.c_void_p, C.c_int]
        L.ref_staged_plane_i16.restype = C.c_int64
        L.ref_staged_plane_i16.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.ref_staged_close.argtypes = [C.c_void_p]
        L.ref_inverse_transform.argtypes = [C.c_int, C.c_void_p]
        L.ref_forward_llf.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_constants.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_default_dq_matrix.restype = C.c_int
        L.ref_default_dq_matrix.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.ref_natural_order.restype = C.c_int
        L.ref_natural_order.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.ref_srgb_quant.restype = C.c_int32
        L.ref_srgb_quant.argtypes = [C.c_float, C.c_int]
        _LIB = L
    return _LIB


def err_str(code):
    return "".join(chr((code >> s) & 0xff) for s in (24, 16, 8, 0)) if code else ""


def decode(data: bytes):
    """Decode through the reference's public API. Returns (rgba[h,w,4] uint8 or None, err4, errstring, stride)."""
    L = lib()
    px = C.POINTER(C.c_uint8)()
    w, h, stride = C.c_int32(), C.c_int32(), C.c_int32()
    err = C.c_uint32()
    msg = C.create_string_buffer(256)
    ok = L.ref_decode_rgba(data, len(data), C.byref(px), C.byref(w), C.byref(h), C.byref(stride), C.byref(err), msg)
    out = None
    if ok:
        raw = np.ctypeslib.as_array(px, shape=(h.value, stride.value)).copy()
        out = raw[:, : w.value * 4].reshape(h.value, w.value, 4).copy()
    if px:
        L.ref_free(px)
    return out, err_str(err.value), msg.value.decode("latin1"), stride.value


def time_decode(data: bytes, reps: int):
    L = lib()
    secs = (C.c_double * reps)()
    good = L.ref_time_decode(data, len(data), reps, secs)
    return good, list(secs)


class Staged:
    """Step-by-step replay of the reference's j40__advance with access to its intermediates."""
    WHAT = dict(blocks=0, varblocks=1, lfindices=2, llf_x=3, llf_y=4, llf_b=5, coef_x=6, coef_y=7, coef_b=8,
                xfromy=9, bfromy=10, sharpness=11)

    def __init__(self, data: bytes):
        self._data = data  # keep alive: the reference does not copy
        err = C.c_uint32()
        self._h = lib().ref_staged_open(data, len(data), C.byref(err))
        self.err = err_str(err.value)
        info = (C.c_int64 * 16)()
        lib().ref_staged_info(self._h, info)
        keys = ["width", "height", "is_modular", "group_size_shift", "num_groups", "num_lf_groups", "num_passes",
                "global_scale", "quant_lf", "nb_block_ctx", "num_hf_presets", "bpp", "xyb_encoded",
                "dct_select_used", "order_used", "num_extra_channels"]
        self.info = dict(zip(keys, list(info)))

    def lf_group_info(self, gg):
        out = (C.c_int32 * 8)()
        if not lib().ref_staged_lf_group_info(self._h, gg, out):
            return None
        keys = ["left", "top", "width", "height", "width8", "height8", "nb_varblocks", "loaded"]
        return dict(zip(keys, list(out)))

    def lf_group_array(self, gg, name):
        gi = self.lf_group_info(gg)
        what = self.WHAT[name]
        w8, h8 = gi["width8"], gi["height8"]
        if what == 0:
            arr = np.zeros((h8, w8), np.int32)
        elif what == 1:
            arr = np.zeros((gi["nb_varblocks"], 2), np.int32)
        elif what == 2:
            arr = np.zeros((h8, w8), np.uint8)
        elif what in (3, 4, 5):
            arr = np.zeros(w8 * h8, np.float32)
        elif what in (6, 7, 8):
            arr = np.zeros(w8 * h8 * 64, np.float32)
        elif what in (9, 10):
            arr = np.zeros(((gi["height"] + 63) // 64, (gi["width"] + 63) // 64), np.int16)
        else:
            arr = np.zeros((h8, w8), np.int16)
        n = lib().ref_staged_lf_group_array(self._h, gg, what, arr.ctypes.data, arr.nbytes)
        if n != arr.nbytes:
            raise RuntimeError(f"oracle array {name} unavailable (got {n}, want {arr.nbytes})")
        return arr

    def advance(self, stage):
        e = lib().ref_staged_advance(self._h, stage)
        self.err = err_str(e)
        return self.err

    def plane_i16(self, c):
        w, h = self.info["width"], self.info["height"]
        arr = np.zeros((h, w), np.int16)
        n = lib().ref_staged_plane_i16(self._h, c, arr.ctypes.data, arr.size)
        if n != arr.size:
            raise RuntimeError("oracle plane unavailable")
        return arr

    def close(self):
        if self._h:
            lib().ref_staged_close(self._h)
            self._h = None

    def __del__(self):
        self.close()


def inverse_transform(dctsel, coeffs):
    buf = np.ascontiguousarray(coeffs, np.float32).copy()
    lib().ref_inverse_transform(dctsel, buf.ctypes.data)
    return buf


def forward_llf(block, log_rows, log_columns):
    buf = np.ascontiguousarray(block, np.float32).copy()
    lib().ref_forward_llf(buf.ctypes.data, log_rows, log_columns)
    return buf


def constants():
    hs = np.zeros(256, np.float32); sc = np.zeros(64, np.float32); afv = np.zeros(256, np.float32)
    lib().ref_constants(hs.ctypes.data, sc.ctypes.data, afv.ctypes.data)
    return hs, sc, afv


def default_dq_matrix(idx):
    buf = np.zeros(65536 * 3, np.float32)
    n = lib().ref_default_dq_matrix(idx, buf.ctypes.data, buf.size)
    return buf[: n * 3].reshape(n, 3).copy()


def natural_order(log_rows, log_columns):
    out = np.zeros(1 << (log_rows + log_columns), np.int32)
    lib().ref_natural_order(log_rows, log_columns, out.ctypes.data)
    return out
