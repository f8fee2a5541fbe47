"""j40_b200 -- Python binding of libj40b200.so, the Blackwell-native JPEG XL decoder behind the j40 C API.

The binding mirrors the reference's public interface (j40.h:233-272): `Image.from_memory / from_file`,
`output_format`, `next_frame`, `current_frame`, `frame_pixels_u8x4`, `error`, `error_string`, `free`,
plus the batch extension of include/j40b.h (`Batch`). Everything runs in the CUDA library; there is no
Python or CPU decoding path -- importing works without a GPU (so that build checks can load the
library and list its symbols) but every decode call fails with error code `!gpu`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libj40b200.so")

J40_U8X4 = 0x0F33
J40_RGBA = 0x1755

# every symbol include/j40.h and include/j40b.h declare
EXPORTED_SYMBOLS = [
    "j40_error", "j40_error_string", "j40_from_memory", "j40_from_file", "j40_output_format", "j40_next_frame",
    "j40_current_frame", "j40_frame_pixels_u8x4", "j40_row_u8x4", "j40_free",
    "j40b_batch_create", "j40b_batch_destroy", "j40b_batch_add", "j40b_batch_upload", "j40b_batch_decode",
    "j40b_batch_wait", "j40b_batch_count", "j40b_batch_error", "j40b_batch_info", "j40b_batch_device_pixels",
    "j40b_batch_read_pixels", "j40b_batch_last_decode_ms", "j40b_batch_kernel_ms", "j40b_batch_stat", "j40b_gpu_available",
    "j40b_batch_mark", "j40b_batch_join", "j40b_batch_mark_ms", "j40b_batch_reset", "j40b_batch_read_all_async",
    "j40b_batch_add_many", "j40b_batch_event_ms", "j40b_batch_after", "j40b_batch_debug_dump", "j40b_batch_write_pam",
]


class _ImageU(C.Union):
    _fields_ = [("inner", C.c_void_p), ("err", C.c_uint32), ("saved_errno", C.c_int)]


class j40_image(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("u", _ImageU)]


class j40_frame(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("reserved", C.c_uint32), ("inner", C.c_void_p)]


class j40_pixels_u8x4(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("stride_bytes", C.c_int32), ("data", C.c_void_p)]


_LIB = None


def lib():
    """Loads libj40b200.so (fails loudly if it has not been built)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make -C j40_b200/csrc` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        L.j40_error.restype = C.c_uint32
        L.j40_error.argtypes = [C.POINTER(j40_image)]
        L.j40_error_string.restype = C.c_char_p
        L.j40_error_string.argtypes = [C.POINTER(j40_image)]
        L.j40_from_memory.restype = C.c_uint32
        L.j40_from_memory.argtypes = [C.POINTER(j40_image), C.c_void_p, C.c_size_t, C.c_void_p]
        L.j40_from_file.restype = C.c_uint32
        L.j40_from_file.argtypes = [C.POINTER(j40_image), C.c_char_p]
        L.j40_output_format.restype = C.c_uint32
        L.j40_output_format.argtypes = [C.POINTER(j40_image), C.c_int32, C.c_int32]
        L.j40_next_frame.restype = C.c_int
        L.j40_next_frame.argtypes = [C.POINTER(j40_image)]
        L.j40_current_frame.restype = j40_frame
        L.j40_current_frame.argtypes = [C.POINTER(j40_image)]
        L.j40_frame_pixels_u8x4.restype = j40_pixels_u8x4
        L.j40_frame_pixels_u8x4.argtypes = [C.POINTER(j40_frame), C.c_int32]
        L.j40_row_u8x4.restype = C.c_void_p
        L.j40_row_u8x4.argtypes = [j40_pixels_u8x4, C.c_int32]
        L.j40_free.argtypes = [C.POINTER(j40_image)]
        L.j40b_batch_create.restype = C.c_void_p
        L.j40b_batch_create.argtypes = [C.c_int]
        L.j40b_batch_destroy.argtypes = [C.c_void_p]
        L.j40b_batch_add.restype = C.c_int
        L.j40b_batch_add.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        for name in ("j40b_batch_upload", "j40b_batch_decode", "j40b_batch_wait", "j40b_batch_count"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p]
        L.j40b_batch_error.restype = C.c_uint32
        L.j40b_batch_error.argtypes = [C.c_void_p, C.c_int]
        L.j40b_batch_info.restype = C.c_int
        L.j40b_batch_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.j40b_batch_device_pixels.restype = C.c_void_p
        L.j40b_batch_device_pixels.argtypes = [C.c_void_p, C.c_int]
        L.j40b_batch_read_pixels.restype = C.c_int
        L.j40b_batch_read_pixels.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.j40b_batch_last_decode_ms.restype = C.c_float
        L.j40b_batch_last_decode_ms.argtypes = [C.c_void_p]
        L.j40b_batch_kernel_ms.restype = C.c_float
        L.j40b_batch_kernel_ms.argtypes = [C.c_void_p, C.c_int]
        L.j40b_batch_after.restype = C.c_int
        L.j40b_batch_after.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.j40b_batch_event_ms.restype = C.c_float
        L.j40b_batch_event_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.j40b_batch_stat.restype = C.c_int64
        L.j40b_batch_stat.argtypes = [C.c_void_p, C.c_int]
        L.j40b_gpu_available.restype = C.c_int
        L.j40b_batch_mark.restype = C.c_int
        L.j40b_batch_mark.argtypes = [C.c_void_p, C.c_int]
        L.j40b_batch_add_many.restype = C.c_int
        L.j40b_batch_add_many.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, C.c_int]
        L.j40b_batch_reset.restype = C.c_int
        L.j40b_batch_reset.argtypes = [C.c_void_p]
        L.j40b_batch_read_all_async.restype = C.c_int
        L.j40b_batch_read_all_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.j40b_batch_join.restype = C.c_int
        L.j40b_batch_join.argtypes = [C.c_void_p, C.c_void_p]
        L.j40b_batch_mark_ms.restype = C.c_float
        L.j40b_batch_mark_ms.argtypes = [C.c_void_p]
        L.j40b_batch_write_pam.restype = C.c_int
        L.j40b_batch_write_pam.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.j40b_batch_debug_dump.restype = C.c_size_t
        L.j40b_batch_debug_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        _LIB = L
    return _LIB


def err_str(code):
    return "".join(chr((code >> s) & 0xFF) for s in (24, 16, 8, 0)) if code else ""


def gpu_available():
    return bool(lib().j40b_gpu_available())


class Image:
    """One image handle, used exactly like the C API (and like the reference's README example)."""

    def __init__(self):
        self._img = j40_image()
        self._buf = None
        self._open = False

    @classmethod
    def from_memory(cls, data: bytes):
        self = cls()
        self._buf = C.create_string_buffer(bytes(data), len(data))  # must outlive the handle (the C API does not copy)
        self.status = lib().j40_from_memory(C.byref(self._img), C.cast(self._buf, C.c_void_p), len(data), None)
        self._open = True
        return self

    @classmethod
    def from_file(cls, path):
        self = cls()
        self.status = lib().j40_from_file(C.byref(self._img), os.fsencode(path))
        self._open = True
        return self

    def output_format(self, channel=J40_RGBA, fmt=J40_U8X4):
        return lib().j40_output_format(C.byref(self._img), channel, fmt)

    def next_frame(self):
        return bool(lib().j40_next_frame(C.byref(self._img)))

    def current_frame(self):
        return lib().j40_current_frame(C.byref(self._img))

    def frame_pixels_u8x4(self, frame=None, channel=J40_RGBA):
        """Returns (array[h, w, 4] uint8 copy, stride_bytes)."""
        if frame is None:
            frame = self.current_frame()
        px = lib().j40_frame_pixels_u8x4(C.byref(frame), channel)
        raw = np.ctypeslib.as_array(C.cast(px.data, C.POINTER(C.c_uint8)), shape=(px.height, px.stride_bytes))
        return raw[:, : px.width * 4].reshape(px.height, px.width, 4).copy(), px.stride_bytes

    def error(self):
        return err_str(lib().j40_error(C.byref(self._img)))

    def error_string(self):
        return lib().j40_error_string(C.byref(self._img)).decode("latin1")

    def free(self):
        if self._open:
            lib().j40_free(C.byref(self._img))
            self._open = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def decode(data: bytes):
    """dj40-style decode through the j40_* API. Returns (rgba or None, err4, error string, stride)."""
    im = Image.from_memory(data)
    im.output_format(J40_RGBA, J40_U8X4)
    out, stride = None, 0
    if im.next_frame():
        out, stride = im.frame_pixels_u8x4()
    err, msg = im.error(), ""
    if err:
        msg = im.error_string()
    im.free()
    return out, err, msg, stride


class Batch:
    """Batch decode on one CUDA device (include/j40b.h)."""

    def __init__(self, device=0):
        self._h = lib().j40b_batch_create(device)
        if not self._h:
            raise RuntimeError("j40_b200: no usable CUDA device (there is no CPU decoding path)")
        self._bufs = []

    def add(self, data: bytes):
        buf = C.create_string_buffer(bytes(data), len(data))
        self._bufs.append(buf)
        return lib().j40b_batch_add(self._h, C.cast(buf, C.c_void_p), len(data))

    def add_many(self, datas, threads=0):
        """Parse many images on `threads` host threads (0 = all cores, at most 16). `datas`: bytes objects, kept
        alive by this object until reset()/close()."""
        n = len(datas)
        keep = [d if isinstance(d, bytes) else bytes(d) for d in datas]
        ptrs = (C.c_void_p * n)(*[C.cast(C.c_char_p(d), C.c_void_p) for d in keep])
        sizes = (C.c_size_t * n)(*[len(d) for d in keep])
        self._bufs.extend(keep)
        return lib().j40b_batch_add_many(self._h, ptrs, sizes, n, threads)

    def upload(self):
        if lib().j40b_batch_upload(self._h) != 0:
            raise RuntimeError("j40_b200: upload failed")

    def reset(self):
        """Forget the images, keep the device and pinned allocations (waits for work in flight)."""
        lib().j40b_batch_reset(self._h)
        self._bufs = []

    def read_all_async(self, out):
        """Enqueue the D2H copy of every image into `out` (uint8 array [n, pitch], ideally pinned) behind the
        decode kernels; complete after wait()."""
        assert out.ndim >= 2 and out.dtype == np.uint8
        pitch = out.strides[0]
        if lib().j40b_batch_read_all_async(self._h, out.ctypes.data, pitch) != 0:
            raise RuntimeError("j40_b200: read_all_async before decode")

    def decode(self):
        lib().j40b_batch_decode(self._h)

    def wait(self):
        return lib().j40b_batch_wait(self._h)

    def count(self):
        return lib().j40b_batch_count(self._h)

    def error(self, i):
        return err_str(lib().j40b_batch_error(self._h, i))

    def info(self, i):
        w, h, s = C.c_int32(), C.c_int32(), C.c_int32()
        lib().j40b_batch_info(self._h, i, C.byref(w), C.byref(h), C.byref(s))
        return w.value, h.value, s.value

    def device_pixels(self, i):
        return lib().j40b_batch_device_pixels(self._h, i)

    def read_pixels(self, i, out=None):
        w, h, s = self.info(i)
        raw = np.empty((h, s), np.uint8) if out is None else out
        if lib().j40b_batch_read_pixels(self._h, i, raw.ctypes.data) != 0:
            return None
        return raw[:, : w * 4].reshape(h, w, 4)

    def debug_dump(self, i, lf_group, what, out):
        """diagnostics: intermediate array `what` (include/j40b.h) of an LF group into the numpy array `out`; bytes written"""
        return int(lib().j40b_batch_debug_dump(self._h, i, lf_group, what, out.ctypes.data, out.nbytes))

    def write_pam(self, i, path):
        """image i as a PAM (P7, RGB_ALPHA) file, copied from device memory without the stride padding"""
        return int(lib().j40b_batch_write_pam(self._h, i, os.fsencode(path)))

    def last_decode_ms(self):
        return float(lib().j40b_batch_last_decode_ms(self._h))

    def kernel_ms(self):
        names = ["lf_group", "hf_group", "back", "back_big", "modular", "render", "lf_image", "lf_hfmeta", "lf_llf"]
        return {n: float(lib().j40b_batch_kernel_ms(self._h, i)) for i, n in enumerate(names)}

    def stat(self, what):
        return int(lib().j40b_batch_stat(self._h, what))

    def mark(self, which):
        lib().j40b_batch_mark(self._h, which)

    def join(self, other):
        lib().j40b_batch_join(self._h, other._h)

    def after(self, other, stage=0):
        return int(lib().j40b_batch_after(self._h, other._h, stage))

    def event_ms(self, ref, which):
        return float(lib().j40b_batch_event_ms(self._h, ref._h, which))

    def mark_ms(self):
        return float(lib().j40b_batch_mark_ms(self._h))

    def close(self):
        if self._h:
            lib().j40b_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
