"""JPEG XL test-stream writer (ctypes binding of tools/streamgen/libjxlgen.so) -- TEST INFRASTRUCTURE.

No JPEG XL encoder exists in this environment, so every test/bench bitstream is produced by
this writer and validated by the oracle.  The writer is lossy and only has to be syntactically exact; the
numeric tables it needs (dequantisation weights, natural coefficient orders, the linear maps of the special
8x8 transforms) are taken from the oracle at start-up so that no decoder code is duplicated here.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_TABLES_READY = False

_LOG_ORDER = [(3, 3), (3, 3), (4, 4), (5, 5), (3, 4), (3, 5), (4, 5), (6, 6), (5, 6), (7, 7), (6, 7), (8, 8), (7, 8)]
_SPECIAL = [1, 2, 3, 12, 13, 14, 15, 16, 17]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libjxlgen.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(path)
        L.jxlgen_last_error.restype = C.c_char_p
        L.jxlgen_set_table.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        for fn in (L.jxlgen_vardct, L.jxlgen_modular):
            fn.restype = C.c_int
            fn.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int64)]
        L.jxlgen_synth.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        L.jxlgen_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _ensure_tables():
    """Feeds the writer with tables derived from the oracle (reference decoder)."""
    global _TABLES_READY
    if _TABLES_READY:
        return
    from oracle import ref
    L = lib()
    for idx in range(17):
        dq = np.ascontiguousarray(ref.default_dq_matrix(idx), np.float32)
        L.jxlgen_set_table(0, idx, dq.ctypes.data, dq.size)
    for idx, (lr, lc) in enumerate(_LOG_ORDER):
        o = np.ascontiguousarray(ref.natural_order(lr, lc), np.int32)
        L.jxlgen_set_table(1, idx, o.ctypes.data, o.size)
    for sel in _SPECIAL:
        m = np.zeros((64, 64), np.float64)  # m[:, k] = samples produced by unit coefficient k
        for k in range(64):
            e = np.zeros(64, np.float32)
            e[k] = 1.0
            m[:, k] = ref.inverse_transform(sel, e)
        fwd = np.ascontiguousarray(np.linalg.inv(m), np.float64)
        L.jxlgen_set_table(2, sel, fwd.ctypes.data, fwd.size)
    _TABLES_READY = True


STAT_KEYS = ["bytes", "hf_symbols", "lf_symbols", "nonzeros", "num_varblocks", "sections", "coef_clusters", "tree_nodes"]


def _run(fn, params, rgb):
    L = lib()
    s = ",".join(f"{k}={int(v) if not isinstance(v, float) else v}" for k, v in params.items())
    out = C.c_void_p()
    size = C.c_size_t()
    stats = (C.c_int64 * 40)()
    rgbp = None
    if rgb is not None:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        assert rgb.shape == (params["height"], params["width"], 3)
        rgbp = rgb.ctypes.data
    ok = fn(s.encode(), rgbp, C.byref(out), C.byref(size), stats)
    if not ok:
        raise RuntimeError("jxlgen: " + L.jxlgen_last_error().decode())
    data = C.string_at(out, size.value)
    L.jxlgen_free(out)
    st = dict(zip(STAT_KEYS, list(stats)[:8]))
    st["transform_hist"] = list(stats)[8:35]
    return data, st


def vardct(width, height, seed=0, rgb=None, **kw):
    """Returns (codestream bytes, stats). Keyword options: see jxlgen_vardct in jxlgen.cc."""
    _ensure_tables()
    p = dict(width=width, height=height, seed=seed)
    p.update(kw)
    return _run(lib().jxlgen_vardct, p, rgb)


def modular(width, height, seed=0, rgb=None, **kw):
    p = dict(width=width, height=height, seed=seed)
    p.update(kw)
    return _run(lib().jxlgen_modular, p, rgb)


def synth(width, height, seed=0):
    out = np.zeros((height, width, 3), np.uint8)
    lib().jxlgen_synth(width, height, seed, out.ctypes.data)
    return out
