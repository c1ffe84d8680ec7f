"""ctypes loaders + helpers shared by the tests (test infrastructure only).

oracle()   -> oracle/libthor_oracle.so  (our plain-C restatement; built on demand with `make -C oracle port`)
ref()      -> oracle/_ref/libthorref.so (the compiled, unmodified reference; None if not built/travelled)
ref_enc(h) -> oracle/_ref/libthorref_enc_{lbd,hbd}.so (file-static reference kernels via trampolines)
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_cache = {}


def oracle():
    if "o" not in _cache:
        so = os.path.join(ORACLE_DIR, "libthor_oracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in ("thor_oracle.c", "thor_oracle.h", "thor_oracle_tmpl.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])
        lib = C.CDLL(so)
        for n in ("orc_ssd_lbd", "orc_ssd_hbd", "orc_dist_8x8_lbd", "orc_dist_8x8_hbd"):
            getattr(lib, n).restype = C.c_uint64
        for n in ("orc_dct_matrix", "orc_zigzag"):
            getattr(lib, n).restype = C.c_void_p
        _cache["o"] = lib
    return _cache["o"]


def ref():
    if "r" not in _cache:
        so = os.path.join(ORACLE_DIR, "_ref", "libthorref.so")
        lib = None
        if os.path.exists(so):
            lib = C.CDLL(so, mode=C.RTLD_GLOBAL)
            C.c_int.in_dll(lib, "use_simd").value = 1
            for n in ("ssd_calc_simd_lbd", "ssd_calc_simd_hbd"):
                getattr(lib, n).restype = C.c_uint64
        _cache["r"] = lib
    return _cache["r"]


def ref_enc(hbd):
    k = "re%d" % hbd
    if k not in _cache:
        so = os.path.join(ORACLE_DIR, "_ref", "libthorref_enc_%s.so" % ("hbd" if hbd else "lbd"))
        lib = None
        if os.path.exists(so) and ref() is not None:
            lib = C.CDLL(so)
            sfx = "hbd" if hbd else "lbd"
            getattr(lib, "ref_ssd_calc_" + sfx).restype = C.c_uint64
            getattr(lib, "ref_set_use_simd_" + sfx)(1)
        _cache[k] = lib
    return _cache[k]


def aligned(shape, dtype, align=64, fill=None):
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    raw = np.zeros(n * dtype.itemsize + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    a = raw[off:off + n * dtype.itemsize].view(dtype).reshape(shape)
    if fill is not None:
        a[...] = fill
    return a


def P(a, off=0):
    """pointer to element `off` (flat index) of numpy array a"""
    return C.c_void_p(a.ctypes.data + off * a.itemsize)


def sdt(hbd):
    return np.uint16 if hbd else np.uint8


def sfx(hbd):
    return "hbd" if hbd else "lbd"


def rand_plane(rng, h, w, bitdepth, hbd, smooth=False):
    """random sample plane, 64-byte aligned; smooth=True gives low-gradient content (forces SAD ties etc.)"""
    a = aligned((h, w), sdt(hbd))
    maxv = (1 << bitdepth) - 1
    if smooth:
        y, x = np.mgrid[0:h, 0:w]
        v = (np.sin(x / 7.0) + np.cos(y / 5.0)) * (maxv / 8.0) + maxv / 2.0 + rng.integers(-3, 4, (h, w))
        a[...] = np.clip(v, 0, maxv).astype(a.dtype)
    else:
        a[...] = rng.integers(0, maxv + 1, (h, w)).astype(a.dtype)
    return a


def ref_frame(hbd):
    """oracle/_ref/libthorref_frame_{lbd,hbd}.so: cdef_search + dist_8x8 trampolines (enc/encode_frame.c included in place)"""
    k = "rf%d" % hbd
    if k not in _cache:
        so = os.path.join(ORACLE_DIR, "_ref", "libthorref_frame_%s.so" % ("hbd" if hbd else "lbd"))
        lib = None
        if os.path.exists(so) and ref() is not None:
            lib = C.CDLL(so)
            getattr(lib, "ref_dist_8x8_" + ("hbd" if hbd else "lbd")).restype = C.c_uint64
        _cache[k] = lib
    return _cache[k]
