"""ctypes driver for oracle/_ref/liblmref.so (the unmodified reference behind a C ABI).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import ctypes as C
import os

import numpy as np

from . import _HERE  # noqa: F401  (also puts the product package on sys.path for lmcu.abi)
from lmcu import abi
from lmcu.scenes import HIT_DTYPE, RAY_DTYPE

_LIB = None


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "liblmref.so"))


def lib():
    global _LIB
    if _LIB is None:
        # liblm.so must be global so that plugins dlopen'ed later resolve lm::comp::detail::reg
        C.CDLL(os.path.join(_HERE, "_ref", "liblm.so"), mode=C.RTLD_GLOBAL)
        _LIB = C.CDLL(os.path.join(_HERE, "_ref", "liblmref.so"), mode=C.RTLD_GLOBAL)
        _LIB.lmref_last_error.restype = C.c_char_p
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference: " + lib().lmref_last_error().decode())


def init(num_threads=-0, verbose=False):
    """num_threads <= 0 means hardware_concurrency() + num_threads (parallel_openmp.cpp:25-29)."""
    _check(lib().lmref_init(int(num_threads), int(verbose)))
    return lib().lmref_num_threads()


_PLUGINS = set()


def load_plugin(path):
    """lm::comp::load_plugin, once per path and process (test modules share one reference instance)."""
    if path in _PLUGINS:
        return
    _check(lib().lmref_load_plugin(path.encode()))
    _PLUGINS.add(path)


class RefScene:
    def __init__(self, builder, accel="accel::sahbvh"):
        self._builder = builder
        self._desc = builder.desc()
        self._h = C.c_void_p()
        secs = C.c_double()
        _check(lib().lmref_scene_create(C.byref(self._desc), accel.encode(), C.byref(self._h), C.byref(secs)))
        self.build_seconds = secs.value

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        secs = C.c_double()
        _check(lib().lmref_accel_intersect(self._h, rays.ctypes.data_as(C.c_void_p), C.c_int64(len(rays)),
                                           hits.ctypes.data_as(C.c_void_p), C.byref(secs)))
        self.last_seconds = secs.value
        return hits

    def render(self, renderer, params):
        """renderer: 'raycast' | 'ao' | 'pt' (or a full key).  Returns (film[h,w,3] f64, elapsed, processed)."""
        key = renderer if "::" in renderer else "renderer::" + renderer
        film = np.zeros((params.height, params.width, 3), dtype=np.float64)
        elapsed, processed = C.c_double(), C.c_longlong()
        _check(lib().lmref_render(self._h, key.encode(), C.byref(params), film.ctypes.data_as(C.c_void_p),
                                  C.byref(elapsed), C.byref(processed)))
        return film, elapsed.value, processed.value

    def save_reset_load(self, path):
        """lm::save_state_to_file(path); lm::reset(); lm::load_state_from_file(path) — func_serial_consistency.py:79-83."""
        _check(lib().lmref_save_reset_load(self._h, str(path).encode()))

    def loc(self):
        """Locator of the lm::Scene asset inside the reference's asset tree."""
        lib().lmref_scene_loc.restype = C.c_char_p
        return lib().lmref_scene_loc(self._h).decode()

    def close(self):
        if self._h:
            lib().lmref_scene_destroy(self._h)
            self._h = C.c_void_p()
