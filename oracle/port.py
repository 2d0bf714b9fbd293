"""ctypes driver for oracle/liblmoracle.so (the C restatement, oracle/lm_oracle.c).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import _HERE
from lmcu.scenes import HIT_DTYPE, RAY_DTYPE

_LIB = None


def build():
    so = os.path.join(_HERE, "liblmoracle.so")
    srcs = [os.path.join(_HERE, "lm_oracle.c"), os.path.join(_HERE, "lm_oracle.h"),
            os.path.join(os.path.dirname(_HERE), "include", "lmcu.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared",
                               "-o", so, srcs[0], "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_scene_create.restype = C.c_void_p
    return _LIB


def rng_uniforms(seed, n):
    out = np.zeros(n)
    lib().orc_rng_fill(C.c_uint32(seed), out.ctypes.data_as(C.c_void_p), n)
    return out


class PortScene:
    def __init__(self, builder):
        self._b = builder
        self._desc = builder.desc()
        self._h = C.c_void_p(lib().orc_scene_create(C.byref(self._desc)))
        if not self._h:
            raise RuntimeError("oracle port: scene creation failed")

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        lib().orc_intersect(self._h, rays.ctypes.data_as(C.c_void_p), C.c_int64(len(rays)), hits.ctypes.data_as(C.c_void_p))
        return hits

    def render(self, renderer, params, num_threads=0):
        film = np.zeros((params.height, params.width, 3), dtype=np.float64)
        fp = film.ctypes.data_as(C.c_void_p)
        if renderer == "raycast":
            lib().orc_render_raycast(self._h, C.byref(params), fp)
        elif renderer == "ao":
            lib().orc_render_ao(self._h, C.byref(params), fp, num_threads)
        else:
            lib().orc_render_pt(self._h, C.byref(params), fp, num_threads)
        return film

    def close(self):
        if self._h:
            lib().orc_scene_destroy(self._h)
            self._h = C.c_void_p()
