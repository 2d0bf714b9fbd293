"""ctypes driver for tests/hostsim (debugging aid: the CUDA kernel bodies compiled by g++).
Not a product path — see hostsim.cpp."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, os.path.join(_ROOT, "lightmetrica-v3_b200"))
from lmcu.scenes import HIT_DTYPE, RAY_DTYPE  # noqa: E402

_LIB = None


def build():
    out = os.path.join(_HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhostsim.so")
    srcs = [os.path.join(_HERE, "hostsim.cpp")] + [
        os.path.join(_ROOT, "lightmetrica-v3_b200", "csrc", f)
        for f in ("common.cuh", "bvh.cuh", "build.cuh", "shade.cuh", "kernels.cuh", "bdpt.cuh", "prep.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared",
                               "-o", so, srcs[0]] + os.environ.get("HOSTSIM_CXXFLAGS", "").split())
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.hostsim_last_error.restype = C.c_char_p
    return _LIB


class SimScene:
    def __init__(self, builder, max_leaf=3, two_level=-1):
        self._b = builder
        self._desc = builder.desc()
        self._h = C.c_void_p()
        nn = C.c_longlong()
        rc = lib().hostsim_scene_create2(C.byref(self._desc), max_leaf, two_level, C.byref(self._h), C.byref(nn))
        if rc:
            raise RuntimeError(lib().hostsim_last_error().decode())
        self.num_nodes = nn.value
        self.fp32 = bool(lib().hostsim_fp32(self._h))
        self.two_level = bool(lib().hostsim_two_level(self._h))

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        npr, tpr = C.c_double(), C.c_double()
        self.overflow = lib().hostsim_intersect(self._h, rays.ctypes.data_as(C.c_void_p), C.c_longlong(len(rays)),
                                                hits.ctypes.data_as(C.c_void_p), C.byref(npr), C.byref(tpr))
        self.nodes_per_ray, self.tris_per_ray = npr.value, tpr.value
        return hits

    def render(self, renderer, params):
        film = np.zeros((params.height, params.width, 3), dtype=np.float32)
        fp = film.ctypes.data_as(C.c_void_p)
        if renderer == "raycast":
            lib().hostsim_raycast(self._h, C.byref(params), fp)
        elif renderer == "ao":
            lib().hostsim_ao(self._h, C.byref(params), fp)
        elif renderer == "bdpt":
            rays = C.c_longlong()
            if lib().hostsim_bdpt(self._h, C.byref(params), fp, C.byref(rays)):
                raise RuntimeError(lib().hostsim_last_error().decode())
            self.rays = rays.value
        elif renderer == "lt":
            rays = C.c_longlong()
            lib().hostsim_lt(self._h, C.byref(params), fp, C.byref(rays))
            self.rays = rays.value
        else:
            rays = C.c_longlong()
            lib().hostsim_pt(self._h, C.byref(params), fp, C.byref(rays))
            self.rays = rays.value
        return film
