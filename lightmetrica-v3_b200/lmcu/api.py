"""Python host side of the CUDA hot path: a ctypes binding of liblmcu.so (include/lmcu.h).

Mirrors the reference's component interface for this path:
  Accel     build(scene) / intersect(ray, tmin, tmax)       /root/reference/include/lm/accel.h:38,70
  Renderer  render() -> {"elapsed", "processed"}            /root/reference/include/lm/renderer.h:25
with batched rays instead of one virtual call per ray.  There is no CPU fallback: importing works
anywhere, but creating a Scene without liblmcu.so or without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .scenes import HIT_DTYPE, RAY_DTYPE

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("LMCU_LIB") or os.path.join(_PKG, "liblmcu.so")   # LMCU_LIB: tuning builds only
_LIB = None

EXPORTS = [
    "lmcu_last_error", "lmcu_device_count", "lmcu_version", "lmcu_scene_create", "lmcu_scene_destroy",
    "lmcu_scene_set_stream", "lmcu_scene_upload", "lmcu_accel_build", "lmcu_accel_intersect", "lmcu_accel_intersect1", "lmcu_accel_intersect_device",
    "lmcu_cache_trim", "lmcu_cache_stats", "lmcu_probe_l2_bandwidth",
    "lmcu_accel_intersect_counted", "lmcu_accel_save_size", "lmcu_accel_save", "lmcu_accel_load", "lmcu_render_raycast", "lmcu_render_ao", "lmcu_render_pt", "lmcu_render_lt", "lmcu_render_bdpt",
    "lmcu_film_device_ptr", "lmcu_film_rescale", "lmcu_film_download",
    "lmcu_film_ipc_export", "lmcu_ipc_open", "lmcu_ipc_close", "lmcu_enable_peer_access", "lmcu_film_reduce_peers", "lmcu_film_reduce_slice",
]


class LmcuError(RuntimeError):
    """Raised for every non-zero status; .code is the LMCU_ERR_* value (lm::Error analogue)."""

    def __init__(self, code, msg):
        super().__init__("lmcu error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LmcuError(-1, "liblmcu.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        _LIB = C.CDLL(LIB_PATH)
        _LIB.lmcu_last_error.restype = C.c_char_p
        _LIB.lmcu_version.restype = C.c_char_p
    return _LIB


def _check(rc):
    if rc != 0:
        raise LmcuError(rc, lib().lmcu_last_error().decode())


def device_count():
    return lib().lmcu_device_count()


def cache_trim(device=0):
    _check(lib().lmcu_cache_trim(int(device)))


def cache_stats(device=0):
    """(bytes, blocks) parked in the library's device-block cache."""
    b, n = C.c_int64(), C.c_int64()
    _check(lib().lmcu_cache_stats(int(device), C.byref(b), C.byref(n)))
    return b.value, n.value


def probe_l2_bandwidth(device=0, nbytes=32 << 20):
    """Measured L2 read bandwidth in GB/s (every CTA streams the same nbytes buffer)."""
    v = C.c_double()
    _check(lib().lmcu_probe_l2_bandwidth(int(device), C.c_int64(nbytes), C.byref(v)))
    return v.value


def ipc_open(device, handle):
    """Maps a film exported by another process (Scene.film_ipc_handle) into this process; returns the device pointer."""
    p = C.c_void_p()
    _check(lib().lmcu_ipc_open(int(device), C.c_char_p(handle), C.byref(p)))
    return p.value


def ipc_close(device, ptr):
    _check(lib().lmcu_ipc_close(int(device), C.c_void_p(ptr)))


def enable_peer_access(device, peer_device):
    _check(lib().lmcu_enable_peer_access(int(device), int(peer_device)))


class Scene:
    """accel::cuda_bvh + renderer::{raycast,ao,pt}_cuda over one uploaded scene on one GPU."""

    def __init__(self, builder=None, device=0, max_leaf_tris=0, accel_blob=None, two_level=0):
        """builder: scene description (None = accel-only scene restored from `accel_blob`).  accel_blob: bytes written by
        save_accel(); when given, the BVH is loaded instead of built (Component::save / load of an Accel).  two_level: 1 = one
        tree per mesh + one over the primitives (instanced scenes), -1 = one tree over the flattened scene, 0 = automatic."""
        self._builder = builder
        self._h = C.c_void_p()
        _check(lib().lmcu_scene_create(int(device), C.byref(self._h)))
        if builder is not None:
            self._desc = builder.desc()
            _check(lib().lmcu_scene_upload(self._h, C.byref(self._desc)))
        elif accel_blob is None:
            raise LmcuError(1, "Scene needs a scene description or an accel blob")
        self.build_stats = abi.LmcuBuildStats()
        if accel_blob is not None:
            _check(lib().lmcu_accel_load(self._h, C.c_char_p(bytes(accel_blob)), C.c_int64(len(accel_blob))))
        else:
            opts = abi.LmcuBuildOptions()
            opts.max_leaf_tris = max_leaf_tris
            opts.two_level = two_level
            _check(lib().lmcu_accel_build(self._h, C.byref(opts), C.byref(self.build_stats)))
        self.last_stats = None

    def save_accel(self):
        """The built BVH as bytes (wide nodes, leaf-ordered triangles, shading records, primitive table, checksum)."""
        n = C.c_int64()
        _check(lib().lmcu_accel_save_size(self._h, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(lib().lmcu_accel_save(self._h, buf, n))
        return buf.raw

    # -- Accel ------------------------------------------------------------------------------------
    def intersect(self, rays, counted=False):
        """rays: structured array RAY_DTYPE (host).  Returns HIT_DTYPE array; primitive == -1 on miss."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        rp, hp = rays.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p)
        if counted:
            npr, tpr = C.c_double(), C.c_double()
            _check(lib().lmcu_accel_intersect_counted(self._h, rp, C.c_int64(len(rays)), hp, C.byref(npr), C.byref(tpr)))
            self.nodes_per_ray, self.tris_per_ray = npr.value, tpr.value
        else:
            _check(lib().lmcu_accel_intersect(self._h, rp, C.c_int64(len(rays)), hp))
        return hits

    def intersect1(self, ray):
        """One ray through the scalar contract (lm::Accel::intersect): RAY_DTYPE record in, HIT_DTYPE record out."""
        rays = np.ascontiguousarray(ray, dtype=RAY_DTYPE).reshape(1)
        hits = np.zeros(1, dtype=HIT_DTYPE)
        _check(lib().lmcu_accel_intersect1(self._h, rays.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p)))
        return hits[0]

    def intersect_device(self, d_rays_ptr, n, d_hits_ptr, stream=0):
        """Device buffers (e.g. torch tensors' data_ptr()); asynchronous on `stream`."""
        _check(lib().lmcu_accel_intersect_device(self._h, C.c_void_p(d_rays_ptr), C.c_int64(n), C.c_void_p(d_hits_ptr),
                                                 C.c_void_p(stream)))

    # -- Renderer ---------------------------------------------------------------------------------
    def render(self, renderer, params, download=True):
        """renderer: 'raycast' | 'ao' | 'pt' | 'lt' | 'bdpt' | 'bdptopt', or one of the reference's single-strategy renderers
        'pt_naive_path' | 'pt_nee_path' | 'lt_nee_path' (= 'bdpt' with params.bdpt_strategies set accordingly, on a copy).
        Returns (film[h,w,3] float32 or None, stats dict)."""
        strategies = {"pt_naive_path": abi.BDPT_PT_NAIVE, "pt_nee_path": abi.BDPT_PT_NEE, "lt_nee_path": abi.BDPT_LT_NEE}
        if renderer in strategies:
            caller_params = params                      # keeps what the copy points at (progress callback) alive
            params = abi.LmcuRenderParams.from_buffer_copy(caller_params)
            params.bdpt_strategies = strategies[renderer]
            renderer = "bdpt"
        fn = {"raycast": lib().lmcu_render_raycast, "ao": lib().lmcu_render_ao, "pt": lib().lmcu_render_pt, "lt": lib().lmcu_render_lt,
              "bdpt": lib().lmcu_render_bdpt, "bdptopt": lib().lmcu_render_bdpt}[renderer]
        film = np.empty((params.height, params.width, 3), dtype=np.float32) if download else None
        st = abi.LmcuRenderStats()
        _check(fn(self._h, C.byref(params), film.ctypes.data_as(C.c_void_p) if download else None, C.byref(st)))
        self.last_stats = st
        return film, {f: getattr(st, f) for f, _ in abi.LmcuRenderStats._fields_}

    def set_stream(self, stream_ptr):
        """Run this scene's kernels on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        _check(lib().lmcu_scene_set_stream(self._h, C.c_void_p(stream_ptr)))

    def film_device(self):
        p, n = C.c_void_p(), C.c_int64()
        _check(lib().lmcu_film_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def film_rescale(self, s):
        _check(lib().lmcu_film_rescale(self._h, C.c_double(s)))

    def film_download(self, width, height):
        film = np.empty((height, width, 3), dtype=np.float32)
        _check(lib().lmcu_film_download(self._h, film.ctypes.data_as(C.c_void_p)))
        return film

    # -- multi-GPU film combine over peer memory (include/lmcu.h) -----------------------------------------
    def film_ipc_handle(self):
        """64-byte CUDA IPC handle of the device film (for a root in another process)."""
        buf = C.create_string_buffer(64)
        _check(lib().lmcu_film_ipc_export(self._h, buf))
        return buf.raw

    def film_reduce_peers(self, peer_ptrs, scale):
        """film = (film + sum of the peer films) * scale in ONE kernel reading the peers over NVLink."""
        arr = (C.c_void_p * max(1, len(peer_ptrs)))(*peer_ptrs)
        _check(lib().lmcu_film_reduce_peers(self._h, arr, len(peer_ptrs), C.c_double(scale)))

    def film_reduce_slice(self, film_ptrs, out_ptr, begin, end, scale):
        """out[begin:end] = scale * sum of films[k][begin:end] on THIS scene's GPU (slice-parallel combine; out and the
        films may be peer-mapped pointers)."""
        arr = (C.c_void_p * len(film_ptrs))(*film_ptrs)
        _check(lib().lmcu_film_reduce_slice(self._h, arr, len(film_ptrs), C.c_void_p(out_ptr), C.c_int64(begin), C.c_int64(end),
                                            C.c_double(scale)))

    def close(self):
        if self._h:
            lib().lmcu_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
