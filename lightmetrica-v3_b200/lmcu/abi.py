"""ctypes mirror of include/lmcu.h (POD structs only; no logic).

Shared by the product binding (lmcu.api -> liblmcu.so) and by the test-only checkers under
oracle/ (reference harness, C port), which consume the same scene description.
"""
import ctypes as C

c_double3 = C.c_double * 3


class LmcuRay(C.Structure):
    _fields_ = [("o", c_double3), ("d", c_double3), ("tmin", C.c_double), ("tmax", C.c_double)]


class LmcuHit(C.Structure):
    _fields_ = [("t", C.c_double), ("u", C.c_double), ("v", C.c_double),
                ("primitive", C.c_int32), ("face", C.c_int32)]


class LmcuMesh(C.Structure):
    _fields_ = [("num_faces", C.c_int32), ("p", C.POINTER(C.c_double)), ("n", C.POINTER(C.c_double)),
                ("t", C.POINTER(C.c_double))]


class LmcuPrimitive(C.Structure):
    _fields_ = [("node_index", C.c_int32), ("mesh", C.c_int32), ("material", C.c_int32),
                ("light", C.c_int32), ("M", C.c_double * 16)]


MAT_DIFFUSE, MAT_GLOSSY, MAT_GLASS, MAT_MIRROR = 0, 1, 2, 3
MAT_MASK, MAT_MIXTURE_OBJ, MAT_MIXTURE_RR, MAT_MIXTURE_MARGINALIZED = 4, 5, 6, 7
LIGHT_AREA, LIGHT_ENVCONST, LIGHT_ENV, LIGHT_POINT, LIGHT_DIRECTIONAL = 0, 1, 2, 3, 4
SAMPLING_NAIVE, SAMPLING_NEE, SAMPLING_MIS = 0, 1, 2
PRIMARY_PIXEL, PRIMARY_IMAGE = 0, 1
SCHED_SAMPLE, SCHED_TIME = 0, 1
BDPT_ALL, BDPT_PT_NAIVE, BDPT_PT_NEE, BDPT_LT_NEE = 0, 1, 2, 3      # LmcuRenderParams::bdpt_strategies


class LmcuTexture(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("c", C.c_int32), ("pad_", C.c_int32),
                ("data", C.POINTER(C.c_float)), ("path", C.c_char_p)]


class LmcuMaterial(C.Structure):
    _fields_ = [("type", C.c_int32), ("map_kd", C.c_int32), ("color", c_double3),
                ("ax", C.c_double), ("ay", C.c_double), ("Ni", C.c_double), ("Ks", c_double3),
                ("alpha_mask", C.c_int32), ("num_entries", C.c_int32),
                ("entry_material", C.POINTER(C.c_int32)), ("entry_weight", C.POINTER(C.c_double))]


class LmcuLight(C.Structure):
    _fields_ = [("type", C.c_int32), ("primitive", C.c_int32), ("node_index", C.c_int32),
                ("env_w", C.c_int32), ("env_h", C.c_int32), ("pad_", C.c_int32),
                ("Ke", c_double3), ("env_rot", C.c_double), ("env_scale", C.c_double),
                ("env_rgb", C.POINTER(C.c_float)), ("env_path", C.c_char_p), ("position", c_double3)]


class LmcuCamera(C.Structure):
    _fields_ = [("position", c_double3), ("center", c_double3), ("up", c_double3),
                ("vfov", C.c_double), ("aspect", C.c_double)]


class LmcuSceneDesc(C.Structure):
    _fields_ = [("num_meshes", C.c_int32), ("num_primitives", C.c_int32),
                ("num_materials", C.c_int32), ("num_lights", C.c_int32),
                ("num_textures", C.c_int32), ("pad_", C.c_int32),
                ("meshes", C.POINTER(LmcuMesh)), ("primitives", C.POINTER(LmcuPrimitive)),
                ("materials", C.POINTER(LmcuMaterial)), ("lights", C.POINTER(LmcuLight)),
                ("camera", C.POINTER(LmcuCamera)), ("textures", C.POINTER(LmcuTexture))]


class LmcuRenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_verts", C.c_int32),
                ("sampling_mode", C.c_int32), ("spp", C.c_int64), ("sample_offset", C.c_int64),
                ("seed", C.c_uint32), ("rescale", C.c_int32), ("use_constant_color", C.c_int32),
                ("visualize_normal", C.c_int32), ("has_color", C.c_int32), ("count_traversal", C.c_int32),
                ("bg_color", c_double3), ("color", c_double3), ("ao_distance", C.c_double),
                ("primary_ray_sampling_mode", C.c_int32), ("scheduler", C.c_int32), ("num_samples", C.c_int64),
                ("samples_per_iter", C.c_int64), ("render_time", C.c_double),
                ("wavefronts", C.c_int32), ("ao_mark_misses", C.c_int32),
                ("progress", C.c_void_p), ("progress_user", C.c_void_p),
                ("min_verts", C.c_int32), ("bdpt_strategies", C.c_int32)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64, C.c_double)


class LmcuRenderStats(C.Structure):
    _fields_ = [("elapsed", C.c_double), ("processed", C.c_int64), ("paths", C.c_int64),
                ("rays", C.c_int64), ("kernel_launches", C.c_int64),
                ("bvh_nodes_per_ray", C.c_double), ("bvh_tris_per_ray", C.c_double),
                ("shadow_nodes_per_ray", C.c_double), ("shadow_tris_per_ray", C.c_double),
                ("extend_seconds", C.c_double), ("shade_seconds", C.c_double), ("shadow_seconds", C.c_double),
                ("extend_launches", C.c_int64), ("extend_rays", C.c_int64), ("shadow_rays", C.c_int64)]


class LmcuBuildOptions(C.Structure):
    _fields_ = [("max_leaf_tris", C.c_int32), ("two_level", C.c_int32), ("reserved_", C.c_int32 * 6)]


class LmcuBuildStats(C.Structure):
    _fields_ = [("elapsed", C.c_double), ("num_triangles", C.c_int64), ("num_wide_nodes", C.c_int64),
                ("node_bytes", C.c_int64), ("triangle_bytes", C.c_int64),
                ("triangles_fp32", C.c_int32), ("two_level", C.c_int32), ("sah_cost", C.c_double)]


def render_params(width, height, spp=1, max_verts=20, seed=42, sampling_mode=SAMPLING_MIS,
                  sample_offset=0, rescale=1, bg_color=(0, 0, 0), use_constant_color=False,
                  visualize_normal=False, color=None, ao_distance=0.2, count_traversal=False,
                  primary_ray_sampling_mode=PRIMARY_PIXEL, scheduler=SCHED_SAMPLE, num_samples=0,
                  samples_per_iter=0, render_time=0.0, wavefronts=0, ao_mark_misses=False, progress=None,
                  min_verts=0, bdpt_strategies=0):
    """Defaults follow the reference renderers' JSON defaults (renderer_pt.cpp:51-77,
    renderer_raycast.cpp:48-60, functest/renderer_ao.cpp:36-40)."""
    p = LmcuRenderParams()
    p.width, p.height, p.max_verts, p.sampling_mode = width, height, max_verts, sampling_mode
    p.spp, p.sample_offset, p.seed, p.rescale = spp, sample_offset, seed, rescale
    p.use_constant_color, p.visualize_normal = int(use_constant_color), int(visualize_normal)
    p.has_color = int(color is not None)
    p.bg_color = c_double3(*bg_color)
    p.color = c_double3(*(color if color is not None else (0, 0, 0)))
    p.ao_distance = ao_distance
    p.count_traversal = int(count_traversal)
    p.primary_ray_sampling_mode, p.scheduler = primary_ray_sampling_mode, scheduler
    p.num_samples, p.samples_per_iter, p.render_time = num_samples, samples_per_iter, render_time
    p.wavefronts = wavefronts
    p.ao_mark_misses = int(ao_mark_misses)
    p.min_verts = min_verts
    p.bdpt_strategies = bdpt_strategies
    if progress is not None:                      # a PROGRESS_FN instance; the params object keeps it alive
        p._progress_ref = progress
        p.progress = C.cast(progress, C.c_void_p)
    return p
