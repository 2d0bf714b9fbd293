"""Edge-case scenes and ray batches for the closest-hit contract (shared by the CPU hostsim test and the -m gpu test)."""
import numpy as np

from lmcu import scenes


def _scene(tris, M=None):
    sb = scenes.SceneBuilder()
    tris = np.asarray(tris, dtype=np.float64).reshape(-1, 3, 3)
    sb.add_primitive(sb.add_mesh(tris, None), sb.diffuse((.5, .5, .5)), M=M)
    sb.set_camera(position=(0, 0, 4), center=(0, 0, 0), up=(0, 1, 0), vfov=40, aspect=1)
    return sb


def _rays(o, d, tmin=1e-3, tmax=1e10):
    o, d = np.asarray(o, dtype=np.float64).reshape(-1, 3), np.asarray(d, dtype=np.float64).reshape(-1, 3)
    r = np.zeros(len(o), dtype=scenes.RAY_DTYPE)
    r["o"], r["d"], r["tmin"], r["tmax"] = o, d, tmin, tmax
    return r


def grid_mesh(n=12, z=0.0):
    """n x n quads in [-1,1]^2 at height z, float-representable vertices"""
    xs = scenes._f32(np.linspace(-1, 1, n + 1))
    tris = []
    for i in range(n):
        for j in range(n):
            a, b, c, d = (xs[i], xs[j], z), (xs[i + 1], xs[j], z), (xs[i + 1], xs[j + 1], z), (xs[i], xs[j + 1], z)
            tris += [(a, b, c), (a, c, d)]
    return np.asarray(tris, dtype=np.float64)


def cases():
    """name -> (SceneBuilder, rays, max fraction of tie / boundary rays)"""
    rng = np.random.RandomState(5)
    out = {}

    # 1. axis-parallel rays (direction components exactly 0) onto a grid: many go through vertices and along edges
    g = grid_mesh(12)
    xs = scenes._f32(np.linspace(-1.2, 1.2, 49))
    ox, oy = np.meshgrid(xs, xs)
    o = np.stack([ox.ravel(), oy.ravel(), np.full(ox.size, 2.0)], -1)
    axis = _rays(o, np.tile([0.0, 0.0, -1.0], (len(o), 1)))
    side = _rays(np.stack([np.full(49, -3.0), xs, np.zeros(49)], -1), np.tile([1.0, 0.0, 0.0], (49, 1)))   # IN the mesh plane
    diag = _rays(o, np.tile([0.0, 1e-3, -1.0], (len(o), 1)))
    out["axis_parallel"] = (_scene(g), np.concatenate([axis, side, diag]), 281)

    # 2. degenerate triangles (zero area: repeated vertex, collinear) and exact duplicates mixed into a soup
    soup = scenes._f32(rng.rand(60, 3, 3) * 2 - 1)
    deg = soup[:20].copy()
    deg[:10, 2] = deg[:10, 1]                                    # repeated vertex
    deg[10:, 2] = deg[10:, 0] + 2 * (deg[10:, 1] - deg[10:, 0])  # collinear
    dup = soup[20:30].copy()                                     # coincident copies: exact ties
    tris = np.concatenate([soup, scenes._f32(deg), dup])
    out["degenerate_and_duplicate"] = (_scene(tris), scenes.random_rays(6000, np.array([-1.5, -1.5, -1.5]), np.array([1.5, 1.5, 1.5]), seed=9), 99)

    # 3. a fan of triangles sharing one centroid region (duplicate Morton codes) + a needle and a huge triangle
    ang = np.linspace(0, 2 * np.pi, 33)
    fan = [((0, 0, 0), (np.cos(a), np.sin(a), 0.1 * k), (np.cos(b), np.sin(b), 0.1 * k)) for k in range(4) for a, b in zip(ang[:-1], ang[1:])]
    extra = [((-50, -50, -2), (50, -50, -2), (0, 80, -2)), ((0.3, 0.3, -1), (0.3000001, 0.3, 1), (0.3, 0.3000001, 1))]
    out["fan_needle_huge"] = (_scene(scenes._f32(np.asarray(fan + extra, dtype=np.float64))),
                              scenes.random_rays(6000, np.array([-1.5, -1.5, -3.0]), np.array([1.5, 1.5, 1.5]), seed=10), 0)

    # 4. far from the origin and non-float-representable (FP64 triangle records): stresses the padded FP32 node test
    M = scenes.translate((10000.25, -5000.5, 2000.125)) @ scenes.rotate(33.0, (1, 2, 3)) @ scenes.scale((1.7, 1.7, 1.7))
    soup4 = scenes._f32(rng.rand(300, 3, 3) * 2 - 1)
    r4 = scenes.random_rays(6000, np.array([-3, -3, -3.]), np.array([3, 3, 3.]), seed=11)
    r4["o"] += np.array([10000.25, -5000.5, 2000.125])
    out["far_offset_rotated"] = (_scene(soup4, M), r4, 0)

    # 5. tiny scale
    soup5 = scenes._f32((rng.rand(300, 3, 3) * 2 - 1) * 1e-3)
    r5 = scenes.random_rays(4000, np.array([-2e-3] * 3), np.array([2e-3] * 3), seed=12)
    r5["tmin"] = 1e-9
    out["tiny_scale"] = (_scene(soup5), r5, 0)

    # 6. ray parameter corner cases on a plain soup: empty interval, zero-length window, tmin = 0 on the surface,
    #    unnormalised directions (t scales with 1/|d|), a hit exactly AT tmax (inclusive) and just before it
    soup6 = scenes._f32(rng.rand(200, 3, 3) * 2 - 1)
    base = scenes.random_rays(3000, np.array([-1.5] * 3), np.array([1.5] * 3), seed=13)
    empty = base.copy(); empty["tmin"] = 2.0; empty["tmax"] = 1.0
    zero = base.copy(); zero["tmax"] = 0.0
    longd = base.copy(); longd["d"] *= 1e3
    shortd = base.copy(); shortd["d"] *= 1e-3
    out["ray_parameters"] = (_scene(soup6), np.concatenate([base, empty, zero, longd, shortd]), 0)
    return out
