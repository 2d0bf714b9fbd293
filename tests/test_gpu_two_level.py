"""The instancing-aware two-level layout (LmcuBuildOptions::two_level; the analogue of accel::embreeinstanced,
/root/reference/plugin/accel_embree/accel_embree_instanced.cpp:96-165) against the flattened tree and the reference.

The contract: the SAME hits.  The bottom levels are walked in object space, but the leaf test transforms a triangle's vertices
with the operations the flattened build applies once (glm's mat4 * vec4 order) and runs the reference's Möller–Trumbore on
them, so (t, u, v) are bit-identical to the flattened tree's and to accel::sahbvh's; shading reads the instance transform
and the mesh arrays instead of per-triangle world-space records."""
import os

import numpy as np
import pytest

import lmcu
from lmcu import abi, scenes
from parity import assert_hit_parity, assert_pt_parity, rmse

pytestmark = pytest.mark.gpu


def _ray_batch(sb, grid, n=60000, w=320, h=180):
    lo, hi = np.array([-1., -.2, -1.]) * grid, np.array([1., 1.2, 1.]) * grid
    return np.concatenate([scenes.camera_rays(sb.camera, w, h), scenes.random_rays(n, lo, hi)])


def _same_hits(a, b):
    """Two trees over the same triangles: every field bit-equal, except rays whose closest hit is shared by two triangles
    (equal t): which of them is reported depends on the order of the walk."""
    bad = (a["primitive"] != b["primitive"]) | (a["face"] != b["face"])
    assert (a["t"][bad] == b["t"][bad]).all() and (a["primitive"][bad] >= 0).all() and (b["primitive"][bad] >= 0).all()
    ok = ~bad
    for k in ("t", "u", "v"):
        assert (a[k][ok] == b[k][ok]).all(), k
    return int(bad.sum())


def _same_image(fa, fb, max_tie_frac=0.0):
    """Deterministic renders of the same hits: equal to FP32 rounding (G2: RMSE <= 1e-6), except at most `max_tie_frac` of
    the pixels, where the ray meets an edge shared by two primitives and the walk's order picks the one reported."""
    diff = np.abs(fa.astype(np.float64) - fb).max(axis=2)
    off = diff > 1e-4
    assert off.mean() <= max_tie_frac, off.sum()
    assert rmse(fa[~off], fb[~off]) <= 1e-6


@pytest.mark.parametrize("grid,num_theta", [(3, 12), (2, 24), (5, 32)])
def test_two_level_hits_equal_flattened_and_reference(reference, grid, num_theta):
    sb = scenes.instanced_spheres(grid=grid, num_theta=num_theta)
    flat, two = lmcu.Scene(sb, two_level=-1), lmcu.Scene(sb, two_level=1)
    assert two.build_stats.two_level == 1 and flat.build_stats.two_level == 0
    assert two.build_stats.num_triangles == flat.build_stats.num_triangles == sb.num_triangles()
    assert two.build_stats.node_bytes + two.build_stats.triangle_bytes < flat.build_stats.node_bytes + flat.build_stats.triangle_bytes
    rays = _ray_batch(sb, grid)
    hf, ht = flat.intersect(rays), two.intersect(rays, counted=True)
    assert two.nodes_per_ray > 0 and two.tris_per_ray > 0
    ties = _same_hits(hf, ht)
    assert ties <= 4, ties                                 # rotated instances: shared edges are met exactly almost never
    assert (ht["primitive"] >= 0).sum() > 0.2 * len(rays)
    rs = reference.RefScene(sb)
    assert_hit_parity(rs.intersect(rays[:120000]), ht[:120000], max_special=4)
    # the scalar lm::Accel::intersect contract and the per-thread walk behind it
    for i in list(range(0, len(rays), 997))[:48]:
        h1 = two.intersect1(rays[i])
        assert h1["primitive"] == ht[i]["primitive"] and h1["face"] == ht[i]["face"] and h1["t"] == ht[i]["t"] \
            and h1["u"] == ht[i]["u"] and h1["v"] == ht[i]["v"]
    rs.close(); flat.close(); two.close()


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_two_level_under_general_affine_transforms(reference, seed):
    """Shear, anisotropic scale 0.02 .. 30 per axis, mirrored instances (negative determinant): the object-space node tests
    stay conservative — no hit of the flattened tree is lost, every field bit-equal — through the queue kernel, and the hits
    are accel::sahbvh's."""
    sb = scenes.affine_instances(seed)
    flat, two = lmcu.Scene(sb, two_level=-1), lmcu.Scene(sb, two_level=1)
    assert two.build_stats.two_level == 1
    rays = np.concatenate([scenes.camera_rays(sb.camera, 400, 400), scenes.random_rays(200000, np.array([-6, -6, -6.]), np.array([6, 6, 6.]), seed=seed)])
    hf, ht = flat.intersect(rays), two.intersect(rays)
    assert _same_hits(hf, ht) <= 2 and (ht["primitive"] >= 0).sum() > 50000
    rs = reference.RefScene(sb)
    assert_hit_parity(rs.intersect(rays[:100000]), ht[:100000], max_special=4)
    p = abi.render_params(256, 256, visualize_normal=True)
    fa, _ = flat.render("raycast", p)
    fb, _ = two.render("raycast", p)
    _same_image(fa, fb, 1e-4)
    rs.close(); flat.close(); two.close()


@pytest.mark.parametrize("nf_list", [[1], [1, 1, 1], [2, 1], [3, 7, 1]])
def test_two_level_ragged_and_tiny_meshes(nf_list):
    """Meshes of 1, 2, 3 and 7 triangles (mesh trees whose root is a leaf cluster, a top level of 3 .. 9 instances)."""
    from test_hostsim import _tiny_instanced
    sb = _tiny_instanced(nf_list)
    flat, two = lmcu.Scene(sb, two_level=-1), lmcu.Scene(sb, two_level=1)
    assert two.build_stats.two_level == 1
    rays = np.concatenate([scenes.camera_rays(sb.camera, 96, 96), scenes.random_rays(20000, np.array([-2, -2, -2.]), np.array([2, 2, 2.]))])
    hf, ht = flat.intersect(rays), two.intersect(rays)
    assert _same_hits(hf, ht) == 0 and (ht["primitive"] >= 0).sum() > 300
    for i in range(0, len(rays), 2503):
        h1 = two.intersect1(rays[i])
        assert h1["primitive"] == ht[i]["primitive"] and h1["t"] == ht[i]["t"]
    flat.close(); two.close()


def test_two_level_tmin_tmax_window_and_any_hit():
    sb = scenes.instanced_spheres(grid=3, num_theta=16)
    flat, two = lmcu.Scene(sb, two_level=-1), lmcu.Scene(sb, two_level=1)
    rays = _ray_batch(sb, 3, n=40000, w=160, h=90)
    rays["tmin"] = 0.7
    rays["tmax"] = 2.9
    _same_hits(flat.intersect(rays), two.intersect(rays))
    # ambient occlusion = per-thread any-hit walks with a short tmax: same image
    p = abi.render_params(160, 90, spp=8, ao_distance=0.6)
    fa, _ = flat.render("ao", p)
    fb, _ = two.render("ao", p)
    assert np.abs(fa - fb).max() <= 1.0 / 8 + 1e-6 and (fa != fb).mean() < 1e-3
    flat.close(); two.close()


@pytest.mark.parametrize("visualize_normal", [False, True])
def test_two_level_raycast_equals_flattened(visualize_normal):
    """Positions, shading normals and materials resolved at the hit from (instance transform, mesh arrays) equal the
    flattened build's per-triangle records (G2's tolerance: RMSE <= 1e-6)."""
    sb = scenes.instanced_spheres(grid=3, num_theta=24)
    flat, two = lmcu.Scene(sb, two_level=-1), lmcu.Scene(sb, two_level=1)
    p = abi.render_params(640, 360, visualize_normal=visualize_normal)
    fa, _ = flat.render("raycast", p)
    fb, _ = two.render("raycast", p)
    _same_image(fa, fb)
    assert fa.std() > 0.05
    flat.close(); two.close()


def test_two_level_textured_meshes_and_area_lights(tmp_path):
    """Texture coordinates read from the mesh arrays; emissive primitives; meshes WITHOUT normals (geometric normal)."""
    sb = scenes.textured_scene(tmp_path, 12, "obj")
    flat, two = lmcu.Scene(sb, two_level=-1), lmcu.Scene(sb, two_level=1)
    assert two.build_stats.two_level == 1
    p = abi.render_params(256, 144)
    fa, _ = flat.render("raycast", p)
    fb, _ = two.render("raycast", p)
    _same_image(fa, fb, 1e-3)              # the shell's walls share edges that the symmetric camera meets exactly
    assert fa.std() > 0.02
    p = abi.render_params(128, 72, spp=64, max_verts=8)
    fa, _ = flat.render("pt", p)
    fb, _ = two.render("pt", p)
    assert rmse(fa, fb) <= 1e-2 * fa.mean(), (rmse(fa, fb), fa.mean())
    flat.close(); two.close()
    # cornell_box(True): boxes through non-trivial transforms, quads without vertex normals, an area light
    sb = scenes.cornell_box(True)
    flat, two = lmcu.Scene(sb, two_level=-1), lmcu.Scene(sb, two_level=1)
    rays = np.concatenate([scenes.camera_rays(sb.camera, 320, 180), scenes.random_rays(30000, np.array([-1, 0, -1.]), np.array([1, 2, 1.]))])
    _same_hits(flat.intersect(rays), two.intersect(rays))
    for mode in (abi.SAMPLING_MIS, abi.SAMPLING_NEE):
        p = abi.render_params(128, 72, spp=64, max_verts=6, sampling_mode=mode)
        fa, _ = flat.render("pt", p)
        fb, _ = two.render("pt", p)
        assert rmse(fa, fb) <= 1e-2 * fa.mean(), (mode, rmse(fa, fb), fa.mean())
    flat.close(); two.close()


@pytest.mark.parametrize("renderer", ["pt", "lt", "bdpt"])
def test_two_level_renderers_equal_flattened(renderer, tmp_path):
    """Same Philox streams, same hits => the same image up to the order of the film's fp32 atomics (and the rare tie)."""
    env = None if renderer == "pt" else scenes.env_light(os.path.join(str(tmp_path), "e.hdr"), 64, 32)
    sb = scenes.instanced_spheres(grid=2, num_theta=24, env=env)
    flat, two = lmcu.Scene(sb, two_level=-1), lmcu.Scene(sb, two_level=1)
    if renderer == "pt":
        p = abi.render_params(192, 108, spp=32, max_verts=10)
    else:
        p = abi.render_params(96, 54, max_verts=6, primary_ray_sampling_mode=abi.PRIMARY_IMAGE, num_samples=96 * 54 * 32)
    fa, sa = flat.render(renderer, p)
    fb, sb_ = two.render(renderer, p)
    assert sa["rays"] == sb_["rays"] or abs(sa["rays"] - sb_["rays"]) <= 1e-4 * sa["rays"]
    assert np.isfinite(fb).all() and fa.mean() > 0.05
    assert rmse(fa, fb) <= 5e-3 * fa.mean(), (rmse(fa, fb), fa.mean())
    flat.close(); two.close()


def test_two_level_pt_vs_reference(reference):
    """G3 against renderer::pt itself (C4's materials under a constant environment)."""
    sb = scenes.instanced_spheres(grid=2, num_theta=24)
    two, rs = lmcu.Scene(sb, two_level=1), reference.RefScene(sb)
    p = abi.render_params(160, 90, spp=128, max_verts=10)
    fb, _ = two.render("pt", p)
    fr, _, _ = rs.render("pt", p)
    assert_pt_parity(fr, fb, 128)
    rs.close(); two.close()


def test_two_level_policy_and_serialisation(monkeypatch):
    sb = scenes.instanced_spheres(grid=2, num_theta=12)
    auto = lmcu.Scene(sb)                                   # 0 = automatic: flattened while it fits
    assert auto.build_stats.two_level == 0
    auto.save_accel()
    auto.close()
    monkeypatch.setenv("LMCU_BUILD_BUDGET_GB", "0.001")     # ... two-level when the flattened build of an instanced scene would not
    auto = lmcu.Scene(sb)
    assert auto.build_stats.two_level == 1
    auto.close()
    plain = lmcu.Scene(scenes.cornell_spheres(16))          # nothing instanced: flattened whatever the budget
    assert plain.build_stats.two_level == 0
    plain.close()
    monkeypatch.delenv("LMCU_BUILD_BUDGET_GB")
    two = lmcu.Scene(sb, two_level=1)
    with pytest.raises(lmcu.LmcuError):                     # rebuilt from the scene, not serialised
        two.save_accel()
    two.close()
    # a singular instance transform cannot be walked in object space: the build falls back to the flattened tree
    sb = scenes.instanced_spheres(grid=2, num_theta=12)
    tri = scenes._f32(np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float64))
    sb.add_primitive(sb.add_mesh(tri, scenes.flat_normals(tri)), sb.diffuse((.5, .5, .5)), M=scenes.scale([1.0, 1.0, 0.0]))
    flat = lmcu.Scene(sb, two_level=1)
    assert flat.build_stats.two_level == 0
    flat.close()


def test_two_level_c4_full_size():
    """C4 (BASELINE configs[3]): 1000 instances of a 49 728-triangle sphere.  Two-level: megabytes instead of 4.75 GB, built
    from the unique triangles; hits equal the flattened tree's on a camera batch."""
    sb = scenes.instanced_spheres(grid=10, num_theta=112)
    two = lmcu.Scene(sb, two_level=1)
    st = two.build_stats
    assert st.two_level == 1 and st.num_triangles == 49_728_002
    assert st.node_bytes + st.triangle_bytes < 100 << 20, (st.node_bytes, st.triangle_bytes)
    rays = scenes.camera_rays(sb.camera, 960, 540)
    ht = two.intersect(rays, counted=True)
    npr, tpr = two.nodes_per_ray, two.tris_per_ray
    film, stats = two.render("pt", abi.render_params(480, 270, spp=16, max_verts=20))
    assert np.isfinite(film).all() and 0.2 < film.mean() < 1.0 + 1e-3       # unit-radiance sky: energy-bounded
    two.close()
    flat = lmcu.Scene(sb)                                   # automatic: the flattened tree (23.6 GB at the build's peak) fits
    assert flat.build_stats.two_level == 0
    hf = flat.intersect(rays, counted=True)
    ties = _same_hits(hf, ht)
    print("C4 two-level: %.2f MB, build %.1f ms, %.2f nodes + %.2f tris per ray (flattened: %.0f MB, %.1f ms, %.2f + %.2f), ties %d"
          % ((st.node_bytes + st.triangle_bytes) / 1e6, st.elapsed * 1e3, npr, tpr,
             (flat.build_stats.node_bytes + flat.build_stats.triangle_bytes) / 1e6, flat.build_stats.elapsed * 1e3,
             flat.nodes_per_ray, flat.tris_per_ray, ties))
    assert ties <= 8
    flat.close()
