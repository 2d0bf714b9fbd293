"""Kernel bodies compiled by g++ (tests/hostsim, a debugging aid — NOT a product path) against the
unmodified reference: catches logic errors in the builder / traversal / shading source before it
reaches a GPU.  The parity tests proper are the -m gpu ones, which run the CUDA build."""
import numpy as np
import pytest

from lmcu import abi, scenes
from parity import assert_hit_parity, assert_pt_parity, classify_hits, rmse


@pytest.fixture(scope="module")
def sim():
    import sim as simmod
    simmod.build()
    return simmod


SPECIAL = {"cornell": 26, "cornell_fp64": 29, "spheres": 15, "heightfield": 0}    # measured: symmetric camera, 160x90 + 8000 random rays


@pytest.mark.parametrize("name,make", [
    ("cornell", lambda: scenes.cornell_box()),
    ("cornell_fp64", lambda: scenes.cornell_box(True)),
    ("spheres", lambda: scenes.cornell_spheres(12)),
    ("heightfield", lambda: scenes.heightfield_scene(48)),
])
def test_closest_hit_matches_sahbvh(reference, sim, name, make):
    sb = make()
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    lo, hi = (np.array([-1, 0, -1.]), np.array([1, 2, 1.])) if "height" not in name else (np.array([-1, -.2, -1.]), np.array([1, .6, 1.]))
    rays = np.concatenate([scenes.camera_rays(sb.camera, 160, 90), scenes.random_rays(8000, lo, hi)])
    hr, hs = rs.intersect(rays), ss.intersect(rays)
    assert ss.overflow == 0
    c = assert_hit_parity(hr, hs, max_special=SPECIAL[name])
    assert c["ties"] + c["ref_boundary_miss"] == SPECIAL[name]      # deterministic batch: the measured count, exactly
    rs.close()


def test_ragged_and_tiny_scenes(reference, sim):
    """1, 2, 3, 4 and 7 triangles (leaf-only roots, a root with one internal child)."""
    for nf in (1, 2, 3, 4, 7):
        sb = scenes.SceneBuilder()
        rng = np.random.RandomState(nf)
        tri = scenes._f32(rng.rand(nf, 3, 3) * 2 - 1)
        sb.add_primitive(sb.add_mesh(tri, scenes.flat_normals(tri)), sb.diffuse((.5, .5, .5)))
        sb.set_camera(position=(0, 0, 4), center=(0, 0, 0), up=(0, 1, 0), vfov=40, aspect=1)
        rs, ss = reference.RefScene(sb), sim.SimScene(sb)
        rays = scenes.camera_rays(sb.camera, 64, 64)
        assert_hit_parity(rs.intersect(rays), ss.intersect(rays), max_special=0)
        rs.close()


def test_tmin_tmax_window(reference, sim):
    sb = scenes.cornell_spheres(8)
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    rays = scenes.random_rays(5000, np.array([-1, 0, -1.]), np.array([1, 2, 1.]), seed=3)
    rays["tmin"] = 0.3
    rays["tmax"] = 0.9
    hr, hs = rs.intersect(rays), ss.intersect(rays)
    assert_hit_parity(hr, hs, max_special=0)
    assert ((hs["t"][hs["primitive"] >= 0] >= 0.3) & (hs["t"][hs["primitive"] >= 0] <= 0.9)).all()
    rs.close()


def test_raycast_image(reference, sim):
    sb = scenes.cornell_spheres(12)
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    rays = scenes.camera_rays(sb.camera, 192, 108)
    bad = classify_hits(rs.intersect(rays), ss.intersect(rays))["bad_mask"].reshape(108, 192)
    for kw in ({}, {"visualize_normal": True}, {"use_constant_color": True}, {"color": (.2, .4, .6)}):
        p = abi.render_params(192, 108, **kw)
        fr, _, _ = rs.render("raycast", p)
        fs = ss.render("raycast", p)
        assert rmse(fs[~bad], fr[~bad]) <= 1e-6
    rs.close()


@pytest.mark.parametrize("mode", [abi.SAMPLING_MIS, abi.SAMPLING_NEE, abi.SAMPLING_NAIVE])
def test_pt_statistics(reference, sim, mode):
    sb = scenes.cornell_spheres(8)
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    p = abi.render_params(96, 54, spp=96, max_verts=12, sampling_mode=mode)
    fr, _, _ = rs.render("pt", p)
    fs = ss.render("pt", p)
    assert_pt_parity(fr, fs, 96, rel_mean=0.02)
    rs.close()


def test_env_light_image(reference, sim, tmp_path):
    """light::env (lat-long .hdr, Dist2 importance sampling, rot + scale): raycast shows the map itself, pt in all
    three sampling modes agrees with the reference."""
    env = scenes.env_light(str(tmp_path / "env.hdr"), w=128, h=64, rot=30.0, scale=1.5)
    sb = scenes.instanced_spheres(grid=2, num_theta=12, env=env)
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    p = abi.render_params(160, 90)
    fr, _, _ = rs.render("raycast", p)
    fs = ss.render("raycast", p)
    d = np.abs(fs - fr).max(axis=2)
    assert (d > 1e-4).mean() < 3e-3 and rmse(fs[d <= 1e-4], fr[d <= 1e-4]) <= 1e-6
    for mode in (abi.SAMPLING_MIS, abi.SAMPLING_NEE, abi.SAMPLING_NAIVE):
        p = abi.render_params(64, 36, spp=64, max_verts=8, sampling_mode=mode)
        fr, _, _ = rs.render("pt", p)
        assert_pt_parity(fr, ss.render("pt", p), 64, rel_mean=0.02)
    rs.close()


@pytest.mark.parametrize("kind", ["obj", "mix"])
def test_textured_and_mixture_materials(reference, sim, tmp_path, kind):
    """mapKd textures, material::mixture_wavefrontobj (+ alpha mask), material::mask and the constant-weight mixtures:
    the shading source against the reference before it reaches a GPU (tests/test_gpu_parity.py has the CUDA run)."""
    sb = scenes.textured_scene(tmp_path, 12, kind)
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    p = abi.render_params(192, 108)
    fr, _, _ = rs.render("raycast", p)
    fs = ss.render("raycast", p)
    d = np.abs(fs - fr).max(axis=2)
    assert (d > 1e-4).mean() < 3e-3 and rmse(fs[d <= 1e-4], fr[d <= 1e-4]) <= 1e-6
    for mode in (abi.SAMPLING_MIS, abi.SAMPLING_NAIVE):
        p = abi.render_params(96, 54, spp=96, max_verts=10, sampling_mode=mode)
        fr, _, _ = rs.render("pt", p)
        assert_pt_parity(fr, ss.render("pt", p), 96, rel_mean=0.02)
    rs.close()


@pytest.mark.parametrize("kind", ["obj", "mix"])
def test_textured_and_mixture_materials_vs_golden(sim, tmp_path, kind):
    """Same source against the committed reference outputs (no reference needed): raycast to G2 tolerances, pt means."""
    import os
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs_materials.npz"))
    sb = scenes.textured_scene(tmp_path, 12, kind)
    ss = sim.SimScene(sb)
    fs = ss.render("raycast", abi.render_params(192, 108))
    fr = golden[kind + "_raycast"]
    d = np.abs(fs - fr).max(axis=2)
    assert (d > 1e-4).mean() < 3e-3 and rmse(fs[d <= 1e-4], fr[d <= 1e-4]) <= 1e-6
    fs = ss.render("pt", abi.render_params(96, 54, spp=64, max_verts=12))
    gm, sm = golden[kind + "_pt_mis"].mean(axis=(0, 1)), fs.mean(axis=(0, 1))
    assert (np.abs(sm - gm) <= 0.03 * gm).all()           # 64 spp against the 1024-spp golden image: means only


def test_point_and_directional_lights(reference, sim):
    sb = scenes.delta_lights_scene(2, 12)
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    for mode in (abi.SAMPLING_MIS, abi.SAMPLING_NEE):
        p = abi.render_params(64, 36, spp=128, max_verts=8, sampling_mode=mode)
        fr, _, _ = rs.render("pt", p)
        assert_pt_parity(fr, ss.render("pt", p), 128, rel_mean=0.015)
    p = abi.render_params(64, 36, spp=4, max_verts=8, sampling_mode=abi.SAMPLING_NAIVE)
    assert ss.render("pt", p).max() == 0.0          # delta lights are invisible to BSDF sampling
    rs.close()


@pytest.mark.parametrize("scene,mode", [("spheres", abi.SAMPLING_MIS), ("spheres", abi.SAMPLING_NEE), ("spheres", abi.SAMPLING_NAIVE),
                                        ("delta", abi.SAMPLING_MIS), ("env", abi.SAMPLING_MIS)])
def test_pt_image_mode_statistics(reference, sim, tmp_path, scene, mode):
    """primary_ray_sampling_mode=image: raster positions drawn over the whole film, NEE from the camera vertex splatted
    at the light's own raster position, film scaled by w*h/num_samples (renderer_pt.cpp:104-106,126-157,310-312);
    equal sample counts, statistical gate G3."""
    if scene == "spheres":
        sb = scenes.cornell_spheres(8)
    elif scene == "delta":
        sb = scenes.delta_lights_scene(2, 12)
    else:
        sb = scenes.instanced_spheres(grid=2, num_theta=12, env=scenes.env_light(str(tmp_path / "env.hdr"), w=64, h=32, rot=30.0, scale=1.5))
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    n = 96 * 54 * 96
    p = abi.render_params(96, 54, max_verts=12, sampling_mode=mode, primary_ray_sampling_mode=abi.PRIMARY_IMAGE, num_samples=n)
    fr, _, processed = rs.render("pt", p)
    assert processed == n
    assert_pt_parity(fr, ss.render("pt", p), 96, rel_mean=0.02)
    rs.close()


def test_edge_cases_closest_hit(reference, sim):
    """Axis-parallel and in-plane rays, degenerate and duplicate triangles, duplicate Morton codes, needle / huge
    triangles, geometry far from the origin under a rotated instance transform, tiny scale, empty / zero / unnormalised
    ray parameters, hits exactly at tmax: (primitive, face, t, u, v) against accel::sahbvh (edge_cases.py)."""
    import edge_cases
    for name, (sb, rays, special) in edge_cases.cases().items():
        rs, ss = reference.RefScene(sb), sim.SimScene(sb)
        hr, hs = rs.intersect(rays), ss.intersect(rays)
        assert ss.overflow == 0, name
        assert_hit_parity(hr, hs, max_special=special)
        # idempotence: a hit found at t is found again with tmax = t (the upper bound is inclusive).  Checked against our
        # own first answer, not the reference's: for an axis-aligned triangle the reference's FP64 slab test and its
        # triangle test round the same t differently, so accel::sahbvh itself loses such a hit when tmax = t
        ok = hs["primitive"] >= 0
        if ok.any():
            at = rays[ok].copy(); at["tmax"] = hs["t"][ok]
            again = ss.intersect(at)
            assert all(np.array_equal(again[k], hs[k][ok]) for k in ("t", "u", "v", "primitive", "face")), name
        rs.close()


def test_scene_description_validation(sim):
    """prepare_scene (csrc/prep.hpp) — the validation lmcu_scene_upload runs before anything reaches the device — refuses
    inconsistent descriptions with the reference's error classes (missing primitives = Unsupported, include/lm/scene.h:232-238;
    bad indices = InvalidArgument).  Runs on the CPU through the hostsim build of the same header."""
    def fails(sb, needle):
        with pytest.raises(RuntimeError) as e:
            sim.SimScene(sb)
        assert needle in str(e.value), str(e.value)

    fails(scenes.SceneBuilder(), "Missing primitives")
    sb = scenes.cornell_box(); sb.primitives[0]["material"] = 99
    fails(sb, "bad material index")
    sb = scenes.cornell_box(); sb.primitives[0]["mesh"] = 99
    fails(sb, "bad mesh index")
    sb = scenes.cornell_box(); sb.primitives[0]["material"] = sb.diffuse((1, 1, 1), map_kd=5)
    fails(sb, "bad mapKd texture index")
    sb = scenes.cornell_box(); sb.primitives[0]["material"] = sb.mixture_rr([(sb.glass(1.5), 1.0)])
    fails(sb, "diffuse, glossy, mirror or mask")
    sb = scenes.cornell_box(); sb.primitives[0]["material"] = sb.mixture_marginalized([(0, -1.0)])
    fails(sb, "weight must be positive")
    sb = scenes.cornell_box(); sb.set_envconst((1, 1, 1)); sb.set_envconst((1, 1, 1))
    fails(sb, "Environment light is already registered")
    sb = scenes.cornell_box(); sb.add_directional_light((1, 1, 1), (0, 0, 0))
    fails(sb, "zero direction")
    sb = scenes.cornell_box(); sb.add_texture(np.zeros((2, 2, 3), np.float32)); sb.textures[0]["data"] = np.zeros((2, 2, 3), np.float32)[:, :, :2]
    with pytest.raises((RuntimeError, AssertionError, ValueError)):
        sim.SimScene(sb)


@pytest.mark.parametrize("scene", ["spheres", "env"])
def test_lt_statistics(reference, sim, scene):
    """The light tracer's device code (lt_generate_body / lt_shade_body: Light::sample_ray of every light type, LE transport
    through the BSDFs, camera connections splatted at their raster position) against the reference's renderer::lt."""
    sb = scenes.cornell_spheres(12) if scene == "spheres" else scenes.instanced_spheres(grid=2, num_theta=10)
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    w, h, spp = 64, 36, 48
    p = abi.render_params(w, h, max_verts=10, num_samples=w * h * spp, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    fr, _, processed = rs.render("renderer::lt", p)
    assert processed == w * h * spp
    fs = ss.render("lt", p)
    assert_pt_parity(fr, fs, spp, rel_mean=0.02)
    rs.close()


@pytest.mark.parametrize("scene", ["spheres", "env"])
def test_bdpt_statistics(reference, sim, scene):
    """The bidirectional path tracer's device code (bdpt.cuh: both subpaths, every connection, power-heuristic weights from
    the per-vertex density factors) against the reference's renderer::bdpt at equal sample counts."""
    sb = scenes.cornell_spheres(12) if scene == "spheres" else scenes.instanced_spheres(grid=2, num_theta=10)
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    w, h, spp = 64, 36, 32
    p = abi.render_params(w, h, max_verts=6, num_samples=w * h * spp, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    fr, _, processed = rs.render("renderer::bdpt", p)
    assert processed == w * h * spp
    fs = ss.render("bdpt", p)
    assert ss.rays > 2 * w * h * spp
    assert_pt_parity(fr, fs, spp, rel_mean=0.02)
    # min_verts / max_verts select path lengths: lengths 2..3 and 4..6 add up to 2..6 (same random numbers)
    lo = ss.render("bdpt", abi.render_params(w, h, max_verts=3, num_samples=w * h * 4, primary_ray_sampling_mode=abi.PRIMARY_IMAGE))
    al = ss.render("bdpt", abi.render_params(w, h, max_verts=3, min_verts=3, num_samples=w * h * 4, primary_ray_sampling_mode=abi.PRIMARY_IMAGE))
    assert 0 < al.mean() < lo.mean()
    rs.close()


@pytest.mark.parametrize("scene", ["spheres", "env", "delta", "obj", "mix"])
@pytest.mark.parametrize("name,mode", [("pt_naive_path", abi.BDPT_PT_NAIVE), ("pt_nee_path", abi.BDPT_PT_NEE),
                                       ("lt_nee_path", abi.BDPT_LT_NEE)])
def test_single_strategy_path_renderers(reference, sim, tmp_path, scene, name, mode):
    """LmcuRenderParams::bdpt_strategies: the bdpt device code restricted to one unweighted strategy family against the
    reference's renderer::pt_naive_path / pt_nee_path / lt_nee_path (src/renderer/renderer_bdpt.cpp:63-288).  The glass
    sphere exercises pt_nee_path's switch to (0, n) behind a specular vertex; also: a constant environment (a light at
    infinity connects through no (1, n-1)), point + directional lights (degenerated endpoints: pt_naive_path stays black
    in the reference too), the OBJ-import material set and the constant-weight mixtures."""
    sb = {"spheres": lambda: scenes.cornell_spheres(12), "env": lambda: scenes.instanced_spheres(grid=2, num_theta=10),
          "delta": lambda: scenes.delta_lights_scene(2, 16), "obj": lambda: scenes.textured_scene(tmp_path, 12, "obj"),
          "mix": lambda: scenes.textured_scene(tmp_path, 12, "mix")}[scene]()
    rs, ss = reference.RefScene(sb), sim.SimScene(sb)
    w, h, spp = 64, 36, 48
    kw = dict(max_verts=6, num_samples=w * h * spp, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    fr, _, processed = rs.render("renderer::" + name, abi.render_params(w, h, **kw))
    assert processed == w * h * spp
    fs = ss.render("bdpt", abi.render_params(w, h, bdpt_strategies=mode, **kw))
    assert fs.mean() > 0 or (scene == "delta" and mode == abi.BDPT_PT_NAIVE and fr.max() == 0)
    assert_pt_parity(fr, fs, spp, rel_mean=0.02)
    if scene == "spheres" and mode != abi.BDPT_LT_NEE:
        # complete estimators of the image renderer::bdpt converges to (lt_nee_path has no strategy for a path whose
        # vertex before the pinhole is specular, and with a light at infinity (1, n-1) never connects: as in the reference)
        fb = ss.render("bdpt", abi.render_params(w, h, **kw))
        assert_pt_parity(fb, fs, spp, rel_mean=0.02)
    rs.close()


@pytest.mark.parametrize("grid,num_theta", [(3, 12), (2, 24)])
def test_two_level_layout_equals_flattened(sim, grid, num_theta):
    """LmcuBuildOptions::two_level (one tree per mesh in object space + one over the instances; leaf test on vertices
    transformed with the flattened build's operations): every hit field, the raycast image and a path-traced image are
    bit-identical to the flattened tree's (sequential CPU build of the kernel bodies: same sample streams, same order)."""
    sb = scenes.instanced_spheres(grid=grid, num_theta=num_theta)
    a, b = sim.SimScene(sb, two_level=-1), sim.SimScene(sb, two_level=1)
    assert b.two_level and not a.two_level and b.num_nodes < a.num_nodes
    lo, hi = np.array([-1., -.2, -1.]) * grid, np.array([1., 1.2, 1.]) * grid
    rays = np.concatenate([scenes.camera_rays(sb.camera, 160, 90), scenes.random_rays(20000, lo, hi)])
    ha, hb = a.intersect(rays), b.intersect(rays)
    assert a.overflow == 0 and b.overflow == 0 and (ha["primitive"] >= 0).sum() > 5000
    for k in ("primitive", "face", "t", "u", "v"):
        assert (ha[k] == hb[k]).all(), k
    rays["tmin"], rays["tmax"] = 0.7, 2.9
    ha, hb = a.intersect(rays), b.intersect(rays)
    for k in ("primitive", "face", "t", "u", "v"):
        assert (ha[k] == hb[k]).all(), k
    p = abi.render_params(64, 36)
    assert np.array_equal(a.render("raycast", p), b.render("raycast", p))
    p = abi.render_params(32, 18, spp=4, max_verts=6)
    fa, fb = a.render("pt", p), b.render("pt", p)
    assert np.array_equal(fa, fb) and fa.mean() > 0.1
    p = abi.render_params(32, 18, spp=2, ao_distance=0.5)
    assert np.array_equal(a.render("ao", p), b.render("ao", p))
    # the light tracer, the bidirectional path tracer and its single-strategy forms read their vertices through the same
    # surface-point code: identical images on both layouts
    for renderer, mode in (("lt", abi.BDPT_ALL), ("bdpt", abi.BDPT_ALL), ("bdpt", abi.BDPT_PT_NAIVE), ("bdpt", abi.BDPT_PT_NEE),
                           ("bdpt", abi.BDPT_LT_NEE)):
        p = abi.render_params(32, 18, max_verts=6, num_samples=32 * 18 * 4, bdpt_strategies=mode, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
        fa, fb = a.render(renderer, p), b.render(renderer, p)
        assert np.array_equal(fa, fb) and fa.mean() > 0.01, (renderer, mode)


def test_two_level_vs_reference_and_policy(reference, sim, tmp_path):
    sb = scenes.instanced_spheres(grid=2, num_theta=12)
    rs, ss = reference.RefScene(sb), sim.SimScene(sb, two_level=1)
    rays = np.concatenate([scenes.camera_rays(sb.camera, 160, 90), scenes.random_rays(8000, np.array([-2, 0, -2.]), np.array([2, 3, 2.]))])
    assert_hit_parity(rs.intersect(rays), ss.intersect(rays), max_special=0)
    rs.close()
    assert not sim.SimScene(sb, two_level=0).two_level                      # automatic: a small scene stays flattened
    # textured meshes (texture coordinates from the mesh arrays) and meshes without vertex normals
    sb = scenes.textured_scene(tmp_path, 8, "obj")
    a, b = sim.SimScene(sb, two_level=-1), sim.SimScene(sb, two_level=1)
    assert b.two_level
    # the shell's walls share edges that the symmetric camera meets exactly: on those rays (equal t on two primitives) the
    # reported triangle depends on the order of the walk — 2 of the 2304 pixels here
    p = abi.render_params(64, 36)
    fa, fb = a.render("raycast", p), b.render("raycast", p)
    assert (fa != fb).any(axis=2).sum() <= 2 and fa.std() > 0.02
    p = abi.render_params(24, 14, spp=4, max_verts=5)
    fa, fb = a.render("pt", p), b.render("pt", p)
    assert (fa != fb).any(axis=2).sum() <= 4 and fa.mean() > 0.05
    # a singular transform cannot be walked in object space: flattened
    tri = scenes._f32(np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float64))
    sb.add_primitive(sb.add_mesh(tri, scenes.flat_normals(tri)), sb.diffuse((.5, .5, .5)), M=scenes.scale([1.0, 1.0, 0.0]))
    assert not sim.SimScene(sb, two_level=1).two_level


@pytest.mark.parametrize("seed", [0, 1, 4])
def test_two_level_under_general_affine_transforms(reference, sim, seed):
    """Shear, anisotropic scale 0.02 .. 30, mirrored instances: the object-space node tests stay conservative (no hit of the
    flattened tree is lost, every field bit-equal) and the hits are the reference's."""
    sb = scenes.affine_instances(seed)
    a, b = sim.SimScene(sb, two_level=-1), sim.SimScene(sb, two_level=1)
    assert b.two_level
    rays = np.concatenate([scenes.camera_rays(sb.camera, 120, 120), scenes.random_rays(30000, np.array([-6, -6, -6.]), np.array([6, 6, 6.]), seed=seed)])
    ha, hb = a.intersect(rays), b.intersect(rays)
    assert a.overflow == 0 and b.overflow == 0 and (ha["primitive"] >= 0).sum() > 10000
    for k in ("primitive", "face", "t", "u", "v"):
        assert (ha[k] == hb[k]).all(), k
    rs = reference.RefScene(sb)
    assert_hit_parity(rs.intersect(rays[:20000]), hb[:20000], max_special=2)
    rs.close()


def _tiny_instanced(nf_list, seed=0):
    rng = np.random.RandomState(seed)
    sb = scenes.SceneBuilder()
    for nf in nf_list:
        tri = scenes._f32(rng.rand(nf, 3, 3) * 2 - 1)
        m = sb.add_mesh(tri, scenes.flat_normals(tri))
        for _ in range(3):
            sb.add_primitive(m, sb.diffuse((.5, .5, .5)), M=scenes.translate(rng.uniform(-1, 1, 3)) @ scenes.rotate(rng.rand() * 360, rng.randn(3)))
    sb.set_camera(position=(0, 0, 6), center=(0, 0, 0), up=(0, 1, 0), vfov=40, aspect=1)
    return sb


@pytest.mark.parametrize("nf_list", [[1], [1, 1, 1], [2, 1], [3, 7, 1]])
def test_two_level_ragged_and_tiny_meshes(sim, nf_list):
    """Meshes of 1, 2, 3 and 7 triangles (mesh trees whose root is a leaf cluster, a top level of 3 .. 9 instances)."""
    sb = _tiny_instanced(nf_list)
    a, b = sim.SimScene(sb, two_level=-1), sim.SimScene(sb, two_level=1)
    assert b.two_level
    rays = np.concatenate([scenes.camera_rays(sb.camera, 96, 96), scenes.random_rays(20000, np.array([-2, -2, -2.]), np.array([2, 2, 2.]))])
    ha, hb = a.intersect(rays), b.intersect(rays)
    assert (ha["primitive"] >= 0).sum() > 300
    for k in ("primitive", "face", "t", "u", "v"):
        assert (ha[k] == hb[k]).all(), k
