"""Drop-in test through the reference's OWN Python API: oracle/_ref/pylm.so is the reference's pybind11 module
(src/pylm.cpp, unmodified), so these tests read like its functests (functest/func_accel_consistency.py,
functest/func_render_instancing.py, example_quad.py): assets created with lm.load_mesh / lm.load_material / ..., the
plugin loaded with lm.comp.load_plugin, and the new components selected by key exactly where a user types 'sahbvh' / 'pt'."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PLUGIN = os.path.join(ROOT, "lightmetrica-v3_b200", "plugin", "lm_cuda")


@pytest.fixture(scope="module")
def lm():
    if not os.path.exists(os.path.join(REF, "pylm.so")) or not os.path.exists(PLUGIN + ".so"):
        pytest.skip("oracle/_ref/pylm.so or the plugin is not built")
    ctypes.CDLL(os.path.join(REF, "liblm.so"), mode=ctypes.RTLD_GLOBAL)
    sys.path.insert(0, REF)
    import pylm
    pylm.init()
    pylm.comp.load_plugin(PLUGIN)
    return pylm


def sphere_mesh(lm, name, n_theta=24, n_phi=48, r=0.4):
    """The numpy sphere of functest/func_render_instancing.py:60-92, as a mesh::raw asset."""
    from lmcu import scenes
    p, n = scenes.uv_sphere(n_theta, n_phi, r)                       # [faces, 3, 3] positions / normals
    ps = p.reshape(-1, 3)
    ns = n.reshape(-1, 3)
    idx = np.arange(len(ps), dtype=np.int32)
    return lm.load_mesh(name, "raw", ps=ps.flatten().tolist(), ns=ns.flatten().tolist(), ts=[0, 0],
                        fs={"p": idx.tolist(), "n": idx.tolist(), "t": np.zeros(len(ps), dtype=np.int32).tolist()})


def build_scene(lm, accel_key):
    lm.reset()
    accel = lm.load_accel("accel", accel_key)
    scene = lm.load_scene("scene", "default", accel=accel.loc())
    camera = lm.load_camera("camera", "pinhole", position=[0, 1, 5], center=[0, 1, 0], up=[0, 1, 0], vfov=30, aspect=16 / 9)
    scene.add_primitive(camera=camera.loc())
    quad = lm.load_mesh("quad", "raw", ps=[-1, 0, -1, 1, 0, -1, 1, 0, 1, -1, 0, 1], ns=[0, 1, 0], ts=[0, 0, 1, 0, 1, 1, 0, 1],
                        fs={"p": [0, 1, 2, 0, 2, 3], "n": [0, 0, 0, 0, 0, 0], "t": [0, 1, 2, 0, 2, 3]})
    white = lm.load_material("white", "diffuse", Kd=[.7, .7, .7])
    scene.add_primitive(mesh=quad.loc(), material=white.loc())
    sphere = sphere_mesh(lm, "sphere")
    mats = [lm.load_material("red", "diffuse", Kd=[.8, .2, .2]), lm.load_material("glass", "glass", Ni=1.5),
            lm.load_material("gold", "glossy", Ks=[.8, .6, .3], ax=.2, ay=.2)]
    for i, x in enumerate((-.9, 0, .9)):
        t = np.eye(4)
        t[:3, 3] = [x, .4 + .1 * i, 0]
        scene.add_transformed_primitive(t, mesh=sphere.loc(), material=mats[i].loc())
    light_mesh = lm.load_mesh("lquad", "raw", ps=[-.5, 2.5, -.5, .5, 2.5, -.5, .5, 2.5, .5, -.5, 2.5, .5], ns=[0, -1, 0], ts=[0, 0],
                              fs={"p": [0, 2, 1, 0, 3, 2], "n": [0] * 6, "t": [0] * 6})
    black = lm.load_material("black", "diffuse", Kd=[0, 0, 0])
    light = lm.load_light("light", "area", Ke=[12, 12, 12], mesh=light_mesh.loc())
    scene.add_primitive(mesh=light_mesh.loc(), material=black.loc(), light=light.loc())
    return scene


def render(lm, renderer_key, film_key="bitmap", w=160, h=90, **kw):
    film = lm.load_film("film", film_key, w=w, h=h)
    renderer = lm.load_renderer("renderer", renderer_key, scene="$.assets.scene", output=film.loc(), **kw)
    out = renderer.render()
    return np.array(film.buffer(), copy=True), out


def test_plugin_registers_its_components_in_the_reference_registry(lm):
    keys = []
    lm.comp.foreach_registered(lambda k: keys.append(k))
    for k in ("accel::cuda_bvh", "renderer::raycast_cuda", "renderer::ao_cuda", "renderer::pt_cuda", "renderer::lt_cuda", "renderer::bdpt_cuda",
              "renderer::bdptopt_cuda", "renderer::pt_naive_path_cuda", "renderer::pt_nee_path_cuda", "renderer::lt_nee_path_cuda", "film::cuda"):
        assert k in keys
    # every renderer the reference registers on this path has a device counterpart under '<key>_cuda'
    for k in keys:
        if k.startswith("renderer::") and not k.endswith("_cuda") and k not in ("renderer::blank", "renderer::volpt", "renderer::volpt_naive"):
            assert k + "_cuda" in keys, k
    # the reference's own pipeline through its Python API (CPU): accel::sahbvh + renderer::raycast
    scene = build_scene(lm, "sahbvh")
    scene.build()
    img, _ = render(lm, "raycast")
    assert img.shape == (90, 160, 3) and img.max() > 0.1


def test_no_gpu_means_a_loud_error_not_a_fallback(lm):
    import lmcu
    if lmcu.device_count() > 0:
        pytest.skip("a GPU is present")
    scene = build_scene(lm, "cuda_bvh")
    with pytest.raises(lm.Exception) as e:                              # lm::Exception as the Python API surfaces it
        scene.build()
    assert "no CUDA device" in str(e.value)


@pytest.mark.gpu
def test_functest_style_consistency_through_the_python_api(lm):
    """func_accel_consistency: the same scene rendered with accel::sahbvh and accel::cuda_bvh under the reference's
    renderer::raycast — RMSE ~ 0; then renderer::pt against renderer::pt_cuda into film::cuda."""
    scene = build_scene(lm, "sahbvh")
    scene.build()
    ref_img, _ = render(lm, "raycast")
    ref_pt, _ = render(lm, "pt", scheduler="sample", spp=128, max_verts=10, seed=1)
    scene = build_scene(lm, "cuda_bvh")
    scene.build()
    img, _ = render(lm, "raycast")                                     # reference renderer, CUDA accel (scalar contract)
    d = np.abs(img - ref_img).max(axis=2)
    assert (d > 1e-12).sum() <= 8                                      # ties on shared edges only
    cu_img, out = render(lm, "raycast_cuda", film_key="cuda")
    assert "elapsed" in out and np.sqrt(np.mean((cu_img - ref_img)[d <= 1e-12] ** 2)) <= 1e-6
    pt, out = render(lm, "pt_cuda", film_key="cuda", scheduler="sample", spp=128, max_verts=10, seed=1)
    assert out["processed"] == 128
    from parity import assert_pt_parity
    assert_pt_parity(ref_pt, pt, 128, rel_mean=0.02)
