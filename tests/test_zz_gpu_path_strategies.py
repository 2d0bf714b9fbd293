"""The reference's single-strategy renderers on the path structure — renderer::pt_naive_path / pt_nee_path / lt_nee_path
(src/renderer/renderer_bdpt.cpp:63-288) — on the device: lmcu_render_bdpt with LmcuRenderParams::bdpt_strategies, and the
plugin keys renderer::{pt_naive_path,pt_nee_path,lt_nee_path,bdptopt}_cuda.  The CPU build of the same device functions is
checked against the reference in tests/test_hostsim.py::test_single_strategy_path_renderers."""
import os
import sys

import numpy as np
import pytest

import lmcu
from lmcu import abi, scenes
from parity import assert_pt_parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "lightmetrica-v3_b200", "plugin", "lm_cuda")
MODES = [("pt_naive_path", abi.BDPT_PT_NAIVE), ("pt_nee_path", abi.BDPT_PT_NEE), ("lt_nee_path", abi.BDPT_LT_NEE)]


@pytest.mark.parametrize("scene", ["spheres", "env"])
def test_strategy_kernels_equal_the_one_thread_form(scene):
    """The wavefront form (k_bdpt_gen enqueues only the side some strategy extends, k_bdpt_lists applies the per-sample
    strategy choice) against the one-loop-per-sample form of the same functions on the CPU (tests/hostsim): same
    counter-RNG draws, so the images agree to float rounding."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
    import sim
    sb = scenes.cornell_spheres(12) if scene == "spheres" else scenes.instanced_spheres(grid=2, num_theta=10)
    sc, ss = lmcu.Scene(sb), sim.SimScene(sb)
    w, h = 48, 27
    for _, mode in MODES:
        for mv, mn in ((7, 0), (4, 3)):
            p = abi.render_params(w, h, max_verts=mv, min_verts=mn, num_samples=w * h * 8, bdpt_strategies=mode,
                                  primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
            fg, st = sc.render("bdpt", p)
            fs = ss.render("bdpt", p)
            assert st["processed"] == w * h * 8
            assert abs(st["rays"] - ss.rays) <= 1e-3 * ss.rays + 2, (mode, mv, st["rays"], ss.rays)
            assert np.allclose(fg, fs, rtol=1e-3, atol=1e-4 * float(max(fs.max(), 1e-3))), (mode, mv)
    sc.close()


@pytest.mark.parametrize("name,mode", MODES)
def test_g3_single_strategy_vs_reference(reference, name, mode):
    """G3 gate against the reference's own renderer at equal sample counts; sample-split additivity; bad arguments."""
    sb = scenes.cornell_spheres(24)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    w, h, spp = 128, 72, 64
    n = w * h * spp
    img = dict(max_verts=7, num_samples=n, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    fr, _, processed = rs.render("renderer::" + name, abi.render_params(w, h, **img))
    fg, st = sc.render("bdpt", abi.render_params(w, h, bdpt_strategies=mode, **img))
    assert processed == st["processed"] == n
    assert fg.mean() > 0.005
    assert_pt_parity(fr, fg, spp, rel_mean=0.02)
    parts = [sc.render("bdpt", abi.render_params(w, h, max_verts=7, num_samples=n // 2, sample_offset=k * (n // 2), rescale=0,
                                                 bdpt_strategies=mode, primary_ray_sampling_mode=abi.PRIMARY_IMAGE))[0] for k in range(2)]
    assert np.allclose((parts[0] + parts[1]) * (w * h / n), fg, rtol=1e-4, atol=1e-5)
    with pytest.raises(lmcu.LmcuError, match="bdpt_strategies"):
        sc.render("bdpt", abi.render_params(w, h, bdpt_strategies=4, **img))
    rs.close(); sc.close()


def test_plugin_path_strategy_renderers(reference):
    """The plugin keys, created by the reference from the reference renderers' own JSON keys, against those renderers."""
    if not os.path.exists(PLUGIN + ".so"):
        pytest.skip("plugin not built")
    reference.load_plugin(PLUGIN)
    sb = scenes.cornell_spheres(24)
    rs = reference.RefScene(sb)
    w, h, spp = 96, 54, 64
    p = abi.render_params(w, h, max_verts=6, num_samples=w * h * spp, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    for name in ("pt_naive_path", "pt_nee_path", "lt_nee_path", "bdptopt"):
        fr, _, pr = rs.render("renderer::" + name, p)
        fc, _, pc = rs.render("renderer::" + name + "_cuda", p)
        assert pr == pc == w * h * spp, name
        assert_pt_parity(fr, fc, spp, rel_mean=0.02)
    rs.close()


def test_python_mirror_takes_the_reference_renderer_names():
    """lmcu.Scene.render('pt_naive_path' | 'pt_nee_path' | 'lt_nee_path' | 'bdptopt') = 'bdpt' with bdpt_strategies set on a
    copy of the caller's parameters (same samples, same image)."""
    sc = lmcu.Scene(scenes.cornell_spheres(12))
    w, h = 64, 36
    kw = dict(max_verts=6, num_samples=w * h * 16, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    for name, mode in MODES + [("bdptopt", abi.BDPT_ALL)]:
        p = abi.render_params(w, h, **kw)
        a, _ = sc.render(name, p)
        b, _ = sc.render("bdpt", abi.render_params(w, h, bdpt_strategies=mode, **kw))
        assert p.bdpt_strategies == 0
        assert a.mean() > 0 and np.allclose(a, b, rtol=1e-4, atol=1e-6), name
    sc.close()
