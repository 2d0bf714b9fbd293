"""oracle/shim/cereal is a functional stand-in: the reference's own archives (include/lm/serial.h) round-trip its asset
tree through lm::save_state_to_file / lm::reset / lm::load_state_from_file exactly as functest/func_serial_consistency.py
does — raycast before and after, RMSE 0.  (The -m gpu twin with accel::cuda_bvh / film::cuda / renderer::*_cuda in the
tree is tests/test_gpu_plugin.py::test_save_state_round_trip_with_cuda_components.)"""
import numpy as np

from lmcu import abi, scenes


def test_save_state_round_trip_reference_components(reference, tmp_path):
    sb = scenes.instanced_spheres(grid=2, num_theta=10)        # instanced transforms, four material kinds, an env light
    rs = reference.RefScene(sb)
    rays = scenes.camera_rays(sb.camera, 64, 36)
    p = abi.render_params(96, 54)
    before_hits = rs.intersect(rays)
    pt_before, _, _ = rs.render("pt", abi.render_params(32, 18, spp=256, max_verts=6, seed=3))
    # leaves a film::bitmap and a renderer::raycast asset in the tree, like the functest (a renderer::pt asset cannot be
    # saved by the reference itself: its scheduler has a locator that renderer::pt's underlying() does not resolve)
    before_img, _, _ = rs.render("raycast", p)
    path = tmp_path / "lm.serialized"
    rs.save_reset_load(path)
    assert path.stat().st_size > 10000
    after_hits = rs.intersect(rays)                            # accel::sahbvh restored from nodes_ / trs_ / indices_, not rebuilt
    after_img, _, _ = rs.render("raycast", p)
    pt_after, _, _ = rs.render("pt", abi.render_params(32, 18, spp=256, max_verts=6, seed=3))
    for k in ("t", "u", "v", "face"):
        assert np.array_equal(before_hits[k], after_hits[k])
    assert np.array_equal(before_hits["primitive"] >= 0, after_hits["primitive"] >= 0)
    assert before_img.max() > 0 and np.array_equal(before_img, after_img)      # RMSE 0
    # materials, lights, camera, Dist tables restored (renderer::pt keeps its thread-local RNGs across renders, so two renders
    # of one process never repeat bit for bit: compare the image means)
    assert abs(pt_before.mean() - pt_after.mean()) < 0.03 * pt_before.mean()
    rs.close()
