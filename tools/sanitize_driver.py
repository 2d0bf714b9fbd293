"""Small driver that launches every kernel family of liblmcu once, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_driver.py
    compute-sanitizer --tool initcheck python tools/sanitize_driver.py
    compute-sanitizer --tool racecheck python tools/sanitize_driver.py        (build sort / collapse, ray pool)
    compute-sanitizer --tool synccheck python tools/sanitize_driver.py

Sizes are tiny (the tools slow kernels 10-100x); what matters is that each code path runs: builder (both
triangle formats), accel batches (closest, counted, scalar), raycast / ao / pt (three sampling modes, image-space
primary rays, sample-count scheduler) / lt / bdpt, every material and light type, textures, the FP64 film fold and
finish, the film combine kernels, accel save / load."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v3_b200"))
import lmcu  # noqa: E402
from lmcu import abi, scenes  # noqa: E402


def run_scene(name, sb, w=48, h=27, spp=2, lt=True, bdpt=True, two_level=-1):
    sc = lmcu.Scene(sb, device=0, two_level=two_level)
    assert sc.build_stats.two_level == (1 if two_level > 0 else 0)
    rays = scenes.camera_rays(sb.camera, w, h)
    hits = sc.intersect(rays)
    sc.intersect(rays[:64], counted=True)
    sc.intersect1(rays[len(rays) // 2])
    if two_level <= 0:                                          # a two-level accel is rebuilt, not serialised
        blob = sc.save_accel()
        sc2 = lmcu.Scene(sb, device=0, accel_blob=blob)
        assert (sc2.intersect(rays)["face"] == hits["face"]).all()
        sc2.close()
    sc.render("raycast", abi.render_params(w, h))
    sc.render("raycast", abi.render_params(w, h, visualize_normal=True))
    sc.render("ao", abi.render_params(w, h, spp=spp))
    sc.render("ao", abi.render_params(w, h, spp=spp, ao_mark_misses=True))
    for mode in (abi.SAMPLING_NAIVE, abi.SAMPLING_NEE, abi.SAMPLING_MIS):
        film, _ = sc.render("pt", abi.render_params(w, h, spp=spp, max_verts=6, sampling_mode=mode))
        assert np.isfinite(film).all()
    sc.render("pt", abi.render_params(w, h, max_verts=6, primary_ray_sampling_mode=abi.PRIMARY_IMAGE,
                                      num_samples=w * h * spp))
    sc.render("pt", abi.render_params(w, h, max_verts=6, scheduler=abi.SCHED_TIME, render_time=0.05))
    if lt:
        film, _ = sc.render("lt", abi.render_params(w, h, max_verts=6, num_samples=w * h * spp,
                                                    primary_ray_sampling_mode=abi.PRIMARY_IMAGE))
        assert np.isfinite(film).all()
    if bdpt:
        for mv, mn in ((6, 0), (3, 3)):
            film, st = sc.render("bdpt", abi.render_params(w, h, max_verts=mv, min_verts=mn, num_samples=w * h * spp,
                                                           primary_ray_sampling_mode=abi.PRIMARY_IMAGE))
            assert np.isfinite(film).all() and st["rays"] > 0
        sc.render("bdpt", abi.render_params(w, h, max_verts=5, scheduler=abi.SCHED_TIME, render_time=0.02, samples_per_iter=1000))
        for name in ("pt_naive_path", "pt_nee_path", "lt_nee_path"):      # single-strategy forms: one-sided walks, per-sample strategy choice
            film, st = sc.render(name, abi.render_params(w, h, max_verts=6, num_samples=w * h * spp, primary_ray_sampling_mode=abi.PRIMARY_IMAGE))
            assert np.isfinite(film).all() and st["rays"] > 0
    # film combine kernels on one device (the peer pointers are this scene's own film)
    sc.render("pt", abi.render_params(w, h, spp=spp, max_verts=4, rescale=0), download=False)
    ptr = sc.film_device()[0]
    sc.film_reduce_peers([], 0.5)
    n = w * h * 3
    from lmcu import distributed as D
    for b, e in D.film_slices(n, 3):
        sc.film_reduce_slice([ptr, ptr], ptr, b, e, 0.5)
    sc.film_download(w, h)
    sc.close()
    print("ok", name, flush=True)


def main():
    tmp = tempfile.mkdtemp()
    # the two-level layout: per-mesh trees + tree over the primitives, object-space walk, surface points through instances
    run_scene("two-level instanced+env", scenes.instanced_spheres(grid=2, num_theta=8, env=scenes.env_light(os.path.join(tmp, "e2.hdr"), 64, 32)),
              two_level=1)
    run_scene("two-level textured obj", scenes.textured_scene(tmp, 8, "obj"), two_level=1)
    run_scene("two-level affine", scenes.affine_instances(1, 12, 6), two_level=1, lt=False, bdpt=False)
    run_scene("two-level cornell_box", scenes.cornell_box(True), two_level=1)
    if len(sys.argv) > 1 and sys.argv[1] == "two_level":
        lmcu.cache_trim(0)
        print("driver done")
        return
    run_scene("cornell_box", scenes.cornell_box(True))
    run_scene("cornell_spheres", scenes.cornell_spheres(8))
    run_scene("instanced+envconst", scenes.instanced_spheres(grid=2, num_theta=8))
    run_scene("instanced+env", scenes.instanced_spheres(grid=2, num_theta=8,
                                                        env=scenes.env_light(os.path.join(tmp, "e.hdr"), 64, 32)))
    run_scene("textured obj", scenes.textured_scene(tmp, 8, "obj"))
    run_scene("textured mix", scenes.textured_scene(tmp, 8, "mix"), lt=False)
    run_scene("delta lights", scenes.delta_lights_scene(2, 8))
    hf = scenes.heightfield_scene(40)
    hf.set_envconst((1, 1, 1))
    run_scene("heightfield", hf)
    lmcu.cache_trim(0)
    print("driver done")


if __name__ == "__main__":
    main()
