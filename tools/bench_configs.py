"""Per-config measurements of SURVEY.md §8(d): C1, C3, C4 on one GPU next to the reference's CPU path on this box's host
cores (C2 / C5 are bench.py's workload).  Diagnostic companion of bench.py, not the bench: writes ONE JSON document.

  python tools/bench_configs.py [--out profiles/r01_configs.json] [--skip c4]

GPU numbers are device time (CUDA events inside liblmcu / torch events around device-resident batches).  CPU numbers are
the unmodified reference (oracle/_ref) under parallel::openmp with every host core, on the bounded samples stated in each
entry (SURVEY.md §8d "CPU baseline timing": the work is reduced so that it finishes, and the reduction is written down).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v3_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import lmcu  # noqa: E402
from lmcu import abi, scenes  # noqa: E402


def device_batch(sc, rays, stream, reps=3):
    d_rays = torch.from_numpy(rays.view(np.float64).reshape(-1, 8)).cuda()
    d_hits = torch.empty((len(rays), 4), dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sc.intersect_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return len(rays) / best / 1e6, d_hits.cpu().numpy().view(scenes.HIT_DTYPE).reshape(-1)


def hit_agreement(hr, hg):
    bad = (hr["primitive"] != hg["primitive"]) | (hr["face"] != hg["face"])
    ties = bad & (hr["t"] == hg["t"])
    same = ~bad & (hr["primitive"] >= 0)
    return {"rays": int(len(hr)), "mismatch": int(bad.sum()), "ties": int(ties.sum()), "other": int((bad & ~ties).sum()),
            "t_u_v_bit_identical": bool(((hr["t"][same] == hg["t"][same]) & (hr["u"][same] == hg["u"][same]) &
                                         (hr["v"][same] == hg["v"][same])).all())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_configs.json"))
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    skip = set(args.skip.split(",")) if args.skip else set()
    from oracle import ref
    have_ref = ref.available()
    threads = ref.init(0) if have_ref else 0
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    out = {"gpu": torch.cuda.get_device_name(0), "host_cores": threads, "cpu": "unmodified reference (oracle/_ref), parallel::openmp"}

    # ---- C1: renderer::raycast of the procedural Cornell box at 1920x1080 -----------------------------------------------
    if "c1" not in skip:
        sb = scenes.cornell_box()
        sc = lmcu.Scene(sb)
        sc.set_stream(stream.cuda_stream)
        p = abi.render_params(1920, 1080)
        best = 1e9
        for _ in range(5):
            fg, st = sc.render("raycast", p, download=False)
            best = min(best, st["elapsed"])
        fg = sc.film_download(1920, 1080)
        rays = scenes.camera_rays(sb.camera, 1920, 1080)
        mr, hg = device_batch(sc, rays, stream)
        e = {"workload": "renderer::raycast, procedural Cornell box (36 triangles), 1920x1080 pixel-centre rays",
             "gpu_raycast_ms": best * 1e3, "gpu_raycast_mrays_per_s": 1920 * 1080 / best / 1e6, "gpu_accel_batch_mrays_per_s": mr}
        if have_ref:
            rs = ref.RefScene(sb)
            fr, el, _ = rs.render("raycast", p)
            fr, el, _ = rs.render("raycast", p)
            hr = rs.intersect(rays)
            d = np.abs(fg - fr).max(axis=2)
            e.update({"cpu_raycast_ms": el * 1e3, "cpu_raycast_mrays_per_s": 1920 * 1080 / el / 1e6,
                      "cpu_accel_mrays_per_s": len(rays) / rs.last_seconds / 1e6, "cpu_sample": "the full config",
                      "hits_vs_sahbvh": hit_agreement(hr, hg),
                      "raycast_rmse_outside_tie_pixels": float(np.sqrt(np.mean((fg[d <= 1e-5] - fr[d <= 1e-5]) ** 2))),
                      "tie_pixel_fraction": float((d > 1e-5).mean())})
            rs.close()
        sc.close()
        out["C1"] = e
        print("C1", json.dumps(e), flush=True)

    # ---- C3: accel batches on the 10 M-triangle heightfield ---------------------------------------------------------------
    if "c3" not in skip:
        t0 = time.time()
        sb = scenes.heightfield_scene(2236)
        sc = lmcu.Scene(sb)
        sc.set_stream(stream.cuda_stream)
        rays = scenes.camera_rays(sb.camera, 4096, 4096)
        coh, hits = device_batch(sc, rays, stream)
        p = sb.meshes[0][0]
        gn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
        gn /= np.linalg.norm(gn, axis=1, keepdims=True)
        bounce = scenes.bounce_rays(rays, hits, lambda h: gn[h["face"]].copy())
        inc, hits_b = device_batch(sc, bounce, stream)
        sc.intersect(rays[:1 << 20], counted=True)
        npr_c, tpr_c = sc.nodes_per_ray, sc.tris_per_ray
        sc.intersect(bounce[:1 << 20], counted=True)
        npr_i, tpr_i = sc.nodes_per_ray, sc.tris_per_ray
        e = {"workload": "closest-hit batches, N=2236 heightfield", "triangles": int(sc.build_stats.num_triangles),
             "gpu_build_ms": sc.build_stats.elapsed * 1e3, "bvh_bytes": int(sc.build_stats.node_bytes + sc.build_stats.triangle_bytes),
             "coherent": {"rays": int(len(rays)), "gpu_mrays_per_s": coh, "nodes_per_ray": npr_c, "tris_per_ray": tpr_c,
                          "bytes_per_ray": 48 + npr_c * 80 + tpr_c * 48},
             "incoherent": {"rays": int(len(bounce)), "gpu_mrays_per_s": inc, "nodes_per_ray": npr_i, "tris_per_ray": tpr_i,
                            "bytes_per_ray": 48 + npr_i * 80 + tpr_i * 48}}
        for k in ("coherent", "incoherent"):
            e[k]["algorithmic_gb_per_s"] = e[k]["gpu_mrays_per_s"] * 1e6 * e[k]["bytes_per_ray"] / 1e9
        sc.close()
        if have_ref and "c3cpu" not in skip:
            # the CPU side runs 2^20 rays per batch on the full scene, its BVH build timed separately
            sbr = sb
            scr = lmcu.Scene(sbr)
            scr.set_stream(stream.cuda_stream)
            rr = scenes.camera_rays(sbr.camera, 1024, 1024)
            g_c, hg = device_batch(scr, rr, stream)
            pr = sbr.meshes[0][0]
            gnr = np.cross(pr[:, 1] - pr[:, 0], pr[:, 2] - pr[:, 0])
            gnr /= np.linalg.norm(gnr, axis=1, keepdims=True)
            rb = scenes.bounce_rays(rr, hg, lambda h: gnr[h["face"]].copy())
            g_i, hgb = device_batch(scr, rb, stream)
            rs = ref.RefScene(sbr)
            hr = rs.intersect(rr)
            cpu_c = len(rr) / rs.last_seconds / 1e6
            hrb = rs.intersect(rb)
            cpu_i = len(rb) / rs.last_seconds / 1e6
            e.update({"cpu_sample": "full 9 999 392-triangle scene, 1024x1024 = 2^20 coherent rays and their bounce rays",
                      "cpu_build_s": rs.build_seconds, "gpu_build_ms_same_scene": scr.build_stats.elapsed * 1e3,
                      "cpu_coherent_mrays_per_s": cpu_c, "cpu_incoherent_mrays_per_s": cpu_i,
                      "gpu_coherent_mrays_per_s_same_scene": g_c, "gpu_incoherent_mrays_per_s_same_scene": g_i,
                      "hits_vs_sahbvh_coherent": hit_agreement(hr, hg), "hits_vs_sahbvh_incoherent": hit_agreement(hrb, hgb)})
            rs.close()
            scr.close()
        e["wall_s"] = time.time() - t0
        out["C3"] = e
        print("C3", json.dumps(e), flush=True)

    # ---- C4: renderer::pt, 1000 instances (49.7 M triangles), glossy / glass / mirror / diffuse, env light, 3840x2160 -------
    if "c4" not in skip:
        t0 = time.time()
        sb = scenes.instanced_spheres(grid=10, num_theta=112)
        sc = lmcu.Scene(sb)
        sc.set_stream(stream.cuda_stream)
        p = abi.render_params(3840, 2160, spp=32, max_verts=20, seed=42)
        sc.render("pt", abi.render_params(3840, 2160, spp=4, max_verts=20, seed=42), download=False)
        _, st = sc.render("pt", p, download=False)
        _, stc = sc.render("pt", abi.render_params(3840, 2160, spp=2, max_verts=20, seed=42, count_traversal=True), download=False)
        b_ray = 48 + stc["bvh_nodes_per_ray"] * 80 + stc["bvh_tris_per_ray"] * 80
        e = {"workload": "renderer::pt, 1000 rotated/scaled instances of a 49 728-triangle sphere + ground, envconst light, "
                         "3840x2160, max_verts 20, mis; 32 spp timed (rate is spp-independent; the config's 256 spp = 8 such passes)",
             "triangles": int(sc.build_stats.num_triangles), "gpu_build_ms": sc.build_stats.elapsed * 1e3,
             "bvh_bytes": int(sc.build_stats.node_bytes + sc.build_stats.triangle_bytes), "fp64_triangles": not sc.build_stats.triangles_fp32,
             "gpu_mpaths_per_s": st["paths"] / st["elapsed"] / 1e6, "gpu_mrays_per_s": st["rays"] / st["elapsed"] / 1e6,
             "extend_share": st["extend_seconds"] / st["elapsed"], "shade_share": st["shade_seconds"] / st["elapsed"],
             "shadow_share": st["shadow_seconds"] / st["elapsed"],
             "nodes_per_ray": stc["bvh_nodes_per_ray"], "tris_per_ray": stc["bvh_tris_per_ray"], "bytes_per_ray": b_ray,
             "extend_algorithmic_gb_per_s": st["extend_rays"] * b_ray / st["extend_seconds"] / 1e9}
        sc.close()
        if have_ref and "c4cpu" not in skip:
            sbr = scenes.instanced_spheres(grid=3, num_theta=112)
            rs = ref.RefScene(sbr)
            pr = abi.render_params(960, 540, spp=4, max_verts=20, seed=42)
            rs.render("pt", abi.render_params(480, 270, spp=1, max_verts=20, seed=42))
            _, el, _ = rs.render("pt", pr)
            scr = lmcu.Scene(sbr)
            _, str_ = scr.render("pt", abi.render_params(3840, 2160, spp=16, max_verts=20, seed=42), download=False)
            e.update({"cpu_sample": "27 instances (1.34 M triangles: the reference's SAH build and memory do not reach 49.7 M), 960x540 @ 4 spp",
                      "cpu_build_s": rs.build_seconds, "cpu_mpaths_per_s": 960 * 540 * 4 / el / 1e6,
                      "gpu_mpaths_per_s_same_27_instance_scene": str_["paths"] / str_["elapsed"] / 1e6})
            scr.close()
            rs.close()
        e["wall_s"] = time.time() - t0
        out["C4"] = e
        print("C4", json.dumps(e), flush=True)

    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
