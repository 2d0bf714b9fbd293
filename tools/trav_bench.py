"""A/B harness for traversal / pipeline tuning knobs on a GPU box (diagnostic, not the bench).

  python tools/trav_bench.py "LMCU_FETCH=12" "LMCU_FETCH=16,LMCU_CTAS=5" "LMCU_WAVES=1,LMCU_PT_POOL=8388608" ...

Knobs liblmcu reads (tuning only): LMCU_FETCH (idle lanes before a warp refills), LMCU_CTAS (traversal CTAs per SM),
LMCU_SHADE_GRID (shade CTAs per SM), LMCU_WAVES (staggered wavefronts), LMCU_PT_POOL (paths in flight per wavefront),
LMCU_LEAF (triangles per leaf slot), LMCU_BIG (size-class threshold denominator, 0 = off), LMCU_TRI64 (1 = FP64 triangle
records, 0 = float records); LMCU_LIB points the binding at another build of the library.

Each argument is a comma-separated list of environment settings applied before the scenes are created
(liblmcu reads its knobs at lmcu_scene_create).  Per setting it prints:
  - C2 renderer::pt 1920x1080 @ SPP: Mpaths/s and the per-kernel device seconds
  - C3 accel batches with DEVICE-resident rays: coherent camera rays and incoherent bounce rays, Mrays/s
"""
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

SPP = int(os.environ.get("TB_SPP", "32"))
C3_N = int(os.environ.get("TB_C3", "2236"))
C3_RES = int(os.environ.get("TB_C3_RES", "2048"))
TB_W, TB_H = int(os.environ.get("TB_W", "1920")), int(os.environ.get("TB_H", "1080"))


def tri_normals(sb):
    p = sb.meshes[0][0]
    gn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    return gn / np.linalg.norm(gn, axis=1, keepdims=True)


def device_batch(sc, rays, reps=3):
    d_rays = torch.from_numpy(rays.view(np.float64).reshape(-1, 8)).cuda()
    d_hits = torch.empty((len(rays), 4), dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()             # a real stream: handle 0 (the legacy default stream) means "the scene's own stream"
    st.wait_stream(torch.cuda.current_stream())
    sc.set_stream(st.cuda_stream)
    best = 1e9
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        sc.intersect_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    hits = d_hits.cpu().numpy().view(scenes.HIT_DTYPE).reshape(-1)
    return len(rays) / best / 1e6, hits


def main():
    settings = sys.argv[1:] or ["LMCU_FETCH=12"]
    c2 = scenes.cornell_spheres(64)
    t = time.time()
    c3 = scenes.heightfield_scene(C3_N) if C3_N > 0 else None
    c3_rays = c3_bounce = None
    ref_hits = {}
    print("scenes ready in %.1fs" % (time.time() - t), flush=True)
    for setting in settings:
        for kv in setting.split(","):
            if kv:
                k, v = kv.split("=")
                os.environ[k] = v
        print("== %s" % setting, flush=True)
        sc = lmcu.Scene(c2)
        for i in range(2):
            _, st = sc.render("pt", abi.render_params(TB_W, TB_H, spp=SPP, max_verts=20), download=False)
        print("  C2 pt %d spp: %.1f Mpaths/s  %.1f Mrays/s | extend %.4f shade %.4f shadow %.4f total %.4f s | launches %d" % (
            SPP, st["paths"] / st["elapsed"] / 1e6, st["rays"] / st["elapsed"] / 1e6, st["extend_seconds"],
            st["shade_seconds"], st["shadow_seconds"], st["elapsed"], st["kernel_launches"]), flush=True)
        film = sc.film_download(TB_W, TB_H)
        print("  C2 film mean %s" % (film.mean(axis=(0, 1)) / SPP), flush=True)
        rays = scenes.camera_rays(c2.camera, 1920, 1080)
        mr, hits = device_batch(sc, rays)
        print("  C2 camera rays (device): %.1f Mrays/s" % mr, flush=True)
        if "c2" in ref_hits:
            same = (hits["primitive"] == ref_hits["c2"]["primitive"]) & (hits["face"] == ref_hits["c2"]["face"]) & (hits["t"] == ref_hits["c2"]["t"])
            print("  C2 hits identical to first setting: %s (%d differ)" % (bool(same.all()), int((~same).sum())))
        else:
            ref_hits["c2"] = hits
        sc.close()
        if c3 is not None:
            sc = lmcu.Scene(c3)
            if c3_rays is None:
                c3_rays = scenes.camera_rays(c3.camera, C3_RES, C3_RES)
            mr, hits = device_batch(sc, c3_rays)
            print("  C3 coherent %d rays (device): %.1f Mrays/s  hit %.3f" % (len(c3_rays), mr, (hits["primitive"] >= 0).mean()), flush=True)
            if c3_bounce is None:
                gn = tri_normals(c3)
                c3_bounce = scenes.bounce_rays(c3_rays, hits, lambda h: gn[h["face"]].copy())
            mr2, hits2 = device_batch(sc, c3_bounce)
            print("  C3 incoherent %d rays (device): %.1f Mrays/s  hit %.3f" % (len(c3_bounce), mr2, (hits2["primitive"] >= 0).mean()), flush=True)
            if "c3" in ref_hits:
                a, b = ref_hits["c3"]
                ok = all(((x["primitive"] == y["primitive"]) & (x["face"] == y["face"]) & (x["t"] == y["t"])).all() for x, y in ((hits, a), (hits2, b)))
                print("  C3 hits identical to first setting: %s" % ok)
            else:
                ref_hits["c3"] = (hits, hits2)
            sc.close()


if __name__ == "__main__":
    main()
