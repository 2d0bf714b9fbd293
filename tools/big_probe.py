"""A/B of the builder's size-class threshold (LMCU_BIG) on C2 and C4: nodes / triangles per ray and pt throughput."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v3_b200"))
import lmcu  # noqa: E402
from lmcu import abi, scenes  # noqa: E402

c2 = scenes.cornell_spheres(64)
c4s = scenes.instanced_spheres(grid=3, num_theta=112)
c4 = scenes.instanced_spheres(grid=10, num_theta=112) if os.environ.get("FULL_C4") else None
for big in sys.argv[1:]:
    os.environ["LMCU_BIG"] = big
    for name, sb, w, h, spp in (("C2", c2, 1920, 1080, 128), ("C4/27", c4s, 1920, 1080, 32), ("C4", c4, 3840, 2160, 16)):
        if sb is None:
            continue
        sc = lmcu.Scene(sb)
        sc.render("pt", abi.render_params(w, h, spp=4, max_verts=20), download=False)
        _, st = sc.render("pt", abi.render_params(w, h, spp=spp, max_verts=20), download=False)
        _, stc = sc.render("pt", abi.render_params(w, h, spp=2, max_verts=20, count_traversal=True), download=False)
        print("LMCU_BIG=%s %s: %.1f Mpaths/s | extend %.4f shade %.4f shadow %.4f | nodes/ray %.2f tris/ray %.2f shadow nodes %.2f tris %.2f | wide nodes %d" % (
            big, name, st["paths"] / st["elapsed"] / 1e6, st["extend_seconds"], st["shade_seconds"], st["shadow_seconds"],
            stc["bvh_nodes_per_ray"], stc["bvh_tris_per_ray"], stc["shadow_nodes_per_ray"], stc["shadow_tris_per_ray"],
            sc.build_stats.num_wide_nodes), flush=True)
        sc.close()
