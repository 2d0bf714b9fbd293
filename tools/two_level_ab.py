"""C4 (1000 instances of a 49 728-triangle sphere) on the flattened tree and on the two-level layout: resident bytes, build time,
node / triangle visits per ray, pt throughput at 3840x2160 and the extend kernel timed alone.  LMCU_LIB selects the build,
TWO_LEVEL_AB_GRID / _SPP shrink the run."""
import json, os, sys
sys.path.insert(0, 'lightmetrica-v3_b200')
import numpy as np
import lmcu
from lmcu import abi, scenes
grid = int(os.environ.get("TWO_LEVEL_AB_GRID", 10))
spp = int(os.environ.get("TWO_LEVEL_AB_SPP", 8))
layouts = [int(x) for x in os.environ.get("TWO_LEVEL_AB_LAYOUTS", "1,-1").split(",")]
sb = scenes.instanced_spheres(grid=grid, num_theta=112)
W, H = 3840, 2160
out = {"lib": os.environ.get("LMCU_LIB", "default"), "grid": grid, "spp": spp}
for tl in layouts:
    sc = lmcu.Scene(sb, two_level=tl)
    bs = sc.build_stats
    r = {"two_level": int(bs.two_level), "build_ms": round(bs.elapsed * 1e3, 2), "resident_mb": round((bs.node_bytes + bs.triangle_bytes) / 1e6, 2),
         "nodes": int(bs.num_wide_nodes)}
    _, stc = sc.render("pt", abi.render_params(W, H, spp=1, max_verts=20, count_traversal=True), download=False)
    r["nodes_per_ray"], r["tris_per_ray"] = round(stc["bvh_nodes_per_ray"], 3), round(stc["bvh_tris_per_ray"], 3)
    sc.render("pt", abi.render_params(W, H, spp=2, max_verts=20), download=False)
    f, st = sc.render("pt", abi.render_params(W, H, spp=spp, max_verts=20))
    r["mpaths_per_s"] = round(st["paths"] / st["elapsed"] / 1e6, 1)
    r["mrays_per_s"] = round(st["rays"] / st["elapsed"] / 1e6, 1)
    r["mean"] = float(f.mean())
    _, solo = sc.render("pt", abi.render_params(W, H, spp=spp, max_verts=20, wavefronts=1), download=False)
    r["solo_mpaths_per_s"] = round(solo["paths"] / solo["elapsed"] / 1e6, 1)
    r["extend_mrays_per_s"] = round(solo["extend_rays"] / solo["extend_seconds"] / 1e6, 1)
    r["extend_share"] = round(solo["extend_seconds"] / solo["elapsed"], 3)
    for k in ("shade_seconds", "shadow_seconds"):
        if k in solo: r[k.replace("_seconds", "_share")] = round(solo[k] / solo["elapsed"], 3)
    out["two_level" if bs.two_level else "flattened"] = r
    sc.close()
print(json.dumps(out))
