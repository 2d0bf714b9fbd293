"""A/B of liblmcu builds on renderer::pt throughput (LMCU_LIB selects the build): C2 at 1920x1080, 256 spp, best of 3; with
AB_C4=1 also C4 (flattened) at 3840x2160, 8 spp."""
import json, os, sys
sys.path.insert(0, 'lightmetrica-v3_b200')
import lmcu
from lmcu import abi, scenes
sc = lmcu.Scene(scenes.cornell_spheres(64))
out = {"lib": os.environ.get("LMCU_LIB", "default")}
sc.render("pt", abi.render_params(1920, 1080, spp=32, max_verts=20), download=False)
best, solo = 0.0, None
for _ in range(3):
    f, st = sc.render("pt", abi.render_params(1920, 1080, spp=256, max_verts=20))
    best = max(best, st["paths"] / st["elapsed"] / 1e6)
_, s1 = sc.render("pt", abi.render_params(1920, 1080, spp=128, max_verts=20, wavefronts=1), download=False)
out.update({"c2_mpaths_per_s": round(best, 1), "mean": float(f.mean()),
            "extend_mrays_per_s": round(s1["extend_rays"] / s1["extend_seconds"] / 1e6, 1),
            "shares": [round(s1[k] / s1["elapsed"], 3) for k in ("extend_seconds", "shade_seconds", "shadow_seconds")]})
sc.close()
if os.environ.get("AB_C4"):
    sc = lmcu.Scene(scenes.instanced_spheres(grid=10, num_theta=112))
    sc.render("pt", abi.render_params(3840, 2160, spp=2, max_verts=20), download=False)
    _, st = sc.render("pt", abi.render_params(3840, 2160, spp=8, max_verts=20), download=False)
    out["c4_mpaths_per_s"] = round(st["paths"] / st["elapsed"] / 1e6, 1)
    sc.close()
print(json.dumps(out))
