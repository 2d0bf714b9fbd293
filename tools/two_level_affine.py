"""Where the two-level layout pays in rays per second too: instances under shear / anisotropic scale (scenes.affine_instances),
whose object-space boxes are far tighter than the world-space boxes of the flattened triangles.  pt at 1920x1080, 16 spp."""
import json, sys
sys.path.insert(0, 'lightmetrica-v3_b200')
import lmcu
from lmcu import abi, scenes
out = {}
for n_inst, nth in ((40, 10), (400, 48)):
    sb = scenes.affine_instances(0, n_inst, nth)
    r = {"triangles": sb.num_triangles()}
    for tl in (-1, 1):
        sc = lmcu.Scene(sb, two_level=tl)
        _, stc = sc.render("pt", abi.render_params(1920, 1080, spp=2, max_verts=10, count_traversal=True), download=False)
        sc.render("pt", abi.render_params(1920, 1080, spp=4, max_verts=10), download=False)
        f, st = sc.render("pt", abi.render_params(1920, 1080, spp=16, max_verts=10))
        r["two_level" if tl > 0 else "flattened"] = {"mpaths_per_s": round(st["paths"] / st["elapsed"] / 1e6, 1), "mrays_per_s": round(st["rays"] / st["elapsed"] / 1e6, 1),
            "nodes_per_ray": round(stc["bvh_nodes_per_ray"], 2), "tris_per_ray": round(stc["bvh_tris_per_ray"], 2), "mean": float(f.mean()),
            "build_ms": round(sc.build_stats.elapsed * 1e3, 2)}
        sc.close()
    out["%d instances, num_theta %d" % (n_inst, nth)] = r
print(json.dumps(out))
