"""A/B of liblmcu builds on renderer::bdpt throughput (LMCU_LIB selects the build): C2 scene, 1920x1080."""
import json, os, sys
sys.path.insert(0, 'lightmetrica-v3_b200')
import lmcu
from lmcu import abi, scenes
sb = scenes.cornell_spheres(64)
sc = lmcu.Scene(sb)
out = {"lib": os.environ.get("LMCU_LIB", "default")}
for mv in (4, 10, 20):
    n = 1920 * 1080 * 2
    kw = dict(max_verts=mv, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    sc.render("bdpt", abi.render_params(1920, 1080, num_samples=n // 4, **kw), download=False)
    f, st = sc.render("bdpt", abi.render_params(1920, 1080, num_samples=n, **kw))
    out["mv%d" % mv] = round(st["paths"] / st["elapsed"] / 1e6, 2)
    out["mean%d" % mv] = float(f.mean())
print(json.dumps(out))
