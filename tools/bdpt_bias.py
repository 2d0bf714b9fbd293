"""bdpt against pt (MIS) on the device at high sample counts: two estimators of the same image, so the channel means must
agree to their Monte-Carlo error (pt itself is within 3e-4 of the reference at 8192 spp, test_g3_pt_high_spp_bias)."""
import json, os, sys, tempfile
import numpy as np
sys.path.insert(0, 'lightmetrica-v3_b200')
import lmcu
from lmcu import abi, scenes
w, h, spp, B = 480, 270, 256, 8
tmp = tempfile.mkdtemp()
for name, sb in (("cornell_spheres", scenes.cornell_spheres(32)), ("instanced_envconst", scenes.instanced_spheres(grid=2, num_theta=24)),
                 ("delta_lights", scenes.delta_lights_scene(2, 16)),
                 ("envmap", scenes.instanced_spheres(grid=2, num_theta=16, env=scenes.env_light(os.path.join(tmp, "e.hdr"), 64, 32)))):
    sc = lmcu.Scene(sb)
    n = w * h * spp
    mb, mp = [], []
    for b in range(B):
        kw = dict(max_verts=8, num_samples=n, sample_offset=b * n, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
        mb.append(sc.render("bdpt", abi.render_params(w, h, **kw))[0].mean(axis=(0, 1), dtype=np.float64))
        mp.append(sc.render("pt", abi.render_params(w, h, **kw))[0].mean(axis=(0, 1), dtype=np.float64))
    mb, mp = np.array(mb), np.array(mp)
    d = mb.mean(0) - mp.mean(0)
    se = np.sqrt(mb.var(0, ddof=1) / B + mp.var(0, ddof=1) / B)
    print(json.dumps({"scene": name, "spp_total": spp * B, "bdpt": mb.mean(0).tolist(), "pt": mp.mean(0).tolist(),
                      "rel_diff": (d / mp.mean(0)).tolist(), "z": (d / se).tolist()}))
    sc.close()
