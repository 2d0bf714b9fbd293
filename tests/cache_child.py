"""Child process of test_block_cache_cap_and_lru (LMCU_CACHE_GB is read when the library first parks a block)."""
import json
import sys

import numpy as np

import lmcu
from lmcu import abi, scenes

cap = int(float(sys.argv[1]) * (1 << 30))
out = {"cap": cap, "parked": []}
sb = scenes.cornell_spheres(16)
rays = scenes.camera_rays(sb.camera, 64, 36)
ref_hits = None
films = []
for k in range(6):
    sc = lmcu.Scene(sb)
    h = sc.intersect(rays)
    if ref_hits is None:
        ref_hits = h
    assert (h["t"] == ref_hits["t"]).all() and (h["primitive"] == ref_hits["primitive"]).all()
    # pools and films of a different size every time: under a small cap the older ones must make room
    f, _ = sc.render("pt", abi.render_params(96 + 32 * k, 54 + 18 * k, spp=8, max_verts=6))
    films.append(float(f.mean()))
    sc.close()
    out["parked"].append(lmcu.cache_stats(0))
out["film_means"] = films
# the sizes of the LAST scene are still parked: building it again allocates nothing new
before = lmcu.cache_stats(0)
sc = lmcu.Scene(sb)
sc.render("pt", abi.render_params(96 + 32 * 5, 54 + 18 * 5, spp=8, max_verts=6))
sc.close()
out["reuse"] = [before, lmcu.cache_stats(0)]
lmcu.cache_trim(0)
out["after_trim"] = lmcu.cache_stats(0)
print(json.dumps(out))
