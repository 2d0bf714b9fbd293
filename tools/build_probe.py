"""Repeated upload + build of the C2 scene next to a resident scene with its path pools (diagnostic): wall and device time per build."""
import os, sys, time
sys.path.insert(0, 'lightmetrica-v3_b200')
import numpy as np
import lmcu
from lmcu import abi, scenes
sb = scenes.cornell_spheres(64)
keep = lmcu.Scene(sb)
keep.render("pt", abi.render_params(1920, 1080, spp=int(os.environ.get("BP_SPP", "64")), max_verts=20), download=False)
wall, dev = [], []
for it in range(int(os.environ.get("BP_ITERS", "60"))):
    t0 = time.perf_counter()
    sc = lmcu.Scene(sb)
    wall.append((time.perf_counter() - t0) * 1e3); dev.append(sc.build_stats.elapsed * 1e3)
    if os.environ.get("BP_RENDER"):
        sc.render("pt", abi.render_params(1920, 1080, spp=8, max_verts=20), download=False)
    sc.close()
wall, dev = np.array(wall), np.array(dev)
print(os.environ.get("LMCU_LIB", "default"), "wall ms: median %.2f max %.1f  >50ms: %d | device ms: median %.2f max %.1f" % (np.median(wall), wall.max(), (wall > 50).sum(), np.median(dev), dev.max()), np.round(np.sort(wall)[-5:], 1))
