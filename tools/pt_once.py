"""One small C2 pt render (for ncu captures): python tools/pt_once.py [spp] [width] [height]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v3_b200"))
import lmcu  # noqa: E402
from lmcu import abi, scenes  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
sc = lmcu.Scene(scenes.cornell_spheres(64))
_, st = sc.render("pt", abi.render_params(w, h, spp=spp, max_verts=20), download=False)
print("pt %dx%d @ %d spp: %.1f Mpaths/s, %d launches" % (w, h, spp, st["paths"] / st["elapsed"] / 1e6, st["kernel_launches"]))
