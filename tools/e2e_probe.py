"""Where the end-to-end step of bench.py spends its wall time (diagnostic)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v3_b200"))
import torch  # noqa: E402
import lmcu  # noqa: E402
from lmcu import abi, scenes  # noqa: E402

sb = scenes.cornell_spheres(64)
keep = lmcu.Scene(sb)                     # the resident scene of bench.py holds its own pools
keep.render("pt", abi.render_params(1920, 1080, spp=4, max_verts=20), download=False)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for it in range(int(os.environ.get("E2E_ITERS", "4"))):
    t = [time.perf_counter()]
    d = sb.desc(); t.append(time.perf_counter())
    sc = lmcu.Scene(sb); t.append(time.perf_counter())
    _, st = sc.render("pt", abi.render_params(1920, 1080, spp=spp, max_verts=20, rescale=0), download=False); t.append(time.perf_counter())
    sc.film_reduce_peers([], 1.0 / spp); t.append(time.perf_counter())
    film = sc.film_download(1920, 1080); t.append(time.perf_counter())
    torch.cuda.synchronize(); sc.close(); t.append(time.perf_counter())
    names = ["desc()", "Scene() upload+build", "render", "reduce", "download", "close"]
    print("iter %d: " % it + "  ".join("%s %.1f ms" % (n, (b - a) * 1e3) for n, a, b in zip(names, t[:-1], t[1:])) +
          "  | device render %.1f ms  build %.1f ms" % (st["elapsed"] * 1e3, sc.build_stats.elapsed * 1e3), flush=True)
