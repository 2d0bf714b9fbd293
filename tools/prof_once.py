"""One short run of a named config for ncu captures; prints how many rays the FIRST launch of the kernel of interest traces.

  python tools/prof_once.py c2     C2 pt 1920x1080 @ 8 spp, one wavefront        -> k_pt_extend_q / k_pt_shade / k_pt_shadow_q
  python tools/prof_once.py c4     C4 pt 3840x2160 @ 2 spp, one wavefront        -> k_pt_extend_q on the 49.7 M-triangle scene
  python tools/prof_once.py c4two  the same on the two-level layout                      -> k_pt_extend_q<TriObj>
  python tools/prof_once.py c3     C3 incoherent bounce batch, device buffers    -> k_intersect_q<TriF32>
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v3_b200"))
import numpy as np  # noqa: E402
import lmcu  # noqa: E402
from lmcu import abi, scenes  # noqa: E402

mode = sys.argv[1]
if mode in ("c2", "c4", "c4two"):
    if mode == "c2":
        sb, w, h, spp = scenes.cornell_spheres(64), 1920, 1080, 48
    else:
        sb, w, h, spp = scenes.instanced_spheres(grid=10, num_theta=112), 3840, 2160, 10
    sc = lmcu.Scene(sb, two_level=1 if mode == "c4two" else 0)
    _, st = sc.render("pt", abi.render_params(w, h, spp=spp, max_verts=20, wavefronts=1), download=False)
    # one wavefront of 16 Mi paths, refilled to capacity every iteration while new paths remain (>= 5 iterations here):
    # launches 1..4 of the extend kernel trace exactly 2^24 rays, a steady-state mix of camera and bounce rays
    first = min(w * h * spp, 1 << 24)
    print(json.dumps({"config": mode, "rays_per_full_launch": first, "mpaths_per_s": st["paths"] / st["elapsed"] / 1e6,
                      "extend_rays": st["extend_rays"], "extend_launches": st["extend_launches"]}))
else:
    import torch
    sb = scenes.heightfield_scene(2236)
    sc = lmcu.Scene(sb)
    rays = scenes.camera_rays(sb.camera, 1024, 1024)
    hits = sc.intersect(rays)
    p = sb.meshes[0][0]
    gn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    gn /= np.linalg.norm(gn, axis=1, keepdims=True)
    bounce = scenes.bounce_rays(rays, hits, lambda hh: gn[hh["face"]].copy())
    bounce = np.concatenate([bounce] * 4)
    d_rays = torch.from_numpy(bounce.view(np.float64).reshape(-1, 8)).cuda()
    d_hits = torch.empty((len(bounce), 4), dtype=torch.float64, device="cuda")
    for _ in range(2):
        sc.intersect_device(d_rays.data_ptr(), len(bounce), d_hits.data_ptr())
    torch.cuda.synchronize()
    print(json.dumps({"config": mode, "rays_per_device_launch": int(len(bounce)), "host_batches_first": [int(len(rays))]}))
