#!/bin/bash
# the Python mirror's renderer names for the path-structure renderers (lmcu.Scene.render) on one B200
mkdir -p gpurun_out
timeout 15 python - > gpurun_out/r2_check4.log 2>&1 <<'PY'
import sys
import numpy as np
sys.path.insert(0, 'lightmetrica-v3_b200')
import lmcu
from lmcu import abi, scenes
sc = lmcu.Scene(scenes.cornell_spheres(12))
w, h = 64, 36
kw = dict(max_verts=6, num_samples=w * h * 16, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
for name, mode in (("pt_naive_path", 1), ("pt_nee_path", 2), ("lt_nee_path", 3), ("bdptopt", 0)):
    p = abi.render_params(w, h, **kw)
    a, _ = sc.render(name, p)
    b, _ = sc.render("bdpt", abi.render_params(w, h, bdpt_strategies=mode, **kw))
    assert p.bdpt_strategies == 0
    assert a.mean() > 0 and np.allclose(a, b, rtol=1e-4, atol=1e-6), name
    print(name, float(a.mean()), "ok")
PY
echo "rc=$?"; cat gpurun_out/r2_check4.log
