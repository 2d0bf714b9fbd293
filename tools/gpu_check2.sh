#!/bin/bash
# perf sanity (C2 pt, bdpt) + plugin / multi / pylm regression subset on one B200
mkdir -p gpurun_out
L=gpurun_out/r2_check2.log
SECONDS=0
python - > $L 2>&1 <<'PY'
import sys
sys.path.insert(0, 'lightmetrica-v3_b200')
import lmcu
from lmcu import abi, scenes
sc = lmcu.Scene(scenes.cornell_spheres(64))
for spp in (256, 1024, 1024):
    _, st = sc.render("pt", abi.render_params(1920, 1080, spp=spp, max_verts=20), download=False)
    print("pt C2 %d spp: %.1f Mpaths/s" % (spp, st["paths"] / st["elapsed"] / 1e6))
n = 1920 * 1080 * 8
for mode in (0, 0, 1, 2, 3):
    _, st = sc.render("bdpt", abi.render_params(1920, 1080, max_verts=10, num_samples=n, bdpt_strategies=mode, primary_ray_sampling_mode=abi.PRIMARY_IMAGE), download=False)
    print("bdpt C2 max_verts 10 strategies %d: %.1f Msamples/s, %d rays" % (mode, n / st["elapsed"] / 1e6, st["rays"]))
PY
cat $L; echo "perf ${SECONDS}s"
timeout 80 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_multi.py tests/test_pylm_dropin.py -x -q -m gpu >> $L 2>&1; echo "subset rc=$? ${SECONDS}s"; tail -4 $L
