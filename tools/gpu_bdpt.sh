#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py tests/test_pylm_dropin.py -m gpu -x -q -k "bdpt or pylm or dropin" > gpurun_out/r2_bdpt_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2_bdpt_tests.log
python - <<'PY'
import sys, json
sys.path.insert(0, 'lightmetrica-v3_b200'); sys.path.insert(0, '.')
import bench, lmcu
from lmcu import abi, scenes
sb = scenes.cornell_spheres(64)
sc = lmcu.Scene(sb)
print(json.dumps(bench.bdpt_rate(sc, abi, sb)))
for mv in (4, 6, 10, 20):
    n = 1920*1080*2
    sc.render("bdpt", abi.render_params(1920,1080,max_verts=mv,num_samples=n//4,primary_ray_sampling_mode=abi.PRIMARY_IMAGE), download=False)
    _, st = sc.render("bdpt", abi.render_params(1920,1080,max_verts=mv,num_samples=n,primary_ray_sampling_mode=abi.PRIMARY_IMAGE), download=False)
    print(mv, st["paths"]/st["elapsed"]/1e6, "Msamples/s", st["rays"]/st["paths"], "rays/sample")
PY
