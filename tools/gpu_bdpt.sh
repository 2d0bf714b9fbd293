#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py -m gpu -x -q -k "bdpt" > gpurun_out/r2_bdpt_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_bdpt_tests.log
python tools/ab_bdpt.py
python - <<'PY'
import sys, json
sys.path.insert(0, 'lightmetrica-v3_b200'); sys.path.insert(0, '.')
import bench, lmcu
from lmcu import abi, scenes
sb = scenes.cornell_spheres(64)
sc = lmcu.Scene(sb)
print(json.dumps(bench.bdpt_rate(sc, abi, sb)))
PY
