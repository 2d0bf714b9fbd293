#!/bin/bash
# the two large GPU test files on 4 xdist workers (one B200) + the bench's bdpt leg
mkdir -p gpurun_out
L=gpurun_out/r2_check3.log
SECONDS=0
timeout 88 python -m pytest tests/test_gpu_parity.py tests/test_gpu_two_level.py -m gpu -n 4 -v -p no:cacheprovider > $L 2>&1; echo "tests rc=$? ${SECONDS}s"
grep -c PASSED $L; grep -n "FAILED\|ERROR" $L | head -20; tail -3 $L
timeout 14 python - > gpurun_out/r2_check3_bdpt.json 2>&1 <<'PY'
import json, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'lightmetrica-v3_b200')
import bench, lmcu
from lmcu import abi, scenes
sb = scenes.cornell_spheres(64)
sc = lmcu.Scene(sb)
print(json.dumps(bench.bdpt_rate(sc, abi, sb)))
PY
echo "bdpt_rate rc=$? ${SECONDS}s"; tail -c 1500 gpurun_out/r2_check3_bdpt.json
