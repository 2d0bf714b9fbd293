#!/bin/bash
# two-level layout: parity tests, then C4 on both layouts (and the 5-CTA build of the two-level kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_two_level.py -x -q -s > gpurun_out/r2_two_level_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2_two_level_tests.log
timeout 600 python tools/two_level_ab.py 2>&1 | tail -3 | tee gpurun_out/r2_two_level_ab.log
if [ -f build_ab/liblmcu_m4.so ]; then TWO_LEVEL_AB_LAYOUTS=1 LMCU_LIB=build_ab/liblmcu_m4.so timeout 600 python tools/two_level_ab.py 2>&1 | tail -3 | tee -a gpurun_out/r2_two_level_ab.log; fi
