#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py -m gpu -x -q -k "bdpt" > gpurun_out/r2_bdpt_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_bdpt_tests.log
python tools/ab_bdpt.py
