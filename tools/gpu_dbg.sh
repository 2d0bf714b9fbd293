#!/bin/bash
mkdir -p gpurun_out
python -X faulthandler -m pytest tests/test_gpu_plugin.py tests/test_pylm_dropin.py -m gpu -x -q -k "lt_cuda or pylm or dropin" > gpurun_out/dbg_c.log 2>&1; echo "lt+pylm rc=$?"; tail -5 gpurun_out/dbg_c.log
python -X faulthandler -m pytest tests/test_pylm_dropin.py -m gpu -x -q > gpurun_out/dbg_d.log 2>&1; echo "pylm alone rc=$?"; tail -5 gpurun_out/dbg_d.log
python -X faulthandler -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py tests/test_pylm_dropin.py -m gpu -x -q -k "bdpt or pylm or dropin" > gpurun_out/dbg_a.log 2>&1; echo "subset rc=$?"; tail -5 gpurun_out/dbg_a.log
