#!/bin/bash
# single-strategy path renderers (LmcuRenderParams::bdpt_strategies) + regression subset + smoke on one B200
mkdir -p gpurun_out
L=gpurun_out/r2_path_strategies.log
SECONDS=0
timeout 100 python -m pytest tests/test_zz_gpu_path_strategies.py -q > $L 2>&1; echo "paths rc=$? ${SECONDS}s"; tail -25 $L
timeout 50 python -m pytest tests/test_gpu_parity.py -x -q -k "bdpt_kernels_equal or bdpt_agrees or accel_save_load or g1_closest_hit_vs_golden or g1_tiny" >> $L 2>&1; echo "subset rc=$? ${SECONDS}s"; tail -4 $L
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "total ${SECONDS}s"
