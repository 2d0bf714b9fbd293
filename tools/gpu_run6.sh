set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -60 > gpurun_out/r2_gpu_tests_v6.log
tail -12 gpurun_out/r2_gpu_tests_v6.log
