set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -40 > gpurun_out/r2_gpu_tests_v10.log
tail -5 gpurun_out/r2_gpu_tests_v10.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v10.json 2> gpurun_out/r2_bench_v10.err
tail -c 800 gpurun_out/r2_bench_v10.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r2_launches_v10.csv python bench.py --steps 1 --warmup 1 --spp 32 --no-configs --no-c5 > gpurun_out/r2_launches_bench.log 2>&1
tail -3 gpurun_out/r2_launches_v10.csv
