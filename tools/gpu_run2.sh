set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_plugin.py -m gpu -q -s -k "batched or missed or scalar_intersect_rate" 2>&1 | tail -30 > gpurun_out/r2_plugin_tests_v2.log
tail -5 gpurun_out/r2_plugin_tests_v2.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err
tail -c 1500 gpurun_out/r2_bench_v2.err
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_pt_extend_q -c 1 -o gpurun_out/r2_prof_c2_extend -f python tools/prof_once.py c2 > gpurun_out/r2_prof_c2.log 2>&1
$NCU -k regex:k_pt_shade -c 1 -o gpurun_out/r2_prof_c2_shade -f python tools/prof_once.py c2 >> gpurun_out/r2_prof_c2.log 2>&1
$NCU -k regex:k_pt_extend_q -c 1 -o gpurun_out/r2_prof_c4_extend -f python tools/prof_once.py c4 > gpurun_out/r2_prof_c4.log 2>&1
$NCU -k regex:k_intersect_q -s 1 -c 1 -o gpurun_out/r2_prof_c3_intersect -f python tools/prof_once.py c3 > gpurun_out/r2_prof_c3.log 2>&1
tail -2 gpurun_out/r2_prof_c2.log gpurun_out/r2_prof_c4.log gpurun_out/r2_prof_c3.log
ls -la gpurun_out/*.ncu-rep | tail -5
