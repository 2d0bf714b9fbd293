set -x
NCU="ncu --set full --clock-control none --import-source on"
LMCU_TRACE=pool $NCU -k regex:k_pt_extend_ -s 4 -c 1 -o gpurun_out/r2_prof_c2_extend_pool -f python tools/prof_once.py c2 > gpurun_out/r2_prof_pool.log 2>&1
LMCU_TRACE=queue $NCU -k regex:k_pt_extend_ -s 4 -c 1 -o gpurun_out/r2_prof_c2_extend_queue -f python tools/prof_once.py c2 > gpurun_out/r2_prof_queue.log 2>&1
$NCU -k regex:k_pt_shade -s 4 -c 1 -o gpurun_out/r2_prof_c2_shade -f python tools/prof_once.py c2 >> gpurun_out/r2_prof_queue.log 2>&1
$NCU -k regex:k_pt_shadow_ -s 4 -c 1 -o gpurun_out/r2_prof_c2_shadow -f python tools/prof_once.py c2 >> gpurun_out/r2_prof_queue.log 2>&1
$NCU -k regex:k_pt_extend_ -s 4 -c 1 -o gpurun_out/r2_prof_c4_extend -f python tools/prof_once.py c4 > gpurun_out/r2_prof_c4.log 2>&1
grep -h config gpurun_out/r2_prof_pool.log gpurun_out/r2_prof_queue.log gpurun_out/r2_prof_c4.log
