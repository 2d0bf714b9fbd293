set -x
TB_SPP=1024 TB_C3=0 python tools/trav_bench.py "LMCU_WAVES=2" "LMCU_WAVES=3" "LMCU_WAVES=2,LMCU_PT_POOL=33554432" "LMCU_WAVES=3,LMCU_PT_POOL=8388608" "LMCU_WAVES=4,LMCU_PT_POOL=8388608" > gpurun_out/r2_ab_waves.log 2>&1
grep "C2 pt\|==" gpurun_out/r2_ab_waves.log
