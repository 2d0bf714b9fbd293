set -x
TB_SPP=1024 TB_C3=0 python tools/trav_bench.py "LMCU_CHUNK=64" "LMCU_CHUNK=256" "LMCU_CHUNK=1024" "LMCU_CHUNK=32" > gpurun_out/r2_ab_chunk2.log 2>&1
TB_W=3840 TB_H=2160 TB_SPP=256 TB_C3=0 python tools/trav_bench.py "LMCU_CHUNK=64" "LMCU_CHUNK=256" >> gpurun_out/r2_ab_chunk2.log 2>&1
grep "C2 pt\|==" gpurun_out/r2_ab_chunk2.log
