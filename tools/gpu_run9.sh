set -x
TB_SPP=512 TB_C3=0 python tools/trav_bench.py "LMCU_FETCH=12" "LMCU_FETCH=8" "LMCU_FETCH=16" > gpurun_out/r2_ab_stream.log 2>&1
LMCU_LIB=$PWD/build_ab/liblmcu_nostream.so TB_SPP=512 TB_C3=0 python tools/trav_bench.py "LMCU_FETCH=12" >> gpurun_out/r2_ab_stream.log 2>&1
TB_W=3840 TB_H=2160 TB_SPP=128 TB_C3=0 python tools/trav_bench.py "LMCU_FETCH=12" >> gpurun_out/r2_ab_stream.log 2>&1
LMCU_LIB=$PWD/build_ab/liblmcu_nostream.so TB_W=3840 TB_H=2160 TB_SPP=128 TB_C3=0 python tools/trav_bench.py "LMCU_FETCH=12" >> gpurun_out/r2_ab_stream.log 2>&1
grep "C2 pt\|==" gpurun_out/r2_ab_stream.log
