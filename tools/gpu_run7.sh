set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "save_load" 2>&1 | tail -3
TB_SPP=128 TB_C3=0 python tools/trav_bench.py "LMCU_WAVES=2" > gpurun_out/r2_ab_shade.log 2>&1
LMCU_LIB=$PWD/build_ab/liblmcu_shade10.so TB_SPP=128 TB_C3=0 python tools/trav_bench.py "LMCU_WAVES=2" >> gpurun_out/r2_ab_shade.log 2>&1
LMCU_LIB=$PWD/build_ab/liblmcu_shade8.so TB_SPP=128 TB_C3=0 python tools/trav_bench.py "LMCU_WAVES=2" >> gpurun_out/r2_ab_shade.log 2>&1
LMCU_LIB=$PWD/build_ab/liblmcu_shade8.so TB_SPP=128 TB_C3=0 python tools/trav_bench.py "LMCU_WAVES=2,LMCU_SHADE_GRID=12" >> gpurun_out/r2_ab_shade.log 2>&1
grep "C2 pt\|==" gpurun_out/r2_ab_shade.log
