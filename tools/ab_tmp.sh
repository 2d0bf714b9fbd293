python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "g1 or edge or additive or scalar or pool" 2>&1 | tail -3
TB_SPP=512 python tools/trav_bench.py "LMCU_WAVES=2" > gpurun_out/r2_ab_defer.log 2>&1
LMCU_LIB=$PWD/build_ab/liblmcu_prev.so TB_SPP=512 python tools/trav_bench.py "LMCU_WAVES=2" >> gpurun_out/r2_ab_defer.log 2>&1
grep "C2 pt\|==\|C3\|C2 camera" gpurun_out/r2_ab_defer.log
