set -x
mkdir -p gpurun_out
LMCU_TRACE=pool python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "g1 or edge or c2_vs_reference or additive or wavefront or fp64 or scalar" 2>&1 | tail -15 > gpurun_out/r2_pool_tests.log
tail -5 gpurun_out/r2_pool_tests.log
TB_SPP=64 python tools/trav_bench.py "LMCU_TRACE=queue" "LMCU_TRACE=pool,LMCU_POOL_FETCH=16" "LMCU_TRACE=pool,LMCU_POOL_FETCH=8" "LMCU_TRACE=pool,LMCU_POOL_FETCH=32" "LMCU_TRACE=pool,LMCU_POOL_FETCH=16,LMCU_WAVES=1" "LMCU_TRACE=queue,LMCU_WAVES=1" > gpurun_out/r2_ab_pool.log 2>&1
cat gpurun_out/r2_ab_pool.log
