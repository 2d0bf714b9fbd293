set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py tests/test_gpu_plugin.py -m gpu -q -s -k "multi or across or two_processes or reduce or save_state or missed or cached" 2>&1 | tail -40 > gpurun_out/r2_multi_tests_2gpu.log
tail -8 gpurun_out/r2_multi_tests_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 1200 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"])
print(json.dumps(d["multi_gpu_check"])); print(json.dumps(d["c5_strong"]))
PY
