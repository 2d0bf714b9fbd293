set -x
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 600 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n8.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"])
print(json.dumps(d["multi_gpu_check"])); print(json.dumps(d["c5_strong"]))
PY
