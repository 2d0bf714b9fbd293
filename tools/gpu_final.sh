#!/bin/bash
# round-end rehearsal: the driver's own bench command on one B200 (tests + smoke: see the first half of the log)
mkdir -p gpurun_out
if [ "$1" != "bench" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_gpu_tests_final.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
fi
SECONDS=0
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'], d['gpu_launches'])
print({k:(v if not isinstance(v,dict) else '...') for k,v in d['roofline'].items()})
print(d.get('parity_c2'))
PY
