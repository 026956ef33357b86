set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_base.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
python tools/rns_bench.py 2>&1 | tail -3 | cut -c1-900
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; tail -c 300 gpurun_out/bench_n1c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1c.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","ntt_64k_per_s","gpu_launches")}, "e2e", d["e2e"]["value"], "c5", d.get("config5"))
PY
