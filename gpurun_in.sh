set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_base.py -x -q -m gpu 2>&1 | tail -3
python tools/ntt_bench.py | cut -c1-420
python bench.py --steps 20 --warmup 3 --no-c5 > gpurun_out/bench_n1d.json 2> gpurun_out/bench_n1d.err; tail -c 300 gpurun_out/bench_n1d.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1d.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","ntt_64k_per_s","gpu_launches")}, "e2e", d["e2e"]["value"], "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_source"][:60], "cpu", d["cpu_baseline"]["value"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c5 > /dev/null 2>&1
