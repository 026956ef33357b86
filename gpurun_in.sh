set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/ntt_bench.py | cut -c1-420
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","ntt_64k_per_s","gpu_launches")}, "e2e", d["e2e"]["value"], "roof", d["roofline"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c5 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:ntt4_pass -s 6 -c 2 -o gpurun_out/r02_gen4b_ntt python tools/ntt_bench.py --one > /dev/null 2>&1
