set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; tail -c 300 gpurun_out/bench_n1b.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1b.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","ntt_64k_per_s","gpu_launches")}, "e2e", d["e2e"])
PY
ncu --set full --import-source on --clock-control none -k regex:"crt_kernel|cyclo_reduce" -s 6 -c 3 -o gpurun_out/r02_rns_full python bench.py --steps 2 --warmup 1 --no-cpu --no-c5 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
