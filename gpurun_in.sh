set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 300 gpurun_out/bench_final_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_final_n1.json').read().strip().splitlines()[-1])
for k in ("metric","value","unit","ms_per_step","steps","warmup","ntt_64k_per_s","gpu_launches","verified","clocks"): print(k, d.get(k))
print("e2e", d["e2e"]); print("roofline", {k:v for k,v in d["roofline"].items() if k!="secondary"}); print("cpu", d["cpu_baseline"]); print("c5", d.get("config5")); print("mulzzx", d.get("e2e_mulzzx")); print("relin", d.get("relin_config3"))
PY
