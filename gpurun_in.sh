set -x
CUHE_B200_LIB=$PWD/variants/p1x6/libcuhe_b200.so python tools/ntt_bench.py | cut -c1-420
python tools/ntt_bench.py | cut -c1-420
python tools/ntt_bench.py --sweep > gpurun_out/ntt_sweep_gen4.json; python - <<PY
import json
d=json.load(open('gpurun_out/ntt_sweep_gen4.json'))
for r in d["rows"]: print(r["N"], r["batch"], round(r["fwd_ms_per_transform"]*1e3,2), "us fwd", round(r["inv_ms_per_transform"]*1e3,2), "us inv")
print("cfg0 latency ms", d["config0_fwd_plus_inv_latency_ms"])
PY
python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
