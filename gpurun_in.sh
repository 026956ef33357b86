set -x
python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; tail -c 200 gpurun_out/bench_n2b.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2b.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","verified","ntt_64k_per_s")}, "e2e", d["e2e"]["value"], "c5", d["config5"]["value"], d["config5"]["verified"])
PY
