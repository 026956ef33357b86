set -x
nvidia-smi topo -m 2>&1 | head -12
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; tail -c 300 gpurun_out/bench_n$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print("N=$n", {k:d.get(k) for k in ("value","ms_per_step","verified","ntt_64k_per_s")}, "e2e", d["e2e"]["value"], "c5", d["config5"])
PY
done
