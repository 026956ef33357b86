set -x
for ch in 3 2; do
CUHE_B200_SHARD_CHUNKS=$ch python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
CUHE_B200_SHARD_CHUNKS=$ch python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-c5 > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2c.json').read().strip().splitlines()[-1])
print("chunks $ch", {k:d.get(k) for k in ("value","ms_per_step","verified")}, "e2e", d["e2e"]["value"])
PY
done
