python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_base.py -x -q -m gpu 2>&1 | tail -2
python tools/ntt_bench.py | cut -c1-330
python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
