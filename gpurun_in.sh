set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "three_reductions or mul_barrett or relin or golden or full_size or large_prime" 2>&1 | tail -8
for m in sparse ntt; do CUHE_B200_REDUCE=$m python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-200; done
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
