set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for t in 1 4 8 16; do ./tools/_mulzzx_bench $t 16; done
CUHE_B200_MULZZX=literal ./tools/_mulzzx_bench 1 8
CUHE_B200_MULZZX=literal ./tools/_mulzzx_bench 8 8
