set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/relin_bench.py
CUHE_B200_RELIN_RING=1 python tools/relin_bench.py
CUHE_B200_RELIN_RING=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "relin" 2>&1 | tail -3
for c in 16 32 64 128; do CUHE_B200_NTT_CHUNK_ROWS=$c python tools/ntt_bench.py; done
CUHE_B200_NTT_CHUNK_ROWS=64 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ntt or mul or relin" 2>&1 | tail -3
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ntt96 -s 6 -c 4 --csv --log-file gpurun_out/ncu_dram_nochunk.csv python tools/ntt_bench.py > /dev/null 2>&1
CUHE_B200_NTT_CHUNK_ROWS=64 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ntt96 -s 48 -c 16 --csv --log-file gpurun_out/ncu_dram_chunk64.csv python tools/ntt_bench.py > /dev/null 2>&1
compute-sanitizer --tool memcheck python tools/one_ntt.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_memcheck_ntt.log
compute-sanitizer --tool racecheck python tools/one_ntt.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_racecheck_ntt.log
compute-sanitizer --tool memcheck python __graft_entry__.py --smoke 2>&1 | tail -6 | tee gpurun_out/sanitizer_memcheck_smoke.log
compute-sanitizer --tool racecheck python __graft_entry__.py --smoke 2>&1 | tail -6 | tee gpurun_out/sanitizer_racecheck_smoke.log
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -2 | cut -c1-1500
