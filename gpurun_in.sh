set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
cat gpurun_out/prince_kat_timings.jsonl
