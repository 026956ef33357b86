set -x
timeout 300 python tools/ntt_bench.py | cut -c1-330
CUHE_B200_NTT_FUSED=8 CUHE_B200_NTT_FUSED_CHUNK=4 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ntt or mul or relin" 2>&1 | tail -3
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:ntt4_ -s 4 -c 3 --csv --log-file gpurun_out/ncu_fused_dram.csv python tools/ntt_bench.py --one > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_fused_dram.csv')) if len(r)>5]
h=next(r for r in rows if "Kernel Name" in r)
for r in rows[rows.index(h)+1:]: print(r[h.index("Kernel Name")][:40], r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
