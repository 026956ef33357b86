set -x
ncu --set full --import-source on --clock-control none -k regex:"cyclo_reduce|icrt_kernel_v3|crt_kernel_v3" -s 4 -c 4 -o gpurun_out/r02_rns3_full python bench.py --steps 2 --warmup 1 --no-cpu --no-c5 > /dev/null 2>&1
ls -la gpurun_out/r02_rns3_full.ncu-rep
