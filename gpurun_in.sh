set -x
compute-sanitizer --tool memcheck python tools/one_ntt.py 2>&1 | grep -E "=========|one_ntt" | tee gpurun_out/r02_gen4_sanitizer_memcheck_ntt.log
compute-sanitizer --tool racecheck python tools/one_ntt.py 2>&1 | grep -E "=========|one_ntt" | tee gpurun_out/r02_gen4_sanitizer_racecheck_ntt.log
compute-sanitizer --tool memcheck python __graft_entry__.py --smoke 2>&1 | grep -E "=========|smoke" | tee gpurun_out/r02_gen4_sanitizer_memcheck_smoke.log
compute-sanitizer --tool racecheck python __graft_entry__.py --smoke 2>&1 | grep -E "=========|smoke" | tee gpurun_out/r02_gen4_sanitizer_racecheck_smoke.log
compute-sanitizer --tool memcheck python tools/one_ntt.py 16384 2>&1 | grep -E "=========|one_ntt" | tee gpurun_out/r02_gen4_sanitizer_memcheck_ntt16k.log
compute-sanitizer --tool racecheck python tools/one_ntt.py 32768 2>&1 | grep -E "=========|one_ntt" | tee gpurun_out/r02_gen4_sanitizer_racecheck_ntt32k.log
