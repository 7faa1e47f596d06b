timeout 900 compute-sanitizer --tool memcheck python profiles/sanitizer_smoke.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r2_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python profiles/sanitizer_smoke.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/r2_sanitizer_racecheck.log
