timeout 900 python profiles/measure_lanes.py gpurun_out/r2_lanes_8gpu.json > gpurun_out/r2_lanes_8gpu.log 2>&1; tail -3 gpurun_out/r2_lanes_8gpu.log
timeout 1200 python profiles/measure_whole_step.py gpurun_out/r2_whole_step_8gpu.json --quick > gpurun_out/r2_whole_step_8gpu.log 2>&1; tail -5 gpurun_out/r2_whole_step_8gpu.log
