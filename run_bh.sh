timeout 600 python -m pytest tests/test_bh_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --workload bh --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bh_group_b.json 2> gpurun_out/r2_bh_group_b.err; tail -c 300 gpurun_out/r2_bh_group_b.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bh_group_b.json').read().strip().splitlines()[-1]); print(d['value'], d['config']['phases_ms_last_step'], d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bh_walk_group -s 1 -c 1 -f -o gpurun_out/r2_bh_group_n1m python bench.py --workload bh --bodies 1048576 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_a.log 2>&1; tail -c 200 gpurun_out/ncu_a.log
