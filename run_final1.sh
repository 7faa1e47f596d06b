timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1]); print('direct', d['value'], d['e2e']['value'], d['config']['phases_ms_last_step'], d['roofline']['frac_issued']); b=d['bh']; print('bh', b['value'], b['e2e']['value'], b['config']['phases_ms_last_step'], b['roofline']['frac'], b['roofline']['lane_use'], b['roofline']['frac_issued'], b['cpu_baseline']); print(d['cpu_baseline']['value'], d['clocks'])"
timeout 300 python bench.py --workload direct --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --opt direct_sym_tile=4096 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tile4096', d['value'])"
