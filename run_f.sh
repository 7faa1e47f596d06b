timeout 1200 python -m pytest tests/test_bh_gpu.py tests/test_direct_gpu.py tests/test_solvers_gpu.py tests/test_stepgraph_gpu.py -x -q -m gpu 2>&1 | tail -4
for w in direct bh; do
timeout 300 python bench.py --workload $w --precision f32 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_${w}_f32.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_${w}_f32.json').read().strip().splitlines()[-1]); print('$w f32', d['value'], d['config']['phases_ms_last_step'])"
done
timeout 300 python profiles/measure_direct_sizes.py 65536 262144 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for r in d['rows']: print(r['precision'], r['bodies'], '%.3g'%r['ordered_pairs'], '%.3g'%r['symmetric_tiles'])"
