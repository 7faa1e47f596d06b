timeout 600 python -m pytest tests/test_bh_gpu.py -x -q -m gpu 2>&1 | tail -3
for p in f64 f32; do
timeout 300 python bench.py --workload bh --precision $p --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bh_group_f_$p.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bh_group_f_$p.json').read().strip().splitlines()[-1]); print('$p', d['value'], d['config']['phases_ms_last_step'])"
done
