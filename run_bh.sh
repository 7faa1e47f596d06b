timeout 600 python -m pytest tests/test_bh_gpu.py -x -q -m gpu 2>&1 | tail -3
for n in 4194304; do
timeout 300 python bench.py --workload bh --bodies $n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bh_group_d_$n.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bh_group_d_$n.json').read().strip().splitlines()[-1]); print(d['value'], d['config']['phases_ms_last_step'], d['roofline']['frac'], d['roofline']['node_visits'], d['roofline']['interactions'], d['roofline']['walk_profile_rank0'])"
done
