for v in "" su2 su4; do
NB200_LIB_VARIANT=$v timeout 300 python bench.py --workload direct --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/tmp.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
python -c "
import json; d=json.loads(open('gpurun_out/tmp.json').read().strip().splitlines()[-1]); print('variant', '$v', '%.4g'%d['value'], d['config']['phases_ms_last_step']['force'])"
done
