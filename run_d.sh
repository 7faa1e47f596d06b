timeout 900 python -m pytest tests/test_direct_gpu.py tests/test_solvers_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python profiles/measure_direct_sizes.py 8192 16384 32768 65536 262144 > gpurun_out/r2_direct_sizes.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_direct_sizes.json'))
for r in d['rows']:
    print(r['precision'], r['bodies'], {k:('%.3g'%v if isinstance(v,float) else v) for k,v in r.items() if k not in ('precision','bodies')})
PY
timeout 300 python bench.py --workload direct --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err; tail -c 300 gpurun_out/tmp.err
python -c "
import json; d=json.loads(open('gpurun_out/tmp.json').read().strip().splitlines()[-1]); print('N=1M', '%.4g'%d['value'], '%.4g'%d['e2e']['value'], d['config']['phases_ms_last_step'], d['roofline']['kernel'])"
