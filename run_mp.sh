N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/mp_check.py 2>&1 | grep -v "^W\|warn" | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; tail -c 400 gpurun_out/r2_bench_${N}gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1]); print('direct', d['value'], d['e2e']['value'], d['config']['phases_ms_last_step']); b=d['bh']; print('bh', b['value'], b['e2e']['value'], b['config']['phases_ms_last_step'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ref', d['value'], d['cpu_baseline']['cores'], d['config']['wall_s'])"
