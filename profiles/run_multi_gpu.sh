# N ranks on one node: tests/mp_check.py (parity of the sharded paths) and the bench line; usage: bash profiles/run_multi_gpu.sh N [nocheck]
N=${1:-2}
if [ "$2" != "nocheck" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/mp_check.py > gpurun_out/r2_mp_check_${N}gpu.log 2>&1; grep "mp_check\|rank . done\|Error\|error" gpurun_out/r2_mp_check_${N}gpu.log | tail -8
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; tail -c 300 gpurun_out/r2_bench_${N}gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1]); print('direct', d['value'], d['e2e']['value'], d['config']['phases_ms_last_step']); b=d['bh']; print('bh', b['value'], b['e2e']['value'], b['config']['phases_ms_last_step'])"
