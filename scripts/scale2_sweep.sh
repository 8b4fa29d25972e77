# developer tool: 2-GPU sweep of the gradient bucket size of bench.py (run under gpurun --gpus 2)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))"; }
python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))"
MICO_BENCH_BUCKET_BLOCKS=5 run 29511 "N=2 bucket5"
MICO_BENCH_BUCKET_BLOCKS=10 run 29512 "N=2 bucket10"
MICO_BENCH_BUCKET_BLOCKS=20 run 29513 "N=2 bucket20"
MICO_BENCH_BUCKET_BLOCKS=40 run 29514 "N=2 bucket40(end)"
