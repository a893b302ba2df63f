# usage: gpurun --gpus N -- 'bash tools/job_scale.sh N'
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2000 --warmup 50 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
echo rc=$?; tail -1 gpurun_out/bench_${N}gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d['strong']), d['e2e']['ms_per_step'], d['e2e']['obs17']['ms_per_step'], json.dumps(d['rollout_collection']['cfg5']) if d.get('rollout_collection') else None)"; tail -5 gpurun_out/bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_multi.py > gpurun_out/pcie_${N}gpu.log 2>&1; cat gpurun_out/pcie_${N}gpu.log | tail -6
