# usage: gpurun --gpus N -- 'bash tools/job_scale2.sh N'   (driver-format lines of both arms at N GPUs)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/driver_ref_${N}gpu.log 2>&1; echo rc=$?; tail -1 gpurun_out/driver_ref_${N}gpu.log | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/driver_b200_${N}gpu.log 2> gpurun_out/driver_b200_${N}gpu.err
echo rc=$?; tail -1 gpurun_out/driver_b200_${N}gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['n_gpus'], json.dumps(d['strong'])[:400], d['e2e']['ms_per_step'], d['e2e']['value'], d['episode_stats'])"; tail -3 gpurun_out/driver_b200_${N}gpu.err
