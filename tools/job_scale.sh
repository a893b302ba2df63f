# usage: bash tools/job_scale.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 2000 --warmup 200 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
echo rc=$?; tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-300
