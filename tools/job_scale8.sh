mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 2000 --warmup 200 > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err
echo rc=$?; tail -1 gpurun_out/bench_8gpu.log | cut -c1-400; tail -5 gpurun_out/bench_8gpu.err
