mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/e2e_multi.py > gpurun_out/e2e_multi.log 2> gpurun_out/e2e_multi.err
echo rc=$?; grep -v "^$" gpurun_out/e2e_multi.log | tail -60; tail -5 gpurun_out/e2e_multi.err
