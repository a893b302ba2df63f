set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-300; tail -3 gpurun_out/bench.err
python gpurun_quick.py > gpurun_out/quick.log 2>&1; cat gpurun_out/quick.log
