set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 500 --warmup 50 --no-cpu-baseline --fused 0 > gpurun_out/bench_e2e.log 2>&1; tail -1 gpurun_out/bench_e2e.log
python gpurun_e2e.py > gpurun_out/e2e_sweep.log 2>&1; cat gpurun_out/e2e_sweep.log
