set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -1 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
python gpurun_quick.py > gpurun_out/quick.log 2>&1; cat gpurun_out/quick.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 2 -o gpurun_out/prof_step python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
