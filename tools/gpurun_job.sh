set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err ) 2>&1 | tail -3; tail -1 gpurun_out/bench.log | cut -c1-200; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 --rollout-steps 0 --post-steps 20 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 --rollout-steps 0 --post-steps 20 > gpurun_out/ncu_launch.log 2>&1
# the step kernel early in the episode (all envs in phase) ...
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 1 -f -o gpurun_out/prof_step python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 --rollout-steps 0 --post-steps 0 > gpurun_out/ncu_full.log 2>&1
# ... and at the steady state the 2000-tick bench spends most of its time in (episodes desynchronised, ~900 auto-resets per tick)
python tools/aged_step.py cfg4 > gpurun_out/aged.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/prof_aged python tools/aged_step.py cfg4 > gpurun_out/ncu_aged.log 2>&1
tail -2 gpurun_out/ncu_full.log gpurun_out/ncu_aged.log
python tools/aged_time.py cfg4,cfg2,cfg3 1 > gpurun_out/aged_time.log 2>&1; cat gpurun_out/aged_time.log
ls -la gpurun_out
