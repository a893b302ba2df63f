set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -1 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 2 -o gpurun_out/prof_step python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
