# round-2 final set: driver-format bench lines, GPU suite, smoke, launch list, ncu captures (steady state step kernel, policy kernel)
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_ref.log 2>&1 ) 2>&1 | tail -3
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_b200.log 2>gpurun_out/driver_b200.err ) 2>&1 | tail -3
tail -1 gpurun_out/driver_b200.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_layout'], d['roofline'].get('frac_traffic'), d['episode_stats']['episodes'], d['done_episodes_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']); print(d['configs']['cfg2@4096']['us_per_tick_graph'], d['configs']['cfg3@262144']['us_per_tick_eager'], d['rollout_collection']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['reference_python']['value'])"
( time python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err ) 2>&1 | tail -3; tail -1 gpurun_out/bench.log | cut -c1-300; tail -3 gpurun_out/bench.err
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 --rollout-steps 0 --post-steps 20 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 --rollout-steps 0 --post-steps 20 > gpurun_out/ncu_launch.log 2>&1
python tools/aged_step.py cfg4 > gpurun_out/aged.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/prof_aged python tools/aged_step.py cfg4 > gpurun_out/ncu_aged.log 2>&1
python tools/policy_step.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:policy_forward -s 3 -c 1 -f -o gpurun_out/prof_policy python tools/policy_step.py > gpurun_out/ncu_policy.log 2>&1
tail -2 gpurun_out/ncu_aged.log gpurun_out/ncu_policy.log
python tools/aged_time.py cfg4,cfg2,cfg3 fp32 > gpurun_out/aged_time.log 2>&1; python tools/aged_time.py cfg4,cfg2,cfg3 fp64 >> gpurun_out/aged_time.log 2>&1; cat gpurun_out/aged_time.log
python tools/baseline_sizes.py > gpurun_out/baseline_sizes.log 2>&1; tail -6 gpurun_out/baseline_sizes.log
