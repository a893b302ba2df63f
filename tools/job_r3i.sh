# final profile set of the round (r02-m)
set -x
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_ref.log 2>&1; tail -1 gpurun_out/driver_ref.log | cut -c1-120
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_b200.log 2>gpurun_out/driver_b200.err
tail -1 gpurun_out/driver_b200.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_layout'], d['roofline']['frac_traffic'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks']['sm_mhz'], d['rollout_collection']['value'], d['configs']['cfg2@4096']['us_per_tick_graph'], d['configs']['cfg3@262144']['us_per_tick_eager'])"
python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_layout'], d['e2e']['value'], d['clocks']['sm_mhz'])"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 --rollout-steps 0 --post-steps 20 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused 0 --e2e-steps 2 --rollout-steps 0 --post-steps 20 > gpurun_out/ncu_launch.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_kernel -s 1 -c 12 --csv --log-file gpurun_out/traffic_cfg4_final.csv python tools/aged_step.py cfg4 > gpurun_out/ncu_traffic.log 2>&1
python tools/aged_time.py cfg4,cfg2,cfg3 fp32 > gpurun_out/aged_time.log 2>&1; python tools/aged_time.py cfg4,cfg2,cfg3 fp64 >> gpurun_out/aged_time.log 2>&1; cat gpurun_out/aged_time.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
