set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_ref.log 2>&1 ) 2>&1 | tail -3
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_b200.log 2>gpurun_out/driver_b200.err ) 2>&1 | tail -3
tail -1 gpurun_out/driver_b200.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_layout'], d['episode_stats'], d['done_episodes_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']); print(d['configs']['cfg2@4096']['us_per_tick_graph'], d['configs']['cfg3@262144']['us_per_tick_eager'], d['rollout_collection']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['reference_python']['value'])"
tail -1 gpurun_out/driver_ref.log | cut -c1-600
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
