# final check of the round's last code state: GPU suite, smoke, driver-format bench of both arms
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_ref.log 2>&1; tail -1 gpurun_out/driver_ref.log | cut -c1-160
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_b200.log 2>gpurun_out/driver_b200.err
tail -1 gpurun_out/driver_b200.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_layout'], d['episode_stats']['episodes'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks']['sm_mhz'], d['rollout_collection']['value'])"
python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_layout'], d['e2e']['value'], d['clocks']['sm_mhz'])"
