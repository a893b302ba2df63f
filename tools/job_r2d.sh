set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_policy.py -q -x > gpurun_out/pytest_policy.log 2>&1; tail -12 gpurun_out/pytest_policy.log
timeout 300 python tools/policy_time.py > gpurun_out/policy_time.log 2>&1; tail -6 gpurun_out/policy_time.log
timeout 1700 python -m pytest tests -m gpu -q --deselect tests/test_policy.py > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
( time python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err ) 2>&1 | tail -3; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['rollout_collection'], indent=1)[:1500]); print(json.dumps(d['cpu_baseline'], indent=1)[:2500]); print(d['value'], d['e2e']['ms_per_step'], d['e2e']['obs17']['ms_per_step'], d['configs'])"; tail -5 gpurun_out/bench.err
ls -la gpurun_out | tail -8
