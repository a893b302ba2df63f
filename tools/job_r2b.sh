set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests/test_policy.py -x -q > gpurun_out/pytest_policy.log 2>&1; tail -30 gpurun_out/pytest_policy.log
timeout 300 python tools/policy_time.py > gpurun_out/policy_time.log 2>&1; cat gpurun_out/policy_time.log | tail -8
timeout 1700 python -m pytest tests -m gpu -q --deselect tests/test_policy.py > gpurun_out/pytest_gpu.log 2>&1; tail -30 gpurun_out/pytest_gpu.log
python tools/aged_time.py cfg4,cfg2,cfg3 fp64 > gpurun_out/aged_time_f64.log 2>&1; cat gpurun_out/aged_time_f64.log
cat > /tmp/f64step.py <<'PY'
import sys, os; sys.path.insert(0, os.getcwd())
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
sim = HlynrSim(config.baseline_config('cfg4'), n_envs=n, warn_dead=False, precision='fp64')
sim.reset(); sim.rollout(1500, None, want_obs=False)
pool = [(torch.rand(n, 6, device='cuda') * 2 - 1) for _ in range(4)]
for k in range(8): sim.step(pool[k % 4], want_terminal_obs=False)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -f -o gpurun_out/prof_f64 python /tmp/f64step.py > gpurun_out/ncu_f64.log 2>&1; tail -2 gpurun_out/ncu_f64.log
ls -la gpurun_out | tail -12
