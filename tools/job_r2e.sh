set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy.py -q -x > gpurun_out/pytest_policy.log 2>&1; tail -5 gpurun_out/pytest_policy.log
timeout 300 python tools/policy_time.py > gpurun_out/policy_time.log 2>&1; tail -6 gpurun_out/policy_time.log
ncu --set full --clock-control none --import-source on -k regex:policy_forward -s 2 -c 1 -f -o gpurun_out/prof_policy python tools/policy_step.py 131072 1 > gpurun_out/ncu_policy.log 2>&1; tail -2 gpurun_out/ncu_policy.log
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/prof_step_r02 python tools/aged_step.py cfg4 > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -f -o gpurun_out/prof_f64_r02 python /dev/stdin > gpurun_out/ncu_f64.log 2>&1 <<'PY'
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
tail -2 gpurun_out/ncu_f64.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
ls -la gpurun_out | tail -10
