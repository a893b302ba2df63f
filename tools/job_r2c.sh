set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy.py tests/test_cuda_parity.py tests/test_vec_env.py tests/test_launcher.py tests/test_rollout.py -m gpu -q -x -k "not mixed_feature or True" > gpurun_out/pytest_sel.log 2>&1; tail -8 gpurun_out/pytest_sel.log
for o in "" "pdl=1"; do HLYNR_OPTS=$o python tools/aged_time.py cfg4,cfg2,cfg3 fp32 >> gpurun_out/aged_time_pdl.log 2>&1; done
for o in "" "pdl=1"; do HLYNR_OPTS=$o python tools/aged_time.py cfg2 fp32 4096 >> gpurun_out/aged_time_pdl.log 2>&1; HLYNR_OPTS=$o python tools/aged_time.py cfg3 fp32 262144 >> gpurun_out/aged_time_pdl.log 2>&1; done
python tools/aged_time.py cfg4,cfg2,cfg3 fp64 >> gpurun_out/aged_time_pdl.log 2>&1
cat gpurun_out/aged_time_pdl.log
timeout 600 python tools/e2e_sweep.py > gpurun_out/e2e_sweep.log 2>&1; cat gpurun_out/e2e_sweep.log
( time python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err ) 2>&1 | tail -3; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['rollout_collection'], indent=1)); print(d['value'], d['e2e']['ms_per_step'], d['e2e']['obs17']['ms_per_step'])"; tail -5 gpurun_out/bench.err
ls -la gpurun_out | tail -8
