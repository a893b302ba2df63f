mkdir -p gpurun_out
python -m pytest tests/test_vec_env.py tests/test_post.py tests/test_rollout.py -m gpu -x -q 2>&1 | tail -4
HLYNR_HOST_TRACE=1 python tools/e2e_trace.py 2>&1 | tail -9
python tools/e2e_pyprof.py 2>&1 | head -3
python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/e2e_breakdown.log
